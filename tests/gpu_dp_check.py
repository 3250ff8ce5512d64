"""2-rank data-parallel parity check (run under torchrun):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/gpu_dp_check.py

With >= 2 GPUs the ranks use one GPU each over NCCL; on a 1-GPU box both ranks share cuda:0 and the collective runs over
gloo (same FusedDataParallel code path: bucket, chunked async all-reduce, deferred EMA), so the reducer's semantics
(reference distributed/distributed.py:64-72 + DDP, train_faceoff_perceptual.py:164-169) are checked wherever the GPU
tests run.

1. one clip per rank through FusedDataParallel with the LPIPS loss on; rank 0 also runs the SAME clips in one process
   (batched == DDP semantics) and compares every gradient and codebook; buffers must be bit-identical
   across ranks.
2. micro-batches: clip A inside ``no_sync()``, clip B outside == the batch of all 2*world clips (gradients x2 because
   DDP averages over ranks only), ONE EMA update from the summed statistics.
3. a train-mode forward under ``torch.no_grad()`` with world > 1 (used to raise KeyError) behaves like the reference:
   statistics all-reduced and EMA applied inside forward.

``--cpu`` (used by ``pytest -m "not gpu"``, tests/test_cpu.py): the same three checks with the ranks on the CPU over gloo and
the kernels replaced by the plain-torch stand-ins of tests/fake_ops.py -- the PRODUCT's host logic (FusedDataParallel: flat
bucket, ready frontier, chunked all-reduce, deferred statistics / EMA, no_sync; the fused VQVAE + LPIPS tapes) without a GPU.
"""
import os
import sys
import warnings

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from faceoff_b200.lpips import VQLPIPS  # noqa: E402
from faceoff_b200.parallel import FusedDataParallel  # noqa: E402
from faceoff_b200.vqvae import VQVAE, _LocalStatSink  # noqa: E402
from oracle import faceoff_oracle as O  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    cpu = "--cpu" in sys.argv
    if cpu:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import fake_ops

        fake_ops.install_for_process()
        torch.cuda.synchronize = lambda *a, **k: None     # nothing asynchronous on this path
        multi_gpu = False
        dev = torch.device("cpu")
        dist.init_process_group("gloo")
    else:
        multi_gpu = torch.cuda.device_count() >= world
        dev = torch.device("cuda", rank if multi_gpu else 0)
        torch.cuda.set_device(dev)
        if multi_gpu:
            dist.init_process_group("nccl", device_id=dev)
        else:
            dist.init_process_group("gloo")
    p = O.init_vqvae_params(seed=0)
    T, R = (2, 32) if cpu else (4, 64)
    img, gt = O.synthetic_clip(2 * world, T, R, R, seed=77)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        vql = VQLPIPS()
    vql.load_state_dict({"perceptual_loss." + k: v for k, v in O.init_lpips_params(seed=1).items()})
    vql = vql.to(dev)

    def new_model(local=False):
        m = VQVAE(in_channel=6)
        m.load_state_dict(p)
        m = m.to(dev).train()
        if local:
            for q in (m.quantize_t, m.quantize_b):
                q.stat_sink = _LocalStatSink()   # single-process replica inside a multi-rank job: no collective
        return m

    def run(model, x, y, clips, net=None, zero=True):
        net = net or model
        if zero:
            model.zero_grad(set_to_none=True)
        out, latent = net.forward_with_ids(x, clips)[:2]
        loss = torch.nn.functional.mse_loss(out[:, :3], y) + latent.mean() + vql(y, out[:, :3])
        loss.backward()
        return loss

    def same_on_all_ranks(model, what):
        for k, v in model.named_buffers():
            # NCCL gathers device tensors (one GPU per rank), gloo host tensors (both ranks on cuda:0)
            mine = v.detach().clone() if multi_gpu else v.detach().cpu()
            lst = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(lst, mine)
            assert all(torch.equal(lst[0], t) for t in lst), f"{what}: buffer {k} differs across ranks"

    def compare(tag, model, ref, grad_scale, tol_g, tol_b):
        worst = 0.0
        for (k, v), (_, r) in zip(model.named_parameters(), ref.named_parameters()):
            e = ((v.grad - grad_scale * r.grad).abs().max() / ((grad_scale * r.grad).abs().max() + 1e-12)).item()
            worst = max(worst, e)
        print(f"{tag}: worst max-normalised grad err {worst:.3e}")
        ok = worst < tol_g
        for (k, v), (_, r) in zip(model.named_buffers(), ref.named_buffers()):
            e = ((v - r).abs().max() / (r.abs().max() + 1e-12)).item()
            print(f"  buffer {k}: rel err {e:.3e}")
            ok = ok and e < tol_b
        return ok

    ok = True
    # ---- 1. one clip per rank, LPIPS on ----
    model = new_model()
    ddp = FusedDataParallel(model)
    sl = slice(rank * T, (rank + 1) * T)
    run(model, img[sl].to(dev), gt[sl].to(dev), 1, ddp)
    torch.cuda.synchronize()
    same_on_all_ranks(model, "step 1")
    if rank == 0:
        ref = new_model(local=True)
        run(ref, img[:world * T].to(dev), gt[:world * T].to(dev), world)
        torch.cuda.synchronize()
        ok = compare("DP vs single-process batched (with LPIPS)", model, ref, 1.0, 1e-4, 1e-5) and ok
    # ---- 2. micro-batches under no_sync ----
    model = new_model()
    ddp = FusedDataParallel(model)
    model.zero_grad(set_to_none=True)
    a = slice(rank * T, (rank + 1) * T)
    b = slice((world + rank) * T, (world + rank + 1) * T)
    with ddp.no_sync():
        run(model, img[a].to(dev), gt[a].to(dev), 1, ddp, zero=False)
    e_mid = model.quantize_b.embed.clone()
    run(model, img[b].to(dev), gt[b].to(dev), 1, ddp, zero=False)
    torch.cuda.synchronize()
    same_on_all_ranks(model, "micro-batches")
    if rank == 0:
        assert torch.equal(e_mid, torch.as_tensor(p["quantize_b.embed"]).to(dev)), "no_sync must not touch the codebooks"
        ref = new_model(local=True)
        run(ref, img.to(dev), gt.to(dev), 2 * world)
        torch.cuda.synchronize()
        # bf16 activations differ slightly between the two micro-batch passes and the batched pass only through fp32
        # summation order, except that micro-batch B sees the SAME (pre-update) codebook as A -- exactly the batched semantics
        ok = compare("no_sync micro-batches vs one batch", model, ref, 2.0, 1e-4, 1e-5) and ok
    # ---- 3. train-mode forward under no_grad with world > 1 ----
    model = new_model()
    ddp = FusedDataParallel(model)
    with torch.no_grad():
        ddp(img[sl].to(dev))
    torch.cuda.synchronize()
    same_on_all_ranks(model, "no_grad forward")
    changed = not torch.equal(model.quantize_t.embed, torch.as_tensor(p["quantize_t.embed"]).to(dev))
    if rank == 0:
        print("no_grad train-mode forward: EMA applied in forward:", changed)
        ok = ok and changed
        print("DP CHECK", "PASS" if ok else "FAIL")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
