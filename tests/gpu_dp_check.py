"""2-rank data-parallel parity check (run under torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/gpu_dp_check.py

Each rank trains on one clip through FusedDataParallel (one fused grad+EMA-statistics all-reduce); rank 0 also runs
the SAME two clips in one process (batched == DDP semantics) and compares gradients and codebooks.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from faceoff_b200.parallel import FusedDataParallel  # noqa: E402
from faceoff_b200.vqvae import VQVAE  # noqa: E402
from oracle import faceoff_oracle as O  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    p = O.init_vqvae_params(seed=0)
    T, R = 4, 64
    img, gt = O.synthetic_clip(world, T, R, R, seed=77)

    def run(model, x, y, clips, net=None):
        net = net or model
        model.zero_grad(set_to_none=True)
        out, latent = net.forward_with_ids(x, clips)[:2]
        loss = torch.nn.functional.mse_loss(out[:, :3], y) + latent.mean()
        loss.backward()
        return loss

    model = VQVAE(in_channel=6)
    model.load_state_dict(p)
    model = model.cuda().train()
    ddp = FusedDataParallel(model)
    sl = slice(rank * T, (rank + 1) * T)
    loss = run(model, img[sl].cuda(), gt[sl].cuda(), 1, ddp)
    torch.cuda.synchronize()
    grads = {k: v.grad.clone() for k, v in model.named_parameters()}
    bufs = {k: v.clone() for k, v in model.named_buffers()}
    # codebooks must be bit-identical across ranks (no buffer broadcast needed)
    for k, v in bufs.items():
        lst = [torch.empty_like(v) for _ in range(world)]
        dist.all_gather(lst, v)
        assert all(torch.equal(lst[0], t) for t in lst), f"buffer {k} differs across ranks"
    ok = True
    if rank == 0:
        ref = VQVAE(in_channel=6)
        ref.load_state_dict(p)
        ref = ref.cuda().train()
        from faceoff_b200.vqvae import _LocalStatSink
        for q in (ref.quantize_t, ref.quantize_b):
            q.stat_sink = _LocalStatSink()   # rank 0 only: must not enter a collective
        run(ref, img.cuda(), gt.cuda(), world)
        torch.cuda.synchronize()
        worst = 0.0
        for k, v in ref.named_parameters():
            e = ((grads[k] - v.grad).abs().max() / (v.grad.abs().max() + 1e-12)).item()
            worst = max(worst, e)
        print(f"DP vs single-process batched: worst max-normalised grad err {worst:.3e}")
        ok = ok and worst < 2e-2
        for k, v in ref.named_buffers():
            e = ((bufs[k] - v).abs().max() / (v.abs().max() + 1e-12)).item()
            print(f"  buffer {k}: rel err {e:.3e}")
            ok = ok and e < 1e-4
        print("DP CHECK", "PASS" if ok else "FAIL")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
