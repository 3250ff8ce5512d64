#!/usr/bin/env python
"""Benchmark of the FaceOff VQVAE-conv3d training step (fwd+bwd) on B200 -- BASELINE.json metric:
"train clips/sec (b32, 256^2, fwd+bwd)".

    python bench.py [--gpus N] [--steps K] [--warmup W]          # our arm (one process per GPU; torchrun for N>1)
    python bench.py --impl reference [--gpus N] ...               # the reference's CPU implementation of the path
    python bench.py --impl eager ...                              # the same functional step in library-eager PyTorch on the GPU

One JSON line is printed by rank 0.  What it carries:

* ``value`` / ``e2e`` / ``roofline``: BASELINE configs[1] -- VQVAE(in_channel=6) train step WITHOUT perceptual loss, global
  batch 32 clips x T=30 frames of 256x256 synthetic data, seeded random weights (the config the metric is quoted on).
* ``lpips_step``: BASELINE configs[2] -- the same step WITH the LPIPS loss (train_faceoff_perceptual.py:32-47,98), the
  step north_star names as the target; timed the same way (device-resident and end-to-end), at every N.
* ``dp_check`` (N > 1): computed during warm-up on separate replicas: after one synchronised data-parallel step (one clip
  per rank, LPIPS on) every rank's codebook buffers and averaged gradient bucket must be BIT-identical, and rank 0
  re-runs the same N clips in one process and compares every gradient (reference DP semantics,
  distributed/distributed.py:64-72 + DDP, train_faceoff_perceptual.py:164-169).
* ``cpu_baseline`` (N = 1): the oracle port on the host cores; ``extra.gpu_eager_baseline`` (N = 1): the same functional
  step as plain PyTorch ops on the GPU (cuDNN/cuBLAS from the installed wheel) -- the number the kernels have to beat.

For N>1 the 32 clips are sharded over the ranks (strong scaling) and the fused EMA+gradient all-reduce runs every step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_FRAMES = 30
RES = 256
GLOBAL_CLIPS = 32
METRIC = "train clips/sec (b32, 256^2, fwd+bwd)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager"])
    ap.add_argument("--clips", type=int, default=GLOBAL_CLIPS, help="global batch in clips")
    ap.add_argument("--frames", type=int, default=T_FRAMES)
    ap.add_argument("--res", type=int, default=RES)
    ap.add_argument("--lpips", type=int, default=0,
                    help="1: make the LPIPS step (configs[2]) the headline value instead of reporting it under lpips_step")
    ap.add_argument("--no-lpips-step", action="store_true", help="skip the configs[2] measurement")
    ap.add_argument("--no-dp-check", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-disc-step", action="store_true")
    ap.add_argument("--cpu-clip-frames", type=int, default=T_FRAMES)
    ap.add_argument("--eager-clips", type=int, default=4)
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


def ncu_evidence():
    """Static pointer to the committed ncu captures (profiles/), not a live measurement."""
    for name in ("r2c_roofline_evidence.json", "r2b_roofline_evidence.json", "r2_roofline_evidence.json", "r1_roofline_evidence.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            return json.load(open(path))
    return None


def workload_config(args, n, lpips):
    return {"workload": "VQVAE-conv3d train step (fwd+bwd), BASELINE configs[1]" + (" + LPIPS (configs[2])" if lpips else ""),
            "global_batch_clips": args.clips, "frames_per_clip": args.frames, "resolution": args.res,
            "in_channel": 6, "embed_dim": 64, "n_embed": 512, "parallelism": f"dp{n}",
            "l2": "inputs (>= 1.5 GB/step) and activations (~30 GB/step) are far larger than the 126 MB L2"}


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_train_step_time(frames: int, res: int, lpips: bool, steps: int, warmup: int):
    """Times the oracle port of the reference path (zero_grad -> fwd -> MSE + latent [+ LPIPS] -> bwd) for ONE clip
    on all host cores.  Returns (seconds per clip, cores)."""
    import torch

    from oracle import faceoff_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    p = O.init_vqvae_params(seed=0)
    lp = O.init_lpips_params(seed=1) if lpips else None
    img, gt = O.synthetic_clip(1, frames, res, res, seed=1234)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.train_step(p, img, gt, n_clips=1, lp=lp)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times) / len(times), cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sec, cores = cpu_train_step_time(args.frames, args.res, bool(args.lpips), args.steps, args.warmup)
    value = 1.0 / sec
    sample = f"1 clip of T={args.frames} frames {args.res}x{args.res} per step (of the {args.clips}-clip batch)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "clips/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus, bool(args.lpips)),
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ library-eager GPU arm
def eager_gpu_baseline(args, dev, clips: int, steps: int = 2, warmup: int = 1):
    """The functional restatement of the reference step (oracle/, plain torch ops) moved to the GPU: what the installed
    wheel's cuDNN / cuBLAS give without any of this repo's kernels (SURVEY 8(d) secondary baseline).  Bounded sample of
    ``clips`` clips per step; fp32 with TF32 off (the reference's numerics), TF32 on (its default on Ampere+) and bf16
    autocast.  Returns {variant: {config: clips/s}}."""
    import torch

    from oracle import faceoff_oracle as O

    p = {k: v.to(dev) for k, v in O.init_vqvae_params(seed=0).items()}
    lp = {k: v.to(dev) for k, v in O.init_lpips_params(seed=1).items()}
    img, gt = O.synthetic_clip(clips, args.frames, args.res, args.res, seed=1234)
    img, gt = img.to(dev), gt.to(dev)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    out = {}
    try:
        for variant in ("fp32_tf32_off", "fp32_tf32_on", "bf16_autocast"):
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = variant == "fp32_tf32_on"
            res = {}
            for name, lpp in (("configs1_no_lpips", None), ("configs2_lpips", lp)):
                def one():
                    if variant == "bf16_autocast":
                        with torch.autocast("cuda", dtype=torch.bfloat16):
                            O.train_step(p, img, gt, n_clips=clips, lp=lpp)
                    else:
                        O.train_step(p, img, gt, n_clips=clips, lp=lpp)

                try:
                    for _ in range(warmup):
                        one()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(steps):
                        one()
                    e1.record()
                    torch.cuda.synchronize()
                    res[name] = round(clips * steps / (e0.elapsed_time(e1) / 1e3), 3)
                except Exception as exc:  # noqa: BLE001  (an OOM of the library path must not kill the bench line)
                    res[name] = f"failed: {type(exc).__name__}: {str(exc)[:120]}"
                    torch.cuda.empty_cache()
            out[variant] = res
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
        torch.cuda.empty_cache()
    return {"unit": "clips/s", "kind": "library-eager PyTorch %s on the GPU (oracle/ functional step, torch ops only)" % torch.__version__,
            "sample": f"{clips} clips of T={args.frames} frames {args.res}x{args.res} per step, {warmup} warm-up + {steps} timed steps",
            "variants": out}


def run_eager(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", 0)
    b = eager_gpu_baseline(args, dev, args.eager_clips, steps=max(1, args.steps), warmup=max(1, args.warmup))
    key = "configs2_lpips" if args.lpips else "configs1_no_lpips"
    v = b["variants"]["bf16_autocast"][key]
    line = {"impl": "eager", "metric": METRIC, "value": v, "unit": "clips/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16 autocast (fp32 variants under extra)", "data": "synthetic",
            "config": workload_config(args, 1, bool(args.lpips)), "extra": {"gpu_eager_baseline": b}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.gpu_index = gpu_index
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if "Active" in v and "Not" not in v:
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ our arm
def summarise_profile(prof, steps, pk):
    """Per-kernel roofline numbers from the live CUDA-event records of faceoff_b200.ops."""
    kern, classes, hbm = {}, {}, {}
    for name, recs in prof.items():
        ms = sum(a.elapsed_time(b) for a, b, _ in recs) / steps
        work = sum(w for _, _, w in recs) / steps
        rec = {"launches_per_step": len(recs) / steps, "ms_per_step": round(ms, 4)}
        if name.startswith("hbm/"):
            gbs = work / (ms / 1e3) / 1e9 if ms > 0 else 0.0
            rec.update(algorithmic_gb_per_step=round(work / 1e9, 4), gb_s=round(gbs, 1), frac_of_hbm_peak=round(gbs / pk["hbm_gbs"], 3))
            hbm[name[4:]] = rec
        else:
            tf = work / (ms / 1e3) / 1e12 if ms > 0 else 0.0
            rec.update(algorithmic_tflop_per_step=round(work / 1e12, 4), tflops=round(tf, 1),
                       frac_of_sustained_peak=round(tf / pk["bf16_tflops_sustained"], 3))
            (classes if "/" in name else kern)[name.split("/", 1)[-1]] = rec
    return kern, classes, hbm


def run_ours(args):
    import torch
    import torch.distributed as dist

    from faceoff_b200 import ops
    from faceoff_b200.losses import mse_loss
    from faceoff_b200.parallel import FusedDataParallel
    from faceoff_b200.vqvae import VQVAE, _LocalStatSink
    from oracle import faceoff_oracle as O  # seeded weights only (plus the cpu_baseline / eager baseline legs)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # same banner line ("NCCL version ..."), nothing else next to the JSON line
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank if world > 1 else 0)
    torch.cuda.set_device(dev)
    assert args.clips % world == 0, "global batch must divide over ranks"
    clips = args.clips // world
    F_ = clips * args.frames
    pk = peaks()

    def new_model():
        m = VQVAE(in_channel=6)
        m.load_state_dict(O.init_vqvae_params(seed=0))
        return m.to(dev).train()

    import warnings

    from faceoff_b200.lpips import VQLPIPS

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        vql = VQLPIPS()
    vql.load_state_dict({"perceptual_loss." + k: v for k, v in O.init_lpips_params(seed=1).items()})
    vql = vql.to(dev)

    def make_step(model, net, n_clips, lpips):
        def step(img_d, gt_d):
            model.zero_grad(set_to_none=True)
            out, latent = net.forward_with_ids(img_d, n_clips)[:2]
            # reference run_step: MSELoss()(out[:, :3], gt) + latent_loss.mean() (+ vqlpips(gt, out[:, :3]))
            loss = mse_loss(out, gt_d) + latent.mean()
            if lpips:
                loss = loss + vql(gt_d, out[:, :3])
            loss.backward()
            return loss
        return step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---------------- dp_check (N > 1): bit-identical replicas + gradients equal to a single-process run ----------------
    dp_check = None
    if world > 1 and not args.no_dp_check:
        T, R = args.frames, args.res

        def clip_of(r):
            g_ = torch.Generator().manual_seed(4321 + r)
            return (torch.empty(T, 6, R, R).uniform_(-1, 1, generator=g_), torch.empty(T, 3, R, R).uniform_(-1, 1, generator=g_))

        m1 = new_model()
        d1 = FusedDataParallel(m1)
        xi, xg = clip_of(rank)
        make_step(m1, d1, 1, True)(xi.to(dev), xg.to(dev))
        torch.cuda.synchronize()
        # bit-level checksums (int32 view summed in int64): the averaged bucket and all six codebook buffers
        sums = [d1._bucket.view(torch.int32).to(torch.int64).sum()]
        sums += [b.contiguous().view(torch.int32).to(torch.int64).sum() for _, b in sorted(m1.named_buffers())]
        mine = torch.stack(sums)
        allv = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        identical = all(torch.equal(allv[0], v) for v in allv)
        worst_norm = worst_max = worst_buf = worst_name = None
        if rank == 0:
            m2 = new_model()
            for q in (m2.quantize_t, m2.quantize_b):
                q.stat_sink = _LocalStatSink()    # this replica must not enter a collective
            pairs = [clip_of(r) for r in range(world)]
            xa = torch.cat([a for a, _ in pairs]).to(dev)
            ga = torch.cat([b for _, b in pairs]).to(dev)
            make_step(m2, m2, world, True)(xa, ga)
            torch.cuda.synchronize()
            g1 = dict(m1.named_parameters())
            worst_norm = worst_max = 0.0
            worst_name = None
            for k, v in m2.named_parameters():
                a, b = g1[k].grad.double(), v.grad.double()
                e_n = abs(a.norm().item() - b.norm().item()) / (b.norm().item() + 1e-30)
                e_m = ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()
                if e_m > worst_max:
                    worst_name = k
                worst_norm, worst_max = max(worst_norm, e_n), max(worst_max, e_m)
            b1 = dict(m1.named_buffers())
            worst_buf = max(((b1[k].double() - v.double()).abs().max() / (v.double().abs().max() + 1e-30)).item()
                            for k, v in m2.named_buffers())
            del m2, xa, ga
        ok = torch.tensor([int(identical and (rank != 0 or (worst_norm <= 1e-3 and worst_max <= 1e-3 and worst_buf <= 1e-5)))],
                          device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        dp_check = {"status": "ok" if ok.item() == 1 else "FAILED", "replicas_bit_identical": identical,
                    "checked": "averaged gradient bucket + 6 codebook buffers (int32-view checksums, all ranks); every "
                               "parameter gradient and codebook vs a single-process run of the same clips on rank 0",
                    "worst_grad_norm_rel_err": worst_norm, "worst_grad_maxnorm_err": worst_max,
                    "worst_grad_tensor": worst_name if rank == 0 else None,
                    "worst_codebook_rel_err": worst_buf, "clips": world, "with_lpips": True,
                    "tolerance": "replicas bit-identical; gradient norm and max-normalised 1e-3, codebooks 1e-5.  The two "
                                 "runs execute the same kernels on bit-identical activations and differ only in how the "
                                 "weight-gradient sums are partitioned: the single-process run accumulates N times more "
                                 "pixels per TMEM accumulator (N x 480 tcgen05 MMAs per split for the Conv3d layers) than a "
                                 "rank does before the fp32 all-reduce, and the tensor core adds products to the fp32 "
                                 "accumulator with truncation, a bias that grows with the chain length: measured 4e-5 at N = 2, "
                                 "3e-4 at N = 8, always on a Conv3d weight and as a norm shrink of the longer chain (DESIGN.md 5)"}
        del m1, d1
        torch.cuda.empty_cache()

    # ---------------- synthetic inputs ----------------
    # The loader's product is uint8 HWC frames (source face, background, ground truth: TemporalAlignment/dataset.py:235-249);
    # the end-to-end leg ships THOSE over PCIe from pinned memory (9 bytes per pixel) and normalises / concatenates on the
    # GPU (faceoff_b200.data.process_data_u8 == utils.process_data + ToTensor + Normalize, bit-exact), SURVEY 8(f4).
    from faceoff_b200.data import process_data_u8

    g = torch.Generator().manual_seed(1234 + rank)
    host_u8 = [[torch.randint(0, 256, (F_, args.res, args.res, 3), generator=g, dtype=torch.uint8).pin_memory()
                for _ in range(3)] for _ in range(2)]          # 2 batches x (source, background, ground truth)
    img, _, gt = process_data_u8(*host_u8[0], device=dev)
    torch.cuda.synchronize()

    def measure(lpips: bool, steps: int, warmup: int, sample_clocks: bool):
        """Device-resident timing + (optionally) the end-to-end leg of one configuration on a fresh model replica."""
        model = new_model()
        net = FusedDataParallel(model) if world > 1 else model
        step = make_step(model, net, clips, lpips)
        torch.cuda.reset_peak_memory_stats(dev)
        ops.PROFILE = None
        for _ in range(warmup):
            step(img, gt)
            torch.cuda.synchronize()
        barrier()
        # The timed region runs WITHOUT the per-launch instrumentation (two CUDA-event records per launch cost the host
        # ~2-3 ms per step -- enough to make a 10 ms step at 4 clips per rank host bound); the per-kernel tables come from a
        # separate instrumented pass below.
        ops.LAUNCHES = 0
        sampler = ClockSampler(local_rank if world > 1 else 0)
        if rank == 0 and sample_clocks:
            sampler.start()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        barrier()
        ev[0].record()
        t_host0 = time.perf_counter()
        for _ in range(steps):
            step(img, gt)
        ev[1].record()
        host_ms = (time.perf_counter() - t_host0) * 1e3 / steps   # host enqueue time per step (no sync inside)
        barrier()
        clocks = sampler.stop() if (rank == 0 and sample_clocks) else None
        total_ms = max_over_ranks(ev[0].elapsed_time(ev[1]))
        launches = ops.LAUNCHES
        # instrumented pass: per-kernel CUDA-event timing (events are created in a first step and recycled)
        prof_steps = min(steps, 4)
        ops.PROFILE = {}
        step(img, gt)
        torch.cuda.synchronize()
        ops.recycle_events(ops.PROFILE)
        ops.PROFILE = {}
        for _ in range(prof_steps):
            step(img, gt)
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        kern, classes, hbm = summarise_profile(prof, prof_steps, pk)
        ops.recycle_events(prof)
        barrier()
        res = {"value": args.clips * steps / (total_ms / 1e3), "ms_per_step": total_ms / steps, "kernels": kern,
               "classes": classes, "hbm": hbm, "launches": launches, "host_ms": host_ms, "clocks": clocks, "e2e": None}

        # ---- end-to-end: host buffers in, loss out, copies inside the timed region ----
        if not args.no_e2e:
            copy_stream = torch.cuda.Stream(device=dev)
            dbuf = [[torch.empty_like(t, device=dev) for t in host_u8[0]] for _ in range(2)]
            fbuf = (torch.empty_like(img), torch.empty_like(gt))       # normalised fp32 inputs of the running step
            ready = [torch.cuda.Event() for _ in range(2)]
            consumed = [torch.cuda.Event() for _ in range(2)]

            def prefetch(i):
                b = i % 2
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[b])
                    for d_, h_ in zip(dbuf[b], host_u8[b]):
                        d_.copy_(h_, non_blocking=True)
                    ready[b].record(copy_stream)

            loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
            for b in range(2):
                consumed[b].record()

            def e2e_loop(n):
                prefetch(0)
                for i in range(n):
                    if i + 1 < n:
                        prefetch(i + 1)
                    b = i % 2
                    torch.cuda.current_stream().wait_event(ready[b])
                    img_d, _, gt_d = process_data_u8(*dbuf[b], out_img=fbuf[0], out_gt=fbuf[1])
                    consumed[b].record()
                    loss = step(img_d, gt_d)
                    loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
                torch.cuda.synchronize()

            e2e_loop(2)
            barrier()
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            e2e_loop(steps)
            e1.record()
            barrier()
            e_ms = max_over_ranks(e0.elapsed_time(e1))
            res["e2e"] = {"value": args.clips * steps / (e_ms / 1e3), "unit": "clips/s",
                          "h2d_bytes_per_step": sum(t.numel() for t in host_u8[0]) * world,
                          "input": "uint8 HWC frames (source, background, ground truth) from pinned memory; normalise + "
                                   "concat on the GPU (process_data_u8)",
                          "d2h_bytes_per_step": 4 * world, "wall_s": time.perf_counter() - t0}
            del dbuf, fbuf
        res["peak_mem_gb"] = torch.cuda.max_memory_allocated(dev) / 2 ** 30
        del model, net, step
        torch.cuda.empty_cache()
        return res

    headline_lpips = bool(args.lpips)
    import contextlib
    # FO_BENCH_SIDE_STREAM=1 (experiments): run the steps on a non-default stream
    side = torch.cuda.Stream(device=dev) if os.environ.get("FO_BENCH_SIDE_STREAM") == "1" else None
    with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
        main = measure(headline_lpips, args.steps, args.warmup, sample_clocks=True)
        other = None
        if not args.no_lpips_step and not headline_lpips:
            other = measure(True, args.steps, max(3, args.warmup), sample_clocks=True)

    def roofline_of(m):
        """Dominant kernel = conv_igemm_kernel over EVERY launch of the step (tensor-bound and HBM-bound layers alike);
        the best layer class and the ncu traffic of its launch shape are given next to it."""
        k = m["kernels"].get("conv_igemm")
        if k is None:
            return None
        evd = ncu_evidence() or {}
        conv_classes = {c: v for c, v in m["classes"].items() if c.startswith("conv")}
        best = max(conv_classes.items(), key=lambda kv: kv[1]["ms_per_step"]) if conv_classes else (None, None)
        return {"kernel": "conv_igemm_kernel (all launches of the step)", "bound": "tensor", "achieved": k["tflops"],
                "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": k["tflops"] / pk["bf16_tflops_sustained"],
                "frac_of_burst_peak": k["tflops"] / pk["bf16_tflops"],
                "traffic": evd.get("dram_bytes_per_launch"), "traffic_note": evd.get("traffic_note"),
                "peak_source": pk["source"] + " (sustained bf16 cuBLAS: the kernel is timed inside a long step)",
                "kernel_timing": "CUDA events around every launch on the launching stream, in an instrumented pass of up to 4 "
                                 "full steps run right after the timed region (the timed region itself carries no per-launch "
                                 "instrumentation)",
                "launches_per_step": k["launches_per_step"], "kernel_ms_per_step": k["ms_per_step"],
                "algorithmic_flop_per_step": k["algorithmic_tflop_per_step"] * 1e12,
                "largest_layer_class": None if best[0] is None else dict(best[1], layer_class=best[0]),
                "ncu_evidence": evd}

    def disc_step_bench(steps=5, warmup=3):
        """BASELINE configs[4] addendum: the MoCoGAN-HD discriminator step of one clip at full size (trainer
        disc_trainers/train_vqvae_perceptual_mocoganhd_disc.py:240-300): video discriminator on the 11 frame pairs
        [1, 6, 11, 256, 256] (fake + real forward, relativistic average LSGAN, backward) and image discriminator on one
        frame pair [1, 6, 256, 256].  Timed in both convolution modes: the product path (im2col + split-bf16 GEMMs on the
        tcgen05 kernel, csrc/disc_gemm.cu) and the exact-fp32 FFMA kernels (csrc/disc.cu)."""
        from faceoff_b200.mocoganhd import content_disc, layers, losses, video_disc

        torch.manual_seed(0)
        d3 = video_disc.ModelD_3d(3, "instance", 2, 1e-4, False, 12).to(dev).train()
        d2 = content_disc.ModelD_img(3, "instance", 2, 1e-4).to(dev).train()
        crit = losses.Relativistic_Average_LSGAN()
        gq = torch.Generator().manual_seed(3)
        v_real = (torch.rand(1, 6, 11, args.res, args.res, generator=gq) * 2 - 1).to(dev)
        v_fake = (torch.rand(1, 6, 11, args.res, args.res, generator=gq) * 2 - 1).to(dev)
        i_real, i_fake = v_real[:, :, 0].contiguous(), v_fake[:, :, 0].contiguous()

        def conv_flops(m, x_shape):
            tot, shp = 0.0, x_shape
            for j in range(5):
                w = getattr(m.netD, f"scale1_layer{j}")[0].weight
                s = 2 if j < 3 else 1
                sp = [(v + 4 - 4) // s + 1 for v in shp]
                tot += 2.0 * w.numel() * float(torch.tensor(sp).prod())
                shp = sp
            return tot

        fl3 = conv_flops(d3, [11, args.res, args.res]) + conv_flops(d3, [11, args.res // 2, args.res // 2])
        fl2 = conv_flops(d2, [args.res, args.res]) + conv_flops(d2, [args.res // 2, args.res // 2])
        flops = (fl3 + fl2) * 2 * 3      # fake + real, forward + data gradient + weight gradient (upper bound: the first layer has no dgrad)

        def one():
            for d, fake, real in ((d3, v_fake, v_real), (d2, i_fake, i_real)):
                o_f, o_r = d(fake), d(real)
                loss = (crit(o_r, o_f, True) + crit(o_f, o_r, False)) * 0.5
                d.optim.zero_grad()
                loss.backward()
            return loss

        def timed():
            for _ in range(warmup):
                one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                one()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps

        ms = timed()
        layers.TENSOR_CORE = False
        try:
            ms32 = timed()
        finally:
            layers.TENSOR_CORE = True
        return {"workload": "MoCoGAN-HD discriminator step, 1 clip: D_3d on [1,6,11,256,256] + D_img on [1,6,256,256], "
                            "fake+real forward, RA-LSGAN, backward (configs[4] component, SURVEY 8(f1))",
                "ms_per_step": ms, "algorithmic_tflop_per_step": flops / 1e12, "tflops": flops / ms / 1e9,
                "dtype": "split-bf16 (hi|lo) GEMMs on tcgen05 with fp32 accumulation over an explicit im2col matrix: 3 MMAs per "
                         "algorithmic product, ~2e-5 against fp64 (tests/test_gpu_disc.py)",
                "fp32_ffma_mode": {"ms_per_step": ms32, "tflops": flops / ms32 / 1e9},
                "steps": steps, "warmup": warmup}

    cpu_baseline = None
    extra = {}
    if rank == 0 and world == 1 and not args.no_disc_step:
        extra["disc_step"] = disc_step_bench()
    if rank == 0 and world == 1:
        if not args.no_cpu_baseline:
            sec, cores = cpu_train_step_time(args.cpu_clip_frames, args.res, headline_lpips, steps=3, warmup=1)
            cpu_baseline = {"value": (args.cpu_clip_frames / args.frames) / sec, "unit": "clips/s", "cores": cores,
                            "kind": "port",
                            "sample": f"1 clip of T={args.cpu_clip_frames} frames {args.res}x{args.res}, 1 warm-up + 3 timed "
                                      f"steps of the oracle port (torch CPU fp32, {cores} threads)"}
        if not args.no_eager:
            extra["gpu_eager_baseline"] = eager_gpu_baseline(args, dev, args.eager_clips)

    if rank == 0:
        def classes_top(m, n=16):
            return dict(sorted(m["classes"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:n])

        line = {
            "metric": METRIC, "value": main["value"], "unit": "clips/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world, headline_lpips), "clocks": main["clocks"], "e2e": main["e2e"],
            "gpu_launches": main["launches"], "roofline": roofline_of(main),
            "roofline_by_layer_class": classes_top(main), "hbm_kernels": main["hbm"],
            "cpu_baseline": cpu_baseline, "kernels": main["kernels"], "peak_mem_gb": main["peak_mem_gb"],
            "host_enqueue_ms_per_step": main["host_ms"],
        }
        if other is not None:
            line["lpips_step"] = {
                "config": workload_config(args, world, True), "value": other["value"], "unit": "clips/s",
                "ms_per_step": other["ms_per_step"], "e2e": other["e2e"], "gpu_launches": other["launches"],
                "clocks": other["clocks"], "roofline": roofline_of(other), "roofline_by_layer_class": classes_top(other, 24),
                "hbm_kernels": other["hbm"], "kernels": other["kernels"], "peak_mem_gb": other["peak_mem_gb"],
                "host_enqueue_ms_per_step": other["host_ms"]}
        if dp_check is not None:
            line["dp_check"] = dp_check
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "eager":
        run_eager(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
