#!/usr/bin/env python
"""Benchmark of the FaceOff VQVAE-conv3d training step (fwd+bwd) on B200 -- BASELINE.json metric:
"train clips/sec (b32, 256^2, fwd+bwd)".

    python bench.py [--gpus N] [--steps K] [--warmup W]          # our arm (one process per GPU; torchrun for N>1)
    python bench.py --impl reference [--gpus N] ...               # the reference's CPU implementation of the path

Workload (config.workload): BASELINE.json configs[1] -- VQVAE(in_channel=6) train step without perceptual loss,
global batch 32 clips x T=30 frames of 256x256 synthetic data, random-init weights (seeded).  For N>1 the 32 clips are
sharded over the ranks (strong scaling) and the fused EMA+gradient all-reduce runs every step.
One JSON line is printed by rank 0 (see the contract in the task description).
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=GLOBAL_CLIPS, help="global batch in clips")
    ap.add_argument("--frames", type=int, default=T_FRAMES)
    ap.add_argument("--res", type=int, default=RES)
    ap.add_argument("--lpips", type=int, default=0, help="add the LPIPS perceptual loss (BASELINE configs[2])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-clip-frames", type=int, default=T_FRAMES)
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


def ncu_evidence():
    """Static pointer to the committed ncu capture of the dominant kernel (profiles/), not a live measurement."""
    path = os.path.join(ROOT, "profiles", "r1_roofline_evidence.json")
    return json.load(open(path)) if os.path.exists(path) else None


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
        "impl": "reference", "metric": "train clips/sec (b32, 256^2, fwd+bwd)", "value": value, "unit": "clips/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n):
    return {"workload": "VQVAE-conv3d train step (fwd+bwd), BASELINE configs[1]" + (" + LPIPS (configs[2])" if args.lpips else ""),
            "global_batch_clips": args.clips, "frames_per_clip": args.frames, "resolution": args.res,
            "in_channel": 6, "embed_dim": 64, "n_embed": 512, "parallelism": f"dp{n}",
            "l2": "inputs (>= 1.5 GB/step) and activations (~30 GB/step) are far larger than the 126 MB L2"}


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
def run_ours(args):
    import torch
    import torch.distributed as dist

    from faceoff_b200 import ops
    from faceoff_b200.losses import mse_loss
    from faceoff_b200.parallel import FusedDataParallel
    from faceoff_b200.vqvae import VQVAE
    from oracle import faceoff_oracle as O  # synthetic data + seeded weights only (and the cpu_baseline leg)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the single JSON line
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank if world > 1 else 0)
    torch.cuda.set_device(dev)
    assert args.clips % world == 0, "global batch must divide over ranks"
    clips = args.clips // world
    F_ = clips * args.frames

    model = VQVAE(in_channel=6)
    model.load_state_dict(O.init_vqvae_params(seed=0))
    model = model.to(dev).train()
    net = FusedDataParallel(model) if world > 1 else model
    vql = None
    if args.lpips:
        import warnings

        from faceoff_b200.lpips import VQLPIPS

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            vql = VQLPIPS()
        vql.load_state_dict({"perceptual_loss." + k: v for k, v in O.init_lpips_params(seed=1).items()})
        vql = vql.to(dev)

    g = torch.Generator().manual_seed(1234 + rank)
    host_img = [torch.empty(F_, 6, args.res, args.res).uniform_(-1, 1, generator=g).pin_memory() for _ in range(2)]
    host_gt = [torch.empty(F_, 3, args.res, args.res).uniform_(-1, 1, generator=g).pin_memory() for _ in range(2)]
    img = host_img[0].to(dev)
    gt = host_gt[0].to(dev)

    def step(img_d, gt_d):
        model.zero_grad(set_to_none=True)
        out, latent = net.forward_with_ids(img_d, clips)[:2]
        # reference run_step: MSELoss()(out[:, :3], gt) + latent_loss.mean() (+ vqlpips(gt, out[:, :3]))
        loss = mse_loss(out, gt_d) + latent.mean()
        if vql is not None:
            loss = loss + vql(gt_d, out[:, :3])
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ----------------
    ops.PROFILE = {}  # the warm-up steps also create the (recycled) timing events, so the timed steps only record them
    for _ in range(args.warmup):
        step(img, gt)
        torch.cuda.synchronize()
        ops.recycle_events(ops.PROFILE)
        ops.PROFILE = {}
    barrier()
    ops.PROFILE = {}  # per-kernel CUDA-event timing inside the timed region
    ops.LAUNCHES = 0
    sampler = ClockSampler(local_rank if world > 1 else 0)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    t_host0 = time.perf_counter()
    for i in range(args.steps):
        step(img, gt)
        ev[i + 1].record()
    host_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps   # host enqueue time per step (no sync inside)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    prof = ops.PROFILE
    ops.PROFILE = None
    launches = ops.LAUNCHES
    tms = torch.tensor([total_ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    total_ms = tms.item()
    value = args.clips * args.steps / (total_ms / 1e3)

    # per-kernel roofline from the live events
    kern = {}
    for name, recs in prof.items():
        ms = sum(a.elapsed_time(b) for a, b, _ in recs)
        work = sum(w for _, _, w in recs)
        kern[name] = {"launches": len(recs), "ms_per_step": ms / args.steps, "work_per_step": work / args.steps}
    pk = peaks()
    groups = {k.split("/", 1)[1]: v for k, v in kern.items() if "/" in k}   # conv_igemm/<class>, wgrad_igemm/<class>
    kern = {k: v for k, v in kern.items() if "/" not in k}
    for gname, gv in groups.items():
        gv["tflops"] = gv["work_per_step"] / (gv["ms_per_step"] / 1e3) / 1e12
    # roofline of the dominant kernel: the conv_igemm launches of the layer class with the largest share of the step
    # (the Conv3d 128->128 latent blocks, SURVEY 8(d): 2*27*128*128 FLOP per output voxel), live CUDA-event time;
    # the aggregate over every conv_igemm launch (incl. the HBM-bound 1x1 / 32-channel layers) is reported next to it
    roofline = None
    conv_groups = {g_: v_ for g_, v_ in groups.items() if g_.startswith("conv")}
    if conv_groups:
        gname, gv = max(conv_groups.items(), key=lambda kv: kv[1]["ms_per_step"])
        ev = ncu_evidence() or {}
        k = kern["conv_igemm"]
        ach_all = k["work_per_step"] / (k["ms_per_step"] / 1e3) / 1e12
        roofline = {"kernel": f"conv_igemm_kernel [{gname}]", "bound": "tensor", "achieved": gv["tflops"],
                    "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": gv["tflops"] / pk["bf16_tflops_sustained"],
                    "frac_of_burst_peak": gv["tflops"] / pk["bf16_tflops"],
                    "traffic": ev.get("dram_bytes_per_launch"),
                    "traffic_note": ev.get("traffic_note"),
                    "peak_source": pk["source"] + " (sustained bf16 cuBLAS: the kernel is timed inside a long step)",
                    "launches_per_step": gv["launches"] / args.steps, "kernel_ms_per_step": gv["ms_per_step"],
                    "algorithmic_flop_per_step": gv["work_per_step"],
                    "all_conv_igemm_launches": {
                        "achieved": ach_all, "frac": ach_all / pk["bf16_tflops_sustained"],
                        "launches_per_step": k["launches"] / args.steps, "kernel_ms_per_step": k["ms_per_step"],
                        "note": "every launch of the kernel in the step, including HBM-bound layers (1x1, 32- and "
                                "6-channel); per layer class see roofline_by_layer_class"},
                    "ncu_evidence": ev}

    # ---------------- end-to-end: host buffers in, loss out, copies inside the timed region ----------------
    e2e = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream(device=dev)
        dbuf = [(torch.empty_like(img), torch.empty_like(gt)) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            b = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[b])
                dbuf[b][0].copy_(host_img[b], non_blocking=True)
                dbuf[b][1].copy_(host_gt[b], non_blocking=True)
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
                loss = step(dbuf[b][0], dbuf[b][1])
                consumed[b].record()
                loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
            torch.cuda.synchronize()

        e2e_loop(2)
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_loop(args.steps)
        e1.record()
        barrier()
        e_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
        e2e = {"value": args.clips * args.steps / (e_ms.item() / 1e3), "unit": "clips/s",
               "h2d_bytes_per_step": (host_img[0].numel() + host_gt[0].numel()) * 4 * world, "d2h_bytes_per_step": 4 * world,
               "wall_s": time.perf_counter() - t0}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, cores = cpu_train_step_time(args.cpu_clip_frames, args.res, bool(args.lpips), steps=2, warmup=1)
        cpu_baseline = {"value": (args.cpu_clip_frames / args.frames) / sec, "unit": "clips/s", "cores": cores,
                        "kind": "port",
                        "sample": f"1 clip of T={args.cpu_clip_frames} frames {args.res}x{args.res}, 1 warm-up + 2 timed steps "
                                  f"of the oracle port (torch CPU fp32, {cores} threads)"}

    if rank == 0:
        line = {
            "metric": "train clips/sec (b32, 256^2, fwd+bwd)", "value": value, "unit": "clips/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "roofline_by_layer_class": {
                g_: {"ms_per_step": round(v_["ms_per_step"], 3), "tflops": round(v_["tflops"], 1),
                     "frac_of_sustained_peak": round(v_["tflops"] / pk["bf16_tflops_sustained"], 3)}
                for g_, v_ in sorted(groups.items(), key=lambda kv: -kv[1]["ms_per_step"])[:14]},
            "cpu_baseline": cpu_baseline, "kernels": kern,
            "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30, "host_enqueue_ms_per_step": host_ms,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
