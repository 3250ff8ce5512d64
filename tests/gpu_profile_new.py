"""Isolated launches of the kernels added in the last third of round 2 (ncu DRAM bytes / CUDA-event times).

    python tests/gpu_profile_new.py [clips]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from faceoff_b200 import ops  # noqa: E402


def timeit(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    clips = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    F_ = clips * 30
    dev = torch.device("cuda", 0)
    dy = (torch.randn(F_, 256, 256, 64, device=dev) * 0.1).to(torch.bfloat16)
    w3 = torch.randn(64, 3, 3, 3, device=dev) * 0.1
    sc = torch.tensor([.458, .448, .45], device=dev)
    ms = timeit(lambda: ops.vgg_first_dgrad(dy, w3, sc))
    gb = (dy.numel() * 2 + F_ * 3 * 65536 * 4) / 1e9
    print(f"vgg_first_dgrad {F_} frames: {ms:.3f} ms, algorithmic {gb:.3f} GB, {gb / ms * 1e3:.0f} GB/s")
    del dy
    for c, r in ((64, 256), (128, 128), (256, 64), (512, 32)):
        f0 = torch.randn(F_, r, r, c, device=dev).relu_().to(torch.bfloat16)
        f1 = torch.randn(F_, r, r, c, device=dev).relu_().to(torch.bfloat16)
        lw = torch.rand(c, device=dev) * 0.1
        g = torch.randn(F_, device=dev)
        pdy = (torch.randn(F_, r // 2, r // 2, c, device=dev) * 0.01).to(torch.bfloat16)
        ms = timeit(lambda: ops.lpips_tap_bwd_pool(f0, f1, lw, g, pdy))
        gb = 6.5 * f0.numel() / 1e9
        y = ops.maxpool2(f0)
        ms2 = timeit(lambda: ops.lpips_tap_bwd(f0, f1, lw, g, ops.maxpool2_bwd(f0, y, pdy)))
        print(f"lpips_tap_bwd_pool c={c} @{r}: {ms:.3f} ms, algorithmic {gb:.3f} GB, {gb / ms * 1e3:.0f} GB/s "
              f"(maxpool2_bwd + lpips_tap_bwd: {ms2:.3f} ms, {13.0 * f0.numel() / 1e9:.3f} GB)")
        del f0, f1, pdy, y


if __name__ == "__main__":
    main()
