"""LPIPS head kernels (reference models/lpips.py:80-93,155-161) at production tap shapes: achieved HBM GB/s.

    python tests/gpu_profile_lpips.py [frames]
Algorithmic bytes: forward reads both bf16 feature maps once; backward reads both and writes one gradient.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from faceoff_b200 import ops  # noqa: E402


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 240
    print(f"frames = {frames}")
    print("| tap | C | HxW | fwd ms | fwd GB/s | bwd ms | bwd GB/s |")
    print("|---:|---:|---:|---:|---:|---:|---:|")
    for tap, (c, s) in enumerate([(64, 256), (128, 128), (256, 64), (512, 32), (512, 16)]):
        torch.manual_seed(tap)
        f0 = torch.randn(frames, s, s, c, device="cuda").to(torch.bfloat16)
        f1 = torch.randn(frames, s, s, c, device="cuda").to(torch.bfloat16)
        w = torch.rand(c, device="cuda")
        out = torch.zeros(frames, device="cuda")
        g = torch.full((frames,), 1.0 / frames, device="cuda")
        ms_f = timeit(lambda: ops.lpips_tap(f0, f1, w, out))
        ms_b = timeit(lambda: ops.lpips_tap_bwd(f0, f1, w, g))
        nbytes = f0.numel() * 2
        print(f"| {tap} | {c} | {s}x{s} | {ms_f:.3f} | {2 * nbytes / ms_f / 1e6:.0f} | {ms_b:.3f} | {3 * nbytes / ms_b / 1e6:.0f} |")
        del f0, f1


if __name__ == "__main__":
    main()
