"""Quantiser sweep (BASELINE configs[3]): n_embed 512/1024/2048 x embed_dim 64/128 at N = 122,880 x clips rows.

    python tests/gpu_profile_vq.py [clips [dim n_embed]]   -> one markdown table row per configuration
Reports the assign kernel (tcgen05 split-bf16 GEMM + argmin + exact re-check) in algorithmic TFLOP/s (2*D*K per row)
and GB/s (4D + 8 bytes per row), and the fused gather/ST/loss/stats kernel in GB/s (8D + 8 (+2D bf16) bytes per row).
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
    clips = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rows = 122880 * clips
    print(f"rows = {rows}")
    print("| dim | n_embed | assign ms | TFLOP/s (2DK/row) | GB/s (4D+8 B/row) | rows re-checked | gather+stats ms | GB/s |")
    print("|---:|---:|---:|---:|---:|---:|---:|---:|")
    only = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else None
    for dim in (64, 128):
        for K in (512, 1024, 2048):
            if only is not None and (dim, K) != only:
                continue
            torch.manual_seed(0)
            x = torch.randn(rows, dim, device="cuda")
            e = torch.randn(dim, K, device="cuda")
            e_split, e_t, e_n2 = ops.vq_prep(e)
            nf = torch.zeros(1, dtype=torch.int32, device="cuda")
            ind = ops.vq_assign(x, e_t, e_split, e_n2, nf)
            ms = timeit(lambda: ops.vq_assign(x, e_t, e_split, e_n2, nf))
            diff = torch.zeros(1, device="cuda")
            counts = torch.zeros(K, device="cuda")
            esum = torch.zeros(dim, K, device="cuda")
            ms2 = timeit(lambda: ops.vq_gather_stats(x, ind, e_t, diff, counts, esum, want_f32=True, want_bf16=True))
            tf = 2.0 * rows * dim * K / ms / 1e9
            gbs = rows * (4 * dim + 8) / ms / 1e6
            gbs2 = rows * (4 * dim + 8 + 4 * dim + 2 * dim) / ms2 / 1e6
            print(f"| {dim} | {K} | {ms:.3f} | {tf:.1f} | {gbs:.0f} | {nf.item()} | {ms2:.3f} | {gbs2:.0f} |")


if __name__ == "__main__":
    main()
