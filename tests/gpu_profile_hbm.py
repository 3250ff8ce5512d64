"""One launch of every HBM-bound kernel of the step at production per-frame shapes (default 8 clips = 240 frames), for
ncu captures (dram__bytes vs algorithmic bytes) and CUDA-event bandwidth numbers.

    python tests/gpu_profile_hbm.py [clips] [--time]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from faceoff_b200 import ops  # noqa: E402


def main():
    clips = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
    timed = "--time" in sys.argv
    F_ = clips * 30
    dev = "cuda"
    torch.manual_seed(0)
    bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)   # noqa: E731
    x = bf(F_, 64, 64, 128)
    h = bf(F_, 64, 64, 32).relu_()
    b128 = torch.zeros(128, device=dev)
    w1 = torch.randn(128, 32, 1, 1, device=dev) * 0.1
    w32 = torch.randn(32, 128, 3, 3, device=dev) * 0.03
    rows = F_ * 64 * 64
    xq = torch.randn(rows, 64, device=dev)
    e = torch.randn(64, 512, device=dev)
    e_split, e_t, e_n2 = ops.vq_prep(e)
    ind = torch.randint(0, 512, (rows,), device=dev)
    gq = bf(rows, 128)
    f0, f1 = bf(F_, 256, 256, 64).relu_(), bf(F_, 256, 256, 64).relu_()
    lw = torch.rand(64, device=dev)
    out = torch.zeros(F_, device=dev)
    g = torch.ones(F_, device=dev)
    dec = torch.randn(F_, 6, 256, 256, device=dev)
    gt = torch.randn(F_, 3, 256, 256, device=dev)
    one = torch.ones(1, device=dev)
    img = torch.rand(F_, 3, 256, 256, device=dev)
    db = torch.zeros(128, device=dev)
    cases = [
        ("conv1x1 32->128 +res raw+relu", (h.numel() + 3 * x.numel()) * 2,
         lambda: ops.conv(ops.FORM_S1, 2, 1, [(h, 32, 0)], w1, 0, 128, bias=b128, addend=x, want_raw=True, want_relu=True)),
        ("dgrad3x3 32->128 +mask+addend", (h.numel() + 3 * x.numel()) * 2,
         lambda: ops.conv(ops.FORM_S1_DGRAD, 2, 3, [(h, 32, 0)], w32, 1, 128, mask=x, addend=x)),
        ("dgrad1x1 128->32 +mask", (x.numel() + 2 * h.numel()) * 2,
         lambda: ops.conv(ops.FORM_S1_DGRAD, 2, 1, [(x, 128, 0)], w1, 1, 32, mask=h)),
        ("vq_gather_stats", rows * (256 + 8 + 256 + 128),
         lambda: ops.vq_gather_stats(xq, ind, e_t, torch.zeros(1, device=dev), torch.zeros(512, device=dev),
                                     torch.zeros(64, 512, device=dev), want_f32=True, want_bf16=True)),
        ("vq_backward", rows * (128 + 256 + 8 + 128),
         lambda: ops.vq_backward(gq, 0, one, xq, ind, e_t, want_f32=False, want_bf16=True)),
        ("lpips_tap C=64", 4 * f0.numel(), lambda: ops.lpips_tap(f0, f1, lw, out)),
        ("lpips_tap_bwd C=64", 6 * f0.numel(), lambda: ops.lpips_tap_bwd(f0, f1, lw, g)),
        ("maxpool2 C=64", 2.5 * f0.numel(), lambda: ops.maxpool2(f0)),
        ("mse", 8 * gt.numel(), lambda: ops.mse_sum(dec, gt)),
        ("mse_grad", 8 * gt.numel() + 4 * dec.numel(), lambda: ops.mse_grad(dec, gt, one, 1e-3)),
        ("im2col3x3", img.numel() * 4 + F_ * 65536 * 64, lambda: ops.im2col3x3(img)),
        ("colsum 128", x.numel() * 2, lambda: ops.colsum(x, 128, db)),
    ]
    for name, nbytes, fn in cases:
        fn()
        torch.cuda.synchronize()
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                fn()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            print(f"{name:34s} {ms:8.3f} ms  {nbytes / ms / 1e6:8.0f} GB/s  (algorithmic {nbytes / 1e9:.3f} GB)")
        else:
            fn()
            torch.cuda.synchronize()
            print(f"{name:34s} algorithmic {nbytes / 1e9:.3f} GB")


if __name__ == "__main__":
    main()
