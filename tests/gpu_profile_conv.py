"""Stand-alone launches of the hot conv / wgrad kernels at production shapes, for ncu captures and A/B timing.

    python tests/gpu_profile_conv.py [conv3d|conv3x3|resblock|wgrad_down|wgrad3d|all] [clips]
"""
import os
import sys
import time

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
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    clips = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    T = 30
    dev = "cuda"
    torch.manual_seed(0)
    if which in ("conv3d", "all"):
        x = torch.randn(clips, T, 64, 64, 128, device=dev).to(torch.bfloat16)
        w = torch.randn(128, 128, 3, 3, 3, device=dev) * 0.02
        b = torch.zeros(128, device=dev)
        fl = 2.0 * clips * T * 64 * 64 * 128 * 128 * 27
        ms = timeit(lambda: ops.conv(ops.FORM_S1, 3, 3, [(x, 128, 0)], w, 0, 128, bias=b, want_raw=False, want_relu=True))
        print(f"conv3d fwd  [{clips}x{T}x64x64x128]: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
        ms = timeit(lambda: ops.conv(ops.FORM_S1_DGRAD, 3, 3, [(x, 128, 0)], w, 1, 128, mask=x, addend=x))
        print(f"conv3d dgrad(+mask+addend): {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
    if which in ("conv3x3", "all"):
        F_ = clips * T
        x = torch.randn(F_, 64, 64, 128, device=dev).to(torch.bfloat16)
        w = torch.randn(128, 128, 3, 3, device=dev) * 0.03
        b = torch.zeros(128, device=dev)
        fl = 2.0 * F_ * 64 * 64 * 128 * 128 * 9
        ms = timeit(lambda: ops.conv(ops.FORM_S1, 2, 3, [(x, 128, 0)], w, 0, 128, bias=b, want_raw=True, want_relu=True))
        print(f"conv3x3 128->128 @64 fwd: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
        w32 = torch.randn(32, 128, 3, 3, device=dev) * 0.03
        fl32 = fl / 4
        ms = timeit(lambda: ops.conv(ops.FORM_S1, 2, 3, [(x, 128, 0)], w32, 0, 32, bias=b, want_raw=False, want_relu=True))
        print(f"conv3x3 128->32 @64 fwd: {ms:.3f} ms  {fl32 / ms / 1e9:.1f} TFLOP/s")
        w1 = torch.randn(128, 32, 1, 1, device=dev) * 0.1
        h = torch.randn(F_, 64, 64, 32, device=dev).to(torch.bfloat16)
        ms = timeit(lambda: ops.conv(ops.FORM_S1, 2, 1, [(h, 32, 0)], w1, 0, 128, bias=b, addend=x, want_raw=True, want_relu=True))
        gb = (h.numel() + 3 * x.numel()) * 2 / 1e9
        print(f"conv1x1 32->128 (+res, raw+relu) @64: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s")
        wd = torch.randn(128, 64, 4, 4, device=dev) * 0.03
        xd = torch.randn(F_, 128, 128, 64, device=dev).to(torch.bfloat16)
        fld = 2.0 * F_ * 64 * 64 * 128 * 64 * 16
        ms = timeit(lambda: ops.conv(ops.FORM_DOWN, 2, 4, [(xd, 64, 0)], wd, 0, 128, bias=b, want_raw=False, want_relu=True))
        print(f"conv4x4s2 64->128 (DOWN) : {ms:.3f} ms  {fld / ms / 1e9:.1f} TFLOP/s")
        wu = torch.randn(128, 64, 4, 4, device=dev) * 0.03
        ms = timeit(lambda: ops.conv(ops.FORM_UP, 2, 4, [(x, 128, 0)], wu, 1, 64, bias=b, want_raw=False, want_relu=True))
        print(f"convT4x4s2 128->64 (UP)  : {ms:.3f} ms  {fld / ms / 1e9:.1f} TFLOP/s")
    if which in ("resblock", "all"):
        F_ = clips * T
        x = torch.randn(F_, 64, 64, 128, device=dev).to(torch.bfloat16)
        dh = torch.randn(F_, 64, 64, 32, device=dev).to(torch.bfloat16)
        w32 = torch.randn(32, 128, 3, 3, device=dev) * 0.03
        ms = timeit(lambda: ops.conv(ops.FORM_S1_DGRAD, 2, 3, [(dh, 32, 0)], w32, 1, 128, mask=x, addend=x))
        gb = (dh.numel() + 3 * x.numel()) * 2 / 1e9
        print(f"resblock dgrad3x3 32->128 (+mask+addend): {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s")
        w1 = torch.randn(128, 32, 1, 1, device=dev) * 0.1
        ms = timeit(lambda: ops.conv(ops.FORM_S1_DGRAD, 2, 1, [(x, 128, 0)], w1, 1, 32, mask=dh))
        gb = (x.numel() + 2 * dh.numel()) * 2 / 1e9
        print(f"resblock dgrad1x1 128->32 (+mask): {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s")
        dw = torch.empty(32, 128, 3, 3, device=dev)
        ms = timeit(lambda: ops.wgrad(ops.FORM_S1, 2, 3, (x, 128, 0), (dh, 32, 0), dw, m_axis=1, q_shift_sign=-1))
        print(f"resblock wgrad3x3 (swapped): {ms:.3f} ms")
        ms = timeit(lambda: ops.wgrad(ops.FORM_S1, 2, 3, (dh, 32, 0), (x, 128, 0), dw, m_axis=0))
        print(f"resblock wgrad3x3 (P=dy): {ms:.3f} ms")
    if which in ("wgrad_down", "all"):
        F_ = clips * T
        dy = torch.randn(F_, 64, 64, 128, device=dev).to(torch.bfloat16)
        xh = torch.randn(F_, 128, 128, 64, device=dev).to(torch.bfloat16)
        dw = torch.empty(128, 64, 4, 4, device=dev)
        db = torch.zeros(128, device=dev)
        fl = 2.0 * F_ * 64 * 64 * 128 * 64 * 16
        ms = timeit(lambda: ops.wgrad(ops.FORM_DOWN, 2, 4, (dy, 128, 0), (xh, 64, 0), dw, m_axis=0, dbias=db))
        print(f"wgrad4x4s2 128x64 (enc conv 64->128 @128): {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
        xl = torch.randn(F_, 64, 64, 128, device=dev).to(torch.bfloat16)
        dyh = torch.randn(F_, 128, 128, 64, device=dev).to(torch.bfloat16)
        dwt = torch.empty(128, 64, 4, 4, device=dev)
        ms = timeit(lambda: ops.wgrad(ops.FORM_DOWN, 2, 4, (xl, 128, 0), (dyh, 64, 0), dwt, m_axis=0))
        print(f"wgrad4x4s2 128x64 (dec convT 128->64 @64): {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
    if which in ("vgg", "all"):
        F_ = clips * T
        x = torch.randn(F_, 256, 256, 64, device=dev).to(torch.bfloat16).relu_()
        w = torch.randn(64, 64, 3, 3, device=dev) * 0.03
        b = torch.zeros(64, device=dev)
        fl = 2.0 * F_ * 256 * 256 * 64 * 64 * 9
        ms = timeit(lambda: ops.conv(ops.FORM_S1, 2, 3, [(x, 64, 0)], w, 0, 64, bias=b, want_raw=False, want_relu=True))
        print(f"vgg conv3x3 64->64 @256 fwd: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
        col = torch.randn(F_, 256, 256, 32, device=dev).to(torch.bfloat16)
        w1 = torch.randn(64, 32, 1, 1, device=dev) * 0.1
        ms = timeit(lambda: ops.conv(ops.FORM_S1, 2, 1, [(col, 32, 0)], w1, 0, 64, bias=b, want_raw=False, want_relu=True))
        gb = (col.numel() + x.numel()) * 2 / 1e9
        print(f"vgg first conv (im2col K=32 -> 64) @256: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s")
        w3 = torch.randn(64, 3, 3, 3, device=dev) * 0.1
        img = torch.rand(F_, 3, 256, 256, device=dev)
        sh3, sc3 = torch.tensor([-.03, -.088, -.188], device=dev), torch.tensor([.458, .448, .45], device=dev)
        ms = timeit(lambda: ops.vgg_first_conv(img, w3, b, sh3, sc3))
        gb = (img.numel() * 4 + x.numel() * 2) / 1e9
        print(f"vgg first conv fused (3 -> 64) @256: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s")
        del img
        ms = timeit(lambda: ops.conv(ops.FORM_S1_DGRAD, 2, 3, [(x, 64, 0)], w3, 1, 3, f32="nchw"))
        print(f"vgg first conv dgrad 64->3 @256 (implicit GEMM, N = 16, bf16 channels-last out): {ms:.3f} ms")
        ms = timeit(lambda: ops.vgg_first_dgrad(x, w3, sc3))
        print(f"vgg first conv dgrad 64->3 @256 (taps on N + shift-add, fp32 NCHW out): {ms:.3f} ms  "
              f"{(x.numel() * 2 + F_ * 3 * 65536 * 4) / 1e9 / ms * 1e3:.0f} GB/s")
        del x, col
        x2 = torch.randn(F_, 128, 128, 64, device=dev).to(torch.bfloat16).relu_()
        w2 = torch.randn(128, 64, 3, 3, device=dev) * 0.03
        b2 = torch.zeros(128, device=dev)
        fl2 = 2.0 * F_ * 128 * 128 * 64 * 128 * 9
        ms = timeit(lambda: ops.conv(ops.FORM_S1, 2, 3, [(x2, 64, 0)], w2, 0, 128, bias=b2, want_raw=False, want_relu=True))
        print(f"vgg conv3x3 64->128 @128 fwd: {ms:.3f} ms  {fl2 / ms / 1e9:.1f} TFLOP/s")
    if which in ("s2", "all"):
        F_ = clips * T
        img = torch.randn(F_, 6, 256, 256, device=dev)
        w6 = torch.randn(64, 6, 4, 4, device=dev) * 0.1
        b = torch.zeros(64, device=dev)
        act = torch.randn(F_, 128, 128, 64, device=dev).to(torch.bfloat16)
        gb = (img.numel() * 4 + act.numel() * 2) / 1e9
        ms = timeit(lambda: ops.s2conv(img, 6, w6, b, relu=True))
        print(f"s2conv 6->64 @256 fwd (no im2col): {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s")
        ms = timeit(lambda: ops.s2conv(img, 6, w6, None, mask=act, addend=act))
        print(f"s2conv dgrad form (+mask+addend): {ms:.3f} ms  {(gb + 2 * act.numel() * 2 / 1e9) / ms * 1e3:.0f} GB/s")
        dw = torch.empty(64, 6, 4, 4, device=dev)
        ms = timeit(lambda: ops.s2wgrad(img, 6, act, dw, dbias=b))
        print(f"s2wgrad 64x6x4x4 (+bias): {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s")
        ms = timeit(lambda: ops.im2col4x4s2(img, 6))
        print(f"   (old) im2col4x4s2: {ms:.3f} ms")
        col = ops.im2col4x4s2(img, 6)
        w2 = torch.randn(64, 128, 1, 1, device=dev) * 0.1
        ms = timeit(lambda: ops.conv(ops.FORM_S1, 2, 1, [(col, 128, 0)], w2, 0, 64, bias=b, want_raw=False, want_relu=True))
        print(f"   (old) 1x1 GEMM K=128 -> 64: {ms:.3f} ms")
        dw2 = torch.empty(64, 128, 1, 1, device=dev)
        ms = timeit(lambda: ops.wgrad(ops.FORM_S1, 2, 1, (act, 64, 0), (col, 128, 0), dw2, m_axis=0, dbias=b))
        print(f"   (old) wgrad1x1 64x128: {ms:.3f} ms")
        ms = timeit(lambda: ops.conv(ops.FORM_S1_DGRAD, 2, 1, [(col, 128, 0)], w2.view(128, 64, 1, 1), 1, 64, mask=act, addend=act))
        print(f"   (old) dgrad1x1 128 -> 64 (+mask+addend): {ms:.3f} ms")
        del img, col, act
    if which in ("wgrad_small", "all"):
        F_ = clips * T
        dy = torch.randn(F_, 64, 64, 128, device=dev).to(torch.bfloat16)
        h = torch.randn(F_, 64, 64, 32, device=dev).to(torch.bfloat16)
        x64 = torch.randn(F_, 64, 64, 64, device=dev).to(torch.bfloat16)
        dw = torch.empty(128, 32, 1, 1, device=dev)
        db = torch.zeros(128, device=dev)
        ms = timeit(lambda: ops.wgrad(ops.FORM_S1, 2, 1, (dy, 128, 0), (h, 32, 0), dw, m_axis=0, dbias=db))
        gb = (dy.numel() + h.numel()) * 2 / 1e9
        print(f"wgrad1x1 128x32 @64 (+bias): {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s")
        dw = torch.empty(128, 64, 1, 1, device=dev)
        ms = timeit(lambda: ops.wgrad(ops.FORM_S1, 2, 1, (dy, 128, 0), (x64, 64, 0), dw, m_axis=0, dbias=db))
        gb = (dy.numel() + x64.numel()) * 2 / 1e9
        print(f"wgrad1x1 128x64 @64 (+bias): {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s")
        dw = torch.empty(128, 64, 3, 3, device=dev)
        fl = 2.0 * F_ * 64 * 64 * 128 * 64 * 9
        ms = timeit(lambda: ops.wgrad(ops.FORM_S1, 2, 3, (dy, 128, 0), (x64, 64, 0), dw, m_axis=0, dbias=db))
        print(f"wgrad3x3 128x64 @64 (+bias): {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
        dw = torch.empty(128, 128, 3, 3, device=dev)
        ms = timeit(lambda: ops.wgrad(ops.FORM_S1, 2, 3, (dy, 128, 0), (dy, 128, 0), dw, m_axis=0, dbias=db))
        print(f"wgrad3x3 128x128 @64 (+bias): {ms:.3f} ms  {2 * fl / ms / 1e9:.1f} TFLOP/s")
    if which in ("wgrad3d", "all"):
        x = torch.randn(clips, T, 64, 64, 128, device=dev).to(torch.bfloat16)
        dy = torch.randn(clips, T, 64, 64, 128, device=dev).to(torch.bfloat16)
        dw = torch.empty(128, 128, 3, 3, 3, device=dev)
        fl = 2.0 * clips * T * 64 * 64 * 128 * 128 * 27
        ms = timeit(lambda: ops.wgrad(ops.FORM_S1, 3, 3, (dy, 128, 0), (x, 128, 0), dw, m_axis=0))
        print(f"conv3d wgrad: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
