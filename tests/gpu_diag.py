"""GPU diagnostics for the individual kernels (run on the B200 box).

    python tests/gpu_diag.py            # run every case, each in its own subprocess with a timeout
    python tests/gpu_diag.py CASE       # run one case in-process

Each case compares a C-ABI call against plain torch fp32 ops on bf16-rounded operands (so the only
differences are accumulation order / output rounding) and prints max abs / relative errors.
"""
import os
import subprocess
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

dev = "cuda"


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def to_cl(x_nchw, cs=None):
    """fp32 NCHW (or NCDHW) -> channels-last bf16 with padded channels (torch ops; test-side only)."""
    c = x_nchw.shape[1]
    cs = cs or (c + 15) // 16 * 16
    if x_nchw.dim() == 4:
        t = x_nchw.permute(0, 2, 3, 1)
    else:
        t = x_nchw.permute(0, 2, 3, 4, 1)
    out = torch.zeros((*t.shape[:-1], cs), dtype=torch.bfloat16, device=x_nchw.device)
    out[..., :c] = t.to(torch.bfloat16)
    return out.contiguous()


def from_cl(t, c):
    t = t[..., :c].to(torch.float32)
    if t.dim() == 4:
        return t.permute(0, 3, 1, 2).contiguous()
    return t.permute(0, 4, 1, 2, 3).contiguous()


def report(name, got, ref, tol=2e-2):
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-12
    ok = err <= tol * scale
    print(f"  [{'OK' if ok else 'FAIL'}] {name}: max|err| {err:.4e}  ref max {scale:.4e}  rel {err / scale:.3e}", flush=True)
    return ok


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev)


# ------------------------------------------------------------------------------------------------
def case_conv_s1():
    from faceoff_b200 import ops
    ok = True
    cfgs = [
        # (n, h, w, cin, cout, k)
        (2, 16, 16, 64, 128, 1),
        (2, 16, 16, 128, 128, 3),
        (3, 32, 32, 128, 32, 3),
        (2, 8, 8, 32, 128, 1),
        (2, 64, 64, 64, 128, 3),
        (1, 32, 32, 6, 64, 3),
        (2, 16, 16, 128, 512, 3),
        (2, 16, 16, 512, 512, 3),
        (1, 64, 64, 3, 64, 3),
        (2, 24, 40, 64, 64, 3),
    ]
    for (n, h, w, cin, cout, k) in cfgs:
        x = bf(rnd(n, cin, h, w, seed=1))
        wt = bf(rnd(cout, cin, k, k, seed=2, scale=(cin * k * k) ** -0.5))
        b = rnd(cout, seed=3)
        ref = F.conv2d(x, wt, b, padding=(k - 1) // 2)
        bp = torch.zeros(ops.pad16(cout), device=dev)
        bp[:cout] = b
        raw, relu, _ = ops.conv(ops.FORM_S1, 2, k, [(to_cl(x), cin, 0)], wt.contiguous(), 0, cout, bias=bp,
                                want_raw=True, want_relu=True)
        torch.cuda.synchronize()
        ok &= report(f"s1 n{n} {h}x{w} {cin}->{cout} k{k} raw", from_cl(raw, cout), ref)
        ok &= report(f"s1 n{n} {h}x{w} {cin}->{cout} k{k} relu", from_cl(relu, cout), ref.relu())
    return ok


def case_conv_s1_epilogue():
    from faceoff_b200 import ops
    ok = True
    n, h, w, cin, cout, k = 2, 16, 16, 64, 128, 3
    x = bf(rnd(n, cin, h, w, seed=1))
    wt = bf(rnd(cout, cin, k, k, seed=2, scale=(cin * k * k) ** -0.5))
    b = rnd(cout, seed=3)
    m = bf(rnd(n, cout, h, w, seed=4))
    a = bf(rnd(n, cout, h, w, seed=5))
    ref = F.conv2d(x, wt, b, padding=1) * (m > 0) + a
    raw, relu, f32 = ops.conv(ops.FORM_S1, 2, k, [(to_cl(x), cin, 0)], wt.contiguous(), 0, cout, bias=b.contiguous(),
                              mask=to_cl(m), addend=to_cl(a), want_raw=True, want_relu=True, f32="cl")
    torch.cuda.synchronize()
    ok &= report("mask+addend raw", from_cl(raw, cout), ref)
    ok &= report("mask+addend relu", from_cl(relu, cout), ref.relu())
    ok &= report("mask+addend f32", f32.permute(0, 3, 1, 2), ref, tol=2e-3)
    # NCHW fp32 output with cout=6
    cout = 6
    wt = bf(rnd(cout, cin, k, k, seed=6, scale=(cin * k * k) ** -0.5))
    b = torch.zeros(16, device=dev)
    b[:cout] = rnd(cout, seed=7)
    ref = F.conv2d(x, wt, b[:cout], padding=1)
    _, _, f32 = ops.conv(ops.FORM_S1, 2, k, [(to_cl(x), cin, 0)], wt.contiguous(), 0, cout, bias=b, want_raw=False,
                         f32="nchw")
    torch.cuda.synchronize()
    ok &= report("nchw f32 cout=6", f32, ref, tol=2e-3)
    # concat of two sources (64 + 128 -> 64, 1x1) and a channel-offset slice
    x0, x1 = bf(rnd(n, 64, h, w, seed=8)), bf(rnd(n, 128, h, w, seed=9))
    wt = bf(rnd(64, 192, 1, 1, seed=10, scale=192 ** -0.5))
    ref = F.conv2d(torch.cat([x0, x1], 1), wt)
    raw, _, _ = ops.conv(ops.FORM_S1, 2, 1, [(to_cl(x0), 64, 0), (to_cl(x1), 128, 0)], wt.contiguous(), 0, 64)
    torch.cuda.synchronize()
    ok &= report("cat 64+128 -> 64 k1", from_cl(raw, 64), ref)
    wide = to_cl(torch.cat([x0, x1], 1))  # 192 storage channels; read slice [64, 192)
    wt2 = bf(rnd(32, 128, 3, 3, seed=11, scale=(128 * 9) ** -0.5))
    ref = F.conv2d(x1, wt2, padding=1)
    raw, _, _ = ops.conv(ops.FORM_S1, 2, 3, [(wide, 128, 64)], wt2.contiguous(), 0, 32)
    torch.cuda.synchronize()
    ok &= report("slice c_off=64 of 192", from_cl(raw, 32), ref)
    return ok


def case_conv3d():
    from faceoff_b200 import ops
    ok = True
    for (n, d, h, w, c) in [(1, 4, 16, 16, 128), (2, 3, 8, 8, 128), (1, 6, 32, 32, 128)]:
        x = bf(rnd(n, c, d, h, w, seed=1))
        wt = bf(rnd(c, c, 3, 3, 3, seed=2, scale=(c * 27) ** -0.5))
        b = rnd(c, seed=3)
        ref = F.conv3d(x, wt, b, padding=1)
        raw, relu, _ = ops.conv(ops.FORM_S1, 3, 3, [(to_cl(x), c, 0)], wt.contiguous(), 0, c, bias=b.contiguous(),
                                want_relu=True)
        torch.cuda.synchronize()
        ok &= report(f"conv3d n{n} d{d} {h}x{w}", from_cl(raw, c), ref)
        ok &= report(f"conv3d relu", from_cl(relu, c), ref.relu())
        # data gradient
        gy = bf(rnd(n, c, d, h, w, seed=4))
        xr = x.clone().requires_grad_(True)
        F.conv3d(xr, wt, None, padding=1).backward(gy)
        dx, _, _ = ops.conv(ops.FORM_S1_DGRAD, 3, 3, [(to_cl(gy), c, 0)], wt.contiguous(), 1, c)
        torch.cuda.synchronize()
        ok &= report(f"conv3d dgrad", from_cl(dx, c), xr.grad)
    return ok


def case_conv_strided():
    from faceoff_b200 import ops
    ok = True
    for (n, h, w, cin, cout) in [(2, 32, 32, 64, 128), (2, 64, 64, 6, 64), (2, 16, 16, 128, 64), (1, 128, 128, 6, 64)]:
        x = bf(rnd(n, cin, h, w, seed=1))
        wt = bf(rnd(cout, cin, 4, 4, seed=2, scale=(cin * 16) ** -0.5))
        b = torch.zeros(ops.pad16(cout), device=dev)
        b[:cout] = rnd(cout, seed=3)
        ref = F.conv2d(x, wt, b[:cout], stride=2, padding=1)
        raw, relu, _ = ops.conv(ops.FORM_DOWN, 2, 4, [(to_cl(x), cin, 0)], wt.contiguous(), 0, cout, bias=b,
                                want_relu=True)
        torch.cuda.synchronize()
        ok &= report(f"down n{n} {h}x{w} {cin}->{cout}", from_cl(raw, cout), ref)
        # its data gradient: UP form with the same weight, n_axis=1
        gy = bf(rnd(*ref.shape, seed=4))
        xr = x.clone().requires_grad_(True)
        F.conv2d(xr, wt, None, stride=2, padding=1).backward(gy)
        dx, _, _ = ops.conv(ops.FORM_UP, 2, 4, [(to_cl(gy), cout, 0)], wt.contiguous(), 1, cin)
        torch.cuda.synchronize()
        ok &= report(f"down-dgrad (UP form) -> {cin}", from_cl(dx, cin), xr.grad)
    for (n, h, w, cin, cout) in [(2, 16, 16, 128, 64), (2, 32, 32, 64, 6), (2, 8, 8, 64, 64)]:
        x = bf(rnd(n, cin, h, w, seed=1))
        wt = bf(rnd(cin, cout, 4, 4, seed=2, scale=(cin * 4) ** -0.5))
        b = torch.zeros(ops.pad16(cout), device=dev)
        b[:cout] = rnd(cout, seed=3)
        ref = F.conv_transpose2d(x, wt, b[:cout], stride=2, padding=1)
        raw, _, _ = ops.conv(ops.FORM_UP, 2, 4, [(to_cl(x), cin, 0)], wt.contiguous(), 1, cout, bias=b)
        torch.cuda.synchronize()
        ok &= report(f"up n{n} {h}x{w} {cin}->{cout}", from_cl(raw, cout), ref)
        _, _, f32 = ops.conv(ops.FORM_UP, 2, 4, [(to_cl(x), cin, 0)], wt.contiguous(), 1, cout, bias=b, want_raw=False,
                             f32="nchw")
        torch.cuda.synchronize()
        ok &= report(f"up nchw f32", f32, ref, tol=2e-3)
        gy = bf(rnd(*ref.shape, seed=4))
        xr = x.clone().requires_grad_(True)
        F.conv_transpose2d(xr, wt, None, stride=2, padding=1).backward(gy)
        dx, _, _ = ops.conv(ops.FORM_DOWN, 2, 4, [(to_cl(gy), cout, 0)], wt.contiguous(), 0, cin)
        torch.cuda.synchronize()
        ok &= report(f"up-dgrad (DOWN form) -> {cin}", from_cl(dx, cin), xr.grad)
    return ok


def case_dgrad_s1():
    from faceoff_b200 import ops
    ok = True
    for (n, h, w, cin, cout, k) in [(2, 16, 16, 128, 32, 3), (2, 16, 16, 32, 128, 1), (2, 32, 32, 64, 128, 3),
                                    (2, 16, 16, 192, 64, 1)]:
        x = bf(rnd(n, cin, h, w, seed=1))
        wt = bf(rnd(cout, cin, k, k, seed=2, scale=(cin * k * k) ** -0.5))
        gy = bf(rnd(n, cout, h, w, seed=4))
        xr = x.clone().requires_grad_(True)
        F.conv2d(xr, wt, None, padding=(k - 1) // 2).backward(gy)
        dx, _, _ = ops.conv(ops.FORM_S1_DGRAD, 2, k, [(to_cl(gy), cout, 0)], wt.contiguous(), 1, cin)
        torch.cuda.synchronize()
        ok &= report(f"dgrad s1 {cin}<-{cout} k{k}", from_cl(dx, cin), xr.grad)
    return ok


def case_wgrad():
    from faceoff_b200 import ops
    ok = True
    for (n, h, w, cin, cout, k) in [(2, 16, 16, 128, 128, 3), (4, 16, 16, 128, 32, 3), (2, 16, 16, 32, 128, 1),
                                    (2, 32, 32, 64, 128, 3), (2, 32, 32, 6, 64, 3), (3, 8, 8, 128, 64, 1)]:
        x = bf(rnd(n, cin, h, w, seed=1))
        wt = bf(rnd(cout, cin, k, k, seed=2)).requires_grad_(True)
        gy = bf(rnd(n, cout, h, w, seed=4))
        F.conv2d(x, wt, None, padding=(k - 1) // 2).backward(gy)
        dw = torch.full_like(wt, 7.0).detach()
        db = torch.full((cout,), 3.0, device=dev)
        ops.wgrad(ops.FORM_S1, 2, k, (to_cl(gy), cout, 0), (to_cl(x), cin, 0), dw, m_axis=0, dbias=db)
        torch.cuda.synchronize()
        ok &= report(f"wgrad s1 {cin}->{cout} k{k} n{n} {h}x{w}", dw, wt.grad)
        ok &= report(f"   fused bias grad", db, gy.sum((0, 2, 3)), tol=1e-4)
    # swapped roles (P = x, Q = dy read at pix - tap): narrow-output layers
    for (n, h, w, cin, cout, k) in [(4, 16, 16, 128, 32, 3), (2, 32, 32, 128, 32, 3), (2, 8, 8, 128, 32, 3)]:
        x = bf(rnd(n, cin, h, w, seed=1))
        wt = bf(rnd(cout, cin, k, k, seed=2)).requires_grad_(True)
        gy = bf(rnd(n, cout, h, w, seed=4))
        F.conv2d(x, wt, None, padding=1).backward(gy)
        dw = torch.zeros_like(wt).detach()
        ops.wgrad(ops.FORM_S1, 2, k, (to_cl(x), cin, 0), (to_cl(gy), cout, 0), dw, m_axis=1, q_shift_sign=-1)
        torch.cuda.synchronize()
        ok &= report(f"wgrad s1 swapped {cin}->{cout} n{n} {h}x{w}", dw, wt.grad)
    # conv3d
    n, d, h, w, c = 1, 4, 16, 16, 128
    x = bf(rnd(n, c, d, h, w, seed=1))
    wt = bf(rnd(c, c, 3, 3, 3, seed=2)).requires_grad_(True)
    gy = bf(rnd(n, c, d, h, w, seed=4))
    F.conv3d(x, wt, None, padding=1).backward(gy)
    dw = torch.zeros_like(wt).detach()
    ops.wgrad(ops.FORM_S1, 3, 3, (to_cl(gy), c, 0), (to_cl(x), c, 0), dw, m_axis=0)
    torch.cuda.synchronize()
    ok &= report("wgrad conv3d", dw, wt.grad)
    # 4x4 stride-2 conv: P = dy (low res), Q = x (hi res)
    for (n, h, w, cin, cout) in [(2, 32, 32, 64, 128), (2, 64, 64, 6, 64)]:
        x = bf(rnd(n, cin, h, w, seed=1))
        wt = bf(rnd(cout, cin, 4, 4, seed=2)).requires_grad_(True)
        y = F.conv2d(x, wt, None, stride=2, padding=1)
        gy = bf(rnd(*y.shape, seed=4))
        y.backward(gy)
        dw = torch.zeros_like(wt).detach()
        db = torch.ones(cout, device=dev)
        ops.wgrad(ops.FORM_DOWN, 2, 4, (to_cl(gy), cout, 0), (to_cl(x), cin, 0), dw, m_axis=0, dbias=db,
                  dbias_accumulate=True)
        torch.cuda.synchronize()
        ok &= report(f"wgrad down {cin}->{cout}", dw, wt.grad)
        ok &= report(f"   fused bias grad (accumulate)", db, 1 + gy.sum((0, 2, 3)), tol=1e-4)
    # transposed conv: P = x (low res), Q = dy (hi res), weight [cin, cout, 4, 4]
    for (n, h, w, cin, cout) in [(2, 16, 16, 128, 64), (2, 32, 32, 64, 6)]:
        x = bf(rnd(n, cin, h, w, seed=1))
        wt = bf(rnd(cin, cout, 4, 4, seed=2)).requires_grad_(True)
        y = F.conv_transpose2d(x, wt, None, stride=2, padding=1)
        gy = bf(rnd(*y.shape, seed=4))
        y.backward(gy)
        dw = torch.zeros_like(wt).detach()
        ops.wgrad(ops.FORM_DOWN, 2, 4, (to_cl(x), cin, 0), (to_cl(gy), cout, 0), dw, m_axis=0)
        torch.cuda.synchronize()
        ok &= report(f"wgrad up {cin}->{cout}", dw, wt.grad)
    return ok


def case_elementwise():
    from faceoff_b200 import ops
    ok = True
    x = rnd(3, 6, 32, 32, seed=1)
    p = ops.pack_nchw(x)
    ok &= report("pack_nchw", from_cl(p, 6), bf(x), tol=1e-6)
    ok &= bool((p[..., 6:] == 0).all())
    shift, scale = torch.tensor([-.03, -.088, -.188], device=dev), torch.tensor([.458, .448, .45], device=dev)
    x3 = rnd(2, 3, 32, 32, seed=2)
    p = ops.pack_nchw(x3, shift=shift, scale=scale)
    ok &= report("pack_nchw scaling", from_cl(p, 3), bf((x3 - shift.view(1, 3, 1, 1)) / scale.view(1, 3, 1, 1)), tol=1e-2)
    u = ops.unpack_nchw(to_cl(x), 6)
    ok &= report("unpack_nchw", u, bf(x), tol=1e-6)
    xx = to_cl(rnd(2, 64, 16, 16, seed=3))
    ok &= report("relu", ops.relu(xx).float(), xx.float().relu(), tol=1e-6)
    out = torch.zeros(64, device=dev)
    ops.colsum(xx, 64, out)
    ok &= report("colsum", out, xx.float().reshape(-1, 64).sum(0), tol=1e-4)
    wide = to_cl(rnd(2, 192, 16, 16, seed=4))
    out = torch.ones(128, device=dev)
    ops.colsum(wide, 128, out, c_off=64, accumulate=True)
    ok &= report("colsum slice+acc", out, 1 + wide.float().reshape(-1, 192)[:, 64:].sum(0), tol=1e-4)
    xm = to_cl(rnd(2, 64, 16, 16, seed=5)).relu()
    y = ops.maxpool2(xm)
    ref = F.max_pool2d(from_cl(xm, 64), 2)
    ok &= report("maxpool2", from_cl(y, 64), ref, tol=1e-6)
    dy = to_cl(rnd(2, 64, 8, 8, seed=6))
    xr = from_cl(xm, 64).requires_grad_(True)
    F.max_pool2d(xr, 2).backward(from_cl(dy, 64))
    dx = ops.maxpool2_bwd(xm, y, dy)
    ok &= report("maxpool2_bwd", from_cl(dx, 64), xr.grad * (xr.detach() > 0), tol=1e-6)
    torch.cuda.synchronize()
    return ok


def case_vq():
    from faceoff_b200 import ops
    from oracle import faceoff_oracle as O
    ok = True
    for (rows, dim, K) in [(4096, 64, 512), (1000, 64, 512), (2048, 128, 1024), (8192, 64, 2048), (300, 128, 512)]:
        x = rnd(rows, dim, seed=1)
        e = rnd(dim, K, seed=2)
        e_split, e_t, e_n2 = ops.vq_prep(e)
        nf = torch.zeros(1, dtype=torch.int32, device=dev)
        ind = ops.vq_assign(x, e_t, e_split, e_n2, nf)
        torch.cuda.synchronize()
        ref_ind, dist = O.quantize_assign(x.double().cpu(), e.double().cpu())
        mism = (ind.cpu() != ref_ind).sum().item()
        d2 = dist.sort(1).values
        gap = ((d2[:, 1] - d2[:, 0]) / d2[:, 0].abs()).min().item()
        good = mism == 0
        ok &= good
        print(f"  [{'OK' if good else 'FAIL'}] vq_assign rows {rows} dim {dim} K {K}: mismatches vs fp64 argmin {mism}, "
              f"flagged {nf.item()}, min rel gap {gap:.2e}", flush=True)
        # gather / stats / loss
        diff = torch.zeros(1, device=dev)
        counts = torch.zeros(K, device=dev)
        esum = torch.zeros(dim, K, device=dev)
        q32, q16 = ops.vq_gather_stats(x, ind, e_t, diff, counts, esum, want_f32=True, want_bf16=True)
        torch.cuda.synchronize()
        xc, ec = x.cpu(), e.cpu()
        qr = F.embedding(ref_ind, ec.t())
        ok &= report("  gather q", q32.cpu(), xc + (qr - xc), tol=1e-6)
        ok &= report("  diff", diff.cpu() / (rows * dim), (qr - xc).pow(2).mean().reshape(1), tol=1e-5)
        c_ref, s_ref = O.quantize_stats(xc, ref_ind, K)
        ok &= report("  counts", counts.cpu(), c_ref, tol=1e-6)
        ok &= report("  embed_sum", esum.cpu(), s_ref, tol=1e-5)
        cs0 = torch.rand(K, device=dev) * 3
        ea0 = e.clone()
        emb, cs, ea = e.clone(), cs0.clone(), ea0.clone()
        ops.vq_ema(emb, cs, ea, counts, esum, 0.99, 1e-5)
        r_e, r_cs, r_ea = O.quantize_ema(ec, cs0.cpu(), ea0.cpu(), c_ref, s_ref)
        ok &= report("  ema embed", emb.cpu(), r_e, tol=1e-5)
        ok &= report("  ema cluster_size", cs.cpu(), r_cs, tol=1e-6)
        gq = rnd(rows, dim, seed=5)
        gd = torch.tensor([3.0], device=dev)
        g32, _ = ops.vq_backward(gq, 0, gd, x, ind, e_t)
        ok &= report("  backward", g32.cpu(), gq.cpu() + 3.0 * 2 * (xc - qr) / (rows * dim), tol=1e-5)
    return ok


def case_lpips():
    from faceoff_b200 import ops
    from oracle import faceoff_oracle as O
    ok = True
    for (n, h, c) in [(2, 32, 64), (2, 16, 128), (3, 16, 256), (2, 8, 512)]:
        f0 = bf(rnd(n, c, h, h, seed=1)).relu()
        f1 = bf(rnd(n, c, h, h, seed=2)).relu()
        w = rnd(c, seed=3).abs()
        f0r = f0.clone().requires_grad_(True)
        d = (O.normalize_tensor(f0r) - O.normalize_tensor(f1)) ** 2
        val = (d * w.view(1, c, 1, 1)).sum(1, keepdim=True).mean([2, 3]).flatten()
        out = torch.zeros(n, device=dev)
        ops.lpips_tap(to_cl(f0), to_cl(f1), w, out)
        ok &= report(f"lpips_tap c{c}", out, val.detach(), tol=1e-4)
        g = rnd(n, seed=4)
        (val * g).sum().backward()
        dd = ops.lpips_tap_bwd(to_cl(f0), to_cl(f1), w, g)
        ok &= report(f"lpips_tap_bwd c{c}", from_cl(dd, c), f0r.grad * (f0 > 0), tol=2e-2)
    torch.cuda.synchronize()
    return ok


def case_lpips_trunk():
    import warnings
    from faceoff_b200 import ops
    from faceoff_b200.graph import Tape
    from faceoff_b200.lpips import LPIPS
    from oracle import faceoff_oracle as O
    ok = True
    g = torch.load(os.path.join(ROOT, "tests", "golden", "golden.pt"), map_location="cpu")["lpips_3x64"]
    lp = O.init_lpips_params(seed=1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = LPIPS()
    m.load_state_dict(lp)
    m = m.cuda().eval()
    for name in ("a", "b"):
        x = g[name]
        s = (x - lp["scaling_layer.shift"]) / lp["scaling_layer.scale"]
        ref = O.vgg_taps(lp, s)
        tape = Tape({k: v for k, v in m.named_parameters()}, need_grad=False)
        _, taps = m._trunk(tape, x.cuda())
        torch.cuda.synchronize()
        for k, (t, r) in enumerate(zip(taps, ref)):
            ok &= report(f"trunk {name} tap{k} {tuple(r.shape)}", from_cl(t.act, r.shape[1]).cpu(), r, tol=3e-2)
    a, b = g["a"].cuda(), g["b"].cuda()
    val = m(a, b)
    ok &= report("lpips value", val.cpu().flatten(), g["val"].flatten(), tol=3e-2)
    # per-tap head on the oracle features (isolates fo_lpips_tap from the trunk)
    sa = (g["a"] - lp["scaling_layer.shift"]) / lp["scaling_layer.scale"]
    sb = (g["b"] - lp["scaling_layer.shift"]) / lp["scaling_layer.scale"]
    ta, tb = O.vgg_taps(lp, sa), O.vgg_taps(lp, sb)
    for k in range(5):
        d = (O.normalize_tensor(ta[k]) - O.normalize_tensor(tb[k])) ** 2
        r = F.conv2d(d, lp[f"lin{k}.model.1.weight"]).mean([2, 3]).flatten()
        out = torch.zeros(3, device=dev)
        ops.lpips_tap(to_cl(ta[k].cuda()), to_cl(tb[k].cuda()), lp[f"lin{k}.model.1.weight"].flatten().cuda().contiguous(), out)
        ok &= report(f"head tap{k}", out.cpu(), r, tol=2e-2)
    return ok


class _RoundBF16(torch.autograd.Function):
    """bf16 storage emulation: round values in forward and gradients in backward."""

    @staticmethod
    def forward(ctx, x):
        return bf(x)

    @staticmethod
    def backward(ctx, g):
        return bf(g)


def case_lpips_grad():
    """How much of the LPIPS input-gradient error is inherent to bf16 activation/gradient storage?"""
    import warnings
    from faceoff_b200.lpips import LPIPS
    from oracle import faceoff_oracle as O
    g = torch.load(os.path.join(ROOT, "tests", "golden", "golden.pt"), map_location="cpu")["lpips_3x64"]
    lp = {k: v.cuda() for k, v in O.init_lpips_params(seed=1).items()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = LPIPS()
    m.load_state_dict(lp)
    m = m.cuda().eval()
    a = g["a"].cuda()
    b = g["b"].cuda().requires_grad_(True)
    m(a, b).mean().backward()
    ours = b.grad.clone()

    def emu_taps(x):
        taps, ci, idx = [], 0, 0
        x = _RoundBF16.apply(x)
        for v in O.VGG_CFG:
            if v == "M":
                x = F.max_pool2d(x, 2, 2)
                idx += 1
            else:
                k = O._vgg_key(O.VGG_CONV_IDX[ci])
                x = _RoundBF16.apply(F.relu(F.conv2d(x, bf(lp[k + ".weight"]), lp[k + ".bias"], padding=1)))
                ci += 1
                idx += 2
            if idx in O.VGG_SLICE_ENDS:
                taps.append(x)
        return taps

    def lp_val(taps_fn, x0, x1):
        s0 = (x0 - lp["scaling_layer.shift"]) / lp["scaling_layer.scale"]
        s1 = (x1 - lp["scaling_layer.shift"]) / lp["scaling_layer.scale"]
        t0, t1 = taps_fn(s0), taps_fn(s1)
        val = 0
        for k in range(5):
            d = (O.normalize_tensor(t0[k]) - O.normalize_tensor(t1[k])) ** 2
            val = val + F.conv2d(d, lp[f"lin{k}.model.1.weight"]).mean([2, 3], keepdim=True)
        return val

    b2 = g["b"].cuda().requires_grad_(True)
    lp_val(lambda s: O.vgg_taps(lp, s), a, b2).mean().backward()
    ref = b2.grad.clone()
    b3 = g["b"].cuda().requires_grad_(True)
    lp_val(emu_taps, a, b3).mean().backward()
    emu = b3.grad.clone()

    def stats(name, x, r):
        nw = ((x - r).norm() / r.norm()).item()
        mx = ((x - r).abs().max() / r.abs().max()).item()
        cos = F.cosine_similarity(x.flatten(), r.flatten(), dim=0).item()
        print(f"  {name}: normwise rel err {nw:.3e}  max-normalised {mx:.3e}  cosine {cos:.6f}", flush=True)
        return nw

    e1 = stats("ours vs fp32 reference", ours, ref)
    e2 = stats("bf16-storage emulation (torch) vs fp32 reference", emu, ref)
    stats("golden (CPU reference) vs GPU fp32 reference", g["grad_b"].cuda(), ref)
    return e1 < max(2.0 * e2, 0.05)


CASES = {k[5:]: v for k, v in list(globals().items()) if k.startswith("case_")}


def main():
    if len(sys.argv) > 1:
        name = sys.argv[1]
        torch.manual_seed(0)
        t0 = time.time()
        ok = CASES[name]()
        torch.cuda.synchronize()
        print(f"== {name}: {'PASS' if ok else 'FAIL'} ({time.time() - t0:.1f}s)", flush=True)
        sys.exit(0 if ok else 1)
    bad = []
    for name in CASES:
        print(f"=== case {name}", flush=True)
        try:
            r = subprocess.run([sys.executable, __file__, name], timeout=300)
            if r.returncode != 0:
                bad.append(name)
        except subprocess.TimeoutExpired:
            print(f"== {name}: TIMEOUT", flush=True)
            bad.append(name)
    print("FAILED CASES:", bad, flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
