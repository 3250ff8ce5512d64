"""CPU stand-ins (plain torch, fp32) for the C-ABI ops the LPIPS path calls -- TEST INFRASTRUCTURE ONLY.

They let ``pytest -m "not gpu"`` drive the PRODUCT's host logic (faceoff_b200/lpips.py + graph.py: the autograd tape, the
order of the recorded backward closures, the lockstep of the two VGG trunks, the hand-over of the pool gradient to the fused
tap kernel, the fp32 NCHW input gradient) without a GPU, and compare the result with the oracle.  Nothing under
faceoff_b200/ imports this file; ``install(monkeypatch)`` swaps the functions in for the duration of one test and counts the
calls per op.  Each stand-in restates what the CUDA kernel of the same name is documented to compute (include/faceoff_b200.h).
"""
import collections

import torch
import torch.nn.functional as F

from faceoff_b200 import ops

CALLS = collections.Counter()
EPS = 1e-10


def _nchw(t, c=None):
    t = t.permute(0, 3, 1, 2)
    return t if c is None else t[:, :c]


def _nhwc(t, cs=None):
    t = t.permute(0, 2, 3, 1).contiguous()
    if cs is not None and cs > t.shape[-1]:
        t = F.pad(t, (0, cs - t.shape[-1]))
    return t


def vgg_first_conv(x, weight, bias, shift=None, scale=None):
    CALLS["vgg_first_conv"] += 1
    if shift is not None:
        x = (x - shift.view(1, 3, 1, 1)) / scale.view(1, 3, 1, 1)
    return _nhwc(F.conv2d(x, weight, bias[:64], padding=1).relu())


def vgg_first_dgrad(dy, weight, scale=None):
    CALLS["vgg_first_dgrad"] += 1
    dx = F.conv_transpose2d(_nchw(dy), weight, padding=1)
    return (dx / scale.view(1, 3, 1, 1) if scale is not None else dx).contiguous()


def _to_nc(t, ndim):
    """channels-last [N, (D,) H, W, C] -> channels-first"""
    return t.permute(0, 3, 1, 2) if ndim == 2 else t.permute(0, 4, 1, 2, 3)


def _to_cl(t, ndim, cs=None):
    t = (t.permute(0, 2, 3, 1) if ndim == 2 else t.permute(0, 2, 3, 4, 1)).contiguous()
    if cs is not None and cs > t.shape[-1]:
        t = F.pad(t, (0, cs - t.shape[-1]))
    return t


def conv(form, ndim, ksize, srcs, weight, n_axis, cout, bias=None, mask=None, addend=None, want_raw=True, want_relu=False,
         f32=None, out_cs=None, **kw):
    """One convolution of the planner's vocabulary on channels-last tensors ([N,H,W,Cs], or [B,T,H,W,Cs] for ndim 3):
    FORM_S1 / FORM_S1_DGRAD = stride 1, pad k//2; FORM_DOWN = Conv2d(4, stride 2, pad 1); FORM_UP = ConvTranspose2d(4, 2, 1).
    ``n_axis`` says which weight axis is the OUTPUT channel: 0 = a Conv weight [out, in, ...] used as a correlation, 1 = a
    weight [in, out, ...] used as its transpose (ConvTranspose / the data gradient of a Conv)."""
    CALLS["conv"] += 1
    x = _to_nc(torch.cat([t[..., off:off + c] for (t, c, off) in srcs], -1), ndim)
    w = weight.detach()
    b = None if bias is None else bias[:cout]
    fwd = F.conv2d if ndim == 2 else F.conv3d
    tr = F.conv_transpose2d if ndim == 2 else F.conv_transpose3d
    if form in (ops.FORM_S1, ops.FORM_S1_DGRAD):
        y = fwd(x, w, b, padding=ksize // 2) if n_axis == 0 else tr(x, w, b, padding=ksize // 2)
    elif form == ops.FORM_DOWN:
        assert n_axis == 0 and ndim == 2 and ksize == 4
        y = F.conv2d(x, w, b, stride=2, padding=1)
    else:
        assert form == ops.FORM_UP and n_axis == 1 and ndim == 2 and ksize == 4
        y = F.conv_transpose2d(x, w, b, stride=2, padding=1)
    assert y.shape[1] == cout, (y.shape, cout)
    if f32 == "nchw":          # fp32 NCHW result next to (or instead of) the channels-last ones; no mask / addend there
        assert mask is None and addend is None and ndim == 2
        return None, None, y.contiguous()
    y = _to_cl(y, ndim, out_cs if out_cs is not None else ops.pad16(cout))
    if mask is not None:
        y = torch.where(mask > 0, y, torch.zeros_like(y))
    if addend is not None:
        y = y + addend
    return (y if want_raw else None), (y.relu() if want_relu else None), (y if f32 == "cl" else None)


def wgrad(form, ndim, ksize, p, q, dweight, m_axis, q_w_off=0, accumulate=False, dbias=None, dbias_accumulate=False,
          q_shift_sign=1, **kw):
    """dweight[m][q_w_off + n][tap] (m_axis 0; transposed for m_axis 1) (+)= sum_pix P[pix][m] * Q[s * pix + sign * (tap - pad)][n]
    with s = 1, pad = k // 2 (FORM_S1) or s = 2, pad = 1 (FORM_DOWN); dbias (+)= column sums of P."""
    CALLS["wgrad"] += 1
    pt = _to_nc(p[0][..., p[2]:p[2] + p[1]], ndim)
    qt = _to_nc(q[0][..., q[2]:q[2] + q[1]], ndim)
    stride, pad = (1, ksize // 2) if form == ops.FORM_S1 else (2, 1)
    size = (p[1], q[1]) + (ksize,) * ndim
    fn = torch.nn.grad.conv2d_weight if ndim == 2 else torch.nn.grad.conv3d_weight
    dw = fn(qt, size, pt, stride=stride, padding=pad)          # [m][n][taps]: Q is the "input", P the "grad_output"
    if q_shift_sign < 0:
        dw = dw.flip(tuple(range(2, 2 + ndim)))
    if m_axis == 0:
        view = dweight[:, q_w_off:q_w_off + q[1]]
    else:
        view, dw = dweight[q_w_off:q_w_off + q[1], :], dw.transpose(0, 1)
    if accumulate:
        view += dw
    else:
        view.copy_(dw)
    if dbias is not None:
        s_ = pt.transpose(0, 1).reshape(p[1], -1).sum(1)
        if dbias_accumulate:
            dbias[:p[1]] += s_
        else:
            dbias[:p[1]] = s_


def colsum(x, c, out, c_off=0, accumulate=False, **kw):
    CALLS["colsum"] += 1
    s_ = x.reshape(-1, x.shape[-1])[:, c_off:c_off + c].sum(0)
    if accumulate:
        out[:c] += s_
    else:
        out[:c] = s_


def relu(x):
    CALLS["relu"] += 1
    return x.relu()


# ---- image-side stride-2 layers (csrc/small_cin.cu) and the GEMM + col2im form of the last ConvTranspose2d
def s2conv(x, c, weight, bias=None, mask=None, addend=None, relu=False):
    CALLS["s2conv"] += 1
    y = _nhwc(F.conv2d(x[:, :c], weight, bias, stride=2, padding=1))
    if mask is not None:
        y = torch.where(mask > 0, y, torch.zeros_like(y))
    if addend is not None:
        y = y + addend
    return y.relu() if relu else y


def s2wgrad(x, c, y, dweight, accumulate=False, dbias=None, dbias_accumulate=False):
    CALLS["s2wgrad"] += 1
    dw = torch.nn.grad.conv2d_weight(x[:, :c], (64, c, 4, 4), _nchw(y), stride=2, padding=1)
    if accumulate:
        dweight += dw
    else:
        dweight.copy_(dw)
    if dbias is not None:
        s_ = y.reshape(-1, 64).sum(0)
        if dbias_accumulate:
            dbias[:64] += s_
        else:
            dbias[:64] = s_


def im2col4x4s2(x, c):
    CALLS["im2col4x4s2"] += 1
    n, _, h, w = x.shape
    cols = F.unfold(x[:, :c], 4, stride=2, padding=1).view(n, c, 16, h // 2, w // 2)     # [N, ch, tap, ho, wo]
    out = torch.zeros(n, h // 2, w // 2, 16, 8)
    out[..., :c] = cols.permute(0, 3, 4, 2, 1)
    return out.view(n, h // 2, w // 2, 128)


def col2im4x4s2(col, bias, c):
    CALLS["col2im4x4s2"] += 1
    n, hi, wi, _ = col.shape
    wid = torch.zeros(16, 8, c, 16)
    for tap in range(16):
        for co in range(c):
            wid[tap, co, co, tap] = 1.0
    out = F.conv_transpose2d(_nchw(col), wid.view(128, c, 4, 4), stride=2, padding=1)
    return (out + bias.view(1, c, 1, 1) if bias is not None else out).contiguous()


def chansum_nchw(x, c, out, accumulate=False):
    CALLS["chansum_nchw"] += 1
    s_ = x[:, :c].sum((0, 2, 3))
    if accumulate:
        out[:c] += s_
    else:
        out[:c] = s_


# ---- vector quantiser (csrc/vq.cu)
def vq_prep(embed):
    CALLS["vq_prep"] += 1
    e_t = embed.t().contiguous()
    return None, e_t, torch.cat([e_t.pow(2).sum(1), e_t.pow(2).sum(1).max().view(1)])


def vq_assign(x, e_t, e_split, e_norm2, n_flagged=None):
    CALLS["vq_assign"] += 1
    d = x.double().pow(2).sum(1, keepdim=True) - 2 * x.double() @ e_t.double().t() + e_t.double().pow(2).sum(1)
    return d.argmin(1)


def vq_gather_stats(x, ind, e_t, diff_sum, counts, embed_sum, want_f32=True, want_bf16=False):
    CALLS["vq_gather_stats"] += 1
    q = e_t[ind]
    diff_sum += (q - x).pow(2).sum()
    if counts is not None:
        counts += torch.bincount(ind, minlength=e_t.shape[0]).to(counts.dtype)
        embed_sum += x.t() @ F.one_hot(ind, e_t.shape[0]).to(x.dtype)
    st = x + (q - x)
    return (st if want_f32 else None), (st.clone() if want_bf16 else None)


def vq_ema(embed, cluster_size, embed_avg, counts, embed_sum, decay, eps):
    CALLS["vq_ema"] += 1
    cluster_size.mul_(decay).add_(counts, alpha=1 - decay)
    embed_avg.mul_(decay).add_(embed_sum, alpha=1 - decay)
    n = cluster_size.sum()
    cs = (cluster_size + eps) / (n + embed.shape[1] * eps) * n
    embed.copy_(embed_avg / cs.unsqueeze(0))


def vq_backward(g_q, g_c_off, g_diff, x, ind, e_t, want_f32=True, want_bf16=False):
    """d/dx of the straight-through output (identity) + g_diff * d/dx mean((q - x)^2)."""
    CALLS["vq_backward"] += 1
    g = torch.zeros_like(x)
    if g_q is not None:
        g = g + g_q[:, g_c_off:g_c_off + x.shape[1]]
    if g_diff is not None:
        g = g + g_diff * 2.0 * (x - e_t[ind]) / x.numel()
    return (g if want_f32 else None), (g.clone() if want_bf16 else None)


def maxpool2(x):
    CALLS["maxpool2"] += 1
    return _nhwc(F.max_pool2d(_nchw(x), 2, 2))


def maxpool2_bwd(x, y, dy):
    """Gradient to the FIRST maximum of each window (row-major), only where x > 0 (x is post-ReLU)."""
    CALLS["maxpool2_bwd"] += 1
    n, h, w, c = x.shape
    win = x.view(n, h // 2, 2, w // 2, 2, c).permute(0, 1, 3, 5, 2, 4).reshape(n, h // 2, w // 2, c, 4)
    first = win.argmax(-1)    # torch.argmax returns the first maximal index on CPU
    onehot = F.one_hot(first, 4).to(x.dtype) * dy.unsqueeze(-1)
    dx = onehot.view(n, h // 2, w // 2, c, 2, 2).permute(0, 1, 4, 2, 5, 3).reshape(n, h, w, c)
    return torch.where(x > 0, dx, torch.zeros_like(dx))


def _tap_value(f0, f1, w):
    a = f0 / (f0.pow(2).sum(-1, keepdim=True).sqrt() + EPS)
    b = f1 / (f1.pow(2).sum(-1, keepdim=True).sqrt() + EPS)
    return ((a - b).pow(2) * w).sum(-1).mean((1, 2))


def lpips_tap(f0, f1, w, out):
    CALLS["lpips_tap"] += 1
    out += _tap_value(f0, f1, w)


def lpips_tap_pool(f0, f1, w, out, pool_f1=False):
    CALLS["lpips_tap_pool"] += 1
    out += _tap_value(f0, f1, w)
    return _nhwc(F.max_pool2d(_nchw(f0), 2, 2)), (_nhwc(F.max_pool2d(_nchw(f1), 2, 2)) if pool_f1 else None)


def lpips_tap_bwd(f0, f1, w, g, addend=None):
    CALLS["lpips_tap_bwd"] += 1
    with torch.enable_grad():
        a = f0.detach().clone().requires_grad_(True)
        (_tap_value(a, f1, w) * g).sum().backward()
    d = torch.nan_to_num(a.grad)
    if addend is not None:
        d = d + addend
    return torch.where(f0 > 0, d, torch.zeros_like(d))


def lpips_tap_bwd_pool(f0, f1, w, g, pool_dy):
    CALLS["lpips_tap_bwd_pool"] += 1
    n = CALLS["lpips_tap_bwd"], CALLS["maxpool2_bwd"], CALLS["maxpool2"]
    d = lpips_tap_bwd(f0, f1, w, g, maxpool2_bwd(f0, maxpool2(f0), pool_dy))
    CALLS["lpips_tap_bwd"], CALLS["maxpool2_bwd"], CALLS["maxpool2"] = n
    return d


def add_grads(a, b):
    CALLS["add_grads"] += 1
    return a + b


def pack_nchw(x, cs=None, shift=None, scale=None):
    CALLS["pack_nchw"] += 1
    if shift is not None:
        x = (x - shift.view(1, -1, 1, 1)) / scale.view(1, -1, 1, 1)
    return _nhwc(x, cs if cs is not None else ops.pad16(x.shape[1]))


def unpack_nchw(x, c):
    CALLS["unpack_nchw"] += 1
    return _nchw(x, c).contiguous()


_ALL = ("s2conv", "s2wgrad", "im2col4x4s2", "col2im4x4s2", "chansum_nchw", "vq_prep", "vq_assign", "vq_gather_stats", "vq_ema",
        "vq_backward", "wgrad", "colsum", "relu", "vgg_first_conv", "vgg_first_dgrad", "conv", "maxpool2", "maxpool2_bwd", "lpips_tap", "lpips_tap_pool", "lpips_tap_bwd",
        "lpips_tap_bwd_pool", "add_grads", "pack_nchw", "unpack_nchw")


def install(monkeypatch):
    CALLS.clear()
    for name in _ALL:
        monkeypatch.setattr(ops, name, globals()[name])


def install_for_process():
    """The same without a pytest fixture: for worker processes that live only for one check (tests/gpu_dp_check.py --cpu)."""
    CALLS.clear()
    for name in _ALL:
        setattr(ops, name, globals()[name])
