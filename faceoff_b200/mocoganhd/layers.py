"""Layer modules of the discriminators: nn.Conv2d / nn.Conv3d / nn.InstanceNorm2d / nn.InstanceNorm3d / nn.LeakyReLU /
nn.AvgPool2d / nn.AvgPool3d subclasses (identical parameters, buffers and state-dict keys) whose forward / backward run
in csrc/disc.cu through the C ABI (fo_dconv_*, fo_instnorm_*, fo_lrelu*, fo_avgpool3*)."""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from .. import _lib as L
from ..ops import _count, _stream


def _req(t: torch.Tensor, what: str):
    if not (t.is_cuda and t.dtype == torch.float32):
        raise L.FaceoffB200Error(f"{what}: fp32 CUDA tensor required (faceoff_b200 has no CPU path), got {t.dtype} on {t.device}")


def _triple(v, n):
    v = tuple(v) if isinstance(v, (tuple, list)) else (v,) * n
    return (1,) * (3 - n) + v if n < 3 else v


def _desc(x_shape, w_shape, stride, padding):
    """DConvDesc for x [N, Cin, (D,) H, W] and w [Cout, Cin, (kD,) kH, kW]."""
    nd = len(x_shape) - 2
    d = L.DConvDesc()
    sp = (1,) * (3 - nd) + tuple(x_shape[2:])
    k = (1,) * (3 - nd) + tuple(w_shape[2:])
    s = _triple(stride, nd)
    p = (0,) * (3 - nd) + tuple(padding if isinstance(padding, (tuple, list)) else (padding,) * nd)
    d.n, d.cin, d.id, d.ih, d.iw = x_shape[0], x_shape[1], sp[0], sp[1], sp[2]
    d.cout = w_shape[0]
    d.kd, d.kh, d.kw = k
    d.sd, d.sh, d.sw = s
    d.pd, d.ph, d.pw = p
    d.od = (d.id + 2 * d.pd - d.kd) // d.sd + 1
    d.oh = (d.ih + 2 * d.ph - d.kh) // d.sh + 1
    d.ow = (d.iw + 2 * d.pw - d.kw) // d.sw + 1
    out_sp = (d.od, d.oh, d.ow)[3 - nd:]
    return d, (x_shape[0], w_shape[0]) + out_sp


# ---------------------------------------------------------------------------------------------------------------------
# Tensor-core path (default): im2col + GEMMs on the tcgen05 implicit-GEMM kernel (1x1 form) in hi|lo split-bf16 arithmetic
# (x_hi w_hi + x_lo w_hi + x_hi w_lo, fp32 accumulation in TMEM; see csrc/disc_gemm.cu).  K = Cin * taps is padded to a
# multiple of 128 (kp).  Accuracy against fp64 convolutions (tests/test_gpu_disc.py): y ~2e-5, dx ~5e-6, dw ~1e-5
# max-normalised -- the forward figure is set by the tensor core's fp32 accumulation over K / 16 chained MMAs, not by the
# split (a three-term split measured the same 2e-5), against ~2e-6 for the FFMA kernels of csrc/disc.cu.  Those remain
# selectable (TENSOR_CORE = False: exact-fp32 parity mode, 4.5x slower); the reference itself runs these convolutions
# through cuDNN, whose default on Ampere-and-later GPUs is TF32 (~1e-3).
# ---------------------------------------------------------------------------------------------------------------------
_KCH = 4096      # K (or position) chunk of one GEMM launch: 3 x 4096 / 64 = 192 of the planner's 256 K steps
TENSOR_CORE = True
TC_PARTS = {"fwd": True, "dgrad": True, "wgrad": True}   # experiments: which of the three GEMMs take the tensor-core path


def _geom(d):
    K = d.cin * d.kd * d.kh * d.kw
    return d.n, d.cout, d.od * d.oh * d.ow, K, (K + 127) // 128 * 128


def _use_tc(d) -> bool:
    return TENSOR_CORE and d.cin * d.kd * d.kh * d.kw >= 64


def _w2(w, cout, K, kp):
    """the parameter as the [cout, kp] GEMM operand (zero columns for k >= K)"""
    w2 = w.detach().reshape(cout, K)
    if kp != K:
        w2 = torch.nn.functional.pad(w2, (0, kp - K))
    return w2


def _tc_forward(x, w, b, d, out_shape):
    from .. import ops
    lib = L.load()
    n, cout, P, K, kp = _geom(d)
    col = torch.empty((n, 1, P, 2 * kp), dtype=torch.bfloat16, device=x.device)
    L.check(lib.fo_dconv_im2col_pairs(C.byref(d), x.data_ptr(), col.data_ptr(), kp, 2, _stream()), "fo_dconv_im2col_pairs")
    _count(1)
    y = torch.empty(out_shape, dtype=torch.float32, device=x.device)
    bias = None
    if b is not None:
        bias = torch.zeros(ops.pad16(cout), dtype=torch.float32, device=x.device)
        bias[:cout] = b.detach()
    for k0 in range(0, kp, _KCH):
        kc = min(_KCH, kp - k0)
        ops.conv(ops.FORM_S1, 2, 1, [(col, kc, k0)],
                 lambda k0=k0, kc=kc: _w2(w, cout, K, kp)[:, k0:k0 + kc].contiguous().view(cout, kc, 1, 1),
                 0, cout, bias=bias if k0 == 0 else None, want_raw=False, f32="nchw", precise=True, out_f32=y,
                 f32_accumulate=k0 > 0, wkey=(w, ("dfwd", k0)))
    return y


def _tc_dgrad(dy, w, d, x_shape):
    from .. import ops
    lib = L.load()
    n, cout, P, K, kp = _geom(d)
    cp = ops.pad16(cout)
    # dy [n, cout, P] -> pairs channels-last [n * P, 2 * cp]
    dyp = torch.empty((n, 1, P, 2 * cp), dtype=torch.bfloat16, device=dy.device)
    L.check(lib.fo_split_f32(dy.data_ptr(), n, cout, P, cout * P, P, 1, dyp.data_ptr(), cp, _stream()), "fo_split_f32")
    _count(1)
    dcol = torch.empty((n, 1, P, kp), dtype=torch.float32, device=dy.device)
    ops.conv(ops.FORM_S1, 2, 1, [(dyp, cout, 0)], lambda: _w2(w, cout, K, kp).contiguous().view(cout, kp, 1, 1), 1, kp,
             want_raw=False, f32="cl", out_cs=kp, precise=True, out_f32=dcol, wkey=(w, "ddgrad"))
    dx = torch.empty(x_shape, dtype=torch.float32, device=dy.device)
    L.check(lib.fo_dconv_col2im(C.byref(d), dcol.data_ptr(), kp, dx.data_ptr(), _stream()), "fo_dconv_col2im")
    _count(1)
    return dx


def _tc_wgrad(x, dy, w, d, has_bias):
    from .. import ops
    lib = L.load()
    n, cout, P, K, kp = _geom(d)
    rows = n * P
    chunks = (rows + _KCH - 1) // _KCH
    pp = chunks * _KCH
    # im2col^T, already packed as the (hi | hi | lo) weight operand of each position chunk
    colT = torch.empty((chunks, kp, 3 * _KCH), dtype=torch.bfloat16, device=x.device)
    L.check(lib.fo_dconv_im2col_t(C.byref(d), x.data_ptr(), colT.data_ptr(), kp, _KCH, chunks, _stream()), "fo_dconv_im2col_t")
    _count(1)
    # dy^T [cout, n * P] as the activation operand ("pixels" = output channels, "channels" = positions)
    dyt = dy if n == 1 else dy.view(n, cout, P).permute(1, 0, 2).reshape(cout, rows).contiguous()
    cpix = ops.pad16(cout) if cout < 16 else cout
    dyp = torch.empty((1, 1, cpix, 2 * pp), dtype=torch.bfloat16, device=x.device)
    if cpix != cout:
        dyp.zero_()
    L.check(lib.fo_split_f32(dyt.data_ptr(), 1, rows, cout, 0, 1, rows, dyp.data_ptr(), pp, _stream()), "fo_split_f32")
    _count(1)
    dw = torch.empty((1, 1, cpix, kp), dtype=torch.float32, device=x.device)
    for i in range(chunks):
        ops.conv(ops.FORM_S1, 2, 1, [(dyp, _KCH, i * _KCH)], None, 0, kp, want_raw=False, f32="cl", out_cs=kp,
                 precise=True, out_f32=dw, f32_accumulate=i > 0, wpacked=colT[i], wkey=(w, "unused"))
    db = None
    if has_bias:
        db = torch.empty(cout, dtype=torch.float32, device=x.device)
        L.check(lib.fo_dconv_dbias(C.byref(d), dy.data_ptr(), db.data_ptr(), _stream()), "fo_dconv_dbias")
        _count(1)
    return dw.view(cpix, kp)[:cout, :K].reshape(w.shape), db


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, stride, padding):
        _req(x, "conv input")
        lib = L.load()
        x = x.contiguous()
        wd = w.detach().contiguous()
        d, out_shape = _desc(x.shape, w.shape, stride, padding)
        ctx.tc = _use_tc(d)
        ctx.w_param = w     # the parameter object keys the packed-weight cache (version-checked) in forward and backward
        if ctx.tc and TC_PARTS["fwd"]:
            y = _tc_forward(x, w, b, d, out_shape)
        else:
            y = torch.empty(out_shape, dtype=torch.float32, device=x.device)
            L.check(lib.fo_dconv_fwd(C.byref(d), x.data_ptr(), wd.data_ptr(), None if b is None else b.detach().data_ptr(),
                                     y.data_ptr(), _stream()), "fo_dconv_fwd")
            _count(1)
        ctx.save_for_backward(x, wd)
        ctx.d, ctx.has_bias = d, b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        d = ctx.d
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            if ctx.tc and TC_PARTS["dgrad"]:
                dx = _tc_dgrad(dy, ctx.w_param, d, x.shape)
            else:
                dx = torch.empty_like(x)
                L.check(lib.fo_dconv_dgrad(C.byref(d), dy.data_ptr(), w.data_ptr(), dx.data_ptr(), _stream()), "fo_dconv_dgrad")
                _count(1)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            if ctx.tc and TC_PARTS["wgrad"]:
                dw, db = _tc_wgrad(x, dy, ctx.w_param, d, ctx.has_bias)
            else:
                dw = torch.empty_like(w)
                db = torch.empty(w.shape[0], dtype=torch.float32, device=w.device) if ctx.has_bias else None
                L.check(lib.fo_dconv_wgrad(C.byref(d), x.data_ptr(), dy.data_ptr(), dw.data_ptr(),
                                           None if db is None else db.data_ptr(), _stream()), "fo_dconv_wgrad")
                _count(2)
        return dx, dw, db, None, None


class Conv2d(nn.Conv2d):
    def forward(self, input):
        return _ConvFn.apply(input, self.weight, self.bias, self.stride, self.padding)


class Conv3d(nn.Conv3d):
    def forward(self, input):
        return _ConvFn.apply(input, self.weight, self.bias, self.stride, self.padding)


class _InstNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, running_mean, running_var, training, momentum, eps):
        _req(x, "instance norm input")
        lib = L.load()
        x = x.contiguous()
        n, c = x.shape[:2]
        plane = x[0, 0].numel()
        y = torch.empty_like(x)
        save = torch.empty(2 * n * c, dtype=torch.float32, device=x.device)
        L.check(lib.fo_instnorm_fwd(x.data_ptr(), y.data_ptr(), n, c, plane, eps, 1.0, int(training), momentum,
                                    None if running_mean is None else running_mean.data_ptr(),
                                    None if running_var is None else running_var.data_ptr(), save.data_ptr(), _stream()),
                "fo_instnorm_fwd")
        _count(2 if (training and running_mean is not None) else 1)
        ctx.save_for_backward(y, save)
        ctx.training = bool(training)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        y, save = ctx.saved_tensors
        dy = dy.contiguous()
        n, c = y.shape[:2]
        dx = torch.empty_like(y)
        L.check(lib.fo_instnorm_bwd(y.data_ptr(), dy.data_ptr(), dx.data_ptr(), n, c, y[0, 0].numel(), 1.0,
                                    int(ctx.training), save.data_ptr(), _stream()), "fo_instnorm_bwd")
        _count(1)
        return dx, None, None, None, None, None


class _InstanceNormMixin:
    """forward of nn.InstanceNorm*d(affine=False): instance statistics in training mode (running estimates updated when
    track_running_stats), the running estimates in eval mode when they are tracked."""

    def forward(self, input):
        if self.affine:
            raise L.FaceoffB200Error("InstanceNorm: affine=True is not used by the reference discriminators")
        use_running = (not self.training) and self.track_running_stats
        # (nn.InstanceNorm*d never advances num_batches_tracked -- only BatchNorm's forward does -- so neither does this)
        momentum = 0.1 if self.momentum is None else self.momentum
        rm = self.running_mean if self.track_running_stats else None
        rv = self.running_var if self.track_running_stats else None
        return _InstNormFn.apply(input, rm, rv, not use_running, momentum, self.eps)


class InstanceNorm2d(_InstanceNormMixin, nn.InstanceNorm2d):
    pass


class InstanceNorm3d(_InstanceNormMixin, nn.InstanceNorm3d):
    pass


class _LReLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, slope):
        _req(x, "leaky relu input")
        lib = L.load()
        x = x.contiguous()
        y = torch.empty_like(x)
        L.check(lib.fo_lrelu(x.data_ptr(), y.data_ptr(), x.numel(), slope, _stream()), "fo_lrelu")
        _count(1)
        ctx.save_for_backward(y)
        ctx.slope = slope
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        (y,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(y)
        L.check(lib.fo_lrelu_bwd(y.data_ptr(), dy.data_ptr(), dx.data_ptr(), y.numel(), ctx.slope, _stream()), "fo_lrelu_bwd")
        _count(1)
        return dx, None


class LeakyReLU(nn.LeakyReLU):
    def forward(self, input):
        return _LReLUFn.apply(input, float(self.negative_slope))


class _AvgPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kd, stride3):
        _req(x, "avg pool input")
        lib = L.load()
        x = x.contiguous()
        nd = x.dim() - 2
        sp = (1,) * (3 - nd) + tuple(x.shape[2:])
        sd, sh, sw = stride3
        od = (sp[0] + 2 * (kd // 2) - kd) // sd + 1
        oh = (sp[1] + 2 - 3) // sh + 1
        ow = (sp[2] + 2 - 3) // sw + 1
        planes = x.shape[0] * x.shape[1]
        y = torch.empty(tuple(x.shape[:2]) + (od, oh, ow)[3 - nd:], dtype=torch.float32, device=x.device)
        L.check(lib.fo_avgpool3(x.data_ptr(), y.data_ptr(), planes, sp[0], sp[1], sp[2], od, oh, ow, kd, sd, sh, sw,
                                _stream()), "fo_avgpool3")
        _count(1)
        ctx.geom = (planes, sp, (od, oh, ow), kd, stride3, x.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        planes, sp, out, kd, (sd, sh, sw), xshape = ctx.geom
        dy = dy.contiguous()
        dx = torch.empty(xshape, dtype=torch.float32, device=dy.device)
        L.check(lib.fo_avgpool3_bwd(dy.data_ptr(), dx.data_ptr(), planes, sp[0], sp[1], sp[2], out[0], out[1], out[2], kd, sd,
                                    sh, sw, _stream()), "fo_avgpool3_bwd")
        _count(1)
        return dx, None, None


def _check_pool(m, nd):
    k = _triple(m.kernel_size, nd)[3 - nd:]
    p = tuple(m.padding) if isinstance(m.padding, (tuple, list)) else (m.padding,) * nd
    if any(v != 3 for v in k) or any(v != 1 for v in p) or m.count_include_pad or m.ceil_mode:
        raise L.FaceoffB200Error("AvgPool: only kernel 3, padding 1, count_include_pad=False (the reference's downsample)")


class AvgPool2d(nn.AvgPool2d):
    def forward(self, input):
        _check_pool(self, 2)
        s = _triple(self.stride, 2)
        return _AvgPoolFn.apply(input, 1, (1, s[1], s[2]))


class AvgPool3d(nn.AvgPool3d):
    def forward(self, input):
        _check_pool(self, 3)
        return _AvgPoolFn.apply(input, 3, _triple(self.stride, 3))
