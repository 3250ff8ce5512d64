"""In-situ per-launch verifier for the tensor-core convolution ops (test infrastructure).

``with OpVerifier() as v:`` wraps ``faceoff_b200.ops.conv`` / ``ops.wgrad``: after every launch the same operation is
recomputed with torch in fp64 ON THE TENSORS THE KERNEL ACTUALLY SAW (sources, weight, bias, mask, addend -- merged from
hi|lo pairs in the verification mode) and the max-normalised error of every output is recorded.  Unlike an end-to-end
comparison this localises an error to ONE launch (layer, form, tiling mode) and is immune to error propagation, so it can be
held to ~5e-5 in the verification mode and to one bf16 rounding in the product mode.

Reference semantics restated here (torch.nn.functional, fp64):
    FORM_S1        conv{2,3}d(x, W, padding=k//2)                      W [N, K, k..]      (n_axis 0)
    FORM_S1_DGRAD  conv_transpose{2,3}d(dy, W, padding=k//2)           W [K, N, k..]      (n_axis 1)
    FORM_DOWN      conv2d(x, W, stride=2, padding=1)                   W [N, K, 4, 4]     (n_axis 0)
    FORM_UP        conv_transpose2d(x, W, stride=2, padding=1)         W [K, N, 4, 4]     (n_axis 1)
(models/vqvae_conv3d_latent.py:86-190 forward ops and their autograd.)
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from faceoff_b200 import ops
from faceoff_b200.ops import FORM_DOWN, FORM_S1, FORM_S1_DGRAD, FORM_UP


def _logical(t: torch.Tensor, c: int, c_off: int) -> torch.Tensor:
    """fp64 [.., c] view of channels c_off..c_off+c of an activation tensor (pairs merged in the verification mode)."""
    if ops.PRECISE:
        half = t.shape[-1] // 2
        return t[..., c_off:c_off + c].double() + t[..., half + c_off:half + c_off + c].double()
    return t[..., c_off:c_off + c].double()


def _to_nc(x: torch.Tensor, ndim: int) -> torch.Tensor:       # channels-last -> N C (D) H W
    return x.permute(0, 4, 1, 2, 3) if ndim == 3 else x.permute(0, 3, 1, 2)


def _to_cl(x: torch.Tensor, ndim: int) -> torch.Tensor:
    return x.permute(0, 2, 3, 4, 1) if ndim == 3 else x.permute(0, 2, 3, 1)


def _maxnorm(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-300)).item()


class OpVerifier:
    def __init__(self, verbose: bool = False):
        self.records = []      # (kind, description, {output: err})
        self.verbose = verbose

    def __enter__(self):
        self._conv, self._wgrad = ops.conv, ops.wgrad
        ver = self

        def conv(form, ndim, ksize, srcs, weight, n_axis, cout, bias=None, mask=None, addend=None, want_raw=True,
                 want_relu=False, f32=None, relu_f32=False, n_scale=None, out_cs=None, out_raw=None, wkey=None):
            raw, relu, of32 = ver._conv(form, ndim, ksize, srcs, weight, n_axis, cout, bias=bias, mask=mask, addend=addend,
                                        want_raw=want_raw, want_relu=want_relu, f32=f32, relu_f32=relu_f32,
                                        n_scale=n_scale, out_cs=out_cs, out_raw=out_raw, wkey=wkey)
            x = torch.cat([_logical(t, c, off) for t, c, off in srcs], -1)
            w = weight.detach().double()
            xn = _to_nc(x, ndim)
            if form == FORM_S1:
                assert n_axis == 0
                y = (F.conv3d if ndim == 3 else F.conv2d)(xn, w, padding=ksize // 2)
            elif form == FORM_S1_DGRAD:
                assert n_axis == 1
                y = (F.conv_transpose3d if ndim == 3 else F.conv_transpose2d)(xn, w, padding=ksize // 2)
            elif form == FORM_DOWN:
                assert n_axis == 0
                y = F.conv2d(xn, w, stride=2, padding=1)
            else:
                assert n_axis == 1
                y = F.conv_transpose2d(xn, w, stride=2, padding=1)
            y = _to_cl(y, ndim)
            if bias is not None:
                y = y + bias.double()[:cout]
            if mask is not None:
                gate = (mask[..., :cout].float() > 0)          # the kernel gates on the (hi part of the) bf16 mask value
                y = torch.where(gate, y, torch.zeros((), dtype=y.dtype, device=y.device))
            if addend is not None:
                y = y + _logical(addend, cout, 0)
            errs = {}
            if raw is not None:
                errs["raw"] = _maxnorm(_logical(raw, cout, 0), y)
            if relu is not None:
                errs["relu"] = _maxnorm(_logical(relu, cout, 0), y.clamp_min(0))
            if of32 is not None:
                got = of32.double()
                got = _to_cl(got, 2) if f32 == "nchw" else got[..., :cout]
                errs["f32"] = _maxnorm(got, y.clamp_min(0) if relu_f32 else y)
            kind = {FORM_S1: "S1", FORM_S1_DGRAD: "S1_DGRAD", FORM_DOWN: "DOWN", FORM_UP: "UP"}[form]
            desc = (f"conv {kind} {ndim}d k{ksize} {[(tuple(t.shape), c, off) for t, c, off in srcs]} -> {cout}"
                    f"{' +bias' if bias is not None else ''}{' +mask' if mask is not None else ''}"
                    f"{' +addend' if addend is not None else ''}")
            ver._add("conv", desc, errs)
            return raw, relu, of32

        def wgrad(form, ndim, ksize, p, q, dweight, m_axis, q_w_off=0, accumulate=False, dbias=None,
                  dbias_accumulate=False, q_shift_sign=1, _raw=False):
            if _raw:      # inner launches of the verification mode's three-launch expansion
                return ver._wgrad(form, ndim, ksize, p, q, dweight, m_axis, q_w_off, accumulate, dbias, dbias_accumulate,
                                  q_shift_sign, _raw)
            before = dweight.detach().double().clone() if accumulate else None
            bias_before = dbias.detach().double().clone() if (dbias is not None and dbias_accumulate) else None
            ver._wgrad(form, ndim, ksize, p, q, dweight, m_axis, q_w_off, accumulate, dbias, dbias_accumulate,
                       q_shift_sign)
            P = _to_nc(_logical(*p), ndim)
            Q = _to_nc(_logical(*q), ndim)
            # dW[m][n][tap] = sum_pix P[pix][m] Q[pix + sign*tap][n]: the weight gradient of y = conv(Q, W) w.r.t. W with
            # dy = P when sign > 0; with sign < 0 the roles of the two tensors are swapped (x = P, dy = Q)
            if form == FORM_S1:
                x_, dy_ = (Q, P) if q_shift_sign >= 0 else (P, Q)
                wshape = (dy_.shape[1], x_.shape[1]) + (ksize,) * ndim
                with torch.enable_grad():       # we are inside autograd's backward, where grad mode is off
                    wz = torch.zeros(wshape, dtype=torch.float64, device=P.device, requires_grad=True)
                    y = (F.conv3d if ndim == 3 else F.conv2d)(x_, wz, padding=ksize // 2)
                    (g,) = torch.autograd.grad(y, wz, dy_)
                # g [dy ch][x ch][taps]; the kernel's [m = P ch][n = Q ch]
                gm = g if q_shift_sign >= 0 else g.transpose(0, 1)
            else:   # FORM_DOWN: P low-res, Q hi-res: dW[m][n] of conv2d(Q, W[m, n], stride 2, pad 1) with dy = P
                with torch.enable_grad():
                    wz = torch.zeros((P.shape[1], Q.shape[1], 4, 4), dtype=torch.float64, device=P.device,
                                     requires_grad=True)
                    y = F.conv2d(Q, wz, stride=2, padding=1)
                    (gm,) = torch.autograd.grad(y, wz, P)
            mc, nc = gm.shape[0], gm.shape[1]
            dw = dweight.detach().double()
            if m_axis == 0:
                got = dw[:mc, q_w_off:q_w_off + nc]
                ref = gm + (before[:mc, q_w_off:q_w_off + nc] if before is not None else 0)
            else:
                got = dw[q_w_off:q_w_off + nc, :mc].transpose(0, 1)
                ref = gm + (before[q_w_off:q_w_off + nc, :mc].transpose(0, 1) if before is not None else 0)
            errs = {"dW": _maxnorm(got, ref)}
            if dbias is not None:
                bref = P.sum(dim=[0] + list(range(2, P.dim()))) + (bias_before[:mc] if bias_before is not None else 0)
                errs["dbias"] = _maxnorm(dbias.detach().double()[:mc], bref)
            kind = "S1" if form == FORM_S1 else "DOWN"
            ver._add("wgrad", f"wgrad {kind} {ndim}d k{ksize} P{tuple(p[0].shape)}c{p[1]}@{p[2]} Q{tuple(q[0].shape)}c{q[1]}@{q[2]} "
                              f"m_axis {m_axis} sign {q_shift_sign}{' acc' if accumulate else ''}", errs)

        ops.conv, ops.wgrad = conv, wgrad
        return self

    def _add(self, kind, desc, errs):
        self.records.append((kind, desc, errs))
        if self.verbose:
            print(f"  [{len(self.records):3d}] " + " ".join(f"{k}={v:.1e}" for k, v in errs.items()) + "  " + desc)

    def __exit__(self, *exc):
        ops.conv, ops.wgrad = self._conv, self._wgrad

    def worst(self):
        w = (0.0, None)
        for kind, desc, errs in self.records:
            for k, v in errs.items():
                if not (v <= w[0]):
                    w = (v, f"{desc} [{k}]")
        return w
