"""GAN losses used by the discriminator trainer: drop-in for the GANLoss / Relativistic_Average_LSGAN classes of
TemporalAlignment/models/mocoganhd_losses.py:52-126 (least-squares variant; the BCE variant is not on the FaceOff path)."""
from __future__ import annotations

import ctypes as C  # noqa: F401

import torch
from torch import nn

from .. import _lib as L
from ..ops import _count, _stream


class _RaLsganFn(torch.autograd.Function):
    """mean((a - mean(b) - target)^2) with gradients to a and b (mocoganhd_losses.py:118-119: nn.MSELoss on
    ``pred - torch.mean(_pred)`` against a constant target)."""

    @staticmethod
    def forward(ctx, a, b, target):
        lib = L.load()
        for t in (a, b):
            if not (t.is_cuda and t.dtype == torch.float32):
                raise L.FaceoffB200Error("GAN loss: fp32 CUDA predictions required (faceoff_b200 has no CPU path)")
        a, b = a.contiguous(), b.contiguous()
        out = torch.empty(2, dtype=torch.float32, device=a.device)
        L.check(lib.fo_ralsgan(a.data_ptr(), a.numel(), b.data_ptr(), b.numel(), float(target), out.data_ptr(), _stream()),
                "fo_ralsgan")
        _count(1)
        ctx.save_for_backward(a, out)
        ctx.target, ctx.b_shape = float(target), b.shape
        return out[0]

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        a, out = ctx.saved_tensors
        g = g.detach().reshape(1).to(torch.float32).contiguous()
        da = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        db = torch.empty(ctx.b_shape, dtype=torch.float32, device=a.device) if ctx.needs_input_grad[1] else None
        m = 1
        for v in ctx.b_shape:
            m *= v
        L.check(lib.fo_ralsgan_bwd(a.data_ptr(), a.numel(), m, ctx.target, out.data_ptr(), g.data_ptr(),
                                   None if da is None else da.data_ptr(), None if db is None else db.data_ptr(), _stream()),
                "fo_ralsgan_bwd")
        _count(1)
        return da, db, None


def _zero_like(pred):
    return torch.zeros(1, dtype=torch.float32, device=pred.device)


class GANLoss(nn.Module):
    """reference :52-106 (use_lsgan=True): MSE between the patch prediction and a constant real / fake label; for the
    multi-scale output the per-scale losses are summed."""

    def __init__(self, use_lsgan=True, target_real_label=1.0, target_fake_label=0.0, tensor=torch.FloatTensor):
        super().__init__()
        if not use_lsgan:
            raise NotImplementedError("only the least-squares GAN loss is on the FaceOff path")
        self.real_label = target_real_label
        self.fake_label = target_fake_label

    def _target(self, target_is_real):
        return self.real_label if target_is_real else self.fake_label

    def __call__(self, input, target_is_real):
        t = self._target(target_is_real)
        if isinstance(input[0], list):
            loss = 0
            for input_i in input:
                pred = input_i[-1]
                loss = loss + _RaLsganFn.apply(pred, _zero_like(pred), t)    # mean(b) = 0: plain LSGAN
            return loss
        return _RaLsganFn.apply(input[-1], _zero_like(input[-1]), t)


class Relativistic_Average_LSGAN(GANLoss):
    """reference :109-126 -- the prediction is compared with the label relative to the mean prediction on the other
    (real resp. fake) batch."""

    def __call__(self, input_1, input_2, target_is_real):
        t = self._target(target_is_real)
        if isinstance(input_1[0], list):
            loss = 0
            for input_i, _input_i in zip(input_1, input_2):
                loss = loss + _RaLsganFn.apply(input_i[-1], _input_i[-1], t)
            return loss
        return _RaLsganFn.apply(input_1[-1], input_2[-1], t)
