"""Reconstruction loss of the training step (reference train_faceoff_perceptual.py:38-40):

    recon_loss = nn.MSELoss()(out[:, :3], gt)

``mse_loss(out, gt)`` is that expression as one fused forward kernel and one fused backward kernel: the channel slice is
folded into the kernels (``gt`` decides how many channels of ``out`` are compared), so autograd neither materialises the
slice nor zero-fills and copies its gradient.  Same value and gradient as the reference expression.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import FaceoffB200Error


class _MSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, gt):
        ctx.save_for_backward(out, gt)
        n = gt.numel()
        ctx.scale = 2.0 / n
        return (ops.mse_sum(out, gt) / n).reshape(())

    @staticmethod
    def backward(ctx, g):
        out, gt = ctx.saved_tensors
        g = g.detach().to(torch.float32).reshape(1).contiguous()
        return ops.mse_grad(out, gt, g, ctx.scale), None


def mse_loss(out: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """mean((out[:, :gt.shape[1]] - gt) ** 2); ``out`` [N, Ca, H, W] and ``gt`` [N, C <= Ca, H, W], fp32 CUDA tensors.
    Only ``out`` receives a gradient (the target of the reference loss is data)."""
    if not (out.is_cuda and gt.is_cuda and out.dtype == torch.float32 and gt.dtype == torch.float32):
        raise FaceoffB200Error("mse_loss: fp32 CUDA tensors required (faceoff_b200 has no CPU path)")
    if out.dim() != 4 or gt.dim() != 4 or out.shape[0] != gt.shape[0] or out.shape[2:] != gt.shape[2:] or \
            gt.shape[1] > out.shape[1] or (out.shape[2] * out.shape[3]) % 4 != 0:
        raise FaceoffB200Error(f"mse_loss: unsupported shapes {tuple(out.shape)} vs {tuple(gt.shape)}")
    if gt.requires_grad:
        raise FaceoffB200Error("mse_loss: the target must not require a gradient")
    return _MSEFn.apply(out.contiguous(), gt.contiguous())
