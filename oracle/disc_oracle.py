"""CPU oracle for the MoCoGAN-HD discriminator step.  TEST INFRASTRUCTURE ONLY (same rules as
faceoff_oracle.py): a functional restatement with torch CPU ops of

    TemporalAlignment/models/mocoganhd_content_disc.py:49-165   (MultiscaleDiscriminator / NLayerDiscriminator, 2-D)
    TemporalAlignment/models/mocoganhd_video_disc.py:55-176     (the 3-D variants)
    TemporalAlignment/models/mocoganhd_losses.py:52-126         (GANLoss, Relativistic_Average_LSGAN, least squares)
    disc_trainers/train_vqvae_perceptual_mocoganhd_disc.py:240-300 (discriminator step)

operating on a plain state_dict with the reference's keys (``scale{i}_layer{j}.{idx}.weight`` ...).  Pinned against the
reference classes themselves by tests/golden/make_golden_disc.py (the reference modules import only torch / numpy).

Constructor arguments: the factory that builds ModelD_img / ModelD_3d is not part of the reference repository (SURVEY
8(f1)); the upstream MoCoGAN-HD defaults are assumed: nc=3, norm 'instance', num_D=2, lr=1e-4, cross_domain=False,
n_frames_G=12 (the trainer samples SAMPLE_FRAMES=12 frames, :164).
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def _conv(x, w, b, stride, ndim):
    return (F.conv2d if ndim == 2 else F.conv3d)(x, w, b, stride=stride, padding=2)   # kw = 4, padw = ceil(3 / 2) = 2


def nlayer_forward(sd: Dict[str, Tensor], prefix: str, x: Tensor, ndim: int, n_layers: int = 3, training: bool = True,
                   new_stats: Dict[str, Tensor] = None) -> List[Tensor]:
    """One N-layer discriminator (content :110-165 / video :119-176); blocks named ``{prefix}{j}`` = scale{i}_layer{j}.
    InstanceNorm(affine=False, track_running_stats=True): instance statistics in training (running estimates updated with
    momentum 0.1 into ``new_stats``), running estimates in eval."""
    feats = []
    for j in range(n_layers + 2):
        p = f"{prefix}{j}"
        stride = 2 if j < n_layers else 1
        x = _conv(x, sd[p + ".0.weight"], sd[p + ".0.bias"], stride, ndim)
        if 0 < j <= n_layers:
            rm, rv = sd[p + ".1.running_mean"].clone(), sd[p + ".1.running_var"].clone()
            x = F.instance_norm(x, rm, rv, None, None, use_input_stats=training, momentum=0.1, eps=1e-5)
            if new_stats is not None and training:
                new_stats[p + ".1.running_mean"], new_stats[p + ".1.running_var"] = rm, rv
        if j <= n_layers:
            x = F.leaky_relu(x, 0.2)
        feats.append(x)
    return feats


def multiscale_forward(sd: Dict[str, Tensor], x: Tensor, ndim: int, num_D: int = 2, n_layers: int = 3, n_frames: int = 11,
                       training: bool = True, new_stats: Dict[str, Tensor] = None, prefix: str = "netD.") -> List[List[Tensor]]:
    """MultiscaleDiscriminator.forward (content :85-105 / video :100-116): scale num_D-1 sees the full resolution, the
    input is average-pooled (3, pad 1, count_include_pad=False; stride 2, or [1, 2, 2] for clips of <= 16 frames) between
    scales."""
    result = []
    for i in range(num_D):
        k = num_D - 1 - i
        result.append(nlayer_forward(sd, f"{prefix}scale{k}_layer", x, ndim, n_layers, training, new_stats))
        if i != num_D - 1:
            if ndim == 2:
                x = F.avg_pool2d(x, 3, stride=2, padding=[1, 1], count_include_pad=False)
            else:
                x = F.avg_pool3d(x, 3, stride=2 if n_frames > 16 else [1, 2, 2], padding=[1, 1, 1], count_include_pad=False)
    return result


def ra_lsgan(out_1: List[List[Tensor]], out_2: List[List[Tensor]], target_is_real: bool) -> Tensor:
    """Relativistic_Average_LSGAN.__call__ (losses :114-120): sum over scales of MSE(pred - mean(other pred), label)."""
    t = 1.0 if target_is_real else 0.0
    loss = 0
    for a, b in zip(out_1, out_2):
        pred = a[-1] - torch.mean(b[-1])
        loss = loss + F.mse_loss(pred, torch.full_like(pred, t))
    return loss


def disc_loss(sd, x_real, x_fake, ndim, n_frames=11, new_stats=None):
    """Discriminator-side loss of the step (trainer :258-266, :282-290): fake first, then real."""
    d_fake = multiscale_forward(sd, x_fake, ndim, n_frames=n_frames, new_stats=new_stats)
    sd2 = dict(sd, **(new_stats or {}))
    d_real = multiscale_forward(sd2, x_real, ndim, n_frames=n_frames, new_stats=new_stats)
    return (ra_lsgan(d_real, d_fake, True) + ra_lsgan(d_fake, d_real, False)) * 0.5, d_real, d_fake


def gen_loss(sd, x_real, x_fake, ndim, n_frames=11):
    """Generator-side adversarial loss (trainer :208-227): fake first, then real."""
    d_fake = multiscale_forward(sd, x_fake, ndim, n_frames=n_frames)
    d_real = multiscale_forward(sd, x_real, ndim, n_frames=n_frames)
    return (ra_lsgan(d_fake, d_real, True) + ra_lsgan(d_real, d_fake, False)) * 0.5
