"""Video (motion) discriminator: drop-in for TemporalAlignment/models/mocoganhd_video_disc.py."""
from __future__ import annotations

import torch
from torch import nn

from . import _nlayer, layers


def get_norm_layer(norm_type="instance"):
    return _nlayer.norm_layer_for(norm_type, 3)


weights_init = _nlayer.init_weights


class NLayerDiscriminator(_nlayer.NLayerDiscriminatorBase):
    NDIM = 3


class MultiscaleDiscriminator(_nlayer.MultiscaleDiscriminatorBase):
    NDIM = 3
    NLAYER = NLayerDiscriminator

    def _make_downsample(self, n_frames):
        # reference :80-89 -- long clips are also halved in time
        stride = 2 if n_frames > 16 else [1, 2, 2]
        return layers.AvgPool3d(3, stride=stride, padding=[1, 1, 1], count_include_pad=False)


class ModelD_3d(nn.Module):
    """reference :8-30 -- classifies [N, C, T, H, W] clips; with cross_domain False the input pairs every frame with the
    first one (2 * nc channels, n_frames_G - 1 frames)."""

    def __init__(self, nc, norm_D_3d, num_D, lr, cross_domain, n_frames_G):
        super().__init__()
        if not cross_domain:
            nc, n_frames_G = nc * 2, n_frames_G - 1
        self.netD = MultiscaleDiscriminator(input_nc=nc, n_frames=n_frames_G, norm_layer=get_norm_layer(norm_D_3d),
                                            num_D=num_D)
        self.netD.apply(weights_init)
        self.optim = torch.optim.Adam(self.netD.parameters(), lr=lr, betas=(0.5, 0.999))

    def forward(self, x):
        return self.netD.forward(x)
