"""Image (content) discriminator: drop-in for TemporalAlignment/models/mocoganhd_content_disc.py."""
from __future__ import annotations

import torch
from torch import nn

from . import _nlayer


def get_norm_layer(norm_type="instance"):
    return _nlayer.norm_layer_for(norm_type, 2)


weights_init = _nlayer.init_weights


class NLayerDiscriminator(_nlayer.NLayerDiscriminatorBase):
    NDIM = 2


class MultiscaleDiscriminator(_nlayer.MultiscaleDiscriminatorBase):
    NDIM = 2
    NLAYER = NLayerDiscriminator


class ModelD_img(nn.Module):
    """reference :8-24 -- classifies a pair of frames (2 * nc channels) as real / fake content; owns its Adam
    optimizer (betas 0.5, 0.999) like the reference."""

    def __init__(self, nc, norm_D_3d, num_D, lr):
        super().__init__()
        self.netD = MultiscaleDiscriminator(input_nc=nc * 2, norm_layer=get_norm_layer(norm_D_3d), num_D=num_D)
        self.netD.apply(weights_init)
        self.optim = torch.optim.Adam(self.netD.parameters(), lr=lr, betas=(0.5, 0.999))

    def forward(self, x):
        return self.netD.forward(x)
