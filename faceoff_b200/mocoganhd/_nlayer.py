"""Shared builder of the 2-D (content) and 3-D (video) multi-scale PatchGAN discriminators.

Structure (reference mocoganhd_content_disc.py:49-165 / mocoganhd_video_disc.py:55-176, pix2pixHD lineage): ``num_D``
N-layer discriminators applied to an input pyramid (AvgPool 3 / pad 1 / count_include_pad=False between the scales); each
N-layer discriminator is conv(k4, s2, p2) + LeakyReLU(0.2), (n_layers - 1) x [conv(k4, s2, p2) + norm + LeakyReLU],
[conv(k4, s1, p2) + norm + LeakyReLU], conv(k4, s1, p2) -> 1 channel.  With ``getIntermFeat`` every block is a separate
``nn.Sequential`` registered as ``scale{i}_layer{j}`` (these names are the checkpoint keys) and forward returns, per
scale, the list of all block outputs (the last one is the patch prediction).
"""
from __future__ import annotations

import functools
import math

from torch import nn

from . import layers


def norm_layer_for(norm_type, ndim):
    """reference get_norm_layer: 'instance' -> InstanceNorm(affine=False, track_running_stats=True)."""
    if norm_type == "instance":
        cls = layers.InstanceNorm2d if ndim == 2 else layers.InstanceNorm3d
        return functools.partial(cls, affine=False, track_running_stats=True)
    if norm_type == "batch":
        raise NotImplementedError("norm 'batch' is not on the FaceOff path (the trainer builds the discriminators with "
                                  "'instance'); only the instance-norm kernels exist")
    raise NotImplementedError("normalization layer [%s] is not found" % norm_type)


def init_weights(m):
    """reference weights_init: Conv weights ~ N(0, 0.02) (biases keep the default init)."""
    name = m.__class__.__name__
    if name.find("Conv") != -1 and hasattr(m, "weight"):
        m.weight.data.normal_(0.0, 0.02)
    elif name.find("BatchNorm") != -1:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)


def nlayer_blocks(ndim, input_nc, ndf, n_layers, norm_layer):
    conv = layers.Conv2d if ndim == 2 else layers.Conv3d
    kw = 4
    padw = int(math.ceil((kw - 1.0) / 2))
    blocks = [[conv(input_nc, ndf, kernel_size=kw, stride=2, padding=padw), layers.LeakyReLU(0.2, True)]]
    nf = ndf
    for _ in range(1, n_layers):
        nf_prev, nf = nf, min(nf * 2, 512)
        blocks.append([conv(nf_prev, nf, kernel_size=kw, stride=2, padding=padw), norm_layer(nf), layers.LeakyReLU(0.2, True)])
    nf_prev, nf = nf, min(nf * 2, 512)
    blocks.append([conv(nf_prev, nf, kernel_size=kw, stride=1, padding=padw), norm_layer(nf), layers.LeakyReLU(0.2, True)])
    blocks.append([conv(nf, 1, kernel_size=kw, stride=1, padding=padw)])
    return blocks


class NLayerDiscriminatorBase(nn.Module):
    NDIM = 2

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=None, getIntermFeat=True):
        super().__init__()
        self.getIntermFeat = getIntermFeat
        self.n_layers = n_layers
        norm_layer = norm_layer or norm_layer_for("instance", self.NDIM)
        blocks = nlayer_blocks(self.NDIM, input_nc, ndf, n_layers, norm_layer)
        if getIntermFeat:
            for n, b in enumerate(blocks):
                setattr(self, "model" + str(n), nn.Sequential(*b))
        else:
            self.model = nn.Sequential(*[m for b in blocks for m in b])

    def forward(self, input):
        if not self.getIntermFeat:
            return self.model(input)
        res = [input]
        for n in range(self.n_layers + 2):
            res.append(getattr(self, "model" + str(n))(res[-1]))
        return res[1:]


class MultiscaleDiscriminatorBase(nn.Module):
    NDIM = 2
    NLAYER = NLayerDiscriminatorBase

    def __init__(self, input_nc, ndf=64, n_layers=3, n_frames=16, norm_layer=None, num_D=2, getIntermFeat=True):
        super().__init__()
        self.num_D = num_D
        self.n_layers = n_layers
        self.getIntermFeat = getIntermFeat
        ndf_max = 64
        norm_layer = norm_layer or norm_layer_for("instance", self.NDIM)
        for i in range(num_D):
            netD = self.NLAYER(input_nc, min(ndf_max, ndf * (2 ** (num_D - 1 - i))), n_layers, norm_layer, getIntermFeat)
            if getIntermFeat:
                for j in range(n_layers + 2):
                    setattr(self, "scale" + str(i) + "_layer" + str(j), getattr(netD, "model" + str(j)))
            else:
                setattr(self, "layer" + str(i), netD.model)
        self.downsample = self._make_downsample(n_frames)

    def _make_downsample(self, n_frames):
        return layers.AvgPool2d(3, stride=2, padding=[1, 1], count_include_pad=False)

    def singleD_forward(self, model, input):
        if not self.getIntermFeat:
            return [model(input)]
        result = [input]
        for m in model:
            result.append(m(result[-1]))
        return result[1:]

    def forward(self, input):
        result = []
        x = input
        for i in range(self.num_D):
            k = self.num_D - 1 - i
            if self.getIntermFeat:
                model = [getattr(self, "scale" + str(k) + "_layer" + str(j)) for j in range(self.n_layers + 2)]
            else:
                model = getattr(self, "layer" + str(k))
            result.append(self.singleD_forward(model, x))
            if i != self.num_D - 1:
                x = self.downsample(x)
        return result
