"""Isolated discriminator conv layers (tensor-core path and fp32 path) against torch fp64 convolutions (test tooling)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from faceoff_b200.mocoganhd import layers  # noqa: E402


def mn(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-300)).item()


cases = [("2d 6->32 s2", 2, (1, 6, 128, 128), 32, 2), ("2d 32->64 s2", 2, (1, 32, 65, 65), 64, 2),
         ("2d 64->128 s2", 2, (1, 64, 33, 33), 128, 2), ("2d 128->256 s1", 2, (1, 128, 17, 17), 256, 1),
         ("2d 256->1 s1", 2, (1, 256, 18, 18), 1, 1), ("3d 32->64 s2", 3, (1, 32, 6, 65, 65), 64, 2),
         ("3d 128->256 s1", 3, (1, 128, 3, 17, 17), 256, 1), ("2d 256->512 s1", 2, (1, 256, 33, 33), 512, 1)]
torch.manual_seed(0)
for name, nd, shp, cout, stride in cases:
    cls = layers.Conv2d if nd == 2 else layers.Conv3d
    m = cls(shp[1], cout, 4, stride=stride, padding=2).cuda()
    x = torch.randn(shp, device="cuda", requires_grad=True)
    conv = F.conv2d if nd == 2 else F.conv3d
    x64 = x.detach().double().requires_grad_(True)
    w64 = m.weight.detach().double().requires_grad_(True)
    y64 = conv(x64, w64, m.bias.detach().double(), stride=stride, padding=2)
    go = torch.randn_like(y64)
    gx64, gw64 = torch.autograd.grad(y64, [x64, w64], go)
    out = []
    for tc in (False, True):
        layers.TENSOR_CORE = tc
        m.zero_grad()
        x.grad = None
        y = m(x)
        y.backward(go.float())
        out.append((mn(y, y64), mn(x.grad, gx64), mn(m.weight.grad, gw64)))
    print(f"{name:18s} fp32: y {out[0][0]:.1e} dx {out[0][1]:.1e} dw {out[0][2]:.1e} | tensor: y {out[1][0]:.1e} dx {out[1][1]:.1e} dw {out[1][2]:.1e}")
