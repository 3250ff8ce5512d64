"""Discriminator convolutions: tensor-core (im2col + split-bf16 GEMM) path against the fp32 CUDA-core kernels, per layer --
max-normalised differences of y / dx / dw / db and CUDA-event times; then the full discriminator step.

    python tests/gpu_profile_disc.py [check|step|all]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from faceoff_b200.mocoganhd import content_disc, layers, losses, video_disc  # noqa: E402

dev = "cuda"


def mn(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-300)).item()


def run_layer(m, xx, go, tc):
    layers.TENSOR_CORE = tc
    for _ in range(2):
        m.zero_grad()
        xx.grad = None
        y = m(xx)
        y.backward(go)
    torch.cuda.synchronize()
    a, b, c = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    m.zero_grad()
    xx.grad = None
    a.record()
    y = m(xx)
    b.record()
    y.backward(go)
    c.record()
    torch.cuda.synchronize()
    return y.detach(), xx.grad.clone(), m.weight.grad.clone(), m.bias.grad.clone(), a.elapsed_time(b), b.elapsed_time(c)


def check():
    torch.manual_seed(0)
    cases = [("3d 64->128 s2", layers.Conv3d, (1, 64, 6, 129, 129), 128, 2),
             ("3d 128->256 s2", layers.Conv3d, (1, 128, 4, 65, 65), 256, 2),
             ("3d 256->512 s1", layers.Conv3d, (1, 256, 3, 33, 33), 512, 1),
             ("2d 64->128 s2", layers.Conv2d, (2, 64, 129, 129), 128, 2),
             ("2d 256->512 s1", layers.Conv2d, (1, 256, 33, 33), 512, 1),
             ("2d 6->64 s2", layers.Conv2d, (1, 6, 256, 256), 64, 2),
             ("2d 512->1 s1", layers.Conv2d, (1, 512, 34, 34), 1, 1),
             ("3d 6->64 s2", layers.Conv3d, (1, 6, 11, 256, 256), 64, 2),
             ("3d 512->1 s1", layers.Conv3d, (1, 512, 4, 34, 34), 1, 1),
             ("2d 6->32 s2 (scale 1)", layers.Conv2d, (1, 6, 128, 128), 32, 2)]
    for name, cls, shp, cout, stride in cases:
        m = cls(shp[1], cout, 4, stride=stride, padding=2).cuda()
        xx = torch.randn(shp, device=dev, requires_grad=True)
        y0 = m(xx)
        go = torch.randn_like(y0)
        ref = run_layer(m, xx, go, False)
        got = run_layer(m, xx, go, True)
        fl = 2.0 * m.weight.numel() * y0[0, 0].numel() * shp[0]
        print(f"{name}: y {mn(got[0], ref[0]):.1e} dx {mn(got[1], ref[1]):.1e} dw {mn(got[2], ref[2]):.1e} db {mn(got[3], ref[3]):.1e} | "
              f"fp32 fwd {ref[4]:.2f} bwd {ref[5]:.2f} ms | tensor fwd {got[4]:.2f} ms ({fl / got[4] / 1e9:.0f} TF/s) "
              f"bwd {got[5]:.2f} ms ({2 * fl / got[5] / 1e9:.0f} TF/s)")


def step():
    torch.manual_seed(0)
    d3 = video_disc.ModelD_3d(3, "instance", 2, 1e-4, False, 12).to(dev).train()
    d2 = content_disc.ModelD_img(3, "instance", 2, 1e-4).to(dev).train()
    crit = losses.Relativistic_Average_LSGAN()
    xr3, xf3 = [(torch.rand(1, 6, 11, 256, 256) * 2 - 1).to(dev) for _ in range(2)]
    xr2, xf2 = [(torch.rand(1, 6, 256, 256) * 2 - 1).to(dev) for _ in range(2)]

    def one():
        for m, xr, xf in ((d3, xr3, xf3), (d2, xr2, xf2)):
            m.zero_grad()
            f, r = m(xf), m(xr)
            loss = (crit(r, f, True) + crit(f, r, False)) * 0.5
            loss.backward()

    for tc in (False, True, "bwd"):
        layers.TENSOR_CORE = bool(tc)
        layers.TC_PARTS["fwd"] = tc is True
        for _ in range(3):
            one()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            one()
        b.record()
        torch.cuda.synchronize()
        print(f"discriminator step (D_3d + D_img, fake+real, fwd+bwd), tensor cores {tc}: {a.elapsed_time(b) / 5:.2f} ms")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("check", "all"):
        check()
    if which in ("step", "all"):
        step()


def kernels():
    """per-kernel time table of one tensor-core discriminator step (torch profiler)"""
    from torch.profiler import ProfilerActivity, profile
    torch.manual_seed(0)
    d3 = video_disc.ModelD_3d(3, "instance", 2, 1e-4, False, 12).to(dev).train()
    d2 = content_disc.ModelD_img(3, "instance", 2, 1e-4).to(dev).train()
    crit = losses.Relativistic_Average_LSGAN()
    xr3, xf3 = [(torch.rand(1, 6, 11, 256, 256) * 2 - 1).to(dev) for _ in range(2)]
    xr2, xf2 = [(torch.rand(1, 6, 256, 256) * 2 - 1).to(dev) for _ in range(2)]

    def one():
        for m, xr, xf in ((d3, xr3, xf3), (d2, xr2, xf2)):
            m.zero_grad()
            f, r = m(xf), m(xr)
            loss = (crit(r, f, True) + crit(f, r, False)) * 0.5
            loss.backward()

    layers.TENSOR_CORE = True
    layers.TC_PARTS["fwd"] = len(sys.argv) > 2 and sys.argv[2] == "fwdtc"
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        one()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "kernels":
    kernels()
