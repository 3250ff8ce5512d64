"""Which discriminator conv path is closer to the truth?  fp64 oracle (torch ops on the GPU, test tooling only) against the
tensor-core split-bf16 path and the fp32 CUDA-core path: every parameter gradient of one discriminator step."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from faceoff_b200.mocoganhd import content_disc, layers, losses, video_disc  # noqa: E402
from oracle import disc_oracle as DO  # noqa: E402


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "img"
    torch.manual_seed(0)
    if kind == "img":
        m = content_disc.ModelD_img(3, "instance", 2, 1e-4).cuda().train()
        shape, ndim = (1, 6, 256, 256), 2
    else:
        m = video_disc.ModelD_3d(3, "instance", 2, 1e-4, False, 12).cuda().train()
        shape, ndim = (1, 6, 11, 256, 256), 3
    xr, xf = [(torch.rand(shape, device="cuda") * 2 - 1) for _ in range(2)]
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    crit = losses.Relativistic_Average_LSGAN()
    grads = {}
    modes = {False: dict(fwd=False, dgrad=False, wgrad=False), True: dict(fwd=True, dgrad=True, wgrad=True),
             "fwd": dict(fwd=True, dgrad=False, wgrad=False), "dgrad": dict(fwd=False, dgrad=True, wgrad=False),
             "wgrad": dict(fwd=False, dgrad=False, wgrad=True)}
    for tc, parts in modes.items():
        layers.TENSOR_CORE = True
        layers.TC_PARTS.update(parts)
        m.load_state_dict(sd0)
        m.zero_grad()
        f, r = m(xf), m(xr)
        loss = (crit(r, f, True) + crit(f, r, False)) * 0.5
        loss.backward()
        grads[tc] = {k: p.grad.clone() for k, p in m.named_parameters()}
    sd64 = {k: (v.double().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.double() if v.dtype.is_floating_point else v)
            for k, v in sd0.items()}
    loss64, _, _ = DO.disc_loss(sd64, xr.double(), xf.double(), ndim, n_frames=11, new_stats={})
    names = [k for k, _ in m.named_parameters()]
    g64 = torch.autograd.grad(loss64, [sd64[k] for k in names])
    print(f"{'parameter':40s} {'fp32 path':>10s} {'tensor path':>12s} {'fwd only':>10s} {'dgrad only':>10s} {'wgrad only':>10s}  (max-normalised error vs fp64)")
    for k, g in zip(names, g64):
        mx = g.abs().max().item() + 1e-300
        e32 = (grads[False][k].double() - g).abs().max().item() / mx
        etc = (grads[True][k].double() - g).abs().max().item() / mx
        rest = " ".join(f"{(grads[q][k].double() - g).abs().max().item() / mx:10.2e}" for q in ("fwd", "dgrad", "wgrad"))
        if k.endswith("weight"):
            print(f"{k:40s} {e32:10.2e} {etc:12.2e} {rest}")


if __name__ == "__main__":
    main()
