"""Generates tests/golden/golden_disc.pt from the UNMODIFIED reference discriminators (run in the build container, where
/root/reference exists):

    python tests/golden/make_golden_disc.py

For the image and the video discriminator: seeded construction (the checksums of every parameter are stored so that the
drop-in modules can prove they initialise identically from the same seed), seeded real / fake inputs, the patch
predictions of both scales, the discriminator-step loss (Relativistic_Average_LSGAN, trainer :258-290), every parameter
gradient norm + a slice, the running statistics after the two training-mode forwards, the generator-side loss and its
gradient w.r.t. the fake input.  Also asserts that oracle/disc_oracle.py reproduces all of it.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "TemporalAlignment", "models"))

import mocoganhd_content_disc as RC  # noqa: E402
import mocoganhd_losses as RL  # noqa: E402
import mocoganhd_video_disc as RV  # noqa: E402

from oracle import disc_oracle as DO  # noqa: E402


def run(kind):
    torch.manual_seed(7)
    if kind == "img":
        m = RC.ModelD_img(3, "instance", 2, 1e-4)
        shape, ndim = (1, 6, 64, 64), 2
    else:
        m = RV.ModelD_3d(3, "instance", 2, 1e-4, False, 12)
        shape, ndim = (1, 6, 5, 48, 48), 3
    m.train()
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(11)
    x_real = torch.rand(shape, generator=g) * 2 - 1
    x_fake = torch.rand(shape, generator=g) * 2 - 1
    crit = RL.Relativistic_Average_LSGAN()
    # ---- discriminator step (fake first, then real: trainer :258-259 / :282-283)
    d_fake = m(x_fake)
    d_real = m(x_real)
    d_loss = (crit(d_real, d_fake, True) + crit(d_fake, d_real, False)) * 0.5
    m.zero_grad()
    d_loss.backward()
    grads = {k: p.grad.clone() for k, p in m.named_parameters()}
    sd1 = {k: v.clone() for k, v in m.state_dict().items()}
    # ---- generator-side loss and gradient w.r.t. the fake input (weights / statistics as after construction)
    m.load_state_dict(sd0)
    xf = x_fake.clone().requires_grad_(True)
    df = m(xf)
    dr = m(x_real)
    g_loss = (crit(df, dr, True) + crit(dr, df, False)) * 0.5
    g_loss.backward()
    # ---- oracle pin
    ns = {}
    o_loss, o_real, o_fake = DO.disc_loss(sd0, x_real, x_fake, ndim, n_frames=11, new_stats=ns)
    torch.testing.assert_close(o_loss, d_loss.detach(), rtol=1e-6, atol=1e-8)
    for a, b in zip(o_real, d_real):
        for u, v in zip(a, b):
            torch.testing.assert_close(u, v.detach(), rtol=1e-5, atol=1e-6)
    for k, v in ns.items():
        torch.testing.assert_close(v, sd1[k], rtol=1e-5, atol=1e-7)
    p0 = {k: v.clone().requires_grad_(k.endswith(("weight", "bias"))) for k, v in sd0.items()}
    ol, _, _ = DO.disc_loss(p0, x_real, x_fake, ndim, n_frames=11, new_stats={})
    ol.backward()
    for k, gr in grads.items():
        torch.testing.assert_close(p0[k].grad, gr, rtol=1e-4, atol=1e-9)
    xo = x_fake.clone().requires_grad_(True)
    og = DO.gen_loss(sd0, x_real, xo, ndim, n_frames=11)
    og.backward()
    torch.testing.assert_close(og.detach(), g_loss.detach(), rtol=1e-6, atol=1e-8)
    torch.testing.assert_close(xo.grad, xf.grad, rtol=1e-4, atol=1e-10)
    return {
        "seed_model": 7, "seed_data": 11, "shape": shape,
        "param_checksums": {k: (v.double().sum().item(), v.double().abs().sum().item()) for k, v in sd0.items()
                            if v.dtype.is_floating_point},
        "pred_real": [s[-1].detach() for s in d_real], "pred_fake": [s[-1].detach() for s in d_fake],
        "feat_real_scale0_layer1": d_real[1][1].detach()[:, :4],     # result[1] is scale 0 (the pooled input)
        "d_loss": d_loss.detach(), "g_loss": g_loss.detach(),
        "grad_norms": {k: v.norm() for k, v in grads.items()},
        "grad_slices": {k: v.flatten()[:64].clone() for k, v in grads.items()},
        "stats_after": {k: v for k, v in sd1.items() if "running" in k or "num_batches" in k},
        "grad_x_fake": xf.grad.clone(),
    }


def main():
    out = {"img": run("img"), "vid": run("vid")}
    path = os.path.join(HERE, "golden_disc.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB; oracle == reference on both discriminators")


if __name__ == "__main__":
    main()
