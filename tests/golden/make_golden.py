"""Generate golden fixtures from the UNMODIFIED reference modules (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference's models.vqvae_conv3d_latent / models.lpips / loss with the two shims of
SURVEY.md section 8(c) (matplotlib stub; LPIPS weight download replaced by seeded random weights), runs them on
seeded inputs/weights, and writes the REFERENCE's outputs to tests/golden/*.pt.  The oracle
(oracle/faceoff_oracle.py) and the CUDA path are then both checked against these files, on any box.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import models.vqvae_conv3d_latent as ref_vq  # noqa
    import models.lpips as ref_lpips  # noqa
    import torchvision

    _orig_vgg16 = torchvision.models.vgg16
    shim = types.SimpleNamespace(vgg16=lambda pretrained=True, **kw: _orig_vgg16(weights=None))
    ref_lpips.models = shim  # only the name looked up at models/lpips.py:118 is replaced
    ref_lpips.LPIPS.load_from_pretrained = lambda self, name="vgg_lpips": None
    # loss.py imports models.discriminator (pure torch) - fine
    import loss as ref_loss  # noqa
    return ref_vq, ref_lpips, ref_loss


def ref_batched_forward(model, img, n_clips):
    """B>1 clips through the reference's own sub-methods."""
    F_ = img.shape[0]
    T = F_ // n_clips
    enc_b, enc_t = model.only_encode(img)

    def to5(x):
        return x.reshape(n_clips, T, *x.shape[1:]).permute(0, 2, 1, 3, 4)

    def to4(x):
        return x.permute(0, 2, 1, 3, 4).reshape(F_, x.shape[1], *x.shape[3:])

    eb = to4(model.conv3d_encoded_b(to5(enc_b)))
    et = to4(model.conv3d_encoded_t(to5(enc_t)))
    quant_t, quant_b, diff, id_t, id_b = model.encode_quantized(eb, et)
    dec = model.decode(quant_t, quant_b)
    return dec, diff, id_t, id_b


def main():
    from oracle import faceoff_oracle as O

    ref_vq, ref_lpips, ref_loss = import_reference()
    torch.manual_seed(0)
    torch.set_num_threads(8)
    out = {}

    # ---- 1. Quantize alone (reference Quantize, training mode), two sizes -------------------
    for tag, (N, D, K) in {"q_small": (1024, 64, 512), "q_d128": (512, 128, 1024)}.items():
        g = torch.Generator().manual_seed(7)
        x = torch.randn(4, N // 4, D, generator=g)
        qm = ref_vq.Quantize(D, K)
        qm.embed.copy_(torch.randn(D, K, generator=g))
        qm.embed_avg.copy_(qm.embed)
        qm.cluster_size.copy_(torch.rand(K, generator=g) * 3)
        e0, c0, a0 = qm.embed.clone(), qm.cluster_size.clone(), qm.embed_avg.clone()
        qm.train()
        xin = x.clone().requires_grad_(True)
        q, diff, ind = qm(xin)
        gq = torch.randn(q.shape, generator=g)
        (q * gq).sum().add(diff * 3.0).backward()
        out[tag] = dict(x=x, embed0=e0, cluster_size0=c0, embed_avg0=a0, gq=gq, quantize=q.detach(),
                        diff=diff.detach(), embed_ind=ind, grad_x=xin.grad, embed1=qm.embed.clone(),
                        cluster_size1=qm.cluster_size.clone(), embed_avg1=qm.embed_avg.clone())
        # oracle must agree
        oq, od, oi, nb, _ = O.quantize_forward(x, e0, c0, a0, True)
        assert torch.equal(oi, ind), tag
        assert torch.allclose(oq, q.detach()) and torch.allclose(od, diff.detach())
        assert torch.allclose(nb[0], qm.embed, rtol=1e-6, atol=1e-7)

    # ---- 2. VQVAE(in_channel=6) train step, no LPIPS (train_faceoff.py:31-44,140-142) ---------
    def vq_case(tag, n_clips, T, H, W, with_lpips):
        p = O.init_vqvae_params(seed=0)
        model = ref_vq.VQVAE(in_channel=6)
        missing = model.load_state_dict(p, strict=True)
        model.train()
        img, gt = O.synthetic_clip(n_clips, T, H, W, seed=1234)
        lp = None
        if with_lpips:
            lp = O.init_lpips_params(seed=1)
            vql = ref_loss.VQLPIPS()
            sd = {"perceptual_loss." + k: v for k, v in lp.items()}
            vql.load_state_dict(sd, strict=True)
        model.zero_grad()
        if n_clips == 1:
            dec, diff = model(img)
            id_t = id_b = None
        else:
            dec, diff, id_t, id_b = ref_batched_forward(model, img, n_clips)
        rec = dec[:, :3]
        recon = torch.nn.MSELoss()(rec, gt)
        latent = diff.mean()
        loss = recon + latent
        perc = None
        if with_lpips:
            perc = vql(gt, rec)
            loss = loss + perc
        loss.backward()
        grads = {k: v.grad.clone() for k, v in model.named_parameters()}
        # oracle on the same state
        o = O.train_step(p, img, gt, n_clips=n_clips, lp=lp)
        o64 = O.train_step(p, img, gt, n_clips=n_clips, lp=lp, dtype=torch.float64)
        assert torch.allclose(o["dec"], dec.detach(), rtol=1e-4, atol=1e-5), tag
        assert torch.allclose(o["loss"], loss.detach(), rtol=1e-5), (tag, o["loss"], loss)
        if id_t is not None:
            assert torch.equal(o["id_t"], id_t) and torch.equal(o["id_b"], id_b)
        worst = 0.0
        for k, gref in grads.items():
            denom = gref.abs().max().item() + 1e-12
            worst = max(worst, (o["grads"][k] - gref).abs().max().item() / denom)
        print(f"[{tag}] oracle-vs-reference worst max-normalised grad err: {worst:.3e}")
        assert worst < 2e-3, worst
        for q in ("quantize_t", "quantize_b"):
            for i, name in enumerate(("embed", "cluster_size", "embed_avg")):
                refbuf = getattr(getattr(model, q), name)
                assert torch.allclose(o["new_buffers"][q][i], refbuf, rtol=1e-5, atol=1e-6), (tag, q, name)
        keep = ["enc_b.blocks.0.weight", "enc_b.blocks.0.bias", "enc_b.blocks.5.conv.1.weight",
                "conv3d_encoded_b.conv3d.0.0.weight", "conv3d_encoded_t.conv3d.2.0.bias",
                "quantize_conv_b.weight", "dec_t.blocks.4.weight", "dec.blocks.6.weight", "dec.blocks.6.bias",
                "upsample_t.weight"]
        out[tag] = dict(
            cfg=dict(n_clips=n_clips, T=T, H=H, W=W, with_lpips=with_lpips, seed_params=0, seed_data=1234,
                     seed_lpips=1),
            loss=loss.detach(), recon_loss=recon.detach(), latent_loss=latent.detach(),
            perceptual_loss=None if perc is None else perc.detach(),
            dec_sample=dec.detach()[:, :, ::8, ::8].clone(), dec_full0=dec.detach()[0].clone(), dec_mean=dec.detach().mean(), dec_std=dec.detach().std(),
            id_t=o["id_t"].to(torch.int16), id_b=o["id_b"].to(torch.int16),
            grads_ref={k: (grads[k] if grads[k].numel() <= 8192 else grads[k][:8, :8].clone()) for k in keep},
            grad_norms_ref={k: v.norm() for k, v in grads.items()},
            grads_f64_norms={k: v.norm().float() for k, v in o64["grads"].items()},
            buffers_ref={f"{q}.{n}": getattr(getattr(model, q), n).clone() for q in ("quantize_t", "quantize_b")
                         for n in ("embed", "cluster_size", "embed_avg")},
        )

    vq_case("vqvae_1x4x64", 1, 4, 64, 64, False)
    vq_case("vqvae_2x3x64_lpips", 2, 3, 64, 64, True)

    # ---- 3. LPIPS alone -------------------------------------------------------------------------
    lp = O.init_lpips_params(seed=1)
    lm = ref_lpips.LPIPS()
    lm.load_state_dict(lp, strict=True)
    lm.eval()
    g = torch.Generator().manual_seed(5)
    a = (torch.rand(3, 3, 64, 64, generator=g) * 2 - 1)
    b = (a + 0.3 * torch.randn(3, 3, 64, 64, generator=g)).clamp(-1, 1)
    bb = b.clone().requires_grad_(True)
    val = lm(a, bb)
    val.mean().backward()
    ov = O.lpips_forward(lp, a, b)
    assert torch.allclose(ov, val.detach(), rtol=1e-5, atol=1e-7)
    out["lpips_3x64"] = dict(a=a, b=b, val=val.detach(), grad_b=bb.grad.clone())

    torch.save(out, os.path.join(HERE, "golden.pt"))
    sz = os.path.getsize(os.path.join(HERE, "golden.pt"))
    print("wrote golden.pt", sz, "bytes")


if __name__ == "__main__":
    main()
