"""GPU parity tests (run on the B200 box: ``pytest -m gpu``).  Everything goes through the C ABI
(faceoff_b200/libfaceoff_b200.so) and is compared with (a) the committed golden vectors produced by the
UNMODIFIED reference (tests/golden/golden.pt) and (b) the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): fp32 path (Quantize) rtol 1e-5, indices bit-exact outside near-ties
(relative gap < 1e-6); bf16 conv path rtol 2e-2 / atol 1e-2, checked as max-normalised error per tensor.
"""
import os
import sys
import warnings

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "golden.pt")

BF16_RTOL, BF16_ATOL = 2e-2, 1e-2
# end-to-end bf16 gradient bounds above the default 2e-2 (measured 6.4e-2 / 9.8e-2 / 3.0e-2): the first encoder layer sits
# behind ~25 bf16-stored activation gradients, upsample_t / the last transposed conv sum ~1e5 products of similar size
BF16_GRAD_ALLOW = {"enc_b.blocks.0.weight": 0.1, "upsample_t.weight": 0.15, "dec.blocks.6.weight": 0.05}


def _golden():
    return torch.load(GOLDEN, map_location="cpu")


def _near_tie_rows(x, embed, thresh=1e-6):
    """rows whose best / second-best distance gap is < thresh relative (fp64) -- excluded per north_star."""
    x = x.double().reshape(-1, embed.shape[0])
    d = x.pow(2).sum(1, keepdim=True) - 2 * x @ embed.double() + embed.double().pow(2).sum(0, keepdim=True)
    s = d.sort(1).values
    return ((s[:, 1] - s[:, 0]) / s[:, 0].abs()) < thresh


def maxnorm_err(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12)).item()


@pytest.mark.parametrize("tag", ["q_small", "q_d128"])
def test_quantize_matches_reference_golden(tag):
    from faceoff_b200.vqvae import Quantize

    g = _golden()[tag]
    D, K = g["embed0"].shape
    q = Quantize(D, K).cuda()
    q.embed.copy_(g["embed0"])
    q.embed_avg.copy_(g["embed_avg0"])
    q.cluster_size.copy_(g["cluster_size0"])
    q.train()
    q.n_flagged = torch.zeros(1, dtype=torch.int32, device="cuda")
    x = g["x"].cuda().requires_grad_(True)
    quant, diff, ind = q(x)
    (quant * g["gq"].cuda()).sum().add(diff * 3.0).backward()
    torch.cuda.synchronize()
    assert ind.dtype == torch.int64 and ind.shape == g["embed_ind"].shape
    tie = _near_tie_rows(g["x"], g["embed0"]).reshape(ind.shape)
    mism = (ind.cpu() != g["embed_ind"]) & ~tie
    assert mism.sum().item() == 0, f"{mism.sum().item()} index mismatches outside near-ties ({tie.sum().item()} near-ties)"
    torch.testing.assert_close(quant.detach().cpu(), g["quantize"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(diff.detach().cpu(), g["diff"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(x.grad.cpu(), g["grad_x"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(q.cluster_size.cpu(), g["cluster_size1"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(q.embed_avg.cpu(), g["embed_avg1"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(q.embed.cpu(), g["embed1"], rtol=1e-5, atol=1e-6)
    print(f"{tag}: rows re-evaluated exactly: {q.n_flagged.item()} / {ind.numel()}")


def test_quantize_eval_mode_leaves_buffers_untouched():
    from faceoff_b200.vqvae import Quantize

    torch.manual_seed(0)
    q = Quantize(64, 512).cuda().eval()
    e0, c0 = q.embed.clone(), q.cluster_size.clone()
    out, diff, ind = q(torch.randn(2, 8, 8, 64, device="cuda"))
    torch.cuda.synchronize()
    assert torch.equal(q.embed, e0) and torch.equal(q.cluster_size, c0)
    # quantize values are codebook rows (straight-through evaluated in fp32: x + (q - x))
    ref = q.embed_code(ind)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("rows,dim,K", [(122880, 64, 512), (30720, 128, 2048), (1, 64, 512), (129, 64, 16)])
def test_vq_assign_bit_exact_vs_fp64(rows, dim, K):
    """Full config-4 style sizes: indices must equal the fp64 argmin (first minimum) outside near-ties."""
    from faceoff_b200 import ops

    gen = torch.Generator().manual_seed(11)
    x = torch.randn(rows, dim, generator=gen)
    e = torch.randn(dim, K, generator=gen)
    xs, es = x.cuda(), e.cuda()
    e_split, e_t, e_n2 = ops.vq_prep(es)
    ind = ops.vq_assign(xs, e_t, e_split, e_n2).cpu()
    d = x.double().pow(2).sum(1, keepdim=True) - 2 * x.double() @ e.double() + e.double().pow(2).sum(0, keepdim=True)
    ref = d.argmin(1)
    assert torch.equal(ind, ref), f"{(ind != ref).sum().item()} mismatches"


@pytest.mark.parametrize("dim,K", [(32, 100), (96, 40), (64, 50), (256, 24)])
def test_quantize_any_codebook_shape_vs_oracle(dim, K):
    """The reference's Quantize takes any (dim, n_embed) (models/vqvae_conv3d_latent.py:34-45).  Shapes outside the
    tensor-core kernel's (dim 64 / 128, n_embed % 16 == 0) run the exact fp64 scan (vq.cu vq_assign_generic_kernel): indices
    bit-exact vs the fp64 argmin, outputs / EMA buffers / input gradient vs the oracle on the same seeded input."""
    from faceoff_b200.vqvae import Quantize
    from oracle import faceoff_oracle as O

    gen = torch.Generator().manual_seed(5)
    x = torch.randn(3, 7, 5, dim, generator=gen)
    q = Quantize(dim, K)
    with torch.no_grad():
        q.embed.copy_(torch.randn(dim, K, generator=gen))
        q.embed_avg.copy_(q.embed)
        q.cluster_size.copy_(torch.rand(K, generator=gen))
    e0, c0, a0 = q.embed.clone(), q.cluster_size.clone(), q.embed_avg.clone()
    gq = torch.randn(3, 7, 5, dim, generator=gen)
    q = q.cuda().train()
    xg = x.cuda().requires_grad_(True)
    quant, diff, ind = q(xg)
    (quant * gq.cuda()).sum().add(diff * 3.0).backward()
    torch.cuda.synchronize()
    xo = x.clone().requires_grad_(True)
    oq, odiff, oind, obuf, _ = O.quantize_forward(xo, e0, c0, a0, training=True)
    (oq * gq).sum().add(odiff * 3.0).backward()
    d = (x.double().reshape(-1, dim).pow(2).sum(1, keepdim=True) - 2 * x.double().reshape(-1, dim) @ e0.double()
         + e0.double().pow(2).sum(0, keepdim=True))
    assert torch.equal(ind.cpu().reshape(-1), d.argmin(1)), "indices differ from the fp64 argmin"
    assert torch.equal(ind.cpu(), oind)
    torch.testing.assert_close(quant.detach().cpu(), oq.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(diff.detach().cpu(), odiff.detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(xg.grad.cpu(), xo.grad, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(q.embed.cpu(), obuf[0], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(q.cluster_size.cpu(), obuf[1], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(q.embed_avg.cpu(), obuf[2], rtol=1e-5, atol=1e-6)


def _load_vqvae(p):
    from faceoff_b200.vqvae import VQVAE

    m = VQVAE(in_channel=6)
    m.load_state_dict(p, strict=True)
    return m.cuda().train()


@pytest.mark.parametrize("tag", ["vqvae_1x4x64", "vqvae_2x3x64_lpips"])
def test_vqvae_train_step_matches_reference_golden(tag):
    from oracle import faceoff_oracle as O

    g = _golden()[tag]
    cfg = g["cfg"]
    p = O.init_vqvae_params(seed=cfg["seed_params"])
    img, gt = O.synthetic_clip(cfg["n_clips"], cfg["T"], cfg["H"], cfg["W"], seed=cfg["seed_data"])
    model = _load_vqvae(p)
    lp = None
    if cfg["with_lpips"]:
        from faceoff_b200.lpips import VQLPIPS

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            vql = VQLPIPS()
        lp = O.init_lpips_params(seed=cfg["seed_lpips"])
        vql.load_state_dict({"perceptual_loss." + k: v for k, v in lp.items()}, strict=True)
        vql = vql.cuda()
    from op_verifier import OpVerifier

    x = img.cuda()
    with OpVerifier() as ver:
        dec, diff, id_t, id_b = model.forward_with_ids(x, clips=cfg["n_clips"])
        rec = dec[:, :3]
        recon = torch.nn.functional.mse_loss(rec, gt.cuda())
        latent = diff.mean()
        loss = recon + latent
        if lp is not None:
            perc = vql(gt.cuda(), rec)
            loss = loss + perc
        loss.backward()
    torch.cuda.synchronize()
    # every tensor-core launch of the PRODUCT (bf16) path against torch fp64 on the operands it actually read: conv outputs
    # are one bf16 rounding away (2^-8 relative to the largest element), weight / bias gradients (fp32 outputs) 2e-5
    worst_conv = max((max(e.values()), d) for k, d, e in ver.records if k == "conv")
    worst_wg = max((max(e.values()), d) for k, d, e in ver.records if k == "wgrad")
    print(f"{tag}: per-launch verifier (bf16 product path): {len(ver.records)} launches; worst conv {worst_conv[0]:.1e} "
          f"({worst_conv[1]}); worst wgrad {worst_wg[0]:.1e} ({worst_wg[1]})")
    assert worst_conv[0] <= 2.0 ** -7 and worst_wg[0] <= 2e-5, (worst_conv, worst_wg)

    # indices: the conv stack is bf16, so a few rows may legitimately flip; they must agree almost everywhere
    agree_t = (id_t.cpu() == g["id_t"].long()).float().mean().item()
    agree_b = (id_b.cpu() == g["id_b"].long()).float().mean().item()
    print(f"{tag}: id agreement top {agree_t:.4f} bottom {agree_b:.4f}")
    assert agree_t > 0.98 and agree_b > 0.98
    assert maxnorm_err(dec[0].cpu(), g["dec_full0"]) < 3e-2
    assert abs(recon.item() - g["recon_loss"].item()) <= BF16_RTOL * abs(g["recon_loss"].item()) + BF16_ATOL
    assert abs(latent.item() - g["latent_loss"].item()) <= BF16_RTOL * abs(g["latent_loss"].item()) + BF16_ATOL
    if lp is not None:
        assert abs(perc.item() - g["perceptual_loss"].item()) <= BF16_RTOL * abs(g["perceptual_loss"].item()) + BF16_ATOL
    # gradients: norm agreement for every parameter + max-normalised error on the stored slices
    grads = {k: v.grad for k, v in model.named_parameters()}
    worst = 0.0
    for k, nref in g["grad_norms_ref"].items():
        assert grads[k] is not None, k
        n = grads[k].norm().item()
        rel = abs(n - nref.item()) / (nref.item() + 1e-12)
        worst = max(worst, rel)
        assert rel < 0.1, (k, n, nref.item())
    for k, gref in g["grads_ref"].items():
        got = grads[k].cpu()
        got = got if got.numel() == gref.numel() else got[:8, :8]
        # per-tensor relative bound (max-normalised 2e-2 = the north_star bf16 rtol); explicit allow-list for the layers
        # whose gradient accumulates the bf16 rounding of the longest chains of stored activation gradients
        err = maxnorm_err(got, gref)
        bound = BF16_GRAD_ALLOW.get(k, BF16_RTOL)
        print(f"  grad {k}: max-normalised err {err:.3e} (bound {bound:g})")
        assert err < bound, (k, err)
    print(f"{tag}: worst grad-norm rel err {worst:.3e}")
    # EMA codebooks
    for k, bref in g["buffers_ref"].items():
        mod, name = k.split(".")
        got = getattr(getattr(model, mod), name).cpu()
        assert maxnorm_err(got, bref) < 2e-2, k


def test_vqvae_forward_signature_and_eval():
    from oracle import faceoff_oracle as O

    p = O.init_vqvae_params(seed=0)
    model = _load_vqvae(p).eval()
    img, _ = O.synthetic_clip(1, 2, 64, 64)
    with torch.no_grad():
        dec, diff = model(img.cuda())
    assert dec.shape == (2, 6, 64, 64) and dec.dtype == torch.float32 and diff.shape == (1,)
    ref = O.vqvae_forward(p, img, n_clips=1, training=False)
    assert maxnorm_err(dec.cpu(), ref["dec"]) < 3e-2
    # eval must not touch the codebooks
    torch.testing.assert_close(model.quantize_t.embed.cpu(), p["quantize_t.embed"])
    # 5-D (batched clips) entry point
    img2, _ = O.synthetic_clip(2, 2, 64, 64)
    with torch.no_grad():
        dec2, _ = model(img2.cuda().view(2, 2, 6, 64, 64))
    assert dec2.shape == (2, 2, 6, 64, 64)


def test_lpips_matches_reference_golden():
    from faceoff_b200.lpips import LPIPS
    from oracle import faceoff_oracle as O

    g = _golden()["lpips_3x64"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = LPIPS()
    m.load_state_dict(O.init_lpips_params(seed=1), strict=True)
    m = m.cuda().eval()
    a = g["a"].cuda()
    b = g["b"].cuda().requires_grad_(True)
    val = m(a, b)
    val.mean().backward()
    torch.cuda.synchronize()
    assert val.shape == (3, 1, 1, 1)
    assert maxnorm_err(val.detach().cpu(), g["val"]) < BF16_RTOL * 2.5
    # The input gradient of a random-weight VGG16 is ill-conditioned (max-pool routing flips under rounding): an
    # idealised bf16-storage emulation in torch shows the same 13% norm-wise error as this implementation, and even
    # cuDNN fp32 vs the CPU golden differs by 4% (tests/gpu_diag.py::lpips_grad).  Check direction + norm.
    gb, gr = b.grad.cpu().flatten(), g["grad_b"].flatten()
    cos = torch.nn.functional.cosine_similarity(gb, gr, dim=0).item()
    nw = ((gb - gr).norm() / gr.norm()).item()
    print(f"lpips input grad: cosine {cos:.5f} norm-wise rel err {nw:.3e}")
    assert cos > 0.98 and nw < 0.2
    # symmetric in value (reference trainer passes (ground_truth, out))
    val2 = m(b.detach(), a)
    assert maxnorm_err(val2.cpu(), val.detach().cpu()) < 2e-2


def test_submodules_drop_in():
    """Encoder / Decoder / ResBlock / Conv3dLatentPostnet stand-alone against the oracle's functions."""
    from faceoff_b200.vqvae import Conv3dLatentPostnet, Decoder, Encoder, ResBlock
    from oracle import faceoff_oracle as O
    import torch.nn.functional as F

    torch.manual_seed(0)
    enc = Encoder(6, 128, 2, 32, 4).cuda()
    x = torch.rand(2, 6, 64, 64, device="cuda") * 2 - 1
    y = enc(x)
    p = {"e." + k: v.detach().cpu() for k, v in enc.state_dict().items()}
    ref = O.encoder(p, "e", x.cpu(), 4)
    assert y.shape == ref.shape and maxnorm_err(y.cpu(), ref) < 3e-2
    dec = Decoder(64, 64, 128, 2, 32, 2).cuda()
    z = torch.randn(2, 64, 8, 8, device="cuda", requires_grad=True)
    out = dec(z)
    p = {"d." + k: v.detach().cpu() for k, v in dec.state_dict().items()}
    zc = z.detach().cpu().requires_grad_(True)
    ref = O.decoder(p, "d", zc, 2)
    assert maxnorm_err(out.detach().cpu(), ref) < 3e-2
    go = torch.randn_like(out)
    out.backward(go)
    ref.backward(go.cpu())
    assert maxnorm_err(z.grad.cpu(), zc.grad) < 5e-2
    rb = ResBlock(128, 32).cuda()
    h = torch.randn(2, 128, 16, 16, device="cuda")
    p = {"r." + k: v.detach().cpu() for k, v in rb.state_dict().items()}
    assert maxnorm_err(rb(h).cpu(), O.resblock(p, "r", h.cpu())) < 3e-2
    c3 = Conv3dLatentPostnet(128).cuda()
    v = torch.randn(1, 128, 3, 8, 8, device="cuda")
    p = {"c." + k: t.detach().cpu() for k, t in c3.state_dict().items()}
    assert maxnorm_err(c3(v).cpu(), O.conv3d_postnet(p, "c", v.cpu())) < 3e-2


def test_clip_independence_property():
    """Size-independent property: a batch of B clips == B single-clip calls (Conv3d must not mix clips)."""
    from oracle import faceoff_oracle as O

    p = O.init_vqvae_params(seed=0)
    model = _load_vqvae(p).eval()
    img, _ = O.synthetic_clip(2, 3, 64, 64, seed=3)
    with torch.no_grad():
        both, _ = model(img.cuda().view(2, 3, 6, 64, 64))
        one0, _ = model(img[:3].cuda())
        one1, _ = model(img[3:].cuda())
    torch.testing.assert_close(both[0], one0, rtol=0, atol=0)
    torch.testing.assert_close(both[1], one1, rtol=0, atol=0)


def test_fused_dp_wrapper_single_rank_and_grad_accumulation():
    """FusedDataParallel at world_size 1 must be a transparent wrapper; two backward passes without zero_grad must add."""
    from faceoff_b200.parallel import FusedDataParallel
    from oracle import faceoff_oracle as O

    p = O.init_vqvae_params(seed=0)
    img, gt = O.synthetic_clip(1, 3, 64, 64, seed=21)

    def step(net, model, zero=True):
        if zero:
            model.zero_grad(set_to_none=True)
        out, latent = net(img.cuda())
        (torch.nn.functional.mse_loss(out[:, :3], gt.cuda()) + latent.mean()).backward()

    plain = _load_vqvae(p)
    step(plain, plain)
    g_plain = {k: v.grad.clone() for k, v in plain.named_parameters()}
    wrapped = _load_vqvae(p)
    ddp = FusedDataParallel(wrapped)
    assert ddp.module is wrapped
    step(ddp, wrapped)
    for k, v in wrapped.named_parameters():
        torch.testing.assert_close(v.grad, g_plain[k], rtol=1e-5, atol=1e-8)
    # accumulation: eval-mode codebooks stay fixed so the two passes are identical => grads double
    acc = _load_vqvae(p)
    acc.quantize_t.eval()
    acc.quantize_b.eval()
    step(acc, acc)
    g1 = {k: v.grad.clone() for k, v in acc.named_parameters()}
    step(acc, acc, zero=False)
    for k, v in acc.named_parameters():
        torch.testing.assert_close(v.grad, 2 * g1[k], rtol=1e-4, atol=1e-7)


def test_quantize_with_dead_codes_of_huge_norm():
    """The EMA renormalisation blows unused codes up by ~1e5 (reference :70-75); the tensor-core error band is per code,
    so such codebooks must neither break bit-exactness nor send every row to the exact re-check."""
    from faceoff_b200 import ops

    gen = torch.Generator().manual_seed(3)
    x = torch.randn(20000, 64, generator=gen) * 0.3
    e = torch.randn(64, 512, generator=gen)
    e[:, 100:] *= 3e5          # 412 dead codes
    xs, es = x.cuda(), e.cuda()
    e_split, e_t, e_n2 = ops.vq_prep(es)
    nf = torch.zeros(1, dtype=torch.int32, device="cuda")
    ind = ops.vq_assign(xs, e_t, e_split, e_n2, nf).cpu()
    d = x.double().pow(2).sum(1, keepdim=True) - 2 * x.double() @ e.double() + e.double().pow(2).sum(0, keepdim=True)
    assert torch.equal(ind, d.argmin(1))
    assert nf.item() < 0.05 * x.shape[0], f"{nf.item()} rows re-checked"


@pytest.mark.parametrize("T,res,cin", [(5, 64, 6), (2, 128, 3)])
def test_vqvae_odd_shapes(T, res, cin):
    from faceoff_b200.vqvae import VQVAE
    from oracle import faceoff_oracle as O

    p = O.init_vqvae_params(seed=2, in_channel=cin)
    m = VQVAE(in_channel=cin)
    m.load_state_dict(p)
    m = m.cuda().train()
    g = torch.Generator().manual_seed(1)
    img = torch.rand(T, cin, res, res, generator=g) * 2 - 1
    dec, diff = m(img.cuda())
    ref = O.vqvae_forward(p, img, n_clips=1, training=True)
    assert dec.shape == (T, cin, res, res)
    assert maxnorm_err(dec.cpu(), ref["dec"]) < 3e-2
    assert abs(diff.item() - ref["diff"].item()) <= 2e-2 * abs(ref["diff"].item()) + 1e-2
    dec.square().mean().backward()
    assert all(q.grad is not None and torch.isfinite(q.grad).all() for q in m.parameters())


def test_full_size_clip_train_step_vs_oracle():
    """BASELINE configs[0] size: one 30-frame 256x256 clip, fwd+bwd, against the CPU oracle run live (a few seconds of
    host time).  Losses within the bf16 tolerance, indices almost everywhere equal, every gradient norm within 10 %."""
    from oracle import faceoff_oracle as O

    p = O.init_vqvae_params(seed=0)
    img, gt = O.synthetic_clip(1, 30, 256, 256, seed=1234)
    model = _load_vqvae(p)
    dec, diff, id_t, id_b = model.forward_with_ids(img.cuda(), clips=1)
    recon = torch.nn.functional.mse_loss(dec[:, :3], gt.cuda())
    (recon + diff.mean()).backward()
    torch.cuda.synchronize()
    torch.set_num_threads(os.cpu_count() or 1)
    o = O.train_step(p, img, gt, n_clips=1)
    assert abs(recon.item() - o["recon_loss"].item()) <= BF16_RTOL * abs(o["recon_loss"].item()) + BF16_ATOL
    assert abs(diff.mean().item() - o["latent_loss"].item()) <= BF16_RTOL * abs(o["latent_loss"].item()) + BF16_ATOL
    agree_t = (id_t.cpu() == o["id_t"]).float().mean().item()
    agree_b = (id_b.cpu() == o["id_b"]).float().mean().item()
    # a code flip (bf16 conv noise on a near-tie) changes the decoded patch, so dec is compared norm-wise and by the
    # fraction of pixels inside the bf16 tolerance rather than by its maximum
    d, r = dec.detach().cpu(), o["dec"]
    nw = ((d - r).norm() / r.norm()).item()
    inside = ((d - r).abs() <= BF16_ATOL + BF16_RTOL * r.abs()).float().mean().item()
    print(f"full-size clip: id agreement top {agree_t:.4f} bottom {agree_b:.4f}; dec norm-wise err {nw:.3e}, "
          f"{100 * inside:.2f}% of pixels inside rtol 2e-2 / atol 1e-2")
    assert agree_t > 0.98 and agree_b > 0.98
    assert nw < 5e-2 and inside > 0.98
    worst = 0.0
    for k, v in model.named_parameters():
        nref = o["grads"][k].norm().item()
        rel = abs(v.grad.norm().item() - nref) / (nref + 1e-12)
        worst = max(worst, rel)
    print(f"full-size clip: worst grad-norm rel err {worst:.3e}")
    assert worst < 0.1
    for q in ("quantize_t", "quantize_b"):
        for i, name in enumerate(("embed", "cluster_size", "embed_avg")):
            assert maxnorm_err(getattr(getattr(model, q), name).cpu(), o["new_buffers"][q][i]) < 2e-2, (q, name)


def test_three_optimizer_steps_track_the_oracle():
    """Adam steps mutate the weights in place: the packed-weight cache must follow (tensor version), the EMA codebooks
    must evolve step by step like the reference.  Losses of 3 consecutive steps vs the oracle run with the same optimizer."""
    from oracle import faceoff_oracle as O

    p0 = O.init_vqvae_params(seed=0)
    img, gt = O.synthetic_clip(1, 2, 64, 64, seed=11)
    # oracle side: functional train_step + torch Adam on a dict of leaf tensors
    p = {k: v.clone() for k, v in p0.items()}
    keys = O.trainable_keys(p)
    leaves = [p[k].requires_grad_(True) for k in keys]
    opt_ref = torch.optim.Adam(leaves, lr=3e-3)
    ref_losses = []
    for _ in range(3):
        o = O.train_step({k: v.detach() for k, v in p.items()}, img, gt, n_clips=1)
        ref_losses.append(o["loss"].item())
        for k, leaf in zip(keys, leaves):
            leaf.grad = o["grads"][k].clone()
        opt_ref.step()
        for q in ("quantize_t", "quantize_b"):
            for i, name in enumerate(("embed", "cluster_size", "embed_avg")):
                p[f"{q}.{name}"] = o["new_buffers"][q][i]
    model = _load_vqvae(p0)
    opt = torch.optim.Adam(model.parameters(), lr=3e-3)
    losses = []
    for _ in range(3):
        model.zero_grad()
        out, latent = model(img.cuda())
        loss = torch.nn.functional.mse_loss(out[:, :3], gt.cuda()) + latent.mean()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    print("losses ours", losses, "oracle", ref_losses)
    assert ref_losses[2] != ref_losses[0]
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 3e-2 * abs(b) + 1e-2


@pytest.mark.parametrize("n,ca,c,h,w", [(5, 3, 3, 64, 64), (3, 6, 3, 32, 16), (1, 3, 3, 2, 2)])
def test_fused_mse_loss_matches_torch(n, ca, c, h, w):
    """mse_loss(out, gt) == nn.MSELoss()(out[:, :c], gt) (reference train_faceoff_perceptual.py:38-40), value and grad
    (fp32: rtol 1e-5; the upstream gradient is a device scalar, here 0.37)."""
    from faceoff_b200.losses import mse_loss

    torch.manual_seed(0)
    out = torch.randn(n, ca, h, w, device="cuda", requires_grad=True)
    gt = torch.randn(n, c, h, w, device="cuda")
    loss = mse_loss(out, gt)
    (loss * 0.37).backward()
    ref_in = out.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.mse_loss(ref_in[:, :c], gt)
    (ref * 0.37).backward()
    torch.testing.assert_close(loss, ref, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(out.grad, ref_in.grad, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_fused_adam_matches_torch_adam(wd):
    """FusedAdam == torch.optim.Adam (reference train_faceoff_perceptual.py:190, lr 3e-4) over 4 steps on
    tensors of awkward sizes (one launch for all of them); fp32, rtol 1e-6.  State dicts are interchangeable."""
    from faceoff_b200.optim import FusedAdam

    torch.manual_seed(0)
    shapes = [(128, 128, 3, 3, 3), (64,), (3, 5, 7), (16385,), (1,)]
    ours = [torch.randn(*s, device="cuda").requires_grad_(True) for s in shapes]
    ref = [p.detach().clone().requires_grad_(True) for p in ours]
    opt = FusedAdam(ours, lr=3e-4, weight_decay=wd)
    opt_ref = torch.optim.Adam(ref, lr=3e-4, weight_decay=wd)
    for it in range(4):
        for p, r in zip(ours, ref):
            g = torch.randn_like(p) * (10.0 ** (it - 2))
            p.grad = g.clone()
            r.grad = g.clone()
        opt.step()
        opt_ref.step()
    for p, r in zip(ours, ref):
        torch.testing.assert_close(p, r, rtol=1e-6, atol=1e-7)
    sd = opt.state_dict()
    opt_ref.load_state_dict(sd)            # same state names / shapes
    for k in ("step", "exp_avg", "exp_avg_sq"):
        assert k in sd["state"][0]


def test_three_fused_adam_steps_track_the_oracle():
    """FusedAdam updates the parameters through raw pointers; the packed bf16 weights cached per (tensor, version) must
    be re-packed after every step (advisor finding r1: stale weights).  Same protocol as the torch.optim.Adam test."""
    from faceoff_b200.optim import FusedAdam
    from oracle import faceoff_oracle as O

    p0 = O.init_vqvae_params(seed=0)
    img, gt = O.synthetic_clip(1, 2, 64, 64, seed=11)
    p = {k: v.clone() for k, v in p0.items()}
    keys = O.trainable_keys(p)
    leaves = [p[k].requires_grad_(True) for k in keys]
    opt_ref = torch.optim.Adam(leaves, lr=3e-3)
    ref_losses = []
    for _ in range(3):
        o = O.train_step({k: v.detach() for k, v in p.items()}, img, gt, n_clips=1)
        ref_losses.append(o["loss"].item())
        for k, leaf in zip(keys, leaves):
            leaf.grad = o["grads"][k].clone()
        opt_ref.step()
        for q in ("quantize_t", "quantize_b"):
            for i, name in enumerate(("embed", "cluster_size", "embed_avg")):
                p[f"{q}.{name}"] = o["new_buffers"][q][i]
    model = _load_vqvae(p0)
    opt = FusedAdam(model.parameters(), lr=3e-3)
    losses = []
    for _ in range(3):
        model.zero_grad()
        out, latent = model(img.cuda())
        loss = torch.nn.functional.mse_loss(out[:, :3], gt.cuda()) + latent.mean()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    print("losses ours (FusedAdam)", losses, "oracle", ref_losses)
    assert ref_losses[2] != ref_losses[0]
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 3e-2 * abs(b) + 1e-2
    # The weights must have MOVED like the oracle's.  (Adam normalises each element's step to ~lr * sign(g), so elements
    # whose gradient is near zero may legitimately step the other way under bf16 noise: compare the update as a whole.)
    params = dict(model.named_parameters())
    for k, leaf in zip(keys, leaves):
        d_ours = (params[k].detach().cpu() - p0[k]).flatten().double()
        d_ref = (leaf.detach() - p0[k]).flatten().double()
        cos = torch.nn.functional.cosine_similarity(d_ours, d_ref, dim=0).item()
        ratio = (d_ours.norm() / d_ref.norm()).item()
        assert cos > 0.8 and 0.85 < ratio < 1.15, (k, cos, ratio)


def test_embed_ind_bit_exact_on_real_activations_after_training():
    """north_star: codebook indices bit-exact on the fp32 path.  The quantiser input here is the REAL pre-quantiser
    activation (fp32 output of quantize_conv_{t,b} after 3 optimizer steps, with evolved EMA codebooks), and the check is
    the oracle's quantize_assign (models/vqvae_conv3d_latent.py:48-54) in fp64 on exactly those rows; near-ties
    (relative gap < 1e-6) are excluded and counted."""
    from oracle import faceoff_oracle as O

    p0 = O.init_vqvae_params(seed=0)
    img, gt = O.synthetic_clip(2, 3, 64, 64, seed=5)
    model = _load_vqvae(p0)
    opt = torch.optim.Adam(model.parameters(), lr=3e-3)
    for _ in range(3):
        model.zero_grad()
        out, latent = model.forward_with_ids(img.cuda(), clips=2)[:2]
        (torch.nn.functional.mse_loss(out[:, :3], gt.cuda()) + latent.mean()).backward()
        opt.step()
    # A randomly initialised network collapses onto one or two codes, which would leave the argmin untested: re-seed both
    # codebooks k-means style from the REAL pre-quantiser activations (distinct rows + 5 % noise), so that hundreds of codes
    # compete at realistic distances.  Done twice because the bottom quantiser's input depends on the top codebook.
    model.eval()
    gen = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(2):
        with torch.no_grad():
            _, _, _, _, pre_t, pre_b = model.forward_with_ids(img.cuda(), clips=2, return_pre=True)
        for q, pre in ((model.quantize_t, pre_t), (model.quantize_b, pre_b)):
            rows = pre.reshape(-1, q.dim)
            pick = torch.randperm(rows.shape[0], device="cuda", generator=gen)[:q.n_embed]
            pick = pick.repeat((q.n_embed + pick.numel() - 1) // pick.numel())[:q.n_embed]
            noise = torch.randn(q.n_embed, q.dim, device="cuda", generator=gen) * 0.05 * rows.std()
            q.embed.copy_((rows[pick] + noise).t())
    e_t0, e_b0 = model.quantize_t.embed.clone(), model.quantize_b.embed.clone()
    with torch.no_grad():
        dec, diff, id_t, id_b, pre_t, pre_b = model.forward_with_ids(img.cuda(), clips=2, return_pre=True)
    torch.cuda.synchronize()
    for name, ids, pre, emb in (("top", id_t, pre_t, e_t0), ("bottom", id_b, pre_b, e_b0)):
        x = pre.reshape(-1, emb.shape[0]).double().cpu()
        ref, _ = O.quantize_assign(x, emb.double().cpu())
        tie = _near_tie_rows(x, emb.cpu())
        mism = (ids.reshape(-1).cpu() != ref) & ~tie
        used = ids.unique().numel()
        print(f"{name}: {ids.numel()} rows, {used} codes in use, {tie.sum().item()} near-ties, "
              f"{mism.sum().item()} mismatches")
        assert mism.sum().item() == 0
        assert used >= 64, "the re-seeded codebook must spread the rows over many codes"


def test_vgg16_forward_returns_the_five_taps():
    """models/lpips.py:139-152: vgg16.forward(X) -> VggOutputs(relu1_2 .. relu5_3), gradients back to X."""
    from faceoff_b200.lpips import vgg16
    from oracle import faceoff_oracle as O

    lp = O.init_lpips_params(seed=1)
    net = vgg16(pretrained=False)
    net.load_state_dict({k[4:]: v for k, v in lp.items() if k.startswith("net.")}, strict=True)
    net = net.cuda()
    gen = torch.Generator().manual_seed(0)
    x = (torch.rand(2, 3, 32, 32, generator=gen) * 2 - 1)
    xc = x.cuda().requires_grad_(True)
    out = net(xc)
    assert out._fields == ("relu1_2", "relu2_2", "relu3_3", "relu4_3", "relu5_3")
    ref = O.vgg_taps(lp, x)
    for a, b in zip(out, ref):
        assert a.shape == b.shape and maxnorm_err(a.detach().cpu(), b) < 3e-2
    (out.relu2_2.square().mean() + out.relu5_3.mean()).backward()
    xr = x.clone().requires_grad_(True)
    r = O.vgg_taps(lp, xr)
    (r[1].square().mean() + r[4].mean()).backward()
    cos = torch.nn.functional.cosine_similarity(xc.grad.cpu().flatten(), xr.grad.flatten(), dim=0).item()
    assert cos > 0.9, cos     # bf16 trunk with random weights: ill-conditioned (see test_lpips_matches_reference_golden)
    # only an early tap used: the deeper (frozen) layers receive no gradient and must be skipped silently
    xc2 = x.cuda().requires_grad_(True)
    net(xc2).relu1_2.mean().backward()
    assert torch.isfinite(xc2.grad).all() and xc2.grad.abs().sum() > 0


def test_decode_code_and_eval_submethods():
    """models/vqvae_conv3d_latent.py:261-295: only_encode / encode_quantized / decode / decode_code (eager composition of
    the drop-in modules) against the oracle; decode_code(id_t, id_b) must reproduce forward()'s reconstruction."""
    from oracle import faceoff_oracle as O

    p = O.init_vqvae_params(seed=0)
    model = _load_vqvae(p).eval()
    img, _ = O.synthetic_clip(1, 2, 64, 64, seed=3)
    with torch.no_grad():
        dec, diff, id_t, id_b = model.forward_with_ids(img.cuda(), clips=1)
        dec2 = model.decode_code(id_t, id_b)
        enc_b, enc_t = model.only_encode(img.cuda())
    assert dec2.shape == dec.shape
    assert maxnorm_err(dec2.cpu(), dec.cpu()) < 2e-2
    ref = O.vqvae_forward(p, img, n_clips=1, training=False)
    ref_dec = O.decoder(p, "dec", torch.cat([O._ct2(p, "upsample_t", torch.nn.functional.embedding(
        ref["id_t"], p["quantize_t.embed"].t()).permute(0, 3, 1, 2)), torch.nn.functional.embedding(
        ref["id_b"], p["quantize_b.embed"].t()).permute(0, 3, 1, 2)], 1), 4)
    agree = (id_t.cpu() == ref["id_t"]).float().mean().item()
    if agree == 1.0 and (id_b.cpu() == ref["id_b"]).all():
        assert maxnorm_err(dec2.cpu(), ref_dec) < 3e-2
    assert maxnorm_err(enc_b.cpu(), O.encoder(p, "enc_b", img, 4)) < 3e-2
    assert enc_t.shape == (2, 128, 8, 8)


def test_validation_forward_T50_no_grad_saves_nothing():
    """SURVEY 8(f3) / train_faceoff_perceptual.py:53-79: eval forward of a T=50 256x256 clip under no_grad; nothing may
    be kept for backward (peak memory stays far below a training forward) and the codebooks stay untouched."""
    from oracle import faceoff_oracle as O

    p = O.init_vqvae_params(seed=0)
    model = _load_vqvae(p).eval()
    g = torch.Generator().manual_seed(50)
    img = torch.rand(50, 6, 256, 256, generator=g) * 2 - 1
    x = img.cuda()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    with torch.no_grad():
        dec, diff = model(x)
    torch.cuda.synchronize()
    peak_eval = torch.cuda.max_memory_allocated() - base
    assert dec.shape == (50, 6, 256, 256) and not dec.requires_grad
    assert torch.equal(model.quantize_t.embed.cpu(), p["quantize_t.embed"])
    # first 3 frames' worth of the clip against the oracle needs the whole clip (Conv3d mixes frames): compare norm-wise
    torch.set_num_threads(os.cpu_count() or 1)
    ref = O.vqvae_forward(p, img, n_clips=1, training=False)
    nw = ((dec.cpu() - ref["dec"]).norm() / ref["dec"].norm()).item()
    print(f"T=50 eval forward: norm-wise err {nw:.3e}, peak extra memory {peak_eval / 2**30:.2f} GiB")
    assert nw < 5e-2
    model.train()
    torch.cuda.reset_peak_memory_stats()
    out = model(x)
    torch.cuda.synchronize()
    peak_train = torch.cuda.max_memory_allocated() - base
    del out
    assert peak_eval < 0.6 * peak_train, (peak_eval, peak_train)


def test_full_size_clip_with_lpips_vs_oracle():
    """BASELINE configs[2] per-rank unit: one 30-frame 256x256 clip, fwd+bwd WITH the LPIPS loss
    (train_faceoff_perceptual.py:32-47,98) against the CPU oracle run live."""
    from faceoff_b200.lpips import VQLPIPS
    from oracle import faceoff_oracle as O

    p = O.init_vqvae_params(seed=0)
    lp = O.init_lpips_params(seed=1)
    img, gt = O.synthetic_clip(1, 30, 256, 256, seed=1234)
    model = _load_vqvae(p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        vql = VQLPIPS()
    vql.load_state_dict({"perceptual_loss." + k: v for k, v in lp.items()}, strict=True)
    vql = vql.cuda()
    dec, diff, id_t, id_b = model.forward_with_ids(img.cuda(), clips=1)
    recon = torch.nn.functional.mse_loss(dec[:, :3], gt.cuda())
    perc = vql(gt.cuda(), dec[:, :3])
    (recon + diff.mean() + perc).backward()
    torch.cuda.synchronize()
    torch.set_num_threads(os.cpu_count() or 1)
    o = O.train_step(p, img, gt, n_clips=1, lp=lp)
    print(f"full-size clip + LPIPS: perceptual {perc.item():.6f} vs oracle {o['perceptual_loss'].item():.6f}; "
          f"recon {recon.item():.6f} vs {o['recon_loss'].item():.6f}")
    assert abs(perc.item() - o["perceptual_loss"].item()) <= BF16_RTOL * abs(o["perceptual_loss"].item()) + BF16_ATOL
    assert abs(recon.item() - o["recon_loss"].item()) <= BF16_RTOL * abs(o["recon_loss"].item()) + BF16_ATOL
    worst = 0.0
    for k, v in model.named_parameters():
        nref = o["grads"][k].norm().item()
        rel = abs(v.grad.norm().item() - nref) / (nref + 1e-12)
        worst = max(worst, rel)
    print(f"full-size clip + LPIPS: worst grad-norm rel err {worst:.3e}")
    assert worst < 0.1


def maxnorm_err64(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-300)).item()


def _precise_module_check(mod, ref_fn, x, tol=5e-5):
    """Stand-alone drop-in module under the verification mode vs its fp64 restatement on the same piecewise-linear branch
    (tests/gate_consistent.py): output, input gradient and every parameter gradient, max-normalised; plus the per-launch
    verifier (tests/op_verifier.py) on every tensor-core launch."""
    from faceoff_b200 import ops
    from gate_consistent import GateRecorder
    from op_verifier import OpVerifier

    mod = mod.cuda()
    xc = x.cuda().requires_grad_(True)
    with ops.precise_mode(), OpVerifier() as ver, GateRecorder() as rec:
        y = mod(xc)
        go = torch.randn(y.shape, generator=torch.Generator().manual_seed(1)).cuda()
        y.backward(go)
    torch.cuda.synchronize()
    w_err, w_desc = ver.worst()
    print(f"     per-launch verifier: {len(ver.records)} launches, worst {w_err:.1e} ({w_desc})")
    p64 = {"m." + k: v.detach().cpu().double().requires_grad_(v.requires_grad) for k, v in mod.state_dict(keep_vars=True).items()}
    x64 = x.double().requires_grad_(True)
    rec.install()
    try:
        r = ref_fn(p64, x64)
        r.backward(go.cpu().double())
    finally:
        rec.uninstall()
    print("     " + rec.summary())
    errs = {"out": maxnorm_err64(y.detach().cpu(), r.detach()), "dx": maxnorm_err64(xc.grad.cpu(), x64.grad)}
    for k, v in mod.named_parameters():
        errs[k] = maxnorm_err64(v.grad.cpu(), p64["m." + k].grad)
    for k, e in errs.items():
        print(f"     {k:28s} {e:.1e}{'   <-- BAD' if e > tol else ''}")
    bad = {k: e for k, e in errs.items() if e > tol}
    if w_err > tol:
        bad["per-launch"] = (w_err, w_desc)
    return bad


def test_precise_mode_conv_layers_vs_fp64():
    """Every convolution form of the hot path through the tensor-core kernels in the verification mode against fp64:
    4x4 stride-2 (+its transposed form), 3x3 halo / 1x1 + residual (ResBlock), 3x3x3 Conv3d; fwd + dgrad + wgrad + bias
    gradients at <= 5e-5 max-normalised (the bf16 product path can only be held to ~3e-2), per launch AND end to end."""
    from faceoff_b200.vqvae import Conv3dLatentPostnet, Decoder, Encoder, ResBlock
    from oracle import faceoff_oracle as O

    gen = torch.Generator().manual_seed(0)
    torch.manual_seed(0)
    cases = [
        ("ResBlock(128, 32)", ResBlock(128, 32), lambda p, x: O.resblock(p, "m", x), torch.randn(2, 128, 16, 16, generator=gen)),
        ("Encoder(6, 128, 2, 32, stride 4)", Encoder(6, 128, 2, 32, 4), lambda p, x: O.encoder(p, "m", x, 4),
         torch.rand(2, 6, 64, 64, generator=gen) * 2 - 1),
        ("Encoder(128, 128, 2, 32, stride 2)", Encoder(128, 128, 2, 32, 2), lambda p, x: O.encoder(p, "m", x, 2),
         torch.randn(2, 128, 16, 16, generator=gen)),
        ("Encoder(128, 128, 0, 32, stride 2)", Encoder(128, 128, 0, 32, 2), lambda p, x: O.encoder(p, "m", x, 2, n_res_block=0),
         torch.randn(2, 128, 16, 16, generator=gen)),
        ("Decoder(64, 64, 128, 2, 32, stride 2)", Decoder(64, 64, 128, 2, 32, 2), lambda p, x: O.decoder(p, "m", x, 2),
         torch.randn(2, 64, 8, 8, generator=gen)),
        ("Decoder(128, 6, 128, 2, 32, stride 4)", Decoder(128, 6, 128, 2, 32, 4), lambda p, x: O.decoder(p, "m", x, 4),
         torch.randn(2, 128, 16, 16, generator=gen)),
        ("Conv3dLatentPostnet(128) 2x3x8x8", Conv3dLatentPostnet(128), lambda p, x: O.conv3d_postnet(p, "m", x),
         torch.randn(2, 128, 3, 8, 8, generator=gen)),
        ("Conv3dLatentPostnet(128) 1x4x16x16", Conv3dLatentPostnet(128), lambda p, x: O.conv3d_postnet(p, "m", x),
         torch.randn(1, 128, 4, 16, 16, generator=gen)),
    ]
    bad = {}
    for name, mod, ref, x in cases:
        print(name)
        b = _precise_module_check(mod, ref, x)
        if b:
            bad[name] = b
    assert not bad, bad


@pytest.mark.parametrize("tag", ["vqvae_1x4x64", "vqvae_2x3x64_lpips"])
def test_precise_mode_every_gradient_vs_fp64_oracle(tag):
    """The verification mode (faceoff_b200.ops.precise_mode: activations as hi|lo bf16 pairs through the SAME planner and
    tcgen05 kernels, 3 MMAs per product) against the oracle in fp64 on the golden training steps:
      * forward: indices identical, losses rtol 1e-5 (LPIPS 1e-4), reconstruction max-normalised 1e-4, EMA codebooks 2e-5;
      * every tensor-core launch verified in place against torch fp64 on its actual operands (<= 5e-5);
      * EVERY parameter gradient max-normalised <= 1e-4 against the fp64 oracle evaluated on the same piecewise-linear
        branch (audited ReLU gate / pooling-winner overrides, tests/gate_consistent.py), printed next to the plain
        comparison and to the error of the reference's own fp32 arithmetic (|ref_fp32 - fp64|).
    A wrong tap, a dropped residual gradient or a mis-scaled bias gradient is orders of magnitude above 1e-4."""
    from faceoff_b200 import ops
    from gate_consistent import GateRecorder
    from op_verifier import OpVerifier
    from oracle import faceoff_oracle as O

    g = _golden()[tag]
    cfg = g["cfg"]
    p = O.init_vqvae_params(seed=cfg["seed_params"])
    img, gt = O.synthetic_clip(cfg["n_clips"], cfg["T"], cfg["H"], cfg["W"], seed=cfg["seed_data"])
    lp = O.init_lpips_params(seed=cfg["seed_lpips"]) if cfg["with_lpips"] else None
    model = _load_vqvae(p)
    vql = None
    if lp is not None:
        from faceoff_b200.lpips import VQLPIPS

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            vql = VQLPIPS()
        vql.load_state_dict({"perceptual_loss." + k: v for k, v in lp.items()}, strict=True)
        vql = vql.cuda()
    with ops.precise_mode(), OpVerifier() as ver, GateRecorder() as rec:
        dec, diff, id_t, id_b = model.forward_with_ids(img.cuda(), clips=cfg["n_clips"])
        rec_ = dec[:, :3]
        recon = torch.nn.functional.mse_loss(rec_, gt.cuda())
        loss = recon + diff.mean()
        perc = None
        if vql is not None:
            perc = vql(gt.cuda(), rec_)
            loss = loss + perc
        loss.backward()
    torch.cuda.synchronize()
    w_err, w_desc = ver.worst()
    print(f"{tag}: per-launch verifier: {len(ver.records)} tensor-core launches, worst {w_err:.1e} ({w_desc})")
    assert w_err <= 5e-5, (w_err, w_desc)
    o64 = O.train_step(p, img, gt, n_clips=cfg["n_clips"], lp=lp, dtype=torch.float64)
    o32 = O.train_step(p, img, gt, n_clips=cfg["n_clips"], lp=lp, dtype=torch.float32)
    rec.install()
    try:
        o64g = O.train_step(p, img, gt, n_clips=cfg["n_clips"], lp=lp, dtype=torch.float64)
    finally:
        rec.uninstall()
    print(f"{tag}: " + rec.summary())
    assert torch.equal(id_t.cpu(), o64["id_t"]) and torch.equal(id_b.cpu(), o64["id_b"]), "indices differ from the fp64 oracle"
    assert torch.equal(id_t.cpu(), g["id_t"].long()) and torch.equal(id_b.cpu(), g["id_b"].long())
    rel = lambda a, b: abs(a - b) / abs(b)   # noqa: E731
    print(f"{tag}: loss {loss.item():.8f} vs fp64 {o64['loss'].item():.8f}")
    assert rel(recon.item(), o64["recon_loss"].item()) < 1e-5
    assert rel(diff.mean().item(), o64["latent_loss"].item()) < 1e-5
    if perc is not None:
        assert rel(perc.item(), o64["perceptual_loss"].item()) < 1e-4, (perc.item(), o64["perceptual_loss"].item())
    assert maxnorm_err(dec.detach().cpu().double(), o64["dec"]) < 1e-4
    worst = 0.0
    print(f"  {'parameter':42s} |ours-fp64| same branch   |ours-fp64| plain   |ref_fp32-fp64|")
    for k, v in model.named_parameters():
        e_same = maxnorm_err64(v.grad.cpu(), o64g["grads"][k])
        e_plain = maxnorm_err64(v.grad.cpu(), o64["grads"][k])
        e_ref32 = maxnorm_err64(o32["grads"][k], o64["grads"][k])
        print(f"  {k:42s} {e_same:.2e}                {e_plain:.2e}            {e_ref32:.2e}")
        worst = max(worst, e_same)
    print(f"{tag}: worst max-normalised gradient error vs the fp64 oracle (same branch): {worst:.2e}")
    assert worst <= 1e-4
    for q in ("quantize_t", "quantize_b"):
        for i, name in enumerate(("embed", "cluster_size", "embed_avg")):
            got = getattr(getattr(model, q), name).cpu().double()
            # (embed_avg = 0.99 old + 0.01 sum can cancel: the fp32 rounding is relative to the terms, hence the atol)
            torch.testing.assert_close(got, o64["new_buffers"][q][i], rtol=2e-5, atol=5e-6)


def test_precise_mode_lpips_value_and_input_gradient():
    """LPIPS alone in the verification mode: value rtol 1e-4, input gradient max-normalised 1e-4 vs the fp64 oracle on the
    same branch (the bf16 product path can only be held to cosine > 0.98 here, see test_lpips_matches_reference_golden)."""
    from faceoff_b200 import ops
    from faceoff_b200.lpips import LPIPS
    from gate_consistent import GateRecorder
    from oracle import faceoff_oracle as O

    g = _golden()["lpips_3x64"]
    lp = O.init_lpips_params(seed=1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = LPIPS()
    m.load_state_dict(lp, strict=True)
    m = m.cuda().eval()
    a = g["a"].cuda()
    b = g["b"].cuda().requires_grad_(True)
    with ops.precise_mode(), GateRecorder() as rec:
        val = m(a, b)
        val.mean().backward()
    torch.cuda.synchronize()
    lp64 = {k: v.double() for k, v in lp.items()}

    def run64():
        b64 = g["b"].double().requires_grad_(True)
        v64 = O.lpips_forward(lp64, g["a"].double(), b64)
        v64.mean().backward()
        return v64.detach(), b64.grad

    v_plain, g_plain = run64()
    rec.install()
    try:
        v_same, g_same = run64()
    finally:
        rec.uninstall()
    print("precise LPIPS: " + rec.summary())
    torch.testing.assert_close(val.detach().cpu().double(), v_plain, rtol=1e-4, atol=1e-8)
    e_same = maxnorm_err64(b.grad.cpu(), g_same)
    e_plain = maxnorm_err64(b.grad.cpu(), g_plain)
    e32 = maxnorm_err64(g["grad_b"], g_plain)
    print(f"precise LPIPS: input-gradient max-normalised err same branch {e_same:.2e}, plain {e_plain:.2e} "
          f"(reference fp32 vs fp64: {e32:.2e})")
    assert e_same <= 1e-4


def test_two_rank_data_parallel_matches_single_process():
    """Two ranks through FusedDataParallel (NCCL on >= 2 GPUs; both ranks on cuda:0 over gloo on a 1-GPU box) against a
    single-process run of the same clips: gradients, codebooks, no_sync micro-batches, no_grad forward
    (tests/gpu_dp_check.py)."""
    import socket
    import subprocess
    import sys

    root = os.path.dirname(HERE)
    with socket.socket() as s_:
        s_.bind(("127.0.0.1", 0))
        port = s_.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "gpu_dp_check.py")],
                       cwd=root, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0 and "DP CHECK PASS" in r.stdout, r.stdout[-3000:]
