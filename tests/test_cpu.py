"""CPU-side tests (no GPU): oracle vs the reference's golden vectors, C-ABI surface, host logic, gloo DP."""
import ctypes
import os
import re
import subprocess
import sys
import warnings

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden", "golden.pt")


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, map_location="cpu")


# ---------------------------------------------------------------------------------------------- oracle pin
@pytest.mark.parametrize("tag", ["q_small", "q_d128"])
def test_oracle_quantize_vs_reference_golden(golden, tag):
    from oracle import faceoff_oracle as O

    g = golden[tag]
    x = g["x"].clone().requires_grad_(True)
    q, diff, ind, nb, _ = O.quantize_forward(x, g["embed0"], g["cluster_size0"], g["embed_avg0"], True)
    (q * g["gq"]).sum().add(diff * 3.0).backward()
    assert torch.equal(ind, g["embed_ind"])
    torch.testing.assert_close(q.detach(), g["quantize"])
    torch.testing.assert_close(diff.detach(), g["diff"])
    torch.testing.assert_close(x.grad, g["grad_x"])
    torch.testing.assert_close(nb[0], g["embed1"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(nb[1], g["cluster_size1"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(nb[2], g["embed_avg1"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("tag", ["vqvae_1x4x64", "vqvae_2x3x64_lpips"])
def test_oracle_train_step_vs_reference_golden(golden, tag):
    from oracle import faceoff_oracle as O

    g = golden[tag]
    cfg = g["cfg"]
    p = O.init_vqvae_params(seed=cfg["seed_params"])
    img, gt = O.synthetic_clip(cfg["n_clips"], cfg["T"], cfg["H"], cfg["W"], seed=cfg["seed_data"])
    lp = O.init_lpips_params(seed=cfg["seed_lpips"]) if cfg["with_lpips"] else None
    o = O.train_step(p, img, gt, n_clips=cfg["n_clips"], lp=lp)
    torch.testing.assert_close(o["loss"], g["loss"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(o["dec"][0], g["dec_full0"], rtol=1e-4, atol=1e-5)
    assert torch.equal(o["id_t"], g["id_t"].long()) and torch.equal(o["id_b"], g["id_b"].long())
    for k, gref in g["grads_ref"].items():
        got = o["grads"][k]
        got = got if got.numel() == gref.numel() else got[:8, :8]
        torch.testing.assert_close(got, gref, rtol=1e-3, atol=1e-6)
    for k, n in g["grad_norms_ref"].items():
        assert abs(o["grads"][k].norm().item() - n.item()) <= 1e-3 * n.item() + 1e-9
    for k, bref in g["buffers_ref"].items():
        mod, name = k.split(".")
        idx = ("embed", "cluster_size", "embed_avg").index(name)
        torch.testing.assert_close(o["new_buffers"][mod][idx], bref, rtol=1e-5, atol=1e-6)


def test_oracle_lpips_vs_reference_golden(golden):
    from oracle import faceoff_oracle as O

    g = golden["lpips_3x64"]
    lp = O.init_lpips_params(seed=1)
    b = g["b"].clone().requires_grad_(True)
    val = O.lpips_forward(lp, g["a"], b)
    val.mean().backward()
    torch.testing.assert_close(val.detach(), g["val"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(b.grad, g["grad_b"], rtol=1e-4, atol=1e-8)


def test_oracle_batched_equals_ddp_semantics():
    """B clips in one call == B ranks with SUM-all-reduced EMA statistics and averaged grads."""
    from oracle import faceoff_oracle as O

    p = O.init_vqvae_params(seed=0)
    img, gt = O.synthetic_clip(2, 2, 32, 32, seed=9)
    both = O.vqvae_forward(p, img, n_clips=2)
    a = O.vqvae_forward(p, img[:2], n_clips=1)
    b = O.vqvae_forward(p, img[2:], n_clips=1)
    torch.testing.assert_close(both["dec"], torch.cat([a["dec"], b["dec"]]), rtol=1e-4, atol=1e-5)
    for q in ("quantize_t", "quantize_b"):
        cnt = a["stats"][q][0] + b["stats"][q][0]
        torch.testing.assert_close(both["stats"][q][0], cnt)
        torch.testing.assert_close(both["stats"][q][1], a["stats"][q][1] + b["stats"][q][1], rtol=1e-4, atol=1e-5)


def test_quantize_properties_hypothesis():
    from hypothesis import given, settings, strategies as st
    from oracle import faceoff_oracle as O

    @settings(max_examples=20, deadline=None)
    @given(st.integers(1, 64), st.sampled_from([4, 8]), st.sampled_from([3, 16]), st.integers(0, 10 ** 6))
    def run(rows, dim, k, seed):
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(rows, dim, generator=g, dtype=torch.float64)
        e = torch.randn(dim, k, generator=g, dtype=torch.float64)
        q, diff, ind, nb, stats = O.quantize_forward(x, e, torch.zeros(k, dtype=torch.float64), e.clone(), True)
        d = ((x[:, :, None] - e[None]) ** 2).sum(1)
        assert torch.equal(ind, d.argmin(1))
        assert abs(stats[0].sum().item() - rows) < 1e-9          # counts bookkeeping
        torch.testing.assert_close(stats[1].sum(1), x.sum(0))     # embed_sum conserves the rows
        torch.testing.assert_close(q, e.t()[ind])                 # straight-through value

    run()


# ---------------------------------------------------------------------------------------------- C ABI
def _header_symbols():
    text = open(os.path.join(ROOT, "include", "faceoff_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fo_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    from faceoff_b200 import _lib

    so = os.path.join(ROOT, "faceoff_b200", "libfaceoff_b200.so")
    if not os.path.exists(so):
        _lib.build()
    lib = ctypes.CDLL(so)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/faceoff_b200.h but not exported"
    # and the Python binding covers all of them
    assert set(syms) == set(_lib.EXPORTED_SYMBOLS)


def test_fastdiv_magic_numbers(tmp_path):
    """Host logic of the conv kernels' tile decode: the multiply-high divider (csrc/igemm.cuh make_fastdiv) equals integer
    division for every divisor 1..4096 (+ a spread up to 2e6) over dividends covering [0, 2^31)."""
    exe = str(tmp_path / "fastdiv_check")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.run(["g++", "-O2", "-I", cuda_inc, os.path.join(ROOT, "tests", "fastdiv_check.cpp"), "-o", exe], check=True)
    res = subprocess.run([exe], stdout=subprocess.PIPE, text=True, timeout=120)
    assert res.returncode == 0 and "bad=0" in res.stdout, res.stdout


def test_wgrad_accumulation_chain_plan(tmp_path):
    """Host logic of the weight-gradient kernel's bounded accumulation chains (csrc/igemm.cuh wgrad_plan_chains + the chunk
    arithmetic of csrc/wgrad_igemm.cu, modelled in tests/wgrad_chain_check.cpp): chunks cover every split's tile range once,
    respect the MMA bound, enumerate the partial slots once, and the bias-column predicate matches a direct search."""
    exe = str(tmp_path / "wgrad_chain_check")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.run(["g++", "-O2", "-I", cuda_inc, os.path.join(ROOT, "tests", "wgrad_chain_check.cpp"), "-o", exe], check=True)
    res = subprocess.run([exe], stdout=subprocess.PIPE, text=True, timeout=300)
    assert res.returncode == 0 and "bad=0" in res.stdout, res.stdout


@pytest.mark.parametrize("fuse", [True, False])
def test_lpips_tape_logic_with_cpu_stand_in_kernels(monkeypatch, fuse):
    """Host logic of the LPIPS path (faceoff_b200/lpips.py + graph.py) on CPU: the product's tape -- two VGG trunks in
    lockstep, taps that take over the max pools next to them (forward and backward), the recorded backward order, the
    fp32 NCHW input gradient -- runs over plain-torch stand-ins of the kernels (tests/fake_ops.py) and must reproduce the
    oracle's value and input gradient; the op counts show which kernels the tape chose."""
    import warnings

    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import fake_ops
    from faceoff_b200 import lpips as L_
    from oracle import faceoff_oracle as O

    fake_ops.install(monkeypatch)
    monkeypatch.setattr(L_, "_FUSE", fuse)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = L_.LPIPS().eval()
    p = O.init_lpips_params(seed=3)
    model.load_state_dict(p)
    g = torch.Generator().manual_seed(7)
    target = torch.rand(2, 3, 32, 32, generator=g) * 2 - 1
    x = (torch.rand(2, 3, 32, 32, generator=g) * 2 - 1).requires_grad_(True)
    val = model(target, x)                       # gradient flows to the second argument, as in the reference trainer
    assert val.shape == (2, 1, 1, 1)
    (val * torch.tensor([1.0, -0.5]).view(2, 1, 1, 1)).sum().backward()
    xo = x.detach().clone().requires_grad_(True)
    ref = O.lpips_forward(p, target, xo)
    (ref * torch.tensor([1.0, -0.5]).view(2, 1, 1, 1)).sum().backward()
    torch.testing.assert_close(val.detach(), ref.detach(), rtol=1e-4, atol=1e-6)
    err = ((x.grad - xo.grad).abs().max() / xo.grad.abs().max()).item()
    assert err < 1e-4, err
    c = fake_ops.CALLS
    if fuse:    # four pooled taps fused both ways, the last tap plain; no stand-alone pool kernels at all
        assert (c["lpips_tap_pool"], c["lpips_tap"], c["lpips_tap_bwd_pool"], c["lpips_tap_bwd"]) == (4, 1, 4, 1)
        assert c["maxpool2"] == 0 and c["maxpool2_bwd"] == 0 and c["vgg_first_dgrad"] == 1 and c["unpack_nchw"] == 0
    else:
        assert (c["lpips_tap_pool"], c["lpips_tap"], c["lpips_tap_bwd_pool"], c["lpips_tap_bwd"]) == (0, 5, 0, 5)
        assert c["maxpool2"] == 8 and c["maxpool2_bwd"] == 4 and c["vgg_first_dgrad"] == 0
    assert c["vgg_first_conv"] == 2


def test_vgg16_forward_tape_logic_with_cpu_stand_in_kernels(monkeypatch):
    """The stand-alone trunk (vgg16.forward, reference models/lpips.py:139-152) over the same stand-ins: five taps and the
    gradient of a weighted sum of them w.r.t. the input equal the oracle's (tap gradients enter through pack_nchw and meet
    the pool gradients in add_grads; the pools keep their own kernels here)."""
    import warnings

    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import fake_ops
    from faceoff_b200 import lpips as L_
    from oracle import faceoff_oracle as O

    fake_ops.install(monkeypatch)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = L_.vgg16(pretrained=False).eval()
    p = {k[len("net."):]: v for k, v in O.init_lpips_params(seed=5).items() if k.startswith("net.")}
    net.load_state_dict(p)
    g = torch.Generator().manual_seed(9)
    x = (torch.rand(2, 3, 32, 32, generator=g) * 2 - 1).requires_grad_(True)
    outs = net(x)
    assert outs._fields == ("relu1_2", "relu2_2", "relu3_3", "relu4_3", "relu5_3")
    sum(((k + 1) * t).sum() for k, t in enumerate(outs)).backward()
    xo = x.detach().clone().requires_grad_(True)
    ref = O.vgg_taps({"net." + k: v for k, v in p.items()}, xo)
    sum(((k + 1) * t).sum() for k, t in enumerate(ref)).backward()
    for a, b in zip(outs, ref):
        torch.testing.assert_close(a.detach(), b.detach(), rtol=1e-4, atol=1e-5)
    err = ((x.grad - xo.grad).abs().max() / xo.grad.abs().max()).item()
    assert err < 1e-4, err
    c = fake_ops.CALLS
    assert c["maxpool2"] == 4 and c["maxpool2_bwd"] == 4 and c["lpips_tap_pool"] == 0 and c["add_grads"] == 4


def _stack_case(name):
    """(product module, oracle function of (params, x), input) for one stand-alone conv stack."""
    from faceoff_b200 import vqvae as V
    from oracle import faceoff_oracle as O

    g = torch.Generator().manual_seed(11)
    if name == "resblock":
        m = V.ResBlock(32, 16)
        return m, (lambda p, x: O.resblock(p, "m", x)), torch.randn(2, 32, 6, 8, generator=g)
    if name.startswith("encoder"):
        stride, nres = (4, 2) if name == "encoder4" else (2, 1) if name == "encoder2" else (2, 0)
        m = V.Encoder(16, 32, nres, 16, stride)
        return m, (lambda p, x: O.encoder(p, "m", x, stride, nres)), torch.randn(2, 16, 16, 8, generator=g)
    if name.startswith("decoder"):
        stride, nres = (4, 2) if name == "decoder4" else (2, 1)
        m = V.Decoder(16, 16, 32, nres, 16, stride)
        return m, (lambda p, x: O.decoder(p, "m", x, stride, nres)), torch.randn(2, 16, 4, 6, generator=g)
    m = V.Conv3dLatentPostnet(16)
    return m, (lambda p, x: O.conv3d_postnet(p, "m", x)), torch.randn(2, 16, 3, 4, 6, generator=g)


@pytest.mark.parametrize("name", ["resblock", "encoder4", "encoder2", "encoder2_nores", "decoder4", "decoder2", "conv3d"])
def test_conv_stack_tape_logic_with_cpu_stand_in_kernels(monkeypatch, name):
    """Host logic of the conv stacks (faceoff_b200/graph.py conv_op / resblock_op + the graph builders of vqvae.py) on CPU:
    forward value, input gradient and EVERY parameter gradient of the stand-alone drop-in modules over plain-torch stand-ins
    of fo_conv_run / fo_wgrad_run / fo_colsum (tests/fake_ops.py) against the oracle's autograd -- ReLU views, residual
    pass-through, fused / column-sum bias gradients, swapped weight-gradient roles of the narrow layers, transposed-conv
    weight layout, the 3-D views.  (``encoder2_nores``: the n_res_block = 0 configuration of the advisor's note.)"""
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import fake_ops

    fake_ops.install(monkeypatch)
    torch.manual_seed(3)
    m, ref_fn, x = _stack_case(name)
    m.train()
    p = {"m." + k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    x1 = x.clone().requires_grad_(True)
    y = m(x1)
    x2 = x.clone().requires_grad_(True)
    yr = ref_fn(p, x2)
    assert y.shape == yr.shape
    gy = torch.randn(yr.shape, generator=torch.Generator().manual_seed(5))
    (y * gy).sum().backward()
    (yr * gy).sum().backward()
    torch.testing.assert_close(y.detach(), yr.detach(), rtol=1e-4, atol=1e-5)
    assert ((x1.grad - x2.grad).abs().max() / x2.grad.abs().max()).item() < 1e-4
    for k, v in m.named_parameters():
        gr = p["m." + k].grad
        assert v.grad is not None, k
        err = ((v.grad - gr).abs().max() / (gr.abs().max() + 1e-30)).item()
        assert err < 1e-4, (k, err)


@pytest.mark.parametrize("clips,with_lpips", [(1, False), (2, False), (1, True)])
def test_vqvae_train_step_tape_logic_with_cpu_stand_in_kernels(monkeypatch, clips, with_lpips):
    """The whole fused training graph of the product VQVAE (faceoff_b200/vqvae.py:_runner + graph.py: both encoders, the Conv3d
    latent blocks on the clip views, the two quantisers with their straight-through / commitment gradients and EMA update,
    torch.cat folded into K loops, both decoders, the image-side layers) on CPU over the stand-in kernels: loss, reconstruction,
    code indices, EVERY parameter gradient and the six codebook buffers after the step against the oracle's train step;
    ``with_lpips``: the north-star target step (+ VQLPIPS(gt, out), train_faceoff_perceptual.py:32-47) -- two gradient sources
    meet at the reconstruction."""
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import fake_ops
    from faceoff_b200.vqvae import VQVAE
    from oracle import faceoff_oracle as O

    fake_ops.install(monkeypatch)
    p = O.init_vqvae_params(seed=2)
    img, gt = O.synthetic_clip(clips, 2, 32, 32, seed=8)
    model = VQVAE(in_channel=6)
    model.load_state_dict(p)
    model.train()
    out, latent, id_t, id_b = model.forward_with_ids(img, clips)
    loss = torch.nn.functional.mse_loss(out[:, :3], gt) + latent.mean()
    lp = None
    if with_lpips:
        from faceoff_b200.lpips import VQLPIPS

        lp = O.init_lpips_params(seed=4)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            vql = VQLPIPS()
        vql.load_state_dict({"perceptual_loss." + k: v for k, v in lp.items()})
        loss = loss + vql(gt, out[:, :3])
    loss.backward()
    o = O.train_step(p, img, gt, n_clips=clips, lp=lp)
    assert torch.equal(id_t, o["id_t"]) and torch.equal(id_b, o["id_b"])
    torch.testing.assert_close(out.detach(), o["dec"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(loss.detach(), o["loss"], rtol=1e-5, atol=1e-7)
    worst = ("", 0.0)
    for k, v in model.named_parameters():
        gr = o["grads"][k]
        assert v.grad is not None, k
        err = ((v.grad - gr).abs().max() / (gr.abs().max() + 1e-30)).item()
        if err > worst[1]:
            worst = (k, err)
    assert worst[1] < 2e-4, worst
    for q in ("quantize_t", "quantize_b"):       # (embed, cluster_size, embed_avg) after the EMA update
        for i, name in enumerate(("embed", "cluster_size", "embed_avg")):
            torch.testing.assert_close(getattr(getattr(model, q), name), o["new_buffers"][q][i], rtol=1e-5, atol=1e-6)
    c = fake_ops.CALLS
    # the image-side layers took the direct kernels (no im2col matrix), the cat of (dec_t, enc_b) never materialised
    assert c["s2conv"] == 2 and c["s2wgrad"] == 2 and c["im2col4x4s2"] == 0 and c["col2im4x4s2"] == 1
    assert c["vq_assign"] == 2 and c["vq_ema"] == 2 and c["vq_backward"] == 2


@pytest.mark.parametrize("dim,K", [(64, 512), (32, 40)])
def test_quantize_module_autograd_with_cpu_stand_in_kernels(monkeypatch, dim, K):
    """The stand-alone Quantize module (its own autograd Function: straight-through + commitment gradient, EMA in train
    mode only, reference :47-83) over the stand-ins against the oracle."""
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import fake_ops
    from faceoff_b200.vqvae import Quantize
    from oracle import faceoff_oracle as O

    fake_ops.install(monkeypatch)
    g = torch.Generator().manual_seed(dim + K)
    q = Quantize(dim, K)
    with torch.no_grad():
        q.embed.copy_(torch.randn(dim, K, generator=g))
        q.embed_avg.copy_(q.embed)
        q.cluster_size.copy_(torch.rand(K, generator=g))
    e0, c0, a0 = q.embed.clone(), q.cluster_size.clone(), q.embed_avg.clone()
    x = torch.randn(2, 5, 3, dim, generator=g)
    gq = torch.randn(2, 5, 3, dim, generator=g)
    x1 = x.clone().requires_grad_(True)
    quant, diff, ind = q.train()(x1)
    (quant * gq).sum().add(diff * 3.0).backward()
    x2 = x.clone().requires_grad_(True)
    oq, odiff, oind, obuf, _ = O.quantize_forward(x2, e0, c0, a0, training=True)
    (oq * gq).sum().add(odiff * 3.0).backward()
    assert torch.equal(ind, oind) and not ind.requires_grad
    torch.testing.assert_close(quant.detach(), oq.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(diff.detach(), odiff.detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(x1.grad, x2.grad, rtol=1e-5, atol=1e-7)
    for got, want in zip((q.embed, q.cluster_size, q.embed_avg), obuf):
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)
    # eval mode: same outputs from the updated codebook, buffers untouched
    e1 = q.embed.clone()
    quant_e, _, ind_e = q.eval()(x)
    assert torch.equal(q.embed, e1) and fake_ops.CALLS["vq_ema"] == 1
    torch.testing.assert_close(quant_e, q.embed_code(ind_e), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("in_channel", [3, 16])
def test_vqvae_other_input_widths_with_cpu_stand_in_kernels(monkeypatch, in_channel):
    """in_channel = 3 takes the direct image-side kernels with 3 channels; in_channel = 16 the generic path (packed
    channels-last input, FORM_DOWN first conv without an input gradient, fp32 NCHW output from the last transposed conv's
    epilogue).  Also: the 5-D input [B, T, C, H, W] (B clips in one call) and a validation-style eval / no_grad forward that
    must leave the codebooks untouched and record nothing."""
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import fake_ops
    from faceoff_b200.vqvae import VQVAE
    from oracle import faceoff_oracle as O

    fake_ops.install(monkeypatch)
    p = O.init_vqvae_params(seed=4, in_channel=in_channel)
    img, gt = O.synthetic_clip(2, 2, 32, 32, seed=9, in_channel=in_channel)
    model = VQVAE(in_channel=in_channel)
    model.load_state_dict(p)
    model.train()
    out, latent = model(img.view(2, 2, in_channel, 32, 32))          # 5-D: two clips of two frames
    assert out.shape == (2, 2, in_channel, 32, 32)
    loss = torch.nn.functional.mse_loss(out.reshape(4, in_channel, 32, 32)[:, :3], gt) + latent.mean()
    loss.backward()
    o = O.train_step(p, img, gt, n_clips=2)
    torch.testing.assert_close(loss.detach(), o["loss"], rtol=1e-5, atol=1e-7)
    for k, v in model.named_parameters():
        gr = o["grads"][k]
        err = ((v.grad - gr).abs().max() / (gr.abs().max() + 1e-30)).item()
        assert err < 2e-4, (k, err)
    c = fake_ops.CALLS
    assert (c["s2conv"] > 0) == (in_channel <= 8) and (c["pack_nchw"] > 0) == (in_channel > 8)
    # validation: eval + no_grad, codebooks untouched, output of the UPDATED codebooks == oracle with the new buffers
    model.eval()
    before = {k: v.clone() for k, v in model.named_buffers()}
    n_ema = c["vq_ema"]
    with torch.no_grad():
        out_e, _ = model(img)
    assert c["vq_ema"] == n_ema and all(torch.equal(v, before[k]) for k, v in model.named_buffers())
    p2 = dict(p)
    for q in ("quantize_t", "quantize_b"):
        for i, name in enumerate(("embed", "cluster_size", "embed_avg")):
            p2[f"{q}.{name}"] = o["new_buffers"][q][i]
    ref = O.vqvae_forward(p2, img, n_clips=1, training=False)       # 4-D input: ONE clip of four frames (reference :247)
    torch.testing.assert_close(out_e, ref["dec"], rtol=1e-4, atol=1e-5)


def test_vqvae_eval_sub_methods_with_cpu_stand_in_kernels(monkeypatch):
    """The reference's eager sub-methods (only_encode / encode_quantized / decode / decode_code,
    models/vqvae_conv3d_latent.py:261-295) compose the drop-in modules one by one; in eval mode they must give the oracle's
    indices and reconstruction, and decode_code(id_t, id_b) must reproduce decode(quant_t, quant_b).  (The reference's
    own ``forward`` also runs the Conv3d latent blocks between encode and quantise; the sub-method chain does not, so the
    oracle is evaluated the same way: encoder -> quantise -> decode.)"""
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import fake_ops
    from faceoff_b200.vqvae import VQVAE
    from oracle import faceoff_oracle as O

    fake_ops.install(monkeypatch)
    p = O.init_vqvae_params(seed=6)
    img, _ = O.synthetic_clip(1, 2, 32, 32, seed=3)
    model = VQVAE(in_channel=6)
    model.load_state_dict(p)
    model.eval()
    with torch.no_grad():
        enc_b, enc_t = model.only_encode(img)
        quant_t, quant_b, diff, id_t, id_b = model.encode_quantized(enc_b, enc_t)
        dec = model.decode(quant_t, quant_b)
        dec_c = model.decode_code(id_t, id_b)
        # the same chain from the oracle's pieces
        ob = O.encoder(p, "enc_b", img, 4)
        ot = O.encoder(p, "enc_t", ob, 2)
        qt_in = O._c2(p, "quantize_conv_t", ot).permute(0, 2, 3, 1)
        oqt, dt, oid_t, _, _ = O.quantize_forward(qt_in, p["quantize_t.embed"], p["quantize_t.cluster_size"],
                                                  p["quantize_t.embed_avg"], training=False)
        oqt = oqt.permute(0, 3, 1, 2)
        odt = O.decoder(p, "dec_t", oqt, 2)
        qb_in = O._c2(p, "quantize_conv_b", torch.cat([odt, ob], 1)).permute(0, 2, 3, 1)
        oqb, db, oid_b, _, _ = O.quantize_forward(qb_in, p["quantize_b.embed"], p["quantize_b.cluster_size"],
                                                  p["quantize_b.embed_avg"], training=False)
        oqb = oqb.permute(0, 3, 1, 2)
        odec = O.decoder(p, "dec", torch.cat([O._ct2(p, "upsample_t", oqt), oqb], 1), 4)
    torch.testing.assert_close(enc_b, ob, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(enc_t, ot, rtol=1e-4, atol=1e-5)
    assert torch.equal(id_t, oid_t) and torch.equal(id_b, oid_b)
    torch.testing.assert_close(diff.reshape(-1), (dt + db).reshape(-1), rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(dec, odec, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dec_c, dec, rtol=1e-5, atol=1e-6)
    for k, v in model.named_buffers():
        assert torch.equal(v, p[k]), f"eval mode touched {k}"


def test_no_cpu_fallback_fails_loudly():
    """Without a GPU every op must raise (no silent eager/CPU path)."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from faceoff_b200 import _lib
    from faceoff_b200.vqvae import Quantize

    q = Quantize(64, 512)
    with pytest.raises((_lib.FaceoffB200Error, RuntimeError, AssertionError)):
        q(torch.randn(4, 64))


def test_loss_and_optimizer_have_no_cpu_path():
    """faceoff_b200.losses.mse_loss / optim.FusedAdam refuse CPU tensors instead of silently computing in PyTorch."""
    from faceoff_b200 import _lib
    from faceoff_b200.losses import mse_loss
    from faceoff_b200.optim import FusedAdam

    with pytest.raises(_lib.FaceoffB200Error):
        mse_loss(torch.randn(2, 3, 8, 8, requires_grad=True), torch.randn(2, 3, 8, 8))
    p = torch.nn.Parameter(torch.randn(10))
    p.grad = torch.randn(10)
    opt = FusedAdam([p], lr=3e-4)
    with pytest.raises(_lib.FaceoffB200Error):
        opt.step()
    with pytest.raises(ValueError):
        FusedAdam([p], lr=-1.0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "faceoff_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src or f == "__init__.py" and "oracle" not in src, f"{f} mentions oracle"


# ---------------------------------------------------------------------------------------------- host logic
def test_state_dict_keys_match_reference_layout():
    from faceoff_b200.vqvae import VQVAE
    from faceoff_b200.lpips import LPIPS
    from oracle import faceoff_oracle as O

    m = VQVAE(in_channel=6)
    p = O.init_vqvae_params()
    assert set(m.state_dict().keys()) == set(p.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(p[k].shape), k
    assert sum(x.numel() for x in m.parameters()) == 4_049_990
    m.load_state_dict({("module." + k)[7:]: v for k, v in p.items()})  # DDP-prefixed checkpoints are stripped by callers
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        l = LPIPS()
    lp = O.init_lpips_params()
    assert set(l.state_dict().keys()) == set(lp.keys())
    assert all(not q.requires_grad for q in l.parameters())
    assert sum(x.numel() for x in l.parameters()) == 14_716_160


def test_constructor_signatures():
    import inspect
    from faceoff_b200 import vqvae

    assert list(inspect.signature(vqvae.Quantize.__init__).parameters)[1:] == ["dim", "n_embed", "decay", "eps"]
    assert list(inspect.signature(vqvae.VQVAE.__init__).parameters)[1:] == [
        "in_channel", "channel", "n_res_block", "n_res_channel", "embed_dim", "n_embed", "decay", "residual"]
    assert list(inspect.signature(vqvae.Encoder.__init__).parameters)[1:] == [
        "in_channel", "channel", "n_res_block", "n_res_channel", "stride"]
    assert list(inspect.signature(vqvae.Decoder.__init__).parameters)[1:] == [
        "in_channel", "out_channel", "channel", "n_res_block", "n_res_channel", "stride"]


def test_distributed_api_noop_without_group():
    from faceoff_b200 import distributed as dist

    assert dist.get_rank() == 0 and dist.get_world_size() == 1 and dist.is_primary()
    t = torch.ones(3)
    assert dist.all_reduce(t) is t
    assert dist.all_gather({"a": 1}) == [{"a": 1}]
    dist.synchronize()
    called = []
    dist.launch(lambda a: called.append(a), 1, args=(7,))
    assert called == [7]


_GLOO_WORKER = r"""
import os, sys, torch
sys.path.insert(0, %(root)r)
import torch.distributed as td
from torch import nn
from faceoff_b200 import distributed as dist
from faceoff_b200.parallel import FusedDataParallel
from faceoff_b200.graph import Tape
from faceoff_b200.vqvae import Quantize
from oracle import faceoff_oracle as O
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
td.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world)
assert dist.get_world_size() == 2 and dist.get_rank() == rank
t = torch.full((4,), float(rank + 1)); dist.all_reduce(t); assert torch.equal(t, torch.full((4,), 3.0))
objs = dist.all_gather({"mse_sum": rank * 1.5, "mse_n": 30})
assert [o["mse_sum"] for o in objs] == [0.0, 1.5]
red = dist.reduce_dict({"a": torch.tensor(float(rank)), "b": torch.tensor(2.0)})
if rank == 0:
    assert abs(red["a"].item() - 0.5) < 1e-6 and abs(red["b"].item() - 2.0) < 1e-6

# ---- the PRODUCT reducer (FusedDataParallel) driven through its tape hooks on CPU tensors.  The only CUDA call it
# ---- makes is Quantize.apply_ema (fo_vq_ema); here that one method is replaced by the oracle's EMA formula.
class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(5, 3); self.b = nn.Linear(3, 2)
        self.quantize_t = Quantize(4, 8)
    def param_forward_order(self):
        return ["a.weight", "a.bias", "b.weight", "b.bias"]
torch.manual_seed(0)
net = Net()
ema_calls = []
def cpu_ema(self, counts, embed_sum):
    ema_calls.append((counts.clone(), embed_sum.clone()))
    e, c, a = O.quantize_ema(self.embed, self.cluster_size, self.embed_avg, counts, embed_sum, self.decay, self.eps)
    self.embed.copy_(e); self.cluster_size.copy_(c); self.embed_avg.copy_(a)
Quantize.apply_ema = cpu_ema
ddp = FusedDataParallel(net, n_chunks=3)
params = dict(net.named_parameters())
q = net.quantize_t

def one_pass(scale, sync_ctx=None):
    tape = Tape(params, need_grad=True)
    ddp.begin_forward(params, net.param_forward_order())
    counts, esum = ddp.stat_buffers(q)
    counts += float(rank + 1) * scale; esum += 10.0 * (rank + 1) * scale     # what the gather kernel does: ADD
    ddp.submit_stats(q)
    ddp.begin_step(tape)
    for name in reversed(net.param_forward_order()):                          # backward order
        g, acc = tape.grad_buffer(name)
        val = torch.full_like(g, float(rank + 1) * scale)
        g.add_(val) if acc else g.copy_(val)
        tape.grad_ready(name)
    ddp.end_step()
    ddp.grads_for_autograd(tape, tuple(params))

one_pass(1.0)
assert len(ddp._works) >= 2, "bucket must be reduced in several chunks"
for n, p_ in params.items():
    assert p_.grad is not None and torch.allclose(p_.grad, torch.full_like(p_, 1.5)), (n, p_.grad)   # mean of 1, 2
assert len(ema_calls) == 1
assert torch.allclose(ema_calls[0][0], torch.full((8,), 3.0)) and torch.allclose(ema_calls[0][1], torch.full((4, 8), 30.0))
# codebooks identical on both ranks
gathered = dist.all_gather(q.embed.clone())
assert torch.equal(gathered[0], gathered[1])
# micro-batches: two passes inside no_sync + one outside = ONE collective, ONE EMA with the summed statistics
net.zero_grad(set_to_none=True)
ema_calls.clear()
with ddp.no_sync():
    one_pass(1.0); one_pass(2.0)
assert not ema_calls and ddp._accumulating
one_pass(4.0)
assert len(ema_calls) == 1 and torch.allclose(ema_calls[0][0], torch.full((8,), 3.0 * 7.0))
for n, p_ in params.items():
    assert torch.allclose(p_.grad, torch.full_like(p_, 1.5 * 7.0)), (n, p_.grad)
# a deferred forward whose backward never runs is settled (all-reduce + EMA) before the next forward starts
ema_calls.clear()
ddp.begin_forward(params, net.param_forward_order())
c, e = ddp.stat_buffers(q); c += 1.0; ddp.submit_stats(q)
ddp.begin_forward(params, net.param_forward_order())
assert len(ema_calls) == 1 and torch.allclose(ema_calls[0][0], torch.full((8,), 2.0))
assert float(ddp.stat_buffers(q)[0].sum()) == 0.0
b0, c0 = ddp.bucket_checksums()
both = dist.all_gather((float(b0), float(c0)))
assert both[0] == both[1]
td.destroy_process_group()
print("ok", rank)
"""


def test_gloo_world2_collectives_and_fused_data_parallel(tmp_path):
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER % {"root": ROOT, "port": port})
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_two_rank_data_parallel_step_on_cpu_stand_in_kernels():
    """Reference data-parallel semantics (distributed/distributed.py:64-72 + DDP, train_faceoff_perceptual.py:164-169) of the
    PRODUCT's FusedDataParallel with the real VQVAE + LPIPS tapes, two ranks over gloo on the CPU (kernels = tests/fake_ops.py):
    one clip per rank == the same two clips in one process (every gradient, codebooks; replicas bit-identical), no_sync
    micro-batches == one batch with ONE EMA update, train-mode forward under no_grad applies the EMA in forward
    (tests/gpu_dp_check.py --cpu; the same script runs on GPUs under ``pytest -m gpu``)."""
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "gpu_dp_check.py"), "--cpu"], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert "DP CHECK PASS" in outs[0], outs[0]


def test_launch_spawns_world2(tmp_path):
    """distributed.launch (reference distributed/launch.py:22-92) with n_gpu_per_machine=2: spawn, tcp rendezvous on
    127.0.0.1, per-machine LOCAL_PROCESS_GROUP, helpers -- exercised with the gloo backend (no GPU here)."""
    code = (f"import sys; sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {HERE!r})\n"
            "from faceoff_b200 import distributed as dist\n"
            "import _launch_worker\n"
            f"if __name__ == '__main__':\n"
            f"    dist.launch(_launch_worker.worker, 2, dist_url='auto', args=({str(tmp_path)!r}, 'x'), backend='gloo')\n")
    script = tmp_path / "l.py"
    script.write_text(code)
    r = subprocess.run([sys.executable, str(script)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=240)
    assert r.returncode == 0, r.stdout[-2000:]
    got = sorted(open(tmp_path / f"rank{i}.txt").read() for i in range(2))
    assert got == ["x 0 2 3.0 1 1", "x 1 2 3.0 1 0"], got
    from faceoff_b200 import distributed as dist

    with pytest.raises(ValueError):
        dist.launch(lambda: None, 2, n_machine=2, dist_url="auto")
    with pytest.raises(ValueError):
        dist.launch(lambda: None, 2, n_machine=2, dist_url="file:///tmp/x")


# ---------------------------------------------------------------------------------------------- discriminators (8(f1))
@pytest.mark.parametrize("kind", ["img", "vid"])
def test_disc_oracle_and_dropin_init_vs_reference_golden(kind):
    """oracle/disc_oracle.py against the committed outputs of the reference discriminators, run on the state_dict of the
    DROP-IN module built from the same seed (which must reproduce the reference's initial weights: checksums)."""
    from faceoff_b200.mocoganhd import content_disc, video_disc
    from oracle import disc_oracle as DO

    g = torch.load(os.path.join(HERE, "golden", "golden_disc.pt"), map_location="cpu")[kind]
    torch.manual_seed(g["seed_model"])
    m = content_disc.ModelD_img(3, "instance", 2, 1e-4) if kind == "img" else video_disc.ModelD_3d(3, "instance", 2, 1e-4, False, 12)
    sd = m.state_dict()
    assert set(k for k, v in sd.items() if v.dtype.is_floating_point) == set(g["param_checksums"])
    for k, (s_, a_) in g["param_checksums"].items():
        assert abs(sd[k].double().sum().item() - s_) < 1e-9 and abs(sd[k].double().abs().sum().item() - a_) < 1e-9 * a_ + 1e-12, k
    gen = torch.Generator().manual_seed(g["seed_data"])
    x_real = torch.rand(g["shape"], generator=gen) * 2 - 1
    x_fake = torch.rand(g["shape"], generator=gen) * 2 - 1
    ns = {}
    loss, d_real, d_fake = DO.disc_loss(sd, x_real, x_fake, 2 if kind == "img" else 3, n_frames=11, new_stats=ns)
    torch.testing.assert_close(loss, g["d_loss"], rtol=1e-6, atol=1e-8)
    for got, ref in zip(d_real, g["pred_real"]):
        torch.testing.assert_close(got[-1], ref, rtol=1e-5, atol=1e-6)
    for k, ref in g["stats_after"].items():
        if "num_batches" not in k:
            torch.testing.assert_close(ns[k], ref, rtol=1e-5, atol=1e-7)
    # the drop-in refuses to compute on the CPU
    from faceoff_b200 import _lib
    with pytest.raises((_lib.FaceoffB200Error, RuntimeError)):
        m(x_real)
