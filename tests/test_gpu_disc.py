"""GPU parity tests of the MoCoGAN-HD discriminator step (BASELINE configs[4]) -- ``pytest -m gpu``.
The CUDA path (faceoff_b200.mocoganhd through the C ABI) against the committed outputs of the UNMODIFIED reference classes
(tests/golden/golden_disc.pt) and against the oracle, in both convolution modes:

* ``tensor`` (default product path): im2col + split-bf16 GEMMs on the tcgen05 kernel (csrc/disc_gemm.cu).  Per layer
  ~2e-5 / 5e-6 / 1e-5 (y / dx / dw, max-normalised against fp64).  End to end the predictions and losses hold the same
  1e-4 as the fp32 mode; individual weight-gradient ELEMENTS may move by up to a few per cent of the tensor's maximum,
  because a forward difference of 2e-5 puts some LeakyReLU inputs on the other side of zero (the fp32 mode itself is 1.6e-3
  away from an fp64 run on one of these tensors: tests/gpu_disc_fp64_check.py) -- the norms of the gradients still agree
  to 2e-3.
* ``fp32``: the FFMA kernels of csrc/disc.cu, rtol 1e-4 throughout."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "golden_disc.pt")


def _mn(a, b):
    """max-normalised error (the fp32 GPU kernels and the reference's CPU kernels sum K up to 32768 products in different
    orders: elements that cancel to ~0 cannot be held to an element-wise rtol)."""
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-300)).item()


import contextlib  # noqa: E402


@contextlib.contextmanager
def _mode(mode):
    from faceoff_b200.mocoganhd import layers
    old = layers.TENSOR_CORE
    layers.TENSOR_CORE = mode == "tensor"
    try:
        yield
    finally:
        layers.TENSOR_CORE = old


def _build(kind):
    from faceoff_b200.mocoganhd import content_disc, video_disc

    g = torch.load(GOLDEN, map_location="cpu")[kind]
    torch.manual_seed(g["seed_model"])
    m = content_disc.ModelD_img(3, "instance", 2, 1e-4) if kind == "img" else video_disc.ModelD_3d(3, "instance", 2, 1e-4, False, 12)
    for k, v in m.state_dict().items():     # same seed => the reference's initial weights (checksums from the reference)
        if v.dtype.is_floating_point:
            s, a = g["param_checksums"][k]
            assert abs(v.double().sum().item() - s) <= 1e-9 + 1e-12 * abs(a) and abs(v.double().abs().sum().item() - a) <= 1e-9 * a + 1e-12, k
    gen = torch.Generator().manual_seed(g["seed_data"])
    x_real = torch.rand(g["shape"], generator=gen) * 2 - 1
    x_fake = torch.rand(g["shape"], generator=gen) * 2 - 1
    return g, m, x_real, x_fake


@pytest.mark.parametrize("mode", ["tensor", "fp32"])
@pytest.mark.parametrize("kind", ["img", "vid"])
def test_discriminator_step_matches_reference_golden(kind, mode):
    """Discriminator step (trainer :240-300): fake then real forward in training mode, relativistic average LSGAN loss,
    backward: patch predictions, loss, every parameter gradient, InstanceNorm running statistics."""
    with _mode(mode):
        _golden_step(kind, mode)


def _golden_step(kind, mode):
    from faceoff_b200.mocoganhd import losses

    tensor = mode == "tensor"
    g, m, x_real, x_fake = _build(kind)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.cuda().train()
    crit = losses.Relativistic_Average_LSGAN()
    d_fake = m(x_fake.cuda())
    d_real = m(x_real.cuda())
    d_loss = (crit(d_real, d_fake, True) + crit(d_fake, d_real, False)) * 0.5
    m.zero_grad()
    d_loss.backward()
    torch.cuda.synchronize()
    assert len(d_real) == 2 and len(d_real[0]) == 5
    for got, ref in zip(d_real, g["pred_real"]):
        assert _mn(got[-1].detach().cpu(), ref) < 2e-4
    for got, ref in zip(d_fake, g["pred_fake"]):
        assert _mn(got[-1].detach().cpu(), ref) < 2e-4
    assert _mn(d_real[1][1].detach().cpu()[:, :4], g["feat_real_scale0_layer1"]) < (2e-4 if tensor else 1e-4)
    torch.testing.assert_close(d_loss.detach().cpu(), g["d_loss"], rtol=1e-4, atol=1e-7)
    worst = 0.0
    for k, p in m.named_parameters():
        n_ref = g["grad_norms"][k].item()
        layer = int(k.split("_layer")[1][0])
        if k.endswith(".0.bias") and layer >= 1:
            # a bias in front of InstanceNorm has no effect on the loss, and the bias of the last conv shifts the real and
            # the fake prediction alike (relativistic loss): these gradients are pure rounding noise (reference: ~1e-6)
            w_norm = g["grad_norms"][k.replace(".bias", ".weight")].item()
            assert p.grad.norm().item() <= 1e-3 * w_norm and n_ref <= 1e-3 * w_norm, k
            continue
        rel = abs(p.grad.norm().item() - n_ref) / (n_ref + 1e-30)
        worst = max(worst, rel)
        sl = p.grad.flatten()[:64].cpu()
        err = (sl - g["grad_slices"][k]).abs().max().item()
        # tensor mode: single elements move when a LeakyReLU gate flips (module docstring); the norm check above holds both
        assert err <= (5e-2 if tensor else 1e-3) * g["grad_slices"][k].abs().max().item() + 1e-5 * n_ref, \
            (k, err, g["grad_slices"][k].abs().max().item(), n_ref, rel)
    print(f"{kind} [{mode}]: worst gradient-norm relative error {worst:.2e}")
    assert worst < (2e-3 if tensor else 5e-4)
    for k, ref in g["stats_after"].items():
        got = m.state_dict()[k].cpu()
        if "num_batches" in k:
            assert got.item() == ref.item(), k
        else:
            torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-6)
    # generator-side loss and its gradient w.r.t. the fake input, from the initial state (trainer :208-227)
    m.load_state_dict(sd0)
    xf = x_fake.cuda().requires_grad_(True)
    df = m(xf)
    dr = m(x_real.cuda())
    g_loss = (crit(df, dr, True) + crit(dr, df, False)) * 0.5
    g_loss.backward()
    torch.testing.assert_close(g_loss.detach().cpu(), g["g_loss"], rtol=1e-4, atol=1e-7)
    ref_gx = g["grad_x_fake"]
    err = ((xf.grad.cpu() - ref_gx).abs().max() / ref_gx.abs().max()).item()
    print(f"{kind} [{mode}]: d(G loss)/d(x_fake) max-normalised error {err:.2e}")
    assert err < (2e-2 if tensor else 5e-4)


@pytest.mark.parametrize("mode", ["tensor", "fp32"])
@pytest.mark.parametrize("kind", ["img", "vid"])
def test_discriminator_eval_mode_uses_running_statistics(kind, mode):
    with _mode(mode):
        _eval_mode(kind)


def _eval_mode(kind):
    from oracle import disc_oracle as DO

    g, m, x_real, _ = _build(kind)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    for k in sd:    # realistic running statistics: the reference's estimates after two training forwards
        if "running" in k:
            sd[k] = g["stats_after"][k].clone()
    m.load_state_dict(sd)
    m = m.cuda().eval()
    with torch.no_grad():
        out = m(x_real.cuda())
    ref = DO.multiscale_forward(sd, x_real, 2 if kind == "img" else 3, n_frames=11, training=False)
    for a, b in zip(out, ref):
        for u, v in zip(a, b):
            assert _mn(u.cpu(), v) < 2e-4
    for k, v in m.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), f"eval forward modified {k}"


@pytest.mark.parametrize("mode", ["tensor", "fp32"])
@pytest.mark.parametrize("shape,cout,k,s,p", [((2, 5, 17, 13), 7, 4, 2, 2), ((1, 6, 9, 9), 64, 4, 1, 2), ((1, 3, 5, 11, 9), 4, 4, 2, 2),
                                              ((2, 4, 3, 6, 7), 5, 4, 1, 2), ((1, 70, 8, 8), 130, 3, 1, 1),
                                              ((1, 64, 33, 33), 128, 4, 2, 2), ((2, 32, 3, 17, 17), 64, 4, 1, 2),
                                              ((1, 256, 18, 18), 1, 4, 1, 2), ((1, 6, 5, 64, 64), 64, 4, 2, 2)])
def test_dconv_forward_dgrad_wgrad_vs_torch_fp64(shape, cout, k, s, p, mode):
    """Both convolution paths on odd sizes / ragged channel counts (partially filled tiles, K padded to 128, several K and
    position chunks, 1-channel outputs) vs torch fp64.  fp32 mode: rtol 1e-4.  tensor mode: max-normalised 6e-5 / 2e-5 /
    6e-5 for y / dx / dw (the tensor core's chained fp32 accumulation)."""
    with _mode(mode):
        _dconv_vs_fp64(shape, cout, k, s, p, mode)


def _dconv_vs_fp64(shape, cout, k, s, p, mode):
    from faceoff_b200.mocoganhd import layers

    gen = torch.Generator().manual_seed(sum(shape) + cout)
    nd = len(shape) - 2
    conv = (layers.Conv2d if nd == 2 else layers.Conv3d)(shape[1], cout, k, stride=s, padding=p).cuda()
    x = torch.randn(shape, generator=gen).cuda().requires_grad_(True)
    y = conv(x)
    go = torch.randn(y.shape, generator=gen).cuda()
    y.backward(go)
    x64 = x.detach().double().cpu().requires_grad_(True)
    w64 = conv.weight.detach().double().cpu().requires_grad_(True)
    b64 = conv.bias.detach().double().cpu().requires_grad_(True)
    r = (F.conv2d if nd == 2 else F.conv3d)(x64, w64, b64, stride=s, padding=p)
    r.backward(go.double().cpu())
    if mode == "tensor":
        assert _mn(y.detach().cpu(), r.detach()) < 6e-5
        assert _mn(x.grad.cpu(), x64.grad) < 2e-5
        assert _mn(conv.weight.grad.cpu(), w64.grad) < 6e-5
        assert _mn(conv.bias.grad.cpu(), b64.grad) < 1e-5
        return
    torch.testing.assert_close(y.detach().cpu().double(), r.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(x.grad.cpu().double(), x64.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(conv.weight.grad.cpu().double(), w64.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(conv.bias.grad.cpu().double(), b64.grad, rtol=1e-4, atol=1e-4)


def test_avgpool_instnorm_lrelu_vs_torch():
    from faceoff_b200.mocoganhd import layers

    gen = torch.Generator().manual_seed(0)
    for shape, mod, ref in (((2, 3, 9, 11), layers.AvgPool2d(3, stride=2, padding=[1, 1], count_include_pad=False),
                             lambda t: F.avg_pool2d(t, 3, 2, [1, 1], count_include_pad=False)),
                            ((1, 2, 5, 8, 7), layers.AvgPool3d(3, stride=[1, 2, 2], padding=[1, 1, 1], count_include_pad=False),
                             lambda t: F.avg_pool3d(t, 3, [1, 2, 2], [1, 1, 1], count_include_pad=False)),
                            ((1, 2, 6, 8, 8), layers.AvgPool3d(3, stride=2, padding=[1, 1, 1], count_include_pad=False),
                             lambda t: F.avg_pool3d(t, 3, 2, [1, 1, 1], count_include_pad=False))):
        x = torch.randn(shape, generator=gen)
        xc = x.cuda().requires_grad_(True)
        y = mod(xc)
        go = torch.randn(y.shape, generator=gen)
        y.backward(go.cuda())
        xr = x.double().requires_grad_(True)
        r = ref(xr)
        r.backward(go.double())
        torch.testing.assert_close(y.detach().cpu().double(), r.detach(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(xc.grad.cpu().double(), xr.grad, rtol=1e-5, atol=1e-6)
    # InstanceNorm (training, running statistics) followed by LeakyReLU
    norm = layers.InstanceNorm3d(5, affine=False, track_running_stats=True).cuda().train()
    act = layers.LeakyReLU(0.2, True)
    x = torch.randn(2, 5, 3, 6, 7, generator=gen) * 3 + 1
    xc = x.cuda().requires_grad_(True)
    y = act(norm(xc))
    go = torch.randn(y.shape, generator=gen)
    y.backward(go.cuda())
    xr = x.double().requires_grad_(True)
    rm, rv = torch.zeros(5, dtype=torch.float64), torch.ones(5, dtype=torch.float64)
    r = F.leaky_relu(F.instance_norm(xr, rm, rv, use_input_stats=True, momentum=0.1, eps=1e-5), 0.2)
    r.backward(go.double())
    torch.testing.assert_close(y.detach().cpu().double(), r.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(xc.grad.cpu().double(), xr.grad, rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(norm.running_mean.cpu().double(), rm, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(norm.running_var.cpu().double(), rv, rtol=1e-5, atol=1e-6)
    assert norm.num_batches_tracked.item() == 0      # like nn.InstanceNorm3d: the counter is never advanced
