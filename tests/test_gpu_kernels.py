"""Isolated GPU tests of the non-GEMM kernels against a torch fp64 restatement on IDENTICAL (bf16-representable) inputs,
so that the bf16 storage of the surrounding conv stack cannot hide an error (run on the B200 box: ``pytest -m gpu``).

fp32-accumulating reductions: rtol 1e-5.  bf16 outputs: at most one bf16 rounding (2^-8 relative) away from the fp64 value.
Index / routing work (max-pool, gather): bit-exact.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SWEEP = [(64, 512), (64, 1024), (64, 2048), (128, 512), (128, 1024), (128, 2048)]   # BASELINE configs[3]


def _bf16(*shape, gen, scale=1.0, relu=False):
    x = torch.randn(*shape, generator=gen) * scale
    if relu:
        x = x.relu()
    return x.bfloat16()


def _one_ulp_close(got_bf16, ref64, extra=0.0):
    """|got - ref| <= one bf16 rounding of ref (2^-8 relative) + extra absolute slack."""
    got = got_bf16.double().cpu()
    tol = ref64.abs() * 2.0 ** -8 + extra
    bad = (got - ref64).abs() > tol
    assert not bad.any(), f"{bad.sum().item()} elements off by more than one bf16 rounding; worst " \
                          f"{((got - ref64).abs() / (ref64.abs() + 1e-30)).max().item():.3e}"


def _within(name, err, tol):
    bad = err > tol
    if bad.any():
        i = (err / tol).argmax()
        raise AssertionError(f"{name}: {bad.sum().item()} of {bad.numel()} elements out of tolerance; worst err "
                             f"{err.flatten()[i].item():.3e} vs allowed {tol.flatten()[i].item():.3e}")


# ------------------------------------------------------------------------------------------------ LPIPS head
def _lpips_tap_ref(f0, f1, w):
    """models/lpips.py:80-93,155-161 for one tap: normalise (eps outside the sqrt), squared diff, 1x1 lin, spatial mean."""
    def nrm(x):
        return x / (x.pow(2).sum(-1, keepdim=True).sqrt() + 1e-10)
    d = (nrm(f0) - nrm(f1)).pow(2)
    return (d * w).sum(-1).mean((1, 2))


@pytest.mark.parametrize("n,h,w,c", [(3, 16, 16, 64), (2, 8, 8, 128), (2, 8, 4, 256), (1, 4, 4, 512), (5, 2, 2, 512)])
def test_lpips_tap_forward_and_backward(n, h, w, c):
    from faceoff_b200 import ops

    gen = torch.Generator().manual_seed(c + n)
    f0 = _bf16(n, h, w, c, gen=gen, relu=True)
    f1 = _bf16(n, h, w, c, gen=gen, relu=True)
    f0[0, 0, 0] = 0          # an all-zero feature vector: the eps-outside-the-sqrt path
    lw = torch.rand(c, generator=gen) * 0.1
    g = torch.randn(n, generator=gen)
    add = _bf16(n, h, w, c, gen=gen, scale=0.01)
    out = torch.zeros(n, device="cuda")
    out += 0.25                                  # the kernel accumulates into out (five taps share it)
    ops.lpips_tap(f0.cuda(), f1.cuda(), lw.cuda(), out)
    a = f0.double().requires_grad_(True)
    ref = _lpips_tap_ref(a, f1.double(), lw.double())
    torch.testing.assert_close(out.cpu().double() - 0.25, ref.detach(), rtol=1e-5, atol=1e-7)
    (ref * g.double()).sum().backward()
    # The kernel returns the gradient w.r.t. the PRE-ReLU activation of the tap (gate f0 > 0; the pool gradient passed as
    # ``addend`` goes through the same gate).  At an all-zero feature vector torch's autograd of sqrt gives NaN (0/0); the
    # kernel defines that gradient as 0.
    zero_pix = (f0.double().pow(2).sum(-1, keepdim=True) == 0)
    assert zero_pix.any() and torch.isnan(a.grad[zero_pix.expand_as(a.grad)]).all()
    gate = (f0.double() > 0)
    for addend in (None, add):
        d = ops.lpips_tap_bwd(f0.cuda(), f1.cuda(), lw.cuda(), g.cuda(), None if addend is None else addend.cuda())
        want = torch.where(gate, torch.nan_to_num(a.grad) + (0 if addend is None else addend.double()),
                           torch.zeros((), dtype=torch.float64))
        assert torch.isfinite(d.float()).all()
        _one_ulp_close(d, want, extra=1e-9)


@pytest.mark.parametrize("n,h,w,c", [(3, 16, 16, 64), (2, 8, 12, 128), (2, 6, 4, 256), (1, 4, 4, 512), (5, 2, 2, 512), (2, 64, 64, 64)])
def test_lpips_tap_backward_with_the_pool_folded_in(n, h, w, c):
    """fo_lpips_tap_bwd_pool == fo_maxpool2_bwd followed by fo_lpips_tap_bwd(addend = its result), bit for bit: coarse
    features (many exact ties inside a pooling window, closed gates) exercise the first-maximum rule."""
    from faceoff_b200 import ops

    gen = torch.Generator().manual_seed(c + n + h)
    f0 = (torch.randn(n, h, w, c, generator=gen).relu() * 4).round().div(4).bfloat16().cuda()    # multiples of 1/4: ties
    f1 = _bf16(n, h, w, c, gen=gen, relu=True).cuda()
    lw = (torch.rand(c, generator=gen) * 0.1).cuda()
    g = torch.randn(n, generator=gen).cuda()
    pool_dy = _bf16(n, h // 2, w // 2, c, gen=gen, scale=0.01).cuda()
    y = ops.maxpool2(f0)
    want = ops.lpips_tap_bwd(f0, f1, lw, g, ops.maxpool2_bwd(f0, y, pool_dy))
    got = ops.lpips_tap_bwd_pool(f0, f1, lw, g, pool_dy)
    torch.cuda.synchronize()
    ties = (f0.view(n, h // 2, 2, w // 2, 2, c).amax((2, 4), keepdim=True) == f0.view(n, h // 2, 2, w // 2, 2, c)).sum((2, 4)) > 1
    assert ties.any(), "the test data must contain ties"
    assert torch.equal(got.view(torch.int16), want.view(torch.int16)), \
        f"{(got.view(torch.int16) != want.view(torch.int16)).sum().item()} elements differ"
    # forward: the tap that also writes the pooled tensor
    out_a = torch.zeros(n, device="cuda")
    out_b = torch.zeros(n, device="cuda")
    ops.lpips_tap(f0, f1, lw, out_a)
    pooled, none1 = ops.lpips_tap_pool(f0, f1, lw, out_b)
    out_c = torch.zeros(n, device="cuda")
    pooled_c, pooled1 = ops.lpips_tap_pool(f0, f1, lw, out_c, pool_f1=True)
    torch.cuda.synchronize()
    assert none1 is None
    assert torch.equal(pooled.view(torch.int16), y.view(torch.int16)), "pooled tensor differs from maxpool2"
    assert torch.equal(pooled_c.view(torch.int16), y.view(torch.int16))
    assert torch.equal(pooled1.view(torch.int16), ops.maxpool2(f1).view(torch.int16)), "pooled f1 differs from maxpool2"
    torch.testing.assert_close(out_b, out_a, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(out_c, out_a, rtol=1e-5, atol=1e-8)


# ------------------------------------------------------------------------------------------------ pooling / layout
@pytest.mark.parametrize("n,h,w,cs", [(2, 8, 8, 64), (1, 4, 6, 128), (3, 2, 2, 16)])
def test_maxpool2_and_backward_bit_exact(n, h, w, cs):
    from faceoff_b200 import ops

    gen = torch.Generator().manual_seed(h * w)
    x = _bf16(n, h, w, cs, gen=gen, relu=True)      # post-ReLU input: plenty of exact ties at 0
    dy = _bf16(n, h // 2, w // 2, cs, gen=gen)
    y = ops.maxpool2(x.cuda())
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    yr = F.max_pool2d(xr, 2, 2)
    assert torch.equal(y.float().cpu(), yr.detach().permute(0, 2, 3, 1))
    yr.backward(dy.float().permute(0, 3, 1, 2))
    # the kernel fuses the ReLU gate of the (post-ReLU) pool input: dx is the gradient w.r.t. the PRE-ReLU value, i.e.
    # torch's routing (first maximum of the window wins) times (x > 0)
    want = xr.grad.permute(0, 2, 3, 1) * (x.float() > 0)
    dx = ops.maxpool2_bwd(x.cuda(), y, dy.cuda())
    assert torch.equal(dx.float().cpu(), want), "max-pool gradient routing (first maximum wins) + ReLU gate"


@pytest.mark.parametrize("n,c,hi,wi", [(2, 6, 8, 8), (1, 3, 4, 16), (1, 8, 2, 2)])
def test_col2im4x4s2_matches_conv_transpose_scatter(n, c, hi, wi):
    """col[n,iy,ix,(ky*4+kx)*8+co] scattered to out[n,co,2iy-1+ky,2ix-1+kx] (+bias) == F.fold of the same columns."""
    from faceoff_b200 import ops

    gen = torch.Generator().manual_seed(wi)
    col = _bf16(n, hi, wi, 128, gen=gen)
    bias = torch.randn(c, generator=gen)
    out = ops.col2im4x4s2(col.cuda(), bias.cuda(), c)
    cols = col.double().view(n, hi * wi, 16, 8)[..., :c]                     # [n, L, taps, co]
    cols = cols.permute(0, 3, 2, 1).reshape(n, c * 16, hi * wi)               # fold wants [n, co*taps, L]
    ref = F.fold(cols, (2 * hi, 2 * wi), kernel_size=4, stride=2, padding=1) + bias.double().view(1, c, 1, 1)
    torch.testing.assert_close(out.cpu().double(), ref, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("n,c,h,w", [(2, 6, 16, 16), (1, 3, 8, 32)])
def test_im2col4x4s2_matches_unfold(n, c, h, w):
    from faceoff_b200 import ops

    gen = torch.Generator().manual_seed(h)
    x = torch.randn(n, c, h, w, generator=gen)
    col = ops.im2col4x4s2(x.cuda(), c).float().cpu()                          # [n, h/2, w/2, 16 taps x 8]
    ref = F.unfold(x.bfloat16().float(), kernel_size=4, stride=2, padding=1)  # [n, c*16, L]
    ref = ref.view(n, c, 16, (h // 2) * (w // 2)).permute(0, 3, 2, 1)         # [n, L, taps, c]
    got = col.view(n, (h // 2) * (w // 2), 16, 8)
    assert torch.equal(got[..., :c], ref) and got[..., c:].abs().max().item() == 0.0


def test_im2col3x3_with_scaling_layer():
    """First VGG conv operand: (x - shift) / scale folded into the column matrix (models/lpips.py:96-103)."""
    from faceoff_b200 import ops

    gen = torch.Generator().manual_seed(2)
    x = torch.rand(2, 3, 8, 12, generator=gen) * 2 - 1
    shift = torch.tensor([-.030, -.088, -.188])
    scale = torch.tensor([.458, .448, .450])
    col = ops.im2col3x3(x.cuda(), shift.cuda(), scale.cuda()).float().cpu()   # [n, h, w, 32]; k = (ky*3+kx)*3 + ch
    xs = (x - shift.view(1, 3, 1, 1)) / scale.view(1, 3, 1, 1)
    ref = F.unfold(xs, kernel_size=3, padding=1).view(2, 3, 9, 8 * 12).permute(0, 3, 2, 1).reshape(2, 8, 12, 27)
    torch.testing.assert_close(col[..., :27], ref.bfloat16().float(), rtol=2.0 ** -7, atol=1e-6)
    assert col[..., 27:].abs().max().item() == 0.0


@pytest.mark.parametrize("n,h,w,scaled", [(2, 6, 256, True), (1, 5, 128, True), (3, 3, 200, False), (1, 1, 384, True)])
def test_vgg_first_conv_fused_kernel(n, h, w, scaled):
    """ScalingLayer + Conv2d(3 -> 64, 3x3, pad 1) + ReLU in one kernel (csrc/small_cin.cu, reference models/lpips.py:96-103,
    119-127) against fp64 on the operands the tensor core sees (scaled image and weights rounded to bf16): products are
    exact, accumulation fp32, output one bf16 rounding.  Widths that are not multiples of the 128-pixel tile included."""
    from faceoff_b200 import ops

    gen = torch.Generator().manual_seed(n * 1000 + h * 10 + w)
    x = torch.rand(n, 3, h, w, generator=gen) * 2 - 1
    wt = torch.randn(64, 3, 3, 3, generator=gen) * 0.2
    b = torch.randn(64, generator=gen) * 0.1
    shift = torch.tensor([-.030, -.088, -.188])
    scale = torch.tensor([.458, .448, .450])
    got = ops.vgg_first_conv(x.cuda(), wt.cuda(), b.cuda(), shift.cuda() if scaled else None, scale.cuda() if scaled else None)
    xs = ((x - shift.view(1, 3, 1, 1)) / scale.view(1, 3, 1, 1)) if scaled else x
    ref = F.conv2d(xs.bfloat16().double(), wt.bfloat16().double(), b.double(), padding=1).relu().permute(0, 2, 3, 1)
    assert got.shape == (n, h, w, 64) and got.dtype == torch.bfloat16
    _one_ulp_close(got, ref, extra=2e-6)


@pytest.mark.parametrize("n,c,ca,h,w", [(2, 6, 6, 8, 256), (1, 3, 3, 6, 512), (3, 6, 8, 4, 200), (2, 6, 6, 10, 16), (1, 6, 6, 2, 1024)])
def test_s2conv_and_s2wgrad_without_im2col(n, c, ca, h, w):
    """Image-side stride-2 layers without an im2col matrix (csrc/small_cin.cu) against fp64 on the bf16-rounded operands:
    forward (+bias, ReLU), the gated / accumulated data-gradient form, and the weight + bias gradient.  Widths that are not
    multiples of the 128-pixel tile, several tiles per row, more storage channels than used."""
    from faceoff_b200 import ops

    gen = torch.Generator().manual_seed(n * 100 + c * 10 + h + w)
    x = torch.randn(n, ca, h, w, generator=gen)
    wt = torch.randn(64, c, 4, 4, generator=gen) * 0.1
    b = torch.randn(64, generator=gen) * 0.1
    xs = x[:, :c].bfloat16().double()
    ref = F.conv2d(xs, wt.bfloat16().double(), None, stride=2, padding=1).permute(0, 2, 3, 1)      # [n, h/2, w/2, 64]
    got = ops.s2conv(x.cuda(), c, wt.cuda(), b.cuda(), relu=True)
    _one_ulp_close(got, (ref + b.double()).relu(), extra=2e-6)
    mask = _bf16(n, h // 2, w // 2, 64, gen=gen)
    add = _bf16(n, h // 2, w // 2, 64, gen=gen)
    got = ops.s2conv(x.cuda(), c, wt.cuda(), None, mask=mask.cuda(), addend=add.cuda())
    _one_ulp_close(got, torch.where(mask.double() > 0, ref, torch.zeros_like(ref)) + add.double(), extra=2e-6)
    y = _bf16(n, h // 2, w // 2, 64, gen=gen)
    dw = torch.full((64, c, 4, 4), 7.0).cuda()
    db = torch.full((64,), -3.0).cuda()
    ops.s2wgrad(x.cuda(), c, y.cuda(), dw, dbias=db)
    cols = F.unfold(xs, kernel_size=4, stride=2, padding=1)                                          # [n, c*16, L]
    dw_ref = torch.einsum("nlo,nkl->ok", y.double().reshape(n, -1, 64), cols).reshape(64, c, 4, 4)
    abs_ref = torch.einsum("nlo,nkl->ok", y.double().abs().reshape(n, -1, 64), cols.abs()).reshape(64, c, 4, 4)
    _within("dweight", (dw.cpu().double() - dw_ref).abs(), 1e-5 * dw_ref.abs() + 3e-7 * abs_ref + 1e-6)
    db_ref = y.double().sum((0, 1, 2))
    _within("dbias", (db.cpu().double() - db_ref).abs(), 1e-5 * db_ref.abs() + 3e-7 * y.double().abs().sum((0, 1, 2)) + 1e-6)
    dw2, db2 = dw.clone(), db.clone()
    ops.s2wgrad(x.cuda(), c, y.cuda(), dw2, accumulate=True, dbias=db2, dbias_accumulate=True)
    torch.testing.assert_close(dw2, 2 * dw, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(db2, 2 * db, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("rows,cs,c_off,c", [(4096, 128, 0, 128), (1000, 192, 64, 64), (7, 32, 0, 6), (70000, 64, 0, 64)])
def test_colsum_bias_gradient(rows, cs, c_off, c):
    from faceoff_b200 import ops

    gen = torch.Generator().manual_seed(rows)
    x = _bf16(rows, cs, gen=gen)
    out = torch.full((c,), 3.0, device="cuda")
    ops.colsum(x.cuda().view(rows, 1, 1, cs), c, out, c_off=c_off, accumulate=True)
    ref = x.double()[:, c_off:c_off + c].sum(0) + 3.0
    torch.testing.assert_close(out.cpu().double(), ref, rtol=1e-5, atol=1e-4)
    ops.colsum(x.cuda().view(rows, 1, 1, cs), c, out, c_off=c_off, accumulate=False)
    torch.testing.assert_close(out.cpu().double(), ref - 3.0, rtol=1e-5, atol=1e-4)


def test_pack_unpack_nchw_roundtrip():
    from faceoff_b200 import ops

    gen = torch.Generator().manual_seed(0)
    x = torch.randn(3, 6, 8, 8, generator=gen)
    p = ops.pack_nchw(x.cuda())
    assert p.shape == (3, 8, 8, 16) and p[..., 6:].abs().max().item() == 0
    assert torch.equal(p[..., :6].float().cpu(), x.bfloat16().float().permute(0, 2, 3, 1))
    assert torch.equal(ops.unpack_nchw(p, 6).cpu(), x.bfloat16().float())


def test_process_data_u8_bit_exact_vs_torchvision_transforms():
    """uint8 frames -> normalised fp32 on the GPU == the reference loader's ToTensor + Normalize(0.5, 0.5)
    (TemporalAlignment/dataset.py:235-249) followed by utils.process_data's channel concat (utils.py:29-38), bit for bit.
    All 256 byte values occur."""
    from faceoff_b200.data import process_data_u8

    gen = torch.Generator().manual_seed(0)
    T, H, W = 3, 16, 24
    frames = [torch.randint(0, 256, (T, H, W, 3), generator=gen, dtype=torch.uint8) for _ in range(3)]
    frames[0].view(-1)[:256] = torch.arange(256, dtype=torch.uint8)

    def reference(fr):       # torchvision.transforms.functional.to_tensor + normalize, per frame, then vstack
        try:
            from torchvision import transforms

            tf = transforms.Compose([transforms.ToPILImage(), transforms.ToTensor(),
                                     transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])
            return torch.vstack([tf(f.numpy()).unsqueeze(0) for f in fr])
        except ImportError:  # no PIL: the same two functional steps written out
            t = fr.permute(0, 3, 1, 2).to(torch.float32).div(255)
            return t.sub_(0.5).div_(0.5)

    src, bg, gtf = (reference(f) for f in frames)
    img_ref = torch.cat([src, bg], 1)
    img, S, gt = process_data_u8(frames[0], frames[1], frames[2], device="cuda")
    assert S == T and img.shape == (T, 6, H, W) and gt.shape == (T, 3, H, W)
    assert torch.equal(img.cpu(), img_ref) and torch.equal(gt.cpu(), gtf)


# ------------------------------------------------------------------------------------------------ quantiser sweep
@pytest.mark.parametrize("dim,n_embed", SWEEP)
def test_vq_assign_bit_exact_on_the_codebook_sweep(dim, n_embed):
    """BASELINE configs[3]: every (embed_dim, n_embed) of the sweep, indices == fp64 argmin (first minimum)."""
    from faceoff_b200 import ops

    gen = torch.Generator().manual_seed(dim + n_embed)
    rows = 40000
    x = torch.randn(rows, dim, generator=gen)
    e = torch.randn(dim, n_embed, generator=gen)
    e_split, e_t, e_n2 = ops.vq_prep(e.cuda())
    ind = ops.vq_assign(x.cuda(), e_t, e_split, e_n2).cpu()
    d = x.double().pow(2).sum(1, keepdim=True) - 2 * x.double() @ e.double() + e.double().pow(2).sum(0, keepdim=True)
    assert torch.equal(ind, d.argmin(1))


@pytest.mark.parametrize("dim,n_embed", SWEEP)
@pytest.mark.parametrize("skew", [False, True])
def test_vq_gather_stats_ema_backward_on_the_codebook_sweep(dim, n_embed, skew):
    """gather + straight-through + diff + EMA statistics + EMA update + backward vs the oracle (fp32 path: rtol 1e-5),
    on every sweep shape (shared-memory and global-atomic variants), uniform and collapsed (3 hot codes) assignments."""
    from faceoff_b200 import ops
    from oracle import faceoff_oracle as O

    gen = torch.Generator().manual_seed(dim * 3 + n_embed)
    rows = 20000
    x = torch.randn(rows, dim, generator=gen)
    e = torch.randn(dim, n_embed, generator=gen)
    ind = torch.randint(0, 3 if skew else n_embed, (rows,), generator=gen)
    cs0 = torch.rand(n_embed, generator=gen) * 5
    ea0 = e * cs0 + 0.1 * torch.randn(dim, n_embed, generator=gen)
    xs, inds = x.cuda(), ind.cuda()
    e_t = e.t().contiguous().cuda()
    diff_sum = torch.zeros(1, device="cuda")
    counts = torch.zeros(n_embed, device="cuda")
    esum = torch.zeros(dim, n_embed, device="cuda")
    q32, q16 = ops.vq_gather_stats(xs, inds, e_t, diff_sum, counts, esum, want_f32=True, want_bf16=True)
    qref = e.t()[ind]
    st = x + (qref - x)                                             # :78, evaluated in fp32 like the reference
    assert torch.equal(q32.cpu(), st)
    assert torch.equal(q16.float().cpu(), st.bfloat16().float())
    torch.testing.assert_close(diff_sum.cpu().double() / x.numel(), (qref.double() - x.double()).pow(2).mean().reshape(1),
                               rtol=1e-5, atol=1e-9)
    c_ref, s_ref = O.quantize_stats(x.double(), ind, n_embed)
    assert torch.equal(counts.cpu().double(), c_ref), "counts are exact integers"
    # fp32 accumulation (any order) of n terms: |err| <= ~eps * sum|x_i|; the reference's fp32 matmul (:61) has the same bound
    abs_sum = (F.one_hot(ind, n_embed).double().t() @ x.double().abs()).t()
    err = (esum.cpu().double() - s_ref).abs()
    _within("embed_sum", err, 1e-5 * s_ref.abs() + 3e-7 * abs_sum + 1e-6)
    # EMA with the (exact) fp64 statistics rounded to fp32
    emb, cs, ea = e.clone().cuda(), cs0.clone().cuda(), ea0.clone().cuda()
    ops.vq_ema(emb, cs, ea, c_ref.float().cuda(), s_ref.float().cuda(), 0.99, 1e-5)
    e1, c1, a1 = O.quantize_ema(e.double(), cs0.double(), ea0.double(), c_ref, s_ref, 0.99, 1e-5)
    torch.testing.assert_close(cs.cpu().double(), c1, rtol=1e-5, atol=1e-7)
    # a1 = 0.99 ea0 + 0.01 s may cancel: the fp32 rounding is relative to the two terms, not to their sum
    mag = 0.99 * ea0.double().abs() + 0.01 * s_ref.abs()
    _within("embed_avg", (ea.cpu().double() - a1).abs(), 1e-5 * a1.abs() + 2e-7 * mag + 1e-9)
    csn = (c1 + 1e-5) / (c1.sum() + n_embed * 1e-5) * c1.sum()
    _within("embed", (emb.cpu().double() - e1).abs(), 2e-5 * e1.abs() + 4e-7 * mag / csn + 1e-9)
    # backward: gx = g_q + g_diff * 2 (x - q) / numel   (autograd of :77-78)
    gq = torch.randn(rows, dim, generator=gen)
    gd = torch.tensor([0.7])
    g32, _ = ops.vq_backward(gq.cuda(), 0, gd.cuda(), xs, inds, e_t)
    gref = gq.double() + 0.7 * 2 * (x.double() - qref.double()) / x.numel()
    torch.testing.assert_close(g32.cpu().double(), gref, rtol=1e-5, atol=1e-9)


# ------------------------------------------------------------------------------------------------ weight gradients
@pytest.mark.parametrize("clips", [1, 32])
def test_wgrad_accumulation_chain_is_bounded(clips):
    """Conv3d 128 -> 128 weight gradient at the BASELINE batch (32 clips x 30 x 64 x 64 positions in ONE launch) against an
    fp64 GEMM of the same bf16 operands: centre tap dW[:, :, 1, 1, 1] = dy^T x.  The tensor core truncates when it adds to
    its fp32 accumulator (csrc/wgrad_igemm.cu "accumulation chains"): without the chunked chains the 32-clip launch is
    2.5e-4 short in norm (3.0e-4 max-normalised); with them it must stay below 1e-4, like a 1-clip launch.  Also checks
    the fused bias gradient (column sums of dy) at rtol 1e-5."""
    from faceoff_b200 import ops
    from faceoff_b200.ops import FORM_S1

    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn(clips, 30, 64, 64, 128, device=dev, generator=g).to(torch.bfloat16)
    dy = (torch.randn(clips, 30, 64, 64, 128, device=dev, generator=g) * 1e-3).to(torch.bfloat16)
    ref = torch.zeros(128, 128, dtype=torch.float64, device=dev)
    bref = torch.zeros(128, dtype=torch.float64, device=dev)
    for c in range(clips):
        d = dy[c].reshape(-1, 128).double()
        ref += d.t() @ x[c].reshape(-1, 128).double()
        bref += d.sum(0)
    dw = torch.empty(128, 128, 3, 3, 3, dtype=torch.float32, device=dev)
    db = torch.empty(128, dtype=torch.float32, device=dev)
    ops.wgrad(FORM_S1, 3, 3, (dy, 128, 0), (x, 128, 0), dw, m_axis=0, dbias=db)
    torch.cuda.synchronize()
    a = dw[:, :, 1, 1, 1].double()
    e_max = ((a - ref).abs().max() / ref.abs().max()).item()
    e_norm = ((a.norm() - ref.norm()) / ref.norm()).item()
    print(f"wgrad3d {clips} clips: max-normalised err {e_max:.3e}, norm rel err {e_norm:+.3e}")
    assert e_max < 1e-4 and abs(e_norm) < 1e-4
    assert ((db.double() - bref).abs().max() / bref.abs().max()).item() < 1e-4


@pytest.mark.parametrize("n,h,w,scaled", [(2, 12, 64, True), (1, 7, 33, True), (3, 5, 200, False), (1, 256, 256, True), (2, 1, 30, False)])
def test_vgg_first_dgrad_taps_on_n_axis(n, h, w, scaled):
    """Data gradient of Conv2d(3 -> 64, 3x3, pad 1) (csrc/small_cin.cu vgg_first_dgrad_kernel: one K = 64 GEMM per pixel tile
    with the taps on the N axis + in-tile shift-add, fp32 NCHW out, ScalingLayer division folded in) against fp64 on the
    operands the tensor core sees (dy bf16, weights rounded to bf16): products exact, fp32 accumulation.  Sizes that are not
    multiples of the 6 x 30 output tile included."""
    from faceoff_b200 import ops

    gen = torch.Generator().manual_seed(n * 1000 + h * 10 + w)
    dy = (torch.randn(n, h, w, 64, generator=gen) * 0.1).bfloat16()
    wt = torch.randn(64, 3, 3, 3, generator=gen) * 0.2
    scale = torch.tensor([.458, .448, .450])
    got = ops.vgg_first_dgrad(dy.cuda(), wt.cuda(), scale.cuda() if scaled else None)
    ref = F.conv_transpose2d(dy.double().permute(0, 3, 1, 2), wt.bfloat16().double(), padding=1)
    if scaled:
        ref = ref / scale.double().view(1, 3, 1, 1)
    assert got.shape == (n, 3, h, w) and got.dtype == torch.float32
    err = (got.double().cpu() - ref).abs().max().item() / ref.abs().max().item()
    print(f"vgg_first_dgrad {n}x{h}x{w}: max-normalised err {err:.3e}")
    assert err < 2e-6
