"""CPU oracle for the FaceOff VQVAE-conv3d (+LPIPS) training-step hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``faceoff_b200/`` may import this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs do, and there only as the checker / the reported CPU baseline.

What it is: a functional restatement (torch CPU ops, fp32 or fp64) of the reference's
algorithm, each function citing the reference file:line it follows.  The arithmetic of the
reference lives in a third-party dependency that is not vendored under /root/reference:
PyTorch (pinned torch==1.13.1 / torchvision==0.14.1, reference environment.yml:66,69); the
container has torch 2.11 with the same operator semantics for every op used here
(conv2d, conv_transpose2d, conv3d, matmul, max, one_hot, embedding, relu, max_pool2d).

Parity pin: the reference ships no tests / golden vectors.  The pin is
therefore the reference itself, imported live in the build container:
``tests/golden/make_golden.py`` runs the unmodified reference modules and this oracle on the same
seeded inputs/weights and commits the reference's outputs as fixtures;
``tests/test_oracle_vs_golden.py`` re-checks this oracle against those fixtures everywhere
(including the GPU box, where /root/reference does not exist).

Batched ("batch B clips") semantics follow SURVEY.md section 8(e): the reference can only push one clip
through ``VQVAE.forward`` (``unsqueeze(0)`` at models/vqvae_conv3d_latent.py:247); its multi-clip
semantics are DDP over clips with SUM-all-reduced EMA statistics.  ``vqvae_forward(n_clips=B)``
reproduces exactly that by reshaping to [B,128,T,H,W] before the Conv3d stacks.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# ----------------------------------------------------------------------------------------------
# Parameter containers: plain dicts keyed exactly like the reference state_dict
# ----------------------------------------------------------------------------------------------


def _conv_init(gen: torch.Generator, cout: int, cin: int, *k: int, transposed: bool = False):
    """torch's default Conv init (kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)))."""
    shape = (cin, cout, *k) if transposed else (cout, cin, *k)
    fan_in = shape[1] * math.prod(k)
    bound = 1.0 / math.sqrt(fan_in)
    w = (torch.rand(shape, generator=gen) * 2 - 1) * bound
    b = (torch.rand(cout, generator=gen) * 2 - 1) * bound
    return w, b


def init_vqvae_params(seed: int = 0, in_channel: int = 6, channel: int = 128, n_res_block: int = 2,
                      n_res_channel: int = 32, embed_dim: int = 64, n_embed: int = 512) -> Dict[str, Tensor]:
    """Random weights with the reference's key names and shapes
    (models/vqvae_conv3d_latent.py:193-231).  Not bit-identical to nn.Module default init order;
    parity tests copy ONE state_dict into both implementations."""
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, Tensor] = {}

    def put(name, wb):
        p[name + ".weight"], p[name + ".bias"] = wb

    def resblocks(prefix, start):
        for i in range(n_res_block):
            put(f"{prefix}.blocks.{start + i}.conv.1", _conv_init(g, n_res_channel, channel, 3, 3))
            put(f"{prefix}.blocks.{start + i}.conv.3", _conv_init(g, channel, n_res_channel, 1, 1))

    # enc_b: Encoder(in, channel, stride 4)  :107-114
    put("enc_b.blocks.0", _conv_init(g, channel // 2, in_channel, 4, 4))
    put("enc_b.blocks.2", _conv_init(g, channel, channel // 2, 4, 4))
    put("enc_b.blocks.4", _conv_init(g, channel, channel, 3, 3))
    resblocks("enc_b", 5)
    # enc_t: Encoder(channel, channel, stride 2)  :116-121
    put("enc_t.blocks.0", _conv_init(g, channel // 2, channel, 4, 4))
    put("enc_t.blocks.2", _conv_init(g, channel, channel // 2, 3, 3))
    resblocks("enc_t", 3)
    put("quantize_conv_t", _conv_init(g, embed_dim, channel, 1, 1))
    # dec_t: Decoder(embed_dim, embed_dim, channel, stride 2)  :139-160
    put("dec_t.blocks.0", _conv_init(g, channel, embed_dim, 3, 3))
    resblocks("dec_t", 1)
    put(f"dec_t.blocks.{2 + n_res_block}", _conv_init(g, embed_dim, channel, 4, 4, transposed=True))
    put("quantize_conv_b", _conv_init(g, embed_dim, embed_dim + channel, 1, 1))
    put("upsample_t", _conv_init(g, embed_dim, embed_dim, 4, 4, transposed=True))
    # dec: Decoder(2*embed_dim, in_channel, channel, stride 4)
    put("dec.blocks.0", _conv_init(g, channel, embed_dim + embed_dim, 3, 3))
    resblocks("dec", 1)
    put(f"dec.blocks.{2 + n_res_block}", _conv_init(g, channel // 2, channel, 4, 4, transposed=True))
    put(f"dec.blocks.{4 + n_res_block}", _conv_init(g, in_channel, channel // 2, 4, 4, transposed=True))
    for stack in ("conv3d_encoded_b", "conv3d_encoded_t"):  # hard-wired 128 channels :230-231
        for i in range(3):
            put(f"{stack}.conv3d.{i}.0", _conv_init(g, 128, 128, 3, 3, 3))
    for q in ("quantize_t", "quantize_b"):  # :42-45
        e = torch.randn(embed_dim, n_embed, generator=g)
        p[f"{q}.embed"] = e
        p[f"{q}.cluster_size"] = torch.zeros(n_embed)
        p[f"{q}.embed_avg"] = e.clone()
    return p


VGG_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512]
# torchvision vgg16.features indices of the 13 convs and which reference slice they fall in
# (models/lpips.py:127-136): slice1 = 0..3, slice2 = 4..8, slice3 = 9..15, slice4 = 16..22, slice5 = 23..29
VGG_CONV_IDX = [0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28]
VGG_SLICE_ENDS = [4, 9, 16, 23, 30]
LPIPS_CHNS = [64, 128, 256, 512, 512]


def _vgg_key(idx: int) -> str:
    s = next(i for i, e in enumerate(VGG_SLICE_ENDS) if idx < e) + 1
    return f"net.slice{s}.{idx}"


def init_lpips_params(seed: int = 1) -> Dict[str, Tensor]:
    """Seeded random LPIPS weights with the reference's key names (models/lpips.py:50-64,115-136).
    The pretrained vgg.pth cannot be downloaded offline; lin weights are made
    non-negative like the real ones."""
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, Tensor] = {}
    cin = 3
    ci = 0
    for v in VGG_CFG:
        if v == "M":
            continue
        # He-normal keeps activations O(1) through 13 layers so every tap is exercised
        w = torch.randn(v, cin, 3, 3, generator=g) * math.sqrt(2.0 / (cin * 9))
        b = torch.randn(v, generator=g) * 0.05
        k = _vgg_key(VGG_CONV_IDX[ci])
        p[k + ".weight"], p[k + ".bias"] = w, b
        cin = v
        ci += 1
    for i, c in enumerate(LPIPS_CHNS):
        p[f"lin{i}.model.1.weight"] = torch.randn(1, c, 1, 1, generator=g).abs() * 0.1
    p["scaling_layer.shift"] = torch.tensor([-.030, -.088, -.188])[None, :, None, None]
    p["scaling_layer.scale"] = torch.tensor([.458, .448, .450])[None, :, None, None]
    return p


# ----------------------------------------------------------------------------------------------
# Quantize  (models/vqvae_conv3d_latent.py:33-83)
# ----------------------------------------------------------------------------------------------


def quantize_assign(flatten: Tensor, embed: Tensor) -> Tuple[Tensor, Tensor]:
    """:48-54 -- dist = |x|^2 - 2 x@E + |e|^2 ; embed_ind = argmax(-dist) (first max wins)."""
    dist = (flatten.pow(2).sum(1, keepdim=True) - 2 * flatten @ embed
            + embed.pow(2).sum(0, keepdim=True))
    _, embed_ind = (-dist).max(1)
    return embed_ind, dist


def quantize_stats(flatten: Tensor, embed_ind: Tensor, n_embed: int) -> Tuple[Tensor, Tensor]:
    """:55,60-61 -- one-hot counts and per-code sums (the inputs of the cross-rank SUM all-reduce :63-64)."""
    onehot = F.one_hot(embed_ind.reshape(-1), n_embed).type(flatten.dtype)
    return onehot.sum(0), flatten.transpose(0, 1) @ onehot


def quantize_ema(embed: Tensor, cluster_size: Tensor, embed_avg: Tensor, onehot_sum: Tensor,
                 embed_sum: Tensor, decay: float = 0.99, eps: float = 1e-5):
    """:66-75 -- EMA update + Laplace-smoothed renormalisation.  Returns new (embed, cluster_size, embed_avg)."""
    n_embed = embed.shape[1]
    cs = cluster_size * decay + onehot_sum * (1 - decay)
    ea = embed_avg * decay + embed_sum * (1 - decay)
    n = cs.sum()
    cs_n = (cs + eps) / (n + n_embed * eps) * n
    return ea / cs_n.unsqueeze(0), cs, ea


def quantize_forward(inp: Tensor, embed: Tensor, cluster_size: Tensor, embed_avg: Tensor,
                     training: bool = True, decay: float = 0.99, eps: float = 1e-5):
    """Quantize.forward :47-80.  Returns (quantize, diff, embed_ind, new_buffers|None, (counts, sums)|None).
    ``quantize`` carries the straight-through graph ``x + (q - x).detach()`` (:78)."""
    dim, n_embed = embed.shape
    flatten = inp.reshape(-1, dim)
    embed_ind, _ = quantize_assign(flatten.detach(), embed)
    ind = embed_ind.view(*inp.shape[:-1])
    q = F.embedding(ind, embed.transpose(0, 1))  # :57,82-83 -- pre-update codebook
    new_buffers = stats = None
    if training:
        stats = quantize_stats(flatten.detach(), embed_ind, n_embed)
        new_buffers = quantize_ema(embed, cluster_size, embed_avg, stats[0], stats[1], decay, eps)
    diff = (q.detach() - inp).pow(2).mean()  # :77
    quantize = inp + (q - inp).detach()  # :78
    return quantize, diff, ind, new_buffers, stats


# ----------------------------------------------------------------------------------------------
# Conv stacks  (models/vqvae_conv3d_latent.py:86-190)
# ----------------------------------------------------------------------------------------------


# Optional probe (tests only): when RELU_PROBE is a list, every ReLU appends (min |x|, max |x|, numel) of its input.  A test
# that compares GRADIENTS tightly uses it to know how close the nearest pre-activation is to the (discontinuous) gate.
RELU_PROBE = None
# Optional hooks (tests only): callables ``hook(x) -> tensor | None`` that may replace relu(x) / max_pool2d(x, 2, 2) by
# the SAME piecewise-linear branch another run took (gates / pooling winners given), see tests/gate_consistent.py.  ReLU and
# max-pool are discontinuous in their gradients: a pre-activation within the forward rounding error of zero (or a pooling
# window whose two largest values are that close) legitimately sends the gradient elsewhere, which no finite-precision
# implementation -- the reference's own fp32 included -- can reproduce bit for bit.
RELU_HOOK = None
MAXPOOL_HOOK = None


def _relu(x):
    if RELU_PROBE is not None:
        a = x.detach().abs()
        RELU_PROBE.append((a.min().item(), a.max().item(), a.numel()))
    if RELU_HOOK is not None:
        y = RELU_HOOK(x)
        if y is not None:
            return y
    return F.relu(x)


def _maxpool2(x):
    if MAXPOOL_HOOK is not None:
        y = MAXPOOL_HOOK(x)
        if y is not None:
            return y
    return F.max_pool2d(x, 2, 2)


def relu_margin(probe):
    """Smallest |pre-activation| / max |pre-activation| over the probed ReLU inputs."""
    return min(mn / mx for mn, mx, _ in probe) if probe else float("inf")


def _c2(p, name, x, stride=1, padding=0):
    return F.conv2d(x, p[name + ".weight"], p[name + ".bias"], stride=stride, padding=padding)


def _ct2(p, name, x):
    return F.conv_transpose2d(x, p[name + ".weight"], p[name + ".bias"], stride=2, padding=1)


def resblock(p, prefix, x):
    """ResBlock :86-101 -- ReLU, conv3x3 C->r, ReLU, conv1x1 r->C, += input."""
    h = _relu(x)
    h = _c2(p, prefix + ".conv.1", h, padding=1)
    h = _relu(h)
    h = _c2(p, prefix + ".conv.3", h)
    return h + x


def encoder(p, prefix, x, stride, n_res_block=2):
    """Encoder :103-131."""
    if stride == 4:
        x = _relu(_c2(p, prefix + ".blocks.0", x, 2, 1))
        x = _relu(_c2(p, prefix + ".blocks.2", x, 2, 1))
        x = _c2(p, prefix + ".blocks.4", x, 1, 1)
        start = 5
    else:
        x = _relu(_c2(p, prefix + ".blocks.0", x, 2, 1))
        x = _c2(p, prefix + ".blocks.2", x, 1, 1)
        start = 3
    for i in range(n_res_block):
        x = resblock(p, f"{prefix}.blocks.{start + i}", x)
    return _relu(x)


def decoder(p, prefix, x, stride, n_res_block=2):
    """Decoder :134-166."""
    x = _c2(p, prefix + ".blocks.0", x, 1, 1)
    for i in range(n_res_block):
        x = resblock(p, f"{prefix}.blocks.{1 + i}", x)
    x = _relu(x)
    x = _ct2(p, f"{prefix}.blocks.{2 + n_res_block}", x)
    if stride == 4:
        x = _ct2(p, f"{prefix}.blocks.{4 + n_res_block}", _relu(x))
    return x


def conv3d_postnet(p, prefix, x):
    """Conv3dLatentPostnet :169-190 -- (Conv3d k3 p1 + ReLU) x2, Conv3d k3 p1.  x: [B,C,T,H,W]."""
    for i in range(3):
        x = F.conv3d(x, p[f"{prefix}.conv3d.{i}.0.weight"], p[f"{prefix}.conv3d.{i}.0.bias"], padding=1)
        if i < 2:
            x = _relu(x)
    return x


def vqvae_forward(p: Dict[str, Tensor], inp: Tensor, n_clips: int = 1, training: bool = True,
                  decay: float = 0.99, eps: float = 1e-5):
    """VQVAE.forward :243-285 for ``n_clips`` clips stacked along dim 0 (inp [B*T, Cin, H, W]).

    Returns dict(dec, diff[1], id_t, id_b, new_buffers{...}, stats{...}).  Does NOT mutate ``p``.
    """
    F_, _, _, _ = inp.shape
    T = F_ // n_clips
    enc_b = encoder(p, "enc_b", inp, 4)  # :237-241
    enc_t = encoder(p, "enc_t", enc_b, 2)

    def to5(x):  # :247 per clip
        return x.reshape(n_clips, T, *x.shape[1:]).permute(0, 2, 1, 3, 4)

    def to4(x):  # :251
        return x.permute(0, 2, 1, 3, 4).reshape(F_, *x.shape[1:2], *x.shape[3:])

    enc_b_c = to4(conv3d_postnet(p, "conv3d_encoded_b", to5(enc_b)))
    enc_t_c = to4(conv3d_postnet(p, "conv3d_encoded_t", to5(enc_t)))

    out = {"new_buffers": {}, "stats": {}}
    # encode_quantized :261-278
    qt_in = _c2(p, "quantize_conv_t", enc_t_c).permute(0, 2, 3, 1)
    quant_t, diff_t, id_t, nb, st = quantize_forward(
        qt_in, p["quantize_t.embed"], p["quantize_t.cluster_size"], p["quantize_t.embed_avg"],
        training, decay, eps)
    out["new_buffers"]["quantize_t"], out["stats"]["quantize_t"] = nb, st
    quant_t = quant_t.permute(0, 3, 1, 2)
    dec_t = decoder(p, "dec_t", quant_t, 2)
    cat_b = torch.cat([dec_t, enc_b_c], 1)
    qb_in = _c2(p, "quantize_conv_b", cat_b).permute(0, 2, 3, 1)
    quant_b, diff_b, id_b, nb, st = quantize_forward(
        qb_in, p["quantize_b.embed"], p["quantize_b.cluster_size"], p["quantize_b.embed_avg"],
        training, decay, eps)
    out["new_buffers"]["quantize_b"], out["stats"]["quantize_b"] = nb, st
    quant_b = quant_b.permute(0, 3, 1, 2)
    # decode :280-285
    up_t = _ct2(p, "upsample_t", quant_t)
    dec = decoder(p, "dec", torch.cat([up_t, quant_b], 1), 4)
    out.update(dec=dec, diff=diff_t.unsqueeze(0) + diff_b.unsqueeze(0), id_t=id_t, id_b=id_b,
               qt_in=qt_in, qb_in=qb_in)
    return out


# ----------------------------------------------------------------------------------------------
# LPIPS  (models/lpips.py:50-161, loss.py:27-33)
# ----------------------------------------------------------------------------------------------


def vgg_taps(p: Dict[str, Tensor], x: Tensor):
    """vgg16 trunk :115-152 -- returns the 5 ReLU taps."""
    taps = []
    ci = 0
    idx = 0
    for v in VGG_CFG:
        if v == "M":
            x = _maxpool2(x)
            idx += 1
        else:
            k = _vgg_key(VGG_CONV_IDX[ci])
            x = _relu(F.conv2d(x, p[k + ".weight"], p[k + ".bias"], padding=1))
            ci += 1
            idx += 2
        if idx in VGG_SLICE_ENDS:
            taps.append(x)
    return taps


def normalize_tensor(x, eps=1e-10):
    """:155-157 -- eps added to the norm, outside the sqrt."""
    return x / (torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True)) + eps)


def lpips_forward(p: Dict[str, Tensor], inp: Tensor, target: Tensor) -> Tensor:
    """LPIPS.forward :80-93 (dropout inactive: .eval() at loss.py:30).  Returns [N,1,1,1]."""
    s0 = (inp - p["scaling_layer.shift"]) / p["scaling_layer.scale"]
    s1 = (target - p["scaling_layer.shift"]) / p["scaling_layer.scale"]
    t0, t1 = vgg_taps(p, s0), vgg_taps(p, s1)
    val = None
    for k in range(5):
        d = (normalize_tensor(t0[k]) - normalize_tensor(t1[k])) ** 2
        r = F.conv2d(d, p[f"lin{k}.model.1.weight"]).mean([2, 3], keepdim=True)
        val = r if val is None else val + r
    return val


def vqlpips(p: Dict[str, Tensor], targets: Tensor, reconstructions: Tensor) -> Tensor:
    """VQLPIPS.forward loss.py:32-33."""
    return lpips_forward(p, targets.contiguous(), reconstructions.contiguous()).mean()


# ----------------------------------------------------------------------------------------------
# Training step  (train_faceoff_perceptual.py:32-47,98-100 ; train_faceoff.py:31-44,140-142)
# ----------------------------------------------------------------------------------------------

PARAM_SUFFIXES = (".weight", ".bias")


def trainable_keys(p: Dict[str, Tensor]):
    return [k for k in p if k.endswith(PARAM_SUFFIXES)]


def train_step(p: Dict[str, Tensor], img: Tensor, gt: Tensor, n_clips: int = 1,
               lp: Optional[Dict[str, Tensor]] = None, latent_w: float = 1.0, perc_w: float = 1.0,
               dtype=torch.float32):
    """zero_grad -> forward -> MSE(out[:, :3], gt) + latent.mean() [+ VQLPIPS(gt, out)] -> backward.

    Returns dict with losses, grads (by state_dict key), new codebook buffers, indices, dec.
    """
    q = {k: v.detach().to(dtype).clone() for k, v in p.items()}
    for k in trainable_keys(q):
        q[k].requires_grad_(True)
    img = img.to(dtype)
    gt = gt.to(dtype)
    out = vqvae_forward(q, img, n_clips=n_clips, training=True)
    rec = out["dec"][:, :3]
    recon_loss = F.mse_loss(rec, gt)
    latent_loss = out["diff"].mean()
    loss = recon_loss + latent_w * latent_loss
    perc = None
    if lp is not None:
        lq = {k: v.to(dtype) for k, v in lp.items()}
        perc = vqlpips(lq, gt, rec)
        loss = loss + perc_w * perc
    loss.backward()
    grads = {k: q[k].grad.detach() for k in trainable_keys(q)}
    return dict(loss=loss.detach(), recon_loss=recon_loss.detach(), latent_loss=latent_loss.detach(),
                perceptual_loss=None if perc is None else perc.detach(), grads=grads,
                new_buffers=out["new_buffers"], stats=out["stats"], id_t=out["id_t"], id_b=out["id_b"],
                dec=out["dec"].detach(), qt_in=out["qt_in"].detach(), qb_in=out["qb_in"].detach())


def synthetic_clip(n_clips: int, T: int, H: int, W: int, seed: int = 1234, in_channel: int = 6):
    """Synthetic stand-in for process_data (utils.py:29-38): img [B*T,6,H,W], gt [B*T,3,H,W] in [-1,1]."""
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(n_clips * T, in_channel, H, W, generator=g) * 2 - 1
    gt = torch.rand(n_clips * T, 3, H, W, generator=g) * 2 - 1
    return img, gt
