"""Drop-in replacements for the reference's models/vqvae_conv3d_latent.py modules.

Same class names, constructor arguments (positional order), forward signatures and state-dict keys as the
reference; all arithmetic runs in the sm_100a kernels behind the C ABI
(include/faceoff_b200.h).  There is no PyTorch fallback: without a B200 + libfaceoff_b200.so every forward raises.

Precision contract: ``Quantize`` is an fp32 path (indices bit-exact outside near-ties); the conv stacks are
bf16 tensor-core kernels with fp32 accumulation (north_star (3)), activations stored bf16 channels-last.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import nn

from . import distributed as dist_fn
from . import ops
from ._lib import FaceoffB200Error
from .graph import (FORM_DOWN, FORM_S1, FORM_UP, Node, Tape, View, conv_op, first_conv_op, last_convT_op,
                    resblock_op)


# ------------------------------------------------------------------------------------------------
# Quantize  (reference :33-83)
# ------------------------------------------------------------------------------------------------
class _StatSink:
    """Where a quantiser's EMA statistics go.  Default: the reference behaviour -- all_reduce(SUM) both
    tensors right away (:63-64) then EMA (:66-75).  (The fused data-parallel reducer, parallel.py, bypasses the sink
    for the forwards it owns and folds the statistics into its single per-step bucket.)"""

    def submit(self, q: "Quantize", counts: torch.Tensor, embed_sum: torch.Tensor):
        dist_fn.all_reduce(counts)
        dist_fn.all_reduce(embed_sum)
        q.apply_ema(counts, embed_sum)


class _LocalStatSink(_StatSink):
    """No collective: EMA from this process's statistics only (single-process runs inside a multi-rank job)."""

    def submit(self, q, counts, embed_sum):
        q.apply_ema(counts, embed_sum)


_default_sink = _StatSink()


def _quantize_forward(q: "Quantize", x32: torch.Tensor, want_bf16: bool, dp=None):
    """x32: fp32 [rows, dim] contiguous.  Returns (q_f32, q_bf16|None, diff_sum[1], ind[rows], e_t).
    ``dp``: the FusedDataParallel whose ``begin_forward`` ran for this forward (statistics go into its bucket and the
    EMA update is deferred to the end of backward), or None: reference behaviour, all_reduce x2 + EMA right here."""
    rows, dim = x32.shape
    e_split, e_t, e_n2 = ops.vq_prep(q.embed)
    ind = ops.vq_assign(x32, e_t, e_split, e_n2, q.n_flagged)
    diff_sum = torch.zeros(1, dtype=torch.float32, device=x32.device)
    counts = embed_sum = None
    if q.training:
        if dp is not None:
            counts, embed_sum = dp.stat_buffers(q)
        else:
            counts = torch.zeros(q.n_embed, dtype=torch.float32, device=x32.device)
            embed_sum = torch.zeros(dim, q.n_embed, dtype=torch.float32, device=x32.device)
    q32, q16 = ops.vq_gather_stats(x32, ind, e_t, diff_sum, counts, embed_sum, want_f32=True, want_bf16=want_bf16)
    if q.training:
        if dp is not None:
            dp.submit_stats(q)
        else:
            q.stat_sink.submit(q, counts, embed_sum)
    return q32, q16, diff_sum, ind, e_t


class _QuantizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp: torch.Tensor, q: "Quantize"):
        x32 = inp.detach().reshape(-1, q.dim).to(torch.float32).contiguous()
        q32, _, diff_sum, ind, e_t = _quantize_forward(q, x32, want_bf16=False)
        ctx.save_for_backward(x32, ind, e_t)
        ctx.in_shape = inp.shape
        diff = (diff_sum / float(x32.numel())).reshape(())
        embed_ind = ind.view(*inp.shape[:-1])
        ctx.mark_non_differentiable(embed_ind)
        return q32.view(inp.shape), diff, embed_ind

    @staticmethod
    def backward(ctx, g_q, g_diff, _g_ind):
        x32, ind, e_t = ctx.saved_tensors
        gq = None if g_q is None else g_q.reshape(-1, x32.shape[1]).to(torch.float32).contiguous()
        gd = None if g_diff is None else g_diff.reshape(1).to(torch.float32).contiguous()
        g32, _ = ops.vq_backward(gq, 0, gd, x32, ind, e_t)
        return g32.view(ctx.in_shape), None


class Quantize(nn.Module):
    """Vector quantiser with EMA codebook (reference :33-83).  forward(input[..., dim]) ->
    (quantize, diff, embed_ind)."""

    def __init__(self, dim, n_embed, decay=0.99, eps=1e-5):
        super().__init__()
        self.dim = dim
        self.n_embed = n_embed
        self.decay = decay
        self.eps = eps
        embed = torch.randn(dim, n_embed)
        self.register_buffer("embed", embed)
        self.register_buffer("cluster_size", torch.zeros(n_embed))
        self.register_buffer("embed_avg", embed.clone())
        self.stat_sink: _StatSink = _default_sink
        self.n_flagged: Optional[torch.Tensor] = None  # optional device int32[1]: rows re-evaluated exactly

    def forward(self, input):
        return _QuantizeFn.apply(input, self)

    def embed_code(self, embed_id):
        # :82-83 -- a plain gather from the (current) codebook; tiny, index plumbing
        return torch.nn.functional.embedding(embed_id, self.embed.transpose(0, 1))

    def apply_ema(self, counts: torch.Tensor, embed_sum: torch.Tensor):
        """EMA + renormalisation (:66-75) with (already all-reduced) statistics, in place on the buffers."""
        ops.vq_ema(self.embed, self.cluster_size, self.embed_avg, counts, embed_sum, self.decay, self.eps)


# ------------------------------------------------------------------------------------------------
# parameter-holder modules: identical module tree => identical state_dict keys
# ------------------------------------------------------------------------------------------------
def _params_of(module: nn.Module, prefix: str = "") -> Dict[str, torch.Tensor]:
    return {prefix + k: v for k, v in module.named_parameters()}


class _GraphFn(torch.autograd.Function):
    """One autograd node for a whole sub-graph executed through the tape."""

    @staticmethod
    def forward(ctx, runner, inp, names, grad_mode, *params):
        # needs_input_grad ignores torch.no_grad(), and grad mode is always off inside forward: the caller passes it in
        need_grad = grad_mode and any(ctx.needs_input_grad)
        tape = Tape(dict(zip(names, params)), need_grad)
        outs, state = runner(tape, inp.detach())
        ctx.tape, ctx.state, ctx.names = tape, state, names
        ctx.inp_needs_grad = inp.requires_grad
        ctx.n_out = len(outs)
        nd = [outs[i] for i in state.get("non_diff", ())]
        if nd:
            ctx.mark_non_differentiable(*nd)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        tape, state = ctx.tape, ctx.state
        if tape is None:
            raise FaceoffB200Error("backward through a faceoff_b200 graph a second time: the fused tape frees its saved "
                                   "activations during the first backward (retain_graph=True is not supported); sum the "
                                   "losses and call backward once, or run forward again")
        dp = state.get("dp")
        if dp is not None:
            dp.begin_step(tape)
        old, ops.PRECISE = ops.PRECISE, tape.precise
        try:
            state["seed"](tape, gouts)
            tape.backward()
            gi = state["input_grad"]() if ctx.inp_needs_grad else None
        finally:
            ops.PRECISE = old
        if dp is not None:
            dp.end_step()
            grads = dp.grads_for_autograd(tape, ctx.names)
        else:
            grads = tuple(tape.grads.get(n) for n in ctx.names)
        ctx.tape = ctx.state = None
        return (None, gi, None, None) + grads


def apply_graph(runner, inp: torch.Tensor, names, params):
    """Run ``runner`` as ONE autograd node; the tape records a backward only when autograd is enabled right now."""
    return _GraphFn.apply(runner, inp, names, torch.is_grad_enabled(), *params)


def _run_graph(module: nn.Module, runner, inp: torch.Tensor):
    ps = _params_of(module)
    names = tuple(ps.keys())
    return apply_graph(runner, inp, names, ps.values())


def _io_wrap(build, c_in: int, c_out_fn):
    """Runner for a stand-alone stack: fp32 NCHW in -> graph -> fp32 NCHW out."""

    def runner(tape: Tape, x: torch.Tensor):
        assert x.dim() == 4, "expected [N, C, H, W]"
        xin = Node(c_in, raw=ops.pack_nchw(x.to(torch.float32)))
        xin.act = None
        out_node, out_relu = build(tape, xin)
        t = out_node.act if out_relu else out_node.raw
        y = ops.unpack_nchw(t, out_node.c)

        def seed(tape_, gouts):
            g = ops.pack_nchw(gouts[0].to(torch.float32), cs=t.shape[-1])
            if out_relu:
                # gradient arrives w.r.t. relu(raw): gate it (rare stand-alone path; index plumbing in torch)
                if tape_.precise:
                    m = t[..., :t.shape[-1] // 2] > 0
                    g = g * torch.cat([m, m], -1)
                else:
                    g = g * (t > 0)
            out_node.g = (g, 0)
            return g

        def input_grad():
            return ops.unpack_nchw(xin.g[0], c_in) if xin.g is not None else None

        return (y,), {"seed": seed, "input_grad": input_grad}

    return runner


class ResBlock(nn.Module):
    def __init__(self, in_channel, channel):
        super().__init__()
        self.conv = nn.Sequential(
            nn.ReLU(),
            nn.Conv2d(in_channel, channel, 3, padding=1),
            nn.ReLU(inplace=True),
            nn.Conv2d(channel, in_channel, 1),
        )
        self.in_channel, self.channel = in_channel, channel

    def forward(self, input):
        def build(tape, x):
            x.act = ops.relu(x.raw)
            h = conv_op(tape, FORM_S1, 3, [View(x, True)], "conv.1", self.channel, want_raw=False, want_relu=True)
            o = conv_op(tape, FORM_S1, 1, [View(h, True)], "conv.3", self.in_channel, residual=x)
            return o, False

        return _run_graph(self, _io_wrap(build, self.in_channel, None), input)[0]


def _encoder_graph(tape: Tape, x, prefix: str, channel: int, n_res_block: int, n_res_channel: int, stride: int,
                   input_needs_grad: bool) -> Node:
    """Encoder (reference :103-131).  Returns a node whose ``act`` is the encoder output (post final ReLU).
    ``x``: a View (which of raw / relu the encoder reads is the CALLER's statement, never guessed), or (stride 4 only)
    the raw fp32 NCHW image with <= 8 channels (im2col first layer, no input grad)."""
    if stride == 4:
        if isinstance(x, torch.Tensor):
            a = first_conv_op(tape, x, prefix + "blocks.0", x.shape[1], channel // 2)
        else:
            a = conv_op(tape, FORM_DOWN, 4, [x], prefix + "blocks.0", channel // 2,
                        want_raw=False, want_relu=True, input_needs_grad=input_needs_grad)
        a = conv_op(tape, FORM_DOWN, 4, [View(a, True)], prefix + "blocks.2", channel, want_raw=False, want_relu=True)
        a = conv_op(tape, FORM_S1, 3, [View(a, True)], prefix + "blocks.4", channel, want_raw=True,
                    want_relu=True)
        start = 5
    else:
        a = conv_op(tape, FORM_DOWN, 4, [x], prefix + "blocks.0", channel // 2, want_raw=False,
                    want_relu=True, input_needs_grad=input_needs_grad)
        a = conv_op(tape, FORM_S1, 3, [View(a, True)], prefix + "blocks.2", channel, want_raw=True, want_relu=True)
        start = 3
    if n_res_block == 0:
        return a
    for i in range(n_res_block):
        a = resblock_op(tape, a, f"{prefix}blocks.{start + i}", channel, n_res_channel, last=(i == n_res_block - 1))
    return a


def _decoder_graph(tape: Tape, srcs, prefix: str, out_channel: int, channel: int, n_res_block: int,
                   n_res_channel: int, stride: int, final_f32: Optional[str] = None) -> Node:
    """Decoder (reference :134-166).  ``srcs``: list of raw views (torch.cat folded into the first conv)."""
    a = conv_op(tape, FORM_S1, 3, srcs, prefix + "blocks.0", channel, want_raw=True, want_relu=True)
    for i in range(n_res_block):
        a = resblock_op(tape, a, f"{prefix}blocks.{1 + i}", channel, n_res_channel, last=(i == n_res_block - 1))
    k = 2 + n_res_block
    if stride == 4:
        a = conv_op(tape, FORM_UP, 4, [View(a, True)], f"{prefix}blocks.{k}", channel // 2, transposed=True,
                    want_raw=False, want_relu=True)
        if final_f32 == "nchw" and out_channel <= 8 and not tape.precise:
            return last_convT_op(tape, View(a, True), f"{prefix}blocks.{k + 2}", out_channel)
        return conv_op(tape, FORM_UP, 4, [View(a, True)], f"{prefix}blocks.{k + 2}", out_channel, transposed=True,
                       want_raw=final_f32 is None, f32=final_f32)
    return conv_op(tape, FORM_UP, 4, [View(a, True)], f"{prefix}blocks.{k}", out_channel, transposed=True,
                   want_raw=final_f32 is None, f32=final_f32)


def _conv3d_graph(tape: Tape, x: View, prefix: str, channels: int, clips: int) -> Node:
    """Conv3dLatentPostnet (reference :169-190)."""
    a = conv_op(tape, FORM_S1, 3, [x], prefix + "conv3d.0.0", channels, want_raw=False, want_relu=True, ndim=3,
                clips=clips)
    a = conv_op(tape, FORM_S1, 3, [View(a, True)], prefix + "conv3d.1.0", channels, want_raw=False, want_relu=True,
                ndim=3, clips=clips)
    return conv_op(tape, FORM_S1, 3, [View(a, True)], prefix + "conv3d.2.0", channels, ndim=3, clips=clips)


class Encoder(nn.Module):
    def __init__(self, in_channel, channel, n_res_block, n_res_channel, stride):
        super().__init__()
        if stride == 4:
            blocks = [
                nn.Conv2d(in_channel, channel // 2, 4, stride=2, padding=1),
                nn.ReLU(inplace=True),
                nn.Conv2d(channel // 2, channel, 4, stride=2, padding=1),
                nn.ReLU(inplace=True),
                nn.Conv2d(channel, channel, 3, padding=1),
            ]
        elif stride == 2:
            blocks = [
                nn.Conv2d(in_channel, channel // 2, 4, stride=2, padding=1),
                nn.ReLU(inplace=True),
                nn.Conv2d(channel // 2, channel, 3, padding=1),
            ]
        else:
            raise ValueError("stride must be 2 or 4")
        for _ in range(n_res_block):
            blocks.append(ResBlock(channel, n_res_channel))
        blocks.append(nn.ReLU(inplace=True))
        self.blocks = nn.Sequential(*blocks)
        self.cfg = (in_channel, channel, n_res_block, n_res_channel, stride)

    def forward(self, input):
        in_channel, channel, n_res_block, n_res_channel, stride = self.cfg

        def build(tape, x):
            return _encoder_graph(tape, View(x, False), "", channel, n_res_block, n_res_channel, stride,
                                  input_needs_grad=input.requires_grad), True

        return _run_graph(self, _io_wrap(build, in_channel, None), input)[0]


class Decoder(nn.Module):
    def __init__(self, in_channel, out_channel, channel, n_res_block, n_res_channel, stride):
        super().__init__()
        blocks = [nn.Conv2d(in_channel, channel, 3, padding=1)]
        for _ in range(n_res_block):
            blocks.append(ResBlock(channel, n_res_channel))
        blocks.append(nn.ReLU(inplace=True))
        if stride == 4:
            blocks.extend([
                nn.ConvTranspose2d(channel, channel // 2, 4, stride=2, padding=1),
                nn.ReLU(inplace=True),
                nn.ConvTranspose2d(channel // 2, out_channel, 4, stride=2, padding=1),
            ])
        elif stride == 2:
            blocks.append(nn.ConvTranspose2d(channel, out_channel, 4, stride=2, padding=1))
        else:
            raise ValueError("stride must be 2 or 4")
        self.blocks = nn.Sequential(*blocks)
        self.cfg = (in_channel, out_channel, channel, n_res_block, n_res_channel, stride)

    def forward(self, input):
        in_channel, out_channel, channel, n_res_block, n_res_channel, stride = self.cfg

        def build(tape, x):
            return _decoder_graph(tape, [View(x, False)], "", out_channel, channel, n_res_block, n_res_channel,
                                  stride), False

        return _run_graph(self, _io_wrap(build, in_channel, None), input)[0]


class Conv3dLatentPostnet(nn.Module):
    """applies a sequence of conv3d layers with relu activation (reference :169-190).  forward(x[N,C,D,H,W])."""

    def __init__(self, channels):
        super().__init__()
        self.conv3d = nn.Sequential(
            self.conv3d_layer(channels=channels),
            self.conv3d_layer(channels=channels),
            self.conv3d_layer(channels=channels, is_final=True),
        )
        self.channels = channels

    def conv3d_layer(self, channels=128, kernel_size=3, padding=1, is_final=False):
        if is_final:
            return nn.Sequential(nn.Conv3d(channels, channels, kernel_size, padding=padding))
        return nn.Sequential(nn.Conv3d(channels, channels, kernel_size, padding=padding), nn.ReLU())

    def forward(self, input):
        assert input.dim() == 5, "expected [N, C, D, H, W]"
        n, c, d, h, w = input.shape
        clips = n

        def runner(tape: Tape, x: torch.Tensor):
            # [N,C,D,H,W] -> frames-major [N*D, C, H, W] (layout plumbing) -> channels-last bf16
            x4 = x.to(torch.float32).permute(0, 2, 1, 3, 4).reshape(n * d, c, h, w)
            xin = Node(c, raw=ops.pack_nchw(x4))
            out = _conv3d_graph(tape, View(xin, False), "", self.channels, clips)
            y = ops.unpack_nchw(out.raw, c).view(n, d, c, h, w).permute(0, 2, 1, 3, 4)

            def seed(tape_, gouts):
                g4 = gouts[0].to(torch.float32).permute(0, 2, 1, 3, 4).reshape(n * d, c, h, w)
                out.g = (ops.pack_nchw(g4, cs=out.raw.shape[-1]), 0)

            def input_grad():
                if xin.g is None:
                    return None
                return ops.unpack_nchw(xin.g[0], c).view(n, d, c, h, w).permute(0, 2, 1, 3, 4)

            return (y,), {"seed": seed, "input_grad": input_grad}

        return _run_graph(self, runner, input)[0]


# ------------------------------------------------------------------------------------------------
# VQVAE  (reference :192-295)
# ------------------------------------------------------------------------------------------------
class VQVAE(nn.Module):
    def __init__(self, in_channel=3, channel=128, n_res_block=2, n_res_channel=32, embed_dim=64, n_embed=512,
                 decay=0.99, residual=False):
        super().__init__()
        self.enc_b = Encoder(in_channel, channel, n_res_block, n_res_channel, stride=4)
        self.enc_t = Encoder(channel, channel, n_res_block, n_res_channel, stride=2)
        self.quantize_conv_t = nn.Conv2d(channel, embed_dim, 1)
        self.quantize_t = Quantize(embed_dim, n_embed)  # NB: decay not forwarded, like the reference (:209)
        self.dec_t = Decoder(embed_dim, embed_dim, channel, n_res_block, n_res_channel, stride=2)
        self.quantize_conv_b = nn.Conv2d(embed_dim + channel, embed_dim, 1)
        self.quantize_b = Quantize(embed_dim, n_embed)
        self.upsample_t = nn.ConvTranspose2d(embed_dim, embed_dim, 4, stride=2, padding=1)
        self.dec = Decoder(embed_dim + embed_dim, in_channel, channel, n_res_block, n_res_channel, stride=4)
        self.residual = residual
        # hard-wired to 128 channels like the reference (:230-231)
        self.conv3d_encoded_b = Conv3dLatentPostnet(128)
        self.conv3d_encoded_t = Conv3dLatentPostnet(128)
        self.cfg = dict(in_channel=in_channel, channel=channel, n_res_block=n_res_block, n_res_channel=n_res_channel,
                        embed_dim=embed_dim, n_embed=n_embed)
        if channel != 128:
            raise ValueError("the reference hard-wires the Conv3d stacks to 128 channels; channel must be 128")

    # -- the fused training graph ----------------------------------------------------------------
    def _runner(self, clips: int, want_ids: bool, want_pre: bool = False):
        cfg = self.cfg
        ch, nrb, nrc, ed = cfg["channel"], cfg["n_res_block"], cfg["n_res_channel"], cfg["embed_dim"]
        cin = cfg["in_channel"]
        model = self

        def runner(tape: Tape, x: torch.Tensor):
            F_ = x.shape[0]
            assert F_ % clips == 0
            dp = getattr(model, "_dp", None)
            if dp is not None and model.training and tape.need_grad:
                dp.begin_forward(tape.params, model.param_forward_order())
            else:
                dp = None
            x = x.to(torch.float32).contiguous()
            small = cin <= 8 and not tape.precise   # im2col / col2im fast path for the 6-channel image-side layers
            xin = x if small else View(Node(cin, raw=ops.pack_nchw(x)), False)
            enc_b = _encoder_graph(tape, xin, "enc_b.", ch, nrb, nrc, 4, input_needs_grad=False)
            enc_t = _encoder_graph(tape, View(enc_b, True), "enc_t.", ch, nrb, nrc, 2, input_needs_grad=True)
            eb_c = _conv3d_graph(tape, View(enc_b, True), "conv3d_encoded_b.", 128, clips)
            et_c = _conv3d_graph(tape, View(enc_t, True), "conv3d_encoded_t.", 128, clips)

            gdiff_holder = {}

            def quantize(qmod: Quantize, pre: Node):
                x32 = pre.f32.view(-1, ed)
                q32, q16, diff_sum, ind, e_t = _quantize_forward(qmod, x32, want_bf16=True, dp=dp)
                node = Node(ed, raw=q16.view(*pre.f32.shape[:-1], -1))
                rec = dict(x32=x32, ind=ind, diff_sum=diff_sum)

                def vq_bwd():  # recorded right after the producing conv => replayed right before its backward
                    gq, g_off = node.g if node.g is not None else (None, 0)
                    _, g16 = ops.vq_backward(None if gq is None else gq.view(-1, gq.shape[-1]), g_off,
                                             gdiff_holder.get("g"), x32, ind, e_t, want_f32=False, want_bf16=True)
                    pre.g = (g16.view(*pre.f32.shape[:-1], -1), 0)

                tape.record(vq_bwd)
                return node, rec

            pre_t = conv_op(tape, FORM_S1, 1, [View(et_c, False)], "quantize_conv_t", ed, want_raw=False, f32="cl")
            qt, rec_t = quantize(model.quantize_t, pre_t)
            dec_t = _decoder_graph(tape, [View(qt, False)], "dec_t.", ed, ch, nrb, nrc, 2)
            pre_b = conv_op(tape, FORM_S1, 1, [View(dec_t, False), View(eb_c, False)], "quantize_conv_b", ed,
                            want_raw=False, f32="cl")
            qb, rec_b = quantize(model.quantize_b, pre_b)
            up_t = conv_op(tape, FORM_UP, 4, [View(qt, False)], "upsample_t", ed, transposed=True)
            out = _decoder_graph(tape, [View(up_t, False), View(qb, False)], "dec.", cin, ch, nrb, nrc, 4,
                                 final_f32="nchw")
            dec = out.f32
            numel_t, numel_b = float(rec_t["x32"].numel()), float(rec_b["x32"].numel())
            diff = rec_t["diff_sum"] / numel_t + rec_b["diff_sum"] / numel_b  # [1]  (:268,276,278)
            def seed(tape_, gouts):
                g_dec, g_diff = gouts[0], gouts[1]
                if g_dec is None:
                    g_dec = torch.zeros_like(dec)
                g_dec = g_dec.to(torch.float32).contiguous()
                if small:
                    out.gdec = g_dec
                else:
                    out.g = (ops.pack_nchw(g_dec), 0)
                if g_diff is not None:
                    gdiff_holder["g"] = g_diff.reshape(1).to(torch.float32).contiguous()

            outs = [dec, diff]
            if want_ids:
                outs += [rec_t["ind"].view(*pre_t.f32.shape[:-1]), rec_b["ind"].view(*pre_b.f32.shape[:-1])]
            if want_pre:    # the fp32 rows the quantisers saw ([F, h, w, embed_dim]); parity tests re-run the argmin on them
                outs += [pre_t.f32, pre_b.f32]
            return tuple(outs), {"seed": seed, "input_grad": lambda: None, "dp": dp,
                                 "non_diff": tuple(range(2, len(outs)))}

        return runner

    def param_forward_order(self):
        """Parameter names in the order the fused graph uses them (the bucket is laid out in reverse)."""
        order = []
        for sub in ("enc_b", "enc_t", "conv3d_encoded_b", "conv3d_encoded_t", "quantize_conv_t", "dec_t",
                    "quantize_conv_b", "upsample_t", "dec"):
            order += [f"{sub}.{k}" for k, _ in getattr(self, sub).named_parameters()]
        return order

    def _forward_clips(self, x4: torch.Tensor, clips: int, want_ids: bool = False, want_pre: bool = False):
        return _run_vqvae(self, self._runner(clips, want_ids, want_pre), x4)

    def forward(self, input):
        """input [T, Cin, H, W] (one clip, reference semantics) or [B, T, Cin, H, W] (B clips, extension).
        Returns (dec like input, diff [1])."""
        if input.dim() == 5:
            b, t = input.shape[:2]
            outs = self._forward_clips(input.reshape(b * t, *input.shape[2:]), b)
            return outs[0].view(b, t, *outs[0].shape[1:]), outs[1]
        outs = self._forward_clips(input, 1)
        return outs[0], outs[1]

    def forward_with_ids(self, input, clips: int = 1, return_pre: bool = False):
        """(dec, diff, id_t, id_b[, pre_t, pre_b]) -- used by parity tests and by the data-parallel trainer.
        ``return_pre`` appends the fp32 pre-quantiser activations (the exact rows the argmin ran on)."""
        return self._forward_clips(input, clips, want_ids=True, want_pre=return_pre)

    # -- reference sub-methods (eager composition of the drop-in modules; used for inference / tests) ----
    def only_encode(self, input):
        enc_b = self.enc_b(input)
        enc_t = self.enc_t(enc_b)
        return enc_b, enc_t

    def encode_quantized(self, enc_b, enc_t):
        quant_t = _conv1x1_module(self.quantize_conv_t, enc_t).permute(0, 2, 3, 1)
        quant_t, diff_t, id_t = self.quantize_t(quant_t)
        quant_t = quant_t.permute(0, 3, 1, 2)
        diff_t = diff_t.unsqueeze(0)
        dec_t = self.dec_t(quant_t)
        enc_b = torch.cat([dec_t, enc_b], 1)
        quant_b = _conv1x1_module(self.quantize_conv_b, enc_b).permute(0, 2, 3, 1)
        quant_b, diff_b, id_b = self.quantize_b(quant_b)
        quant_b = quant_b.permute(0, 3, 1, 2)
        diff_b = diff_b.unsqueeze(0)
        return quant_t, quant_b, diff_t + diff_b, id_t, id_b

    def decode(self, quant_t, quant_b):
        upsample_t = _convT_module(self.upsample_t, quant_t)
        quant = torch.cat([upsample_t, quant_b], 1)
        return self.dec(quant)

    def decode_code(self, code_t, code_b):
        quant_t = self.quantize_t.embed_code(code_t).permute(0, 3, 1, 2)
        quant_b = self.quantize_b.embed_code(code_b).permute(0, 3, 1, 2)
        return self.decode(quant_t, quant_b)


def _run_vqvae(model: VQVAE, runner, x4: torch.Tensor):
    ps = _params_of(model)
    return apply_graph(runner, x4, tuple(ps.keys()), ps.values())


def _conv1x1_module(m: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """Stand-alone 1x1 Conv2d through the same kernels (used by the eager sub-method path)."""
    cout, cin = m.weight.shape[:2]

    def build(tape, xin):
        return conv_op(tape, FORM_S1, 1, [View(xin, False)], "", cout), False

    return _run_graph_named(m, build, cin, x)


def _convT_module(m: nn.ConvTranspose2d, x: torch.Tensor) -> torch.Tensor:
    cin, cout = m.weight.shape[:2]

    def build(tape, xin):
        return conv_op(tape, FORM_UP, 4, [View(xin, False)], "", cout, transposed=True), False

    return _run_graph_named(m, build, cin, x)


def _run_graph_named(m: nn.Module, build, cin: int, x: torch.Tensor):
    ps = {"." + k: v for k, v in m.named_parameters()}
    names = tuple(ps.keys())
    return apply_graph(_io_wrap(build, cin, None), x, names, ps.values())[0]
