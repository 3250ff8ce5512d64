"""A minimal tape over the C-ABI ops: forward builds nodes + backward closures, backward replays the tape.

Why not torch.autograd per op: the backward of every conv here is *fused* -- the data-gradient kernel applies
the ReLU gate of the consumer's input and accumulates into an existing gradient in its epilogue, torch.cat
gradients are channel slices of one wide tensor, and weight gradients go straight into fp32 PyTorch-layout
buffers.  torch.autograd only sees one Function per module call (vqvae.py / lpips.py).

A Node is one activation in channels-last bf16 ([F, H, W, Cs]):
    raw  : pre-activation value (or None if never needed)
    act  : relu(raw)            (or None)
    g    : gradient w.r.t. *raw*, as (tensor, channel_offset); consumers of the relu view gate with act > 0.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from .ops import FORM_DOWN, FORM_S1, FORM_S1_DGRAD, FORM_UP, pad16


class Node:
    __slots__ = ("raw", "act", "c", "g", "f32", "gdec", "g32", "fuse_pool", "pool_dy", "pool_follows", "pooled")

    def __init__(self, c: int, raw=None, act=None, f32=None):
        self.c = c
        self.raw = raw
        self.act = act
        self.f32 = f32
        self.g: Optional[Tuple[torch.Tensor, int]] = None
        self.g32: Optional[torch.Tensor] = None      # gradient already in fp32 NCHW (image-side input of the LPIPS trunk)
        self.fuse_pool = False                       # LPIPS tap whose backward also does the following max pool's (lpips.py)
        self.pool_dy: Optional[torch.Tensor] = None  # ... the pooled gradient left for it by the pool's backward
        self.pool_follows = False                    # set by the VGG trunk on a tap node whose next layer is a max pool
        self.pooled: Optional[torch.Tensor] = None   # maxpool2(act) if the tap's forward kernel already produced it

    @property
    def any(self) -> torch.Tensor:
        return self.raw if self.raw is not None else self.act

    @property
    def cs(self) -> int:
        return self.any.shape[-1]


class View:
    """A consumer's view of a node: the raw value or its ReLU."""
    __slots__ = ("node", "relu")

    def __init__(self, node: Node, relu: bool):
        self.node = node
        self.relu = relu

    @property
    def t(self) -> torch.Tensor:
        t = self.node.act if self.relu else self.node.raw
        assert t is not None, "requested view was not materialised"
        return t


class Tape:
    def __init__(self, params: Dict[str, torch.Tensor], need_grad: bool):
        self.params = params
        self.need_grad = need_grad
        self.grads: Dict[str, torch.Tensor] = {}
        self._bwd: List[Callable[[], None]] = []
        self.precise = ops.PRECISE     # verification mode (ops.PRECISE) is a property of the tape: backward re-enters it
        self.dp = None                 # optional FusedDataParallel: grads are born in its flat bucket
        self._fresh: set = set()       # pre-allocated gradient buffers not written yet

    def record(self, fn: Callable[[], None]):
        if self.need_grad:
            self._bwd.append(fn)

    def backward(self):
        old, ops.PRECISE = ops.PRECISE, self.precise
        try:
            for fn in reversed(self._bwd):
                fn()
        finally:
            ops.PRECISE = old
        self._bwd.clear()

    def grad_buffer(self, name: str) -> Tuple[torch.Tensor, bool]:
        """fp32 gradient buffer for a parameter and whether to accumulate into it."""
        g = self.grads.get(name)
        if g is None:
            g = torch.empty_like(self.params[name], dtype=torch.float32)
            self.grads[name] = g
            return g, False
        if name in self._fresh:
            self._fresh.discard(name)
            return g, False
        return g, True

    def grad_ready(self, name: str):
        if self.dp is not None:
            self.dp.grad_ready(name)


_DGRAD_FORM = {FORM_S1: FORM_S1_DGRAD, FORM_DOWN: FORM_UP, FORM_UP: FORM_DOWN}


def _padded_bias(tape: Tape, name: str, cout: int) -> torch.Tensor:
    b = tape.params[name]
    if b.numel() == pad16(cout):
        return b
    key = ("bias_pad", name)
    cache = tape.__dict__.setdefault("_cache", {})
    hit = cache.get(key)
    if hit is None:
        hit = torch.zeros(pad16(cout), dtype=torch.float32, device=b.device)
        hit[:cout] = b.detach()
        cache[key] = hit
    return hit


def _as5(t: torch.Tensor, clips: int) -> torch.Tensor:
    f = t.shape[0]
    return t.view(clips, f // clips, *t.shape[1:])


def conv_op(tape: Tape, form: int, ksize: int, srcs: Sequence[View], wname: str, cout: int, *, transposed: bool = False,
            want_raw: bool = True, want_relu: bool = False, residual: Optional[Node] = None, f32: Optional[str] = None,
            ndim: int = 2, clips: int = 1, input_needs_grad: bool = True, bias: bool = True,
            param_grad: bool = True) -> Node:
    """One convolution layer (forward now, backward recorded).

    ``wname`` is the state-dict prefix ("enc_b.blocks.0"): weight = wname.weight, bias = wname.bias.
    ``transposed``: the parameter is a ConvTranspose2d weight [Cin, Cout, 4, 4] (form must be FORM_UP).
    ``residual``: node whose raw value is added to the output (ResBlock ``out += input``).
    """
    w = tape.params[wname + ".weight"]
    b = _padded_bias(tape, wname + ".bias", cout) if bias else None
    n_axis = 1 if transposed else 0

    def shaped(t):
        return _as5(t, clips) if ndim == 3 else t

    src_list = [(shaped(v.t), v.node.c, 0) for v in srcs]
    raw, relu, of32 = ops.conv(form, ndim, ksize, src_list, w, n_axis, cout, bias=b,
                               addend=None if residual is None else shaped(residual.raw),
                               want_raw=want_raw, want_relu=want_relu, f32=f32)

    def flat(t):
        return None if t is None else (t.view(-1, *t.shape[2:]) if ndim == 3 else t)

    out = Node(cout, raw=flat(raw), act=flat(relu), f32=of32 if f32 == "nchw" else flat(of32))

    def backward():
        if out.g is None and not param_grad:
            return      # frozen layer (LPIPS trunk) that no gradient reached: nothing to do
        assert out.g is not None, f"no gradient reached {wname}"
        gt, g_off = out.g
        dy = (shaped(gt), cout, g_off)
        # weight gradient (one launch per concatenated source) with the bias gradient fused into the first launch when
        # P = dy (the tensor core sums dy's columns with an all-ones operand); ConvTranspose2d keeps the column-sum kernel
        gw, acc = tape.grad_buffer(wname + ".weight") if param_grad else (None, False)
        gb, bacc = tape.grad_buffer(wname + ".bias") if (bias and param_grad) else (None, False)
        fused_bias_ok = form == FORM_S1 or (form == FORM_DOWN and 4 * pad16(srcs[0].node.c) + 16 <= 512)
        if gb is not None and not fused_bias_ok:
            ops.colsum(gt, cout, gb, c_off=g_off, accumulate=bacc)
            gb = None
        w_off = 0
        for v, s in zip(srcs, src_list):
            if not param_grad:
                break
            if form == FORM_S1 and len(srcs) == 1 and pad16(cout) < pad16(v.node.c):
                # narrow output (ResBlock 3x3 C->32): make the wide tensor x the M side of the MMA and read dy at
                # pix - tap; the (tiny) bias gradient then comes from the column-sum kernel
                if gb is not None:
                    ops.colsum(gt, cout, gb, c_off=g_off, accumulate=bacc)
                ops.wgrad(FORM_S1, ndim, ksize, s, dy, gw, m_axis=1, q_w_off=0, accumulate=acc, q_shift_sign=-1)
            elif form == FORM_S1:
                ops.wgrad(FORM_S1, ndim, ksize, dy, s, gw, m_axis=0, q_w_off=w_off, accumulate=acc, dbias=gb,
                          dbias_accumulate=bacc)
            elif form == FORM_DOWN:
                ops.wgrad(FORM_DOWN, 2, 4, dy, s, gw, m_axis=0, q_w_off=w_off, accumulate=acc, dbias=gb,
                          dbias_accumulate=bacc)
            else:  # FORM_UP / ConvTranspose2d weight [cin, cout]: P = x (low res), Q = dy (hi res)
                assert len(srcs) == 1
                ops.wgrad(FORM_DOWN, 2, 4, s, dy, gw, m_axis=0, q_w_off=0, accumulate=acc)
            gb = None
            w_off += v.node.c
        if param_grad:
            tape.grad_ready(wname + ".weight")
            if bias:
                tape.grad_ready(wname + ".bias")
        # residual pass-through
        if residual is not None:
            assert g_off == 0
            if residual.g is None:
                residual.g = (gt, 0)
            else:
                residual.g = (ops.add_grads(residual.g[0], gt), 0)
        # data gradient
        if not input_needs_grad:
            return
        cin_total = sum(v.node.c for v in srcs)
        dform = _DGRAD_FORM[form]
        dn_axis = 0 if transposed else 1
        if len(srcs) == 1:
            v = srcs[0]
            nd = v.node
            mask = shaped(nd.act) if v.relu else None
            addend = None
            if nd.g is not None:
                assert nd.g[1] == 0 and nd.g[0].shape[-1] == nd.cs
                addend = shaped(nd.g[0])
            dx, _, _ = ops.conv(dform, ndim, ksize, [dy], w, dn_axis, cin_total, mask=mask, addend=addend,
                                out_cs=nd.cs)
            nd.g = (flat(dx), 0)
        else:
            dx, _, _ = ops.conv(dform, ndim, ksize, [dy], w, dn_axis, cin_total)
            dx = flat(dx)
            off = 0
            for v in srcs:
                assert v.node.g is None and not v.relu, "torch.cat sources must have a single consumer"
                v.node.g = (dx, off)
                off += v.node.c

    tape.record(backward)
    return out


def _s2_direct(img: torch.Tensor, c_img: int, c_feat: int) -> bool:
    """The image-side stride-2 layers run without an im2col matrix when the shapes are the kernels' (csrc/small_cin.cu)."""
    return (not ops.PRECISE and c_img in (3, 6) and c_feat == 64 and img.dtype == torch.float32 and img.dim() == 4
            and img.shape[2] % 2 == 0 and img.shape[3] % 2 == 0)


def first_conv_op(tape: Tape, x_nchw: torch.Tensor, wname: str, cin: int, cout: int) -> Node:
    """Conv2d(cin<=8 -> cout, 4, stride 2, pad 1) + ReLU on the raw fp32 NCHW input (reference
    models/vqvae_conv3d_latent.py:109).  The input needs no gradient.  3 / 6 input channels and 64 output channels (the
    reference's configuration): the kernels of csrc/small_cin.cu build the im2col tile in shared memory, forward and
    weight gradient read the image itself.  Otherwise (and in verification mode, whose activations are hi|lo pairs):
    explicit im2col (K = 16 taps x 8 = 128) + a 1x1 GEMM on the tcgen05 kernel."""
    w = tape.params[wname + ".weight"]          # [cout, cin, 4, 4]
    if _s2_direct(x_nchw, cin, cout):
        act = ops.s2conv(x_nchw, cin, w.detach(), tape.params[wname + ".bias"].detach(), relu=True)
        out = Node(cout, act=act)

        def backward_direct():
            gt, g_off = out.g
            if g_off != 0 or gt.shape[-1] != cout:
                gt = gt[..., g_off:g_off + cout].contiguous()
            gw, acc_w = tape.grad_buffer(wname + ".weight")
            gb, acc_b = tape.grad_buffer(wname + ".bias")
            ops.s2wgrad(x_nchw, cin, gt.view(act.shape), gw, accumulate=acc_w, dbias=gb, dbias_accumulate=acc_b)
            tape.grad_ready(wname + ".weight")
            tape.grad_ready(wname + ".bias")

        tape.record(backward_direct)
        return out
    b = _padded_bias(tape, wname + ".bias", cout)
    col = ops.im2col4x4s2(x_nchw, cin)          # [F, H/2, W/2, 128]
    w2 = torch.zeros(cout, 16, 8, dtype=torch.float32, device=w.device)
    w2[:, :, :cin] = w.detach().reshape(cout, cin, 16).permute(0, 2, 1)
    w2 = w2.view(cout, 128, 1, 1)
    _, act, _ = ops.conv(FORM_S1, 2, 1, [(col, 128, 0)], w2, 0, cout, bias=b, want_raw=False, want_relu=True,
                         wkey=(w, "im2col"))
    out = Node(cout, act=act)

    def backward():
        gt, g_off = out.g
        gb, acc = tape.grad_buffer(wname + ".bias")
        dw2 = torch.empty(cout, 128, 1, 1, dtype=torch.float32, device=w.device)
        # bias gradient fused into the weight-gradient launch (all-ones MMA operand)
        ops.wgrad(FORM_S1, 2, 1, (gt, cout, g_off), (col, 128, 0), dw2, m_axis=0, dbias=gb, dbias_accumulate=acc)
        gw, acc = tape.grad_buffer(wname + ".weight")
        dw = dw2.view(cout, 16, 8)[:, :, :cin].permute(0, 2, 1).reshape(cout, cin, 4, 4)
        if acc:
            gw.add_(dw)
        else:
            gw.copy_(dw)
        tape.grad_ready(wname + ".weight")
        tape.grad_ready(wname + ".bias")

    tape.record(backward)
    return out


def last_convT_op(tape: Tape, src: View, wname: str, cout: int) -> Node:
    """ConvTranspose2d(cin -> cout<=8, 4, stride 2, pad 1) producing the fp32 NCHW reconstruction (reference :154-156):
    1x1 GEMM to the [pixels, 16 taps x 8] column matrix (bf16) + col2im scatter (+bias).  Backward: im2col of the
    incoming NCHW gradient, then 1x1 data / weight gradients.  ``out.gdec`` must hold the fp32 NCHW gradient."""
    w = tape.params[wname + ".weight"]          # [cin, cout, 4, 4]
    b = tape.params[wname + ".bias"]
    cin = w.shape[0]
    w2 = torch.zeros(16, 8, cin, dtype=torch.float32, device=w.device)
    w2[:, :cout] = w.detach().reshape(cin, cout, 16).permute(2, 1, 0)
    w2 = w2.view(128, cin, 1, 1)                # Conv weight [Cout'=128, Cin]
    x = src.t
    col, _, _ = ops.conv(FORM_S1, 2, 1, [(x, cin, 0)], w2, 0, 128, wkey=(w, "col2im"))
    dec = ops.col2im4x4s2(col, b.detach().contiguous(), cout)
    del col
    out = Node(cout, f32=dec)
    out.gdec = None

    def backward():
        g = out.gdec
        assert g is not None, "no gradient reached the reconstruction"
        gb, acc = tape.grad_buffer(wname + ".bias")
        ops.chansum_nchw(g, cout, gb, accumulate=acc)
        nd = src.node
        if _s2_direct(g, cout, cin) and g.is_contiguous() and x.is_contiguous() and x.shape[-1] == cin and nd.cs == cin:
            # both gradients straight from the fp32 NCHW output gradient (no im2col matrix): the ConvTranspose2d weight
            # [cin, cout, 4, 4] is the Conv2d weight of its own data gradient
            gw, acc = tape.grad_buffer(wname + ".weight")
            xa = x.view(g.shape[0], g.shape[2] // 2, g.shape[3] // 2, cin)
            ops.s2wgrad(g, cout, xa, gw, accumulate=acc)
            tape.grad_ready(wname + ".weight")
            tape.grad_ready(wname + ".bias")
            addend = None
            if nd.g is not None:
                addend = nd.g[0] if (nd.g[1] == 0 and nd.g[0].shape[-1] == cin) else nd.g[0][..., nd.g[1]:nd.g[1] + cin].contiguous()
                addend = addend.view(xa.shape)
            dx = ops.s2conv(g, cout, w.detach(), None, mask=nd.act.view(xa.shape) if src.relu else None, addend=addend)
            nd.g = (dx.view(x.shape), 0)
            return
        dcol = ops.im2col4x4s2(g, cout)         # [F, h, w, 128]
        dw2 = torch.empty(128, cin, 1, 1, dtype=torch.float32, device=w.device)
        ops.wgrad(FORM_S1, 2, 1, (dcol, 128, 0), (x, cin, 0), dw2, m_axis=0)
        gw, acc = tape.grad_buffer(wname + ".weight")
        dw = dw2.view(16, 8, cin)[:, :cout].permute(2, 1, 0).reshape(cin, cout, 4, 4)
        if acc:
            gw.add_(dw)
        else:
            gw.copy_(dw)
        tape.grad_ready(wname + ".weight")
        tape.grad_ready(wname + ".bias")
        nd = src.node
        addend = nd.g[0] if nd.g is not None else None
        dx, _, _ = ops.conv(FORM_S1_DGRAD, 2, 1, [(dcol, 128, 0)], w2, 1, cin, mask=nd.act if src.relu else None,
                            addend=addend, out_cs=nd.cs, wkey=(w, "col2im"))
        nd.g = (dx, 0)

    tape.record(backward)
    return out


def resblock_op(tape: Tape, x: Node, prefix: str, channel: int, n_res_channel: int, last: bool) -> Node:
    """ResBlock (reference models/vqvae_conv3d_latent.py:86-101): relu -> 3x3 -> relu -> 1x1 -> += input.
    ``last``: the block is followed by the stack's final in-place ReLU, so only relu(out) is materialised."""
    h = conv_op(tape, FORM_S1, 3, [View(x, True)], prefix + ".conv.1", n_res_channel, want_raw=False, want_relu=True)
    return conv_op(tape, FORM_S1, 1, [View(h, True)], prefix + ".conv.3", channel, residual=x, want_raw=not last,
                   want_relu=True)
