"""Drop-in replacements for the reference's models/lpips.py (LPIPS, ScalingLayer, NetLinLayer, vgg16) and
loss.py:VQLPIPS.  Same constructor arguments, forward signatures and state-dict keys; the VGG16 trunk runs on
the tcgen05 implicit-GEMM conv kernel, the LPIPS head is one fused bandwidth-bound reduction per tap
(csrc/lpips.cu), and only the ``input`` side keeps activations for backward (all LPIPS parameters are frozen,
reference models/lpips.py:63-64, so there is no weight gradient).
"""
from __future__ import annotations

import os
import warnings
from collections import namedtuple

import torch
from torch import nn

from . import ops
from .graph import FORM_S1, FORM_S1_DGRAD, Node, Tape, View, conv_op
from ._lib import FaceoffB200Error
from .vqvae import _params_of, apply_graph

_VGG_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512]
_SLICE_ENDS = [4, 9, 16, 23, 30]  # reference models/lpips.py:127-136


CKPT_MAP = {"vgg_lpips": "vgg.pth"}
MD5_MAP = {"vgg_lpips": "d507d7349b931f0638a25a48a722f98a"}   # reference models/lpips.py:20-22


def get_ckpt_path(name, root, check=False):
    """Path of the LPIPS checkpoint if it is already on disk (reference models/lpips.py:40-48 would download it; this
    environment has no network), else None.  ``check`` verifies the reference's md5."""
    assert name in CKPT_MAP
    path = os.path.join(root, CKPT_MAP[name])
    if not os.path.exists(path):
        return None
    if check:
        import hashlib

        with open(path, "rb") as f:
            md5 = hashlib.md5(f.read()).hexdigest()
        if md5 != MD5_MAP[name]:
            raise FaceoffB200Error(f"{path}: md5 {md5} != {MD5_MAP[name]}")
    return path


class ScalingLayer(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("shift", torch.Tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer("scale", torch.Tensor([.458, .448, .450])[None, :, None, None])

    def forward(self, inp):
        return (inp - self.shift) / self.scale


class NetLinLayer(nn.Module):
    """A single linear layer which does a 1x1 conv (parameter holder; evaluated inside the fused tap kernel)."""

    def __init__(self, chn_in, chn_out=1, use_dropout=False):
        super().__init__()
        layers = [nn.Dropout()] if use_dropout else []
        layers += [nn.Conv2d(chn_in, chn_out, 1, stride=1, padding=0, bias=False)]
        self.model = nn.Sequential(*layers)


class vgg16(nn.Module):
    """VGG16 feature trunk split in the 5 LPIPS slices; module indices follow torchvision's vgg16.features so
    the state-dict keys (net.slice{k}.{idx}.weight) match the reference (models/lpips.py:115-136)."""

    def __init__(self, requires_grad=False, pretrained=True):
        super().__init__()
        self.N_slices = 5
        self.trunk_loaded = False   # True once real (non-random) trunk weights were loaded
        slices = [nn.Sequential() for _ in range(5)]
        idx, cin, which = 0, 3, 0
        self.layout = []  # (kind, key, cout)
        for v in _VGG_CFG:
            if v == "M":
                if idx >= _SLICE_ENDS[which]:
                    which += 1
                slices[which].add_module(str(idx), nn.MaxPool2d(kernel_size=2, stride=2))
                self.layout.append(("pool", None, cin))
                idx += 1
            else:
                if idx >= _SLICE_ENDS[which]:
                    which += 1
                slices[which].add_module(str(idx), nn.Conv2d(cin, v, 3, padding=1))
                slices[which].add_module(str(idx + 1), nn.ReLU(inplace=True))
                self.layout.append(("conv", f"slice{which + 1}.{idx}", v))
                idx += 2
                cin = v
            if idx in _SLICE_ENDS:
                self.layout.append(("tap", None, cin))
        self.slice1, self.slice2, self.slice3, self.slice4, self.slice5 = slices
        if pretrained:
            self._load_torchvision_features()
        if not requires_grad:
            for p in self.parameters():
                p.requires_grad = False

    def _load_torchvision_features(self):
        """The reference builds the trunk from ``torchvision.models.vgg16(pretrained=True).features``
        (models/lpips.py:118).  There is no network here, so only an already cached torchvision checkpoint
        (``<torch hub>/checkpoints/vgg16-*.pth``) can be used; otherwise the trunk stays RANDOM and every consumer is
        told so loudly (``trunk_loaded`` stays False, LPIPS warns)."""
        ckpt_dir = os.path.join(torch.hub.get_dir(), "checkpoints")
        cands = sorted(f for f in (os.listdir(ckpt_dir) if os.path.isdir(ckpt_dir) else []) if f.startswith("vgg16-"))
        if not cands:
            return
        self.load_torchvision_state_dict(torch.load(os.path.join(ckpt_dir, cands[0]), map_location="cpu"))

    def load_torchvision_state_dict(self, sd):
        """Load a torchvision ``vgg16`` (or ``vgg16().features``) state_dict: key ``features.{idx}.weight`` ->
        ``slice{k}.{idx}.weight`` with k the LPIPS slice holding module ``idx`` (reference models/lpips.py:127-136)."""
        own = dict(self.named_parameters())
        hit = 0
        for k, v in sd.items():
            k = k[len("features."):] if k.startswith("features.") else k
            parts = k.split(".")
            if len(parts) != 2 or not parts[0].isdigit():
                continue
            idx = int(parts[0])
            if idx >= _SLICE_ENDS[-1]:
                continue
            which = next(i for i, e in enumerate(_SLICE_ENDS) if idx < e) + 1
            name = f"slice{which}.{idx}.{parts[1]}"
            if name in own:
                with torch.no_grad():
                    own[name].copy_(v)
                hit += 1
        if hit != len(own):
            raise FaceoffB200Error(f"vgg16: state_dict covered {hit} of {len(own)} trunk tensors")
        self.trunk_loaded = True

    def forward(self, X):
        """[N,3,H,W] (already scaled) -> VggOutputs(relu1_2, relu2_2, relu3_3, relu4_3, relu5_3), NCHW fp32
        (reference models/lpips.py:139-152); gradients flow back to X."""
        net = self

        def runner(tape: Tape, x: torch.Tensor):
            holders = []

            def tap_hook(k, node):
                h = {}
                holders.append(h)

                def tap_bwd():
                    g = h.get("g")
                    if g is None:
                        return
                    g = ops.pack_nchw(g, cs=node.act.shape[-1])
                    # the gradient arrives w.r.t. the tap = relu(raw), node.g is w.r.t. raw: close the tap's own ReLU gate
                    # (the pool gradient it meets in add_grads is gated already); rare stand-alone path, plain torch
                    t = node.act
                    if tape.precise:
                        m = t[..., :t.shape[-1] // 2] > 0
                        g = g * torch.cat([m, m], -1)
                    else:
                        g = g * (t > 0)
                    node.g = (g if node.g is None else ops.add_grads(node.g[0], g), 0)

                tape.record(tap_bwd)

            xin, taps = _vgg_trunk(tape, net.layout, "", x, None, None, tap_hook)
            outs = tuple(ops.unpack_nchw(t.act, t.c) for t in taps)

            def seed(tape_, gouts):
                for h, g in zip(holders, gouts):
                    if g is not None:
                        h["g"] = g.to(torch.float32).contiguous()

            def input_grad():
                if xin.g32 is not None:
                    return xin.g32
                return ops.unpack_nchw(xin.g[0], 3) if xin.g is not None else None

            return outs, {"seed": seed, "input_grad": input_grad}

        ps = _params_of(self)
        outs = apply_graph(runner, X, tuple(ps.keys()), ps.values())
        vgg_outputs = namedtuple("VggOutputs", ["relu1_2", "relu2_2", "relu3_3", "relu4_3", "relu5_3"])
        return vgg_outputs(*outs)


# Experiments only: FO_LPIPS_FUSE=0 keeps the separate kernels (maxpool2 / maxpool2_bwd next to the tap kernels, the first
# conv's data gradient as an implicit GEMM) for same-box A/B timing.
_FUSE = os.environ.get("FO_LPIPS_FUSE", "1") != "0"


def _vgg_trunk(tape: Tape, layout, prefix: str, x: torch.Tensor, shift, scale, tap_hook=None):
    """(Optional scaling layer +) VGG16 trunk on channels-last bf16.  ``tap_hook(k, node)`` is called at each of the
    five taps, at the tape position of the tap (so whatever it records is replayed after the following pool's
    backward).  Returns (input node, tap nodes)."""
    gen = _vgg_trunk_gen(tape, layout, prefix, x, shift, scale)
    try:
        while True:
            k, node = next(gen)
            if tap_hook is not None:
                tap_hook(k, node)
    except StopIteration as stop:
        return stop.value


def _vgg_trunk_gen(tape: Tape, layout, prefix: str, x: torch.Tensor, shift, scale):
    """The trunk as a generator: yields (k, node) at each tap and resumes with the layer after it, so that two trunks (the
    two sides of LPIPS) can run in lockstep and one kernel can serve the tap and both following pools.  Its return value
    (StopIteration.value) is (input node, tap nodes)."""
    x = x.to(torch.float32).contiguous()
    xin = Node(3)
    cur, cur_relu = xin, False
    taps = []
    first = True
    for li, (kind, key, ch) in enumerate(layout):
        if kind == "conv" and first and tape.precise:
            # verification mode: the plain implicit-GEMM path on the 3-channel (hi|lo pair) input
            first = False
            xin.raw = ops.pack_nchw(x, shift=shift, scale=scale)
            cur = conv_op(tape, FORM_S1, 3, [View(xin, False)], prefix + key, ch, want_raw=False, want_relu=True,
                          param_grad=False)
            cur_relu = True
        elif kind == "conv" and first:
            # first conv (3 -> 64): one kernel from the fp32 NCHW image (ScalingLayer folded in, A tile built in shared
            # memory, weights resident); no im2col matrix in HBM (csrc/small_cin.cu)
            first = False
            w = tape.params[prefix + key + ".weight"]
            b = tape.params[prefix + key + ".bias"]
            if ch == 64 and x.shape[1] == 3:
                act = ops.vgg_first_conv(x.contiguous(), w.detach().contiguous(), b.detach(), shift, scale)
            else:   # other widths: explicit im2col (K = 27 -> 32) + 1x1 GEMM
                col = ops.im2col3x3(x, shift, scale)
                w2 = torch.zeros(ch, 32, dtype=torch.float32, device=w.device)
                w2[:, :27] = w.detach().permute(0, 2, 3, 1).reshape(ch, 27)
                _, act, _ = ops.conv(FORM_S1, 2, 1, [(col, 32, 0)], w2.view(ch, 32, 1, 1), 0, ch, bias=b,
                                     want_raw=False, want_relu=True, wkey=(w, "im2col3x3"))
                del col
            cur = Node(ch, act=act)

            def first_bwd(node=cur, w=w):
                if node.g is None:
                    return
                gt, g_off = node.g
                if _FUSE and node.c == 64 and g_off == 0 and gt.shape[-1] == 64 and gt.is_contiguous() and w.is_contiguous():
                    # taps on the N axis + in-tile shift-add, fp32 NCHW out with the ScalingLayer's division folded in
                    # (csrc/small_cin.cu vgg_first_dgrad_kernel)
                    sc = None if scale is None else scale.detach().reshape(-1).to(torch.float32).contiguous()
                    xin.g32 = ops.vgg_first_dgrad(gt, w.detach(), sc)
                    return
                dx, _, _ = ops.conv(FORM_S1_DGRAD, 2, 3, [(gt, node.c, g_off)], w, 1, 3, out_cs=16)
                xin.g = (dx, 0)

            tape.record(first_bwd)
            cur_relu = True
        elif kind == "conv":
            cur = conv_op(tape, FORM_S1, 3, [View(cur, cur_relu)], prefix + key, ch, want_raw=False, want_relu=True,
                          param_grad=False)
            cur_relu = True
        elif kind == "tap":
            cur.pool_follows = li + 1 < len(layout) and layout[li + 1][0] == "pool"
            yield len(taps), cur
            taps.append(cur)
        else:  # 2x2 max pool (the tap's forward kernel may have produced it already: ops.lpips_tap_pool)
            src = cur
            pooled = Node(ch, raw=src.pooled if src.pooled is not None else ops.maxpool2(src.act))
            src.pooled = None

            def pool_bwd(src=src, pooled=pooled):
                if pooled.g is None:
                    return
                assert src.g is None
                pg, p_off = pooled.g
                if src.fuse_pool and p_off == 0 and pg.shape == pooled.raw.shape and pg.is_contiguous():
                    src.pool_dy = pg      # the tap's backward (recorded just before this pool) folds the pool gradient in
                    return
                src.g = (ops.maxpool2_bwd(src.act, pooled.raw, pg), 0)

            tape.record(pool_bwd)
            cur, cur_relu = pooled, False
    return xin, taps


def normalize_tensor(x, eps=1e-10):
    norm_factor = torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True))
    return x / (norm_factor + eps)


def spatial_average(x, keepdim=True):
    return x.mean([2, 3], keepdim=keepdim)


class LPIPS(nn.Module):
    """Learned perceptual metric.  forward(input, target) -> [N,1,1,1] (reference models/lpips.py:80-93)."""

    def __init__(self, use_dropout=True):
        super().__init__()
        self.scaling_layer = ScalingLayer()
        self.chns = [64, 128, 256, 512, 512]
        self.net = vgg16(pretrained=True, requires_grad=False)
        self.lin0 = NetLinLayer(self.chns[0], use_dropout=use_dropout)
        self.lin1 = NetLinLayer(self.chns[1], use_dropout=use_dropout)
        self.lin2 = NetLinLayer(self.chns[2], use_dropout=use_dropout)
        self.lin3 = NetLinLayer(self.chns[3], use_dropout=use_dropout)
        self.lin4 = NetLinLayer(self.chns[4], use_dropout=use_dropout)
        self.load_from_pretrained()
        for param in self.parameters():
            param.requires_grad = False

    def load_from_pretrained(self, name="vgg_lpips"):
        """The reference downloads vgg.pth (models/lpips.py:66-69) -- to my reading a file holding the five ``lin*``
        layers only; its trunk comes from torchvision's ImageNet checkpoint (:118).  There is no network here: the file
        is loaded if it already sits at the reference's relative path, and the module says loudly which parts are
        still random."""
        ckpt = get_ckpt_path(name, "taming/modules/autoencoder/lpips")
        if ckpt is None:
            warnings.warn(f"LPIPS: {CKPT_MAP[name]} not found under taming/modules/autoencoder/lpips and downloads are "
                          "disabled; the lin layers" + ("" if self.net.trunk_loaded else " AND the VGG16 trunk") +
                          " stay randomly initialised (load_state_dict / net.load_torchvision_state_dict explicitly)")
            return
        self._load_checkpoint(ckpt)

    def _load_checkpoint(self, ckpt):
        sd = torch.load(ckpt, map_location=torch.device("cpu"))
        res = self.load_state_dict(sd, strict=False)
        if any(k.startswith("net.") for k in sd) and not any(k.startswith("net.") for k in res.missing_keys):
            self.net.trunk_loaded = True
        missing_lin = [k for k in res.missing_keys if k.startswith("lin")]
        if missing_lin:
            raise FaceoffB200Error(f"LPIPS: {ckpt} does not hold {missing_lin}")
        if not self.net.trunk_loaded:
            warnings.warn(f"LPIPS: {ckpt} carries no VGG16 trunk (net.*) and no cached torchvision vgg16 checkpoint was "
                          "found: the trunk is RANDOM and the perceptual loss meaningless until "
                          "net.load_torchvision_state_dict(torchvision vgg16 state_dict) is called")

    @classmethod
    def from_pretrained(cls, name="vgg_lpips"):
        """reference models/lpips.py:71-78"""
        if name != "vgg_lpips":
            raise NotImplementedError
        model = cls()
        ckpt = get_ckpt_path(name, "taming/modules/autoencoder/lpips")
        if ckpt is not None:
            model._load_checkpoint(ckpt)
        return model

    def _trunk_gen(self, tape: Tape, x: torch.Tensor):
        """The trunk as a generator over its taps (see ``_vgg_trunk_gen``)."""
        shift = self.scaling_layer.shift.reshape(-1).contiguous()
        scale = self.scaling_layer.scale.reshape(-1).contiguous()
        return _vgg_trunk_gen(tape, self.net.layout, "net.", x, shift, scale)

    def _trunk(self, tape: Tape, x: torch.Tensor, tap_hook=None):
        """Scaling layer + VGG16 trunk on channels-last bf16 (see ``_vgg_trunk``)."""
        shift = self.scaling_layer.shift.reshape(-1).contiguous()
        scale = self.scaling_layer.scale.reshape(-1).contiguous()
        return _vgg_trunk(tape, self.net.layout, "net.", x, shift, scale, tap_hook)

    def forward(self, input, target):
        """LPIPS distance [N,1,1,1].  The value is symmetric in (input, target); gradients flow to whichever side
        requires them (the reference trainer passes (ground_truth, reconstruction), train_faceoff_perceptual.py:42)."""
        gi = torch.is_grad_enabled() and input.requires_grad
        gt_ = torch.is_grad_enabled() and target.requires_grad
        if gt_ and not gi:
            return self._distance(target, input)
        if gi and gt_:
            # d/d(input) with target fixed + d/d(target) with input fixed; the value is counted once
            return (self._distance(input, target.detach()) + self._distance(target, input.detach())
                    - self._distance(input.detach(), target.detach()))
        return self._distance(input, target)

    def _distance(self, diff_side, fixed_side):
        """Differentiable w.r.t. ``diff_side`` only."""
        lins = [self.lin0, self.lin1, self.lin2, self.lin3, self.lin4]
        ws = [l.model[-1].weight.reshape(-1).to(torch.float32).contiguous() for l in lins]
        model = self

        def runner(tape: Tape, x: torch.Tensor):
            n = x.shape[0]
            # fixed side: no gradient, nothing kept but the five taps.  Its trunk runs in lockstep with the differentiated
            # side's (a generator advanced to tap k inside the other side's tap hook), so that ONE kernel computes the tap
            # and both sides' following max pools.
            t_tape = Tape(tape.params, need_grad=False)
            fixed = model._trunk_gen(t_tape, fixed_side.detach())
            val = torch.zeros(n, dtype=torch.float32, device=x.device)
            g_holder = {}

            def tap_hook(k, node):
                k1, node1 = next(fixed)
                assert k1 == k
                f1, w = node1.act, ws[k]
                # a tap that feeds a max pool takes over the pools, forward (ops.lpips_tap_pool writes both pooled tensors)
                # and backward (ops.lpips_tap_bwd_pool): one pass over the feature maps instead of two each way; the
                # verification mode keeps the separate kernels
                node.fuse_pool = (_FUSE and node.pool_follows and not tape.precise and node.act.shape[1] % 2 == 0
                                  and node.act.shape[2] % 2 == 0 and node.act.is_contiguous() and f1.is_contiguous())
                if node.fuse_pool:
                    node.pooled, node1.pooled = ops.lpips_tap_pool(node.act, f1, w, val, pool_f1=node1.pool_follows)
                else:
                    ops.lpips_tap(node.act, f1, w, val)

                def tap_bwd():
                    if node.pool_dy is not None:
                        assert node.g is None
                        node.g = (ops.lpips_tap_bwd_pool(node.act, f1, w, g_holder["g"], node.pool_dy), 0)
                        node.pool_dy = None
                        return
                    addend = node.g[0] if node.g is not None else None
                    node.g = (ops.lpips_tap_bwd(node.act, f1, w, g_holder["g"], addend), 0)

                tape.record(tap_bwd)

            xin, _ = model._trunk(tape, x, tap_hook)
            fixed.close()

            def seed(tape_, gouts):
                g = gouts[0]
                g_holder["g"] = (torch.zeros(n, device=x.device) if g is None
                                 else g.reshape(n).to(torch.float32).contiguous())

            def input_grad():
                if xin.g32 is not None:
                    return xin.g32           # already divided by the ScalingLayer's scale
                if xin.g is None:
                    return None
                return ops.unpack_nchw(xin.g[0], 3) / model.scaling_layer.scale

            return (val.view(n, 1, 1, 1),), {"seed": seed, "input_grad": input_grad}

        ps = _params_of(self)
        return apply_graph(runner, diff_side, tuple(ps.keys()), ps.values())[0]


class VQLPIPS(nn.Module):
    """reference loss.py:27-33"""

    def __init__(self):
        super().__init__()
        self.perceptual_loss = LPIPS().eval()

    def forward(self, targets, reconstructions):
        return self.perceptual_loss(targets.contiguous(), reconstructions.contiguous()).mean()
