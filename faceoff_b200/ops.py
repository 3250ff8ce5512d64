"""Thin tensor-level wrappers over the C ABI (faceoff_b200/_lib.py).

Activations are channels-last bf16 tensors shaped [N, H, W, Cs] (2-D) or [N, D, H, W, Cs] (3-D),
contiguous, Cs a multiple of 16; logical channel counts are passed explicitly.  Everything here runs on the
current CUDA stream and never synchronises the host.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib as L
from ._lib import FORM_DOWN, FORM_S1, FORM_S1_DGRAD, FORM_UP, ConvDesc, Src, WgradDesc  # noqa: F401


# Instrumentation used by bench.py: when PROFILE is a dict, every tensor-core launch is bracketed by CUDA events on the
# launching stream (name -> [(start, stop, algorithmic FLOPs)]); LAUNCHES counts kernels launched through this module.
PROFILE: Optional[dict] = None
LAUNCHES = 0


def _count(n: int):
    global LAUNCHES
    LAUNCHES += n


_event_pool: list = []   # timing events are recycled: creating two CUDA events per launch costs more host time than the launch


def _take_event():
    return _event_pool.pop() if _event_pool else torch.cuda.Event(enable_timing=True)


def recycle_events(profile: dict):
    """Give the events of a finished PROFILE dict back to the pool (after their times were read)."""
    seen = set()
    for recs in profile.values():
        for a, b, _ in recs:
            if id(a) not in seen:
                seen.add(id(a))
                _event_pool.extend((a, b))


class _Timed:
    __slots__ = ("name", "work", "a", "b")

    def __init__(self, name: str, work: float):
        self.name, self.work = name, work

    def __enter__(self):
        if PROFILE is not None:
            self.a = _take_event()
            self.b = _take_event()
            self.a.record()

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.b.record()
            PROFILE.setdefault(self.name, []).append((self.a, self.b, self.work))


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> C.c_void_p:
    """cudaStream_t of PyTorch's current stream on the current device (raw handle: ~20x cheaper than building a
    torch.cuda.Stream object per launch)."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def pad16(c: int) -> int:
    return (c + 15) // 16 * 16


# ------------------------------------------------------------------------------------------------
# Verification mode (tests only): fp32-accurate activations as hi|lo bf16 pairs through the SAME tensor-core kernels
# ------------------------------------------------------------------------------------------------
# With PRECISE set, every activation / activation-gradient tensor produced by this module holds an fp32 value per
# logical channel as two bf16 numbers: [.., 2*Cp] with hi = bf16(v) in channels [0, Cp) and lo = bf16(v - hi) in
# [Cp, 2*Cp) (Cp = pad16(C)).  A convolution lists each source three times (hi, lo, hi) against weights packed as
# (w_hi, w_hi, w_lo), i.e. hi*w_hi + lo*w_hi + hi*w_lo with fp32 accumulation (the term lo*w_lo ~ 2^-18 is dropped);
# weight gradients are three accumulating launches (P_hi Q_hi + P_lo Q_hi + P_hi Q_lo).  The planner, K-step tables,
# packed-weight order, halo / pair / parity addressing and the epilogue fusions are exactly the product's, so a wrong
# tap, a missed residual or a mis-scaled bias gradient shows up at the 1e-4 level against an fp64 CPU restatement instead of
# hiding inside the bf16 tolerance.  ~3-4x slower than the product path; never enabled outside tests.
PRECISE = False


class precise_mode:
    """``with ops.precise_mode(): loss = step(...); loss.backward()`` -- see PRECISE above."""

    def __enter__(self):
        global PRECISE
        self._old, PRECISE = PRECISE, True
        return self

    def __exit__(self, *exc):
        global PRECISE
        PRECISE = self._old


def split_cl(x: torch.Tensor, c: Optional[int] = None, cp: Optional[int] = None) -> torch.Tensor:
    """fp32 channels-last [.., Cs] (first c channels) -> hi|lo pairs bf16 [.., 2*cp]."""
    lib = L.load()
    assert x.dtype == torch.float32 and x.is_contiguous()
    cs = x.shape[-1]
    c = cs if c is None else c
    cp = pad16(c) if cp is None else cp
    rows = x.numel() // cs
    out = torch.empty((*x.shape[:-1], 2 * cp), dtype=torch.bfloat16, device=x.device)
    L.check(lib.fo_split_f32(x.data_ptr(), 1, c, rows, 0, 1, cs, out.data_ptr(), cp, _stream()), "fo_split_f32")
    _count(1)
    return out


def merge_cl(t: torch.Tensor, c: int, c_off: int = 0) -> torch.Tensor:
    """hi|lo pairs bf16 [.., 2*cp] -> fp32 channels-last [.., c] (logical channels c_off .. c_off + c)."""
    lib = L.load()
    assert t.dtype == torch.bfloat16 and t.is_contiguous()
    cp = t.shape[-1] // 2
    rows = t.numel() // t.shape[-1]
    out = torch.empty((*t.shape[:-1], c), dtype=torch.float32, device=t.device)
    L.check(lib.fo_merge_f32(t.data_ptr() + 2 * c_off, 1, c, rows, cp, out.data_ptr(), 0, 1, c, _stream()), "fo_merge_f32")
    _count(1)
    return out


def add_grads(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Sum of two activation gradients of the same layout (rare path: a node with several raw consumers)."""
    if PRECISE:
        c = a.shape[-1] // 2
        return split_cl(merge_cl(a, c) + merge_cl(b, c), c, c)
    return a + b


# ------------------------------------------------------------------------------------------------
# workspaces (per device, grown on demand; all users are ordered on the current stream)
# ------------------------------------------------------------------------------------------------
_ws: Dict[Tuple[int, str], torch.Tensor] = {}


def workspace(nbytes: int, device, tag: str = "main") -> torch.Tensor:
    key = (torch.device(device).index or 0, tag)
    t = _ws.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws[key] = t
    return t


# ------------------------------------------------------------------------------------------------
# convolution family
# ------------------------------------------------------------------------------------------------
SrcT = Tuple[torch.Tensor, int, int]  # (tensor, logical channels, channel offset)

_wcache: Dict[tuple, tuple] = {}


def _fill_src(dst: Src, s: SrcT):
    t, c, c_off = s
    assert t.dtype == torch.bfloat16 and t.is_contiguous() and t.is_cuda, "activations must be contiguous bf16 CUDA"
    dst.ptr, dst.c, dst.cs, dst.c_off = t.data_ptr(), c, t.shape[-1], c_off


def _conv_desc(form: int, ndim: int, ksize: int, srcs: Sequence[SrcT], cout: int) -> ConvDesc:
    d = ConvDesc()
    t0 = srcs[0][0]
    d.form, d.ndim, d.ksize = form, ndim, ksize
    if ndim == 3:
        d.n, d.d, d.h, d.w = t0.shape[0], t0.shape[1], t0.shape[2], t0.shape[3]
    else:
        d.n, d.d, d.h, d.w = t0.shape[0], 1, t0.shape[1], t0.shape[2]
    d.n_src = len(srcs)
    for i, s in enumerate(srcs):
        assert s[0].shape[:-1] == t0.shape[:-1]
        _fill_src(d.src[i], s)
    d.cout = cout
    return d


def out_spatial(form: int, shape: Sequence[int]) -> Tuple[int, ...]:
    """Leading (non-channel) output shape for an input of leading shape ``shape``."""
    if form == FORM_DOWN:
        return (shape[0], shape[1] // 2, shape[2] // 2)
    if form == FORM_UP:
        return (shape[0], shape[1] * 2, shape[2] * 2)
    return tuple(shape)


def packed_weights(desc: ConvDesc, weight, n_axis: int, n_scale: Optional[torch.Tensor], key_extra,
                   wkey=None, cache: bool = True) -> torch.Tensor:
    """bf16 K-step-ordered copy of ``weight``, cached per (parameter storage, version, role).  ``wkey`` =
    (source parameter, tag) must be given when ``weight`` is a temporary derived from a parameter (its own
    data_ptr/version say nothing about staleness); ``weight`` may then be a callable that builds the temporary (only
    called on a miss).  ``cache=False``: the operand is an activation playing the weight role (never cached)."""
    lib = L.load()
    if cache:
        src_t, tag = (weight, None) if wkey is None else wkey
        key = (src_t.data_ptr(), tag, desc.form, desc.ndim, desc.ksize, n_axis, key_extra, _p(n_scale))
        ver = src_t._version
        hit = _wcache.get(key)
        # the entry is only valid for the SAME live tensor object: addresses (and version counters) are recycled by the
        # caching allocator once a parameter dies
        if hit is not None and hit[0] == ver and hit[2]() is src_t:
            return hit[1]
    if callable(weight):
        weight = weight()
    nbytes = lib.fo_conv_wpacked_bytes(C.byref(desc))
    if nbytes == 0:
        raise L.FaceoffB200Error("fo_conv_wpacked_bytes: " + lib.fo_last_error().decode())
    wp = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=weight.device)
    assert weight.dtype == torch.float32 and weight.is_contiguous()
    L.check(lib.fo_conv_pack_weights(C.byref(desc), weight.data_ptr(), weight.shape[0], weight.shape[1], n_axis,
                                     _p(n_scale), wp.data_ptr(), _stream()), "fo_conv_pack_weights")
    _count(1)
    if cache:
        _wcache[key] = (ver, wp, weakref.ref(src_t))
    return wp


def conv(form: int, ndim: int, ksize: int, srcs: Sequence[SrcT], weight: torch.Tensor, n_axis: int, cout: int,
         bias: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
         addend: Optional[torch.Tensor] = None, want_raw: bool = True, want_relu: bool = False,
         f32: Optional[str] = None, relu_f32: bool = False, n_scale: Optional[torch.Tensor] = None,
         out_cs: Optional[int] = None, out_raw: Optional[torch.Tensor] = None, wkey=None,
         precise: Optional[bool] = None, out_f32: Optional[torch.Tensor] = None, f32_accumulate: bool = False,
         cache_weights: bool = True, wpacked: Optional[torch.Tensor] = None):
    """Run one implicit-GEMM convolution.  Returns (raw_bf16|None, relu_bf16|None, f32|None).

    f32: None | 'cl' (channels-last fp32 [.., out_cs]) | 'nchw' (fp32 [N, cout, H, W]).
    precise: None = the module-level verification switch; True = hi|lo split-bf16 arithmetic for THIS call (the
    fp32-accurate GEMMs of the discriminator path, faceoff_b200.mocoganhd).  out_f32 / f32_accumulate: write (or add)
    the fp32 result into an existing tensor (a GEMM whose K dimension is split over several launches).
    wpacked: the weight operand already in the packed layout of this descriptor (``weight`` is then ignored).
    """
    lib = L.load()
    precise = PRECISE if precise is None else precise
    if precise:
        # precise = True / 2: (hi, lo, hi) sources against (w_hi, w_hi, w_lo) weights -- 16 mantissa bits per operand.
        # precise = 3: operands split into three bf16 terms (24 bits), the six products of total order <= 2:
        #   x0 w0 + x0 w1 + x1 w0 + x0 w2 + x1 w1 + x2 w0   (fp32-accurate; one logical source only: 6 tensor maps).
        # The K axis of the weight is the one that is not N.
        three = precise == 3
        k_axis = 1 - n_axis
        w_src = weight
        logical = list(srcs)
        assert not three or len(logical) == 1
        x_sel = (0, 0, 1, 0, 1, 2) if three else (0, 1, 0)
        w_sel = (0, 1, 0, 2, 1, 0) if three else (0, 0, 1)

        def weight():   # built only when the packed-weight cache misses
            ws = (w_src() if callable(w_src) else w_src).detach()
            terms, r = [], ws
            for _ in range(3 if three else 2):
                t_ = r.bfloat16().float()
                terms.append(t_)
                r = r - t_
            parts, o = [], 0
            for (_, c, _) in logical:
                parts += [terms[j].narrow(k_axis, o, c) for j in w_sel]
                o += c
            return torch.cat(parts, k_axis).contiguous()

        srcs3 = []
        for (t, c, off) in srcs:
            part = t.shape[-1] // (3 if three else 2)
            srcs3 += [(t, c, off + j * part) for j in x_sel]
        assert wkey is not None or not callable(w_src), "a lazily built weight needs a cache key"
        wkey = (w_src, "precise") if wkey is None else (wkey[0], ("precise", precise, wkey[1]))
        logical_srcs, srcs = srcs, srcs3
        assert n_scale is None
    d = _conv_desc(form, ndim, ksize, srcs, cout)
    # the packed column order depends on the planner's tiling mode, which depends on the geometry
    key_extra = tuple((s[1], s[0].shape[-1], s[2]) for s in srcs) + (cout, d.n, d.d, d.h, d.w)
    if wpacked is not None:
        need = lib.fo_conv_wpacked_bytes(C.byref(d))
        assert wpacked.dtype == torch.bfloat16 and wpacked.is_contiguous() and wpacked.numel() * 2 == need, \
            (wpacked.shape, need)
        wp = wpacked
    else:
        wp = packed_weights(d, weight, n_axis, n_scale, key_extra, wkey, cache=cache_weights)
    d.wpacked = wp.data_ptr()
    t0 = srcs[0][0]
    lead = out_spatial(form, t0.shape[:-1])
    ocs = out_cs if out_cs is not None else pad16(cout) * (2 if precise and f32 is None else 1)
    dev = t0.device
    raw = relu = of32 = None
    if f32 == "nchw":
        of32 = out_f32 if out_f32 is not None else torch.empty((lead[0], cout, lead[1], lead[2]), dtype=torch.float32, device=dev)
        assert of32.is_contiguous() and of32.numel() == lead[0] * cout * lead[1] * lead[2]
        d.out_f32_nchw = 1
    else:
        if want_raw:
            raw = out_raw if out_raw is not None else torch.empty((*lead, ocs), dtype=torch.bfloat16, device=dev)
        if want_relu:
            relu = torch.empty((*lead, ocs), dtype=torch.bfloat16, device=dev)
        if f32 == "cl":
            of32 = out_f32 if out_f32 is not None else torch.empty((*lead, ocs), dtype=torch.float32, device=dev)
            assert of32.is_contiguous() and of32.numel() == ocs * t0.numel() // t0.shape[-1]
    if precise and (raw is not None or relu is not None or mask is not None or addend is not None):
        assert ocs == 2 * pad16(cout) and f32 is None, "verification mode: bf16 epilogue tensors are hi|lo pairs"
        d.split_out = 1
    elif ocs > pad16(cout):
        # channels beyond the computed ones are never written by the kernel
        for t in (raw, relu):
            if t is not None and t is not out_raw:
                t[..., pad16(cout):].zero_()
    d.out_cs = ocs
    d.bias = _p(bias)
    d.mask = _p(mask)
    d.addend = _p(addend)
    d.out_bf16, d.out_relu, d.out_f32 = _p(raw), _p(relu), _p(of32)
    d.relu_f32 = int(relu_f32)
    d.out_f32_accumulate = int(f32_accumulate)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= pad16(cout), "bias must be fp32 padded to 16"
    for t in (mask, addend):
        if t is not None:
            assert t.dtype == torch.bfloat16 and t.is_contiguous() and tuple(t.shape) == (*lead, ocs), (t.shape, lead, ocs)
    # algorithmic FLOPs: 2 * positions * cin * cout * taps (positions: outputs, or inputs for the scatter form)
    cin = sum(s_[1] for s_ in (logical_srcs if precise else srcs))
    taps = 16 if form in (FORM_DOWN, FORM_UP) else ksize ** ndim
    n_pos = 1
    for v_ in (lead if form != FORM_UP else t0.shape[:-1]):
        n_pos *= int(v_)
    flops = 2.0 * n_pos * cin * cout * taps
    with _Timed("conv_igemm", flops):
        L.check(lib.fo_conv_run(C.byref(d), _stream()), "fo_conv_run")
    if PROFILE is not None:   # same events, second key: per layer class (for the roofline break-down)
        kind = ("conv3d" if ndim == 3 else f"conv{ksize}x{ksize}" if form in (FORM_S1, FORM_S1_DGRAD) else
                "conv4x4s2" if form == FORM_DOWN else "convT4x4s2") + f"_{cin}to{cout}"
        PROFILE.setdefault("conv_igemm/" + kind, []).append(PROFILE["conv_igemm"][-1])
    _count(1)
    return raw, relu, of32


def wgrad(form: int, ndim: int, ksize: int, p: SrcT, q: SrcT, dweight: torch.Tensor, m_axis: int, q_w_off: int = 0,
          accumulate: bool = False, dbias: Optional[torch.Tensor] = None, dbias_accumulate: bool = False,
          q_shift_sign: int = 1, _raw: bool = False):
    """dweight (fp32, PyTorch layout) (+)= sum_pix P (x) Q.  form: FORM_S1 or FORM_DOWN (Q = hi-res side)."""
    if PRECISE and not _raw:
        # P_hi Q_hi + P_lo Q_hi + P_hi Q_lo; the bias gradient (column sums of P) takes hi from the first launch, lo from the second
        hp, hq = p[0].shape[-1] // 2, q[0].shape[-1] // 2
        p_lo, q_lo = (p[0], p[1], p[2] + hp), (q[0], q[1], q[2] + hq)
        wgrad(form, ndim, ksize, p, q, dweight, m_axis, q_w_off, accumulate, dbias, dbias_accumulate, q_shift_sign, True)
        wgrad(form, ndim, ksize, p_lo, q, dweight, m_axis, q_w_off, True, dbias, True, q_shift_sign, True)
        wgrad(form, ndim, ksize, p, q_lo, dweight, m_axis, q_w_off, True, None, False, q_shift_sign, True)
        return
    lib = L.load()
    g = WgradDesc()
    pt = p[0]
    g.form, g.ndim, g.ksize = form, ndim, ksize
    if ndim == 3:
        g.n, g.d, g.h, g.w = pt.shape[0], pt.shape[1], pt.shape[2], pt.shape[3]
    else:
        g.n, g.d, g.h, g.w = pt.shape[0], 1, pt.shape[1], pt.shape[2]
    _fill_src(g.p, p)
    _fill_src(g.q, q)
    assert dweight.dtype == torch.float32 and dweight.is_contiguous()
    g.dweight, g.dimA, g.dimB = dweight.data_ptr(), dweight.shape[0], dweight.shape[1]
    g.m_axis, g.q_w_off, g.accumulate = m_axis, q_w_off, int(accumulate)
    g.q_shift_sign = q_shift_sign
    g.dbias, g.dbias_accumulate = _p(dbias), int(dbias_accumulate)
    need = lib.fo_wgrad_workspace_bytes(C.byref(g))
    if need == 0:
        raise L.FaceoffB200Error("fo_wgrad_workspace_bytes: " + lib.fo_last_error().decode())
    ws = workspace(need, pt.device, "wgrad")
    g.workspace, g.workspace_bytes = ws.data_ptr(), ws.numel()
    taps = 16 if form == FORM_DOWN else ksize ** ndim
    n_pos = 1
    for v_ in pt.shape[:-1]:
        n_pos *= int(v_)
    with _Timed("wgrad_igemm", 2.0 * n_pos * p[1] * q[1] * taps):
        L.check(lib.fo_wgrad_run(C.byref(g), _stream()), "fo_wgrad_run")
    if PROFILE is not None:
        kind = ("wgrad3d" if ndim == 3 else f"wgrad{ksize}x{ksize}" if form == FORM_S1 else "wgrad4x4s2") + f"_{p[1]}x{q[1]}"
        PROFILE.setdefault("wgrad_igemm/" + kind, []).append(PROFILE["wgrad_igemm"][-1])
    _count(2)      # wgrad_igemm + wgrad_finalize (which also reduces the bias partials)


def colsum(x: torch.Tensor, c: int, out: torch.Tensor, c_off: int = 0, accumulate: bool = False, _raw: bool = False):
    """out[:c] (+)= x.reshape(-1, Cs)[:, c_off:c_off+c].sum(0)   (bias gradient)."""
    if PRECISE and not _raw:
        colsum(x, c, out, c_off, accumulate, True)
        colsum(x, c, out, c_off + x.shape[-1] // 2, True, True)
        return
    lib = L.load()
    cs = x.shape[-1]
    rows = x.numel() // cs
    need = lib.fo_colsum_workspace_bytes(cs)
    ws = workspace(need, x.device, "colsum")
    with _Timed("hbm/colsum", 2.0 * rows * cs):
        L.check(lib.fo_colsum(x.data_ptr(), rows, cs, c_off, c, out.data_ptr(), int(accumulate), ws.data_ptr(),
                              ws.numel(), _stream()), "fo_colsum")
    _count(2)


def pack_nchw(x: torch.Tensor, cs: Optional[int] = None, shift: Optional[torch.Tensor] = None,
              scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """NCHW fp32 -> channels-last bf16 [N,H,W,cs] (zero padded channels)."""
    lib = L.load()
    assert x.dtype == torch.float32 and x.dim() == 4
    x = x.contiguous()
    n, c, h, w = x.shape
    if PRECISE:
        if shift is not None:
            x = ((x - shift.view(1, -1, 1, 1)[:, :c]) / scale.view(1, -1, 1, 1)[:, :c]).contiguous()
        cp = pad16(c) if cs is None else cs // 2
        out = torch.empty((n, h, w, 2 * cp), dtype=torch.bfloat16, device=x.device)
        L.check(lib.fo_split_f32(x.data_ptr(), n, c, h * w, c * h * w, h * w, 1, out.data_ptr(), cp, _stream()),
                "fo_split_f32")
        _count(1)
        return out
    cs = cs or pad16(c)
    out = torch.empty((n, h, w, cs), dtype=torch.bfloat16, device=x.device)
    with _Timed("hbm/pack_nchw", n * h * w * (4.0 * c + 2.0 * cs)):
        L.check(lib.fo_pack_nchw(x.data_ptr(), out.data_ptr(), n, c, h * w, cs, _p(shift), _p(scale), _stream()),
                "fo_pack_nchw")
    _count(1)
    return out


def unpack_nchw(x: torch.Tensor, c: int) -> torch.Tensor:
    """channels-last bf16 [N,H,W,cs] -> NCHW fp32 [N,c,H,W]."""
    lib = L.load()
    n, h, w, cs = x.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    if PRECISE:
        L.check(lib.fo_merge_f32(x.data_ptr(), n, c, h * w, cs // 2, out.data_ptr(), c * h * w, h * w, 1, _stream()),
                "fo_merge_f32")
        _count(1)
        return out
    with _Timed("hbm/unpack_nchw", n * h * w * (4.0 * c + 2.0 * cs)):
        L.check(lib.fo_unpack_nchw(x.data_ptr(), out.data_ptr(), n, c, h * w, cs, _stream()), "fo_unpack_nchw")
    _count(1)
    return out


def im2col4x4s2(x: torch.Tensor, c: int) -> torch.Tensor:
    """NCHW fp32 [N, Ca, H, W] (first c <= 8 channels) -> bf16 [N, H/2, W/2, 128], k = (ky*4+kx)*8 + ch."""
    lib = L.load()
    assert x.dtype == torch.float32 and x.dim() == 4 and x.is_contiguous()
    n, ca, h, w = x.shape
    out = torch.empty((n, h // 2, w // 2, 128), dtype=torch.bfloat16, device=x.device)
    with _Timed("hbm/im2col4x4s2", n * h * w * 4.0 * c + out.numel() * 2.0):
        L.check(lib.fo_im2col4x4s2(x.data_ptr(), out.data_ptr(), n, ca, c, h, w, _stream()), "fo_im2col4x4s2")
    _count(1)
    return out


def im2col3x3(x: torch.Tensor, shift: Optional[torch.Tensor] = None, scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """NCHW fp32 [N, c<=3, H, W] -> bf16 [N, H, W, 32], k = (ky*3+kx)*3 + ch (27 valid), optional (x - shift) / scale."""
    lib = L.load()
    assert x.dtype == torch.float32 and x.dim() == 4 and x.is_contiguous()
    n, c, h, w = x.shape
    out = torch.empty((n, h, w, 32), dtype=torch.bfloat16, device=x.device)
    with _Timed("hbm/im2col3x3", n * h * w * 4.0 * c + out.numel() * 2.0):
        L.check(lib.fo_im2col3x3(x.data_ptr(), out.data_ptr(), n, c, h, w, _p(shift), _p(scale), _stream()),
                "fo_im2col3x3")
    _count(1)
    return out


def vgg_first_conv(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, shift: Optional[torch.Tensor] = None,
                   scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """relu(conv3x3((x - shift) / scale, weight) + bias): NCHW fp32 [N, 3, H, W] -> bf16 channels-last [N, H, W, 64]."""
    lib = L.load()
    assert x.dtype == torch.float32 and x.dim() == 4 and x.is_contiguous() and x.shape[1] == 3
    assert tuple(weight.shape) == (64, 3, 3, 3) and weight.is_contiguous() and bias.numel() >= 64
    n, _, h, w = x.shape
    out = torch.empty((n, h, w, 64), dtype=torch.bfloat16, device=x.device)
    with _Timed("hbm/vgg_first_conv", n * h * w * (12.0 + 128.0)):
        L.check(lib.fo_vgg_first_conv(x.data_ptr(), n, h, w, weight.data_ptr(), bias.data_ptr(), _p(shift), _p(scale),
                                      out.data_ptr(), _stream()), "fo_vgg_first_conv")
    _count(1)
    return out


def vgg_first_dgrad(dy: torch.Tensor, weight: torch.Tensor, scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Data gradient of conv3x3(3 -> 64): dy bf16 channels-last [N, H, W, 64] -> fp32 NCHW [N, 3, H, W], divided by
    ``scale`` ([3], the ScalingLayer's) if given."""
    lib = L.load()
    assert dy.dtype == torch.bfloat16 and dy.dim() == 4 and dy.shape[-1] == 64 and dy.is_contiguous()
    assert tuple(weight.shape) == (64, 3, 3, 3) and weight.dtype == torch.float32 and weight.is_contiguous()
    assert scale is None or (scale.dtype == torch.float32 and scale.numel() == 3 and scale.is_contiguous())
    n, h, w, _ = dy.shape
    dx = torch.empty((n, 3, h, w), dtype=torch.float32, device=dy.device)
    with _Timed("hbm/vgg_first_dgrad", n * h * w * (128.0 + 12.0)):
        L.check(lib.fo_vgg_first_dgrad(dy.data_ptr(), n, h, w, weight.data_ptr(), _p(scale), dx.data_ptr(), _stream()),
                "fo_vgg_first_dgrad")
    _count(1)
    return dx


def s2conv(x: torch.Tensor, c: int, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
           mask: Optional[torch.Tensor] = None, addend: Optional[torch.Tensor] = None, relu: bool = False) -> torch.Tensor:
    """Conv2d(c -> 64, 4, stride 2, pad 1) of the fp32 NCHW tensor x [N, Ca, H, W] (first c channels) without an im2col
    matrix -> bf16 channels-last [N, H/2, W/2, 64]; epilogue: + bias, ReLU gate by ``mask``, + ``addend``, optional ReLU."""
    lib = L.load()
    assert x.dtype == torch.float32 and x.dim() == 4 and x.is_contiguous()
    n, ca, h, w = x.shape
    assert tuple(weight.shape) == (64, c, 4, 4) and weight.dtype == torch.float32 and weight.is_contiguous()
    out = torch.empty((n, h // 2, w // 2, 64), dtype=torch.bfloat16, device=x.device)
    for t in (mask, addend):
        assert t is None or (t.dtype == torch.bfloat16 and t.is_contiguous() and tuple(t.shape) == tuple(out.shape))
    nbytes = n * h * w * 4.0 * c + out.numel() * 2.0 * (1 + (mask is not None) + (addend is not None))
    with _Timed("hbm/s2conv", nbytes):
        L.check(lib.fo_s2conv(x.data_ptr(), n, ca, c, h, w, weight.data_ptr(), _p(bias), _p(mask), _p(addend),
                              out.data_ptr(), int(relu), _stream()), "fo_s2conv")
    _count(1)
    return out


def s2wgrad(x: torch.Tensor, c: int, y: torch.Tensor, dweight: torch.Tensor, accumulate: bool = False,
            dbias: Optional[torch.Tensor] = None, dbias_accumulate: bool = False):
    """dweight [64, c, 4, 4] (+)= sum_pixels y (x) im2col4x4s2(x); dbias [64] (+)= sum_pixels y.  x fp32 NCHW, y bf16
    channels-last [N, H/2, W/2, 64]."""
    lib = L.load()
    assert x.dtype == torch.float32 and x.dim() == 4 and x.is_contiguous()
    n, ca, h, w = x.shape
    assert y.dtype == torch.bfloat16 and y.is_contiguous() and tuple(y.shape) == (n, h // 2, w // 2, 64)
    assert tuple(dweight.shape) == (64, c, 4, 4) and dweight.dtype == torch.float32 and dweight.is_contiguous()
    need = lib.fo_s2wgrad_workspace_bytes()
    ws = workspace(need, x.device, "s2wgrad")
    with _Timed("hbm/s2wgrad", n * h * w * 4.0 * c + y.numel() * 2.0):
        L.check(lib.fo_s2wgrad(x.data_ptr(), n, ca, c, h, w, y.data_ptr(), dweight.data_ptr(), int(accumulate), _p(dbias),
                               int(dbias_accumulate), ws.data_ptr(), ws.numel(), _stream()), "fo_s2wgrad")
    _count(2)


def col2im4x4s2(col: torch.Tensor, bias: Optional[torch.Tensor], c: int) -> torch.Tensor:
    """bf16 [N, Hi, Wi, 128] (k = tap*8 + co) -> NCHW fp32 [N, c, 2Hi, 2Wi] (+ bias)."""
    lib = L.load()
    n, hi, wi, k = col.shape
    assert k == 128 and col.dtype == torch.bfloat16 and col.is_contiguous()
    out = torch.empty((n, c, 2 * hi, 2 * wi), dtype=torch.float32, device=col.device)
    with _Timed("hbm/col2im4x4s2", col.numel() * 2.0 + out.numel() * 4.0):
        L.check(lib.fo_col2im4x4s2(col.data_ptr(), _p(bias), out.data_ptr(), n, c, hi, wi, _stream()),
                "fo_col2im4x4s2")
    _count(1)
    return out


def chansum_nchw(x: torch.Tensor, c: int, out: torch.Tensor, accumulate: bool = False):
    """out[:c] (+)= x[:, :c].sum((0, 2, 3)) for NCHW fp32 x."""
    lib = L.load()
    assert x.dtype == torch.float32 and x.is_contiguous()
    n, ca, h, w = x.shape
    with _Timed("hbm/chansum_nchw", 4.0 * n * c * h * w):
        L.check(lib.fo_chansum_nchw(x.data_ptr(), n, ca, c, h * w, out.data_ptr(), int(accumulate), _stream()),
                "fo_chansum_nchw")
    _count(1)


def u8hwc_to_nchw(x: torch.Tensor, out: torch.Tensor, c_off: int = 0, mean: float = 0.5, std: float = 0.5):
    """uint8 frames [N, H, W, 3] -> out[:, c_off:c_off+3] = ((x / 255) - mean) / std, out fp32 NCHW [N, C, H, W]."""
    lib = L.load()
    assert x.dtype == torch.uint8 and x.is_cuda and x.is_contiguous() and x.dim() == 4 and x.shape[-1] == 3
    n, h, w, _ = x.shape
    assert out.dtype == torch.float32 and out.is_contiguous() and out.shape[0] == n and tuple(out.shape[2:]) == (h, w)
    with _Timed("hbm/u8hwc_to_nchw", 15.0 * n * h * w):
        L.check(lib.fo_u8hwc_to_nchw(x.data_ptr(), out.data_ptr(), n, h * w, out.shape[1], c_off, mean, std, _stream()),
                "fo_u8hwc_to_nchw")
    _count(1)
    return out


def relu(x: torch.Tensor) -> torch.Tensor:
    if PRECISE:
        c = x.shape[-1] // 2
        return split_cl(merge_cl(x, c).clamp_min_(0), c, c)
    lib = L.load()
    y = torch.empty_like(x)
    L.check(lib.fo_relu(x.data_ptr(), y.data_ptr(), x.numel(), _stream()), "fo_relu")
    _count(1)
    return y


def maxpool2(x: torch.Tensor) -> torch.Tensor:
    lib = L.load()
    n, h, w, cs = x.shape
    if PRECISE:
        c = cs // 2
        x32 = merge_cl(x, c)
        y32 = torch.empty((n, h // 2, w // 2, c), dtype=torch.float32, device=x.device)
        L.check(lib.fo_maxpool2_f32(x32.data_ptr(), y32.data_ptr(), n, h, w, c, _stream()), "fo_maxpool2_f32")
        _count(1)
        return split_cl(y32, c, c)
    y = torch.empty((n, h // 2, w // 2, cs), dtype=torch.bfloat16, device=x.device)
    with _Timed("hbm/maxpool2", 2.5 * x.numel()):
        L.check(lib.fo_maxpool2(x.data_ptr(), y.data_ptr(), n, h, w, cs, _stream()), "fo_maxpool2")
    _count(1)
    return y


def maxpool2_bwd(x: torch.Tensor, y: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    lib = L.load()
    n, h, w, cs = x.shape
    if PRECISE:
        c = cs // 2
        x32, y32, dy32 = merge_cl(x, c), merge_cl(y, c), merge_cl(dy, c)
        dx32 = torch.empty_like(x32)
        L.check(lib.fo_maxpool2_bwd_f32(x32.data_ptr(), y32.data_ptr(), dy32.data_ptr(), dx32.data_ptr(), n, h, w, c,
                                        _stream()), "fo_maxpool2_bwd_f32")
        _count(1)
        return split_cl(dx32, c, c)
    dx = torch.empty_like(x)
    with _Timed("hbm/maxpool2_bwd", 5.0 * x.numel()):
        L.check(lib.fo_maxpool2_bwd(x.data_ptr(), y.data_ptr(), dy.data_ptr(), dx.data_ptr(), n, h, w, cs, _stream()),
                "fo_maxpool2_bwd")
    _count(1)
    return dx


# ------------------------------------------------------------------------------------------------
# vector quantiser
# ------------------------------------------------------------------------------------------------
def vq_prep(embed: torch.Tensor):
    """embed fp32 [dim, n_embed] -> (e_split bf16 (split codebook [n_embed, 2*dim] + augmented K slice, flat), e_t fp32
    [n_embed, dim], e_norm2 fp32 [n_embed+1])."""
    lib = L.load()
    dim, n_embed = embed.shape
    dev = embed.device
    e_split = torch.empty(lib.fo_vq_split_elems(dim, n_embed), dtype=torch.bfloat16, device=dev)
    e_t = torch.empty((n_embed, dim), dtype=torch.float32, device=dev)
    e_norm2 = torch.empty(n_embed + 1, dtype=torch.float32, device=dev)
    L.check(lib.fo_vq_prep(embed.data_ptr(), dim, n_embed, e_split.data_ptr(), e_t.data_ptr(), e_norm2.data_ptr(),
                           _stream()), "fo_vq_prep")
    _count(2)
    return e_split, e_t, e_norm2


def vq_assign(x: torch.Tensor, e_t: torch.Tensor, e_split: torch.Tensor, e_norm2: torch.Tensor,
              n_flagged: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x fp32 [rows, dim] -> embed_ind int64 [rows]; e_t / e_split / e_norm2 from vq_prep."""
    lib = L.load()
    rows, dim = x.shape
    n_embed = e_t.shape[0]
    assert x.dtype == torch.float32 and x.is_contiguous() and e_t.shape[1] == dim
    ind = torch.empty(rows, dtype=torch.int64, device=x.device)
    need = lib.fo_vq_assign_workspace_bytes(rows, dim)
    ws = workspace(need, x.device, "vq")
    with _Timed("vq_assign", 2.0 * rows * dim * n_embed):
        L.check(lib.fo_vq_assign(x.data_ptr(), rows, dim, n_embed, e_t.data_ptr(), e_split.data_ptr(),
                                 e_norm2.data_ptr(), ind.data_ptr(), _p(n_flagged), ws.data_ptr(), ws.numel(),
                                 _stream()), "fo_vq_assign")
    _count(2)
    return ind


def vq_gather_stats(x: torch.Tensor, ind: torch.Tensor, e_t: torch.Tensor, diff_sum: torch.Tensor,
                    counts: Optional[torch.Tensor], embed_sum: Optional[torch.Tensor], want_f32: bool = True,
                    want_bf16: bool = False):
    lib = L.load()
    rows, dim = x.shape
    n_embed = e_t.shape[0]
    precise_pairs = PRECISE and want_bf16
    if precise_pairs:           # the bf16 copy for the decoder becomes a hi|lo pair tensor made from the fp32 result
        want_f32, want_bf16 = True, False
    q32 = torch.empty_like(x) if want_f32 else None
    q16 = torch.empty((rows, dim), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    # algorithmic bytes per row: read x 4D + index 8, write q 4D (+ 2D for the bf16 copy)
    with _Timed("hbm/vq_gather_stats", rows * (4.0 * dim + 8 + (4.0 * dim if want_f32 else 0) + (2.0 * dim if want_bf16 else 0))):
        need = lib.fo_vq_gather_scratch_bytes(dim, n_embed) if counts is not None else 0
        scratch = workspace(need, x.device, "vq_stats") if need else None
        L.check(lib.fo_vq_gather_stats(x.data_ptr(), ind.data_ptr(), rows, dim, n_embed, e_t.data_ptr(), _p(q32),
                                       _p(q16), diff_sum.data_ptr(), _p(counts), _p(embed_sum), _p(scratch), _stream()),
                "fo_vq_gather_stats")
    _count(1)
    if precise_pairs:
        q16 = split_cl(q32, dim)
    return q32, q16


def vq_ema(embed, cluster_size, embed_avg, counts, embed_sum, decay: float, eps: float):
    lib = L.load()
    dim, n_embed = embed.shape
    L.check(lib.fo_vq_ema(embed.data_ptr(), cluster_size.data_ptr(), embed_avg.data_ptr(), counts.data_ptr(),
                          embed_sum.data_ptr(), dim, n_embed, decay, 1 - decay, eps, _stream()), "fo_vq_ema")
    _count(1)


def vq_backward(g_q: Optional[torch.Tensor], g_c_off: int, g_diff: Optional[torch.Tensor], x: torch.Tensor,
                ind: torch.Tensor, e_t: torch.Tensor, want_f32: bool = True, want_bf16: bool = False):
    lib = L.load()
    rows, dim = x.shape
    n_embed = e_t.shape[0]
    if PRECISE and (want_bf16 or (g_q is not None and g_q.dtype == torch.bfloat16)):
        # pair tensors in and out: merge -> the fp32 kernel -> split
        gq32 = None if g_q is None else (merge_cl(g_q, dim, g_c_off) if g_q.dtype == torch.bfloat16 else g_q)
        g32, _ = vq_backward(gq32, 0 if g_q is None or g_q.dtype == torch.bfloat16 else g_c_off, g_diff, x, ind, e_t,
                             want_f32=True, want_bf16=False)
        return (g32 if want_f32 else None), (split_cl(g32, dim) if want_bf16 else None)
    g32 = torch.empty_like(x) if want_f32 else None
    g16 = torch.empty((rows, dim), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    is_bf16 = int(g_q is not None and g_q.dtype == torch.bfloat16)
    g_cs = g_q.shape[-1] if g_q is not None else dim
    gbytes = 0.0 if g_q is None else (2.0 if is_bf16 else 4.0) * dim
    with _Timed("hbm/vq_backward", rows * (gbytes + 4.0 * dim + 8 + (4.0 * dim if want_f32 else 0) + (2.0 * dim if want_bf16 else 0))):
        L.check(lib.fo_vq_backward(_p(g_q), is_bf16, g_cs, g_c_off, _p(g_diff), x.data_ptr(), ind.data_ptr(),
                                   e_t.data_ptr(), rows, dim, n_embed, _p(g32), _p(g16), _stream()), "fo_vq_backward")
    _count(1)
    return g32, g16


# ------------------------------------------------------------------------------------------------
# LPIPS head
# ------------------------------------------------------------------------------------------------
def lpips_tap(f0: torch.Tensor, f1: torch.Tensor, w: torch.Tensor, out: torch.Tensor):
    """out[n] += mean_hw sum_c w_c (norm(f0) - norm(f1))^2 ; f0, f1 bf16 [N,H,W,C]."""
    lib = L.load()
    n, h, wd, c = f0.shape
    if PRECISE:
        L.check(lib.fo_lpips_tap_split(f0.data_ptr(), f1.data_ptr(), w.data_ptr(), n, h * wd, c // 2, out.data_ptr(),
                                       _stream()), "fo_lpips_tap_split")
        _count(1)
        return
    with _Timed("hbm/lpips_tap", 4.0 * f0.numel()):
        L.check(lib.fo_lpips_tap(f0.data_ptr(), f1.data_ptr(), w.data_ptr(), n, h * wd, c, out.data_ptr(), _stream()),
                "fo_lpips_tap")
    _count(1)


def lpips_tap_pool(f0: torch.Tensor, f1: torch.Tensor, w: torch.Tensor, out: torch.Tensor, pool_f1: bool = False):
    """lpips_tap that also returns maxpool2(f0) (and maxpool2(f1) if ``pool_f1``, else None): the taps feed 2x2 max pools
    (product mode only)."""
    lib = L.load()
    n, h, wd, c = f0.shape
    assert not PRECISE and h % 2 == 0 and wd % 2 == 0 and f0.is_contiguous() and f1.is_contiguous() and f1.shape == f0.shape
    pooled = torch.empty((n, h // 2, wd // 2, c), dtype=f0.dtype, device=f0.device)
    pooled1 = torch.empty_like(pooled) if pool_f1 else None
    with _Timed("hbm/lpips_tap_pool", (4.5 + (0.5 if pool_f1 else 0.0)) * f0.numel()):
        L.check(lib.fo_lpips_tap_pool(f0.data_ptr(), f1.data_ptr(), w.data_ptr(), n, h, wd, c, out.data_ptr(),
                                      pooled.data_ptr(), _p(pooled1), _stream()), "fo_lpips_tap_pool")
    _count(1)
    return pooled, pooled1


def lpips_tap_bwd(f0: torch.Tensor, f1: torch.Tensor, w: torch.Tensor, g: torch.Tensor,
                  addend: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = L.load()
    n, h, wd, c = f0.shape
    d = torch.empty_like(f0)
    if PRECISE:
        L.check(lib.fo_lpips_tap_bwd_split(f0.data_ptr(), f1.data_ptr(), w.data_ptr(), g.data_ptr(), n, h * wd, c // 2,
                                           d.data_ptr(), _p(addend), _stream()), "fo_lpips_tap_bwd_split")
        _count(1)
        return d
    with _Timed("hbm/lpips_tap_bwd", (6.0 + (2.0 if addend is not None else 0.0)) * f0.numel()):
        L.check(lib.fo_lpips_tap_bwd(f0.data_ptr(), f1.data_ptr(), w.data_ptr(), g.data_ptr(), n, h * wd, c,
                                     d.data_ptr(), _p(addend), _stream()), "fo_lpips_tap_bwd")
    _count(1)
    return d


def lpips_tap_bwd_pool(f0: torch.Tensor, f1: torch.Tensor, w: torch.Tensor, g: torch.Tensor,
                       pool_dy: torch.Tensor) -> torch.Tensor:
    """lpips_tap_bwd for a tap that feeds a 2x2 max pool, with maxpool2_bwd(f0, ., pool_dy) folded in (product mode only)."""
    lib = L.load()
    n, h, wd, c = f0.shape
    assert not PRECISE and tuple(pool_dy.shape) == (n, h // 2, wd // 2, c) and pool_dy.is_contiguous() and f0.is_contiguous()
    d = torch.empty_like(f0)
    with _Timed("hbm/lpips_tap_bwd_pool", 6.5 * f0.numel()):
        L.check(lib.fo_lpips_tap_bwd_pool(f0.data_ptr(), f1.data_ptr(), w.data_ptr(), g.data_ptr(), n, h, wd, c, d.data_ptr(),
                                          pool_dy.data_ptr(), _stream()), "fo_lpips_tap_bwd_pool")
    _count(1)
    return d


# ------------------------------------------------------------------------------------------------
# reconstruction loss
# ------------------------------------------------------------------------------------------------
def mse_sum(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """sum((a[:, :C] - b)^2) for fp32 NCHW a [N, Ca, H, W], b [N, C, H, W] -> fp32 scalar tensor [1]."""
    lib = L.load()
    n, ca, h, w = a.shape
    c = b.shape[1]
    out = torch.zeros(1, dtype=torch.float32, device=a.device)
    with _Timed("hbm/mse", 8.0 * b.numel()):
        L.check(lib.fo_mse(a.data_ptr(), b.data_ptr(), n, ca, c, h * w, out.data_ptr(), _stream()), "fo_mse")
    _count(1)
    return out


def mse_grad(a: torch.Tensor, b: torch.Tensor, g: torch.Tensor, scale: float) -> torch.Tensor:
    """g * scale * (a[:, :C] - b) as an fp32 tensor shaped like a (channels >= C zero); g is a device scalar."""
    lib = L.load()
    n, ca, h, w = a.shape
    c = b.shape[1]
    grad = torch.empty_like(a)
    with _Timed("hbm/mse_grad", 8.0 * b.numel() + 4.0 * a.numel()):
        L.check(lib.fo_mse_grad(a.data_ptr(), b.data_ptr(), n, ca, c, h * w, g.data_ptr(), float(scale),
                                grad.data_ptr(), _stream()), "fo_mse_grad")
    _count(1)
    return grad


# ------------------------------------------------------------------------------------------------
# optimizer
# ------------------------------------------------------------------------------------------------
def adam_chunk_elems() -> int:
    return int(L.load().fo_adam_chunk_elems())


def adam_step(table: torch.Tensor, chunks: torch.Tensor, n_chunks: int, lr: float, beta1: float, beta2: float,
              eps: float, weight_decay: float, step: int, grad_scale: float = 1.0):
    """One fused Adam update over every tensor listed in ``table`` (device int64 [n, 5]: param, grad, exp_avg,
    exp_avg_sq pointers and numel) using the chunk map ``chunks`` (device int32 [n_chunks, 2])."""
    lib = L.load()
    L.check(lib.fo_adam_step(table.data_ptr(), chunks.data_ptr(), n_chunks, lr, beta1, beta2, eps, weight_decay, step,
                             grad_scale, _stream()), "fo_adam_step")
    _count(1)
