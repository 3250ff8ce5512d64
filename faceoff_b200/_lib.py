"""ctypes binding of libfaceoff_b200.so (C ABI in include/faceoff_b200.h).

The product path has NO fallback: if the shared library is missing or no sm_100 GPU is present,
every op raises.  ``build()`` compiles the library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# FACEOFF_B200_LIB selects another build of the same C ABI (A/B timing of kernel variants); default: the in-tree build
_SO = os.environ.get("FACEOFF_B200_LIB") or os.path.join(_HERE, "libfaceoff_b200.so")
_lock = threading.Lock()
_lib = None

FORM_S1, FORM_S1_DGRAD, FORM_DOWN, FORM_UP = 0, 1, 2, 3


class FaceoffB200Error(RuntimeError):
    pass


class Src(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("c", C.c_int), ("cs", C.c_int), ("c_off", C.c_int)]


class ConvDesc(C.Structure):
    _fields_ = [
        ("form", C.c_int), ("ndim", C.c_int), ("ksize", C.c_int),
        ("n", C.c_int), ("d", C.c_int), ("h", C.c_int), ("w", C.c_int),
        ("n_src", C.c_int), ("src", Src * 6), ("cout", C.c_int),
        ("wpacked", C.c_void_p), ("bias", C.c_void_p), ("mask", C.c_void_p), ("addend", C.c_void_p),
        ("out_bf16", C.c_void_p), ("out_relu", C.c_void_p), ("out_f32", C.c_void_p),
        ("out_cs", C.c_int), ("out_f32_nchw", C.c_int), ("relu_f32", C.c_int), ("split_out", C.c_int),
        ("out_f32_accumulate", C.c_int),
    ]


class DConvDesc(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("n", "cin", "id", "ih", "iw", "cout", "od", "oh", "ow", "kd", "kh", "kw", "sd", "sh",
                                       "sw", "pd", "ph", "pw")]


class WgradDesc(C.Structure):
    _fields_ = [
        ("form", C.c_int), ("ndim", C.c_int), ("ksize", C.c_int),
        ("n", C.c_int), ("d", C.c_int), ("h", C.c_int), ("w", C.c_int),
        ("p", Src), ("q", Src), ("dweight", C.c_void_p), ("dimA", C.c_int), ("dimB", C.c_int),
        ("m_axis", C.c_int), ("q_w_off", C.c_int), ("q_shift_sign", C.c_int), ("accumulate", C.c_int),
        ("dbias", C.c_void_p), ("dbias_accumulate", C.c_int),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


def build(verbose: bool = False) -> str:
    """Compile libfaceoff_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise FaceoffB200Error("building libfaceoff_b200.so failed")
    return _SO


_SIGS = {
    "fo_last_error": (C.c_char_p, []),
    "fo_version": (C.c_int, []),
    "fo_init": (C.c_int, []),
    "fo_conv_wpacked_bytes": (C.c_size_t, [C.POINTER(ConvDesc)]),
    "fo_conv_pack_weights": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "fo_conv_run": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    "fo_wgrad_workspace_bytes": (C.c_size_t, [C.POINTER(WgradDesc)]),
    "fo_wgrad_run": (C.c_int, [C.POINTER(WgradDesc), C.c_void_p]),
    "fo_pack_nchw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p]),
    "fo_unpack_nchw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fo_u8hwc_to_nchw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                   C.c_void_p]),
    "fo_relu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "fo_colsum": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                            C.c_size_t, C.c_void_p]),
    "fo_colsum_workspace_bytes": (C.c_size_t, [C.c_int]),
    "fo_im2col4x4s2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fo_im2col3x3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p]),
    "fo_vgg_first_conv": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "fo_vgg_first_dgrad": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fo_s2conv": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "fo_s2wgrad_workspace_bytes": (C.c_size_t, []),
    "fo_s2wgrad": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                             C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "fo_col2im4x4s2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fo_chansum_nchw": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "fo_maxpool2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fo_maxpool2_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p]),
    "fo_vq_split_elems": (C.c_size_t, [C.c_int, C.c_int]),
    "fo_vq_prep": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fo_vq_assign_workspace_bytes": (C.c_size_t, [C.c_size_t, C.c_int]),
    "fo_vq_assign": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "fo_vq_gather_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "fo_vq_gather_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fo_vq_ema": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float,
                            C.c_float, C.c_float, C.c_void_p]),
    "fo_vq_backward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fo_lpips_tap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "fo_lpips_tap_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "fo_lpips_tap_pool": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "fo_lpips_tap_bwd_pool": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "fo_lpips_tap_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p]),
    "fo_lpips_tap_bwd_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "fo_split_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_longlong,
                               C.c_void_p, C.c_int, C.c_void_p]),
    "fo_merge_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_longlong,
                               C.c_longlong, C.c_void_p]),
    "fo_maxpool2_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fo_maxpool2_bwd_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p]),
    "fo_dconv_fwd": (C.c_int, [C.POINTER(DConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fo_dconv_dgrad": (C.c_int, [C.POINTER(DConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fo_dconv_wgrad": (C.c_int, [C.POINTER(DConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fo_dconv_im2col_pairs": (C.c_int, [C.POINTER(DConvDesc), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "fo_dconv_im2col_t": (C.c_int, [C.POINTER(DConvDesc), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fo_dconv_col2im": (C.c_int, [C.POINTER(DConvDesc), C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]),
    "fo_dconv_dbias": (C.c_int, [C.POINTER(DConvDesc), C.c_void_p, C.c_void_p, C.c_void_p]),
    "fo_instnorm_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_float, C.c_float, C.c_int,
                                  C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fo_instnorm_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_float, C.c_int,
                                  C.c_void_p, C.c_void_p]),
    "fo_lrelu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_void_p]),
    "fo_lrelu_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_void_p]),
    "fo_avgpool3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong] + [C.c_int] * 10 + [C.c_void_p]),
    "fo_avgpool3_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong] + [C.c_int] * 10 + [C.c_void_p]),
    "fo_ralsgan": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "fo_ralsgan_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "fo_mse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "fo_adam_chunk_elems": (C.c_int, []),
    "fo_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                               C.c_int, C.c_float, C.c_void_p]),
    "fo_mse_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float,
                              C.c_void_p, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)


_ready = False   # fo_init succeeded in this process: load() is then a plain global read (it runs once per op)


def load(init: bool = True):
    """Load the shared library (and, with ``init``, require an sm_100 device)."""
    global _lib, _ready
    if _ready:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(_SO):
                raise FaceoffB200Error(
                    f"{_SO} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(faceoff_b200 has no CPU/PyTorch fallback)")
            lib = C.CDLL(_SO)
            for name, (res, args) in _SIGS.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    if init:
        rc = _lib.fo_init()
        if rc != 0:
            raise FaceoffB200Error(f"fo_init failed ({rc}): {_lib.fo_last_error().decode()}")
        _ready = True
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = _lib.fo_last_error().decode() if _lib is not None else "library not loaded"
        raise FaceoffB200Error(f"{what} failed ({rc}): {msg}")
