"""ctypes binding of `libaclip_b200.so` (the C ABI declared in include/aclip_b200.h).

There is no fallback: if the library cannot be loaded, every operator raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
# ACLIP_LIB: load another build of the same library (A/B timing of kernel changes on one box)
LIB_PATH = Path(os.environ.get("ACLIP_LIB") or _PKG / "lib" / "libaclip_b200.so")

ACLIP_OK = 0
ACT_NONE, ACT_QUICKGELU, ACT_LEAKYRELU = 0, 1, 2

vp, fp_ = C.c_void_p, C.c_void_p  # device pointers travel as integers


class AclipError(RuntimeError):
    """A C-ABI call returned a negative AclipStatus."""


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", vp), ("w", vp),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("lda", C.c_int), ("ldw", C.c_int),
        ("a_plane_stride", C.c_longlong), ("w_plane_stride", C.c_longlong),
        ("passes", C.c_int), ("a_mode", C.c_int),
        ("conv_c", C.c_int), ("conv_h", C.c_int), ("conv_w", C.c_int), ("conv_s", C.c_int),
        ("bias", vp), ("residual", vp),
        ("res_mod", C.c_int), ("ldr", C.c_int), ("act", C.c_int),
        ("out_f32", vp), ("out_split", vp),
        ("split_plane_stride", C.c_longlong), ("ldc", C.c_int), ("ld_split", C.c_int),
        ("row_group", C.c_int), ("row_group_stride", C.c_int), ("row_offset", C.c_int),
        ("max_ctas", C.c_int), ("kernel", C.c_int),
        ("out_scale", C.c_float), ("out_enc", C.c_int),
        ("gather", vp), ("gather_signal", C.c_int), ("gather_row0", C.c_longlong),
        ("tile", C.c_int),
    ]


class VitBlock(C.Structure):
    _fields_ = [(n, vp) for n in ("ln1_g", "ln1_b", "ln2_g", "ln2_b", "qkv_w", "qkv_b", "out_w",
                                  "out_b", "fc_w", "fc_b", "proj_w", "proj_b")] + \
               [(n, C.c_float) for n in ("qkv_s", "out_s", "fc_s", "proj_s")] + \
               [("out_w16", vp), ("fc_wmx", vp), ("proj_wmx", vp)]


class VitWeights(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("width", "layers", "heads", "patch", "resolution",
                                       "output_dim")] + \
               [(n, vp) for n in ("conv1_w", "class_embedding", "positional_embedding", "ln_pre_g",
                                  "ln_pre_b", "ln_post_g", "ln_post_b", "proj_w")] + \
               [("blocks", C.POINTER(VitBlock)), ("conv1_s", C.c_float), ("proj_s", C.c_float)]


class AxialAttnWeights(C.Structure):
    _fields_ = [(n, vp) for n in ("norm_g", "norm_b", "qkv_w", "out_w", "out_b")]


class ConvFFWeights(C.Structure):
    _fields_ = [(n, vp) for n in ("g", "b", "conv1_w", "conv1_b", "conv2_w", "conv2_b",
                                  "conv1_w8", "conv2_w8")] + \
               [("conv1_s", C.c_float), ("conv2_s", C.c_float)]


class TemporalWeights(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("feature_dim", "num_dirs", "emb", "depth", "heads",
                                       "num_segments", "seg_length", "concat", "ldf")] + \
               [(n, vp) for n in ("ncentroid", "selector_w", "selector_b", "proj_w", "proj_b", "pos")] + \
               [("attn", C.POINTER(AxialAttnWeights)), ("ff", C.POINTER(ConvFFWeights))] + \
               [(n, vp) for n in ("head_ln_g", "head_ln_b", "head_w")] + \
               [("head_bias", C.c_float)]


class PeerGather(C.Structure):
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("rows_per_rank", C.c_longlong),
                ("width", C.c_int), ("rows", vp * 8), ("flags", vp * 8), ("epoch", C.c_uint),
                ("counter", vp)]


class TimingRow(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_longlong), ("ms", C.c_double),
                ("flops", C.c_double), ("bytes", C.c_double)]


# name -> (restype, argtypes); every symbol include/aclip_b200.h declares
SIGNATURES = {
    "aclip_version": (C.c_int, []),
    "aclip_last_error": (C.c_char_p, []),
    "aclip_launch_count": (C.c_longlong, []),
    "aclip_saturation_count": (C.c_longlong, [C.c_int]),
    "aclip_note_launches": (C.c_longlong, [C.c_longlong]),
    "aclip_timing_enable": (C.c_int, [C.c_int]),
    "aclip_timing_collect": (C.c_int, [C.POINTER(TimingRow), C.c_int]),
    "aclip_split_f32": (C.c_int, [vp, C.c_longlong, C.c_int, C.c_int, vp, C.c_int, C.c_longlong, vp]),
    "aclip_encode_f16f8": (C.c_int, [vp, C.c_longlong, C.c_int, C.c_int, vp, C.c_int, C.c_longlong,
                                     C.c_int, C.c_int, C.c_int, vp]),
    "aclip_f16mx_bytes": (C.c_longlong, [C.c_longlong, C.c_int]),
    "aclip_encode_f16mx": (C.c_int, [vp, C.c_longlong, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp]),
    "aclip_center_regroup": (C.c_int, [vp, C.c_longlong, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp,
                                       C.c_int, C.c_longlong, vp]),
    "aclip_patchify": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                 C.POINTER(C.c_float), vp, C.c_longlong, C.c_int, vp]),
    "aclip_resize_crop_u8": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp,
                                       C.c_int, vp, vp, C.c_int, vp, vp, vp]),
    "aclip_gemm": (C.c_int, [C.POINTER(GemmArgs), vp]),
    "aclip_layernorm": (C.c_int, [vp, C.c_longlong, C.c_int, C.c_longlong, vp, vp, C.c_float,
                                  C.c_int, vp, C.c_longlong, vp, C.c_longlong, C.c_longlong, C.c_int,
                                  vp]),
    "aclip_vit_attention": (C.c_int, [vp, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, vp,
                                      C.c_longlong, C.c_int, C.c_int, C.c_int, vp]),
    "aclip_axial_attention": (C.c_int, [vp, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, vp, C.c_longlong, vp]),
    "aclip_vit_workspace_bytes": (C.c_size_t, [C.POINTER(VitWeights), C.c_int]),
    "aclip_vit_forward": (C.c_int, [C.POINTER(VitWeights), vp, C.c_int, C.c_longlong, C.c_int,
                                    C.POINTER(C.c_float), C.POINTER(C.c_float), vp, vp,
                                    C.c_size_t, C.c_int, vp]),
    "aclip_vit_forward_ex": (C.c_int, [C.POINTER(VitWeights), vp, C.c_int, C.c_longlong, C.c_int,
                                       C.POINTER(C.c_float), C.POINTER(C.c_float), vp, vp,
                                       C.c_size_t, C.c_int, vp, vp]),
    "aclip_temporal_workspace_bytes": (C.c_size_t, [C.POINTER(TemporalWeights), C.c_longlong]),
    "aclip_temporal_forward": (C.c_int, [C.POINTER(TemporalWeights), vp, C.c_longlong, C.c_int, vp,
                                         vp, vp, vp, C.c_size_t, C.c_int, vp]),
    "aclip_temporal_forward_ex": (C.c_int, [C.POINTER(TemporalWeights), vp, C.c_longlong, C.c_int, vp,
                                            vp, vp, vp, C.c_size_t, C.c_int, C.POINTER(PeerGather), vp]),
    "aclip_temporal_core_forward": (C.c_int, [C.POINTER(TemporalWeights), vp, C.c_longlong, C.c_int, vp,
                                              vp, C.c_size_t, C.c_int, vp]),
    "aclip_peer_wait": (C.c_int, [vp, C.c_int, C.c_uint, vp]),
    "aclip_peer_signal": (C.c_int, [vp, vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library, building it with nvcc first if it is not there yet."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        from . import build as _build  # raises if nvcc is unavailable

        _build.build()
    try:
        lib = C.CDLL(str(LIB_PATH))
    except OSError as exc:  # pragma: no cover - depends on the machine
        raise AclipError(
            f"cannot load {LIB_PATH}: {exc}. The sm_100a CUDA library is required; "
            "there is no CPU or PyTorch fallback for the hot path.") from exc
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != ACLIP_OK:
        msg = load().aclip_last_error()
        raise AclipError(f"aclip status {rc}: {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(load().aclip_launch_count())


def saturation_count(reset: bool = False) -> int:
    """Threads that stored an activation beyond the fp16 range of the f16f8 / f16 encodings on the
    current CUDA device since the last reset (synchronises the device).  0 on a healthy run."""
    n = int(load().aclip_saturation_count(int(reset)))
    if n < 0:
        check(n)
    return n


def timing_enable(on: bool) -> None:
    check(load().aclip_timing_enable(int(on)))


def timing_collect() -> dict:
    """{kernel kind: {launches, ms, flops, bytes}} since the last collect (synchronises)."""
    rows = (TimingRow * 16)()
    n = load().aclip_timing_collect(rows, 16)
    if n < 0:
        check(n)
    return {rows[i].name.decode(): {"launches": int(rows[i].launches), "ms": rows[i].ms,
                                    "flops": rows[i].flops, "bytes": rows[i].bytes}
            for i in range(n) if rows[i].launches > 0}
