"""ctypes binding of `libaclip_b200.so` (the C ABI declared in include/aclip_b200.h).

There is no fallback: if the library cannot be loaded the import of any operator raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libaclip_b200.so"

ACLIP_OK = 0
ACT_NONE, ACT_QUICKGELU, ACT_LEAKYRELU = 0, 1, 2


class AclipError(RuntimeError):
    """A C-ABI call returned a negative AclipStatus."""


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("w", C.c_void_p),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("lda", C.c_int), ("ldw", C.c_int),
        ("a_plane_stride", C.c_longlong), ("w_plane_stride", C.c_longlong),
        ("passes", C.c_int), ("a_mode", C.c_int),
        ("conv_c", C.c_int), ("conv_h", C.c_int), ("conv_w", C.c_int), ("conv_s", C.c_int),
        ("bias", C.c_void_p), ("residual", C.c_void_p),
        ("res_mod", C.c_int), ("ldr", C.c_int), ("act", C.c_int),
        ("out_f32", C.c_void_p), ("out_split", C.c_void_p),
        ("split_plane_stride", C.c_longlong), ("ldc", C.c_int),
        ("row_group", C.c_int), ("row_group_stride", C.c_int), ("row_offset", C.c_int),
        ("max_ctas", C.c_int),
    ]


_lib = None


def _declare(lib: C.CDLL) -> None:
    lib.aclip_version.restype = C.c_int
    lib.aclip_last_error.restype = C.c_char_p
    lib.aclip_launch_count.restype = C.c_longlong
    lib.aclip_split_f32.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p,
                                    C.c_int, C.c_longlong, C.c_void_p]
    lib.aclip_gemm.argtypes = [C.POINTER(GemmArgs), C.c_void_p]
    for name in ("aclip_split_f32", "aclip_gemm"):
        getattr(lib, name).restype = C.c_int


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
    _declare(lib)
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != ACLIP_OK:
        msg = load().aclip_last_error()
        raise AclipError(f"aclip status {rc}: {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(load().aclip_launch_count())
