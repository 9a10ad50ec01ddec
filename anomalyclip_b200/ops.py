"""Python handles on the building-block operators of the C ABI.

torch is used for device memory and streams only; every function here enqueues hand-written
sm_100a kernels through `libaclip_b200.so` and raises if that library is missing.
A "split" tensor is a bf16 tensor of shape [2, rows, ld]: plane 0 = hi, plane 1 = lo.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _lib
from ._lib import ACT_LEAKYRELU, ACT_NONE, ACT_QUICKGELU, GemmArgs  # noqa: F401


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA float32 tensor")
    return t if t.is_contiguous() else t.contiguous()


def split(x: torch.Tensor, ld_out: Optional[int] = None,
          out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [rows, cols] -> split bf16 [2, rows, ld_out] (zero padded columns)."""
    x = _f32c(x, "x")
    rows, cols = x.reshape(-1, x.shape[-1]).shape
    ld = ld_out if ld_out is not None else (cols + 7) // 8 * 8
    if out is None:
        out = torch.empty((2, rows, ld), dtype=torch.bfloat16, device=x.device)
    lib = _lib.load()
    _lib.check(lib.aclip_split_f32(x.data_ptr(), rows, cols, cols, out.data_ptr(), ld,
                                   out.stride(0), _stream()))
    return out


ACT_EXP = (4, 7, 0)   # (e_main, e_res, e_coarse) of f16f8 activations (csrc/split.cuh)
WGT_EXP_RES = 4


class F16F8:
    """An fp32 [rows, ld] tensor in the f16f8 operand encoding (include/aclip_b200.h): one uint8
    buffer of 4 * rows * ld bytes = fp16 plane H | e4m3 plane L | e4m3 plane C.  `exp` is e_main."""

    def __init__(self, rows: int, ld: int, device, exp: int = ACT_EXP[0]) -> None:
        if ld % 16 != 0:
            raise ValueError("f16f8 rows must be a multiple of 16 elements wide")
        self.rows, self.ld, self.exp = rows, ld, exp
        self.buf = torch.empty(4 * rows * ld, dtype=torch.uint8, device=device)

    @property
    def plane_stride(self) -> int:
        return self.rows * self.ld

    def data_ptr(self) -> int:
        return self.buf.data_ptr()

    def planes(self):
        """(H fp16, L e4m3, C e4m3) views [rows, ld]."""
        P = self.plane_stride
        return (self.buf[: 2 * P].view(torch.float16).reshape(self.rows, self.ld),
                self.buf[2 * P: 3 * P].view(torch.float8_e4m3fn).reshape(self.rows, self.ld),
                self.buf[3 * P:].view(torch.float8_e4m3fn).reshape(self.rows, self.ld))

    def decode(self, e_res: int = ACT_EXP[1]) -> torch.Tensor:
        """H + L back to fp64 values (what the GEMM effectively multiplies)."""
        h, l, _ = self.planes()
        return (h.double() + l.to(torch.float32).double() * 2.0 ** -e_res) * 2.0 ** -self.exp


_E2M1 = (0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0)


class F16MX:
    """An fp32 [rows, ld] tensor in the f16mx operand encoding (csrc/mx.cuh): one zero-initialised
    uint8 buffer = fp16 plane H | packed e2m1 planes L4, C4 | chunked UE8M0 scale bytes."""

    def __init__(self, rows: int, ld: int, device, exp: int = ACT_EXP[0]) -> None:
        if ld % 64 != 0:
            raise ValueError("f16mx rows must be a multiple of 64 elements wide")
        self.rows, self.ld, self.exp = rows, ld, exp
        self.row_blocks = (rows + 127) // 128
        self.buf = torch.zeros(3 * rows * ld + (ld // 64) * self.row_blocks * 512, dtype=torch.uint8, device=device)

    @property
    def plane_stride(self) -> int:
        return self.rows * self.ld

    def data_ptr(self) -> int:
        return self.buf.data_ptr()

    def planes(self):
        """(H fp16 [rows, ld], L4 values, C4 values as fp64 [rows, ld] with their block scales applied)."""
        P, rows, ld = self.plane_stride, self.rows, self.ld
        h = self.buf[: 2 * P].view(torch.float16).reshape(rows, ld)
        grid = torch.tensor(_E2M1 + tuple(-v for v in _E2M1), dtype=torch.float64, device=self.buf.device)
        sf = self.buf[3 * P:].reshape(ld // 64, self.row_blocks, 32, 4, 4).double()   # [atom][block][m & 31][m >> 5][byte]
        m = torch.arange(rows, device=self.buf.device)
        out = []
        for plane, byte0 in ((0, 0), (1, 2)):
            q = self.buf[2 * P + plane * (P // 2): 2 * P + (plane + 1) * (P // 2)].reshape(rows, ld // 2)
            nib = torch.stack((q & 15, q >> 4), dim=-1).reshape(rows, ld).long()
            vals = grid[nib]
            s = sf[:, m >> 7, m & 31, (m >> 5) & 3][..., byte0:byte0 + 2]            # [atom][rows][2]
            s = torch.exp2(s - 127.0).permute(1, 0, 2).reshape(rows, ld // 32)       # per 32-value block
            out.append(vals * s.repeat_interleave(32, dim=1))
        return h, out[0], out[1]

    def decode(self) -> torch.Tensor:
        """H + L4 back to fp64 values (what the GEMM effectively multiplies)."""
        h, l4, _ = self.planes()
        return (h.double() + l4) * 2.0 ** -self.exp


def encode_f16mx(x: torch.Tensor, *, weight: bool = False, ld_out: Optional[int] = None) -> F16MX:
    """fp32 [rows, cols] -> f16mx planes; weight=True picks the per-tensor exponent."""
    x = _f32c(x, "x")
    rows, cols = x.reshape(-1, x.shape[-1]).shape
    ld = ld_out if ld_out is not None else (cols + 63) // 64 * 64
    out = F16MX(rows, ld, x.device, weight_exponent(x) if weight else ACT_EXP[0])
    lib = _lib.load()
    assert lib.aclip_f16mx_bytes(rows, ld) == out.buf.numel()
    _lib.check(lib.aclip_encode_f16mx(x.data_ptr(), rows, cols, cols, out.data_ptr(), ld, out.exp, _stream()))
    return out


def weight_exponent(w: torch.Tensor) -> int:
    """e_main of a weight tensor: max|w| * 2^e in (2^14, 2^15]  (fp16 main plane stays finite)."""
    m = float(w.abs().max())
    if not (m > 0 and math.isfinite(m)):
        return 0
    return max(-30, min(30, 15 - math.ceil(math.log2(m))))   # the range aclip_encode_f16f8 accepts


def encode_f16f8(x: torch.Tensor, *, weight: bool = False, ld_out: Optional[int] = None) -> F16F8:
    """fp32 [rows, cols] -> f16f8 planes; weight=True picks the per-tensor exponent."""
    x = _f32c(x, "x")
    rows, cols = x.reshape(-1, x.shape[-1]).shape
    ld = ld_out if ld_out is not None else (cols + 15) // 16 * 16
    if weight:
        e = weight_exponent(x)
        exps = (e, WGT_EXP_RES, e - ACT_EXP[1])
    else:
        exps = ACT_EXP
    out = F16F8(rows, ld, x.device, exps[0])
    lib = _lib.load()
    _lib.check(lib.aclip_encode_f16f8(x.data_ptr(), rows, cols, cols, out.data_ptr(), ld,
                                      out.plane_stride, exps[0], exps[1], exps[2], _stream()))
    return out


def encode_f16(x: torch.Tensor) -> torch.Tensor:
    """fp32 [rows, cols] -> the "f16" activation encoding fp16(x * 2^4) (tests build operands with
    this torch cast; on the path the producing kernels write it)."""
    return (x.to(torch.float32) * 2.0 ** ACT_EXP[0]).to(torch.float16).contiguous()


def decode_f16(h: torch.Tensor) -> torch.Tensor:
    return h.double() * 2.0 ** -ACT_EXP[0]


def gemm(a: torch.Tensor, w: torch.Tensor, *, bias: Optional[torch.Tensor] = None,
         act: int = ACT_NONE, residual: Optional[torch.Tensor] = None, res_mod: int = 0,
         out_f32: Optional[torch.Tensor] = None, out_split: Optional[torch.Tensor] = None,
         want_split: bool = False, passes: int = 3, K: Optional[int] = None,
         conv: Optional[tuple] = None, row_map: Optional[tuple] = None, out_rows: Optional[int] = None,
         max_ctas: int = 0, kernel: int = 0, out_enc: int = 0, tile: int = 0) -> torch.Tensor:
    """out = act(A @ W^T + bias) + residual on the tcgen05 GEMM.

    a: split [2, M, lda] (linear) or split NHWC grid [2, S, H, W, C] with conv=(S, H, W, C).
    w: split [2, N, ldw].
    passes=2: a and w are `F16F8` operands; out_enc=1 with want_split returns an `F16F8` output.
    passes=4: a is an fp16 tensor [M, lda] (`encode_f16`; conv: [S*H*W, C]), w an `F16F8` weight (its
    fp16 plane is read); out_enc=2 with want_split returns an fp16 tensor [rows, N].
    """
    lib = _lib.load()
    g = GemmArgs()
    if passes == 4:
        if not (isinstance(w, F16F8) and isinstance(a, torch.Tensor) and a.dtype == torch.float16):
            raise TypeError("gemm(passes=4) needs an fp16 activation tensor and an F16F8 weight")
        M, N = a.shape[0], w.rows
        Kk = K if K is not None else a.shape[1]
        g.lda, g.ldw = a.stride(0), w.ld
        if conv is not None:
            S, H, Wd, Cc = conv
            if a.shape[0] != S * H * Wd or a.shape[1] != Cc:
                raise ValueError("gemm(passes=4, conv): the fp16 grid must be [S*H*W, C]")
            Kk = 9 * Cc
            g.a_mode, g.conv_s, g.conv_h, g.conv_w, g.conv_c = 1, S, H, Wd, Cc
        g.a_plane_stride, g.w_plane_stride = a.numel(), w.plane_stride
        g.out_scale = 2.0 ** -(ACT_EXP[0] + w.exp)
        dev = a.device
    elif passes == 7:        # f16mx operands (fp16 plane + two MXFP4 cross-term planes)
        if not (isinstance(a, F16MX) and isinstance(w, F16MX)):
            raise TypeError("gemm(passes=7) needs F16MX operands")
        if conv is not None:
            raise ValueError("gemm(passes=7): linear operands only")
        M, N = a.rows, w.rows
        Kk = K if K is not None else a.ld
        g.lda, g.ldw = a.ld, w.ld
        g.a_plane_stride, g.w_plane_stride = a.plane_stride, w.plane_stride
        g.out_scale = 2.0 ** -(a.exp + w.exp)
        dev = a.buf.device
    elif passes in (2, 6):   # 6: f16f8 operands without the weight-residual cross term
        if not (isinstance(a, F16F8) and isinstance(w, F16F8)):
            raise TypeError("gemm(passes=2 / 6) needs F16F8 operands")
        M, N = a.rows, w.rows
        Kk = K if K is not None else a.ld
        g.lda, g.ldw = a.ld, w.ld
        if conv is not None:   # a: rows = S*H*W grid cells of C channels (NHWC)
            S, H, Wd, Cc = conv
            if a.rows != S * H * Wd or a.ld != Cc:
                raise ValueError("gemm(passes=2, conv): the F16F8 grid must be [S*H*W, C]")
            Kk = 9 * Cc
            g.a_mode, g.conv_s, g.conv_h, g.conv_w, g.conv_c = 1, S, H, Wd, Cc
        g.a_plane_stride, g.w_plane_stride = a.plane_stride, w.plane_stride
        g.out_scale = 2.0 ** -(a.exp + w.exp)
        dev = a.buf.device
    else:
        N, dev = w.shape[1], a.device
        g.ldw = w.stride(1)
        g.a_plane_stride, g.w_plane_stride = a.stride(0), w.stride(0)
        if conv is not None:
            S, H, Wd, Cc = conv
            M, Kk = S * H * Wd, 9 * Cc
            g.a_mode, g.conv_s, g.conv_h, g.conv_w, g.conv_c = 1, S, H, Wd, Cc
            g.lda = Cc
        else:
            M = a.shape[1]
            Kk = K if K is not None else a.shape[2]
            g.lda = a.stride(1)
    g.a, g.w = a.data_ptr(), w.data_ptr()
    g.M, g.N, g.K = M, N, Kk
    g.passes = passes
    g.out_enc = out_enc
    g.bias = _ptr(bias)
    g.act = act
    g.res_mod = res_mod
    if residual is not None:
        g.residual, g.ldr = residual.data_ptr(), residual.stride(-2)
    rows = out_rows if out_rows is not None else M
    if out_f32 is None and out_split is None:
        if want_split and out_enc == 3:
            out_split = F16MX(rows, N, dev)
        elif want_split and out_enc == 1:
            out_split = F16F8(rows, N, dev)
        elif want_split and out_enc == 2:
            out_split = torch.empty((rows, N), dtype=torch.float16, device=dev)
        elif want_split:
            out_split = torch.empty((2, rows, N), dtype=torch.bfloat16, device=dev)
        else:
            out_f32 = torch.empty((rows, N), dtype=torch.float32, device=dev)
    if out_f32 is not None:
        g.out_f32, g.ldc = out_f32.data_ptr(), out_f32.stride(-2)
    if isinstance(out_split, (F16F8, F16MX)):
        g.out_split, g.ld_split = out_split.data_ptr(), out_split.ld
        g.split_plane_stride = out_split.plane_stride
    elif out_split is not None and out_split.dtype == torch.float16:
        g.out_split, g.ld_split = out_split.data_ptr(), out_split.stride(0)
        g.split_plane_stride = out_split.numel()
    elif out_split is not None:
        g.out_split, g.ld_split = out_split.data_ptr(), out_split.stride(1)
        g.split_plane_stride = out_split.stride(0)
    if row_map is not None:
        g.row_group, g.row_group_stride, g.row_offset = row_map
    g.max_ctas = max_ctas
    g.kernel = kernel
    g.tile = tile
    _lib.check(lib.aclip_gemm(C.byref(g), _stream()))
    return out_f32 if out_f32 is not None else out_split


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, eps: float = 1e-5,
              chan_mode: bool = False, want_f32: bool = True, want_split: bool = False,
              out_enc: int = 0, out_split=None):
    """Row LayerNorm (chan_mode: axial_attention's ChanLayerNorm, eps added to the std).
    out_enc=1: the split output is an `F16F8` activation tensor; out_enc=2: an fp16 tensor;
    out_enc=3: an `F16MX` tensor (D % 256 == 0, no fp32 output).  out_split: write into this
    tensor of the matching kind instead of allocating one."""
    x = _f32c(x, "x")
    rows, D = x.reshape(-1, x.shape[-1]).shape
    out_f32 = torch.empty((rows, D), dtype=torch.float32, device=x.device) if want_f32 else None
    if want_split and out_split is None:
        out_split = F16F8(rows, D, x.device) if out_enc == 1 else \
            F16MX(rows, D, x.device) if out_enc == 3 else \
            torch.empty((rows, D), dtype=torch.float16, device=x.device) if out_enc == 2 else \
            torch.empty((2, rows, D), dtype=torch.bfloat16, device=x.device)
    plane = 0 if out_split is None else (out_split.plane_stride if out_enc in (1, 3) else
                                         out_split.numel() if out_enc == 2 else out_split.stride(0))
    lib = _lib.load()
    _lib.check(lib.aclip_layernorm(
        x.data_ptr(), rows, D, D, _f32c(gamma, "gamma").data_ptr(), _f32c(beta, "beta").data_ptr(),
        eps, 1 if chan_mode else 0, _ptr(out_f32), D,
        out_split.data_ptr() if out_split is not None else None, D, plane, out_enc, _stream()))
    if want_f32 and want_split:
        return out_f32, out_split
    return out_f32 if want_f32 else out_split


def vit_attention(qkv_split: torch.Tensor, B: int, L: int, heads: int, kernel: int = 0,
                  out_enc: int = 0):
    """qkv_split: [2, B*L, 3*heads*64] -> split [2, B*L, heads*64] (out_enc=1: `F16F8`).
    out_enc=2: qkv is an fp16 tensor [B*L, 3*heads*64] (`encode_f16`), the result fp16 [B*L, heads*64]."""
    W = heads * 64
    lib = _lib.load()
    if out_enc == 2:
        if qkv_split.dtype != torch.float16 or qkv_split.dim() != 2:
            raise TypeError("vit_attention(out_enc=2) needs an fp16 [B*L, 3W] tensor")
        out = torch.empty((B * L, W), dtype=torch.float16, device=qkv_split.device)
        _lib.check(lib.aclip_vit_attention(qkv_split.data_ptr(), qkv_split.numel(),
                                           qkv_split.stride(0), B, L, heads, out.data_ptr(),
                                           out.numel(), W, kernel, 2, _stream()))
        return out
    if out_enc == 1:
        out = F16F8(B * L, W, qkv_split.device)
        plane, ld = out.plane_stride, W
    else:
        out = torch.empty((2, B * L, W), dtype=torch.bfloat16, device=qkv_split.device)
        plane, ld = out.stride(0), out.stride(1)
    _lib.check(lib.aclip_vit_attention(qkv_split.data_ptr(), qkv_split.stride(0),
                                       qkv_split.stride(1), B, L, heads, out.data_ptr(),
                                       plane, ld, kernel, out_enc, _stream()))
    return out


def axial_attention(qkv: torch.Tensor, sub_videos: int, n: int, l: int, heads: int,
                    axis: int) -> torch.Tensor:
    """qkv: fp32 [sub_videos*n*l, 3E] in sub-video order -> split [2, rows, E]."""
    qkv = _f32c(qkv, "qkv")
    E = qkv.shape[1] // 3
    out = torch.empty((2, qkv.shape[0], E), dtype=torch.bfloat16, device=qkv.device)
    lib = _lib.load()
    _lib.check(lib.aclip_axial_attention(qkv.data_ptr(), sub_videos, n, l, E, heads, axis,
                                         out.data_ptr(), out.stride(0), _stream()))
    return out


def center(x: torch.Tensor, centroid: torch.Tensor, regroup: Optional[tuple] = None,
           ld_out: Optional[int] = None) -> torch.Tensor:
    """(x - centroid) as split rows; regroup=(n, s, l) also reorders "(b n s l)" -> "(b s) n l"."""
    x = _f32c(x, "x")
    rows, D = x.shape
    ld = ld_out if ld_out is not None else D
    out = torch.zeros((2, rows, ld), dtype=torch.bfloat16, device=x.device)
    n, s, l = regroup if regroup is not None else (1, 1, 1)
    lib = _lib.load()
    _lib.check(lib.aclip_center_regroup(x.data_ptr(), rows, D, _f32c(centroid, "centroid").data_ptr(),
                                        n, s, l, out.data_ptr(), ld, out.stride(0), _stream()))
    return out
