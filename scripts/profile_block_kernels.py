#!/usr/bin/env python
"""Profiling aid for ncu: the four ViT-block GEMMs (in_proj, out_proj, c_fc, c_proj), the attention
and one LayerNorm at the bench's micro-batch shape (256 frames x 197 tokens), one warm-up pass and
one pass between cudaProfilerStart/Stop.  Launch order inside the profiled range:
  gemm(qkv), attention, gemm(out_proj), layernorm, gemm(c_fc), gemm(c_proj)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anomalyclip_b200 import ops  # noqa: E402

PASSES = 2 if "--passes=2" in sys.argv else 3   # 2: f16f8 operands (fp16 + e4m3 cross terms)
B, L, W = 256, 197, 768
M = B * L
dev = "cuda"
torch.manual_seed(0)
enc_a = ops.encode_f16f8 if PASSES == 2 else ops.split
enc_w = (lambda t: ops.encode_f16f8(t, weight=True)) if PASSES == 2 else ops.split
ENC = 1 if PASSES == 2 else 0
h = enc_a(torch.randn(M, W, device=dev))
x = torch.randn(M, W, device=dev)
# out_proj keeps bf16 hi/lo operands in every mode (csrc/vit.cu)
w_qkv, w_out = enc_w(torch.randn(3 * W, W, device=dev) * 0.03), ops.split(torch.randn(W, W, device=dev) * 0.03)
w_fc, w_proj = enc_w(torch.randn(4 * W, W, device=dev) * 0.03), enc_w(torch.randn(W, 4 * W, device=dev) * 0.03)
b3, b1, b4 = torch.randn(3 * W, device=dev), torch.randn(W, device=dev), torch.randn(4 * W, device=dev)
g, be = torch.ones(W, device=dev), torch.zeros(W, device=dev)
qkv = torch.empty(2, M, 3 * W, dtype=torch.bfloat16, device=dev)
fc = ops.F16F8(M, 4 * W, dev) if PASSES == 2 else torch.empty(2, M, 4 * W, dtype=torch.bfloat16, device=dev)


def block():
    ops.gemm(h, w_qkv, bias=b3, out_split=qkv, passes=PASSES)
    o = ops.vit_attention(qkv, B, L, 12)
    ops.gemm(o, w_out, bias=b1, residual=x, out_f32=x, passes=3)
    hh = ops.layernorm(x, g, be, want_f32=False, want_split=True, out_enc=ENC)
    ops.gemm(hh, w_fc, bias=b4, act=ops.ACT_QUICKGELU, out_split=fc, passes=PASSES, out_enc=ENC)
    ops.gemm(fc, w_proj, bias=b1, residual=x, out_f32=x, passes=PASSES)


block()
torch.cuda.synchronize()
torch.cuda.profiler.start()
block()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
