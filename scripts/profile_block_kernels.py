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

B, L, W = 256, 197, 768
M = B * L
dev = "cuda"
torch.manual_seed(0)
h = ops.split(torch.randn(M, W, device=dev))
x = torch.randn(M, W, device=dev)
w_qkv, w_out = ops.split(torch.randn(3 * W, W, device=dev) * 0.03), ops.split(torch.randn(W, W, device=dev) * 0.03)
w_fc, w_proj = ops.split(torch.randn(4 * W, W, device=dev) * 0.03), ops.split(torch.randn(W, 4 * W, device=dev) * 0.03)
b3, b1, b4 = torch.randn(3 * W, device=dev), torch.randn(W, device=dev), torch.randn(4 * W, device=dev)
g, be = torch.ones(W, device=dev), torch.zeros(W, device=dev)
qkv = torch.empty(2, M, 3 * W, dtype=torch.bfloat16, device=dev)
fc = torch.empty(2, M, 4 * W, dtype=torch.bfloat16, device=dev)


def block():
    ops.gemm(h, w_qkv, bias=b3, out_split=qkv)
    o = ops.vit_attention(qkv, B, L, 12)
    ops.gemm(o, w_out, bias=b1, residual=x, out_f32=x)
    hh = ops.layernorm(x, g, be, want_f32=False, want_split=True)
    ops.gemm(hh, w_fc, bias=b4, act=ops.ACT_QUICKGELU, out_split=fc)
    ops.gemm(fc, w_proj, bias=b1, residual=x, out_f32=x)


block()
torch.cuda.synchronize()
torch.cuda.profiler.start()
block()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
