#!/usr/bin/env python
"""Profiling aid for ncu (round 2): the kernels of one ViT block at the bench's micro-batch shape
(256 frames x 197 tokens) in the operand mode given by --mode (7 = mixed with the MLP pair on f16mx operands, the bench
default on the synthetic checkpoint; 5 = mixed; 4 = fp16 one pass; 2 = f16f8), one warm-up pass and one pass between
cudaProfilerStart/Stop.  Launch order inside the profiled range:
  layernorm(ln_1), gemm(in_proj), attention, gemm(out_proj), layernorm(ln_2), gemm(c_fc), gemm(c_proj)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anomalyclip_b200 import ops  # noqa: E402

MODE = int(next((a.split("=")[1] for a in sys.argv if a.startswith("--mode=")), "5"))
P_ATT = 4 if MODE in (5, 7) else MODE
P_MLP = 2 if MODE == 5 else MODE      # 7: f16mx operands for the MLP pair
B, L, W = 256, 197, 768
M = B * L
dev = "cuda"
torch.manual_seed(0)
x = torch.randn(M, W, device=dev)
wq = lambda n, k: ops.encode_f16f8(torch.randn(n, k, device=dev) * 0.03, weight=True)  # noqa: E731
w_qkv, w_out, w_fc, w_proj = wq(3 * W, W), wq(W, W), wq(4 * W, W), wq(W, 4 * W)
w_out3 = ops.split(torch.randn(W, W, device=dev) * 0.03)
b3, b1, b4 = torch.randn(3 * W, device=dev), torch.randn(W, device=dev), torch.randn(4 * W, device=dev)
g, be = torch.ones(W, device=dev), torch.zeros(W, device=dev)
ENC = {4: 2, 2: 1, 7: 3}
if MODE == 7:
    wmx = lambda n, k: ops.encode_f16mx(torch.randn(n, k, device=dev) * 0.03, weight=True)  # noqa: E731
    w_fc, w_proj = wmx(4 * W, W), wmx(W, 4 * W)


def block():
    h = ops.layernorm(x, g, be, want_f32=False, want_split=True, out_enc=ENC[P_ATT])
    if P_ATT == 4:
        qkv = ops.gemm(h, w_qkv, bias=b3, passes=4, want_split=True, out_enc=2)
        o = ops.vit_attention(qkv, B, L, 12, out_enc=2)
        ops.gemm(o, w_out, bias=b1, residual=x, out_f32=x, passes=4)
    else:
        qkv = ops.gemm(h, w_qkv, bias=b3, passes=2, want_split=True)
        o = ops.vit_attention(qkv, B, L, 12)
        ops.gemm(o, w_out3, bias=b1, residual=x, out_f32=x, passes=3)
    hh = ops.layernorm(x, g, be, want_f32=False, want_split=True, out_enc=ENC[P_MLP])
    fc = ops.gemm(hh, w_fc, bias=b4, act=ops.ACT_QUICKGELU, passes=P_MLP, want_split=True, out_enc=ENC[P_MLP])
    ops.gemm(fc, w_proj, bias=b1, residual=x, out_f32=x, passes=P_MLP)


block()
torch.cuda.synchronize()
torch.cuda.profiler.start()
block()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
