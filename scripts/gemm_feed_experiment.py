#!/usr/bin/env python
"""What bounds the two-pass (f16f8) GEMM?  c_proj at the bench shape (M = 50 432, N = 768, K = 3 072)
with parts of the MMA work switched off by the profiling switches of csrc/gemm.cuh (results are WRONG
by construction; ACLIP_PROFILING_EXPERIMENTS=1 ACLIP_GEMM_DEBUG=<mask>): 0 = normal, 1 = operand
feed only (no MMAs), 2 = fp16 MMAs only, 4 = e4m3 MMAs only.  All four move the same bytes into
shared memory; they differ in how many bytes the tensor core reads back out of it."""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anomalyclip_b200 import ops  # noqa: E402

M, N, K = 256 * 197, 768, 3072
dev = "cuda"
torch.manual_seed(0)
a = ops.encode_f16f8(torch.randn(M, K, device=dev))
w = ops.encode_f16f8(torch.randn(N, K, device=dev) * 0.03, weight=True)
x = torch.randn(M, N, device=dev)
b = torch.randn(N, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
fn = lambda: ops.gemm(a, w, bias=b, residual=x, out_f32=x, passes=2)  # noqa: E731
for _ in range(3):
    fn()
ts = []
for _ in range(15):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(json.dumps({"debug_mask": os.environ.get("ACLIP_GEMM_DEBUG", "0"), "c_proj_f16f8_us": round(statistics.median(ts), 1)}))
