#!/usr/bin/env python
"""Per-kernel device times of one ViT block (256 frames x 197 tokens) in a given operand mode,
CUDA events, L2 flushed before every timed launch, median of --iters.  With ACLIP_LIB pointing at
another build of the library this is the same-box A/B of a kernel change:
    python scripts/time_block_modes.py --mode=5; ACLIP_LIB=/path/other.so python scripts/time_block_modes.py --mode=5"""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anomalyclip_b200 import ops  # noqa: E402

arg = lambda k, d: next((a.split("=")[1] for a in sys.argv if a.startswith(f"--{k}=")), d)  # noqa: E731
MODE, ITERS = int(arg("mode", "5")), int(arg("iters", "15"))
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

h = ops.layernorm(x, g, be, want_f32=False, want_split=True, out_enc=ENC[P_ATT])
if P_ATT == 4:
    qkv = ops.gemm(h, w_qkv, bias=b3, passes=4, want_split=True, out_enc=2)
    o = ops.vit_attention(qkv, B, L, 12, out_enc=2)
else:
    qkv = ops.gemm(h, w_qkv, bias=b3, passes=2, want_split=True)
    o = ops.vit_attention(qkv, B, L, 12)
hh = ops.layernorm(x, g, be, want_f32=False, want_split=True, out_enc=ENC[P_MLP])
fc = ops.gemm(hh, w_fc, bias=b4, act=ops.ACT_QUICKGELU, passes=P_MLP, want_split=True, out_enc=ENC[P_MLP])

steps = {
    "ln_1": lambda: ops.layernorm(x, g, be, want_f32=False, want_split=True, out_enc=ENC[P_ATT], out_split=h),
    "in_proj": (lambda: ops.gemm(h, w_qkv, bias=b3, passes=4, out_split=qkv, out_enc=2)) if P_ATT == 4 else
               (lambda: ops.gemm(h, w_qkv, bias=b3, passes=2, out_split=qkv)),
    "attention": (lambda: ops.vit_attention(qkv, B, L, 12, out_enc=2)) if P_ATT == 4 else
                 (lambda: ops.vit_attention(qkv, B, L, 12)),
    "out_proj": (lambda: ops.gemm(o, w_out, bias=b1, residual=x, out_f32=x, passes=4)) if P_ATT == 4 else
                (lambda: ops.gemm(o, w_out3, bias=b1, residual=x, out_f32=x, passes=3)),
    "ln_2": lambda: ops.layernorm(x, g, be, want_f32=False, want_split=True, out_enc=ENC[P_MLP], out_split=hh),
    "c_fc": lambda: ops.gemm(hh, w_fc, bias=b4, act=ops.ACT_QUICKGELU, passes=P_MLP, out_split=fc, out_enc=ENC[P_MLP]),
    "c_proj": lambda: ops.gemm(fc, w_proj, bias=b1, residual=x, out_f32=x, passes=P_MLP),
}
out = {}
for name, fn in steps.items():
    for _ in range(3):
        fn()
    ts = []
    for _ in range(ITERS):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    out[name] = round(statistics.median(ts), 1)
out["block_us"] = round(sum(out.values()), 1)
print(json.dumps({"mode": MODE, "lib": os.environ.get("ACLIP_LIB", "current"), "us": out}))
