#!/usr/bin/env python
"""The two MLP GEMMs of a ViT-B/16 block (256 frames x 197 tokens) on f16mx operands (passes = 7,
1.5 pass-equivalents) against f16f8 (passes = 2, 2 pass-equivalents) and fp16 alone (passes = 4):
CUDA events, L2 flushed before every timed launch, median."""
import os, sys, statistics, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anomalyclip_b200 import ops
dev = "cuda"
M, W = 256 * 197, 768
torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, n=15):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return round(statistics.median(ts), 1)
x = torch.randn(M, W, device=dev)
hdn = torch.randn(M, 4 * W, device=dev)
w_fc, w_pr = torch.randn(4 * W, W, device=dev) * 0.03, torch.randn(W, 4 * W, device=dev) * 0.03
b4, b1 = torch.randn(4 * W, device=dev), torch.randn(W, device=dev)
res = torch.randn(M, W, device=dev)
out = {}
xm, hm = ops.encode_f16mx(x), ops.encode_f16mx(hdn)
wm_fc, wm_pr = ops.encode_f16mx(w_fc, weight=True), ops.encode_f16mx(w_pr, weight=True)
hid_mx = ops.F16MX(M, 4 * W, dev)
out["mx"] = {"c_fc": t(lambda: ops.gemm(xm, wm_fc, bias=b4, act=ops.ACT_QUICKGELU, passes=7, out_split=hid_mx, out_enc=3)),
             "c_proj": t(lambda: ops.gemm(hm, wm_pr, bias=b1, residual=res, out_f32=res, passes=7))}
x8, h8 = ops.encode_f16f8(x), ops.encode_f16f8(hdn)
w8_fc, w8_pr = ops.encode_f16f8(w_fc, weight=True), ops.encode_f16f8(w_pr, weight=True)
hid8 = ops.F16F8(M, 4 * W, dev)
out["f16f8"] = {"c_fc": t(lambda: ops.gemm(x8, w8_fc, bias=b4, act=ops.ACT_QUICKGELU, passes=2, out_split=hid8, out_enc=1)),
                "c_proj": t(lambda: ops.gemm(h8, w8_pr, bias=b1, residual=res, out_f32=res, passes=2))}
x16, h16 = ops.encode_f16(x), ops.encode_f16(hdn)
hid16 = torch.empty(M, 4 * W, dtype=torch.float16, device=dev)
out["fp16"] = {"c_fc": t(lambda: ops.gemm(x16, w8_fc, bias=b4, act=ops.ACT_QUICKGELU, passes=4, out_split=hid16, out_enc=2)),
               "c_proj": t(lambda: ops.gemm(h16, w8_pr, bias=b1, residual=res, out_f32=res, passes=4))}
print(json.dumps({"us_per_launch": out, "shape": "c_fc 50432x3072x768, c_proj 50432x768x3072"}))
