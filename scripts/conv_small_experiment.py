#!/usr/bin/env python
"""Experiment: where does the one-sub-video conv2 GEMM (M = 512, N = 256, K = 9 216, fp16 operands)
spend its ~50 us?  conv vs linear addressing of the same shape, fewer CTAs, four times the tiles."""
import sys, os, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anomalyclip_b200 import ops
dev='cuda'
S,H,W,C,N=1,32,16,1024,256
torch.manual_seed(0)
x=torch.randn(S*H*W, C, device=dev)
wt=torch.randn(N, 9*C, device=dev)*0.02
a16=ops.encode_f16(x); w8=ops.encode_f16f8(wt, weight=True)
res=torch.randn(S*H*W, N, device=dev)
alin=ops.encode_f16(torch.randn(S*H*W, 9*C, device=dev))
def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)*1e3)
    return statistics.median(ts)
print('conv2 fp16 (M=512,N=256,K=9216)', t(lambda: ops.gemm(a16, w8, conv=(S,H,W,C), passes=4, residual=res, out_f32=res)))
print('linear fp16 same shape        ', t(lambda: ops.gemm(alin, w8, passes=4, residual=res, out_f32=res)))
for mc in (16, 8, 4):
    print('conv2 max_ctas', mc, t(lambda: ops.gemm(a16, w8, conv=(S,H,W,C), passes=4, residual=res, out_f32=res, max_ctas=mc)))
S=4
x=torch.randn(S*H*W, C, device=dev); a16=ops.encode_f16(x); res=torch.randn(S*H*W, N, device=dev)
print('conv2 fp16 4 sub-videos (64 tiles)', t(lambda: ops.gemm(a16, w8, conv=(S,H,W,C), passes=4, residual=res, out_f32=res)))
