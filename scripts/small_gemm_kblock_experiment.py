#!/usr/bin/env python
"""Experiment: what bounds the K loop of a SMALL GEMM (M = 512, N = 256, K = 9 216, fp16 operands,
one tile per CTA)?  128 x 64 tiles (16 CTAs; 24 KB and 4 MMAs per 64-wide K atom, KATOMS atoms per
barrier round) against 64 x 32 tiles (64 CTAs, 12 KB per atom, four atoms per round).  Run three
times: ACLIP_GEMM_DEBUG unset / 1 (no MMAs: feed + barriers only) / 2 (one MMA per atom instead of
four), each with ACLIP_PROFILING_EXPERIMENTS=1.  Times are per launch with 20 launches queued back
to back (no host latency in the number).  Record: profiles/r2_small_gemm_round_trip_experiments.txt
(taken when every tile carried one atom per round)."""Experiment: what bounds the K loop of a SMALL GEMM (M = 512, N = 256, K = 9 216, fp16 operands,
one tile per CTA)?  Per K block a CTA moves 24 KB (128 x 64 tile) or 12 KB (64 x 32 tile) and issues
4 MMAs.  Run three times: ACLIP_GEMM_DEBUG unset / 1 (no MMAs: feed only) / 2 (one MMA per K block
instead of four), each with ACLIP_PROFILING_EXPERIMENTS=1.  Times are per launch with 20 launches
queued back to back (no host latency in the number)."""Experiment: what bounds the K loop of a SMALL GEMM (M = 512, N = 256, K = 9 216, fp16 operands,
one tile per CTA)?  128 x 64 tiles (16 CTAs; 24 KB and 4 MMAs per 64-wide K atom, KATOMS atoms per
barrier round) against 64 x 32 tiles (64 CTAs, 12 KB per atom, four atoms per round).  Run three
times: ACLIP_GEMM_DEBUG unset / 1 (no MMAs: feed + barriers only) / 2 (one MMA per atom instead of
four), each with ACLIP_PROFILING_EXPERIMENTS=1.  Times are per launch with 20 launches queued back
to back (no host latency in the number).  Record: profiles/r2_small_gemm_round_trip_experiments.txt
(taken when every tile carried one atom per round)."""
import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anomalyclip_b200 import ops
dev = 'cuda'
M, N, K = 512, 256, 9216
torch.manual_seed(0)
a = ops.encode_f16(torch.randn(M, K, device=dev))
w = ops.encode_f16f8(torch.randn(N, K, device=dev) * 0.02, weight=True)
out = torch.empty(M, N, device=dev)
x = ops.encode_f16(torch.randn(M, 1024, device=dev))


def t(fn, reps=20, n=15):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / reps)
    return statistics.median(ts)


dbg = os.environ.get("ACLIP_GEMM_DEBUG", "0")
for tile, name in ((1, "128x64 tiles (16 CTAs)"), (2, "64x32 tiles (64 CTAs)")):
    for kk in (9216, 2304):
        us = t(lambda: ops.gemm(a, w, passes=4, out_f32=out, kernel=1, tile=tile, K=kk))
        print(f"debug={dbg} linear {name} K={kk}: {us:.1f} us/launch -> {us * 1e3 / (kk // 64):.0f} ns per K block")
    us = t(lambda: ops.gemm(x, w, conv=(1, 32, 16, 1024), passes=4, out_f32=out, kernel=1, tile=tile))
    print(f"debug={dbg} conv3x3 {name} K=9216: {us:.1f} us/launch -> {us * 1e3 / 144:.0f} ns per K block")
