#!/usr/bin/env python
"""Op-level timing of the hot kernels at the ViT-B/16 micro-batch shapes (B frames x 197 tokens).

    python scripts/bench_ops.py [--frames 256] [--iters 20]

CUDA events on the current stream, L2 flushed between iterations, prints one JSON line per op.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anomalyclip_b200 import ops  # noqa: E402


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        total += a.elapsed_time(b)
    return total / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    B, L, W = args.frames, 197, 768
    M = B * L
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.manual_seed(0)
    res = []

    def gemm_case(name, N, K, **kw):
        a = ops.split(torch.randn(M, K, device=dev))
        w = ops.split(torch.randn(N, K, device=dev) * 0.03)
        bias = torch.randn(N, device=dev)
        extra = {}
        if kw.get("residual"):
            x = torch.randn(M, N, device=dev)
            extra = dict(residual=x, out_f32=x)
        elif kw.get("split"):
            extra = dict(out_split=torch.empty(2, M, N, dtype=torch.bfloat16, device=dev))
        else:
            extra = dict(out_f32=torch.empty(M, N, device=dev))
        for kern in (2, 4, 1):
            try:
                ms = timeit(lambda: ops.gemm(a, w, bias=bias, act=kw.get("act", 0), kernel=kern, **extra),
                            args.iters, flush)
            except Exception as exc:  # noqa: BLE001 - e.g. an older library build without this kernel
                print(f"gemm_{name} kernel {kern}: {exc}", file=sys.stderr)
                continue
            fl = 2.0 * M * N * K
            res.append({"op": f"gemm_{name}", "kernel": kern, "ms": round(ms, 4),
                        "algo_tflops": round(fl / ms / 1e9, 1), "issued_tflops": round(3 * fl / ms / 1e9, 1)})
            print(json.dumps(res[-1]), flush=True)
        # f16f8 operands: fp16 main product + e4m3 cross terms (2 pass-equivalents), CTA-pair kernel
        a8 = ops.encode_f16f8(torch.randn(M, K, device=dev))
        w8 = ops.encode_f16f8(torch.randn(N, K, device=dev) * 0.03, weight=True)
        if kw.get("split"):
            extra = dict(out_split=ops.F16F8(M, N, dev), out_enc=1) if kw.get("act") else \
                dict(out_split=torch.empty(2, M, N, dtype=torch.bfloat16, device=dev))
        for kern in (2, 4):
            try:
                ms = timeit(lambda: ops.gemm(a8, w8, bias=bias, act=kw.get("act", 0), passes=2, kernel=kern,
                                             **extra), args.iters, flush)
            except Exception as exc:  # noqa: BLE001
                print(f"gemm_{name} f16f8 kernel {kern}: {exc}", file=sys.stderr)
                continue
            fl = 2.0 * M * N * K
            res.append({"op": f"gemm_{name}", "kernel": f"f16f8/{kern}", "ms": round(ms, 4),
                        "algo_tflops": round(fl / ms / 1e9, 1),
                        "bf16_pass_equiv_tflops": round(2 * fl / ms / 1e9, 1)})
            print(json.dumps(res[-1]), flush=True)

    if not args.only or "gemm" in args.only:
        gemm_case("qkv", 3 * W, W, split=True)
        gemm_case("out", W, W, residual=True)
        gemm_case("fc", 4 * W, W, split=True, act=ops.ACT_QUICKGELU)
        gemm_case("proj", W, 4 * W, residual=True)
    if not args.only or "attn" in args.only:
        qkv = ops.split(torch.randn(M, 3 * W, device=dev))
        kerns = (2, 1, 17, 18, 20, 23) if os.environ.get("ACLIP_PROFILING_EXPERIMENTS") == "1" else (2, 1)
        for kern in kerns:
            ms = timeit(lambda: ops.vit_attention(qkv, B, L, 12, kernel=kern), args.iters, flush)
            fl = 4.0 * B * 12 * L * L * 64
            res.append({"op": "vit_attention", "kernel": kern, "ms": round(ms, 4),
                        "algo_tflops": round(fl / ms / 1e9, 1)})
            print(json.dumps(res[-1]), flush=True)
    if not args.only or "ln" in args.only:
        x = torch.randn(M, W, device=dev)
        g, b = torch.ones(W, device=dev), torch.zeros(W, device=dev)
        ms = timeit(lambda: ops.layernorm(x, g, b, want_f32=False, want_split=True), args.iters, flush)
        res.append({"op": "layernorm_split", "ms": round(ms, 4), "gbs": round(M * W * 8 / ms / 1e6, 1)})
        print(json.dumps(res[-1]), flush=True)


if __name__ == "__main__":
    main()
