#!/usr/bin/env python
"""BASELINE.json configs[4]: ViT-block microbench, seq_len 197, d 768, batch sweep.

Times aclip_vit_forward with 1 and 3 transformer layers on the same frames and reports the
per-block time (difference / 2), the algorithmic TFLOP/s (2.908 GFLOP per frame per block,
SURVEY 8d) and the tensor-pipe issue rate in bf16-pass equivalents (--passes 2: f16f8 operands, one
fp16 pass + two e4m3 half-passes for in_proj / c_fc / c_proj, three bf16 passes for out_proj and the
attention; --passes 3: three bf16 passes everywhere) against the measured bf16 peak.  CUDA events,
3 warm-ups, L2 flushed between iterations.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anomalyclip_b200.engine import PackedVit, VitEncoder  # noqa: E402
from anomalyclip_b200.synthetic import make_frames_u8, make_vit_weights  # noqa: E402

BLOCK_GFLOP = 2.9079


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="32,64,128,256,512,1024,2048,4096")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--passes", type=int, choices=(2, 3), default=2)
    args = ap.parse_args()
    # bf16-pass equivalents issued per algorithmic flop of one block: in_proj 697 + c_fc 930 + c_proj
    # 930 MFLOP at `passes`, out_proj 232 + attention 119 MFLOP always at 3
    issue = (2557 * args.passes + 351 * 3) / 2908
    dev = torch.device("cuda")
    peak = 1408.1
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops"]
    except Exception:  # noqa: BLE001
        pass
    enc = {n: VitEncoder(PackedVit(make_vit_weights(layers=n), dev, passes=args.passes), micro_batch=1 << 20,
                         passes=args.passes) for n in (1, 3)}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    base = make_frames_u8(64, seed=1).to(dev)
    for B in [int(b) for b in args.batches.split(",")]:
        frames = base.repeat((B + 63) // 64, 1, 1, 1)[:B].contiguous()
        out = torch.empty(B, 512, device=dev)
        t = {}
        for n, e in enc.items():
            for _ in range(3):
                e(frames, out)
            torch.cuda.synchronize()
            tot = 0.0
            for _ in range(args.iters):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); e(frames, out); b.record()
                torch.cuda.synchronize()
                tot += a.elapsed_time(b)
            t[n] = tot / args.iters
        block_ms = (t[3] - t[1]) / 2
        tf = BLOCK_GFLOP * B / block_ms  # GFLOP / ms = TFLOP/s
        print(json.dumps({"batch": B, "passes": args.passes, "block_ms": round(block_ms, 4),
                          "algo_tflops": round(tf, 1), "issued_tflops": round(issue * tf, 1),
                          "issued_frac_of_burst_peak": round(issue * tf / peak, 3),
                          "frames_per_s_12_blocks": round(B / (12 * block_ms) * 1e3)}), flush=True)


if __name__ == "__main__":
    main()
