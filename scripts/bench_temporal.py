#!/usr/bin/env python
"""BASELINE.json configs[1]: UCF-Crime-shaped pre-extracted features (T = 32 segments x 16 rows,
14 classes) through selector + temporal transformer + head, sub-video batch sweep.
Prints rows/s, algorithmic TFLOP/s (20.2 MFLOP per row) and the per-kernel split."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anomalyclip_b200 import _lib, synthetic as syn  # noqa: E402
from anomalyclip_b200.engine import PackedTemporal, TemporalScorer  # noqa: E402

MFLOP_PER_ROW = {"ucfcrime": 20.2, "shanghaitech": 40.2, "xdviolence": 5.1}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="ucfcrime")
    ap.add_argument("--batches", default="1,8,64,512,2048")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--passes", type=lambda v: v if v == "auto" else int(v), choices=(2, 3, 4, "auto"), default="auto",
                    help="conv feed-forward operand mode: 3 split-bf16, 2 f16f8 for chunks of >= 8 sub-videos (E = 256), "
                         "4 fp16 one pass, auto = 4 if the calibration accepts it")
    args = ap.parse_args()
    dev = torch.device("cuda")
    cfg = syn.PRESETS[args.preset]
    packed = PackedTemporal(syn.make_state_dict(cfg, with_vit=False), dev, num_classes=cfg.num_classes,
                            normal_id=cfg.normal_id, emb_size=cfg.emb_size, depth=cfg.depth,
                            heads=cfg.heads, num_segments=cfg.num_segments, seg_length=cfg.seg_length,
                            concat_features=cfg.concat_features)
    packed.set_directions(syn.make_text_features(cfg), syn.make_ncentroid(cfg))
    scorer = TemporalScorer(packed, passes=args.passes, max_chunk_sub_videos=1024)
    scorer.packed.set_directions(syn.make_text_features(cfg).to(dev), syn.make_ncentroid(cfg).to(dev))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for B in [int(b) for b in args.batches.split(",")]:
        feats = torch.randn(B * cfg.unit, 512, device=dev) * 0.5
        for _ in range(3):
            scorer(feats, 1)
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(args.iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); scorer(feats, 1); b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        ms = tot / args.iters
        _lib.timing_enable(True)
        scorer(feats, 1)
        torch.cuda.synchronize()
        _lib.timing_enable(False)
        kinds = {k: round(v["ms"], 3) for k, v in _lib.timing_collect().items()}
        rows = B * cfg.unit
        print(json.dumps({"preset": args.preset, "passes": args.passes, "sub_videos": B, "ms": round(ms, 3),
                          "rows_per_s": round(rows / ms * 1e3),
                          "algo_tflops": round(rows * MFLOP_PER_ROW[args.preset] / ms / 1e3, 1),
                          "kernels_ms": kinds}), flush=True)


if __name__ == "__main__":
    main()
