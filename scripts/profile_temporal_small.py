#!/usr/bin/env python
"""Profiling aid: ONE ShanghaiTech-shaped sub-video (512 rows) through selector + temporal + head
with direct launches (no CUDA graph), one warm-up and one pass between cudaProfilerStart/Stop."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anomalyclip_b200 import synthetic as syn  # noqa: E402
from anomalyclip_b200.engine import PackedTemporal, TemporalScorer  # noqa: E402

dev = torch.device("cuda")
name = next((a.split("=")[1] for a in sys.argv if a.startswith("--preset=")), "shanghaitech")
units = int(next((a.split("=")[1] for a in sys.argv if a.startswith("--units=")), "1"))
cfg = syn.PRESETS[name]
packed = PackedTemporal(syn.make_state_dict(cfg, with_vit=False), dev, num_classes=cfg.num_classes,
                        normal_id=cfg.normal_id, emb_size=cfg.emb_size, depth=cfg.depth, heads=cfg.heads,
                        num_segments=cfg.num_segments, seg_length=cfg.seg_length,
                        concat_features=cfg.concat_features)
packed.set_directions(syn.make_text_features(cfg).to(dev), syn.make_ncentroid(cfg).to(dev))
scorer = TemporalScorer(packed, passes=4, graph_max_sub_videos=0)
x = torch.randn(units * cfg.unit, 512, device=dev) * 0.5
scorer(x, 1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
scorer(x, 1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
