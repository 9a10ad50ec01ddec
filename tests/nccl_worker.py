"""torchrun worker of test_two_rank_nccl_sharding_matches_single_rank (not a test module)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anomalyclip_b200 import synthetic as syn  # noqa: E402
from anomalyclip_b200.distributed import run_sharded  # noqa: E402
from anomalyclip_b200.engine import PackedTemporal, TemporalScorer  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    cfg = syn.PRESETS["xdviolence"]
    packed = PackedTemporal(syn.make_state_dict(cfg, with_vit=False), dev, num_classes=cfg.num_classes,
                            normal_id=cfg.normal_id, emb_size=cfg.emb_size, depth=cfg.depth,
                            heads=cfg.heads, num_segments=cfg.num_segments, seg_length=cfg.seg_length,
                            concat_features=cfg.concat_features)
    packed.set_directions(syn.make_text_features(cfg), syn.make_ncentroid(cfg))
    scorer = TemporalScorer(packed)
    units = 5
    feats = syn.make_features(cfg, units, seed=3).reshape(-1, 512).to(dev)

    def compute(start, count):
        sim, s, pr = scorer(feats[start * cfg.unit:(start + count) * cfg.unit].contiguous(), 1)
        return torch.cat((s[:, None], pr), 1)

    rows = run_sharded(units, cfg.unit, compute)
    sim, s, pr = scorer(feats, 1)
    assert torch.equal(rows, torch.cat((s[:, None], pr), 1)), "sharded result differs from single-rank"

    # fused all-gather over NVLink peer memory: the head kernel stores the rows into every rank's
    # buffer; compare with the NCCL all-gather of the same rows over several calls (both parities)
    from anomalyclip_b200.distributed import PeerRowGather, gather_rows
    world = dist.get_world_size()
    per_rank = 2 * cfg.unit
    peer = PeerRowGather(per_rank, cfg.num_classes, dev)
    for it in range(5):
        mine = syn.make_features(cfg, 2, seed=50 + 10 * it + rank).reshape(-1, 512).to(dev)
        sim, s, pr = scorer(mine, 1, peer=peer)
        fused = peer.wait().clone()
        ref = gather_rows(torch.cat((s[:, None], pr), 1), [per_rank] * world)
        assert torch.equal(fused, ref), f"fused peer gather differs from NCCL at call {it}"
    dist.destroy_process_group()
    print("rank", rank, "ok")


if __name__ == "__main__":
    main()
