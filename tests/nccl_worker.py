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
    peer.check()

    # configs[3] in miniature: the FRAMES of a batch of units sharded over the ranks (image encoder
    # by frame, feature rows all-gathered by the projection GEMM's epilogue, temporal stage by unit,
    # score rows gathered by the head kernel) against the same frames scored by this rank alone --
    # bit for bit, over several calls (both buffer parities), units < ranks and units > ranks
    from anomalyclip_b200.distributed import FrameShardedScorer
    from anomalyclip_b200.models import AnomalyCLIP
    small = dict(width=256, layers=2, patch=16, resolution=32, output_dim=512)
    import anomalyclip_b200.models as models
    models.ARCHS["test-tiny"] = dict(resolution=32, patch=16, width=256, layers=2, embed_dim=512,
                                     text_width=512, text_layers=1, text_heads=8, context_length=77,
                                     vocab_size=64)
    net = AnomalyCLIP(arch="test-tiny", classnames=[f"c{i}" for i in range(cfg.num_classes)],
                      emb_size=cfg.emb_size, depth=cfg.depth, heads=cfg.heads, dim_heads=None,
                      num_segments=cfg.num_segments, seg_length=cfg.seg_length,
                      concat_features=cfg.concat_features, normal_id=cfg.normal_id, stride=1,
                      load_from_features=False, ncrops=1, build_text_tower=False, micro_batch=192)
    sd = syn.make_state_dict(cfg, with_vit=False)
    sd.update({"image_encoder." + k: v for k, v in syn.make_vit_weights(seed=11, **small).items()})
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    net.set_text_features(syn.make_text_features(cfg))
    net.to(dev).eval()
    m = syn.make_ncentroid(cfg).to(dev)
    for units in (1, 2, 3):                      # 1 unit: rank 1 only signals; 3: uneven blocks
        total = units * cfg.unit
        sharded = FrameShardedScorer(net, total, dev)
        first, count = sharded.frame_block()
        for it in range(4):
            g = torch.Generator().manual_seed(900 + 10 * units + it)
            frames = torch.randint(0, 256, (total, 3, 32, 32), dtype=torch.uint8, generator=g).to(dev)
            rows = sharded(frames[first:first + count], m).clone()
            _, sc = net(frames.reshape(units, cfg.unit, 3, 32, 32), None, m, 1, True)
            alone = torch.cat((sc[:, None], net.class_probs), 1)
            assert rows.shape == alone.shape, (rows.shape, alone.shape)
            assert torch.equal(rows, alone), f"frame-sharded result differs (units={units}, call {it})"
        sharded.check()

    # (f3) ncentroid as a sharded reduction: each rank streams ITS share of the normal videos through
    # the image encoder, one all-reduce of (sum, count) -- against the mean over all videos on one rank
    from anomalyclip_b200.module import AnomalyCLIPModule
    module = AnomalyCLIPModule(net, num_classes=cfg.num_classes)
    vids = []
    for v in range(5):
        g = torch.Generator().manual_seed(40 + v)
        n_real = 300 + 37 * v                                      # real frames (the rest is padding)
        vids.append((torch.randint(0, 256, (1, cfg.unit, 3, 32, 32), dtype=torch.uint8, generator=g),
                     torch.zeros(1, n_real, dtype=torch.long)))
    mine = vids[rank::world]
    got = module.compute_ncentroid(mine, load_from_features=False)
    feats_all = torch.cat([net.image_encoder(x.reshape(-1, 3, 32, 32)[: l.shape[1]].to(dev)) for x, l in vids])
    want = feats_all.double().mean(0).float()
    err = ((got.double() - want.double()).norm() / want.double().norm()).item()
    assert err < 1e-6, f"sharded ncentroid differs from the single-rank mean: {err:.3e}"

    # a peer that never shows up is REPORTED, not waited for forever and not papered over: rank 1 skips
    # one exchange; rank 0's stream-side wait gives up after its bound (ACLIP_PEER_WAIT_CYCLES, set
    # to ~1 s by the test) and the next check() raises
    from anomalyclip_b200._lib import AclipError
    late = PeerRowGather(cfg.unit, cfg.num_classes, dev)
    if rank == 0:
        one = syn.make_features(cfg, 1, seed=77).reshape(-1, 512).to(dev)
        scorer(one, 1, peer=late)
        late.wait()
        torch.cuda.synchronize()
        assert late.timed_out() == [1]
        try:
            late.check()
        except AclipError as exc:
            assert "rank(s) [1]" in str(exc)
        else:
            raise AssertionError("the timed-out exchange was not reported")
        assert late.timed_out() == []          # marks are cleared once reported
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")


if __name__ == "__main__":
    main()
