"""The CPU oracle against the golden vectors produced by the reference's own modules
(tests/golden/make_golden.py).  Runs without a GPU."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import anomalyclip_oracle as oracle

GOLD = Path(__file__).resolve().parent / "golden"


def _load(name):
    z = np.load(GOLD / name)
    t = {k: torch.from_numpy(z[k]) for k in z.files}
    w = {k[2:]: v for k, v in t.items() if k.startswith("w.")}
    return t, w


def test_vit_matches_reference_module():
    t, w = _load("vit_small.npz")
    out = oracle.vit_forward(w, t["frames"], heads=2)
    torch.testing.assert_close(out, t["out"], rtol=1e-4, atol=1e-5)


def test_selector_matches_reference_module():
    t, _ = _load("selector.npz")
    out = oracle.selector_forward(t["feats"], t["text"], t["ncentroid"], int(t["normal_id"]),
                                  t["bn_mean"], t["bn_var"])
    torch.testing.assert_close(out, t["out"], rtol=1e-5, atol=1e-6)
    assert out.shape == (48, 13)


def test_head_matches_reference_module():
    t, w = _load("head.npz")
    out = oracle.classification_head(t["x"], w, "")
    torch.testing.assert_close(out, t["out"], rtol=1e-5, atol=1e-6)


def test_temporal_wiring_matches_reference_module():
    """Pins temporal_model.py:42-77 (projection, regrouping, classifier, key names); the axial
    transformer underneath is the restated stand-in on both sides (parity unpinned upstream)."""
    t, w = _load("temporal.npz")
    n, l, s, b, E, in_dim, depth, heads = (int(v) for v in t["cfg"])
    out = oracle.temporal_forward(t["x"], w, s, n, l, depth, heads)
    torch.testing.assert_close(out, t["out"], rtol=1e-4, atol=1e-6)


def test_regroup_is_the_einops_pattern():
    """temporal_model.py:46-53: "(b n s l) d -> b n s l d" then "b n s l d -> (b s) n l d"."""
    from einops import rearrange

    b, n, s, l = 2, 4, 3, 2
    x = torch.arange(b * n * s * l, dtype=torch.float32).unsqueeze(1)
    ref = rearrange(rearrange(x, "(b n s l) d -> b n s l d", n=n, s=s, l=l),
                    "b n s l d -> (b s) n l d")
    mine = x.reshape(b, n, s, l, 1).permute(0, 2, 1, 3, 4).reshape(b * s, n, l, 1)
    assert torch.equal(ref, mine)
    # sub-video 0 of video 0 takes rows [0..l), [s*l .. s*l + l), ...
    assert ref[0, :, :, 0].flatten().tolist()[:4] == [0.0, 1.0, 6.0, 7.0]


def test_full_forward_shapes_and_postprocess():
    torch.manual_seed(0)
    C, normal_id, n, l, s, E, depth, heads = 6, 2, 4, 2, 2, 16, 1, 2
    from tests.util_weights import make_temporal_weights

    w = make_temporal_weights(in_dim=32 + (C - 1), emb=E, depth=depth, heads=heads, n=n, l=l,
                              num_classes=C, seed=3)
    x = torch.randn(1, 1, n * s * l, 32)
    sim, scores = oracle.anomaly_clip_forward(
        w, x, 0.1 * torch.randn(32), torch.randn(C, 32), segment_size=s, normal_id=normal_id,
        num_segments=n, seg_length=l, depth=depth, heads=heads, concat_features=True)
    assert sim.shape == (n * s * l, C - 1) and scores.shape == (n * s * l,)
    probs, sc = oracle.test_step_postprocess(sim, scores, num_labels=13)
    assert probs.shape == (13, C - 1) and sc.shape == (13,)
    assert torch.all((scores > 0) & (scores < 1))
    torch.testing.assert_close(probs.sum(1), sc)


@pytest.mark.parametrize("frames,expect", [(1, 512), (512, 512), (513, 1024), (1500, 1536)])
def test_padded_length(frames, expect):
    assert oracle.padded_length(frames, 32, 16) == expect
