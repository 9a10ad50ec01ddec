"""Frame-batch assembly (test mode) against the oracle's restatement of the reference loops."""
import numpy as np
import pytest
import torch

from anomalyclip_b200 import data
from oracle import anomalyclip_oracle as oracle


@pytest.mark.parametrize("frames", [1, 15, 16, 511, 512, 513, 1000, 1536, 4097])
@pytest.mark.parametrize("stride", [1, 2])
def test_index_plan_matches_reference_loops(frames, stride):
    idx, seg = data.test_mode_indices(frames, 32, 16, stride)
    ref, seg_ref = oracle.test_mode_frame_indices(frames, 32, 16, stride)
    assert idx.tolist() == ref and seg == seg_ref
    assert len(idx) * stride == data.padded_length(frames, 32, 16, stride)
    assert idx.max() < frames


def test_labels_and_dataset_tuple(tmp_path):
    rng = np.random.default_rng(0)
    frames, ncrops = 700, 1
    feats = rng.standard_normal((frames * ncrops, 512)).astype(np.float32)
    np.save(tmp_path / "Abuse001_x264.npy", feats)
    (tmp_path / "test.txt").write_text("Abuse001_x264.npy 0 699 3\n")
    ds = data.FeatureVideoDataset.from_annotation_file(
        str(tmp_path / "test.txt"), str(tmp_path), num_segments=32, seg_length=16, stride=1,
        ncrops=ncrops, normal_id=7, annotations={"Abuse001_x264": [100, 200, 650, 10_000]})
    x, labels, label, segment_size, path = ds[0]
    assert x.shape == (1, 1024, 512) and segment_size == 2 and label == 3
    assert labels.tolist() == oracle.frame_labels(frames, 0, 3, 7, [100, 200, 650, 10_000])
    ref_idx, _ = oracle.test_mode_frame_indices(frames, 32, 16, 1)
    assert torch.equal(x[0], torch.from_numpy(feats)[ref_idx])
    # default collate with batch_size_test = 1 gives the shape test_step expects
    batch = torch.utils.data.default_collate([ds[0]])
    assert batch[0].shape == (1, 1, 1024, 512) and int(batch[3][0]) == 2


def test_gather_raw_frames_wraps_around():
    frames = torch.arange(600, dtype=torch.uint8).view(600, 1, 1, 1).expand(600, 3, 2, 2)
    batch, seg = data.gather_test_frames(frames, 32, 16)
    assert batch.shape == (1024, 3, 2, 2) and seg == 2
    assert batch[:, 0, 0, 0].tolist() == [(i % 600) % 256 for i in range(1024)]


@pytest.mark.parametrize("h,w", [(240, 320), (360, 640), (480, 856), (224, 224), (120, 160), (300, 225), (225, 400)])
def test_resize_crop_plan_is_bit_exact_with_pillow(h, w):
    """The oracle's integer restatement of Pillow's bicubic resample + torchvision's centre crop
    against Pillow/torchvision themselves (what the reference's dataset runs)."""
    import torchvision.transforms as T
    from PIL import Image
    rng = np.random.default_rng(h * 1000 + w)
    frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = T.Compose([T.Resize(224, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(224)])(
        Image.fromarray(frame))
    ref = np.asarray(ref).transpose(2, 0, 1)
    plan = data.resize_crop_plan(h, w, 224)
    got = oracle.resize_center_crop_u8(frame, plan)
    assert got.shape == (3, 224, 224)
    assert np.array_equal(got, ref)


def test_feature_file_roundtrip(tmp_path):
    """save_features writes the `.npy` layout the feature dataset reads back (frame-major rows,
    crops adjacent), for 1 and 10 crops; the dataset then pads / wraps the frames as in test mode."""
    import torch
    from anomalyclip_b200.data import FeatureVideoDataset, VideoRecord, save_features, test_mode_indices
    torch.manual_seed(0)
    for ncrops, frames in ((1, 700), (10, 37)):
        feats = torch.randn(frames, ncrops, 512)
        name = save_features(str(tmp_path / f"video_{ncrops}"), feats, ncrops=ncrops)
        assert name.endswith(".npy")
        ds = FeatureVideoDataset([VideoRecord(name, 1, frames, 3)], num_segments=32, seg_length=16,
                                 ncrops=ncrops, normal_id=7)
        x, labels, label, segment_size, path = ds[0]
        idx, s = test_mode_indices(frames, 32, 16)
        assert segment_size == s and x.shape == (ncrops, len(idx), 512) and labels.shape[0] == frames
        assert torch.equal(x, feats[torch.from_numpy(idx)].permute(1, 0, 2))
        # the crop-major tensor the dataset returns can be written back unchanged
        again = save_features(str(tmp_path / f"again_{ncrops}"), feats.permute(1, 0, 2), ncrops=ncrops)
        import numpy as np
        assert np.array_equal(np.load(again), np.load(name))
