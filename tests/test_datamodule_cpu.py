"""The test-mode data path (reference: src/data/anomaly_clip_datamodule.py, video_dataset.py,
feature_dataset.py, src/utils/augmentations.py) on small synthetic videos written to a temp dir.

Checked three ways: against the oracle's restatement of the reference loops, against the
torchvision/PIL transform the reference composes, and -- when /root/reference is present (the build
container) -- against the reference's own dataset classes loaded from their files."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest
import torch
import torchvision.transforms as T
from PIL import Image

from anomalyclip_b200 import data
from anomalyclip_b200.datamodule import AnomalyCLIPDataModule
from oracle import anomalyclip_oracle as oracle

REF = Path("/root/reference/src/data/components")
MEAN, STD = [0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711]


def _reference_transform():
    """get_augmentations(224, 1) of src/utils/augmentations.py:21-34, composed from torchvision."""
    per_frame = T.Compose([T.Resize(224, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(224),
                           T.ToTensor(), T.Normalize(MEAN, STD)])
    return lambda imgs: torch.stack([per_frame(im) for im in imgs])


def _load_reference(name):
    path = REF / f"{name}.py"
    if not path.exists():
        return None
    spec = importlib.util.spec_from_file_location(f"_ref_{name}", path)
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except ImportError:   # feature_dataset.py pulls in src.utils (Lightning / Hydra), absent here
        return None
    return mod


@pytest.fixture()
def frame_videos(tmp_path):
    """Two videos of PNG frames (lossless, so decoding is exact): 40 frames 60x80 starting at frame
    id 1, 23 frames 90x70 starting at frame id 5; plus annotation files."""
    rng = np.random.default_rng(0)
    spec = {"Abuse/Abuse001_x264": (1, 40, 3, (60, 80)), "Normal/Normal007_x264": (5, 27, 7, (90, 70))}
    for name, (start, end, _, (h, w)) in spec.items():
        d = tmp_path / "frames" / name
        d.mkdir(parents=True)
        for f in range(start, end + 1):
            Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)).save(d / f"{f:06d}.png")
    (tmp_path / "test.txt").write_text("".join(f"{n} {s} {e} {c}\n" for n, (s, e, c, _) in spec.items()))
    (tmp_path / "normal.txt").write_text("Normal/Normal007_x264 5 27 7\n")
    (tmp_path / "temporal.txt").write_text("Abuse001_x264.mp4 Abuse 10 20 33 36\nNormal007_x264.mp4 Normal\n")
    return tmp_path, spec


def test_frame_dataset_matches_the_reference_transform_and_loops(frame_videos):
    root, spec = frame_videos
    kw = dict(root_path=str(root / "frames"), annotationfile_path=str(root / "test.txt"), normal_id=7,
              num_segments=2, frames_per_segment=4, imagefile_template="{:06d}.png", test_mode=True,
              temporal_annotation_file=str(root / "temporal.txt"))
    ds_n = data.FrameVideoDataset(output="normalised", **kw)
    ds_u = data.FrameVideoDataset(output="uint8", **kw)
    ds_r = data.FrameVideoDataset(output="raw", **kw)
    transform = _reference_transform()
    ref_mod = _load_reference("video_dataset")
    ref_ds = ref_mod.VideoFrameDataset(transform=transform, **kw) if ref_mod is not None else None
    assert len(ds_n) == 2
    for i, (name, (start, end, label, (h, w))) in enumerate(spec.items()):
        frames = end - start + 1
        idx, seg = oracle.test_mode_frame_indices(frames, 2, 4, 1)
        imgs = [Image.open(root / "frames" / name / f"{f + start:06d}.png").convert("RGB") for f in idx]
        want = transform(imgs)
        x, labels, lab, segment_size, path = ds_n[i]
        assert x.dtype == torch.float32 and torch.equal(x, want)
        assert lab == label and segment_size == seg and path == str(root / "frames" / name)
        intervals = [10, 20, 33, 36] if label == 3 else []
        assert labels.tolist() == oracle.frame_labels(frames, start, label, 7, intervals)
        # uint8 frames: the same pixels before ToTensor + Normalize (those run on the GPU)
        u = ds_u[i][0]
        assert u.dtype == torch.uint8 and u.shape == (len(idx), 3, 224, 224)
        assert torch.equal((u.float() / 255 - torch.tensor(MEAN).view(1, 3, 1, 1)) / torch.tensor(STD).view(1, 3, 1, 1), want)
        r = ds_r[i][0]
        assert r.shape == (len(idx), h, w, 3) and np.array_equal(r[0].numpy(), np.asarray(imgs[0]))
        if ref_ds is not None:  # the reference's own class on the same files
            rx, rlabels, rlab, rseg, rpath = ref_ds[i]
            assert torch.equal(x, rx) and labels.tolist() == rlabels.tolist()
            assert (rlab, rseg, rpath) == (lab, segment_size, path)


def test_frame_dataset_ncrops_only_shortens_the_labels_like_the_reference(frame_videos):
    """video_dataset.py:323: with ncrops = c the raw-frame dataset returns labels for num_frames // c
    frames and the same images."""
    root, spec = frame_videos
    kw = dict(root_path=str(root / "frames"), annotationfile_path=str(root / "test.txt"), normal_id=7,
              num_segments=2, frames_per_segment=4, imagefile_template="{:06d}.png", test_mode=True,
              temporal_annotation_file=str(root / "temporal.txt"))
    one, two = data.FrameVideoDataset(ncrops=1, **kw), data.FrameVideoDataset(ncrops=2, **kw)
    ref_mod = _load_reference("video_dataset")
    ref = ref_mod.VideoFrameDataset(transform=_reference_transform(), ncrops=2, **kw) if ref_mod is not None else None
    for i, (name, (start, end, label, _)) in enumerate(spec.items()):
        frames = end - start + 1
        x1, l1, *_ = one[i]
        x2, l2, lab, seg, path = two[i]
        assert torch.equal(x1, x2) and len(l2) == frames // 2 and l2.tolist() == l1.tolist()[: frames // 2]
        if ref is not None:
            rx, rl, rlab, rseg, rpath = ref[i]
            assert rl.tolist() == l2.tolist() and (rlab, rseg, rpath) == (lab, seg, path)
    with pytest.raises(ValueError):
        data.FrameVideoDataset(ncrops=0, **kw)


def test_frame_dataset_refuses_what_is_out_of_scope(frame_videos):
    root, _ = frame_videos
    kw = dict(root_path=str(root / "frames"), annotationfile_path=str(root / "test.txt"), normal_id=7)
    with pytest.raises(NotImplementedError):
        data.FrameVideoDataset(test_mode=False, **kw)
    with pytest.raises(ValueError):
        data.FrameVideoDataset(output="jpeg", **kw)


def test_datamodule_feature_mode(tmp_path):
    rng = np.random.default_rng(1)
    (tmp_path / "feats" / "Abuse").mkdir(parents=True)
    f1 = rng.standard_normal((700, 512)).astype(np.float32)
    f2 = rng.standard_normal((100, 512)).astype(np.float32)
    np.save(tmp_path / "feats" / "Abuse" / "Abuse001_x264.npy", f1)
    np.save(tmp_path / "feats" / "Normal_100.npy", f2)
    (tmp_path / "test.txt").write_text("Abuse/Abuse001_x264 0 699 3\nNormal_100 0 99 7\n")   # no suffix, as in the reference
    (tmp_path / "normal.txt").write_text("Normal_100 0 99 7\n")
    (tmp_path / "temporal.txt").write_text("Abuse001_x264.mp4 Abuse 100 200\nNormal_100.mp4 Normal\n")
    dm = AnomalyCLIPDataModule(
        num_workers=0, pin_memory=False, num_segments=32, seg_length=16, batch_size=64, batch_size_test=1,
        num_classes=14, input_size=224, load_from_features=True, frames_root=str(tmp_path / "feats"),
        normal_id=7, image_tmpl="{:06d}.jpg", stride=1, ncrops=1,
        annotation_file_normal=str(tmp_path / "normal.txt"), annotation_file_test=str(tmp_path / "test.txt"),
        annotation_file_temporal_test=str(tmp_path / "temporal.txt"), labels_file=None, visualize=False)
    assert dm.num_classes == 14
    with pytest.raises(RuntimeError):
        dm.test_dataloader()
    dm.setup("test")
    batches = list(dm.test_dataloader())
    assert len(batches) == 2 and len(list(dm.train_dataloader_test_mode())) == 1
    x, labels, label, segment_size, path = batches[0]
    assert x.shape == (1, 1, 1024, 512) and int(segment_size[0]) == 2 and int(label[0]) == 3
    assert labels.shape == (1, 700) and labels[0, 100:201].eq(3).all() and int(labels.sum()) == 3 * 101 + 7 * 599
    assert path[0].endswith("Abuse001_x264.npy")
    idx, _ = oracle.test_mode_frame_indices(700, 32, 16, 1)
    assert torch.equal(x[0, 0], torch.from_numpy(f1)[idx])
    with pytest.raises(NotImplementedError):
        dm.train_dataloader()
    ref_mod = _load_reference("feature_dataset")
    if ref_mod is not None:   # the reference's own feature dataset on the same files
        ref = ref_mod.VideoFrameDataset(root_path=str(tmp_path / "feats"), annotationfile_path=str(tmp_path / "test.txt"),
                                        normal_id=7, num_segments=32, frames_per_segment=16, test_mode=True,
                                        ncrops=1, temporal_annotation_file=str(tmp_path / "temporal.txt"))
        for i in range(2):
            rx, rlabels, rlab, rseg, rpath = ref[i]
            ox, olabels, olab, oseg, opath = dm.test_data[i]
            assert torch.equal(ox, rx) and olabels.tolist() == rlabels.tolist()
            assert (olab, oseg, opath) == (rlab, rseg, rpath)


def test_datamodule_frame_mode_and_target_path(frame_videos):
    import importlib
    root, spec = frame_videos
    cls = getattr(importlib.import_module("src.data.anomaly_clip_datamodule"), "AnomalyCLIPDataModule")
    assert cls is AnomalyCLIPDataModule     # configs/data/*.yaml:1 `_target_`
    dm = cls(num_segments=2, seg_length=4, batch_size_test=1, num_classes=14, input_size=224,
             load_from_features=False, frames_root=str(root / "frames"), normal_id=7,
             image_tmpl="{:06d}.png", stride=1, ncrops=1, annotation_file_normal=str(root / "normal.txt"),
             annotation_file_test=str(root / "test.txt"),
             annotation_file_temporal_test=str(root / "temporal.txt"))
    dm.setup()
    frames, labels, label, segment_size, path = next(iter(dm.test_dataloader()))
    assert frames.dtype == torch.uint8 and frames.shape == (1, 40, 3, 224, 224)   # 40 frames = 5 x (2*4)
    assert int(segment_size[0]) == 5 and labels.shape == (1, 40)
    normal = next(iter(dm.train_dataloader_test_mode()))
    assert normal[0].shape == (1, 24, 3, 224, 224) and int(normal[2][0]) == 7      # 23 frames padded to 24


def _write_feature_videos(root: Path):
    rng = np.random.default_rng(2)
    (root / "feats").mkdir(exist_ok=True)
    vids = {"Abuse001": (600, 3, [50, 300]), "Fight002": (90, 5, [10, 60]), "Normal003": (130, 7, []),
            "Normal004": (75, 7, [])}
    feats = {}
    for name, (frames, label, iv) in vids.items():
        f = rng.standard_normal((frames, 512)).astype(np.float32)
        if iv:
            f[iv[0]:iv[1] + 1, :8] += 3.0 * label        # make the anomalous frames separable
        feats[name] = f
        np.save(root / "feats" / f"{name}.npy", f)
    (root / "test.txt").write_text("".join(f"{n} 0 {v[0] - 1} {v[1]}\n" for n, v in vids.items()))
    (root / "normal.txt").write_text("Normal003 0 129 7\nNormal004 0 74 7\n")
    (root / "temporal.txt").write_text("".join(
        f"{n}.mp4 x {' '.join(map(str, v[2]))}\n" for n, v in vids.items()))
    return vids, feats


def _stub_module_and_datamodule(root: Path):
    """A small CPU stand-in with the AnomalyCLIP call signature: the tests below check the host
    flow (datasets, centroid, trimming, sharding, metrics), not the kernels."""
    from torch import nn
    from anomalyclip_b200.module import AnomalyCLIPModule

    class Net(nn.Module):
        embedding_dim, normal_id = 512, 7

        def __init__(self):
            super().__init__()
            self.temporal_model = nn.Linear(1, 1)      # gives the module its device
            self.class_probs = None
            self.calls = []

        def forward(self, x, labels, ncentroid, segment_size, test_mode):
            assert test_mode and x.shape[-2] == 512 * segment_size
            self.calls.append((tuple(x.shape), int(labels.shape[0]), segment_size))
            z = x.reshape(-1, 512) - ncentroid.to(x.dtype)
            scores = torch.sigmoid(z[:, :8].mean(1))
            sim = z[:, :13]
            self.class_probs = torch.softmax(sim, 1) * scores[:, None]
            return sim, scores

    dm = AnomalyCLIPDataModule(num_segments=32, seg_length=16, batch_size_test=1, num_classes=14,
                               load_from_features=True, frames_root=str(root / "feats"), normal_id=7,
                               annotation_file_normal=str(root / "normal.txt"),
                               annotation_file_test=str(root / "test.txt"),
                               annotation_file_temporal_test=str(root / "temporal.txt"))
    net = Net()
    return AnomalyCLIPModule(net, num_classes=14, save_dir=str(root / "out")), dm, net


def _direct_metrics(vids, feats, centroid):
    from anomalyclip_b200.metrics import frame_metrics
    scores, probs, labels = [], [], []
    for name, (frames, label, iv) in vids.items():
        z = torch.from_numpy(feats[name]) - centroid.float()
        s = torch.sigmoid(z[:, :8].mean(1))
        scores.append(s); probs.append(torch.softmax(z[:, :13], 1) * s[:, None])
        labels.append(torch.from_numpy(data.frame_labels(frames, 0, label, 7, iv)))
    return frame_metrics(torch.cat(scores), torch.cat(probs), torch.cat(labels), 7)


def test_evaluate_runs_the_reference_test_sequence(tmp_path):
    """`evaluate(module, datamodule)` = datamodule.setup -> ncentroid from the normal training videos
    -> one test_step per video (padded rows trimmed to the real frames) -> test_epoch_end metrics."""
    from anomalyclip_b200.eval import evaluate
    vids, feats = _write_feature_videos(tmp_path)
    module, dm, net = _stub_module_and_datamodule(tmp_path)
    metrics = evaluate(module, dm)
    normal = torch.from_numpy(np.concatenate([feats["Normal003"], feats["Normal004"]])).double()
    assert torch.allclose(module.ncentroid.double().cpu(), normal.mean(0), atol=1e-6)
    assert [c[1:] for c in net.calls] == [(600, 2), (90, 1), (130, 1), (75, 1)]   # real frames, segment_size
    ref = _direct_metrics(vids, feats, module.ncentroid.cpu())
    assert set(metrics) == {f"test/{k}" for k in ref}
    for k, v in ref.items():
        assert abs(metrics[f"test/{k}"] - v) < 1e-6, k
    assert metrics["test/AUC"] > 0.9 and (tmp_path / "out" / "metrics.json").is_file()


def _eval_worker(rank, world, port, root, q):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from anomalyclip_b200.eval import evaluate
    module, dm, net = _stub_module_and_datamodule(Path(root))
    metrics = evaluate(module, dm)
    q.put((rank, metrics, module.ncentroid.double().tolist(), [c[1] for c in net.calls]))
    dist.destroy_process_group()


def test_evaluate_shards_the_videos_over_ranks_gloo(tmp_path):
    """world_size 2 on CPU: videos are dealt round-robin, the centroid is the all-reduced mean, and both
    ranks end up with the metrics of the whole test set, identical to the single-process run."""
    import socket
    import torch.multiprocessing as mp
    from anomalyclip_b200.eval import evaluate
    vids, feats = _write_feature_videos(tmp_path)
    module, dm, _ = _stub_module_and_datamodule(tmp_path)
    single = evaluate(module, dm)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_eval_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {r: (m, c, calls) for r, m, c, calls in (q.get(timeout=180) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # rank 0 saw videos 0 and 2 (the two test_step calls after its one centroid video), rank 1 videos 1 and 3
    assert got[0][2] == [600, 130] and got[1][2] == [90, 75]
    for r in (0, 1):
        assert torch.allclose(torch.tensor(got[r][1], dtype=torch.float64), module.ncentroid.double().cpu(), atol=1e-6)
        assert set(got[r][0]) == set(single)
        for k, v in single.items():
            assert abs(got[r][0][k] - v) < 1e-6, (r, k)


@pytest.mark.skipif(not Path("/root/reference/configs").is_dir(),
                    reason="reads the reference's own YAML files (present in the build container)")
@pytest.mark.parametrize("data_name,model_name,classes,emb,depth,concat", [
    ("ucfcrime", "anomaly_clip_ucfcrime", 14, 256, 1, False),
    ("shanghaitech", "anomaly_clip_shanghaitech", None, 256, 2, True),
    ("xdviolence", "anomaly_clip_xdviolence", 7, 128, 1, False)])
def test_reference_yaml_configs_instantiate_the_b200_classes(tmp_path, data_name, model_name, classes, emb,
                                                             depth, concat):
    """The reference's config files, composed without Hydra, build this repository's datamodule and
    module: every key is accepted verbatim, `${data.x}` interpolations resolve, partials stay partials."""
    import functools
    from anomalyclip_b200.config import instantiate, load_eval_config
    from anomalyclip_b200.models import AnomalyCLIP
    from anomalyclip_b200.module import AnomalyCLIPModule
    from anomalyclip_b200.training_stubs import ComputeLoss, WarmupCosineAnnealingLR
    labels = {"ucfcrime": "ucf_labels.csv", "shanghaitech": "sht_labels.csv", "xdviolence": "xd_labels.csv"}
    cfg = load_eval_config("/root/reference/configs", data=data_name, model=model_name, overrides=[
        f"data.labels_file=/root/reference/data/{labels[data_name]}", f"data.frames_root={tmp_path}",
        "data.num_workers=0", "model.net.build_text_tower=false"])
    assert cfg["model"]["net"]["normal_id"] == cfg["data"]["normal_id"]          # ${data.normal_id}
    assert cfg["model"]["net"]["labels_file"].endswith(labels[data_name])
    dm = instantiate(cfg["data"])
    assert isinstance(dm, AnomalyCLIPDataModule) and dm.hparams.num_segments == 32
    module = instantiate(cfg["model"])
    assert isinstance(module, AnomalyCLIPModule) and isinstance(module.net, AnomalyCLIP)
    assert isinstance(module.criterion, ComputeLoss)
    assert isinstance(module.optimizer, functools.partial) and module.optimizer.func is torch.optim.AdamW
    assert module.scheduler.func is WarmupCosineAnnealingLR
    net = module.net
    assert (net.emb_size, net.depth, bool(net.concat_features)) == (emb, depth, concat)
    assert net.normal_id == cfg["data"]["normal_id"] and len(net.classnames) == cfg["data"]["num_classes"]
    if classes is not None:
        assert len(net.classnames) == classes
    assert module.num_classes == cfg["data"]["num_classes"]
