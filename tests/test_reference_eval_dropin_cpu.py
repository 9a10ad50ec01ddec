"""The reference's UNMODIFIED `src/eval.py` (Hydra main -> evaluate(cfg) -> Trainer.test) driving
this repository's `AnomalyCLIPModule` / `AnomalyCLIPDataModule` through the `src.*` import paths of
its configs.  hydra / omegaconf / pyrootutils / pytorch_lightning are not installed in
the build image: tests/stubs holds minimal stand-ins (Lightning 1.8 hook order: test_epoch_end(outputs)
THEN on_test_epoch_end()).  The net is a CPU stand-in -- the kernels are covered by the `-m gpu`
tests; this one covers the boundary: import shadowing, `_target_` resolution, checkpoint load,
`on_test_start` centroid side-car, per-video `test_step`, epoch-end hooks, metrics."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from tests.test_datamodule_cpu import _direct_metrics, _write_feature_videos

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")

pytestmark = pytest.mark.skipif(not (REF / "src" / "eval.py").is_file(),
                                reason="runs the reference's own entry script (present in the build container)")


def _run_dropin(tmp_path: Path, lightning_env=None):
    vids, feats = _write_feature_videos(tmp_path)
    ckpt = tmp_path / "train_logs" / "run_2024" / "last.ckpt"
    ckpt.parent.mkdir(parents=True, exist_ok=True)
    torch.save({"state_dict": {"net.temporal_model.weight": torch.ones(1, 1),
                               "net.temporal_model.bias": torch.zeros(1)},
                "hyper_parameters": {}}, ckpt)
    cfg = {
        "ckpt_path": str(ckpt), "task_name": "eval", "tags": ["dev"], "logger": None, "extras": None,
        "paths": {"output_dir": str(tmp_path / "hydra_out")},
        "data": {"_target_": "src.data.anomaly_clip_datamodule.AnomalyCLIPDataModule",
                 "num_segments": 32, "seg_length": 16, "batch_size_test": 1, "num_classes": 14,
                 "load_from_features": True, "frames_root": str(tmp_path / "feats"), "normal_id": 7,
                 "num_workers": 0, "pin_memory": False,
                 "annotation_file_normal": str(tmp_path / "normal.txt"),
                 "annotation_file_test": str(tmp_path / "test.txt"),
                 "annotation_file_temporal_test": str(tmp_path / "temporal.txt")},
        "model": {"_target_": "src.models.anomaly_clip_module.AnomalyCLIPModule",
                  "net": {"_target_": "aclip_standin_net.Net", "arch": "ViT-B/16"},
                  "optimizer": None, "scheduler": None, "loss": None,
                  "num_classes": 14, "save_dir": str(tmp_path / "out")},
        "trainer": {"_target_": "pytorch_lightning.trainer.Trainer", "accelerator": "cpu", "devices": 1},
    }
    cfg_file, log_file = tmp_path / "cfg.json", tmp_path / "trainer_log.json"
    cfg_file.write_text(json.dumps(cfg))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(ROOT / "tests" / "stubs"), str(ROOT)]),
               ACLIP_TEST_CFG=str(cfg_file), ACLIP_TEST_TRAINER_LOG=str(log_file),
               ACLIP_RUNS_DIR=str(tmp_path / "runs"))
    env.pop("ACLIP_REFERENCE_ROOT", None)
    res = subprocess.run([sys.executable, "-m", "anomalyclip_b200.dropin", str(REF / "src" / "eval.py")],
                         cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return vids, feats, json.loads(log_file.read_text()), tmp_path / "runs" / "run_2024"


def test_unmodified_reference_eval_script_drives_the_b200_classes(tmp_path):
    vids, feats, log, run_dir = _run_dropin(tmp_path)
    # the reference's own Trainer.test sequence reached the B200 classes through the `src.*` paths
    assert log["model_class"] == "anomalyclip_b200.module.AnomalyCLIPModule"
    assert log["datamodule_class"] == "anomalyclip_b200.datamodule.AnomalyCLIPDataModule"
    assert log["calls"] == ["on_test_start", "test_step x4", "test_epoch_end", "on_test_epoch_end"]
    # on_test_start found no side-car, averaged the normal training videos and saved the centroid in
    # the checkpoint's run directory (anomaly_clip_module.py:406-445)
    side_car = run_dir / "ncentroid.pt"
    assert side_car.is_file()
    centroid = torch.load(side_car)
    normal = torch.from_numpy(np.concatenate([feats["Normal003"], feats["Normal004"]])).double()
    assert torch.allclose(centroid.double(), normal.mean(0), atol=1e-6)
    # per-video test_step with the padded rows trimmed; the second epoch-end hook of Lightning 1.8
    # (on_test_epoch_end after test_epoch_end) did not break or overwrite the metrics
    assert log["net_calls"][-4:] == [[600, 2], [90, 1], [130, 1], [75, 1]]
    ref = _direct_metrics(vids, feats, centroid)
    assert set(log["metrics"]) == {f"test/{k}" for k in ref}
    for k, v in ref.items():
        assert abs(log["metrics"][f"test/{k}"] - v) < 1e-6, k
    saved = json.loads((tmp_path / "out" / "metrics.json").read_text())
    assert saved == log["metrics"]


def test_second_run_reuses_the_centroid_side_car(tmp_path):
    _, _, first, run_dir = _run_dropin(tmp_path)
    stamp = (run_dir / "ncentroid.pt").stat().st_mtime_ns
    _, _, second, _ = _run_dropin(tmp_path)
    assert (run_dir / "ncentroid.pt").stat().st_mtime_ns == stamp      # loaded, not recomputed
    assert len(first["net_calls"]) == 4 and len(second["net_calls"]) == 4
    assert second["metrics"] == first["metrics"]
