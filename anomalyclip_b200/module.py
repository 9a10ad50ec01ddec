"""`AnomalyCLIPModule`: the reference's LightningModule test path
(/root/reference/src/models/anomaly_clip_module.py:118-132,406-498) over the B200 net.

Lightning is optional: when `pytorch_lightning` is importable the class derives from
`LightningModule` (so `Trainer.test(model, datamodule, ckpt_path)` in the reference's
`src/eval.py` drives it unchanged); otherwise it is a plain `nn.Module` with the same hooks, which
is what the tests and the benchmark in this repository call directly.
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import Any, Dict, List, Optional

import torch
from torch import nn

try:  # Lightning is not installed in the build image; tests drive this branch with a stub package
    import pytorch_lightning as _pl
    from pytorch_lightning import LightningModule as _Base
    _HAVE_LIGHTNING = True
    try:
        _LIGHTNING_MAJOR = int(str(getattr(_pl, "__version__", "1.8.3")).split(".")[0])
    except ValueError:
        _LIGHTNING_MAJOR = 1
except Exception:  # noqa: BLE001
    _Base = nn.Module
    _HAVE_LIGHTNING = False
    _LIGHTNING_MAJOR = 0


class AnomalyCLIPModule(_Base):
    def __init__(self, net: nn.Module, optimizer: Any = None, scheduler: Any = None,
                 loss: Any = None, **kwargs: Any) -> None:
        super().__init__()
        if _HAVE_LIGHTNING:  # pragma: no cover
            self.save_hyperparameters(logger=False, ignore=["net"])
        self.net = net
        self.criterion, self.optimizer, self.scheduler = loss, optimizer, scheduler
        self.num_classes = kwargs.get("num_classes")
        self.save_dir = kwargs.get("save_dir")
        self.ncentroid: Optional[torch.Tensor] = None
        self.last_metrics: Dict[str, float] = {}
        self.labels: List[torch.Tensor] = []
        self.abnormal_scores: List[torch.Tensor] = []
        self.class_probs: List[torch.Tensor] = []

    # anomaly_clip_module.py:118-132 -- same positional call into the net
    def forward(self, image_features: torch.Tensor, labels, ncentroid: torch.Tensor,
                segment_size: int = 1, test_mode: bool = False):
        return self.net(image_features, labels, ncentroid, segment_size, test_mode)

    @property
    def _device(self) -> torch.device:
        return next(self.net.temporal_model.parameters()).device

    # ---- ncentroid side-car (anomaly_clip_module.py:406-445)
    def _centroid_dir(self) -> Path:
        trainer = getattr(self, "trainer", None) if _HAVE_LIGHTNING else None
        ckpt = getattr(trainer, "ckpt_path", None) if trainer is not None else None
        if ckpt:
            run = os.path.normpath(Path(ckpt).parent).split(os.path.sep)[-1]
            return Path(os.environ.get("ACLIP_RUNS_DIR", "/usr/src/app/logs/train/runs")) / run
        return Path(self.save_dir or ".")

    @torch.no_grad()
    def compute_ncentroid(self, loader, load_from_features: bool) -> torch.Tensor:
        """Mean feature of the normal training videos, streamed in test mode (:419-441).  Under
        torch.distributed each rank streams ITS shard of the loader and the (sum, count) pairs
        are combined with one all-reduce."""
        dev = self._device
        total = torch.zeros(self.net.embedding_dim, dtype=torch.float64, device=dev)
        count = 0
        for batch in loader:
            x, nlabels = batch[0], batch[1]
            n_real = int(nlabels.reshape(-1).shape[0])
            if load_from_features:
                feats = x.reshape(-1, x.shape[-1])[:n_real].to(dev)
            else:
                c, h, w = x.shape[-3:]
                feats = self.net.image_encoder(x.reshape(-1, c, h, w)[:n_real].to(dev))
            total += feats.double().sum(dim=0)
            count += feats.shape[0]
        from .distributed import sharded_mean
        return sharded_mean(total, count)

    def on_test_start(self) -> None:
        if self.ncentroid is not None:
            return
        d = self._centroid_dir()
        f = d / "ncentroid.pt"
        if f.is_file():
            self.ncentroid = torch.load(f, map_location="cpu")
            return
        trainer = getattr(self, "trainer", None) if _HAVE_LIGHTNING else None
        if trainer is None:
            raise RuntimeError(f"{f} not found: set module.ncentroid or call compute_ncentroid()")
        dm = trainer.datamodule  # pragma: no cover - needs Lightning
        self.ncentroid = self.compute_ncentroid(dm.train_dataloader_test_mode(),
                                                dm.hparams.load_from_features)
        d.mkdir(parents=True, exist_ok=True)
        torch.save(self.ncentroid.cpu(), f)

    # ---- per-video step (anomaly_clip_module.py:459-498)
    @torch.no_grad()
    def test_step(self, batch: Any, batch_idx: int = 0) -> Dict[str, torch.Tensor]:
        image_features, labels, label, segment_size, path = batch
        dev = self._device
        image_features = image_features.to(dev, non_blocking=True)
        labels = labels.squeeze(0).to(dev)
        if self.ncentroid.device != dev:   # once: a side-car centroid is loaded on the CPU
            self.ncentroid = self.ncentroid.to(dev)
        if torch.is_tensor(segment_size):
            segment_size = int(segment_size.reshape(-1)[0])
        similarity, abnormal_scores = self.forward(image_features, labels, self.ncentroid,
                                                   segment_size, test_mode=True)
        # softmax(similarity) * score comes fused out of the head kernel (:473-477)
        class_probs = self.net.class_probs
        num_labels = labels.shape[0]                                           # :480-483
        out = {"abnormal_scores": abnormal_scores[:num_labels], "labels": labels,
               "class_probs": class_probs[:num_labels]}
        self.labels.append(labels.cpu())
        self.abnormal_scores.append(out["abnormal_scores"].cpu())
        self.class_probs.append(out["class_probs"].cpu())
        return out

    def predict_step(self, batch: Any, batch_idx: int = 0, dataloader_idx: int = 0):
        out = self.test_step(batch, batch_idx)
        self.labels.pop(), self.abnormal_scores.pop(), self.class_probs.pop()
        return out

    # ---- metrics (anomaly_clip_module.py:501-619, the numeric part)
    def finish_test_epoch(self) -> Dict[str, float]:
        """Metrics over everything `test_step` accumulated, then clear the accumulators.  Called
        again on empty accumulators (Lightning 1.8 fires `test_epoch_end(outputs)` and then
        `on_test_epoch_end()`) it returns the metrics of the epoch just finished."""
        from .metrics import frame_metrics

        if not self.labels:
            return self.last_metrics
        labels = torch.cat(self.labels)
        scores = torch.cat(self.abnormal_scores)
        probs = torch.cat(self.class_probs)
        metrics = {f"test/{k}": v for k, v in
                   frame_metrics(scores, probs, labels, self.net.normal_id).items()}
        if self.save_dir:
            Path(self.save_dir).mkdir(parents=True, exist_ok=True)
            with open(Path(self.save_dir) / "metrics.json", "w") as fp:
                json.dump(metrics, fp, indent=4, sort_keys=True)
        self.labels.clear(), self.abnormal_scores.clear(), self.class_probs.clear()
        self.last_metrics = metrics
        return metrics

    # One epoch-end hook per Lightning generation: 1.x (the reference pins 1.8.3) calls
    # `test_epoch_end(outputs)`; 2.x rejects a module that overrides it and calls
    # `on_test_epoch_end()` instead.  Without Lightning the 1.x name is kept for direct callers.
    if _LIGHTNING_MAJOR >= 2:
        def on_test_epoch_end(self) -> None:
            self.finish_test_epoch()
    else:
        def test_epoch_end(self, outputs: Any = None) -> Dict[str, float]:
            return self.finish_test_epoch()
