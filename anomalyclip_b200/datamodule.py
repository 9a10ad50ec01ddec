"""`AnomalyCLIPDataModule`: the reference's LightningDataModule
(/root/reference/src/data/anomaly_clip_datamodule.py:12-213) for the inference path.

Same constructor keywords (the keys of configs/data/*.yaml), same attributes (`test_data`,
`train_data_normal_test_mode`), same loaders (`test_dataloader`, `val_dataloader`,
`train_dataloader_test_mode`) yielding the reference's test-mode 5-tuples, batch_size_test videos
per batch under torch's default collate (anomaly_clip_datamodule.py:175-203).  The training
datasets / loaders (random segment sampling, abnormal + normal halves) are out of scope and raise.

Raw-frame videos leave the loader as uint8 (T, 3, S, S) by default: the reference's ToTensor +
Normalize run on the GPU inside the image encoder (4x less host->device traffic, same arithmetic);
`frame_output="normalised"` restores the reference's fp32 tensors.

Lightning is optional, exactly as for `AnomalyCLIPModule`.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any, Dict, Optional, Tuple

from torch.utils.data import DataLoader, Dataset

from .data import FeatureVideoDataset, FrameVideoDataset, read_annotation_file, read_temporal_annotations

try:  # pragma: no cover - Lightning is not installed in the build image
    from pytorch_lightning import LightningDataModule as _Base
    _HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    _Base = object
    _HAVE_LIGHTNING = False

# keys of configs/data/*.yaml the test path reads, with the reference's defaults where it has any
_DEFAULTS: Dict[str, Any] = dict(
    num_segments=32, seg_length=16, batch_size_test=1, num_classes=None, input_size=224,
    load_from_features=True, frames_root=None, frames_root_val=None, annotations_root=None,
    normal_id=0, image_tmpl="{:06d}.jpg", stride=1, ncrops=1, annotation_file_anomaly=None,
    annotation_file_normal=None, annotation_file_test=None, annotation_file_temporal_test=None,
    labels_file=None, spatialannotationdir_path=None, visualize=False, frame_output="uint8")


class AnomalyCLIPDataModule(_Base):
    def __init__(self, data_dir: str = "data/", train_val_test_split: Tuple[int, int, int] = (55_000, 5_000, 10_000),
                 batch_size: int = 64, num_workers: int = 0, pin_memory: bool = False, **kwargs: Any) -> None:
        super().__init__()
        hp = dict(_DEFAULTS, data_dir=data_dir, train_val_test_split=train_val_test_split,
                  batch_size=batch_size, num_workers=num_workers, pin_memory=pin_memory)
        hp.update(kwargs)
        if _HAVE_LIGHTNING:  # pragma: no cover
            self.save_hyperparameters(hp, logger=False)
        else:
            self.hparams = SimpleNamespace(**hp)
        self.train_data_normal: Optional[Dataset] = None
        self.train_data_anomaly: Optional[Dataset] = None
        self.test_data: Optional[Dataset] = None
        self.train_data_normal_test_mode: Optional[Dataset] = None

    @property
    def num_classes(self):  # anomaly_clip_datamodule.py:68-70
        return self.hparams.num_classes

    def prepare_data(self) -> None:
        pass

    def _test_mode_dataset(self, annotation_file: str, temporal_annotation_file: Optional[str]) -> Dataset:
        h = self.hparams
        if h.load_from_features:   # data/components/feature_dataset.py
            return FeatureVideoDataset(
                read_annotation_file(annotation_file, h.frames_root), num_segments=h.num_segments,
                seg_length=h.seg_length, stride=h.stride, ncrops=h.ncrops, normal_id=h.normal_id,
                annotations=read_temporal_annotations(temporal_annotation_file))
        return FrameVideoDataset(    # data/components/video_dataset.py
            root_path=h.frames_root, annotationfile_path=annotation_file, normal_id=h.normal_id,
            num_segments=h.num_segments, frames_per_segment=h.seg_length, imagefile_template=h.image_tmpl,
            test_mode=True, ncrops=h.ncrops, temporal_annotation_file=temporal_annotation_file,
            labels_file=h.labels_file, stride=h.stride, input_size=h.input_size, output=h.frame_output)

    def setup(self, stage: Optional[str] = None) -> None:
        """anomaly_clip_datamodule.py:79-141: the two test-mode datasets (the evaluation set and the
        normal training videos `on_test_start` averages into `ncentroid`)."""
        h = self.hparams
        if self.test_data is None and h.annotation_file_test:
            self.test_data = self._test_mode_dataset(h.annotation_file_test, h.annotation_file_temporal_test)
        if self.train_data_normal_test_mode is None and h.annotation_file_normal:
            self.train_data_normal_test_mode = self._test_mode_dataset(h.annotation_file_normal, None)

    def _loader(self, dataset: Optional[Dataset]) -> DataLoader:
        if dataset is None:
            raise RuntimeError("AnomalyCLIPDataModule: call setup() with the annotation files configured first")
        h = self.hparams
        return DataLoader(dataset=dataset, batch_size=h.batch_size_test, num_workers=h.num_workers,
                          pin_memory=h.pin_memory, shuffle=False, drop_last=False)

    def train_dataloader(self):
        raise NotImplementedError("AnomalyCLIPDataModule: the training loaders are out of scope of the "
                                  "B200 inference path")

    def val_dataloader(self) -> DataLoader:        # :165-173
        return self._loader(self.test_data)

    def test_dataloader(self) -> DataLoader:       # :175-183
        return self._loader(self.test_data)

    def train_dataloader_test_mode(self) -> DataLoader:   # :185-193
        return self._loader(self.train_data_normal_test_mode)

    def teardown(self, stage: Optional[str] = None) -> None:
        pass

    def state_dict(self) -> Dict[str, Any]:
        return {}

    def load_state_dict(self, state_dict: Dict[str, Any]) -> None:
        pass
