"""data.load_from_features=False (anomaly_clip_datamodule.py:86-87)."""
from anomalyclip_b200.data import FrameVideoDataset as VideoFrameDataset  # noqa: F401
