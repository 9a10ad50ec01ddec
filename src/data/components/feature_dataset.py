"""data.load_from_features=True (anomaly_clip_datamodule.py:84-85)."""
from anomalyclip_b200.data import FeatureVideoDataset as VideoFrameDataset  # noqa: F401
