from anomalyclip_b200.datamodule import AnomalyCLIPDataModule  # noqa: F401  (configs/data/*.yaml:1)
