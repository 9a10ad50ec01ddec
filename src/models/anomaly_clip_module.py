from anomalyclip_b200.module import AnomalyCLIPModule  # noqa: F401  (configs/model/*.yaml:1)
