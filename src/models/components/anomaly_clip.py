from anomalyclip_b200.models import AnomalyCLIP  # noqa: F401  (configs/model/*.yaml:16)
