from anomalyclip_b200.models import ClassificationHead  # noqa: F401
