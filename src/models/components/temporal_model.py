from anomalyclip_b200.models import TemporalModel  # noqa: F401
