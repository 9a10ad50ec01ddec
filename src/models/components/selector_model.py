from anomalyclip_b200.models import SelectorModel  # noqa: F401
