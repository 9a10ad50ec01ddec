"""hydra.utils.instantiate: builds `_target_` nodes (nested first, `_partial_` honoured); keyword
overrides are merged into the top node like Hydra does."""
from omegaconf import OmegaConf


def instantiate(node, **kwargs):
    from anomalyclip_b200.config import instantiate as _instantiate   # same semantics, PyYAML-free here
    plain = OmegaConf.to_container(node)
    return _instantiate(plain, **kwargs)
