"""Test stand-in for omegaconf (see tests/stubs/README.md): attribute-style nested dicts."""
from contextlib import contextmanager


class DictConfig(dict):
    def __init__(self, data=None):
        super().__init__()
        for k, v in (data or {}).items():
            self[k] = _wrap(v)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as exc:
            raise AttributeError(name) from exc

    def __setattr__(self, name, value):
        self[name] = _wrap(value)


class ListConfig(list):
    pass


def _wrap(v):
    if isinstance(v, DictConfig):
        return v
    if isinstance(v, dict):
        return DictConfig(v)
    if isinstance(v, (list, tuple)):
        return ListConfig(_wrap(x) for x in v)
    return v


def _plain(v):
    if isinstance(v, dict):
        return {k: _plain(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_plain(x) for x in v]
    return v


class OmegaConf:
    @staticmethod
    def create(data=None):
        return DictConfig(data or {})

    @staticmethod
    def to_container(cfg, resolve=True):
        return _plain(cfg)

    @staticmethod
    def to_yaml(cfg, resolve=True):
        import yaml
        return yaml.safe_dump(_plain(cfg))


@contextmanager
def open_dict(cfg):
    yield cfg
