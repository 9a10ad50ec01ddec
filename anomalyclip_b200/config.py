"""Hydra-free composition of the reference's own YAML configs.

`src/eval.py` of the reference lets Hydra compose `configs/eval.yaml` (defaults: `data: ucfcrime`,
`model: anomaly_clip_ucfcrime`, ...) and `hydra.utils.instantiate` the `data` and `model` nodes.  The
inference path needs exactly those two nodes, so this module does the same with PyYAML only: read
the two files, apply `key.path=value` overrides, resolve `${data.x}` style interpolations
(`${oc.env:VAR}` too; Hydra-runtime resolvers are left alone) and build the `_target_` objects,
honouring `_partial_`.  Every key of the reference's files is passed through verbatim.

    cfg = load_eval_config("/path/to/AnomalyCLIP/configs", data="ucfcrime",
                           model="anomaly_clip_ucfcrime", overrides=["data.frames_root=/data/feats"])
    datamodule, module = instantiate(cfg["data"]), instantiate(cfg["model"])
"""
from __future__ import annotations

import functools
import importlib
import os
import re
from pathlib import Path
from typing import Any, Dict, Iterable

import yaml

_INTERP = re.compile(r"\$\{([^${}]+)\}")
_RESERVED = ("_target_", "_partial_", "_recursive_", "_convert_", "_args_")


def _set(cfg: Dict[str, Any], dotted: str, value: Any) -> None:
    node = cfg
    keys = dotted.split(".")
    for k in keys[:-1]:
        node = node.setdefault(k, {})
        if not isinstance(node, dict):
            raise KeyError(f"override '{dotted}': '{k}' is not a mapping")
    node[keys[-1]] = value


def _lookup(cfg: Dict[str, Any], dotted: str) -> Any:
    node: Any = cfg
    for k in dotted.split("."):
        if not isinstance(node, dict) or k not in node:
            raise KeyError(f"interpolation ${{{dotted}}} cannot be resolved")
        node = node[k]
    return node


def _resolve(value: Any, root: Dict[str, Any], depth: int = 0) -> Any:
    if depth > 16:
        raise RecursionError("interpolation loop in the configuration")
    if isinstance(value, dict):
        return {k: _resolve(v, root, depth) for k, v in value.items()}
    if isinstance(value, list):
        return [_resolve(v, root, depth) for v in value]
    if not isinstance(value, str):
        return value

    def one(expr: str) -> Any:
        if expr.startswith("oc.env:"):
            name, _, default = expr[len("oc.env:"):].partition(",")
            if name in os.environ:
                return os.environ[name]
            if default:
                return default
            raise KeyError(f"environment variable {name} is not set")
        if ":" in expr:                    # hydra:..., now:...: not available outside Hydra
            return "${" + expr + "}"
        return _resolve(_lookup(root, expr), root, depth + 1)

    whole = _INTERP.fullmatch(value)
    if whole:                              # "${data.num_classes}" keeps the referenced value's type
        return one(whole.group(1))
    return _INTERP.sub(lambda m: str(one(m.group(1))), value)


def load_eval_config(configs_dir: str, data: str = "ucfcrime", model: str = "anomaly_clip_ucfcrime",
                     overrides: Iterable[str] = ()) -> Dict[str, Any]:
    """{'data': ..., 'model': ...} composed from `<configs_dir>/data/<data>.yaml` and
    `<configs_dir>/model/<model>.yaml` like the defaults list of configs/eval.yaml does."""
    root = Path(configs_dir)
    cfg: Dict[str, Any] = {}
    for group, name in (("data", data), ("model", model)):
        f = root / group / (name if name.endswith(".yaml") else name + ".yaml")
        cfg[group] = yaml.safe_load(f.read_text()) or {}
    for item in overrides:
        key, sep, raw = item.partition("=")
        if not sep:
            raise ValueError(f"override '{item}' is not of the form key.path=value")
        _set(cfg, key.strip(), yaml.safe_load(raw))
    return _resolve(cfg, cfg)


def instantiate(node: Any, **extra: Any) -> Any:
    """Build the object a `_target_` node describes (nested `_target_` nodes first, `_partial_: true`
    gives a functools.partial), the way hydra.utils.instantiate does for these configs."""
    if isinstance(node, list):
        return [instantiate(v) for v in node]
    if not isinstance(node, dict):
        return node
    kwargs = {k: instantiate(v) for k, v in node.items() if k not in _RESERVED}
    if "_target_" not in node:
        return kwargs
    kwargs.update(extra)
    module, _, name = str(node["_target_"]).rpartition(".")
    target = getattr(importlib.import_module(module), name)
    if node.get("_partial_", False):
        return functools.partial(target, **kwargs)
    return target(**kwargs)
