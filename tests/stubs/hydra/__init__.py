"""Test stand-in for hydra (see tests/stubs/README.md).  `@hydra.main` hands the decorated function
the configuration found in the JSON file $ACLIP_TEST_CFG instead of composing YAML groups."""
import functools
import json
import os

from omegaconf import DictConfig

from . import utils  # noqa: F401


def main(version_base=None, config_path=None, config_name=None):
    def deco(fn):
        @functools.wraps(fn)
        def run():
            with open(os.environ["ACLIP_TEST_CFG"]) as fp:
                return fn(DictConfig(json.load(fp)))
        return run
    return deco
