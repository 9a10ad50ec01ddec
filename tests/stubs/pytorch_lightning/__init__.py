"""Test stand-in for pytorch-lightning 1.8.x (the version the reference pins, requirements.txt:5):
just enough of LightningModule / LightningDataModule / Trainer.test to drive the reference's
`src/eval.py` with the hook order of the 1.8 evaluation loop:

    datamodule.setup("test") -> load checkpoint -> model.on_test_start()
    -> test_step(batch, i) per batch -> test_epoch_end(outputs) -> on_test_epoch_end() -> on_test_end()
"""
from types import SimpleNamespace

import torch
from torch import nn

from . import loggers, utilities  # noqa: F401
from .trainer import Trainer  # noqa: F401

__version__ = "1.8.3"


class _HParams(dict):
    __getattr__ = dict.__getitem__


class _HparamsMixin:
    def save_hyperparameters(self, *args, logger=True, ignore=None, **kwargs):
        hp = _HParams()
        for a in args:
            if isinstance(a, dict):
                hp.update(a)
        object.__setattr__(self, "_hparams", hp)

    @property
    def hparams(self):
        return getattr(self, "_hparams", _HParams())


class LightningModule(_HparamsMixin, nn.Module):
    def __init__(self):
        super().__init__()
        self.trainer = None

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def log(self, *args, **kwargs):
        pass

    # hooks the 1.8 loop fires; a module overrides what it needs
    def on_test_start(self):
        pass

    def on_test_epoch_end(self):
        pass

    def on_test_end(self):
        pass


class LightningDataModule(_HparamsMixin):
    def __init__(self):
        super().__init__()

    def setup(self, stage=None):
        pass


class Callback:
    pass
