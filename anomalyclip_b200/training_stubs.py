"""Import-path keepers for the two training-side classes the reference's model configs name.

`hydra.utils.instantiate(cfg.model)` in `src/eval.py` builds the `loss` and (as a partial) the
`scheduler` entries of configs/model/*.yaml even for an evaluation run, and Lightning checkpoints
pickle them inside `hyper_parameters`; so `src.models.components.loss.ComputeLoss` and
`src.models.components.scheduler.WarmupCosineAnnealingLR` have to resolve for `eval.py` to be a
drop-in.  Training itself is out of scope of the B200 inference path:

* `ComputeLoss` keeps the configured hyper-parameters and refuses to be called;
* `WarmupCosineAnnealingLR` is a complete learning-rate schedule (polynomial warm-up to the base
  rate, then half-cosine decay to `final_factor` x base; scheduler.py:22-68 defines the same
  curve), since a schedule is a few lines and a checkpoint's `lr_schedulers` state refers to it.
"""
from __future__ import annotations

import math
from typing import Any, List, Sequence, Union

from torch.optim.lr_scheduler import _LRScheduler

Number = Union[int, float]


class ComputeLoss:
    """configs/model/*.yaml `loss:` (loss.py:21-50 lists the keywords)."""

    def __init__(self, normal_id: int, num_topk: int, lambda_dir_abn: float, lambda_dir_nor: float,
                 lambda_topk_abn: float, lambda_bottomk_abn: float, lambda_topk_nor: float,
                 lambda_smooth: float, lambda_sparse: float, frames_per_segment: int,
                 num_segments: int, **extra: Any) -> None:
        self.normal_id, self.num_topk = normal_id, num_topk
        self.lambda_dir_abn, self.lambda_dir_nor = lambda_dir_abn, lambda_dir_nor
        self.lambda_topk_abn, self.lambda_bottomk_abn = lambda_topk_abn, lambda_bottomk_abn
        self.lambda_topk_nor = lambda_topk_nor
        self.lambda_smooth, self.lambda_sparse = lambda_smooth, lambda_sparse
        self.frames_per_segment, self.num_segments = frames_per_segment, num_segments
        self.extra = extra

    def __call__(self, *args: Any, **kwargs: Any):
        raise NotImplementedError("ComputeLoss: the training objective is out of scope of the B200 "
                                  "inference path; only its configuration is kept")


def _per_group(value: Union[Number, Sequence[Number]], groups: int) -> List[Number]:
    if isinstance(value, (int, float)):
        return [value] * groups
    value = list(value)
    if len(value) != groups:
        raise ValueError(f"expected {groups} values (one per parameter group), got {len(value)}")
    return value


class WarmupCosineAnnealingLR(_LRScheduler):
    """lr(epoch) per parameter group:
         epoch <  warmup_epochs : warmup_lr + (base - warmup_lr) * (epoch / warmup_epochs) ** power
         epoch >= warmup_epochs : base * (final + (1 - final) * (1 + cos(pi * p)) / 2),
                                  p = min(1, (epoch - warmup_epochs) / (total_epoch - warmup_epochs))"""

    def __init__(self, optimizer, total_epoch: int, successor=None, final_factor: float = 0,
                 warmup_epochs: Union[Number, Sequence[Number]] = 0,
                 warmup_powers: Union[Number, Sequence[Number]] = 1,
                 warmup_lrs: Union[Number, Sequence[Number]] = 0, last_epoch: int = -1) -> None:
        groups = len(optimizer.param_groups)
        self.total_epoch, self.final_factor, self.successor = total_epoch, final_factor, successor
        self.warmup_epochs = _per_group(warmup_epochs, groups)
        self.warmup_powers = _per_group(warmup_powers, groups)
        self.warmup_lrs = _per_group(warmup_lrs, groups)
        super().__init__(optimizer, last_epoch)

    def get_lr(self) -> List[float]:
        lrs = []
        for base, warm, power, start in zip(self.base_lrs, self.warmup_epochs, self.warmup_powers,
                                            self.warmup_lrs):
            if self.last_epoch < warm:
                lrs.append(start + (base - start) * (self.last_epoch / warm) ** power)
            else:
                span = max(self.total_epoch - warm, 1e-12)
                p = min((self.last_epoch - warm) / span, 1.0)
                lrs.append(base * (self.final_factor + (1 - self.final_factor) * (1 + math.cos(math.pi * p)) / 2))
        return lrs
