"""Frame-level metrics of `test_epoch_end` (src/models/anomaly_clip_module.py:501-619) on torch
tensors (CPU or GPU), without torchmetrics: binary AUROC / average precision of the anomaly
score, per-class one-vs-rest AUROC / AP of the class probabilities, top-k accuracy.
Tie handling follows sklearn (thresholds at distinct score values)."""
from __future__ import annotations

from typing import Dict

import torch


def _curve_counts(scores: torch.Tensor, target: torch.Tensor):
    """Cumulative TP / FP at every distinct threshold, scores descending."""
    order = torch.argsort(scores, descending=True, stable=True)
    s, t = scores[order], target[order].to(torch.float64)
    distinct = torch.nonzero(s[1:] != s[:-1]).flatten()
    idx = torch.cat((distinct, torch.tensor([s.numel() - 1], device=s.device)))
    tps = torch.cumsum(t, 0)[idx]
    fps = (idx + 1).to(torch.float64) - tps
    return tps, fps


def binary_auroc(scores: torch.Tensor, target: torch.Tensor) -> float:
    target = target.to(torch.bool)
    n_pos, n_neg = int(target.sum()), int((~target).sum())
    if n_pos == 0 or n_neg == 0:
        return float("nan")
    tps, fps = _curve_counts(scores.to(torch.float64), target)
    zero = tps.new_zeros(1)
    tpr = torch.cat((zero, tps)) / n_pos
    fpr = torch.cat((zero, fps)) / n_neg
    return float(torch.trapz(tpr, fpr))


def binary_average_precision(scores: torch.Tensor, target: torch.Tensor) -> float:
    target = target.to(torch.bool)
    n_pos = int(target.sum())
    if n_pos == 0:
        return float("nan")
    tps, fps = _curve_counts(scores.to(torch.float64), target)
    precision = tps / (tps + fps)
    recall = tps / n_pos
    prev = torch.cat((recall.new_zeros(1), recall[:-1]))
    return float(((recall - prev) * precision).sum())


def expand_class_probs(class_probs: torch.Tensor, scores: torch.Tensor, normal_id: int) -> torch.Tensor:
    """(N, C-1) abnormal-class probabilities -> (N, C) with the normal class re-inserted as
    1 - score at column `normal_id` (anomaly_clip_module.py:537-546)."""
    normal = (1.0 - scores).unsqueeze(1)
    return torch.cat((class_probs[:, :normal_id], normal, class_probs[:, normal_id:]), dim=1)


def frame_metrics(scores: torch.Tensor, class_probs: torch.Tensor, labels: torch.Tensor,
                  normal_id: int) -> Dict[str, float]:
    binary = labels != normal_id
    out = {"AUC": binary_auroc(scores, binary), "AP": binary_average_precision(scores, binary)}
    probs = expand_class_probs(class_probs, scores, normal_id)
    aucs, aps = [], []
    for c in range(probs.shape[1]):
        if c == normal_id:
            continue
        t = labels == c
        if 0 < int(t.sum()) < t.numel():
            aucs.append(binary_auroc(probs[:, c], t))
            aps.append(binary_average_precision(probs[:, c], t))
    if aucs:
        out["mAUC"] = sum(aucs) / len(aucs)
        out["mAP"] = sum(aps) / len(aps)
    abn = binary
    if abn.any():
        top = probs[abn].topk(min(5, probs.shape[1]), dim=1).indices
        out["top1"] = float((top[:, 0] == labels[abn]).double().mean())
        out["top5"] = float((top == labels[abn].unsqueeze(1)).any(dim=1).double().mean())
    return out
