"""metrics.py against sklearn; sharded mean under gloo."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from sklearn.metrics import average_precision_score, roc_auc_score

from anomalyclip_b200 import metrics
from tests.test_host_cpu import _free_port


@pytest.mark.parametrize("ties", [False, True])
def test_binary_auroc_and_ap_match_sklearn(ties):
    rng = np.random.default_rng(3)
    y = rng.integers(0, 2, 5000)
    s = rng.random(5000) + 0.3 * y
    if ties:
        s = np.round(s, 1)
    a = metrics.binary_auroc(torch.from_numpy(s), torch.from_numpy(y))
    p = metrics.binary_average_precision(torch.from_numpy(s), torch.from_numpy(y))
    assert abs(a - roc_auc_score(y, s)) < 1e-12
    assert abs(p - average_precision_score(y, s)) < 1e-12


def test_frame_metrics_and_class_expansion():
    torch.manual_seed(0)
    n, C, normal_id = 400, 6, 2
    labels = torch.randint(0, C, (n,))
    scores = torch.rand(n) * 0.5 + 0.5 * (labels != normal_id)
    probs = torch.softmax(torch.randn(n, C - 1), 1) * scores[:, None]
    full = metrics.expand_class_probs(probs, scores, normal_id)
    assert full.shape == (n, C) and torch.allclose(full[:, normal_id], 1 - scores)
    assert torch.equal(full[:, 3], probs[:, 2])           # classes >= normal_id shift by one
    m = metrics.frame_metrics(scores, probs, labels, normal_id)
    assert abs(m["AUC"] - roc_auc_score((labels != normal_id).numpy(), scores.numpy())) < 1e-12
    assert 0 <= m["top1"] <= m["top5"] <= 1 and "mAUC" in m


def _mean_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from anomalyclip_b200.distributed import sharded_mean
    rows = torch.arange(12, dtype=torch.float32).reshape(6, 2)[rank::world]   # rank's shard
    q.put((rank, sharded_mean(rows.sum(0), rows.shape[0]).tolist()))
    dist.destroy_process_group()


def test_sharded_mean_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mean_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    expect = torch.arange(12, dtype=torch.float32).reshape(6, 2).mean(0).tolist()
    assert got[0] == expect and got[1] == expect
