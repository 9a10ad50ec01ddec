"""Lightning-free driver of the reference's evaluation flow.

`src/eval.py:70-75` of the reference instantiates the datamodule and the module from the Hydra
config and calls `Trainer.test(model, datamodule, ckpt_path)`, which runs, in this order,
`datamodule.setup`, `on_test_start` (the `ncentroid` side-car, anomaly_clip_module.py:406-445), one
`test_step` per video of `test_dataloader()` (:459-498) and `test_epoch_end` (:501-619).  With
Lightning installed that call works unchanged on `AnomalyCLIPModule` / `AnomalyCLIPDataModule`;
`evaluate()` is the same sequence as a plain function for environments without Lightning (this
repository's image, the benchmark box)."""
from __future__ import annotations

from typing import Dict, Optional

import torch


@torch.no_grad()
def evaluate(module, datamodule, ncentroid: Optional[torch.Tensor] = None,
             checkpoint: Optional[str] = None) -> Dict[str, float]:
    """Returns the metric dict of `test_epoch_end` (test/AUC, test/AP, test/mAUC, ...).

    checkpoint: optional Lightning `.ckpt` (or plain state_dict file); its `state_dict` is loaded
    into the module first, as `Trainer.test(ckpt_path=...)` does.
    ncentroid: use this centroid instead of the side-car / the normal training videos.

    Under `torch.distributed` (one process per GPU) the videos are dealt round-robin to the ranks,
    every rank runs `test_step` on its own videos and the per-frame outputs are exchanged once at the
    end, so every rank returns the metrics of the WHOLE test set (the reference's test path is
    `@rank_zero_only`, anomaly_clip_module.py:458,500: one GPU does all the work)."""
    import torch.distributed as dist
    from torch.utils.data import DataLoader, Subset

    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    sidecar = None
    if checkpoint is not None:
        load_checkpoint(module, checkpoint)
        # the centroid side-car lives in the run directory of the checkpoint, as in the reference's
        # on_test_start (anomaly_clip_module.py:407-417) -- never in the current directory
        sidecar = _sidecar_path(module, checkpoint)
    module.eval()
    datamodule.setup("test")
    if ncentroid is not None:
        module.ncentroid = ncentroid
    if module.ncentroid is None and sidecar is not None and sidecar.is_file():
        module.ncentroid = torch.load(sidecar, map_location="cpu")
    if module.ncentroid is None:
        loader = datamodule.train_dataloader_test_mode()
        if world > 1:                               # each rank streams its share; sharded_mean combines
            mine = list(range(rank, len(loader.dataset), world))
            loader = DataLoader(Subset(loader.dataset, mine), batch_size=loader.batch_size,
                                num_workers=loader.num_workers, pin_memory=loader.pin_memory)
        module.ncentroid = module.compute_ncentroid(loader, datamodule.hparams.load_from_features)
        if sidecar is not None and rank == 0:       # :444-445
            sidecar.parent.mkdir(parents=True, exist_ok=True)
            torch.save(module.ncentroid.cpu(), sidecar)
    dev = module._device
    if module.ncentroid.device != dev:              # once, not per video
        module.ncentroid = module.ncentroid.to(dev)
    loader = datamodule.test_dataloader()
    if world == 1:
        for i, batch in enumerate(loader):
            module.test_step(batch, i)
        _raise_on_saturation(dev)
        return module.finish_test_epoch()

    mine = list(range(rank, len(loader.dataset), world))
    shard = DataLoader(Subset(loader.dataset, mine), batch_size=loader.batch_size,
                       num_workers=loader.num_workers, pin_memory=loader.pin_memory)
    for i, batch in enumerate(shard):
        module.test_step(batch, i)
    # one exchange of the per-video outputs; then the lists are put back in dataset order so that
    # every rank computes exactly what a single process would
    local = list(zip(mine, module.labels, module.abnormal_scores, module.class_probs))
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    merged = sorted((item for part in gathered for item in part), key=lambda t: t[0])
    module.labels = [t[1] for t in merged]
    module.abnormal_scores = [t[2] for t in merged]
    module.class_probs = [t[3] for t in merged]
    _raise_on_saturation(dev)
    return module.finish_test_epoch()


# Keys a reference checkpoint may lack without consequence for the test path: the text tower is
# only needed when the module was built with it, `logit_scale` and BatchNorm's batch counter are
# unused at inference (selector_model.py:22,30).
_OPTIONAL_KEYS = ("net.selector_model.logit_scale", "net.selector_model.bn_layer.num_batches_tracked")


def load_checkpoint(module, checkpoint: str) -> None:
    """Load a Lightning `.ckpt` (or a plain state_dict file) into the module and REFUSE a partial
    load: the mirror initialises both towers randomly (the reference calls `clip.load`), so a
    key-name mismatch or a features-only checkpoint on a raw-frame configuration would otherwise
    evaluate random weights and print plausible metrics."""
    from ._lib import AclipError
    state = torch.load(checkpoint, map_location="cpu", weights_only=False)
    state = state.get("state_dict", state)
    missing, unexpected = module.load_state_dict(state, strict=False)
    net = module.net
    needed = []
    for k in missing:
        if k in _OPTIONAL_KEYS:
            continue
        if k.startswith("net.image_encoder.") and getattr(net, "load_from_features", False):
            continue                                     # the encoder never runs on pre-extracted features
        if k.startswith(("net.text_encoder.", "net.prompt_learner.", "net.token_embedding.")) and \
                getattr(net, "_text_key", None) == "explicit":
            continue                                     # text features were supplied explicitly
        needed.append(k)
    if needed or unexpected:
        raise AclipError(
            f"checkpoint {checkpoint} does not match the module: {len(needed)} required key(s) missing "
            f"(e.g. {needed[:3]}), {len(unexpected)} unexpected (e.g. {list(unexpected)[:3]})")


def _sidecar_path(module, checkpoint: str):
    import os
    from pathlib import Path
    run = os.path.normpath(Path(checkpoint).resolve().parent).split(os.path.sep)[-1]
    root = os.environ.get("ACLIP_RUNS_DIR")
    base = Path(root) / run if root else Path(checkpoint).resolve().parent
    return base / "ncentroid.pt"


def _raise_on_saturation(dev) -> None:
    """The fp16-based operand encodings clamp out-of-range activations; never report metrics from a
    run where that happened."""
    from . import _lib
    if dev.type != "cuda":
        return
    with torch.cuda.device(dev):
        n = _lib.saturation_count(reset=True)
    if n:
        raise _lib.AclipError(f"{n} activations left the fp16 range of the operand encoding during this "
                              "evaluation: rerun with passes=3 (split-bf16 operands)")


def main(argv=None) -> Dict[str, float]:
    """`python -m anomalyclip_b200.eval --configs <AnomalyCLIP>/configs --data ucfcrime --model
    anomaly_clip_ucfcrime --ckpt last.ckpt [key.path=value ...]`: the reference's `python src/eval.py
    data=... model=... ckpt_path=...` without Hydra / Lightning (config.py composes the same files)."""
    import argparse
    import json

    from .config import instantiate, load_eval_config

    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument("--configs", required=True, help="the reference's configs/ directory")
    ap.add_argument("--data", default="ucfcrime")
    ap.add_argument("--model", default="anomaly_clip_ucfcrime")
    ap.add_argument("--ckpt", default=None, help="Lightning checkpoint (state_dict with the reference's key names)")
    ap.add_argument("--random-weights", action="store_true",
                    help="evaluate the randomly initialised module (plumbing checks only); without this "
                         "flag --ckpt is required")
    ap.add_argument("--device", default="cuda")
    ap.add_argument("overrides", nargs="*", help="key.path=value, e.g. data.frames_root=/data/UCF/features")
    args = ap.parse_args(argv)
    if args.ckpt is None and not args.random_weights:
        ap.error("--ckpt is required (or pass --random-weights for a plumbing run on random weights)")
    cfg = load_eval_config(args.configs, args.data, args.model, args.overrides)
    datamodule, module = instantiate(cfg["data"]), instantiate(cfg["model"])
    module.to(args.device)
    metrics = evaluate(module, datamodule, checkpoint=args.ckpt)
    print(json.dumps(metrics, indent=2, sort_keys=True))
    return metrics


if __name__ == "__main__":
    main()
