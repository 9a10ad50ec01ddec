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
    ncentroid: use this centroid instead of the side-car / the normal training videos."""
    if checkpoint is not None:
        state = torch.load(checkpoint, map_location="cpu")
        module.load_state_dict(state.get("state_dict", state), strict=False)
    module.eval()
    datamodule.setup("test")
    if ncentroid is not None:
        module.ncentroid = ncentroid
    if module.ncentroid is None:
        try:
            module.on_test_start()                      # side-car file, if the run directory has one
        except RuntimeError:
            module.ncentroid = module.compute_ncentroid(datamodule.train_dataloader_test_mode(),
                                                        datamodule.hparams.load_from_features)
    for i, batch in enumerate(datamodule.test_dataloader()):
        module.test_step(batch, i)
    return module.test_epoch_end()
