import torch


class Trainer:
    """Trainer.test of pytorch-lightning 1.8 reduced to its call sequence (see the package docstring)."""

    def __init__(self, logger=None, **kwargs):
        self.logger, self.kwargs = logger, kwargs
        self.callback_metrics = {}
        self.datamodule = None
        self.ckpt_path = None
        self.calls = []

    def test(self, model=None, datamodule=None, ckpt_path=None, dataloaders=None):
        self.datamodule, self.ckpt_path = datamodule, ckpt_path
        model.trainer = self
        datamodule.setup("test")
        if ckpt_path is not None:
            ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
            model.load_state_dict(ckpt["state_dict"])
        model.eval()
        with torch.no_grad():
            self.calls.append("on_test_start")
            model.on_test_start()
            outputs = []
            for i, batch in enumerate(datamodule.test_dataloader()):
                outputs.append(model.test_step(batch, i))
            self.calls.append(f"test_step x{len(outputs)}")
            self.calls.append("test_epoch_end")
            result = model.test_epoch_end(outputs)            # 1.8: first the *_epoch_end(outputs) hook ...
            self.calls.append("on_test_epoch_end")
            model.on_test_epoch_end()                          # ... then on_test_epoch_end()
            model.on_test_end()
        self.callback_metrics = dict(result or getattr(model, "last_metrics", {}))
        import json
        import os
        if os.environ.get("ACLIP_TEST_TRAINER_LOG"):
            with open(os.environ["ACLIP_TEST_TRAINER_LOG"], "w") as fp:
                json.dump({"calls": self.calls, "metrics": self.callback_metrics,
                           "model_class": type(model).__module__ + "." + type(model).__name__,
                           "datamodule_class": type(datamodule).__module__ + "." + type(datamodule).__name__,
                           "net_calls": [list(c[1:]) for c in getattr(model.net, "calls", [])]}, fp)
        return [self.callback_metrics]
