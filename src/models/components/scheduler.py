from anomalyclip_b200.training_stubs import WarmupCosineAnnealingLR  # noqa: F401  (configs/model/*.yaml `scheduler._target_`)
