from anomalyclip_b200.training_stubs import ComputeLoss  # noqa: F401  (configs/model/*.yaml `loss._target_`)
