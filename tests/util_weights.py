"""Re-export of the seeded synthetic weight generators for the tests."""
from anomalyclip_b200.synthetic import (PRESETS, PathConfig, make_features, make_frames_u8,  # noqa: F401
                                        make_ncentroid, make_state_dict, make_temporal_weights,
                                        make_text_features, make_vit_weights, normalise_frames)
