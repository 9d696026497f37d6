"""TEST INFRASTRUCTURE - re-export of the deterministic weight / input generators (oa_transformer_b200/synth.py)
so that make_golden.py, the tests and the product benchmarks draw identical tensors."""
from oa_transformer_b200.synth import (dual_encoder_spec, fill_seeded, text_tower_spec,  # noqa: F401
                                       video_tower_spec)
