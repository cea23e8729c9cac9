"""Re-exports the synthetic weight / signal generators (they live in the product package because bench.py and
users without a checkpoint need them too).  TEST INFRASTRUCTURE ONLY (oracle)."""
from targetdiarization_b200.synth import (random_eres2netv2_state_dict, random_state_dict,  # noqa: F401
                                          synthetic_mixture)
