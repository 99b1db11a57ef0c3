"""Drop-in alias: ``dia.layers`` of the reference maps onto ``dia_tts_prune_b200.layers``."""
from dia_tts_prune_b200.layers import *  # noqa: F401,F403
from dia_tts_prune_b200 import layers as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
