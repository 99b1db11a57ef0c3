"""Drop-in alias: ``dia.audio`` of the reference maps onto ``dia_tts_prune_b200.audio``."""
from dia_tts_prune_b200.audio import *  # noqa: F401,F403
from dia_tts_prune_b200 import audio as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
