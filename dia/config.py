"""Drop-in alias: ``dia.config`` of the reference maps onto ``dia_tts_prune_b200.config``."""
from dia_tts_prune_b200.config import *  # noqa: F401,F403
from dia_tts_prune_b200 import config as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
