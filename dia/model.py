"""Drop-in alias: ``dia.model`` of the reference maps onto ``dia_tts_prune_b200.model``."""
from dia_tts_prune_b200.model import *  # noqa: F401,F403
from dia_tts_prune_b200 import model as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
