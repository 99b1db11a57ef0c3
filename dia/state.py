"""Drop-in alias: ``dia.state`` of the reference maps onto ``dia_tts_prune_b200.state``."""
from dia_tts_prune_b200.state import *  # noqa: F401,F403
from dia_tts_prune_b200 import state as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
