"""Drop-in alias: ``dia.pruning_utils`` of the reference."""
from dia_tts_prune_b200.pruning_utils import *  # noqa: F401,F403
