"""``import dia`` drop-in for babybirdprd/dia-tts-prune: the same public names, served by the
B200-native implementation in ``dia_tts_prune_b200``."""
from dia_tts_prune_b200.model import Dia

__all__ = ["Dia"]
