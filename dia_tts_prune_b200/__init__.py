"""dia_tts_prune_b200 - B200-native (sm_100a) decode path for Dia-1.6B.

A drop-in for the ``dia`` package of babybirdprd/dia-tts-prune on the autoregressive decode path:
``Dia`` (``from_pretrained`` / ``from_local`` / ``generate`` / ``save_audio``), the ``DiaModel`` /
``DecoderLayer`` module boundaries and the ``KVCache`` state keep their reference interfaces, while
every decode step runs in hand-written CUDA kernels behind the C ABI of ``include/dia_b200.h``.
"""

from .config import DiaConfig, dia_1_6b_config, tiny_config  # noqa: F401

__all__ = ["Dia", "DiaConfig", "dia_1_6b_config", "tiny_config"]


def __getattr__(name):
    if name == "Dia":
        from .model import Dia
        return Dia
    raise AttributeError(name)
