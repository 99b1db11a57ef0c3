"""Dia configuration schema (drop-in for the reference's ``dia/config.py``).

Same field names, defaults, validation rules and JSON round trip as the
reference (``DataConfig`` dia/config.py:24-60, ``EncoderConfig`` :63-78,
``DecoderConfig`` :81-102, ``ModelConfig`` :105-128, ``DiaConfig`` :134-207),
so a ``config.json`` written by either side loads on the other.  Every kernel
shape in ``csrc/`` is derived from these values.
"""

from __future__ import annotations

import json
import os
from pathlib import Path
from typing import Annotated

from pydantic import BaseModel, BeforeValidator, Field, ValidationError


def _round_up_128(v: int) -> int:
    return -(-int(v) // 128) * 128


_Len128 = Annotated[int, BeforeValidator(_round_up_128)]
_DEFAULT_DELAYS = (0, 8, 9, 10, 11, 12, 13, 14, 15)


class DataConfig(BaseModel, frozen=True):
    # text_length / audio_length are rounded UP to a multiple of 128 before validation
    text_length: _Len128 = Field(gt=0, multiple_of=128)
    audio_length: _Len128 = Field(gt=0, multiple_of=128)
    channels: int = Field(default=9, gt=0, multiple_of=1)
    text_pad_value: int = 0
    audio_eos_value: int = 1024
    audio_pad_value: int = 1025
    audio_bos_value: int = 1026
    delay_pattern: list[Annotated[int, Field(ge=0)]] = Field(default_factory=lambda: list(_DEFAULT_DELAYS))

    def __hash__(self) -> int:
        return hash((self.text_length, self.audio_length, self.channels, self.text_pad_value, self.audio_pad_value,
                     self.audio_bos_value, self.audio_eos_value, tuple(self.delay_pattern)))


class EncoderConfig(BaseModel, frozen=True):
    n_layer: int = Field(gt=0)
    n_embd: int = Field(gt=0)
    n_hidden: int = Field(gt=0)
    n_head: int = Field(gt=0)
    head_dim: int = Field(gt=0)


class DecoderConfig(BaseModel, frozen=True):
    n_layer: int = Field(gt=0)
    n_embd: int = Field(gt=0)
    n_hidden: int = Field(gt=0)
    gqa_query_heads: int = Field(gt=0)
    kv_heads: int = Field(gt=0)
    gqa_head_dim: int = Field(gt=0)
    cross_query_heads: int = Field(gt=0)
    cross_head_dim: int = Field(gt=0)


class ModelConfig(BaseModel, frozen=True):
    encoder: EncoderConfig
    decoder: DecoderConfig
    src_vocab_size: int = Field(default=128, gt=0)
    tgt_vocab_size: int = Field(default=1028, gt=0)
    dropout: float = Field(default=0.0, ge=0.0, lt=1.0)
    normalization_layer_epsilon: float = Field(default=1.0e-5, ge=0.0)
    weight_dtype: str = Field(default="float32")
    rope_min_timescale: int = Field(default=1)
    rope_max_timescale: int = Field(default=10_000)


class DiaConfig(BaseModel, frozen=True):
    version: str = Field(default="1.0")
    model: ModelConfig
    data: DataConfig
    model_type: str = Field(default="dia")
    architectures: list[str] = Field(default_factory=lambda: ["DiaModel"])

    def save(self, path: str | Path) -> None:
        """Write the config as JSON (a ``.json`` suffix is enforced, parents created)."""
        p = Path(path)
        if p.suffix != ".json":
            p = p.with_suffix(".json")
        os.makedirs(p.parent, exist_ok=True)
        p.write_text(self.model_dump_json(indent=2), encoding="utf-8")

    @classmethod
    def load(cls, path: str | Path) -> "DiaConfig | None":
        """Load + validate; returns None when the file does not exist, raises
        ``pydantic.ValidationError`` on schema violations (reference behaviour)."""
        p = Path(path)
        if not p.is_file():
            print(f"Config file not found at: {p}")
            return None
        if p.suffix != ".json":
            print(f"Warning: Config file does not have .json extension: {p}")
        try:
            return cls.model_validate_json(p.read_text(encoding="utf-8"))
        except ValidationError as e:
            print(f"Configuration validation error loading {p}: {e}")
            raise


def dia_1_6b_config() -> DiaConfig:
    """The public nari-labs/Dia-1.6B configuration."""
    return DiaConfig(
        model=ModelConfig(
            encoder=EncoderConfig(n_layer=12, n_embd=1024, n_hidden=4096, n_head=16, head_dim=128),
            decoder=DecoderConfig(n_layer=18, n_embd=2048, n_hidden=8192, gqa_query_heads=16, kv_heads=4,
                                  gqa_head_dim=128, cross_query_heads=16, cross_head_dim=128),
            src_vocab_size=256, tgt_vocab_size=1028),
        data=DataConfig(text_length=1024, audio_length=3072, channels=9),
    )


def tiny_config(n_layer: int = 2, audio_length: int = 256, text_length: int = 128, width: int = 1) -> DiaConfig:
    """A small configuration with the same head geometry (head_dim 128, GQA
    4:1, 9 codebooks of 1028) used by fast parity tests.  ``width`` scales the decoder (width 2: d_model 1024, 8 query
    heads - wide enough for the 512-row granularity of the K-row compaction)."""
    return DiaConfig(
        model=ModelConfig(
            encoder=EncoderConfig(n_layer=2, n_embd=256, n_hidden=512, n_head=2, head_dim=128),
            decoder=DecoderConfig(n_layer=n_layer, n_embd=512 * width, n_hidden=1024 * width, gqa_query_heads=4 * width,
                                  kv_heads=width, gqa_head_dim=128, cross_query_heads=4 * width, cross_head_dim=128),
            src_vocab_size=256, tgt_vocab_size=1028),
        data=DataConfig(text_length=text_length, audio_length=audio_length, channels=9),
    )


def config_to_json(cfg: DiaConfig) -> str:
    return json.dumps(cfg.model_dump())
