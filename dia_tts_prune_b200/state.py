"""Per-generation inference state (drop-in for the reference's ``dia/state.py``).

Same class names, constructor arguments and fields as the reference
(``create_attn_mask`` dia/state.py:8-39, ``EncoderInferenceState`` :42-69, ``KVCache`` :72-109,
``DecoderInferenceState`` :112-169, ``DecoderOutput`` :172-208), with the intended behaviour where
the shipped code is broken: ``KVCache.prefill`` returns ``(k, v)`` and
``DecoderOutput.get_tokens_at(step)`` returns the 1-D row.

The decode kernels read and write ``KVCache.k`` / ``KVCache.v`` in place through their data
pointers, so those stay ordinary ``[2, heads, max_len, 128]`` tensors the caller can index.
Self caches are always float32 (a bf16 cache cannot meet the parity bar).
"""

from __future__ import annotations

from dataclasses import dataclass, field

import torch

from .config import DiaConfig


def create_attn_mask(q_padding_mask_1d: torch.Tensor, k_padding_mask_1d: torch.Tensor, device: torch.device,
                     is_causal: bool = False) -> torch.Tensor:
    """bool [B, 1, Tq, Tk]: a query attends a key iff both are real tokens or both are padding;
    optionally lower-triangular."""
    if q_padding_mask_1d.shape[0] != k_padding_mask_1d.shape[0]:
        raise AssertionError("Query and key batch dimensions must match")
    q = q_padding_mask_1d[:, :, None]
    k = k_padding_mask_1d[:, None, :]
    allowed = q == k                       # (real, real) or (pad, pad)
    if is_causal:
        tq, tk = q_padding_mask_1d.shape[1], k_padding_mask_1d.shape[1]
        if tq != tk:
            raise AssertionError("Causal mask requires query and key sequence lengths to be equal")
        allowed = allowed & torch.ones((tq, tk), dtype=torch.bool, device=device).tril()
    return allowed[:, None, :, :]


@dataclass
class EncoderInferenceState:
    max_seq_len: int
    device: torch.device
    positions: torch.Tensor
    padding_mask: torch.Tensor
    attn_mask: torch.Tensor
    # extensions (host-side mirrors, optional): valid text bytes per batch row; scratch shared by the layers
    valid_lens: list | None = None
    extras: dict = field(default_factory=dict)

    @classmethod
    def new(cls, config: DiaConfig, cond_src: torch.Tensor) -> "EncoderInferenceState":
        device = cond_src.device
        n = config.data.text_length
        positions = torch.arange(n, dtype=torch.float32, device=device)[None, :].expand(2, -1)
        padding_mask = (cond_src != config.data.text_pad_value).to(device).expand(2, -1)
        return cls(max_seq_len=n, device=device, positions=positions, padding_mask=padding_mask,
                   attn_mask=create_attn_mask(padding_mask, padding_mask, device, is_causal=False))


class KVCache:
    """Pre-allocated key/value store for one attention layer, batch fixed at 2 (CFG)."""

    def __init__(self, num_heads: int, max_len: int, head_dim: int, dtype: torch.dtype, device: torch.device,
                 k: torch.Tensor | None = None, v: torch.Tensor | None = None):
        shape = (2, num_heads, max_len, head_dim)
        self.k = torch.zeros(shape, dtype=dtype, device=device) if k is None else k
        self.v = torch.zeros(shape, dtype=dtype, device=device) if v is None else v
        self.current_idx = 0

    @classmethod
    def from_kv(cls, k: torch.Tensor, v: torch.Tensor) -> "KVCache":
        return cls(k.shape[1], k.shape[2], k.shape[3], k.dtype, k.device, k=k, v=v)

    def update(self, k: torch.Tensor, v: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """Append one step at ``current_idx`` and return the attended prefix views."""
        i = self.current_idx
        self.k[:, :, i:i + 1, :] = k
        self.v[:, :, i:i + 1, :] = v
        self.current_idx = i + 1
        return self.k[:, :, :i + 1, :], self.v[:, :, :i + 1, :]

    def prefill(self, k: torch.Tensor, v: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """Write slots [0, n) and leave ``current_idx = n - 1`` - the reference's value, which makes
        the first decode step overwrite the last prefilled slot."""
        n = k.shape[2]
        self.k[:, :, :n, :] = k
        self.v[:, :, :n, :] = v
        self.current_idx = n - 1
        return k, v


@dataclass
class DecoderInferenceState:
    device: torch.device
    dtype: torch.dtype
    enc_out: torch.Tensor
    enc_positions: torch.Tensor
    dec_positions: torch.Tensor
    dec_cross_attn_mask: torch.Tensor
    self_attn_cache: list[KVCache]
    cross_attn_cache: list[KVCache]
    # host-side mirrors so that the step path never reads a device tensor back
    step_from: int = 0
    step_to: int = 1
    text_len: int | None = None
    extras: dict = field(default_factory=dict)

    @classmethod
    def new(cls, config: DiaConfig, enc_state: EncoderInferenceState, enc_out: torch.Tensor,
            dec_cross_attn_cache: list[KVCache], compute_dtype: torch.dtype) -> "DecoderInferenceState":
        device = enc_out.device
        dec = config.model.decoder
        tgt_mask = torch.ones((2, 1), dtype=torch.bool, device=device)
        cross_mask = create_attn_mask(tgt_mask, enc_state.padding_mask, device, is_causal=False)
        caches = [KVCache(dec.kv_heads, config.data.audio_length, dec.gqa_head_dim, torch.float32, device)
                  for _ in range(dec.n_layer)]
        return cls(device=device, dtype=compute_dtype, enc_out=enc_out, enc_positions=enc_state.positions,
                   dec_positions=torch.zeros((2, 1), dtype=torch.int32, device=device),
                   dec_cross_attn_mask=cross_mask, self_attn_cache=caches, cross_attn_cache=dec_cross_attn_cache)

    def prepare_step(self, step_from: int, step_to: int | None = None) -> None:
        if step_to is None:
            step_to = step_from + 1
        self.step_from, self.step_to = int(step_from), int(step_to)
        self.dec_positions = torch.arange(step_from, step_to, dtype=torch.int32, device=self.device)[None, :].expand(2, -1)


@dataclass
class DecoderOutput:
    generated_tokens: torch.Tensor
    prefill_step: int

    @classmethod
    def new(cls, config: DiaConfig, device: torch.device) -> "DecoderOutput":
        grid = torch.full((config.data.audio_length, config.data.channels), -1, dtype=torch.int32, device=device)
        return cls(generated_tokens=grid, prefill_step=0)

    def get_tokens_at(self, step_from: int, step_to: int | None = None) -> torch.Tensor:
        if step_to is None:
            return self.generated_tokens[step_from]
        return self.generated_tokens[step_from:step_to, :]

    def update_one(self, dec_out: torch.Tensor, step: int, apply_mask: bool = False) -> None:
        new = dec_out.to(self.generated_tokens.dtype)
        row = self.generated_tokens[step:step + 1, :]
        self.generated_tokens[step:step + 1, :] = torch.where(row == -1, new, row) if apply_mask else new

    def prefill(self, dec_out: torch.Tensor, prefill_step: int) -> None:
        self.generated_tokens[0:dec_out.shape[0], :] = dec_out
        self.prefill_step = prefill_step
