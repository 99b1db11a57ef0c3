"""numpy restatement of the codebook delay / revert gathers (dia/audio.py).

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.

PINNED against (a) the unpatched reference functions imported from
/root/reference by oracle/validate_against_reference.py on random grids, and
(b) the known-answer vector of SURVEY.md Appendix C committed as
tests/golden/delay_known_answer.json.
"""

from __future__ import annotations

import numpy as np


def build_delay_indices(B: int, T: int, C: int, delay_pattern) -> tuple[np.ndarray, np.ndarray]:
    """dia/audio.py:6-41.  t_idx[b,t,c] = t - delay[c] (int32, unclamped);
    indices[B*T*C, 3] = (b, clamp(t_idx, 0, T-1), c) as int64."""
    delay = np.asarray(delay_pattern, dtype=np.int32)
    t_idx = np.broadcast_to(np.arange(T, dtype=np.int32)[None, :, None], (B, T, C)) - delay.reshape(1, 1, C)
    b_idx = np.broadcast_to(np.arange(B, dtype=np.int32).reshape(B, 1, 1), (B, T, C))
    c_idx = np.broadcast_to(np.arange(C, dtype=np.int32).reshape(1, 1, C), (B, T, C))
    t_cl = np.clip(t_idx, 0, T - 1)
    idx = np.stack([b_idx.reshape(-1), t_cl.reshape(-1), c_idx.reshape(-1)], axis=1).astype(np.int64)
    return t_idx.astype(np.int32), idx


def apply_audio_delay(audio: np.ndarray, pad_value: int, bos_value: int, delay_pattern) -> np.ndarray:
    """dia/audio.py:44-85.  out[b,t,c] = BOS if t-delay[c] < 0, PAD if
    t-delay[c] >= T (unreachable), else audio[b, t-delay[c], c].  dtype kept."""
    B, T, C = audio.shape
    t_idx, idx = build_delay_indices(B, T, C, delay_pattern)
    gathered = audio[idx[:, 0], idx[:, 1], idx[:, 2]].reshape(audio.shape)
    out = np.where(t_idx < 0, np.asarray(bos_value, dtype=audio.dtype),
                   np.where(t_idx >= T, np.asarray(pad_value, dtype=audio.dtype), gathered))
    return out.astype(audio.dtype)


def build_revert_indices(B: int, T: int, C: int, delay_pattern) -> tuple[np.ndarray, np.ndarray]:
    """dia/audio.py:88-122.  t_idx[b,t,c] = min(t + delay[c], T-1) (int64)."""
    delay = np.asarray(delay_pattern, dtype=np.int64)
    t_idx = np.minimum(np.broadcast_to(np.arange(T, dtype=np.int64)[None, :, None], (B, T, C)) + delay.reshape(1, 1, C),
                       T - 1)
    b_idx = np.broadcast_to(np.arange(B, dtype=np.int64).reshape(B, 1, 1), (B, T, C))
    c_idx = np.broadcast_to(np.arange(C, dtype=np.int64).reshape(1, 1, C), (B, T, C))
    idx = np.stack([b_idx.reshape(-1), t_idx.reshape(-1), c_idx.reshape(-1)], axis=1).astype(np.int64)
    return t_idx.astype(np.int64), idx


def revert_audio_delay(audio: np.ndarray, pad_value: int, delay_pattern, T: int) -> np.ndarray:
    """dia/audio.py:125-163.  out[b,t,c] = audio[b, min(t+delay[c], Tin-1), c],
    replaced by PAD where the (already clamped) index is >= T."""
    B, Tin, C = audio.shape
    t_idx, idx = build_revert_indices(B, Tin, C, delay_pattern)
    gathered = audio[idx[:, 0], idx[:, 1], idx[:, 2]].reshape(audio.shape)
    return np.where(t_idx >= T, np.asarray(pad_value, dtype=audio.dtype), gathered).astype(audio.dtype)
