"""Codebook delay pattern and its inverse (drop-in for the reference's ``dia/audio.py``).

Same five functions and signatures: ``build_delay_indices`` (dia/audio.py:6-41),
``apply_audio_delay`` (:44-85), ``build_revert_indices`` (:88-122), ``revert_audio_delay``
(:125-163), ``decode`` (:166-185).

The reference materialises ``[B*T*C, 3]`` int64 gather indices and indexes through them; here
all four token-grid functions run as coalesced integer gather kernels (``csrc/aux_kernels.cu``)
that compute ``t -/+ delay[c]`` in registers.  The index tensors are still produced (same
dtypes and values) because they are part of the interface, but ``apply``/``revert`` only read
the per-channel delays back out of ``t_idx``.  Results are bit-exact with the reference.
"""

from __future__ import annotations

import ctypes as C
import typing as tp

import torch

from . import _lib


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("dia_tts_prune_b200.audio needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _delays(delay_pattern, C_) -> C.Array:
    if len(delay_pattern) != C_:
        raise ValueError(f"delay_pattern has {len(delay_pattern)} entries for {C_} channels")
    if C_ > _lib.MAX_CHANNELS:
        raise NotImplementedError(f"at most {_lib.MAX_CHANNELS} channels")
    return (C.c_int32 * C_)(*[int(d) for d in delay_pattern])


def _stream(dev) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def build_delay_indices(B: int, T: int, C_: int, delay_pattern: tp.List[int]) -> tp.Tuple[torch.Tensor, torch.Tensor]:
    """(t_idx int32 [B,T,C] = t - delay[c], indices int64 [B*T*C,3] = (b, clamp(t_idx,0,T-1), c))."""
    dev = _device()
    t_idx = torch.empty((B, T, C_), dtype=torch.int32, device=dev)
    idx = torch.empty((B * T * C_, 3), dtype=torch.int64, device=dev)
    _lib.check(_lib.load().dia_b200_build_delay_indices(C.c_void_p(t_idx.data_ptr()), C.c_void_p(idx.data_ptr()),
                                                        B, T, C_, _delays(delay_pattern, C_), _stream(dev)),
               "build_delay_indices")
    return t_idx, idx


def build_revert_indices(B: int, T: int, C_: int, delay_pattern: tp.List[int]) -> tp.Tuple[torch.Tensor, torch.Tensor]:
    """(t_idx int64 [B,T,C] = min(t + delay[c], T-1), indices int64 [B*T*C,3])."""
    dev = _device()
    t_idx = torch.empty((B, T, C_), dtype=torch.int64, device=dev)
    idx = torch.empty((B * T * C_, 3), dtype=torch.int64, device=dev)
    _lib.check(_lib.load().dia_b200_build_revert_indices(C.c_void_p(t_idx.data_ptr()), C.c_void_p(idx.data_ptr()),
                                                         B, T, C_, _delays(delay_pattern, C_), _stream(dev)),
               "build_revert_indices")
    return t_idx, idx


def _as_i32_cuda(x: torch.Tensor) -> torch.Tensor:
    if x.ndim != 3:
        raise ValueError(f"expected a [B, T, C] token grid, got {tuple(x.shape)}")
    return x.to(device=_device(), dtype=torch.int32).contiguous()


def delay_apply(audio_BxTxC: torch.Tensor, delay_pattern, pad_value: int, bos_value: int) -> torch.Tensor:
    """out[b,t,c] = BOS if t < delay[c] else audio[b, t - delay[c], c]   (kernel entry point)."""
    x = _as_i32_cuda(audio_BxTxC)
    B, T, C_ = x.shape
    out = torch.empty_like(x)
    _lib.check(_lib.load().dia_b200_delay_apply_i32(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), B, T, C_,
                                                    _delays(delay_pattern, C_), int(pad_value), int(bos_value),
                                                    _stream(x.device)), "delay_apply")
    return out.to(device=audio_BxTxC.device, dtype=audio_BxTxC.dtype)


def delay_revert(audio_BxTxC: torch.Tensor, delay_pattern, pad_value: int, T_orig: int) -> torch.Tensor:
    """out[b,t,c] = audio[b, min(t + delay[c], T-1), c]   (kernel entry point)."""
    x = _as_i32_cuda(audio_BxTxC)
    B, T, C_ = x.shape
    out = torch.empty_like(x)
    _lib.check(_lib.load().dia_b200_delay_revert_i32(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), B, T, C_,
                                                     _delays(delay_pattern, C_), int(pad_value), int(T_orig),
                                                     _stream(x.device)), "delay_revert")
    return out.to(device=audio_BxTxC.device, dtype=audio_BxTxC.dtype)


def finalize_codes(codes_TxC: torch.Tensor, delay_pattern, pad_value: int, codebook_size: int = 1024) -> torch.Tensor:
    """Token half of ``Dia._generate_output`` (dia/model.py:504-533) in one kernel:
    revert, drop the last max(delay) rows, zero out-of-range codes -> int32 [1, C, T - max(delay)]."""
    x = codes_TxC.to(device=_device(), dtype=torch.int32).contiguous()
    T, C_ = x.shape
    t_out = max(T - max(delay_pattern), 0)
    out = torch.empty((1, C_, t_out), dtype=torch.int32, device=x.device)
    _lib.check(_lib.load().dia_b200_finalize_codes_i32(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), T, C_,
                                                       _delays(delay_pattern, C_), int(pad_value), int(codebook_size),
                                                       _stream(x.device)), "finalize_codes")
    return out


def apply_audio_delay(audio_BxTxC: torch.Tensor, pad_value: int, bos_value: int,
                      precomp: tp.Tuple[torch.Tensor, torch.Tensor]) -> torch.Tensor:
    t_idx, _ = precomp
    if tuple(t_idx.shape) != tuple(audio_BxTxC.shape):
        raise ValueError("precomputed indices do not match the audio grid")
    if audio_BxTxC.numel() == 0:
        return audio_BxTxC.clone()
    delays = (-t_idx[0, 0, :]).tolist()                  # t_idx[b, 0, c] = -delay[c]
    return delay_apply(audio_BxTxC, delays, pad_value, bos_value)


def revert_audio_delay(audio_BxTxC: torch.Tensor, pad_value: int, precomp: tp.Tuple[torch.Tensor, torch.Tensor],
                       T: int) -> torch.Tensor:
    t_idx, _ = precomp
    if tuple(t_idx.shape) != tuple(audio_BxTxC.shape):
        raise ValueError("precomputed indices do not match the audio grid")
    if audio_BxTxC.numel() == 0:
        return audio_BxTxC.clone()
    delays = t_idx[0, 0, :].tolist()                     # min(delay[c], T-1): equivalent under the clamp
    return delay_revert(audio_BxTxC, delays, pad_value, T)


@torch.no_grad()
@torch.inference_mode()
def decode(model, audio_codes):
    """DAC glue (third-party codec, outside the kernel scope): codes [1, C, T] -> waveform."""
    if len(audio_codes) != 1:
        raise ValueError(f"Expected one frame, got {len(audio_codes)}")
    try:
        z = model.quantizer.from_codes(audio_codes)
        return model.decode(z[0])
    except Exception as e:
        print(f"Error in decode method: {str(e)}")
        raise
