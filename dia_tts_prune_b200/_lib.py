"""ctypes bindings of the C ABI in ``include/dia_b200.h``.

The product path has no fallback: if ``libdia_b200.so`` is missing (and cannot be
built) or a call fails, a ``RuntimeError`` is raised.
"""

from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

from . import build as _build

MAX_CHANNELS = 16
MAX_UTTERANCES = 8

# enum dia_b200_buffer
BUF_X, BUF_LOGITS, BUF_PRED, BUF_TIMING, BUF_CTA_TIMING = 0, 6, 7, 8, 9
E_OK, E_INVAL, E_CUDA, E_NOMEM, E_STATE, E_UNSUPPORTED = 0, -1, -2, -3, -4, -5


class Shape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_layer", "d_model", "n_hidden", "q_heads", "kv_heads", "cross_heads",
                                         "channels", "vocab", "max_audio_len", "max_text_len", "eos_value",
                                         "pad_value", "bos_value")] + \
               [("delay_pattern", C.c_int32 * MAX_CHANNELS), ("norm_eps", C.c_float), ("sparse24", C.c_int32),
                ("k_rows", C.c_int32 * 7)]


class GenParams(C.Structure):
    _fields_ = [("cfg_scale", C.c_float), ("temperature", C.c_float), ("top_p", C.c_float), ("top_k", C.c_int32),
                ("max_tokens", C.c_int32), ("prefill_step", C.c_int32), ("first_slot", C.c_int32),
                ("reserved", C.c_int32), ("seed", C.c_uint64)]


class GenStatus(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("dec_step", "finished", "eos_detected", "eos_countdown", "bos_countdown",
                                         "steps_run", "device_error", "reserved")]


_vp, _i, _i32p, _fp = C.c_void_p, C.c_int, C.c_void_p, C.c_void_p
_SIGNATURES = {
    "dia_b200_abi_version": (C.c_int, []),
    "dia_b200_error_string": (C.c_char_p, [_i]),
    "dia_b200_last_cuda_error": (C.c_char_p, []),
    "dia_b200_launch_count": (C.c_int64, []),
    "dia_b200_engine_create": (_i, [C.POINTER(Shape), _i, _i, C.POINTER(_vp)]),
    "dia_b200_engine_destroy": (_i, [_vp]),
    "dia_b200_engine_num_ctas": (_i, [_vp]),
    "dia_b200_engine_weight_stream_bytes": (C.c_int64, [_vp]),
    "dia_b200_load_decoder_weights": (_i, [_vp, C.POINTER(_vp), _i, _i, _vp]),
    "dia_b200_set_rope_table": (_i, [_vp, _vp, _vp, _i]),
    "dia_b200_set_row_map": (_i, [_vp, _i, C.POINTER(C.c_int32), _vp]),
    "dia_b200_bind_caches": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _i, _i, _vp]),
    "dia_b200_decode_step": (_i, [_vp, _i32p, _i, _i, _fp, _vp]),
    "dia_b200_decoder_layer_step": (_i, [_vp, _i, _fp, _fp, _i, _i, _vp]),
    "dia_b200_embed_sum": (_i, [_vp, _i32p, _i, _fp, _vp]),
    "dia_b200_head_sample": (_i, [_vp, _fp, C.c_float, C.c_float, C.c_float, _i, C.c_uint64, C.c_uint64, _i32p, _fp,
                                   _vp]),
    "dia_b200_generate_begin": (_i, [_vp, _i32p, C.POINTER(GenParams), _vp]),
    "dia_b200_generate_steps": (_i, [_vp, _i, _vp]),
    "dia_b200_generate_status": (_i, [_vp, C.POINTER(GenStatus), _vp]),
    "dia_b200_delay_apply_i32": (_i, [_i32p, _i32p, _i, _i, _i, C.POINTER(C.c_int32), C.c_int32, C.c_int32, _vp]),
    "dia_b200_delay_revert_i32": (_i, [_i32p, _i32p, _i, _i, _i, C.POINTER(C.c_int32), C.c_int32, _i, _vp]),
    "dia_b200_finalize_codes_i32": (_i, [_i32p, _i32p, _i, _i, C.POINTER(C.c_int32), C.c_int32, _i, _vp]),
    "dia_b200_build_delay_indices": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(C.c_int32), _vp]),
    "dia_b200_build_revert_indices": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(C.c_int32), _vp]),
    "dia_b200_dense_prepare_weight": (_i, [_vp, _i, _vp, _i, _i, _vp]),
    "dia_b200_dense_workspace_bytes": (C.c_size_t, [_i, _i]),
    "dia_b200_dense_forward": (_i, [_fp, _vp, _fp, _vp, _i, _i, _i, _vp]),
    "dia_b200_engine_create_batched": (_i, [C.POINTER(Shape), _i, _i, _i, C.POINTER(_vp)]),
    "dia_b200_engine_max_utterances": (_i, [_vp]),
    "dia_b200_batch_bind_caches": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _i, _i,
                                        _vp]),
    "dia_b200_batch_decode_step": (_i, [_vp, _i, _i32p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _fp, _vp]),
    "dia_b200_batch_generate_begin": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(GenParams), _vp]),
    "dia_b200_batch_generate_steps": (_i, [_vp, _i, _vp]),
    "dia_b200_batch_generate_status": (_i, [_vp, C.POINTER(GenStatus), _vp]),
    "dia_b200_dense_forward_fused": (_i, [_fp, _fp, C.c_float, _vp, _fp, _fp, _vp, _i, _i, _i, _vp]),
    "dia_b200_attention_rows": (_i, [_fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, C.POINTER(C.c_int32), _vp]),
    "dia_b200_rope_rows": (_i, [_fp, _fp, _fp, _fp, _i32p, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "dia_b200_rmsnorm_rows": (_i, [_fp, _fp, C.c_float, _fp, _i, _i, _vp]),
    "dia_b200_silu_mul": (_i, [_fp, _fp, _i, _i, _vp]),
    "dia_b200_embed_rows": (_i, [_fp, _i32p, _fp, _i, _i, _i, _vp]),
    "dia_b200_debug_run_stages": (_i, [_vp, _i32p, _i, _i, _i, _i, _i, _vp]),
    "dia_b200_debug_enable_timing": (_i, [_vp, _i]),
    "dia_b200_debug_last_device_error": (_i, [_vp, C.POINTER(C.c_int32), _i]),
    "dia_b200_debug_read": (_i, [_vp, _i, _vp, C.c_size_t, _vp]),
}

_lib = None


def header_symbols() -> list[str]:
    """Every function the public header declares (used by the CPU-side export test)."""
    text = (Path(_build.INCLUDE) / "dia_b200.h").read_text()
    return sorted(set(re.findall(r"\b(dia_b200_[a-z0-9_]+)\s*\(", text)))


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building if necessary) the shared library and attach signatures."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    override = os.environ.get("DIA_B200_LIB")            # A/B timing of two builds on one box (tools/ab.sh)
    if override:
        path, build_if_missing = Path(override), False
    if build_if_missing and _build.is_stale():
        try:
            _build.build()
        except Exception as e:  # the prebuilt .so travels to the GPU box; nvcc is there too, but be explicit
            if not path.exists():
                raise RuntimeError(f"libdia_b200.so is missing and could not be built: {e}") from e
    if not path.exists():
        raise RuntimeError(f"{path} not found: run `python -m dia_tts_prune_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(str(path))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.dia_b200_abi_version() != 1:
        raise RuntimeError("libdia_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    """Non-zero ABI codes become exceptions (the reference raises from Python)."""
    if rc == 0:
        return
    lib = load()
    msg = lib.dia_b200_error_string(rc).decode()
    if rc == E_CUDA:
        msg += f" [{lib.dia_b200_last_cuda_error().decode()}]"
    if rc == E_UNSUPPORTED:
        raise NotImplementedError(f"{what}: {msg}")
    if rc == E_INVAL:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: {msg}")
