"""Python face of the C ABI: one ``DecodeEngine`` per GPU wraps ``dia_b200_engine``.

Pure plumbing: tensors are handed over as ``data_ptr()`` + the current CUDA stream.  All
arithmetic happens in ``csrc/`` kernels; there is no eager / CPU fallback here.
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .config import DiaConfig
from .synthetic import is_dense_kernel


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(t.data_ptr())


def rope_tables(config: DiaConfig, n_pos: int, head_dim: int = 128) -> tuple[torch.Tensor, torch.Tensor]:
    """sin/cos of ``int32 position * inv_freq`` computed with the reference's own CPU expressions
    (RotaryEmbedding, dia/layers.py:126-132,146,161-162) so that the table equals what the fp32
    CPU reference uses, bit for bit."""
    half = head_dim // 2
    fraction = (2.0 * torch.arange(0, half)) / head_dim
    m = config.model
    inv_freq = (1.0 / (m.rope_min_timescale * (m.rope_max_timescale / m.rope_min_timescale) ** fraction)).to(torch.float32)
    pos = torch.arange(n_pos, dtype=torch.int32).unsqueeze(-1)
    freqs = (pos * inv_freq).to(torch.float32)
    return torch.sin(freqs).contiguous(), torch.cos(freqs).contiguous()


def decoder_tensor_names(config: DiaConfig) -> list[str]:
    """Names (relative to the Decoder module) in the order ``dia_b200_load_decoder_weights`` expects."""
    names = [f"embeddings.{c}.weight" for c in range(config.data.channels)]
    for i in range(config.model.decoder.n_layer):
        p = f"layers.{i}."
        names += [p + "pre_sa_norm.weight", p + "pre_ca_norm.weight", p + "pre_mlp_norm.weight",
                  p + "self_attention.q_proj.weight", p + "self_attention.k_proj.weight",
                  p + "self_attention.v_proj.weight", p + "self_attention.o_proj.weight",
                  p + "cross_attention.q_proj.weight", p + "cross_attention.o_proj.weight",
                  p + "mlp.wi_fused.weight", p + "mlp.wo.weight"]
    return names + ["norm.weight", "logits_dense.weight"]


class DecodeEngine:
    def __init__(self, config: DiaConfig, device: torch.device | str | int = "cuda", n_ctas: int = 0,
                 n_hidden: int | None = None, sparse24: bool = False, k_rows: dict | None = None):
        """``n_hidden``: MLP width of the weights that will be loaded, when a structurally pruned checkpoint was
        compacted (``pruning_utils.plan_mlp_compaction``); defaults to the configuration's.  ``sparse24``: every dense
        kernel is 2:4-sparse along its input axis (``pruning_utils.is_2to4``): the engine streams compressed slabs
        (0.5625 of the bytes) and multiplies with ``mma.sp``; ``load_weights`` rejects a model that is not."""
        if not torch.cuda.is_available():
            raise RuntimeError("dia_tts_prune_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.config = config
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        d, dt = config.model.decoder, config.data
        if d.gqa_head_dim != 128 or d.cross_head_dim != 128:
            raise NotImplementedError("kernels are specialised for head_dim 128")
        sh = _lib.Shape()
        self.n_hidden = int(n_hidden) if n_hidden else d.n_hidden
        sh.n_layer, sh.d_model, sh.n_hidden = d.n_layer, d.n_embd, self.n_hidden
        sh.q_heads, sh.kv_heads, sh.cross_heads = d.gqa_query_heads, d.kv_heads, d.cross_query_heads
        sh.channels, sh.vocab = dt.channels, config.model.tgt_vocab_size
        sh.max_audio_len, sh.max_text_len = dt.audio_length, dt.text_length
        sh.eos_value, sh.pad_value, sh.bos_value = dt.audio_eos_value, dt.audio_pad_value, dt.audio_bos_value
        for i, v in enumerate(dt.delay_pattern):
            sh.delay_pattern[i] = v
        sh.norm_eps = config.model.normalization_layer_epsilon
        self.sparse24 = bool(sparse24)
        sh.sparse24 = 1 if sparse24 else 0
        # K-row compaction of a structurally pruned checkpoint: {GEMM family: contraction length} (pruning_utils.plan_row_compaction)
        self.k_rows = {int(k): int(v) for k, v in (k_rows or {}).items()}
        for fam, kk in self.k_rows.items():
            sh.k_rows[fam] = kk
        self._h = C.c_void_p()
        _lib.check(self.lib.dia_b200_engine_create(C.byref(sh), self.device.index, n_ctas, C.byref(self._h)),
                   "engine_create")
        self.C, self.V, self.D, self.L = dt.channels, config.model.tgt_vocab_size, d.n_embd, d.n_layer
        self.n_ctas = self.lib.dia_b200_engine_num_ctas(self._h)
        self.weight_stream_bytes = int(self.lib.dia_b200_engine_weight_stream_bytes(self._h))
        sin, cos = rope_tables(config, max(dt.audio_length, dt.text_length) + 1)
        _lib.check(self.lib.dia_b200_set_rope_table(self._h, _ptr(sin), _ptr(cos), sin.shape[0]), "set_rope_table")
        self._keep = []        # tensors whose memory the engine points into
        self._bound = None
        self.weights_version = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.dia_b200_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights -----------------------------------------------------------------------------
    def load_weights(self, tensors: dict[str, torch.Tensor]) -> None:
        """``tensors``: Decoder-relative name -> CUDA tensor (dense kernels all fp32 or all bf16)."""
        names = decoder_tensor_names(self.config)
        dense = [n for n in names if is_dense_kernel(n)]
        dts = {tensors[n].dtype for n in dense}
        if len(dts) != 1 or next(iter(dts)) not in (torch.float32, torch.bfloat16):
            raise ValueError(f"dense kernels must be uniformly float32 or bfloat16, got {dts}")
        dense_dtype = 1 if next(iter(dts)) == torch.bfloat16 else 0
        staged = []
        for n in names:
            t = tensors[n].detach()
            if n not in dense:
                t = t.to(torch.float32)
            t = t.to(self.device).contiguous()
            staged.append(t)
        arr = (C.c_void_p * len(staged))(*[t.data_ptr() for t in staged])
        _lib.check(self.lib.dia_b200_load_decoder_weights(self._h, arr, len(staged), dense_dtype, _stream(self.device)),
                   "load_decoder_weights")
        torch.cuda.current_stream(self.device).synchronize()    # sources may be freed after this

    def set_row_maps(self, maps: dict[int, torch.Tensor]) -> None:
        """``{GEMM family: int32 [layers, K]}`` from ``pruning_utils.compact_rows`` (position of every input element in
        the compacted contraction, -1 = dropped)."""
        for fam, m in maps.items():
            m = m.to(dtype=torch.int32, device="cpu").contiguous()
            _lib.check(self.lib.dia_b200_set_row_map(self._h, int(fam), C.cast(m.data_ptr(), C.POINTER(C.c_int32)),
                                                     _stream(self.device)), "set_row_map")

    # ---- per-utterance binding -------------------------------------------------------------------
    def bind(self, self_caches, cross_caches, text_len: int) -> None:
        L = self.L
        if len(self_caches) != L or len(cross_caches) != L:
            raise ValueError("cache lists must have one entry per decoder layer")
        keep = []
        ptrs = [[], [], [], []]
        d = self.config.model.decoder
        for sc, cc in zip(self_caches, cross_caches):
            for t, shape in ((sc.k, (2, d.kv_heads, self.config.data.audio_length, 128)), (sc.v, None)):
                if t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device:
                    raise ValueError("self KV cache tensors must be contiguous float32 CUDA tensors")
                if shape is not None and tuple(t.shape) != shape:
                    raise ValueError(f"self KV cache shape {tuple(t.shape)} != {shape}")
            ck, cv = cc.k, cc.v
            if ck.dtype != torch.float32 or not ck.is_contiguous():
                ck = ck.to(torch.float32).contiguous()
            if cv.dtype != torch.float32 or not cv.is_contiguous():
                cv = cv.to(torch.float32).contiguous()
            if tuple(ck.shape) != (2, d.cross_query_heads, self.config.data.text_length, 128):
                raise ValueError(f"cross KV cache shape {tuple(ck.shape)} unexpected")
            keep += [sc.k, sc.v, ck, cv]
            for lst, t in zip(ptrs, (sc.k, sc.v, ck, cv)):
                lst.append(t.data_ptr())
        arrs = [(C.c_void_p * L)(*p) for p in ptrs]
        _lib.check(self.lib.dia_b200_bind_caches(self._h, arrs[0], arrs[1], arrs[2], arrs[3], L, int(text_len),
                                                 _stream(self.device)), "bind_caches")
        self._keep = keep
        self._bound = tuple(ptrs[0] + ptrs[2]) + (int(text_len),)

    def bound_key(self):
        return self._bound

    # ---- operator boundaries -----------------------------------------------------------------------
    def decode_step(self, tokens_2xC: torch.Tensor, pos: int, slot: int, out: torch.Tensor | None = None) -> torch.Tensor:
        tok = tokens_2xC.to(device=self.device, dtype=torch.int32).contiguous()
        if out is None:
            out = torch.empty((2, self.C, self.V), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.dia_b200_decode_step(self._h, _ptr(tok), int(pos), int(slot), _ptr(out),
                                                 _stream(self.device)), "decode_step")
        return out

    def layer_step(self, layer: int, x_2xD: torch.Tensor, pos: int, slot: int) -> torch.Tensor:
        x = x_2xD.to(device=self.device, dtype=torch.float32).contiguous()
        out = torch.empty_like(x)
        _lib.check(self.lib.dia_b200_decoder_layer_step(self._h, int(layer), _ptr(x), _ptr(out), int(pos), int(slot),
                                                        _stream(self.device)), "decoder_layer_step")
        return out

    def embed_sum(self, tokens_NxC: torch.Tensor) -> torch.Tensor:
        tok = tokens_NxC.to(device=self.device, dtype=torch.int32).contiguous()
        n = tok.shape[0]
        out = torch.empty((n, self.D), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.dia_b200_embed_sum(self._h, _ptr(tok), n, _ptr(out), _stream(self.device)), "embed_sum")
        return out

    def head_sample(self, logits_2xCxV: torch.Tensor, cfg_scale: float, temperature: float, top_p: float, top_k: int,
                    seed: int = 0, draw: int = 0, want_probs: bool = False):
        lg = logits_2xCxV.to(device=self.device, dtype=torch.float32).contiguous()
        pred = torch.empty((self.C,), dtype=torch.int32, device=self.device)
        probs = torch.empty((self.C, self.V), dtype=torch.float32, device=self.device) if want_probs else None
        _lib.check(self.lib.dia_b200_head_sample(self._h, _ptr(lg), float(cfg_scale), float(temperature), float(top_p),
                                                 int(top_k or 0), int(seed) & (2 ** 64 - 1), int(draw), _ptr(pred),
                                                 _ptr(probs) if want_probs else None, _stream(self.device)),
                   "head_sample")
        return (pred, probs) if want_probs else pred

    # ---- the device-resident generate loop ---------------------------------------------------------------
    def generate_begin(self, grid: torch.Tensor, prefill_step: int, first_slot: int, max_tokens: int, cfg_scale: float,
                       temperature: float, top_p: float, top_k: int, seed: int) -> None:
        if grid.dtype != torch.int32 or not grid.is_contiguous() or grid.device != self.device:
            raise ValueError("token grid must be a contiguous int32 CUDA tensor")
        gp = _lib.GenParams(float(cfg_scale), float(temperature), float(top_p), int(top_k or 0), int(max_tokens),
                            int(prefill_step), int(first_slot), 0, int(seed) & (2 ** 64 - 1))
        _lib.check(self.lib.dia_b200_generate_begin(self._h, _ptr(grid), C.byref(gp), _stream(self.device)),
                   "generate_begin")
        self._keep.append(grid)

    def generate_steps(self, n_steps: int) -> None:
        _lib.check(self.lib.dia_b200_generate_steps(self._h, int(n_steps), _stream(self.device)), "generate_steps")

    def status(self) -> _lib.GenStatus:
        st = _lib.GenStatus()
        _lib.check(self.lib.dia_b200_generate_status(self._h, C.byref(st), _stream(self.device)), "generate_status")
        if st.device_error:
            raise RuntimeError(f"device-side watchdog / state error code {st.device_error}")
        return st

    # ---- debugging hooks used by the parity tests -----------------------------------------------------------
    def run_stages(self, tokens_2xC, stage_begin: int, stage_end: int, pos: int, slot: int, cooperative: bool = True):
        tok = None if tokens_2xC is None else tokens_2xC.to(device=self.device, dtype=torch.int32).contiguous()
        _lib.check(self.lib.dia_b200_debug_run_stages(self._h, _ptr(tok) if tok is not None else None, stage_begin,
                                                      stage_end, pos, slot, 1 if cooperative else 0,
                                                      _stream(self.device)), "debug_run_stages")

    def read_buffer(self, which: int) -> torch.Tensor:
        if which == _lib.BUF_LOGITS:
            out = torch.empty((2, self.C, self.V), dtype=torch.float32)
        elif which == _lib.BUF_PRED:
            out = torch.empty((self.C,), dtype=torch.int32)
        elif which == _lib.BUF_X:
            out = torch.empty((self.D, 2), dtype=torch.float32)
        else:
            raise ValueError(f"unknown buffer id {which}")
        _lib.check(self.lib.dia_b200_debug_read(self._h, which, _ptr(out), out.numel() * out.element_size(),
                                                _stream(self.device)), "debug_read")
        return out.t().contiguous() if which == _lib.BUF_X else out                       # x -> [2, D]

    def last_device_error(self, full: bool = False):
        """[code, block, thread, info, 4 detail words] a kernel watchdog left in pinned host memory (works after
        a failed launch).  full=True: also {(block, warp): (site, info)} for every warp that was waiting."""
        n = 16 + 2 * 12 * self.n_ctas
        out = (C.c_int32 * n)()
        self.lib.dia_b200_debug_last_device_error(self._h, out, n)
        head = list(out[:8])
        if not full:
            return head
        where = {}
        for b in range(self.n_ctas):
            for w in range(12):
                s, i = out[16 + 2 * (b * 12 + w)], out[16 + 2 * (b * 12 + w) + 1]
                if s:
                    where[(b, w)] = (s, i)
        return head, where

    def enable_timing(self, on: bool = True, cta: int = 0) -> None:
        _lib.check(self.lib.dia_b200_debug_enable_timing(self._h, 1 + cta if on else 0), "debug_enable_timing")

    def read_timing(self, n_steps: int) -> torch.Tensor:
        """int64 [n_steps, stages, 16] SM-clock stamps of CTA 0, thread 0: 0 stage start, 1 setup done (GEMM) /
        inputs loaded (attention), 2 main loop done, 3 cross-warp reduce done, 4 stage done, 6 first input
        words arrived (GEMM stages only)."""
        S = 8 * self.L + 3
        out = torch.empty((16, S, 16), dtype=torch.int64)
        _lib.check(self.lib.dia_b200_debug_read(self._h, _lib.BUF_TIMING, _ptr(out), out.numel() * 8,
                                                _stream(self.device)), "debug_read")
        return out[:n_steps]

    def read_cta_timing(self) -> torch.Tensor:
        """int64 [stages, n_ctas]: %globaltimer (ns) at which every CTA finished each stage of step 1 of the last
        timed launch (needs a launch of >= 2 steps)."""
        S = 8 * self.L + 3
        out = torch.empty((S, self.n_ctas), dtype=torch.int64)
        _lib.check(self.lib.dia_b200_debug_read(self._h, _lib.BUF_CTA_TIMING, _ptr(out), out.numel() * 8,
                                                _stream(self.device)), "debug_read")
        return out


class BatchDecodeEngine:
    """N utterances per launch (``dia_b200_engine_create_batched``): 2N batch rows share one pass over the weights on the
    tcgen05 step kernel.  Each utterance binds its own ``KVCache`` lists, token grid and text length."""

    def __init__(self, config: DiaConfig, device: torch.device | str | int = "cuda", max_utterances: int = _lib.MAX_UTTERANCES,
                 n_ctas: int = 0, n_hidden: int | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("dia_tts_prune_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.config = config
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        d, dt = config.model.decoder, config.data
        sh = _lib.Shape()
        self.n_hidden = int(n_hidden) if n_hidden else d.n_hidden
        sh.n_layer, sh.d_model, sh.n_hidden = d.n_layer, d.n_embd, self.n_hidden
        sh.q_heads, sh.kv_heads, sh.cross_heads = d.gqa_query_heads, d.kv_heads, d.cross_query_heads
        sh.channels, sh.vocab = dt.channels, config.model.tgt_vocab_size
        sh.max_audio_len, sh.max_text_len = dt.audio_length, dt.text_length
        sh.eos_value, sh.pad_value, sh.bos_value = dt.audio_eos_value, dt.audio_pad_value, dt.audio_bos_value
        for i, v in enumerate(dt.delay_pattern):
            sh.delay_pattern[i] = v
        sh.norm_eps = config.model.normalization_layer_epsilon
        sh.sparse24 = 0
        self.sparse24 = False
        self._h = C.c_void_p()
        _lib.check(self.lib.dia_b200_engine_create_batched(C.byref(sh), self.device.index, n_ctas, int(max_utterances),
                                                           C.byref(self._h)), "engine_create_batched")
        self.max_utterances = int(max_utterances)
        self.C, self.V, self.D, self.L = dt.channels, config.model.tgt_vocab_size, d.n_embd, d.n_layer
        self.n_ctas = self.lib.dia_b200_engine_num_ctas(self._h)
        self.weight_stream_bytes = int(self.lib.dia_b200_engine_weight_stream_bytes(self._h))
        sin, cos = rope_tables(config, max(dt.audio_length, dt.text_length) + 1)
        _lib.check(self.lib.dia_b200_set_rope_table(self._h, _ptr(sin), _ptr(cos), sin.shape[0]), "set_rope_table")
        self._keep: dict = {}
        self.n_active = 0

    close = DecodeEngine.close
    __del__ = DecodeEngine.__del__
    load_weights = DecodeEngine.load_weights
    last_device_error = DecodeEngine.last_device_error

    def bind(self, utterance: int, self_caches, cross_caches, text_len: int) -> None:
        L = self.L
        d = self.config.model.decoder
        keep, ptrs = [], [[], [], [], []]
        for sc, cc in zip(self_caches, cross_caches):
            for t in (sc.k, sc.v):
                if t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device or \
                        tuple(t.shape) != (2, d.kv_heads, self.config.data.audio_length, 128):
                    raise ValueError("self KV cache tensors must be contiguous float32 [2, kv_heads, audio_length, 128] CUDA tensors")
            ck, cv = cc.k, cc.v
            if ck.dtype != torch.float32 or not ck.is_contiguous():
                ck = ck.to(torch.float32).contiguous()
            if cv.dtype != torch.float32 or not cv.is_contiguous():
                cv = cv.to(torch.float32).contiguous()
            if tuple(ck.shape) != (2, d.cross_query_heads, self.config.data.text_length, 128):
                raise ValueError(f"cross KV cache shape {tuple(ck.shape)} unexpected")
            keep += [sc.k, sc.v, ck, cv]
            for lst, t in zip(ptrs, (sc.k, sc.v, ck, cv)):
                lst.append(t.data_ptr())
        arrs = [(C.c_void_p * L)(*p) for p in ptrs]
        _lib.check(self.lib.dia_b200_batch_bind_caches(self._h, int(utterance), arrs[0], arrs[1], arrs[2], arrs[3], L,
                                                       int(text_len), _stream(self.device)), "batch_bind_caches")
        self._keep[int(utterance)] = keep

    enable_timing = DecodeEngine.enable_timing

    def read_profile(self) -> dict:
        """SM-clock totals of CTA 0 over the last launch (after ``enable_timing(True)``): where the MMA thread and math
        thread 0 spent their time."""
        out = torch.empty((32,), dtype=torch.int64)
        _lib.check(self.lib.dia_b200_debug_read(self._h, _lib.BUF_TIMING, _ptr(out), 32 * 8, _stream(self.device)), "debug_read")
        v = out.tolist()
        names = ["issuer4_total", "issuer4_wait_bfull", "issuer4_wait_ring", "", "act_wait_ready", "act_wait_bempty", "", "",
                 "math_total", "math_wait_ready", "math_epilogue_release", "math_rms_gather", "math_wait_acc_full",
                 "math_epilogue", "math_stage_end_barrier", "math_attention", "math_embed_sample", "math_epilogue_tmem_loads"]
        return {n: v[i] for i, n in enumerate(names) if n}

    def decode_step(self, tokens_UxC: torch.Tensor, pos: list[int], slot: list[int]) -> torch.Tensor:
        """``Decoder.decode_step`` for U utterances: int32 [U, C] -> float32 logits [2U, C, V]."""
        tok = tokens_UxC.to(device=self.device, dtype=torch.int32).contiguous()
        U = tok.shape[0]
        out = torch.empty((2 * U, self.C, self.V), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.dia_b200_batch_decode_step(self._h, U, _ptr(tok), (C.c_int32 * U)(*[int(x) for x in pos]),
                                                       (C.c_int32 * U)(*[int(x) for x in slot]), _ptr(out),
                                                       _stream(self.device)), "batch_decode_step")
        return out

    def generate_begin(self, grids: list[torch.Tensor], prefill_steps: list[int], first_slots: list[int], max_tokens: int,
                       cfg_scale: float, temperature: float, top_p: float, top_k: int | None, seeds: list[int]) -> None:
        U = len(grids)
        for g in grids:
            if g.dtype != torch.int32 or not g.is_contiguous() or g.device != self.device:
                raise ValueError("token grids must be contiguous int32 CUDA tensors")
        gps = (_lib.GenParams * U)(*[_lib.GenParams(float(cfg_scale), float(temperature), float(top_p), int(top_k or 0),
                                                    int(max_tokens), int(prefill_steps[u]), int(first_slots[u]), 0,
                                                    int(seeds[u]) & (2 ** 64 - 1)) for u in range(U)])
        ptrs = (C.c_void_p * U)(*[g.data_ptr() for g in grids])
        _lib.check(self.lib.dia_b200_batch_generate_begin(self._h, U, ptrs, gps, _stream(self.device)), "batch_generate_begin")
        self._keep["grids"] = list(grids)
        self.n_active = U

    def generate_steps(self, n_steps: int) -> None:
        _lib.check(self.lib.dia_b200_batch_generate_steps(self._h, int(n_steps), _stream(self.device)), "batch_generate_steps")

    def status(self) -> list:
        st = (_lib.GenStatus * self.n_active)()
        _lib.check(self.lib.dia_b200_batch_generate_status(self._h, st, _stream(self.device)), "batch_generate_status")
        if st[0].device_error:
            raise RuntimeError(f"device-side watchdog / state error code {st[0].device_error}: {self.last_device_error()}")
        return list(st)


def launch_count() -> int:
    return int(_lib.load().dia_b200_launch_count())


# ---- tcgen05 GEMM for dense layers with more than one row (encoder, cross-KV precompute, prefill) ---------------------
def dense_prepare_weight(w_KxN: torch.Tensor) -> torch.Tensor:
    """K-major bf16 copy ``[N, K]`` of a DenseGeneral kernel viewed as ``[K, N]`` (float32 or bfloat16, CUDA)."""
    lib = _lib.load()
    if not w_KxN.is_cuda or w_KxN.dim() != 2 or w_KxN.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("expected a 2-D float32 / bfloat16 CUDA tensor")
    w = w_KxN.contiguous()
    K, N = w.shape
    wt = torch.empty((N, K), dtype=torch.bfloat16, device=w.device)
    _lib.check(lib.dia_b200_dense_prepare_weight(_ptr(w), 1 if w.dtype == torch.bfloat16 else 0, _ptr(wt), K, N,
                                                 _stream(w.device)), "dense_prepare_weight")
    return wt


def dense_supported(M: int, N: int, K: int) -> bool:
    return M > 0 and N > 0 and N % 4 == 0 and K >= 64 and K % 64 == 0


def dense_forward(x_MxK: torch.Tensor, wt_NxK: torch.Tensor, norm_weight: torch.Tensor | None = None,
                  eps: float = 1e-5, residual: torch.Tensor | None = None) -> torch.Tensor:
    """``residual + rmsnorm(x; norm_weight) @ W`` in float32 accuracy on the tcgen05 tensor cores (both extras
    optional); ``wt_NxK`` from :func:`dense_prepare_weight`.  The result aliases ``residual`` when one is given."""
    lib = _lib.load()
    x = x_MxK.to(torch.float32).contiguous()
    M, K = x.shape
    N = wt_NxK.shape[0]
    if residual is not None:
        if residual.dtype != torch.float32 or not residual.is_contiguous() or tuple(residual.shape) != (M, N):
            raise ValueError("residual must be a contiguous float32 [M, N] tensor")
        y = residual
    else:
        y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    nw = None
    if norm_weight is not None:
        nw = norm_weight.detach().to(torch.float32).contiguous()
        if nw.numel() != K:
            raise ValueError("norm weight length must equal K")
    ws = torch.empty((int(lib.dia_b200_dense_workspace_bytes(M, K)),), dtype=torch.uint8, device=x.device)
    _lib.check(lib.dia_b200_dense_forward_fused(_ptr(x), _ptr(nw) if nw is not None else None, float(eps), _ptr(wt_NxK),
                                                _ptr(residual) if residual is not None else None, _ptr(y), _ptr(ws),
                                                M, N, K, _stream(x.device)), "dense_forward")
    return y


# ---- the rest of the T > 1 passes: attention, RoPE + cache layout, norms, gate, embedding (csrc/prefill_kernels.cu) -----
_ROPE_DEV: dict = {}


def rope_tables_device(config: DiaConfig, device: torch.device) -> tuple[torch.Tensor, torch.Tensor]:
    """The host-made sin / cos tables (:func:`rope_tables`) on ``device``, one copy per (config, device)."""
    n_pos = max(config.data.audio_length, config.data.text_length) + 1
    m = config.model
    key = (str(device), n_pos, m.rope_min_timescale, m.rope_max_timescale)
    if key not in _ROPE_DEV:
        sin, cos = rope_tables(config, n_pos)
        _ROPE_DEV[key] = (sin.to(device), cos.to(device))
    return _ROPE_DEV[key]


def positions_i32(n: int, batch: int, device: torch.device, start: int = 0) -> torch.Tensor:
    """int32 [batch * n] positions start .. start + n - 1 per batch row, built on the host (no device kernel)."""
    return torch.arange(start, start + n, dtype=torch.int32).repeat(batch).to(device)


def rope_rows(src_BTxHd: torch.Tensor, pos: torch.Tensor | None, B: int, T: int, H: int, tables, rotate: bool = True,
              cache: torch.Tensor | None = None, cache_t0: int = 0) -> torch.Tensor:
    """RotaryEmbedding on ``[B*T, H*128]`` rows; in place, or scattered into ``cache [B, H, Tmax, 128]`` at t0."""
    lib = _lib.load()
    sin, cos = tables
    dst = src_BTxHd if cache is None else cache
    _lib.check(lib.dia_b200_rope_rows(_ptr(src_BTxHd), _ptr(dst), _ptr(sin), _ptr(cos),
                                      _ptr(pos) if pos is not None else None, B, T, H, 1 if rotate else 0,
                                      0 if cache is None else 1, 0 if cache is None else cache.shape[2], int(cache_t0),
                                      sin.shape[0], _stream(src_BTxHd.device)), "rope_rows")
    return dst


def attention_rows(q_BTHd: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, Tk: int, mode: int,
                   n_valid: list[int] | None) -> torch.Tensor:
    """fp32 attention for T > 1 query rows; ``k_cache`` / ``v_cache`` are ``[B, Hkv, Tmax, 128]``."""
    lib = _lib.load()
    B, Tq, Hq, d = q_BTHd.shape
    assert d == 128 and k_cache.shape[0] == B and k_cache.is_contiguous() and v_cache.is_contiguous()
    out = torch.empty_like(q_BTHd)
    nv = (C.c_int32 * B)(*[int(x) for x in n_valid]) if n_valid is not None else None
    _lib.check(lib.dia_b200_attention_rows(_ptr(q_BTHd), _ptr(k_cache), _ptr(v_cache), _ptr(out), B, Tq, int(Tk), Hq,
                                           k_cache.shape[1], k_cache.shape[2], int(mode), nv, _stream(q_BTHd.device)),
               "attention_rows")
    return out


def rmsnorm_rows(x_MxD: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    lib = _lib.load()
    x = x_MxD.to(torch.float32).contiguous()
    y = torch.empty_like(x)
    w = weight.detach().to(torch.float32).contiguous()
    _lib.check(lib.dia_b200_rmsnorm_rows(_ptr(x), _ptr(w), float(eps), _ptr(y), x.shape[0], x.shape[1], _stream(x.device)),
               "rmsnorm_rows")
    return y


def silu_mul(gu_Mx2xF: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    M, two, F = gu_Mx2xF.shape
    assert two == 2 and gu_Mx2xF.is_contiguous() and gu_Mx2xF.dtype == torch.float32
    h = torch.empty((M, F), dtype=torch.float32, device=gu_Mx2xF.device)
    _lib.check(lib.dia_b200_silu_mul(_ptr(gu_Mx2xF), _ptr(h), M, F, _stream(h.device)), "silu_mul")
    return h


def embed_rows(table_VxD: torch.Tensor, ids_i32: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    t = table_VxD.detach()
    assert t.dtype == torch.float32 and t.is_contiguous() and ids_i32.dtype == torch.int32 and ids_i32.is_contiguous()
    n = ids_i32.numel()
    out = torch.empty((n, t.shape[1]), dtype=torch.float32, device=t.device)
    _lib.check(lib.dia_b200_embed_rows(_ptr(t), _ptr(ids_i32), _ptr(out), n, t.shape[0], t.shape[1], _stream(t.device)),
               "embed_rows")
    return out
