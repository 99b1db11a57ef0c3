"""Model modules (drop-in for the reference's ``dia/layers.py``).

Same class names, constructor signatures and state_dict keys / shapes as the reference
(``DenseGeneral`` dia/layers.py:17-66, ``MlpBlock`` :69-105, ``RotaryEmbedding`` :108-173,
``Attention`` :176-346, ``EncoderLayer`` :349-416, ``Encoder`` :419-462, ``DecoderLayer`` :465-584,
``Decoder`` :587-766, ``DiaModel`` :769-807), so reference checkpoints, ``prune`` masks on
``.weight`` and PEFT target names keep working.

What runs where
  * ``Decoder.decode_step`` and ``DecoderLayer.forward`` at T=1 - the autoregressive hot path - run
    entirely in the hand-written sm_100a kernels of ``csrc/`` through the C ABI (``engine.py``).
    There is no eager fallback for them: without a CUDA device they raise.
  * ``Encoder``, ``Decoder.precompute_cross_attn_cache`` and the prompt prefill (``Decoder.forward``)
    are once-per-utterance work.  Every dense layer of these T > 1 passes runs on
    the tcgen05 / TMEM GEMM of ``csrc/gemm_tcgen05.cu`` (``DenseGeneral._forward_tcgen05``); their attention is
    library SDPA in fp32.  The results feed the decode kernels through the KV caches.
  * Pruned checkpoints: ``Decoder.engine()`` drops exactly-dead MLP neurons from the weight stream and streams a
    2:4 model compressed on ``mma.sp`` (``pruning_utils``).

Numerics: dense kernels may be stored bf16 (``compute_dtype``), but activations, residual stream,
norms, softmax and KV caches are always fp32.
"""

from __future__ import annotations

import contextlib
import math
import os
import weakref
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor
from torch.nn import RMSNorm

try:  # hub mixin gives DiaModel.from_pretrained / save_pretrained exactly like the reference
    from huggingface_hub import PyTorchModelHubMixin
except Exception:  # pragma: no cover - hub not installed
    class PyTorchModelHubMixin:  # type: ignore
        def __init_subclass__(cls, **kw):
            super().__init_subclass__()

from .config import DiaConfig
from .state import DecoderInferenceState, EncoderInferenceState, KVCache


@contextlib.contextmanager
def _exact_fp32():
    """fp32 library GEMMs without TF32 (the K/V they produce are re-read by every decode step)."""
    if not torch.cuda.is_available():
        yield
        return
    a, b = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = a, b


class DenseGeneral(nn.Module):
    """Bias-free generalised dense layer with a JAX-layout kernel ``in_shapes + out_features``."""

    def __init__(self, in_shapes: tuple[int, ...], out_features: tuple[int, ...], axis: tuple[int, ...] = (-1,),
                 weight_dtype: torch.dtype | None = None, device: torch.device | None = None):
        super().__init__()
        self.in_shapes, self.out_features, self.axis = in_shapes, out_features, axis
        self.kernel_shape = self.in_shapes + self.out_features
        self.weight = nn.Parameter(torch.empty(self.kernel_shape, dtype=weight_dtype, device=device))

    # the K-major bf16 copy the tcgen05 GEMM reads is keyed by (data_ptr, version): drop it whenever the parameter object
    # is replaced (torch.nn.utils.prune, load_state_dict(assign=True)) or moved (.to / .cpu / .cuda) - a new tensor can
    # land on the very address the old one had
    def __setattr__(self, name, value):
        if name in ("weight", "weight_orig", "weight_mask"):
            self.__dict__.pop("_b200_wt", None)
        super().__setattr__(name, value)

    def __delattr__(self, name):
        if name in ("weight", "weight_orig", "weight_mask"):
            self.__dict__.pop("_b200_wt", None)
        super().__delattr__(name)

    def _apply(self, fn, *a, **kw):
        self.__dict__.pop("_b200_wt", None)
        return super()._apply(fn, *a, **kw)

    def forward(self, inputs: Tensor) -> Tensor:
        """``tensordot(x, W)`` over ``axis`` (dia/layers.py:55-66).  On a CUDA device every call runs on the tcgen05
        GEMM of ``csrc/gemm_tcgen05.cu``; there is no library fallback on that device.  On the CPU (state-dict
        tooling, pruning, the CPU test suite) and under autograd (fine-tuning is outside this path) it is the
        reference's own tensordot."""
        n = len(self.axis)
        x_axes = tuple(a if a >= 0 else inputs.ndim + a for a in self.axis)
        w = self.weight
        if inputs.is_cuda and w.is_cuda and not (torch.is_grad_enabled() and (w.requires_grad or inputs.requires_grad)):
            return self._forward_tcgen05(inputs, x_axes)
        out = torch.tensordot(inputs.to(torch.float32), w.to(torch.float32), dims=(x_axes, tuple(range(n))))
        return out.to(inputs.dtype)

    # ---- sm_100a path (encoder, cross-attention K/V precompute, prompt prefill) ----------------------------------
    def kernel_weight_kmajor(self) -> Tensor:
        """The K-major bf16 copy ``[N, K]`` of the kernel the tcgen05 GEMM reads (made once per weight version)."""
        from . import engine as _engine
        w = self.weight
        K, N = math.prod(self.in_shapes), math.prod(self.out_features)
        key = (w.data_ptr(), w._version, w.dtype, w.device)
        cached = getattr(self, "_b200_wt", None)
        if cached is None or cached[0] != key:
            canonicalize_dense_kernel_(self)
            w = self.weight
            key = (w.data_ptr(), w._version, w.dtype, w.device)
            cached = (key, _engine.dense_prepare_weight(w.detach().reshape(K, N)))
            self._b200_wt = cached
        return cached[1]

    def _forward_tcgen05(self, inputs: Tensor, x_axes: tuple[int, ...], norm_weight: Tensor | None = None,
                         eps: float = 1e-5, residual: Tensor | None = None) -> Tensor:
        """``x @ W`` on the tcgen05 tensor cores with the activations split into three bf16 terms (float32-operand
        accuracy).  ``norm_weight``: RMSNorm of the input rows fused into the split pass; ``residual``: added in the
        epilogue (the result aliases it)."""
        from . import engine as _engine
        n = len(x_axes)
        if x_axes != tuple(range(inputs.ndim - n, inputs.ndim)):
            raise NotImplementedError("DenseGeneral on sm_100a contracts the trailing axes only")
        K, N = math.prod(self.in_shapes), math.prod(self.out_features)
        M = inputs.numel() // K if K else 0
        lead = inputs.shape[: inputs.ndim - n]
        if M == 0:
            return torch.zeros((*lead, *self.out_features), dtype=torch.float32, device=inputs.device)
        if not _engine.dense_supported(M, N, K):
            raise NotImplementedError(f"dense layer [{M} x {K}] . [{K} x {N}] does not fit the tcgen05 tiling "
                                      f"(K must be a multiple of 64, N of 4)")
        y = _engine.dense_forward(inputs.reshape(M, K), self.kernel_weight_kmajor(), norm_weight=norm_weight, eps=eps,
                                  residual=None if residual is None else residual.reshape(M, N))
        return y.reshape(*lead, *self.out_features)


_ROUNDING_WARNED = False


@torch.no_grad()
def canonicalize_dense_kernel_(m: "DenseGeneral") -> bool:
    """The sm_100a kernels stream DenseGeneral kernels as bf16 (activations, norms, softmax and caches stay fp32).
    A float16 kernel is widened to float32 first (exact); a kernel that is not bf16-representable - a real float32
    or float16 checkpoint - is rounded to bf16 ONCE, in place, with a warning, so that the encoder, the prompt
    prefill and the decode engine all see the very same numbers.  Returns True if values changed."""
    global _ROUNDING_WARNED
    w = m.weight
    if w.dtype == torch.bfloat16:
        return False
    if w.dtype == torch.float16:
        w.data = w.data.to(torch.float32)
    r = w.data.to(torch.bfloat16).to(torch.float32)
    if torch.equal(r, w.data):
        return False
    if not _ROUNDING_WARNED:
        import warnings
        warnings.warn("dia_tts_prune_b200: DenseGeneral kernels are not bf16-representable; rounding them to bf16 once "
                      "(the sm_100a path keeps weights in bf16 and everything else in fp32: logits differ from the "
                      "float32 reference by ~1e-2, as with the reference's own compute_dtype='bfloat16')", stacklevel=3)
        _ROUNDING_WARNED = True
    w.data.copy_(r)
    return True


def canonicalize_dense_kernels_(model: nn.Module) -> int:
    """:func:`canonicalize_dense_kernel_` over every DenseGeneral of ``model``; returns how many were rounded."""
    return sum(1 for m in model.modules() if isinstance(m, DenseGeneral) and canonicalize_dense_kernel_(m))


class MlpBlock(nn.Module):
    def __init__(self, embed_dim: int, intermediate_dim: int, compute_dtype: torch.dtype):
        super().__init__()
        self.dtype = compute_dtype
        self.wi_fused = DenseGeneral((embed_dim,), (2, intermediate_dim), axis=(-1,), weight_dtype=compute_dtype)
        self.wo = DenseGeneral((intermediate_dim,), (embed_dim,), axis=(-1,), weight_dtype=compute_dtype)

    def forward(self, x: Tensor) -> Tensor:
        if x.is_cuda and not torch.is_grad_enabled():
            return self._forward_b200(x, None, 0.0, None)
        gu = self.wi_fused(x)                               # [..., 2, F]: gate columns first, then up
        return self.wo(F.silu(gu[..., 0, :].float()).to(x.dtype) * gu[..., 1, :])

    def _forward_b200(self, x: Tensor, norm: RMSNorm | None, eps: float, residual: Tensor | None) -> Tensor:
        """``residual + wo(silu(gate) * up)`` with ``[gate | up] = wi_fused(norm(x))`` on our kernels: RMSNorm in the
        split pass of the first GEMM, the gate as one elementwise kernel, the residual add in the second epilogue."""
        from . import engine as _k
        lead = x.shape[:-1]
        gu = self.wi_fused._forward_tcgen05(x, (x.ndim - 1,), norm_weight=None if norm is None else norm.weight, eps=eps)
        h = _k.silu_mul(gu.reshape(-1, 2, gu.shape[-1]))
        return self.wo._forward_tcgen05(h.reshape(*lead, -1), (x.ndim - 1,), residual=residual)


class RotaryEmbedding(nn.Module):
    """Half-split RoPE, theta = position * max_timescale^(-2i/d) (intended semantics of
    dia/layers.py:126-173; the shipped forward has a shape bug)."""

    def __init__(self, embedding_dims: int, min_timescale: int = 1, max_timescale: int = 10000,
                 dtype: torch.dtype = torch.float32):
        super().__init__()
        if embedding_dims % 2:
            raise ValueError("Embedding dim must be even for RoPE.")
        self.embedding_dims, self.min_timescale, self.max_timescale = embedding_dims, min_timescale, max_timescale
        self.compute_dtype = dtype
        fraction = (2.0 * torch.arange(0, embedding_dims // 2)) / embedding_dims
        inv = 1.0 / (min_timescale * (max_timescale / min_timescale) ** fraction)
        self.register_buffer("inv_freq", inv.to(torch.float32), persistent=False)

    def forward(self, inputs: Tensor, position: Tensor) -> Tensor:
        theta = (position[..., None, None] * self.inv_freq.to(position.device)).to(torch.float32)   # [B,T,1,d/2]
        sin, cos = torch.sin(theta), torch.cos(theta)
        a, b = inputs.to(torch.float32).chunk(2, dim=-1)
        return torch.cat((a * cos - b * sin, a * sin + b * cos), dim=-1)


class Attention(nn.Module):
    """Library (fp32) attention used off the hot path: encoder, cross-KV precompute, prompt prefill."""

    def __init__(self, config: DiaConfig, q_embed_dim: int, kv_embed_dim: int, num_query_heads: int,
                 num_kv_heads: int, head_dim: int, compute_dtype: torch.dtype, is_cross_attn: bool = False,
                 out_embed_dim: int | None = None):
        super().__init__()
        if num_query_heads % num_kv_heads:
            raise ValueError(f"num_query_heads ({num_query_heads}) must be divisible by num_kv_heads ({num_kv_heads})")
        self.num_query_heads, self.num_kv_heads, self.head_dim = num_query_heads, num_kv_heads, head_dim
        self.is_cross_attn = is_cross_attn
        self.output_dim = out_embed_dim if out_embed_dim is not None else q_embed_dim
        self.projected_query_dim = num_query_heads * head_dim
        self.num_gqa_groups = num_query_heads // num_kv_heads
        self.compute_dtype = compute_dtype
        self.q_proj = DenseGeneral((q_embed_dim,), (num_query_heads, head_dim), axis=(-1,), weight_dtype=compute_dtype)
        self.k_proj = DenseGeneral((kv_embed_dim,), (num_kv_heads, head_dim), axis=(-1,), weight_dtype=compute_dtype)
        self.v_proj = DenseGeneral((kv_embed_dim,), (num_kv_heads, head_dim), axis=(-1,), weight_dtype=compute_dtype)
        self.o_proj = DenseGeneral((num_query_heads, head_dim), (self.output_dim,), axis=(-2, -1),
                                   weight_dtype=compute_dtype)
        self.rotary_emb = RotaryEmbedding(head_dim, config.model.rope_min_timescale, config.model.rope_max_timescale,
                                          dtype=compute_dtype)

    def forward(self, Xq: Tensor, Xkv: Tensor, q_positions: Tensor, kv_positions: Tensor | None = None,
                attn_mask: Tensor | None = None, cache: KVCache | None = None, prefill: bool = False,
                is_causal: bool = False) -> Tensor:
        if kv_positions is None:
            kv_positions = q_positions
        q = self.rotary_emb(self.q_proj(Xq.float()), q_positions).transpose(1, 2)          # [B, Hq, Tq, d]
        if self.is_cross_attn and cache is not None:
            # K/V were precomputed from the encoder output; the reference's per-step re-projection of
            # Xkv here is dead work (its result is discarded, dia/layers.py:274-275,284-287) and is omitted
            k, v = cache.k, cache.v
        else:
            k = self.rotary_emb(self.k_proj(Xkv.float()), kv_positions).transpose(1, 2)
            v = self.v_proj(Xkv.float()).transpose(1, 2)
            if cache is not None:
                k, v = cache.prefill(k, v) if prefill else cache.update(k, v)
        if self.num_gqa_groups > 1:
            k = k.repeat_interleave(self.num_gqa_groups, dim=1)
            v = v.repeat_interleave(self.num_gqa_groups, dim=1)
        out = F.scaled_dot_product_attention(q, k.float(), v.float(), attn_mask=attn_mask,
                                             is_causal=is_causal and not self.is_cross_attn, dropout_p=0.0)
        if attn_mask is not None and attn_mask.dtype == torch.bool:
            # a query with no allowed key contributes exact zeros (the reference's CPU SDPA behaviour,
            # SURVEY.md Appendix C Q7) - made explicit because GPU SDPA backends may return NaN there
            out = torch.where(attn_mask.any(dim=-1, keepdim=True), out, torch.zeros((), dtype=out.dtype, device=out.device))
        return self.o_proj(out.transpose(1, 2).contiguous()).to(Xq.dtype)

    def _forward_b200(self, x: Tensor, norm: RMSNorm, eps: float, positions_i32: Tensor, tables, mode: int,
                      n_valid: list[int] | None, cache: KVCache | None = None, prefill: bool = False) -> Tensor:
        """``x + o_proj(attention(...))`` for T > 1 rows on our kernels (no library call): RMSNorm fused into the split
        pass of the q / k / v GEMMs, RoPE fused with the scatter into the cache layout, fp32 flash-style attention,
        residual add in the o_proj epilogue.  ``mode``: 0 causal self, 1 pad-partitioned self, 2 cross over the valid
        prefix of ``cache`` (precomputed K/V).  Self-attention with ``cache`` + ``prefill`` writes slots [0, T)."""
        from . import engine as _k
        B, T, _ = x.shape
        H, Hkv = self.num_query_heads, self.num_kv_heads
        ax = (2,)
        q = self.q_proj._forward_tcgen05(x, ax, norm_weight=norm.weight, eps=eps)            # [B, T, H, 128]
        _k.rope_rows(q.reshape(B * T, H * 128), positions_i32, B, T, H, tables)
        if self.is_cross_attn:
            if cache is None:
                raise NotImplementedError("cross-attention needs the precomputed K/V cache")
            kc, vc, Tk = cache.k, cache.v, cache.k.shape[2]
        else:
            k = self.k_proj._forward_tcgen05(x, ax, norm_weight=norm.weight, eps=eps)        # [B, T, Hkv, 128]
            v = self.v_proj._forward_tcgen05(x, ax, norm_weight=norm.weight, eps=eps)
            if cache is not None:
                if not prefill:
                    raise NotImplementedError("T > 1 rows append through prefill only")
                kc, vc = cache.k, cache.v
                cache.current_idx = T - 1                                                  # KVCache.prefill (Appendix C Q2)
            else:
                kc = torch.empty((B, Hkv, T, 128), dtype=torch.float32, device=x.device)
                vc = torch.empty_like(kc)
            _k.rope_rows(k.reshape(B * T, Hkv * 128), positions_i32, B, T, Hkv, tables, cache=kc)
            _k.rope_rows(v.reshape(B * T, Hkv * 128), None, B, T, Hkv, tables, rotate=False, cache=vc)
            Tk = T
        o = _k.attention_rows(q, kc, vc, Tk, mode, n_valid)
        return self.o_proj._forward_tcgen05(o, (2, 3), residual=x)


class EncoderLayer(nn.Module):
    def __init__(self, config: DiaConfig, compute_dtype: torch.dtype):
        super().__init__()
        self.config, self.compute_dtype = config, compute_dtype
        e, eps = config.model.encoder, config.model.normalization_layer_epsilon
        self.pre_sa_norm = RMSNorm(e.n_embd, eps=eps)
        self.self_attention = Attention(config, e.n_embd, e.n_embd, e.n_head, e.n_head, e.head_dim, compute_dtype,
                                        is_cross_attn=False, out_embed_dim=e.n_embd)
        self.post_sa_norm = RMSNorm(e.n_embd, eps=eps)
        self.mlp = MlpBlock(e.n_embd, e.n_hidden, compute_dtype)

    def forward(self, x: Tensor, state: EncoderInferenceState) -> Tensor:
        if x.is_cuda and not torch.is_grad_enabled():
            from . import engine as _k
            eps = self.config.model.normalization_layer_epsilon
            B, T, _ = x.shape
            x = x.to(torch.float32).contiguous().clone()                                  # updated in place below
            pos = state.extras.get("pos_i32")
            if pos is None or pos.numel() != B * T:
                pos = state.extras["pos_i32"] = _k.positions_i32(T, B, x.device)
            x = self.self_attention._forward_b200(x, self.pre_sa_norm, eps, pos, _k.rope_tables_device(self.config, x.device),
                                                  1, _encoder_valid_lens(state))
            return self.mlp._forward_b200(x, self.post_sa_norm, eps, x)
        h = self.pre_sa_norm(x.float())
        x = x + self.self_attention(h, h, state.positions, state.positions, attn_mask=state.attn_mask)
        return x + self.mlp(self.post_sa_norm(x.float()))


class Encoder(nn.Module):
    def __init__(self, config: DiaConfig, compute_dtype: torch.dtype):
        super().__init__()
        self.config, self.compute_dtype = config, compute_dtype
        e = config.model.encoder
        self.embedding = nn.Embedding(config.model.src_vocab_size, e.n_embd)
        self.layers = nn.ModuleList([EncoderLayer(config, compute_dtype) for _ in range(e.n_layer)])
        self.norm = RMSNorm(e.n_embd, eps=config.model.normalization_layer_epsilon)

    def forward(self, x_ids: Tensor, state: EncoderInferenceState) -> Tensor:
        if x_ids.is_cuda and not torch.is_grad_enabled():
            from . import engine as _k
            B, T = x_ids.shape
            x = _k.embed_rows(self.embedding.weight, x_ids.reshape(-1).to(torch.int32).contiguous()).reshape(B, T, -1)
            for layer in self.layers:
                x = layer(x, state)
            return _k.rmsnorm_rows(x.reshape(B * T, -1), self.norm.weight,
                                   self.config.model.normalization_layer_epsilon).reshape(B, T, -1)
        with _exact_fp32():
            x = self.embedding(x_ids).float()
            for layer in self.layers:
                x = layer(x, state)
            return self.norm(x)


def _encoder_valid_lens(state: EncoderInferenceState) -> list[int]:
    """Valid text bytes per batch row (a prefix: pad = 0 never occurs inside utf-8 text); read back once per state."""
    lens = getattr(state, "valid_lens", None)
    if lens is None:
        pm = state.padding_mask
        lens = [int(v) for v in pm.sum(dim=-1).tolist()]
        for b, n in enumerate(lens):
            if not bool(pm[b, :n].all().item()):
                raise NotImplementedError("text with embedded pad bytes is not supported by the attention kernels")
        state.valid_lens = lens
    return lens


def _contract_dims(name: str) -> int:
    """Leading (input) axes of a decoder DenseGeneral kernel: 2 for the attention output projections
    ``[heads, head_dim, D]``, 1 for everything else."""
    return 2 if name.endswith("o_proj.weight") else 1


class DecoderLayer(nn.Module):
    def __init__(self, config: DiaConfig, compute_dtype: torch.dtype):
        super().__init__()
        self.config, self.compute_dtype = config, compute_dtype
        d, e, eps = config.model.decoder, config.model.encoder, config.model.normalization_layer_epsilon
        self.pre_sa_norm = RMSNorm(d.n_embd, eps=eps)
        self.pre_ca_norm = RMSNorm(d.n_embd, eps=eps)
        self.pre_mlp_norm = RMSNorm(d.n_embd, eps=eps)
        self.self_attention = Attention(config, d.n_embd, d.n_embd, d.gqa_query_heads, d.kv_heads, d.gqa_head_dim,
                                        compute_dtype, is_cross_attn=False, out_embed_dim=d.n_embd)
        self.cross_attention = Attention(config, d.n_embd, e.n_embd, d.cross_query_heads, d.cross_query_heads,
                                         d.cross_head_dim, compute_dtype, is_cross_attn=True, out_embed_dim=d.n_embd)
        self.mlp = MlpBlock(d.n_embd, d.n_hidden, compute_dtype)
        self._owner = None       # weakref to the Decoder, set by it
        self._index = -1

    def forward(self, x: Tensor, state: DecoderInferenceState, self_attn_cache: KVCache | None = None,
                cross_attn_cache: KVCache | None = None, prefill: bool = False) -> Tensor:
        if not prefill and x.shape[1] == 1:
            # hot path: the 8 fused stages of this layer in the persistent step kernel
            dec = self._owner() if self._owner is not None else None
            if dec is None:
                raise RuntimeError("DecoderLayer used for decoding outside a Decoder: no engine to run on")
            eng = dec._engine_for(state)
            if self_attn_cache is not state.self_attn_cache[self._index]:
                raise NotImplementedError("decode expects the layer's own cache from `state`")
            slot = self_attn_cache.current_idx
            y = eng.layer_step(self._index, x[:, 0, :], state.step_from, slot)
            self_attn_cache.current_idx = slot + 1
            return y[:, None, :].to(x.dtype)
        if x.is_cuda and not torch.is_grad_enabled():
            # prompt prefill (T > 1) on our kernels
            from . import engine as _k
            eps = self.config.model.normalization_layer_epsilon
            B, T, _ = x.shape
            tables = _k.rope_tables_device(self.config, x.device)
            pos = state.extras.get("prefill_pos")
            if pos is None or pos.numel() != B * T:
                pos = _k.positions_i32(state.step_to - state.step_from, B, x.device, start=state.step_from)
                state.extras["prefill_pos"] = pos
            if state.text_len is None:
                raise NotImplementedError("prefill needs DecoderInferenceState.text_len (valid text bytes)")
            x = x.to(torch.float32).contiguous()
            x = self.self_attention._forward_b200(x, self.pre_sa_norm, eps, pos, tables, 0, None,
                                                  cache=self_attn_cache, prefill=True)
            x = self.cross_attention._forward_b200(x, self.pre_ca_norm, eps, pos, tables, 2,
                                                   [0] * (B - 1) + [state.text_len], cache=cross_attn_cache)
            return self.mlp._forward_b200(x, self.pre_mlp_norm, eps, x)
        # CPU (state-dict tooling and the CPU test suite): the reference's own formulation
        h = self.pre_sa_norm(x.float())
        x = x + self.self_attention(h, h, state.dec_positions, state.dec_positions, attn_mask=None,
                                    cache=self_attn_cache, prefill=prefill, is_causal=prefill)
        h = self.pre_ca_norm(x.float())
        x = x + self.cross_attention(h, state.enc_out, state.dec_positions, state.enc_positions,
                                     attn_mask=state.dec_cross_attn_mask, cache=cross_attn_cache)
        return x + self.mlp(self.pre_mlp_norm(x.float()))


class Decoder(nn.Module):
    def __init__(self, config: DiaConfig, compute_dtype: torch.dtype):
        super().__init__()
        self.config, self.compute_dtype = config, compute_dtype
        d = config.model.decoder
        self.num_channels, self.num_layers = config.data.channels, d.n_layer
        self.embeddings = nn.ModuleList([nn.Embedding(config.model.tgt_vocab_size, d.n_embd)
                                         for _ in range(self.num_channels)])
        self.layers = nn.ModuleList([DecoderLayer(config, compute_dtype) for _ in range(self.num_layers)])
        self.norm = RMSNorm(d.n_embd, eps=config.model.normalization_layer_epsilon)
        self.logits_dense = DenseGeneral((d.n_embd,), (self.num_channels, config.model.tgt_vocab_size), axis=(-1,),
                                         weight_dtype=compute_dtype)
        for i, layer in enumerate(self.layers):
            layer._owner, layer._index = weakref.ref(self), i
        self._engine = None
        self._engine_sig = None
        self.compact_pruned_mlp = True      # drop exactly-dead MLP neurons from the engine's weight stream
        # drop all-zero input rows of the other kernels (K-row compaction); DIA_NO_ROW_COMPACTION=1 is the A/B switch of bench.py
        self.compact_pruned_rows = os.environ.get("DIA_NO_ROW_COMPACTION", "0") != "1"
        self.use_sparse24 = True            # stream 2:4 checkpoints compressed (mma.sp) when every kernel qualifies
        self.register_load_state_dict_post_hook(lambda m, k: m.invalidate_engine())

    # ---- engine management -------------------------------------------------------------------------
    def invalidate_engine(self) -> None:
        """Drop the repacked device copy of the weights (call after modifying parameters in place)."""
        self._engine_sig = None

    def _apply(self, fn, *a, **kw):                      # .to() / .cuda() / .float() move or retype parameters
        self._engine_sig = None
        return super()._apply(fn, *a, **kw)

    def _weights_signature(self):
        w = self.logits_dense.weight
        return (w.device, w.dtype, w.data_ptr(), sum(p._version for p in self.parameters()))

    def engine(self):
        """The bound ``DecodeEngine`` (weights repacked on first use and after any change)."""
        from .engine import DecodeEngine, decoder_tensor_names
        dev = self.logits_dense.weight.device
        if dev.type != "cuda":
            raise RuntimeError("the Dia decode path needs the model on a CUDA device (sm_100a); no CPU fallback")
        sig = self._weights_signature()
        if self._engine is not None and self._engine_sig == sig and self._engine.device == dev:
            return self._engine
        # (re)pack: a checkpoint whose MLP was structurally pruned (zero hidden neurons) gets a narrower engine
        from .pruning_utils import compact_mlp, is_2to4, plan_mlp_compaction
        from .synthetic import is_dense_kernel
        import torch.nn.utils.prune as _prune
        if any(_prune.is_pruned(m) for m in self.modules() if isinstance(m, DenseGeneral)):
            raise RuntimeError("pruning masks are still attached (parameters are named weight_orig): call "
                               "pruning_utils.make_pruning_permanent(model) before decoding")
        canonicalize_dense_kernels_(self)
        sig = self._weights_signature()
        sd = dict(self.named_parameters())
        tensors = {n: sd[n].detach() for n in decoder_tensor_names(self.config)}
        d = self.config.model.decoder
        plan = plan_mlp_compaction(tensors, d.n_layer, d.n_hidden) if self.compact_pruned_mlp else None
        width = plan[0] if plan is not None else d.n_hidden
        if plan is not None:
            tensors = compact_mlp(tensors, plan)
        # a 2:4 checkpoint (every dense kernel) streams compressed and runs on mma.sp
        sparse = bool(self.use_sparse24) and all(is_2to4(t, _contract_dims(n)) for n, t in tensors.items()
                                                 if is_dense_kernel(n))
        # `--prune-dim 0` checkpoints: all-zero INPUT rows of the other kernels are dropped as well (K-row compaction)
        from .pruning_utils import compact_rows, plan_row_compaction
        rplan = plan_row_compaction(tensors, d.n_layer) if (self.compact_pruned_rows and not sparse) else {}
        maps = {}
        if rplan:
            tensors, maps = compact_rows(tensors, rplan)
        k_rows = {fam: w for fam, (w, _) in rplan.items()}
        if self._engine is None or self._engine.device != dev or self._engine.n_hidden != width or \
                self._engine.sparse24 != sparse or self._engine.k_rows != k_rows:
            if self._engine is not None:
                self._engine.close()
            self._engine = DecodeEngine(self.config, dev, n_hidden=width, sparse24=sparse, k_rows=k_rows)
        self._engine.load_weights(tensors)
        if maps:
            self._engine.set_row_maps(maps)
        self._engine_sig = sig
        return self._engine

    def batch_engine(self, max_utterances: int = 8):
        """The ``BatchDecodeEngine`` (N utterances per launch on the tcgen05 step kernel), repacked like :meth:`engine`."""
        from .engine import BatchDecodeEngine, decoder_tensor_names
        dev = self.logits_dense.weight.device
        if dev.type != "cuda":
            raise RuntimeError("the Dia decode path needs the model on a CUDA device (sm_100a); no CPU fallback")
        sig = (self._weights_signature(), int(max_utterances))
        if getattr(self, "_bengine", None) is not None and self._bengine_sig == sig and self._bengine.device == dev:
            return self._bengine
        import torch.nn.utils.prune as _prune
        if any(_prune.is_pruned(m) for m in self.modules() if isinstance(m, DenseGeneral)):
            raise RuntimeError("pruning masks are still attached: call pruning_utils.make_pruning_permanent(model) first")
        canonicalize_dense_kernels_(self)
        sig = (self._weights_signature(), int(max_utterances))
        from .pruning_utils import compact_mlp, plan_mlp_compaction
        sd = dict(self.named_parameters())
        tensors = {n: sd[n].detach() for n in decoder_tensor_names(self.config)}
        d = self.config.model.decoder
        plan = plan_mlp_compaction(tensors, d.n_layer, d.n_hidden) if self.compact_pruned_mlp else None
        width = plan[0] if plan is not None else d.n_hidden
        if plan is not None:
            tensors = compact_mlp(tensors, plan)
        old = getattr(self, "_bengine", None)
        if old is None or old.device != dev or old.n_hidden != width or old.max_utterances != int(max_utterances):
            if old is not None:
                old.close()
            self._bengine = BatchDecodeEngine(self.config, dev, max_utterances=int(max_utterances), n_hidden=width)
        self._bengine.load_weights(tensors)
        self._bengine_sig = sig
        return self._bengine

    def _engine_for(self, state: DecoderInferenceState):
        eng = self.engine()
        if state.text_len is None:
            row = state.dec_cross_attn_mask[1].reshape(-1)
            n = int(row.sum().item())
            if bool(state.dec_cross_attn_mask[0].any().item()) or not bool(row[:n].all().item()):
                raise NotImplementedError("cross-attention mask must be (all-pad row 0, valid-prefix row 1)")
            state.text_len = n
        key = tuple(c.k.data_ptr() for c in state.self_attn_cache) + \
            tuple(c.k.data_ptr() for c in state.cross_attn_cache) + (state.text_len,)
        if eng.bound_key() != key:
            for c in state.cross_attn_cache:            # the kernels need contiguous fp32 [2,H,S,128]
                if c.k.dtype != torch.float32 or not c.k.is_contiguous():
                    c.k = c.k.to(torch.float32).contiguous()
                if c.v.dtype != torch.float32 or not c.v.is_contiguous():
                    c.v = c.v.to(torch.float32).contiguous()
            eng.bind(state.self_attn_cache, state.cross_attn_cache, state.text_len)
        return eng

    # ---- once per utterance ---------------------------------------------------------------------------
    def precompute_cross_attn_cache(self, enc_out: Tensor, enc_positions: Tensor) -> list[KVCache]:
        out: list[KVCache] = []
        if enc_out.is_cuda and not torch.is_grad_enabled():
            from . import engine as _k
            B, S, _ = enc_out.shape
            x = enc_out.to(torch.float32).contiguous()
            tables, pos = _k.rope_tables_device(self.config, x.device), _k.positions_i32(S, B, x.device)
            for layer in self.layers:
                ca = layer.cross_attention
                H = ca.num_kv_heads
                k = torch.empty((B, H, S, 128), dtype=torch.float32, device=x.device)
                v = torch.empty_like(k)
                _k.rope_rows(ca.k_proj._forward_tcgen05(x, (2,)).reshape(B * S, H * 128), pos, B, S, H, tables, cache=k)
                _k.rope_rows(ca.v_proj._forward_tcgen05(x, (2,)).reshape(B * S, H * 128), None, B, S, H, tables,
                             rotate=False, cache=v)
                out.append(KVCache.from_kv(k, v))
            return out
        with torch.no_grad(), _exact_fp32():
            x = enc_out.float()
            for layer in self.layers:
                ca = layer.cross_attention
                k = ca.rotary_emb(ca.k_proj(x), enc_positions).transpose(1, 2).contiguous()
                v = ca.v_proj(x).transpose(1, 2).contiguous()
                out.append(KVCache.from_kv(k, v))
        return out

    def precompute_cross_attn_cache_live(self, enc_live: Tensor, text_length: int) -> list[KVCache]:
        """Same caches as ``precompute_cross_attn_cache`` for the entries that are ever observed.  ``enc_live`` is
        the encoder output of the valid text bytes of the conditional row, ``[1, n, E]``.  The reference projects
        all 2 x text_length encoder positions in every layer, but the unconditional row is fully masked and so are
        the pad positions of the conditional row (overwriting them with noise leaves the
        logits bit-identical); here those entries are zeros and only ``n`` rows are projected."""
        out: list[KVCache] = []
        n = enc_live.shape[1]
        if enc_live.is_cuda and not torch.is_grad_enabled():
            from . import engine as _k
            x = enc_live.to(torch.float32).contiguous()
            tables, pos_i = _k.rope_tables_device(self.config, x.device), _k.positions_i32(n, 1, x.device)
            for layer in self.layers:
                ca = layer.cross_attention
                H = ca.num_kv_heads
                k = torch.zeros((2, H, text_length, 128), dtype=torch.float32, device=x.device)
                v = torch.zeros_like(k)
                _k.rope_rows(ca.k_proj._forward_tcgen05(x, (2,)).reshape(n, H * 128), pos_i, 1, n, H, tables, cache=k[1:])
                _k.rope_rows(ca.v_proj._forward_tcgen05(x, (2,)).reshape(n, H * 128), None, 1, n, H, tables,
                             rotate=False, cache=v[1:])
                out.append(KVCache.from_kv(k, v))
            return out
        pos = torch.arange(n, dtype=torch.float32, device=enc_live.device)[None, :]
        with torch.no_grad(), _exact_fp32():
            x = enc_live.float()
            for layer in self.layers:
                ca = layer.cross_attention
                k_live = ca.rotary_emb(ca.k_proj(x), pos).transpose(1, 2)           # [1, H, n, d]
                v_live = ca.v_proj(x).transpose(1, 2)
                k = torch.zeros((2, k_live.shape[1], text_length, k_live.shape[3]), dtype=k_live.dtype, device=x.device)
                v = torch.zeros_like(k)
                k[1, :, :n] = k_live[0]
                v[1, :, :n] = v_live[0]
                out.append(KVCache.from_kv(k, v))
        return out

    # ---- the hot path -------------------------------------------------------------------------------------
    def decode_step(self, tgt_ids_Bx1xC: Tensor, state: DecoderInferenceState) -> Tensor:
        B, T, C = tgt_ids_Bx1xC.shape
        assert T == 1, "decode_step expects T=1"
        assert C == self.num_channels, "Input channels mismatch"
        assert B == 2, "decode_step expects the CFG batch of 2"
        eng = self._engine_for(state)
        slot = state.self_attn_cache[0].current_idx
        logits = eng.decode_step(tgt_ids_Bx1xC.reshape(2, C), state.step_from, slot)
        for c in state.self_attn_cache:
            c.current_idx = slot + 1
        return logits[:, None, :, :]

    def forward(self, tgt_ids_BxTxC: Tensor, state: DecoderInferenceState, want_logits: bool = True) -> Tensor:
        """Prompt prefill / teacher-forced pass over T positions (dia/layers.py:722-766).  ``want_logits=False`` skips
        the logits head, whose output ``Dia._prepare_generation`` discards (dia/model.py:420-421)."""
        B, T, C = tgt_ids_BxTxC.shape
        assert C == self.num_channels, "Input channels mismatch"
        if tgt_ids_BxTxC.is_cuda and not torch.is_grad_enabled():
            # own kernels end to end: embedding gather-sum, 18 x (fused-norm GEMMs, RoPE + cache scatter, fp32 attention,
            # gate, residual epilogues), final norm fused into the logits GEMM
            x = self.engine().embed_sum(tgt_ids_BxTxC.reshape(B * T, C)).reshape(B, T, -1)
            state.extras.pop("prefill_pos", None)
            for i, layer in enumerate(self.layers):
                x = layer(x, state, self_attn_cache=state.self_attn_cache[i],
                          cross_attn_cache=state.cross_attn_cache[i], prefill=True)
            if not want_logits:
                return None
            return self.logits_dense._forward_tcgen05(x, (2,), norm_weight=self.norm.weight,
                                                      eps=self.config.model.normalization_layer_epsilon)
        with _exact_fp32():
            x = None
            for i, emb in enumerate(self.embeddings):
                e = emb(tgt_ids_BxTxC[..., i]).float()
                x = e if x is None else x + e
            for i, layer in enumerate(self.layers):
                x = layer(x, state, self_attn_cache=state.self_attn_cache[i],
                          cross_attn_cache=state.cross_attn_cache[i], prefill=True)
            return self.logits_dense(self.norm(x.float())).to(torch.float32)


class DiaModel(nn.Module, PyTorchModelHubMixin, repo_url="https://github.com/nari-labs/dia",
               pipeline_tag="text-to-speech", license="apache-2.0",
               coders={DiaConfig: (lambda c: c.model_dump(), lambda d: DiaConfig.model_validate(d))}):
    def __init__(self, config: DiaConfig, compute_dtype: Optional[torch.dtype] = None):
        super().__init__()
        self.config = config
        if compute_dtype is None:
            compute_dtype = {"bfloat16": torch.bfloat16, "float16": torch.float16}.get(
                getattr(config.model, "weight_dtype", "float32"), torch.float32)
            print(f"Inferred compute_dtype: {compute_dtype} from config")
        self.encoder = Encoder(config, compute_dtype)
        self.decoder = Decoder(config, compute_dtype)
