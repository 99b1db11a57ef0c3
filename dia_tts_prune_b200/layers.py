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

    def forward(self, inputs: Tensor) -> Tensor:
        n = len(self.axis)
        x_axes = tuple(a if a >= 0 else inputs.ndim + a for a in self.axis)
        y = self._forward_tcgen05(inputs, x_axes)
        if y is not None:
            return y
        out = torch.tensordot(inputs.to(torch.float32), self.weight.to(torch.float32), dims=(x_axes, tuple(range(n))))
        return out.to(inputs.dtype)

    # ---- sm_100a path for more than a few rows (encoder, cross-attention K/V precompute, prompt prefill) ----------
    def _forward_tcgen05(self, inputs: Tensor, x_axes: tuple[int, ...]) -> Tensor | None:
        """``x @ W`` on the tcgen05 tensor cores with the activations split into three bf16 terms (float32-operand
        accuracy).  Used when the contracted axes are the trailing ones, the shape fits the 128 x 128 x 64 tiling and
        the kernel is bfloat16 (or float32 holding bfloat16-representable values, as every kernel that went through
        the bf16 checkpoint path is); otherwise ``None`` and the caller runs the reference's float32 tensordot."""
        w = self.weight
        if not (inputs.is_cuda and w.is_cuda) or torch.is_grad_enabled() and (w.requires_grad or inputs.requires_grad):
            return None
        n = len(x_axes)
        if x_axes != tuple(range(inputs.ndim - n, inputs.ndim)):
            return None
        K = math.prod(self.in_shapes)
        N = math.prod(self.out_features)
        M = inputs.numel() // K if K else 0
        from . import engine as _engine
        if M < 8 or not _engine.dense_supported(M, N, K):
            return None
        key = (w.data_ptr(), w._version, w.dtype, w.device)
        cached = getattr(self, "_b200_wt", None)
        if cached is None or cached[0] != key:
            w2 = w.detach().reshape(K, N)
            ok = w.dtype == torch.bfloat16 or (w.dtype == torch.float32 and
                                               bool(torch.equal(w2, w2.to(torch.bfloat16).to(torch.float32))))
            cached = (key, _engine.dense_prepare_weight(w2) if ok else None)
            self._b200_wt = cached
        if cached[1] is None:
            return None
        y = _engine.dense_forward(inputs.reshape(M, K), cached[1])
        return y.reshape(*inputs.shape[: inputs.ndim - n], *self.out_features).to(inputs.dtype)


class MlpBlock(nn.Module):
    def __init__(self, embed_dim: int, intermediate_dim: int, compute_dtype: torch.dtype):
        super().__init__()
        self.dtype = compute_dtype
        self.wi_fused = DenseGeneral((embed_dim,), (2, intermediate_dim), axis=(-1,), weight_dtype=compute_dtype)
        self.wo = DenseGeneral((intermediate_dim,), (embed_dim,), axis=(-1,), weight_dtype=compute_dtype)

    def forward(self, x: Tensor) -> Tensor:
        gu = self.wi_fused(x)                               # [..., 2, F]: gate columns first, then up
        return self.wo(F.silu(gu[..., 0, :].float()).to(x.dtype) * gu[..., 1, :])


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
        with _exact_fp32():
            x = self.embedding(x_ids).float()
            for layer in self.layers:
                x = layer(x, state)
            return self.norm(x)


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
        # prompt prefill (T > 1): library path, fp32
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
        if self._engine is None or self._engine.device != dev or self._engine.n_hidden != width or \
                self._engine.sparse24 != sparse:
            if self._engine is not None:
                self._engine.close()
            self._engine = DecodeEngine(self.config, dev, n_hidden=width, sparse24=sparse)
        self._engine.load_weights(tensors)
        self._engine_sig = sig
        return self._engine

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

    def forward(self, tgt_ids_BxTxC: Tensor, state: DecoderInferenceState) -> Tensor:
        """Prompt prefill / teacher-forced pass over T positions (library path, fp32)."""
        B, T, C = tgt_ids_BxTxC.shape
        assert C == self.num_channels, "Input channels mismatch"
        with _exact_fp32():
            if tgt_ids_BxTxC.is_cuda:
                x = self.engine().embed_sum(tgt_ids_BxTxC.reshape(B * T, C)).reshape(B, T, -1)
            else:
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
