"""Functional torch-CPU fp32 restatement of the reference's Dia inference path.

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.  Never imported by the
product package.

The reference (babybirdprd/dia-tts-prune) is pure PyTorch; every arithmetic op
on its decode path is an ATen library call.  This file
restates the *algorithm* - which ops, on which operands, in which order - as
plain functions over a flat ``params`` dict that uses the reference's
state_dict keys, calling the same ATen entry points
(``torch.tensordot``, ``F.rms_norm``, ``F.scaled_dot_product_attention``,
``F.embedding``, ``F.silu``) so that on the same CPU it reproduces the
reference bit for bit.

PINNED: ``oracle/validate_against_reference.py`` (run in the build container,
where /root/reference exists) imports the reference with the three hot-path
patches of SURVEY.md Appendix B and checks this file against it: encoder
output, cross-KV, per-step logits, KV-cache contents, greedy code stream, the
prompt/prefill path and the sampling filter.  The same script writes the
golden fixtures under tests/golden/ that travel to the GPU box.

Every function cites the reference lines it follows (paths relative to
/root/reference).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# parameter naming (dia/layers.py module registration order)
# ----------------------------------------------------------------------------

def param_names(cfg) -> list[str]:
    """state_dict keys in ``DiaModel.named_parameters()`` order.

    Order follows attribute assignment order in the reference constructors:
    Encoder dia/layers.py:429-441, EncoderLayer :360-383, Attention :204-227,
    MlpBlock :77-90, Decoder :601-628, DecoderLayer :479-528.
    """
    names = ["encoder.embedding.weight"]
    for i in range(cfg.model.encoder.n_layer):
        p = f"encoder.layers.{i}."
        names += [p + "pre_sa_norm.weight"]
        names += [p + f"self_attention.{w}_proj.weight" for w in "qkvo"]
        names += [p + "post_sa_norm.weight", p + "mlp.wi_fused.weight", p + "mlp.wo.weight"]
    names += ["encoder.norm.weight"]
    names += [f"decoder.embeddings.{c}.weight" for c in range(cfg.data.channels)]
    for i in range(cfg.model.decoder.n_layer):
        p = f"decoder.layers.{i}."
        names += [p + "pre_sa_norm.weight", p + "pre_ca_norm.weight", p + "pre_mlp_norm.weight"]
        names += [p + f"self_attention.{w}_proj.weight" for w in "qkvo"]
        names += [p + f"cross_attention.{w}_proj.weight" for w in "qkvo"]
        names += [p + "mlp.wi_fused.weight", p + "mlp.wo.weight"]
    names += ["decoder.norm.weight", "decoder.logits_dense.weight"]
    return names


def param_shapes(cfg) -> dict[str, tuple[int, ...]]:
    """Shapes of every parameter (DenseGeneral kernel = in_shapes + out_features,
    dia/layers.py:44-51; embeddings :429-434,601-610; RMSNorm :360-382,479-493)."""
    e, d, m, dt = cfg.model.encoder, cfg.model.decoder, cfg.model, cfg.data
    shp: dict[str, tuple[int, ...]] = {"encoder.embedding.weight": (m.src_vocab_size, e.n_embd)}
    for i in range(e.n_layer):
        p = f"encoder.layers.{i}."
        shp[p + "pre_sa_norm.weight"] = (e.n_embd,)
        for w in "qkv":
            shp[p + f"self_attention.{w}_proj.weight"] = (e.n_embd, e.n_head, e.head_dim)
        shp[p + "self_attention.o_proj.weight"] = (e.n_head, e.head_dim, e.n_embd)
        shp[p + "post_sa_norm.weight"] = (e.n_embd,)
        shp[p + "mlp.wi_fused.weight"] = (e.n_embd, 2, e.n_hidden)
        shp[p + "mlp.wo.weight"] = (e.n_hidden, e.n_embd)
    shp["encoder.norm.weight"] = (e.n_embd,)
    for c in range(dt.channels):
        shp[f"decoder.embeddings.{c}.weight"] = (m.tgt_vocab_size, d.n_embd)
    for i in range(d.n_layer):
        p = f"decoder.layers.{i}."
        for n in ("pre_sa_norm", "pre_ca_norm", "pre_mlp_norm"):
            shp[p + n + ".weight"] = (d.n_embd,)
        shp[p + "self_attention.q_proj.weight"] = (d.n_embd, d.gqa_query_heads, d.gqa_head_dim)
        shp[p + "self_attention.k_proj.weight"] = (d.n_embd, d.kv_heads, d.gqa_head_dim)
        shp[p + "self_attention.v_proj.weight"] = (d.n_embd, d.kv_heads, d.gqa_head_dim)
        shp[p + "self_attention.o_proj.weight"] = (d.gqa_query_heads, d.gqa_head_dim, d.n_embd)
        shp[p + "cross_attention.q_proj.weight"] = (d.n_embd, d.cross_query_heads, d.cross_head_dim)
        shp[p + "cross_attention.k_proj.weight"] = (e.n_embd, d.cross_query_heads, d.cross_head_dim)
        shp[p + "cross_attention.v_proj.weight"] = (e.n_embd, d.cross_query_heads, d.cross_head_dim)
        shp[p + "cross_attention.o_proj.weight"] = (d.cross_query_heads, d.cross_head_dim, d.n_embd)
        shp[p + "mlp.wi_fused.weight"] = (d.n_embd, 2, d.n_hidden)
        shp[p + "mlp.wo.weight"] = (d.n_hidden, d.n_embd)
    shp["decoder.norm.weight"] = (d.n_embd,)
    shp["decoder.logits_dense.weight"] = (d.n_embd, dt.channels, m.tgt_vocab_size)
    return shp


# ----------------------------------------------------------------------------
# elementary ops
# ----------------------------------------------------------------------------

def dense(x: torch.Tensor, w: torch.Tensor, n_contract: int = 1) -> torch.Tensor:
    """DenseGeneral.forward, dia/layers.py:55-66: contract the last
    ``n_contract`` axes of x with the first ``n_contract`` axes of the
    [in..., out...] kernel; result cast back to x's dtype."""
    xa = tuple(range(x.ndim - n_contract, x.ndim))
    wa = tuple(range(n_contract))
    return torch.tensordot(x.to(w.dtype), w, dims=(xa, wa)).to(x.dtype)


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """torch.nn.RMSNorm applied to ``x.to(float32)`` (dia/layers.py:541,560,579,714)."""
    return F.rms_norm(x.to(torch.float32), (x.shape[-1],), w, eps)


def rope_inv_freq(head_dim: int, min_ts: int, max_ts: int) -> torch.Tensor:
    """RotaryEmbedding.__init__ timescales, dia/layers.py:126-132 (same
    expression, so the same fp32 bits)."""
    half = head_dim // 2
    fraction = (2.0 * torch.arange(0, half)) / head_dim
    inv_freq = 1.0 / (min_ts * (max_ts / min_ts) ** fraction)
    return inv_freq.to(torch.float32)


def rope(x: torch.Tensor, position: torch.Tensor, inv_freq: torch.Tensor) -> torch.Tensor:
    """RotaryEmbedding.forward with patch B1 of SURVEY Appendix B: the rotation
    of dia/layers.py:165-173 with theta = position * inv_freq broadcast over
    the head axis (x is [B, T, N, H], position is [B, T])."""
    pos = position.unsqueeze(-1).unsqueeze(-1)
    freqs = pos * inv_freq
    sin = torch.sin(freqs.to(torch.float32))
    cos = torch.cos(freqs.to(torch.float32))
    x1, x2 = torch.chunk(x.to(torch.float32), 2, dim=-1)
    return torch.cat((x1 * cos - x2 * sin, x1 * sin + x2 * cos), dim=-1)


def rope_table(max_pos: int, head_dim: int, min_ts: int, max_ts: int) -> tuple[torch.Tensor, torch.Tensor]:
    """sin/cos of ``int32 position * inv_freq`` for positions 0..max_pos-1 -
    exactly the values ``rope`` uses for decoder positions (int32 tensor times
    fp32 tensor promotes to fp32, dia/state.py:167-169 + dia/layers.py:146)."""
    inv = rope_inv_freq(head_dim, min_ts, max_ts)
    pos = torch.arange(max_pos, dtype=torch.int32).unsqueeze(-1)
    freqs = (pos * inv).to(torch.float32)
    return torch.sin(freqs), torch.cos(freqs)


def mlp(x: torch.Tensor, wi: torch.Tensor, wo: torch.Tensor) -> torch.Tensor:
    """MlpBlock.forward, dia/layers.py:92-105 (gate = [...,0,:], up = [...,1,:])."""
    fused = dense(x, wi)
    gate, up = fused[..., 0, :], fused[..., 1, :]
    hidden = torch.mul(F.silu(gate.to(torch.float32)).to(x.dtype), up)
    return dense(hidden, wo)


def create_attn_mask(q_pad: torch.Tensor, k_pad: torch.Tensor, is_causal: bool = False) -> torch.Tensor:
    """dia/state.py:8-39: attend iff both non-pad or both pad (+ tril)."""
    pq, pk = q_pad.unsqueeze(2), k_pad.unsqueeze(1)
    mask = (pq & pk) | ((~pq) & (~pk))
    if is_causal:
        tq, tk = q_pad.shape[1], k_pad.shape[1]
        mask = mask & torch.tril(torch.ones((tq, tk), dtype=torch.bool))
    return mask.unsqueeze(1)


# ----------------------------------------------------------------------------
# state
# ----------------------------------------------------------------------------

class KV:
    """dia/state.py:72-109 with patch B3 (prefill returns k, v)."""

    def __init__(self, heads: int, max_len: int, head_dim: int, k=None, v=None):
        self.k = torch.zeros((2, heads, max_len, head_dim)) if k is None else k
        self.v = torch.zeros((2, heads, max_len, head_dim)) if v is None else v
        self.current_idx = 0

    def update(self, k, v):
        i = self.current_idx
        self.k[:, :, i:i + 1, :] = k
        self.v[:, :, i:i + 1, :] = v
        self.current_idx = i + 1
        return self.k[:, :, :i + 1, :], self.v[:, :, :i + 1, :]

    def prefill(self, k, v):
        n = k.shape[2]
        self.k[:, :, :n, :] = k
        self.v[:, :, :n, :] = v
        self.current_idx = n - 1
        return k, v


@dataclass
class DecState:
    """dia/state.py:112-169 (DecoderInferenceState)."""
    enc_out: torch.Tensor
    enc_positions: torch.Tensor
    dec_positions: torch.Tensor
    cross_mask: torch.Tensor
    self_cache: list
    cross_cache: list

    def prepare_step(self, step_from: int, step_to: int | None = None):
        if step_to is None:
            step_to = step_from + 1
        self.dec_positions = torch.arange(step_from, step_to, dtype=torch.int32).unsqueeze(0).expand(2, -1)


# ----------------------------------------------------------------------------
# attention / layers
# ----------------------------------------------------------------------------

def attention(P, pre: str, cfg, xq, xkv, q_pos, kv_pos, mask, cache, *, cross: bool, heads: int,
              kv_heads: int, prefill: bool = False, is_causal: bool = False, dead_cross_kv: bool = True):
    """Attention.forward, dia/layers.py:238-346.

    ``dead_cross_kv`` keeps the reference's per-step re-projection of the whole
    encoder output in cross-attention (:274-275,279), whose result is dropped
    at :284-287.  It has no effect on outputs; it is kept (default) so that a
    timed run of this port costs what the reference costs.
    """
    inv = rope_inv_freq(P[pre + "q_proj.weight"].shape[-1], cfg.model.rope_min_timescale, cfg.model.rope_max_timescale)
    q = rope(dense(xq, P[pre + "q_proj.weight"]), q_pos, inv)
    if cross:
        if dead_cross_kv:
            k_dead = dense(xkv, P[pre + "k_proj.weight"])
            _ = dense(xkv, P[pre + "v_proj.weight"])
            _ = rope(k_dead, kv_pos, inv)
        ak, av = cache.k, cache.v
    else:
        k = rope(dense(xkv, P[pre + "k_proj.weight"]), kv_pos, inv)
        v = dense(xkv, P[pre + "v_proj.weight"])
        kc, vc = k.transpose(1, 2), v.transpose(1, 2)
        if cache is None:
            ak, av = kc, vc
        elif prefill:
            ak, av = cache.prefill(kc, vc)
        else:
            ak, av = cache.update(kc, vc)
    aq = q.transpose(1, 2)
    groups = heads // kv_heads
    if groups > 1:
        ak = ak.repeat_interleave(groups, dim=1)
        av = av.repeat_interleave(groups, dim=1)
    out = F.scaled_dot_product_attention(aq, ak, av, attn_mask=mask, is_causal=is_causal and not cross, dropout_p=0.0)
    out = out.transpose(1, 2).contiguous()
    return dense(out, P[pre + "o_proj.weight"], n_contract=2)


def encoder_forward(P, cfg, ids: torch.Tensor, positions: torch.Tensor, attn_mask: torch.Tensor) -> torch.Tensor:
    """Encoder.forward + EncoderLayer.forward, dia/layers.py:385-416,445-462."""
    e, eps = cfg.model.encoder, cfg.model.normalization_layer_epsilon
    x = F.embedding(ids, P["encoder.embedding.weight"])
    for i in range(e.n_layer):
        p = f"encoder.layers.{i}."
        h = rms_norm(x, P[p + "pre_sa_norm.weight"], eps)
        x = x + attention(P, p + "self_attention.", cfg, h, h, positions, positions, attn_mask, None,
                          cross=False, heads=e.n_head, kv_heads=e.n_head)
        h = rms_norm(x, P[p + "post_sa_norm.weight"], eps)
        x = x + mlp(h, P[p + "mlp.wi_fused.weight"], P[p + "mlp.wo.weight"])
    return rms_norm(x, P["encoder.norm.weight"], eps)


def precompute_cross_kv(P, cfg, enc_out: torch.Tensor, enc_positions: torch.Tensor) -> list[KV]:
    """Decoder.precompute_cross_attn_cache, dia/layers.py:632-669."""
    d = cfg.model.decoder
    inv = rope_inv_freq(d.cross_head_dim, cfg.model.rope_min_timescale, cfg.model.rope_max_timescale)
    out = []
    for i in range(d.n_layer):
        p = f"decoder.layers.{i}.cross_attention."
        k = rope(dense(enc_out, P[p + "k_proj.weight"]), enc_positions, inv).transpose(1, 2)
        v = dense(enc_out, P[p + "v_proj.weight"]).transpose(1, 2)
        out.append(KV(k.shape[1], k.shape[2], k.shape[3], k=k, v=v))
    return out


def decoder_layer(P, cfg, i: int, x, st: DecState, prefill: bool, dead_cross_kv: bool = True):
    """DecoderLayer.forward, dia/layers.py:530-584."""
    d, eps = cfg.model.decoder, cfg.model.normalization_layer_epsilon
    p = f"decoder.layers.{i}."
    h = rms_norm(x, P[p + "pre_sa_norm.weight"], eps)
    x = x + attention(P, p + "self_attention.", cfg, h, h, st.dec_positions, st.dec_positions, None,
                      st.self_cache[i], cross=False, heads=d.gqa_query_heads, kv_heads=d.kv_heads,
                      prefill=prefill, is_causal=prefill)
    h = rms_norm(x, P[p + "pre_ca_norm.weight"], eps)
    x = x + attention(P, p + "cross_attention.", cfg, h, st.enc_out, st.dec_positions, st.enc_positions,
                      st.cross_mask, st.cross_cache[i], cross=True, heads=d.cross_query_heads,
                      kv_heads=d.cross_query_heads, dead_cross_kv=dead_cross_kv)
    h = rms_norm(x, P[p + "pre_mlp_norm.weight"], eps)
    return x + mlp(h, P[p + "mlp.wi_fused.weight"], P[p + "mlp.wo.weight"])


def embed_sum(P, cfg, ids: torch.Tensor) -> torch.Tensor:
    """dia/layers.py:691-696 / :737-742: sequential sum over the 9 codebooks."""
    x = None
    for c in range(cfg.data.channels):
        e = F.embedding(ids[..., c], P[f"decoder.embeddings.{c}.weight"])
        x = e if x is None else x + e
    return x


def decoder_forward(P, cfg, ids: torch.Tensor, st: DecState, prefill: bool, dead_cross_kv: bool = True):
    """Decoder.decode_step (prefill=False, T=1) dia/layers.py:671-720 and
    Decoder.forward (prefill=True) :722-766.  Returns fp32 [B, T, C, V]."""
    x = embed_sum(P, cfg, ids)
    for i in range(cfg.model.decoder.n_layer):
        x = decoder_layer(P, cfg, i, x, st, prefill, dead_cross_kv)
    x = rms_norm(x, P["decoder.norm.weight"], cfg.model.normalization_layer_epsilon)
    return dense(x, P["decoder.logits_dense.weight"]).to(torch.float32)


# ----------------------------------------------------------------------------
# logits post-processing and sampling (dia/model.py)
# ----------------------------------------------------------------------------

def cfg_combine_and_mask(cfg, logits_2xCxV: torch.Tensor, cfg_scale: float) -> torch.Tensor:
    """dia/model.py:447-478: guided = cond + s*(cond-uncond) with row0=uncond,
    row1=cond; EOS forbidden on channels>0; PAD and BOS forbidden everywhere;
    ids >= vocab masked only if vocab <= EOS+1 (false for 1028: id 1027 stays
    legal)."""
    uncond, cond = logits_2xCxV[0], logits_2xCxV[1]
    out = cond + cfg_scale * (cond - uncond)
    dt = cfg.data
    if out.shape[0] > 1:
        out[1:, dt.audio_eos_value] = -torch.inf
    out[:, dt.audio_pad_value] = -torch.inf
    out[:, dt.audio_bos_value] = -torch.inf
    if cfg.model.tgt_vocab_size <= dt.audio_eos_value + 1:
        out[:, cfg.model.tgt_vocab_size:] = -torch.inf
    return out


def filtered_probs(logits: torch.Tensor, temperature: float, top_p: float, top_k: int | None) -> torch.Tensor:
    """_sample_next_token up to the final softmax, dia/model.py:43-73."""
    logits = logits / temperature
    if top_k is not None and top_k > 0:
        vals, _ = torch.topk(logits, k=top_k, dim=-1)
        logits = logits.masked_fill(logits < vals[..., -1].unsqueeze(-1), -torch.inf)
    if top_p < 1.0:
        probs = torch.softmax(logits, dim=-1)
        sp, si = torch.sort(probs, dim=-1, descending=True)
        remove = torch.cumsum(sp, dim=-1) > top_p
        remove = torch.roll(remove, shifts=1, dims=-1)
        remove[..., 0] = False
        remove = torch.zeros_like(remove).scatter(dim=-1, index=si, src=remove)
        logits = logits.masked_fill(remove, -torch.inf)
    return torch.softmax(logits, dim=-1)


def sample_next_token(logits: torch.Tensor, temperature: float, top_p: float, top_k: int | None,
                      generator: torch.Generator | None = None) -> torch.Tensor:
    """_sample_next_token, dia/model.py:32-82."""
    if temperature == 0.0:
        return torch.argmax(logits, dim=-1)
    probs = filtered_probs(logits, temperature, top_p, top_k)
    if torch.all(torch.isclose(probs.sum(dim=-1), torch.tensor(0.0))):
        return torch.argmax(logits, dim=-1)
    return torch.multinomial(probs, num_samples=1, generator=generator).squeeze(-1)


# ----------------------------------------------------------------------------
# token-grid preparation (dia/model.py) - uses the numpy delay oracle
# ----------------------------------------------------------------------------

def encode_text(cfg, text: str) -> torch.Tensor:
    """Dia._prepare_text_input, dia/model.py:254-289."""
    raw = text.encode("utf-8").replace(b"[S1]", b"\x01").replace(b"[S2]", b"\x02")
    toks = list(raw)[: cfg.data.text_length]
    out = torch.full((1, cfg.data.text_length), cfg.data.text_pad_value, dtype=torch.long)
    out[0, : len(toks)] = torch.tensor(toks, dtype=torch.long)
    return out


def effective_text(text: str, audio_prompt_text: str | None) -> str:
    """Closing-tag heuristic of Dia.generate, dia/model.py:686-696."""
    t = audio_prompt_text.strip() + " " + text.strip() if audio_prompt_text else text.strip()
    s1, s2 = t.rfind("[S1]"), t.rfind("[S2]")
    if s1 > s2 and not t.endswith("[S2]"):
        t += " [S2]"
    elif s2 > s1 and not t.endswith("[S1]"):
        t += " [S1]"
    elif s1 == -1 and s2 == -1 and t:
        t += " [S2]"
    return t


def prepare_audio_prompt(cfg, prompt: torch.Tensor | None) -> tuple[torch.Tensor, int]:
    """Dia._prepare_audio_prompt, dia/model.py:291-353: [BOS] + prompt + PAD x
    max_delay, then the delay pattern.  Returns (delayed [T,C] int32, prefill_step)."""
    from . import delay_oracle

    dt = cfg.data
    rows = [torch.full((1, dt.channels), dt.audio_bos_value, dtype=torch.int32)]
    step = 1
    if prompt is not None:
        if prompt.ndim == 3 and prompt.shape[0] == 1:
            prompt = prompt.squeeze(0)
        assert prompt.ndim == 2
        step += prompt.shape[0]
        rows.append(prompt.to(torch.int32))
    rows.append(torch.full((max(dt.delay_pattern), dt.channels), dt.audio_pad_value, dtype=torch.int32))
    grid = torch.cat(rows, dim=0)
    delayed = delay_oracle.apply_audio_delay(grid.unsqueeze(0).numpy(), dt.audio_pad_value, dt.audio_bos_value,
                                             dt.delay_pattern)
    return torch.from_numpy(delayed[0]), step


def finalize_codes(cfg, generated: torch.Tensor, codebook_size: int = 1024) -> torch.Tensor:
    """Token half of Dia._generate_output, dia/model.py:504-533: revert the
    delay, drop the last max_delay rows, zero codes outside [0, size-1],
    transpose to [1, C, T].  (The DAC decode that follows is third-party and
    out of scope.)"""
    from . import delay_oracle

    dt = cfg.data
    T = generated.shape[0]
    rev = delay_oracle.revert_audio_delay(generated.unsqueeze(0).numpy(), dt.audio_pad_value, dt.delay_pattern, T)
    rev = torch.from_numpy(rev)[:, : T - max(dt.delay_pattern), :].clone()
    rev[(rev < 0) | (rev > codebook_size - 1)] = 0
    return rev.transpose(1, 2)


# ----------------------------------------------------------------------------
# the generate loop (dia/model.py:631-846; semantics in SURVEY Appendix C)
# ----------------------------------------------------------------------------

@dataclass
class GenTrace:
    codes: torch.Tensor | None = None            # generated_tokens[prefill_step : dec_step+1]  (pre-revert)
    grid: torch.Tensor | None = None             # whole token grid [audio_length, C]
    prefill_step: int = 0
    last_step: int = 0
    logits: dict = field(default_factory=dict)   # step -> fp32 [2, C, V] raw decoder logits
    guided: dict = field(default_factory=dict)   # step -> fp32 [C, V] after CFG + masks
    preds: list = field(default_factory=list)    # raw pred_C per step (before EOS forcing)
    margins: list = field(default_factory=list)  # per-step top1-top2 of guided logits, [C]
    step_seconds: list = field(default_factory=list)


def prepare_generation(P, cfg, text: str, prompt: torch.Tensor | None, dead_cross_kv: bool = True):
    """Dia._prepare_generation, dia/model.py:355-427 (patches B2/B3)."""
    dt, d = cfg.data, cfg.model.decoder
    cond = encode_text(cfg, text)
    enc_in = torch.cat([torch.full_like(cond, dt.text_pad_value), cond], dim=0)
    delayed, prefill_step = prepare_audio_prompt(cfg, prompt)

    positions = torch.arange(dt.text_length, dtype=torch.float32).unsqueeze(0).expand(2, -1)   # state.py:57-59
    pad_mask = enc_in != dt.text_pad_value                                                     # state.py:60
    enc_out = encoder_forward(P, cfg, enc_in, positions, create_attn_mask(pad_mask, pad_mask))
    cross = precompute_cross_kv(P, cfg, enc_out, positions)
    cross_mask = create_attn_mask(torch.ones((2, 1), dtype=torch.bool), pad_mask)               # state.py:139-140
    st = DecState(enc_out=enc_out, enc_positions=positions,
                  dec_positions=torch.zeros((2, 1), dtype=torch.int32), cross_mask=cross_mask,
                  self_cache=[KV(d.kv_heads, dt.audio_length, d.gqa_head_dim) for _ in range(d.n_layer)],
                  cross_cache=cross)
    grid = torch.full((dt.audio_length, dt.channels), -1, dtype=torch.int32)                    # state.py:177-188
    grid[: delayed.shape[0]] = delayed                                                          # state.py:205-208
    if prefill_step > 1:
        st.prepare_step(0, prefill_step - 1)
        toks = grid[0: prefill_step - 1].unsqueeze(0).expand(2, -1, -1)
        decoder_forward(P, cfg, toks, st, prefill=True, dead_cross_kv=dead_cross_kv)
    return st, grid, prefill_step


def generate(P, cfg, text: str, max_tokens: int | None = None, cfg_scale: float = 3.0, temperature: float = 1.3,
             top_p: float = 0.95, cfg_filter_top_k: int = 35, audio_prompt: torch.Tensor | None = None,
             audio_prompt_text: str | None = None, seed: int | None = None, dead_cross_kv: bool = True,
             keep_logits_at: set | None = None, teacher: torch.Tensor | None = None,
             time_steps: bool = False) -> GenTrace:
    """Dia.generate up to (not including) the DAC decode, dia/model.py:631-846.

    ``teacher``: optional int32 token grid; when given, after each step the
    oracle's own prediction is still recorded in ``preds`` but the grid row is
    overwritten from ``teacher`` (teacher forcing, used for logits parity after
    a near-tie divergence).
    """
    import time

    if audio_prompt is not None and not audio_prompt_text:
        raise ValueError("`audio_prompt_text` is required when `audio_prompt` is provided.")
    gen = None
    if seed is not None:
        gen = torch.Generator().manual_seed(seed)
    dt = cfg.data
    eos, pad, delays = dt.audio_eos_value, dt.audio_pad_value, dt.delay_pattern
    D = max(delays)
    max_tokens = dt.audio_length if max_tokens is None else max_tokens
    with torch.inference_mode():
        st, grid, P0 = prepare_generation(P, cfg, effective_text(text, audio_prompt_text), audio_prompt, dead_cross_kv)
        tr = GenTrace(prefill_step=P0)
        dec_step, bos_cd, eos_seen, eos_cd = P0 - 1, D, False, -1
        while dec_step < max_tokens - 1:
            t0 = time.perf_counter()
            cur = dec_step + 1
            st.prepare_step(cur)
            toks = grid[cur - 1].unsqueeze(0).unsqueeze(0).expand(2, 1, -1)
            logits = decoder_forward(P, cfg, toks, st, prefill=False, dead_cross_kv=dead_cross_kv)[:, -1]
            guided = cfg_combine_and_mask(cfg, logits, cfg_scale)
            pred = sample_next_token(guided, temperature, top_p, cfg_filter_top_k, gen)
            if keep_logits_at is not None and cur in keep_logits_at:
                tr.logits[cur] = logits.clone()
                tr.guided[cur] = guided.clone()
            top2 = torch.topk(guided, 2, dim=-1).values
            tr.margins.append((top2[:, 0] - top2[:, 1]).clone())
            tr.preds.append(pred.clone())
            pred = pred.clone()
            if not eos_seen and pred[0] == eos:
                eos_seen, eos_cd = True, D
            if eos_cd > 0:
                s = D - eos_cd
                for c, dl in enumerate(delays):
                    if s == dl:
                        pred[c] = eos
                    elif s > dl and pred[c] != eos:
                        pred[c] = pad
                eos_cd -= 1
            bos_cd = max(0, bos_cd - 1)
            row = pred.to(torch.int32)
            if bos_cd > 0:
                grid[cur] = torch.where(grid[cur] == -1, row, grid[cur])
            else:
                grid[cur] = row
            if teacher is not None:
                grid[cur] = teacher[cur]
            if time_steps:
                tr.step_seconds.append(time.perf_counter() - t0)
            if eos_cd == 0:
                break
            if cur >= max_tokens - D - 1 and not eos_seen:
                eos_seen, eos_cd = True, D
            dec_step += 1
        tr.grid, tr.last_step = grid, dec_step
        tr.codes = grid[P0: dec_step + 1].clone()
    return tr
