"""Seeded synthetic weights and transcripts (there is no network for checkpoints).

The weight recipe is the one SURVEY.md section 8(d) fixes, shared by the oracle runs
that produced tests/golden/ and by every GPU test / bench run: iterate the
parameters in ``named_parameters()`` order with ONE seeded CPU generator,
RMSNorm weights = 1, embeddings ~ N(0,1), dense kernels ~ N(0, fan_in^-1/2),
then round every non-norm tensor to a bf16-representable fp32 value - so the
bf16 weight stream the kernels read holds exactly the oracle's numbers.
"""

from __future__ import annotations

import hashlib

import torch


def is_dense_kernel(name: str) -> bool:
    """True for DenseGeneral kernels (the tensors that may be stored bf16); norms and embeddings stay fp32."""
    return name.endswith(("_proj.weight", "mlp.wi_fused.weight", "mlp.wo.weight", "logits_dense.weight"))


def cast_dense_kernels_(model, dtype) -> None:
    """Store every DenseGeneral kernel of ``model`` in ``dtype`` (what Dia(config, "bfloat16") does)."""
    for n, p in model.named_parameters():
        if is_dense_kernel(n):
            p.data = p.data.to(dtype)


def fan_in_of(name: str, shape) -> int:
    # o_proj kernels contract (heads, head_dim); every other kernel contracts axis 0
    return int(shape[0]) * int(shape[1]) if "o_proj" in name else int(shape[0])


@torch.no_grad()
def init_synthetic_(named_params, seed: int = 1234) -> None:
    """In-place init of an iterable of (name, fp32 CPU tensor) in the given order."""
    g = torch.Generator().manual_seed(seed)
    for name, p in named_params:
        if "norm" in name:
            p.fill_(1.0)
            continue
        if "embedding" in name:
            p.normal_(0.0, 1.0, generator=g)
        else:
            p.normal_(0.0, fan_in_of(name, p.shape) ** -0.5, generator=g)
        p.copy_(p.to(torch.bfloat16).to(torch.float32))


def synthetic_state_dict(shapes: dict, seed: int = 1234) -> dict:
    """``shapes``: ordered name -> shape mapping (named_parameters order)."""
    sd = {n: torch.empty(s, dtype=torch.float32) for n, s in shapes.items()}
    init_synthetic_(sd.items(), seed)
    return sd


def weights_fingerprint(sd: dict, names=None) -> str:
    """sha256 over the bf16 bit patterns of a few tensors - lets a GPU box
    verify it regenerated the very weights the golden fixtures were made with."""
    h = hashlib.sha256()
    for n in (names or sorted(sd)):
        t = sd[n].detach().to("cpu", torch.float32).contiguous()
        h.update(n.encode())
        h.update(t.to(torch.bfloat16).view(torch.int16).numpy().tobytes())
    return h.hexdigest()


def synthetic_transcript(i: int, min_bytes: int = 60, max_bytes: int = 200) -> str:
    """Deterministic alternating [S1]/[S2] dialogue of 60-200 bytes (config 5)."""
    words = ("dia is an open weights text to dialogue model you get full control over scripts and voices "
             "the quick brown fox jumps over the lazy dog while rain keeps falling on the quiet harbor town "
             "please remember to bring the blue notebook when we meet at the station tomorrow morning").split()
    state = (i * 2654435761 + 12345) & 0xFFFFFFFF

    def nxt():
        nonlocal state
        state = (state * 1664525 + 1013904223) & 0xFFFFFFFF
        return state >> 8

    target = min_bytes + nxt() % (max_bytes - 12 - min_bytes + 1)      # a word adds at most 11 bytes
    out, spk, left = "[S1]", 1, 4 + nxt() % 6
    while len(out.encode()) < target:
        if left == 0:
            spk = 3 - spk
            out += f". [S{spk}]"
            left = 4 + nxt() % 6
        out += " " + words[nxt() % len(words)]
        left -= 1
    return out + "."


DEFAULT_TRANSCRIPT = ("[S1] Dia is an open weights text to dialogue model. "
                      "[S2] You get full control over scripts and voices. [S1]")
