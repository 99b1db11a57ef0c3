"""Weight producers for the pruned variants (drop-in for the reference's ``dia/pruning_utils.py``).

Thin wrappers over ``torch.nn.utils.prune`` applied to every ``DenseGeneral`` - masks only, shapes
never change (dia/pruning_utils.py:13-179) - plus the 2:4 mask the reference does not have
(SURVEY.md 8(d), config 4).  After ``make_pruning_permanent`` the Decoder's repacked weight copy is invalidated and rebuilt on the next
decode call.  The rebuild physically drops MLP hidden neurons whose contribution is exactly zero
(:func:`plan_mlp_compaction` - the "reduced MLP width" of a ``--prune-dim 0`` checkpoint, offline_prune.py:43):
the decode kernels then stream (and the roofline counts) only the live part of ``wi_fused`` / ``wo``.  Other
pruned zeros stream through the decode kernels like any other value.
"""

from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.utils.prune as prune

from .layers import DenseGeneral

DEFAULT_PRUNABLE_MODULES = (DenseGeneral,)


def get_prunable_modules(model: nn.Module, module_types=DEFAULT_PRUNABLE_MODULES, parameter_name: str = "weight"):
    """[(module, parameter_name)] for every module of the given types that owns the parameter."""
    out = []
    for _, m in model.named_modules():
        if isinstance(m, tuple(module_types)) and getattr(m, parameter_name, None) is not None:
            out.append((m, parameter_name))
    if not out:
        print(f"Warning: no prunable modules of types {module_types} with parameter '{parameter_name}'")
    return out


def apply_unstructured_pruning(model: nn.Module, amount: float, module_types=DEFAULT_PRUNABLE_MODULES):
    """Global L1-magnitude pruning over all selected kernels (dia/pruning_utils.py:42-62)."""
    if not 0.0 <= amount <= 1.0:
        raise ValueError("amount must be in [0, 1]")
    params = get_prunable_modules(model, module_types)
    if params and amount > 0:
        prune.global_unstructured(params, pruning_method=prune.L1Unstructured, amount=amount)
    return model


def apply_structured_pruning(model: nn.Module, amount: float, dim: int = 0, n: int = 2,
                             module_types=DEFAULT_PRUNABLE_MODULES):
    """Per-module Ln structured pruning along ``dim`` (dia/pruning_utils.py:64-119).  With the
    [in..., out...] kernel layout dim=0 zeroes INPUT slices: hidden neurons for ``mlp.wo``."""
    if not 0.0 <= amount <= 1.0:
        raise ValueError("amount must be in [0, 1]")
    for m, name in get_prunable_modules(model, module_types):
        w = getattr(m, name)
        if dim >= w.ndim:
            print(f"Warning: skipping module with {w.ndim}-D weight for dim={dim}")
            continue
        if amount > 0:
            prune.ln_structured(m, name=name, amount=amount, n=n, dim=dim)
    return model


@torch.no_grad()
def mask_2to4(w: torch.Tensor, K: int) -> torch.Tensor:
    """bool mask (shape of ``w``) keeping the 2 largest-magnitude of every 4 consecutive K (input) entries per output
    column of the kernel viewed as [K, N].  Ties keep the lower K index; the comparison runs on integer keys
    (magnitude bits, then position), so the mask does not depend on the machine that computes it."""
    w2 = w.detach().reshape(K // 4, 4, -1).to(torch.float32)
    mag = w2.abs().contiguous().view(torch.int32).to(torch.int64)          # fp32 bit pattern: monotone in |w|
    key = mag * 4 + (3 - torch.arange(4, dtype=torch.int64, device=w.device))[None, :, None]
    keep = key.topk(2, dim=1).indices
    return torch.zeros_like(w2, dtype=torch.bool).scatter_(1, keep, True).reshape(w.shape)


@torch.no_grad()
def apply_2to4_pruning(model: nn.Module, module_types=DEFAULT_PRUNABLE_MODULES):
    """Keep the 2 largest-magnitude of every 4 consecutive K (input) entries per output column of each
    kernel viewed as [K, N]; applied as a custom mask so that
    ``make_pruning_permanent`` / ``check_pruning_sparsity`` treat it like the other modes."""
    for m, name in get_prunable_modules(model, module_types):
        w = getattr(m, name)
        n_in = len(m.in_shapes)
        K = 1
        for s in w.shape[:n_in]:
            K *= s
        if K % 4:
            continue
        prune.custom_from_mask(m, name=name, mask=mask_2to4(w, K).to(w.dtype))
    return model


@torch.no_grad()
def is_2to4(w: torch.Tensor, n_in_axes: int = 1) -> bool:
    """Every 4 consecutive K (input) entries of every output column of the kernel viewed as [K, N] hold at most 2
    non-zeros - what :func:`apply_2to4_pruning` produces and ``mma.sp`` requires."""
    K = 1
    for s_ in w.shape[:n_in_axes]:
        K *= s_
    if K % 4:
        return False
    return bool(((w.reshape(K // 4, 4, -1) != 0).sum(dim=1) <= 2).all().item())


def make_pruning_permanent(model: nn.Module, module_types=DEFAULT_PRUNABLE_MODULES):
    """Fold masks into the weights and drop the reparametrisation (dia/pruning_utils.py:122-151)."""
    for m, name in get_prunable_modules(model, module_types, parameter_name="weight"):
        if prune.is_pruned(m):
            try:
                prune.remove(m, name)
            except ValueError:
                pass
    for m in model.modules():
        if hasattr(m, "invalidate_engine"):
            m.invalidate_engine()
    return model


def check_pruning_sparsity(model: nn.Module, module_types=DEFAULT_PRUNABLE_MODULES) -> float:
    """Global fraction of exact zeros over the selected kernels (dia/pruning_utils.py:153-179)."""
    zeros = total = 0
    for m, name in get_prunable_modules(model, module_types):
        w = getattr(m, name)
        zeros += int((w == 0).sum().item())
        total += w.numel()
    sparsity = zeros / total if total else 0.0
    print(f"Global sparsity: {100.0 * sparsity:.2f}% ({zeros}/{total})")
    return sparsity


# ---- physical compaction of the MLP for the decode engine ---------------------------------------------------------------
def engine_mlp_width(n_live: int) -> int:
    """Smallest hidden width >= ``n_live`` the step kernel accepts (every contraction length splits into 8 warp
    slices of whole 256-row fetch units, or one shorter unit that is a multiple of 64 rows)."""
    n_live = max(int(n_live), 1)
    if n_live <= 2048:
        return -(-n_live // 512) * 512
    return -(-n_live // 2048) * 2048


@torch.no_grad()
def mlp_live_neurons(wi_fused: torch.Tensor, wo: torch.Tensor) -> torch.Tensor:
    """bool [F]: hidden neuron j of ``wo(silu(gate) * up)`` (dia/layers.py:92-105) can be non-zero.  It is dead when
    row j of ``wo`` is all zeros (what structured ``dim=0`` pruning of ``mlp.wo`` produces), or when its gate or
    up column of ``wi_fused [D, 2, F]`` is all zeros (silu(0) * up = gate * 0 = 0)."""
    live_o = (wo != 0).any(dim=1)
    live_g = (wi_fused[:, 0, :] != 0).any(dim=0)
    live_u = (wi_fused[:, 1, :] != 0).any(dim=0)
    return live_o & live_g & live_u


@torch.no_grad()
def plan_mlp_compaction(params: dict[str, torch.Tensor], n_layer: int, n_hidden: int,
                        prefix: str = "layers.") -> tuple[int, list[torch.Tensor]] | None:
    """Decide the hidden width the decode engine is built with.  Returns ``(F_eff, [index tensor per layer])`` -
    the neurons each layer keeps, ascending, padded with dead ones up to the common width - or ``None`` when
    nothing can be dropped.  Dropping a dead neuron removes terms that are exactly zero, so the layer output is
    unchanged up to fp32 summation order."""
    lives = []
    for i in range(n_layer):
        lives.append(mlp_live_neurons(params[f"{prefix}{i}.mlp.wi_fused.weight"], params[f"{prefix}{i}.mlp.wo.weight"]))
    width = engine_mlp_width(max(int(l.sum().item()) for l in lives))
    if width >= n_hidden:
        return None
    keep = []
    for live in lives:
        idx_live = torch.nonzero(live, as_tuple=False).flatten()
        idx_dead = torch.nonzero(~live, as_tuple=False).flatten()
        idx = torch.cat([idx_live, idx_dead[: width - idx_live.numel()]])
        keep.append(torch.sort(idx).values)
    return width, keep


@torch.no_grad()
def compact_mlp(params: dict[str, torch.Tensor], plan: tuple[int, list[torch.Tensor]],
                prefix: str = "layers.") -> dict[str, torch.Tensor]:
    """The parameter dict with every ``mlp.wi_fused [D, 2, F]`` / ``mlp.wo [F, D]`` reduced to the planned neurons."""
    _, keep = plan
    out = dict(params)
    for i, idx in enumerate(keep):
        wi, wo = params[f"{prefix}{i}.mlp.wi_fused.weight"], params[f"{prefix}{i}.mlp.wo.weight"]
        idx = idx.to(wi.device)
        out[f"{prefix}{i}.mlp.wi_fused.weight"] = wi.index_select(2, idx).contiguous()
        out[f"{prefix}{i}.mlp.wo.weight"] = wo.index_select(0, idx).contiguous()
    return out


# ---- K-row compaction: `--prune-dim 0` zeroes whole INPUT rows of a kernel (offline_prune.py:43, dia/pruning_utils.py:64-119) ----
# GEMM families of the decode engine (csrc/engine_internal.h): the kernels that make up each, viewed as [K, N]
K_FAMILIES = {0: ("self_attention.q_proj.weight", "self_attention.k_proj.weight", "self_attention.v_proj.weight"),
              1: ("self_attention.o_proj.weight",), 2: ("cross_attention.q_proj.weight",),
              3: ("cross_attention.o_proj.weight",), 4: ("mlp.wi_fused.weight",)}
K_LOGITS = 6


def _rows_live(w: torch.Tensor, n_in_axes: int) -> torch.Tensor:
    K = 1
    for s_ in w.shape[:n_in_axes]:
        K *= s_
    return (w.reshape(K, -1) != 0).any(dim=1)


@torch.no_grad()
def plan_row_compaction(params: dict[str, torch.Tensor], n_layer: int, prefix: str = "layers.") -> dict:
    """Which all-zero input rows the decode engine can drop, per GEMM family: ``{family: (K_eff, [index tensor per
    layer])}`` for the families worth compacting (the fused q/k/v projection keeps a row that is live in any of the
    three).  ``K_eff`` is the common contraction length the kernel accepts (``engine_mlp_width``); layers with fewer live
    rows are padded with dead ones (zero weights: exact).  mlp-out is handled by :func:`plan_mlp_compaction`."""
    plan = {}
    for fam, names in K_FAMILIES.items():
        n_in = 2 if fam in (1, 3) else 1
        lives = []
        for i in range(n_layer):
            live = None
            for nm in names:
                l_ = _rows_live(params[f"{prefix}{i}.{nm}"], n_in)
                live = l_ if live is None else (live | l_)
            lives.append(live)
        K = lives[0].numel()
        width = engine_mlp_width(max(int(l_.sum().item()) for l_ in lives))
        if width >= K:
            continue
        keep = []
        for live in lives:
            a = torch.nonzero(live, as_tuple=False).flatten()
            b = torch.nonzero(~live, as_tuple=False).flatten()
            keep.append(torch.sort(torch.cat([a, b[: width - a.numel()]])).values)
        plan[fam] = (width, keep)
    live = _rows_live(params["logits_dense.weight"], 1)
    width = engine_mlp_width(int(live.sum().item()))
    if width < live.numel():
        a = torch.nonzero(live, as_tuple=False).flatten()
        b = torch.nonzero(~live, as_tuple=False).flatten()
        plan[K_LOGITS] = (width, [torch.sort(torch.cat([a, b[: width - a.numel()]])).values])
    return plan


@torch.no_grad()
def compact_rows(params: dict[str, torch.Tensor], plan: dict, prefix: str = "layers.") -> tuple[dict, dict]:
    """(parameter dict with the planned kernels reduced to their kept input rows, ``{family: int32 map [layers, K]}``
    with the compacted position of every input element or -1) - what ``dia_b200_set_row_map`` takes."""
    out = dict(params)
    maps = {}
    for fam, (width, keep) in plan.items():
        names = ("logits_dense.weight",) if fam == K_LOGITS else K_FAMILIES[fam]
        n_in = 2 if fam in (1, 3) else 1
        rows = []
        for i, idx in enumerate(keep):
            for nm in names:
                key = nm if fam == K_LOGITS else f"{prefix}{i}.{nm}"
                w = params[key]
                K = 1
                for s_ in w.shape[:n_in]:
                    K *= s_
                w2 = w.reshape(K, *w.shape[n_in:]).index_select(0, idx.to(w.device)).contiguous()
                out[key] = w2                                   # [K_eff, out...]: the engine reads it as [K_eff, N]
            m = torch.full((K,), -1, dtype=torch.int32)
            m[idx.cpu()] = torch.arange(idx.numel(), dtype=torch.int32)
            rows.append(m)
        maps[fam] = torch.stack(rows).contiguous()
    return out, maps
