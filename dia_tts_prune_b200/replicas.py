"""Multi-GPU = replicas over independent utterances: one process per GPU, utterance
i goes to rank i % world, no collective on the data path.  The only communication is the end-of-run
reduction of (frames, seconds) used for the throughput line, over NCCL on GPUs or gloo on CPU."""

from __future__ import annotations

import torch
import torch.distributed as dist


def assign(n_items: int, world: int, rank: int) -> list[int]:
    """Static round-robin shard of utterance indices."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world / rank")
    return list(range(rank, n_items, world))


def reduce_throughput(frames: float, seconds: float, device=None, group=None) -> tuple[float, float]:
    """(sum of frames over ranks, max of seconds over ranks); identity when not distributed."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(frames), float(seconds)
    t = torch.tensor([float(frames), float(seconds)], dtype=torch.float64, device=device)
    s, m = t.clone(), t.clone()
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    return s[0].item(), m[1].item()


def gather_codes(codes: list[torch.Tensor], group=None) -> list:
    """End-of-run gather of every rank's (index, code tensor) pairs onto all ranks (host objects)."""
    if not (dist.is_available() and dist.is_initialized()):
        return [codes]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, [c.cpu() if isinstance(c, torch.Tensor) else c for c in codes], group=group)
    return out
