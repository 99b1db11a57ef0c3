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


def shard(items: list, world: int, rank: int) -> list:
    """This rank's round-robin share of ``items`` (utterance i goes to rank i % world)."""
    return [items[i] for i in assign(len(items), world, rank)]


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


def run_sharded(generate, transcripts: list[str], world: int = 1, rank: int = 0, sync=None,
                frames_of=None) -> tuple[dict, float, float]:
    """BASELINE.json configs[4]: a batch of independent transcripts over the replicas.  Rank ``rank`` runs
    ``generate(text)`` (e.g. ``lambda t: dia.generate(t, output="codes")``) for its round-robin shard, one utterance
    after the other; returns ({global index: result}, frames produced by this rank, seconds this rank took).
    ``sync`` (e.g. ``torch.cuda.synchronize``) is called before each clock read; ``frames_of(result)`` counts the
    frames of one result (default: its leading dimension).  Feed the last two numbers to :func:`reduce_throughput`
    for the whole-job line."""
    import time
    mine = assign(len(transcripts), world, rank)
    out, frames = {}, 0.0
    if sync is not None:
        sync()
    t0 = time.perf_counter()
    for i in mine:
        res = generate(transcripts[i])
        out[i] = res
        if frames_of is not None:
            frames += float(frames_of(res))
        elif res is not None and hasattr(res, "shape"):
            frames += float(res.shape[0])
    if sync is not None:
        sync()
    return out, frames, time.perf_counter() - t0
