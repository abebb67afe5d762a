"""Multi-GPU plumbing.  The path shards by image with no data-path collective (the reference's
only scale-out is a SLURM array over image chunks, src/foho/guidance/run.py:178-185): rank r
of W takes ``sorted(images)[r::W]``; one all_gather of timing scalars at the end
(NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import os
from typing import Dict, List, Sequence

import torch
import torch.distributed as dist


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_images(images: Sequence[str], rank: int, world: int) -> List[str]:
    """Deterministic, disjoint, exhaustive: sorted(images)[rank::world]."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("invalid rank/world")
    return sorted(images)[rank::world]


def task_chunk(chunks: Sequence[Sequence[str]], rank: int) -> List[str]:
    """The reference's task_list_file semantics with SLURM_ARRAY_TASK_ID -> rank (run.py:178-183)."""
    return list(chunks[rank])


def gather_timings(local: Dict[str, float], device=None) -> List[Dict[str, float]]:
    """all_gather of a small dict of floats (same keys on every rank)."""
    keys = sorted(local)
    if not (dist.is_available() and dist.is_initialized()):
        return [dict(local)]
    t = torch.tensor([float(local[k]) for k in keys], dtype=torch.float64, device=device or "cpu")
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [dict(zip(keys, o.tolist())) for o in out]


def aggregate_throughput(per_rank: List[Dict[str, float]]) -> float:
    """Whole-job units/s: all units processed divided by the slowest rank's time."""
    units = sum(r["units"] for r in per_rank)
    secs = max(r["seconds"] for r in per_rank)
    return units / secs if secs > 0 else 0.0
