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
    """(rank, world) this process shards its images by.  An initialised process group wins; otherwise the
    RANK / WORLD_SIZE variables count only when the process was started by torchrun (TORCHELASTIC_RUN_ID) or
    sharding was asked for explicitly (FOHO_B200_SHARD=1) -- a stray RANK inherited from the environment
    ``foho.main`` runs in must not make a stage silently skip images."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    if "TORCHELASTIC_RUN_ID" in os.environ or os.environ.get("FOHO_B200_SHARD", "0") not in ("0", "", "false", "False"):
        return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    return 0, 1


def default_device() -> str:
    """One process per GPU: ``cuda:$LOCAL_RANK`` (cuda:0 outside torchrun)."""
    return f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}"


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


def parse_cpulist(text: str) -> List[int]:
    """'0-3,8,10-11' -> [0,1,2,3,8,10,11] (the format of /sys/devices/system/node/node*/cpulist)."""
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def bind_to_gpu_numa(device_index: int, sysfs: str = "/sys") -> Dict[str, object]:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host buffer
    is allocated, so that every rank's H2D/D2H staging memory is local to its GPU's PCIe root (one
    process per GPU; first-touch places the pages).  Best effort: returns what it found and did."""
    info: Dict[str, object] = {"numa_node": None, "bound": False}
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(os.path.join(sysfs, "bus/pci/devices", bdf, "numa_node")) as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open(os.path.join(sysfs, f"devices/system/node/node{node}/cpulist")) as f:
            cpus = parse_cpulist(f.read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["bound"] = True
            info["cpus"] = len(allowed)
    except Exception as e:      # no sysfs entry, no permission, old torch: stay unbound
        info["error"] = type(e).__name__
    return info
