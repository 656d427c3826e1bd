"""Multi-GPU plumbing: groups are independent regressions, so the path shards by GROUP with no
data-path collective.  One process per GPU (torch.distributed, NCCL over NVLink on the
GPU box, gloo in the CPU tests); the only exchange is the final gather of the per-rank coefficient (or
prediction) chunks."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_groups(offsets: np.ndarray, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous group ranges [g0, g1) per rank, balanced by cumulative ROW count (the streaming
    kernel's cost is bytes, not groups).  Every group lands on exactly one rank; ranks may be empty."""
    offsets = np.asarray(offsets, dtype=np.int64)
    G = len(offsets) - 1
    total = int(offsets[-1])
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        g = int(np.searchsorted(offsets, target, side="left"))
        if g > 0 and g <= G and abs(offsets[g - 1] - target) <= abs(offsets[min(g, G)] - target):
            g -= 1  # nearest group boundary
        g = min(max(g, bounds[-1]), G)
        bounds.append(g)
    bounds.append(G)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def gather_group_results(local, shards: List[Tuple[int, int]], group=None, out=None):
    """all-gather of per-rank [g_local, k] chunks into the full [G, k] tensor on every rank.
    Equal shards: ONE all_gather_into_tensor straight into `out` (pre-allocated by steady-state callers).
    Ragged shards are padded to the largest one (a single fixed-size all_gather_into_tensor — over
    NVLink/NVSwitch with NCCL — then compacted)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    k = local.shape[1]
    gmax = max(b - a for a, b in shards)
    if all(b - a == gmax for a, b in shards):
        if out is None:
            out = torch.empty((world * gmax, k), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.zeros((gmax, k), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * gmax, k), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    out = out.view(world, gmax, k)
    return torch.cat([out[r, : shards[r][1] - shards[r][0]] for r in range(world)], dim=0)
