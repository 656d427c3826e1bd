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


class PeerGather:
    """Fused gather of per-rank coefficient chunks through NVLink peer memory (no collective per step).

    Every rank owns a full [total_groups, n_coef] f64 buffer (cudaMalloc, exported with CUDA IPC); after
    `attach()` the engine's `least_squares_coefficients` on a DEVICE frame stores rows
    [group_base, group_base + n_groups) into EVERY rank's buffer straight from the solver warps
    (`b200ols_set_peer_gather`).  A rank may read its buffer after all ranks synchronised (stream sync +
    barrier), exactly as after a collective."""

    def __init__(self, engine, total_groups: int, n_coef: int, group_base: int, group=None):
        import torch.distributed as dist
        self.engine, self.total_groups, self.n_coef, self.group_base = engine, total_groups, n_coef, group_base
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise ValueError("peer gather supports up to 8 ranks (one NVSwitch domain)")
        self.nbytes = total_groups * n_coef * 8
        self.local = engine.device_alloc(self.nbytes)
        handles = [None] * self.world
        dist.all_gather_object(handles, engine.ipc_export(self.local), group=group)
        self.peers = [self.local if r == self.rank else engine.ipc_open(handles[r]) for r in range(self.world)]
        dist.barrier(group)
        self.group = group

    def attach(self):
        self.engine.set_peer_gather(self.peers, self.group_base, self.total_groups)

    def detach(self):
        self.engine.set_peer_gather([], 0, 0)

    def read(self) -> np.ndarray:
        out = np.empty((self.total_groups, self.n_coef))
        self.engine.copy_to_host(out, self.local)
        return out

    def close(self):
        import torch.distributed as dist
        self.detach()
        self.engine.synchronize()
        dist.barrier(self.group)
        for r, p in enumerate(self.peers):
            if r != self.rank:
                self.engine.ipc_close(p)
        dist.barrier(self.group)
        self.engine.device_free(self.local)
