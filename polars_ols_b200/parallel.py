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

    def __init__(self, engine, total_groups: int, n_coef: int, group_base: int, group=None, n_buffers: int = 1):
        import torch.distributed as dist
        self.engine, self.total_groups, self.n_coef, self.group_base = engine, total_groups, n_coef, group_base
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise ValueError("peer gather supports up to 8 ranks (one NVSwitch domain)")
        self.nbytes = total_groups * n_coef * 8
        self.n_buffers = n_buffers                                  # > 1: gathers of consecutive steps alternate buffers
        self.buf_stride = (self.nbytes + 255) // 256 * 256
        self.flag_off = self.buf_stride * n_buffers                 # 8 uint64 completion flags behind the coefficient rows
        self.local = engine.device_alloc(self.flag_off + 64)
        engine.zero_device(self.local + self.flag_off, 64)
        self.step = 0
        handles = [None] * self.world
        dist.all_gather_object(handles, engine.ipc_export(self.local), group=group)
        self.peers = [self.local if r == self.rank else engine.ipc_open(handles[r]) for r in range(self.world)]
        dist.barrier(group)
        self.group = group

    def attach(self, buffer: int = 0):
        self.engine.set_peer_gather([p + buffer * self.buf_stride for p in self.peers], self.group_base, self.total_groups)
        if not getattr(self, "_flags_set", False):
            self.engine.set_peer_flags([p + self.flag_off for p in self.peers], self.rank)
            self._flags_set = True

    def step_signal(self):
        """behind a step's kernels: tell every rank that this rank's rows of the step have been stored"""
        self.step += 1
        self.engine.peer_step_signal(self.step)
        return self.step

    def step_signal_wait_previous(self):
        """one launch behind a step's kernels: signal this step, wait until every rank has signalled the PREVIOUS one"""
        self.step += 1
        self.engine.peer_step_signal_wait(self.step, self.step - 1)
        return self.step

    def arm_step(self):
        """BEFORE a step's coefficient call: the fused-gather kernel itself signals this step and waits for the previous one
        (its last solver warp; no extra launch)"""
        self.step += 1
        self.engine.peer_arm_step(self.step, self.step - 1)
        return self.step

    def step_wait(self, step: int):
        """behind later work: the stream continues only when EVERY rank has signalled `step` (its buffer is complete)"""
        self.engine.peer_step_wait(step)

    def step_complete(self):
        """enqueue the per-step completion behind the step's kernels: when it retires, EVERY rank's rows of this step are
        in this rank's buffer (release / acquire flags over NVLink, no collective)"""
        self.step += 1
        self.engine.peer_step_complete(self.step)

    def detach(self):
        self.engine.set_peer_gather([], 0, 0)
        self.engine.set_peer_flags([], 0)
        self._flags_set = False

    def read(self, buffer: int = 0) -> np.ndarray:
        out = np.empty((self.total_groups, self.n_coef))
        self.engine.copy_to_host(out, self.local + buffer * self.buf_stride)
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


# ----------------------------------------------------------------------------------------------------
# time-axis sharding of ONE long series (rls / rolling_ols; SURVEY.md §8e "Single long series")
# ----------------------------------------------------------------------------------------------------
# rank r owns rows [r0, r1) of the series.
#   rolling_ols : every output depends on a bounded history, so a rank also LOADS a halo of earlier rows
#                 [h, r0) and drops their outputs — no exchange at all.  `rolling_halo_start` picks h so that
#                 the rows >= r0 come out exactly as in the whole series, including the reference's warm-up /
#                 min_periods / null-window rules (src/least_squares.rs:848-1032, restated in
#                 csrc/moving_core.cuh); it degrades to h = 0 (load everything) when no finite halo is exact.
#   rls         : the state is an affine recurrence in information form, so each rank computes
#                 the map of its own rows (b200ols_recursive_least_squares_state, zero entering state), the
#                 ranks all-gather these k*k + k + 1 doubles — the path's one real exchange step — and each
#                 rank composes the maps of the ranks before it into the state ENTERING its shard.  Exact for
#                 every forgetting factor, lambda = 1 included.
ROW_ALIGN = 64  # shard / halo starts are multiples of 64 rows: 16-byte aligned columns, byte-aligned bitmaps


def shard_rows(n_rows: int, world_size: int, align: int = ROW_ALIGN) -> List[Tuple[int, int]]:
    """Contiguous row ranges [r0, r1) per rank, equal up to the alignment; ranks may be empty."""
    bounds = [min(n_rows, (n_rows * r // world_size) // align * align) for r in range(world_size)] + [n_rows]
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def _bitmap_rows(bitmap, a: int, b: int) -> np.ndarray:
    """bool validity of rows [a, b) from an Arrow bitmap (numpy uint8 or torch uint8 tensor, offset 0)."""
    lo, hi = a // 8, (b + 7) // 8
    raw = bitmap[lo:hi]
    if not isinstance(raw, np.ndarray):
        raw = raw.cpu().numpy()
    bits = np.unpackbits(raw, bitorder="little")
    return bits[a - lo * 8: b - lo * 8].astype(bool)


def fit_mask_columns(batch, null_policy: str):
    """columns whose nulls invalidate a row for rls / rolling (src/expressions.rs:594-701)."""
    if null_policy in ("zero", "ignore"):
        return []
    if null_policy == "drop_y_zero_x":
        return [batch.target]
    return [batch.target, *batch.features]


def valid_rows_fn(cols):
    """-> f(a, b) = bool array of the rows of [a, b) that enter the fit, or None when no column has nulls."""
    cols = [c for c in cols if c.validity is not None]
    if not cols:
        return None

    def f(a: int, b: int) -> np.ndarray:
        m = _bitmap_rows(cols[0].validity, a, b)
        for c in cols[1:]:
            m &= _bitmap_rows(c.validity, a, b)
        return m

    return f


def rolling_halo_start(valid, start: int, window: int, min_periods: int, fixed_window: bool,
                       align: int = ROW_ALIGN) -> int:
    """First row h <= start a rank must load so that rolling_ols over rows [h, ...) reproduces, for every row
    >= start, what the whole series gives.  `valid`: f(a, b) -> bool array, or None (no nulls)."""
    if start <= 0:
        return 0
    W, mp = int(window), int(min_periods)
    ones = lambda a, b: np.ones(b - a, dtype=bool)  # noqa: E731
    v_of = valid or ones
    if fixed_window:
        # rows of a warm-up longer than the window are never subtracted by the reference (the state keeps them
        # for the rest of the series): no finite halo reproduces that
        cnt, pos, mpv = 0, 0, None
        while pos < start and mpv is None:
            blk = v_of(pos, min(start, pos + (1 << 20)))
            c = np.cumsum(blk) + cnt
            hit = np.nonzero(c >= mp)[0]
            if len(hit):
                mpv = pos + int(hit[0]) + 1
            cnt, pos = int(c[-1]) if len(c) else cnt, pos + len(blk)
        if mpv is None or mpv > W:
            return 0
    H = (max(W, mp) + (W + 1 if fixed_window else 0) + align - 1) // align * align
    while True:
        h = max(0, (start - H) // align * align)
        if h == 0:
            return 0
        v = v_of(h, start)
        if not fixed_window:
            # window = last W valid rows: the halo must hold W of them (and finish the local warm-up)
            if int(v.sum()) >= max(W, mp):
                return h
        else:
            n = start - h
            c = np.concatenate([[0], np.cumsum(v)])
            ok = n >= W + 1 and c[min(W, n)] >= mp          # saturated window; local warm-up inside its first W rows
            if ok:
                # a row j (local index >= W) whose coefficients the reference refreshes: a window change with
                # enough valid rows in (j - W, j] — the forward-fill source of the rows that follow
                j = np.arange(W, n)
                cntj = c[j + 1] - c[j + 1 - W]
                ok = bool(np.any((v[j] | v[j - W]) & (cntj >= mp)))
            if ok:
                return h
        H *= 2


def rls_prior_information(n_coef: int, initial_state_covariance: float, initial_state_mean=None) -> np.ndarray:
    """(A0, b0) = (I / p0, theta0 / p0) flattened [n_coef^2 + n_coef] (src/least_squares.rs:505-511)."""
    p0 = 10.0 if initial_state_covariance is None else float(initial_state_covariance)
    a0 = np.eye(n_coef) / p0
    th = np.zeros(n_coef) if initial_state_mean is None else np.broadcast_to(np.asarray(initial_state_mean, float), (n_coef,))
    return np.concatenate([a0.reshape(-1), th / p0])


def rls_entering_state(maps: np.ndarray, prior: np.ndarray, rank: int) -> np.ndarray:
    """maps[q] = [A_q, b_q, D_q] of shard q (zero entering state); state entering shard `rank` =
    composition of the maps of shards 0..rank-1 applied to the prior."""
    st = np.array(prior, dtype=np.float64)
    for q in range(rank):
        st = maps[q][-1] * st + maps[q][:-1]
    return st


def slice_col(c, a: int, b: int):
    """rows [a, b) of a Col as a view (a must be a multiple of 8 when the column has a bitmap)."""
    from .engine import Col
    validity = None
    if c.validity is not None:
        assert a % 8 == 0
        validity = c.validity[a // 8: (b + 7) // 8]
    return Col(c.values[a:b], validity)


def slice_batch(batch, a: int, b: int):
    from .engine import Batch
    return Batch(slice_col(batch.target, a, b), [slice_col(f, a, b) for f in batch.features],
                 None if batch.weights is None else slice_col(batch.weights, a, b), batch.add_intercept)


def _all_gather_object(obj, group=None):
    import torch.distributed as dist
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, obj, group=group)
    return out


def _all_gather_array(arr: np.ndarray, group=None):
    """the path's one exchange for a time-sharded rls: k*k + k + 1 doubles per rank.  One tensor all-gather (NCCL on the
    GPU box, gloo in the CPU tests) instead of a pickled object collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    a = np.ascontiguousarray(arr, dtype=np.float64)
    t = torch.as_tensor(a.reshape(-1), device=dev)
    out = torch.empty(world * t.numel(), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, t, group=group)
    return list(out.cpu().numpy().reshape((world,) + a.shape))


def time_sharded(expr, frame, rank: int, world_size: int, engine=None, exchange=None, group=None):
    """Evaluate an rls / rolling_ols expression (no `.over()`) on this rank's time shard of one long series.

    `frame` holds the whole series (host arrays or CUDA tensors; only this rank's rows and halo are touched).
    Returns (Result for rows [r0, r1), (r0, r1)).  `exchange(obj) -> list over ranks` defaults to
    torch.distributed.all_gather_object (NCCL or gloo); assemble the full column with `gather_rows`."""
    from . import _lib as L
    from .engine import get_engine
    from .least_squares import Result

    if expr.kind not in ("recursive_least_squares", "rolling_least_squares") or expr._over:
        raise ValueError("time_sharded evaluates rls / rolling_ols expressions without .over()")
    b, names = expr.batch(frame)
    n = len(b.target)
    r0, r1 = shard_rows(n, world_size)[rank]
    if engine is None:
        dev = b.target.values.device.index if b.target.is_device else 0
        engine = get_engine(dev or 0, torch_stream=b.target.is_device)
    mode = L.MODE[expr.mode]
    F = len(names)
    kw = expr.kwargs
    if expr.kind == "rolling_least_squares":
        fixed = kw.null_policy not in ("drop", "drop_zero", "drop_y_zero_x")   # src/least_squares.rs:947-950
        mp = min(F, kw.window_size) if kw.min_periods is None else int(kw.min_periods)
        h = rolling_halo_start(valid_rows_fn(fit_mask_columns(b, kw.null_policy)), r0, kw.window_size, mp, fixed)
        v, m = engine.rolling_least_squares(slice_batch(b, h, r1), kw.to_c(), mode)
        v, m = v[r0 - h:], (None if m is None else m[r0 - h:])
    else:
        sub = slice_batch(b, r0, r1)
        zero = np.zeros((1, F * F + F))
        ckw, keep = expr.rls_c_kwargs(F, zero)
        local = engine.recursive_least_squares_state(sub, ckw, keep)[0]          # this shard's affine map
        maps = np.stack((exchange or (lambda o: _all_gather_array(o, group)))(local))
        info = None
        if rank > 0:
            # quirk A.5.1: the prior mean is honoured in coefficients mode only (src/expressions.rs:604-610 vs :636)
            mean = kw.initial_state_mean if expr.mode == "coefficients" else None
            prior = rls_prior_information(F, kw.initial_state_covariance, mean)
            info = rls_entering_state(maps, prior, rank)[None, :]
        ckw, keep = expr.rls_c_kwargs(F, info)
        v, m = engine.recursive_least_squares(sub, ckw, mode, keep)
    fields = names if expr.mode == "coefficients" else None
    return Result(expr.output_name, v, m, fields), (r0, r1)


def gather_rows(local, shards: List[Tuple[int, int]], group=None):
    """all-gather of per-rank row chunks ([n_local] or [n_local, k] tensors) into the full column."""
    flat = local.reshape(local.shape[0], -1)
    out = gather_group_results(flat, shards, group)
    return out.reshape((out.shape[0],) + tuple(local.shape[1:]))
