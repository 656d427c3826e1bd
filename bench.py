#!/usr/bin/env python
"""bench.py — headline benchmark of the polars_ols hot path on B200 (contract: see the task statement).

Workload (BASELINE.json configs[1], "C2"): ridge(alpha=1e-3) coefficients `.over(group)` on
10,000 groups x 1,000 rows x 8 features, f64.  One "step" = one pass of the hot path over one such
batch (per GPU: weak scaling, groups are independent so ranks shard by group; the only collective is
the final NCCL all-gather of the coefficient chunks).

    python bench.py --gpus 1 --steps 20 --warmup 5            # our arm (CUDA)
    python bench.py --impl reference --steps 5 --warmup 1      # CPU restatement of the reference (oracle port)

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

G, N_PER, K = 10_000, 1_000, 8
ALPHA = 1e-3
METRIC = "group regressions/sec (f64, 10k groups x 1k rows x 8 feat)"
UNIT = "regressions/s"
ALG_BYTES_PER_REGRESSION = N_PER * (K + 1) * 8 + K * 8   # inputs read once + coefficients out
WORKLOAD = "C2: ridge(alpha=1e-3) coefficients .over(group), 10000 groups x 1000 rows x 8 features, f64"


def make_data(seed: int):
    """SURVEY.md §8d generator: X ~ N(0,1), beta_g = 1 + 0.25 N(0,1), y = X beta_g + N(0, 0.1); contiguous groups."""
    rng = np.random.default_rng(seed)
    n = G * N_PER
    x = rng.standard_normal((K, n))
    beta = 1.0 + 0.25 * rng.standard_normal((G, K))
    y = np.einsum("kgn,gk->gn", x.reshape(K, G, N_PER), beta).reshape(-1) + 0.1 * rng.standard_normal(n)
    offsets = np.arange(G + 1, dtype=np.int64) * N_PER
    return x, y, offsets


def bind_to_gpu_numa_node(local_rank: int):
    """N > 1: run this rank's host threads (staging memcpys, pinned allocations by first touch) on the NUMA node its GPU
    hangs off.  Best effort; returns a short description for the JSON line."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text())
        if node < 0:
            return f"gpu {bdf}: no NUMA affinity reported"
        cpus = []
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return f"gpu {bdf} -> NUMA node {node}, {len(allowed)} cpus"
    except Exception as exc:  # noqa: BLE001
        return f"not bound ({type(exc).__name__})"


def make_config(world: int, notes: str = ""):
    """the SAME dict (keys and values) for both arms, so that the driver's config comparison holds"""
    return {"workload": WORKLOAD, "l2": "inputs 720 MB per step > 126 MB L2 (no flush needed)",
            "parallelism": f"groups sharded over {world} GPU(s), weak scaling (every rank a full 10k-group batch)",
            "inputs": "resident in HBM (value) / pinned host memory (e2e)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(2)
            except Exception:
                self.proc.kill()
            self.t.join(1)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_run(x, y, offsets, steps: int, warmup: int, min_seconds: float = 0.0, tune_threads: bool = True):
    """times the CPU restatement of the reference's per-group path (oracle/ols_oracle.c, OpenMP over groups =
    polars' rayon pool) on the host cores.  Each step = the full 10k-group batch."""
    import ctypes as C
    from oracle import lib as oracle_lib
    L = oracle_lib()
    cols = [np.ascontiguousarray(y)] + [np.ascontiguousarray(x[i]) for i in range(K)]
    arr = (C.c_void_p * (K + 1))(*[c.ctypes.data for c in cols])
    out = np.empty((G, K))
    threads = L.orc_max_threads()
    n_threads = [0]  # 0 = the OpenMP default

    def step():
        L.orc_grouped_least_squares_coefficients(arr, K, offsets.ctypes.data, G, 0, ALPHA, 0.0, 1000, 1e-5, 0, n_threads[0], out.ctypes.data)

    if tune_threads:
        # give the CPU arm its best thread count: all logical CPUs vs one thread per physical core (SMT rarely helps
        # this memory-bound loop); 3 untimed batches each, keep the faster
        ncpu = os.cpu_count() or threads
        best = None
        for cand in sorted({ncpu, max(1, ncpu // 2), threads}, reverse=True):
            n_threads[0] = cand
            step()
            t0 = time.perf_counter()
            for _ in range(3):
                step()
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, cand)
        n_threads[0] = threads = best[1]
    for _ in range(warmup):
        step()
    times = []
    t_all = time.perf_counter()
    while len(times) < steps or (time.perf_counter() - t_all) < min_seconds:
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if len(times) >= 10_000:
            break
    return times, threads, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host->device->host leg (default min(steps, 10))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-every", type=int, default=4,
                    help="event-time the Gram kernel on every n-th step of the timed region; the event pair costs ~5 us of a 0.12 ms step")
    ap.add_argument("--tile-rows", type=int, default=0)
    ap.add_argument("--warps", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C4", "C4rls", "C5"],
                    help="BASELINE.json config; C2 (default) is the metric's own configuration, the others print the same line "
                         "shape through tools/bench_configs.py (N = 1)")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N > 1: fused P2P stores from the solver warps (default) or an NCCL all-gather per step")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if a.config != "C2":   # the other BASELINE.json configs: single GPU, same line shape
        if rank != 0:
            return 0
        if a.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the CPU port of this config is timed inside the GPU arm's "
                              "cpu_baseline (python bench.py --config %s)" % a.config}))
            return 0
        sys.path.insert(0, str(ROOT / "tools"))
        from bench_configs import run_config
        print(json.dumps(run_config(a.config, steps=a.steps if a.config != "C1" else max(a.steps, 200), warmup=a.warmup,
                                    with_cpu=not a.no_cpu_baseline)))
        return 0

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if a.impl == "reference":
        if rank != 0:
            return 0
        x, y, offsets = make_data(0)
        times, threads, _ = cpu_port_run(x, y, offsets, a.steps, a.warmup)
        t = sum(times)
        val = G * len(times) / t
        line = {
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": len(times),
            "warmup": a.warmup, "ms_per_step": 1e3 * t / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(a.gpus),
            "notes": "CPU restatement (port) of the reference's per-group path (oracle/ols_oracle.c, OpenMP over groups); the "
                     "Rust crate cannot be built here (no cargo).  One 10k-group batch per step whatever --gpus is: the "
                     "metric is a rate (regressions/s), the GPU arm at N > 1 runs N such batches per step.",
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{len(times)} x full batch of {G} groups ({G * N_PER} rows), OpenMP over groups"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (CUDA)
    import torch
    import polars_ols_b200 as pls
    from polars_ols_b200 import _lib as L
    from polars_ols_b200.parallel import PeerGather, gather_group_results

    torch.cuda.set_device(local_rank)
    dist = None
    numa = None
    if world > 1:
        import torch.distributed as dist
        numa = bind_to_gpu_numa_node(local_rank)
        # the staging ring's copy threads share the host with the other ranks
        os.environ.setdefault("B200OLS_STAGE_THREADS", str(max(2, min(8, (os.cpu_count() or 16) // (2 * world)))))
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    x, y, offsets = make_data(rank)  # weak scaling: every rank owns a full, different 10k-group batch
    kw = pls.OLSKwargs(alpha=ALPHA, l1_ratio=0.0).to_c()

    # ---- value: inputs resident in HBM -----------------------------------------------------------------
    xd = torch.as_tensor(x, device=dev)          # [K, N] slab: each feature column contiguous (SoA)
    yd = torch.as_tensor(y, device=dev)
    # two output buffers: with N > 1 the all-gather of step i (NCCL, side stream) overlaps the kernel of step i+1
    coefs = [torch.empty((G, K), dtype=torch.float64, device=dev) for _ in range(2)]
    coef = coefs[0]
    stream = torch.cuda.current_stream(dev).cuda_stream or 1
    eng = pls.Engine(local_rank, stream)
    if a.tile_rows or a.warps or a.ctas_per_sm:
        eng.set_tuning(a.tile_rows, a.warps, a.ctas_per_sm)
    batch = pls.Batch(pls.Col(yd), [pls.Col(xd[i]) for i in range(K)], offsets=offsets)
    peer = None
    if world > 1 and a.gather == "peer":
        # fused gather: beta goes from the solver warps into every rank's [world*G, K] buffer over NVLink
        try:
            peer = PeerGather(eng, world * G, K, rank * G, n_buffers=2)
            peer.attach(0)
        except Exception as exc:  # no P2P / IPC on this box: fall back to the NCCL all-gather (same results)
            peer = None
            a.gather = "nccl"
            print(f"[bench] peer gather unavailable ({exc}); using NCCL", file=sys.stderr)
    # in peer mode the only outputs are the gathered buffers (no local [G, K] copy)
    steps_fn = [eng.prepare_least_squares(batch, kw, L.COEFFICIENTS, None if peer is not None else c_) for c_ in coefs]
    shards = [(r * G, (r + 1) * G) for r in range(world)]
    comm = torch.cuda.Stream(device=dev) if (world > 1 and peer is None) else None
    ev_done = [torch.cuda.Event() for _ in range(2)]     # kernel i finished writing coefs[i & 1]
    ev_gath = [torch.cuda.Event() for _ in range(2)]     # gather of coefs[i & 1] finished reading it
    gathered = [torch.empty((world * G, K), dtype=torch.float64, device=dev) if world > 1 else None]

    def full_step(i):
        b = i & 1
        if world == 1:
            steps_fn[b]()
            return
        if peer is not None:
            # double-buffered fused gather: step i stores into buffer i & 1 of every rank and signals; the wait for step
            # i - 1 (all ranks' rows of that step have landed here) is enqueued behind this step's kernel
            peer.attach(b)
            peer.arm_step()      # the kernel's last solver warp release-signals step i and acquire-waits step i - 1
            steps_fn[b]()
            return
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ev_gath[b])          # WAR: the gather issued two steps ago has consumed this buffer
        steps_fn[b]()
        ev_done[b].record(cur)
        comm.wait_event(ev_done[b])
        with torch.cuda.stream(comm):
            gather_group_results(coefs[b], shards, out=gathered[0])  # ONE NCCL all-gather of the coefficient chunks
            ev_gath[b].record(comm)

    def drain():
        if comm is not None:
            torch.cuda.current_stream(dev).wait_stream(comm)
        if peer is not None and peer.step > 0:
            peer.step_wait(peer.step)           # the last step's gather is complete too

    for i in range(a.warmup):
        full_step(i)
    drain()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = eng.launch_count
    eng.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        if peer is not None:
            # device-side rendezvous right before the start event: every rank signals and waits for all (flags over
            # NVLink), so the timed regions of all ranks begin at the same instant instead of carrying the host-side
            # skew of N processes leaving the barrier (20 steps x 0.13 ms are 2.6 ms in all)
            peer.step_complete()
        e0.record()
        for i in range(a.steps):
            if a.profile_every > 1:
                eng.set_profiling(i % a.profile_every == 0)
            full_step(i)
        drain()                              # every step's gather has completed inside the timed region
        e1.record()
        torch.cuda.synchronize()
    if dist:
        dist.barrier()
    coef = coefs[(a.steps - 1) & 1]
    ms = e0.elapsed_time(e1)
    kern_ms = eng.profile_drain()
    eng.set_profiling(False)
    launches = eng.launch_count - launches0 + (a.steps if (world > 1 and peer is None) else 0)
    if peer is not None:
        # every rank's shard must have landed in this rank's buffer (all ranks synchronised at the barrier above)
        assert not eng.peer_timed_out(), "a peer never signalled step completion"
        full = peer.read((a.steps - 1) & 1)
        assert np.isfinite(full).all() and (np.abs(full).sum(axis=1) > 0).all(), "peer gather incomplete"
        coef = torch.as_tensor(full[rank * G:(rank + 1) * G], device=dev)
        peer.close()
    t_local = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    ms_max = float(t_local.item())
    value = world * G * a.steps / (ms_max * 1e-3)

    # correctness guard on the timed output (normal equations hold), cheap and outside the timed region
    xg = xd.T.reshape(G, N_PER, K)[:64]
    r = yd.reshape(G, N_PER)[:64] - (xg * coef[:64, None, :]).sum(-1)
    grad = torch.einsum("gnk,gn->gk", xg, r) - ALPHA * coef[:64]
    assert float(grad.abs().max()) < 1e-6, "timed kernel output violates the normal equations"

    # ---- e2e: host buffers through the public C ABI, H2D + D2H inside the timed region -------------------
    e2e_steps = a.e2e_steps or min(a.steps, 10)
    heng = pls.Engine(local_rank)
    hx = heng.pinned_empty((K, G * N_PER))
    hy = heng.pinned_empty((G * N_PER,))
    hx[:] = x
    hy[:] = y
    hcoef = heng.pinned_empty((G, K))
    hbatch = pls.Batch(pls.Col(hy), [pls.Col(hx[i]) for i in range(K)], offsets=offsets)
    hstep = heng.prepare_least_squares(hbatch, kw, L.COEFFICIENTS, hcoef)
    for _ in range(3):
        hstep()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        hstep()   # returns when the coefficients are in host memory
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * G * e2e_steps / float(te.item())
    assert np.allclose(hcoef, coef.cpu().numpy(), rtol=1e-9, atol=1e-12)

    # ---- e2e_api: the drop-in call itself — Frame.select(col("y").least_squares.ridge(...).over("group")) on PAGEABLE
    # numpy columns plus an int64 key column: key upload + device group plan (b200ols_group_plan_build) + column upload
    # through the engine's pinned staging ring + kernel + coefficients back, all inside the timed region
    names = [f"x{i}" for i in range(K)]
    key = np.repeat(np.arange(G, dtype=np.int64), N_PER)
    fr = pls.Frame({"y": y, "group": key, **{nm: x[i] for i, nm in enumerate(names)}})
    expr = pls.col("y").least_squares.ridge(*names, alpha=ALPHA, mode="coefficients").over("group")

    def api_leg(frame, steps):
        for _ in range(2):
            frame.select(expr, engine=heng)
        if dist:
            dist.barrier()
        t0_ = time.perf_counter()
        for _ in range(steps):
            res = frame.select(expr, engine=heng)["coefficients"]
        dt = time.perf_counter() - t0_
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return world * G * steps / float(tt.item()), 1e3 * float(tt.item()) / steps, res

    api_value, api_ms, res = api_leg(fr, e2e_steps)
    assert np.allclose(res.to_numpy(), hcoef, rtol=1e-9, atol=1e-12)
    plan_ms = heng.group_plan([key]).device_ms
    # the same frame with its rows shuffled (tests/test_ols.py:39-40 shape: random, non-contiguous groups): the plan
    # now sorts 10M keys and the columns are gathered on the device
    e2e_api_shuffled = None
    if rank == 0 and world == 1:
        perm = np.random.default_rng(1).permutation(G * N_PER)
        frs = pls.Frame({"y": y[perm], "group": key[perm], **{nm: x[i][perm] for i, nm in enumerate(names)}})
        sv, sms, sres = api_leg(frs, max(3, e2e_steps // 2))
        assert np.allclose(sres.to_numpy(), hcoef, rtol=1e-9, atol=1e-12)
        e2e_api_shuffled = {"value": sv, "unit": UNIT, "ms_per_step": sms,
                            "plan_device_ms": heng.group_plan([key[perm]]).device_ms}
        del frs, perm

    # ---- strong scaling (N > 1): ONE 10k-group frame sharded by group over the N GPUs (N PCIe links for the e2e leg) --
    strong = None
    if world > 1:
        from polars_ols_b200.parallel import shard_groups
        xs, ys, offs = (x, y, offsets) if rank == 0 else make_data(0)      # the same frame on every rank
        g0, g1 = shard_groups(offs, world)[rank]
        r0, r1 = int(offs[g0]), int(offs[g1])
        loffs = np.ascontiguousarray(offs[g0:g1 + 1] - r0)
        seng = pls.Engine(local_rank, stream)
        sxd = torch.as_tensor(np.ascontiguousarray(xs[:, r0:r1]), device=dev)
        syd = torch.as_tensor(np.ascontiguousarray(ys[r0:r1]), device=dev)
        sbatch = pls.Batch(pls.Col(syd), [pls.Col(sxd[i]) for i in range(K)], offsets=loffs)
        speer = PeerGather(seng, G, K, g0, n_buffers=2) if peer is not None else None
        scoef = torch.empty((g1 - g0, K), dtype=torch.float64, device=dev)
        if speer is not None:
            speer.attach(0)
        sstep = seng.prepare_least_squares(sbatch, kw, L.COEFFICIENTS, None if speer is not None else scoef)
        sgath = torch.empty((G, K), dtype=torch.float64, device=dev)
        sshards = shard_groups(offs, world)

        def strong_step():
            if speer is not None:
                speer.attach(speer.step & 1)
                speer.arm_step()
                sstep()
            else:
                sstep()
                gather_group_results(scoef, sshards, out=sgath if all(b_ - a_ == sshards[0][1] - sshards[0][0] for a_, b_ in sshards) else None)

        for _ in range(a.warmup):
            strong_step()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if speer is not None:
            speer.step_complete()               # device-side rendezvous, as in the weak-scaling loop
        s0.record()
        for _ in range(a.steps):
            strong_step()
        if speer is not None:
            speer.step_wait(speer.step)
        s1.record()
        torch.cuda.synchronize(); dist.barrier()
        ts = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        if speer is not None:
            assert not seng.peer_timed_out()
            full = speer.read((speer.step - 1) & 1)
            speer.close()
            if rank == 0:   # rank 0 holds the identical frame in full: every shard must have landed, and be right
                assert np.allclose(full, hcoef, rtol=1e-9, atol=1e-12), "strong-scaling gather differs from the single-GPU result"
        # e2e: every rank uploads ITS shard from pinned host memory and reads its coefficient rows back
        shx = seng.pinned_empty((K, r1 - r0)); shy = seng.pinned_empty((r1 - r0,)); shc = seng.pinned_empty((g1 - g0, K))
        shx[:] = xs[:, r0:r1]; shy[:] = ys[r0:r1]
        heng2 = pls.Engine(local_rank)
        hb2 = pls.Batch(pls.Col(shy), [pls.Col(shx[i]) for i in range(K)], offsets=loffs)
        hs2 = heng2.prepare_least_squares(hb2, kw, L.COEFFICIENTS, shc)
        for _ in range(3):
            hs2()
        dist.barrier()
        t0_ = time.perf_counter()
        for _ in range(e2e_steps):
            hs2()
        tse = torch.tensor([time.perf_counter() - t0_], dtype=torch.float64, device=dev)
        dist.all_reduce(tse, op=dist.ReduceOp.MAX)
        strong = {"scaling": "strong", "frame": "ONE 10k-group frame, groups sharded over the ranks by cumulative rows",
                  "value": G * a.steps / (float(ts.item()) * 1e-3), "ms_per_step": float(ts.item()) / a.steps,
                  "e2e": {"value": G * e2e_steps / float(tse.item()), "ms_per_step": 1e3 * float(tse.item()) / e2e_steps,
                          "h2d_bytes_per_step_per_rank": int(shx.nbytes + shy.nbytes)}, "unit": UNIT}
        del sxd, syd

    # ---- roofline of the dominant kernel (row-streaming Gram + fused solve) ---------------------------------
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    alg_bytes = G * ALG_BYTES_PER_REGRESSION
    k_avg_ms = float(np.mean(kern_ms)) if len(kern_ms) else float("nan")
    achieved = alg_bytes / (k_avg_ms * 1e-3) / 1e9
    traffic = None
    tp = ROOT / "profiles" / "gram_traffic.json"
    if tp.exists():
        traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "gram_cta_kernel<double,1> (TMA bulk-copy pipeline + DMMA Gram + fused warp Cholesky solve)",
                "kernel_ms_avg": k_avg_ms, "kernel_launches_timed": int(len(kern_ms)), "timed_every_nth_step": int(max(1, a.profile_every)),
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "kernel_share_of_step": k_avg_ms * a.steps / ms if ms > 0 else None}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on the host cores, bounded sample ---------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        times, threads, cpu_out = cpu_port_run(x, y, offsets, 3, 1, min_seconds=10.0)
        cpu = {"value": G * len(times) / sum(times), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{len(times)} x full batch of {G} groups (~{sum(times):.0f} s of CPU work), OpenMP over groups"}
        err = float(np.max(np.abs(cpu_out - hcoef) / (1e-3 + np.abs(cpu_out))))
        cpu["max_rel_err_vs_gpu"] = err

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": make_config(world),
            "notes": ("" if world == 1 else
                      ("coefficient chunks gathered by P2P stores from the solver warps into every rank's buffer (NVLink peer memory, "
                       "fused into the kernel, two alternating buffers); every step signals a release flag to all ranks and the acquire-wait "
                       "for step i - 1 are done by the Gram kernel's last solver warp itself, INSIDE the timed loop (b200ols_peer_arm_step): the "
                       "timed region ends only when every rank's rows of every step have landed everywhere"
                       if a.gather == "peer" else
                       "NCCL all-gather of coefficient chunks per step (side stream, overlapped with the next step's kernel)")),
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(hx.nbytes + hy.nbytes + offsets.nbytes),
                    "d2h_bytes_per_step": int(hcoef.nbytes), "steps": e2e_steps, "ms_per_step": 1e3 * float(te.item()) / e2e_steps},
            "e2e_api": {"value": api_value, "unit": UNIT, "ms_per_step": api_ms, "steps": e2e_steps,
                        "call": "Frame.select(col('y').least_squares.ridge(*x, alpha=1e-3, mode='coefficients').over('group'))",
                        "inputs": "pageable numpy f64 columns + int64 key column; group plan on the device",
                        "h2d_bytes_per_step": int(x.nbytes + y.nbytes + key.nbytes), "d2h_bytes_per_step": int(hcoef.nbytes + 2 * 8 * (G + 1)),
                        "plan_device_ms": plan_ms, "shuffled_rows": e2e_api_shuffled},
            "e2e_roofline": {"bound": "pcie", "achieved": (hx.nbytes + hy.nbytes) / (1e6 * 1e3 * float(te.item()) / e2e_steps),
                             "peak": 63.0, "unit": "GB/s", "peak_source": "PCIe Gen5 x16, 32 GT/s x 16 lanes x 128b/130b",
                             "frac": (hx.nbytes + hy.nbytes) / (1e6 * 1e3 * float(te.item()) / e2e_steps) / 63.0},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
        if strong is not None:
            line["strong"] = strong
        if numa is not None:
            line["numa"] = numa
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
