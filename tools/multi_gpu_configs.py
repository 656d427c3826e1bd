"""BASELINE.json's multi-GPU configs under torchrun: C3 (wls + elastic_net predictions, 100k groups x 256 x 16 f32, groups
sharded over the ranks) and C5 (lasso coefficients, 1000 groups x 10k x 64 f64).  STRONG scaling: the frame is fixed, rank
r owns a contiguous range of groups (parallel.shard_groups), no data-path collective; the only exchange is the final
NCCL all-gather of the per-rank output chunks, timed separately.  Device timing, max over ranks.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/multi_gpu_configs.py --out profiles/r02_c3_c5_n8.json
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polars_ols_b200 as pls  # noqa: E402
from polars_ols_b200 import _lib as L  # noqa: E402
from polars_ols_b200.parallel import gather_group_results, gather_rows, shard_groups  # noqa: E402


def gen(n, k, G, dtype, seed, dev):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(k, n, dtype=dtype, device=dev, generator=g)
    beta = 1 + 0.25 * torch.randn(G, k, dtype=torch.float64, device=dev, generator=g)
    per = n // G
    y = torch.empty(n, dtype=torch.float64, device=dev)
    step = max(1, G // 16)
    for g0 in range(0, G, step):
        g1 = min(G, g0 + step)
        xs = x[:, g0 * per:g1 * per].T.reshape(g1 - g0, per, k).to(torch.float64)
        y[g0 * per:g1 * per] = (xs * beta[g0:g1, None, :]).sum(-1).reshape(-1)
    y += 0.1 * torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    return x, y.to(dtype)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = pls.Engine(lr, torch.cuda.current_stream(dev).cuda_stream or 1)
    report = {"world": world}

    def run(name, G, per, k, dtype, kw, mode, weights, alg_bytes_per_group):
        offs = np.arange(G + 1, dtype=np.int64) * per
        g0, g1 = shard_groups(offs, world)[rank]
        Gl = g1 - g0
        x, y = gen(Gl * per, k, Gl, dtype, 100 + rank, dev)          # this rank's groups only
        w = (torch.rand(Gl * per, dtype=dtype, device=dev) + 0.05) if weights else None
        b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)], weights=pls.Col(w) if weights else None,
                      offsets=np.arange(Gl + 1, dtype=np.int64) * per)
        shards = shard_groups(offs, world)

        def step():
            return eng.least_squares(b, kw, mode, want_validity=False)[0]

        for _ in range(3):
            out = step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        for _ in range(a.reps):
            out = step()
        ev[1].record()
        if world > 1:   # the path's only exchange: final gather of the output chunks
            for _ in range(a.reps):
                full = gather_group_results(out, shards) if mode == L.COEFFICIENTS else gather_rows(out, [(s0 * per, s1 * per) for s0, s1 in shards])
        ev[2].record()
        torch.cuda.synchronize()
        t = torch.tensor([ev[0].elapsed_time(ev[1]) / a.reps, ev[1].elapsed_time(ev[2]) / a.reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, gms = float(t[0]), float(t[1])
        report[name] = {"groups": G, "groups_per_rank": Gl, "ms_per_call": ms, "regressions_per_s": G / (ms * 1e-3),
                        "GBps_whole_job": G * alg_bytes_per_group / ms / 1e6, "final_gather_ms": gms if world > 1 else None,
                        "regressions_per_s_incl_gather": G / ((ms + gms) * 1e-3)}
        del x, y, w, out

    run("C3 wls+elastic_net predictions 100kx256x16 f32", 100_000, 256, 16, torch.float32,
        pls.OLSKwargs(alpha=1e-3, l1_ratio=0.5).to_c(), L.PREDICTIONS, True, 256 * 18 * 4 + 256 * 8)
    run("C5 lasso coefficients 1000x10kx64 f64", 1000, 10_000, 64, torch.float64,
        pls.OLSKwargs(alpha=1e-4, l1_ratio=1.0).to_c(), L.COEFFICIENTS, False, 10_000 * 65 * 8 + 512)
    if rank == 0:
        print(json.dumps(report))
        if a.out:
            with open(a.out, "w") as f:
                json.dump(report, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
