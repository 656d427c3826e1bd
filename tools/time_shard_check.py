"""Multi-GPU check of the time-axis sharding of ONE long series: run under torchrun with N
ranks; every rank evaluates its shard of an rls and a rolling_ols expression, rank 0 compares the gathered
result with the single-GPU evaluation of the whole series and prints device timings.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/time_shard_check.py --rows 20000000
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polars_ols_b200 import Frame, col  # noqa: E402
from polars_ols_b200.parallel import gather_rows, shard_rows, time_sharded  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=20_000_000)
    ap.add_argument("--features", type=int, default=6)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator(device=dev).manual_seed(0)           # same series on every rank
    n, k = a.rows, a.features
    cols = {f"x{j}": torch.randn(n, generator=g, device=dev, dtype=torch.float64) for j in range(k)}
    y = sum(cols.values()) + 0.1 * torch.randn(n, generator=g, device=dev, dtype=torch.float64)
    fr = Frame({**cols, "y": y})
    names = list(cols)
    report = {"rows": n, "features": k, "world": world}
    for label, expr in (("rolling", col("y").least_squares.rolling_ols(*names, window_size=252, min_periods=k, mode="coefficients")),
                        ("rls", col("y").least_squares.rls(*names, half_life=252.0, mode="coefficients"))):
        exch = (lambda o: [o]) if world == 1 else None
        ms = []
        for _ in range(a.reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res, (r0, r1) = time_sharded(expr, fr, rank, world, exchange=exch)
            torch.cuda.synchronize()
            t = torch.tensor([time.perf_counter() - t0], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms.append(float(t) * 1e3)
        report[label + "_ms_sharded"] = round(min(ms), 3)
        full = gather_rows(res.values, shard_rows(n, world)) if world > 1 else res.values
        if rank == 0:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            whole = expr.evaluate(fr).values
            torch.cuda.synchronize()
            report[label + "_ms_one_gpu"] = round((time.perf_counter() - t0) * 1e3, 3)
            skip = 1024
            err = ((full[skip:] - whole[skip:]).abs() / (1e-3 + whole[skip:].abs())).max().item()
            report[label + "_max_rel_err_vs_one_gpu"] = err
            assert err < 1e-6, (label, err)
            del whole
        del full, res
    if rank == 0:
        print(json.dumps(report))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
