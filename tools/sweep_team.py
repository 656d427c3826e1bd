"""Sweep over group lengths of (a) fused vs separate (batch_solve) normal-equation solve and (b) the gram_cta
team size (consumer warps per segment); device-resident inputs, ridge coefficients.  Prints ms per call (CUDA
events over 10 calls) and the streaming kernel's own time (profile_drain).  GPU box only."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import polars_ols_b200 as pls  # noqa: E402
from polars_ols_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda", 0)
out = {}
cases = [("f32", 16, 256, 100_000, True), ("f64", 8, 64, 160_000, False), ("f64", 8, 256, 40_000, False),
         ("f64", 8, 512, 20_000, False), ("f64", 8, 1000, 10_000, False), ("f64", 16, 1000, 10_000, False),
         ("f64", 16, 4000, 2_500, False)]
if len(sys.argv) > 1:
    cases = [cases[int(i)] for i in sys.argv[1].split(",")]
engines = {}
for name, fuse, multi in (("fused", "0", "1"), ("batch", str(1 << 40), "0"), ("multi", str(1 << 40), "1")):
    os.environ["B200OLS_FUSE_MIN_BYTES"] = fuse
    os.environ["B200OLS_MULTI"] = multi
    engines[name] = pls.Engine(0, 1)
del os.environ["B200OLS_FUSE_MIN_BYTES"], os.environ["B200OLS_MULTI"]
engines["default"] = pls.Engine(0, 1)
for dt, k, n, G, weighted in cases:
    tdt = torch.float32 if dt == "f32" else torch.float64
    g = torch.Generator(device=dev).manual_seed(1)
    N = n * G
    x = torch.randn(k, N, dtype=tdt, device=dev, generator=g)
    y = x.sum(0) + 0.1 * torch.randn(N, dtype=tdt, device=dev, generator=g)
    w = torch.rand(N, dtype=tdt, device=dev, generator=g) + 0.1 if weighted else None
    b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)], None if w is None else pls.Col(w),
                  offsets=np.arange(G + 1, dtype=np.int64) * n)
    kw = pls.OLSKwargs(alpha=1e-3, l1_ratio=0.0).to_c()
    esz = 4 if dt == "f32" else 8
    gb = (N * (k + 1 + (1 if weighted else 0)) * esz + G * k * 8) / 1e9
    ref = None
    for fuse, eng in engines.items():
        coef = torch.empty((G, k), dtype=torch.float64, device=dev)
        call = eng.prepare_least_squares(b, kw, L.COEFFICIENTS, coef)
        for team in ((0,) if fuse in ("default", "multi") else (0, 2)):
            eng.set_tuning(0, 0, team)
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            eng.set_profiling(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                call()
            e1.record()
            torch.cuda.synchronize()
            ms_call = e0.elapsed_time(e1) / 10
            ms = float(np.median(eng.profile_drain()))
            eng.set_profiling(False)
            cc = coef.cpu().numpy().copy()
            if ref is None:
                ref = cc
            err = float(np.abs(cc - ref).max())
            key = f"{dt} k={k} n={n} G={G} solve={fuse} team={team}"
            out[key] = {"ms_per_call": round(ms_call, 4), "gram_ms": round(ms, 4), "GBps_call": round(gb / ms_call * 1e3, 1),
                        "max_abs_diff": err}
            print(key, out[key], flush=True)
        del call, coef
    del x, y, w, b
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/sweep_team.json").write_text(json.dumps(out, indent=1))
