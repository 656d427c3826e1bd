"""Tuning sweep of the row-streaming Gram kernel on the C2 workload (run on the GPU box):
python tools/sweep_gram.py [--out gpurun_out/sweep.json]"""
import argparse, json, sys, itertools
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import polars_ols_b200 as pls
from polars_ols_b200 import _lib as L
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/sweep.json")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--variants", default="0,1,2,3")
a = ap.parse_args()
x, y, offsets = bench.make_data(0)
dev = torch.device("cuda", 0)
xd, yd = torch.as_tensor(x, device=dev), torch.as_tensor(y, device=dev)
coef = torch.empty((bench.G, bench.K), dtype=torch.float64, device=dev)
kw = pls.OLSKwargs(alpha=bench.ALPHA, l1_ratio=0.0).to_c()
eng = pls.Engine(0, 1)
batch = pls.Batch(pls.Col(yd), [pls.Col(xd[i]) for i in range(bench.K)], offsets=offsets)
step = eng.prepare_least_squares(batch, kw, L.COEFFICIENTS, coef)
res = []
alg = bench.G * bench.ALG_BYTES_PER_REGRESSION
configs = [(3, 0, t, w, 1) for t, w in itertools.product([256, 336, 504, 512, 1000, 1024], [0, 2, 3, 4])]
configs += [(0, 0, t, w, c) for t, w, c in itertools.product([64, 128, 256], [4, 6, 8, 12], [1, 2])]
configs += [(1, u, 0, w, c) for u, w, c in itertools.product([1, 2], [8], [2, 3, 4])]
configs += [(2, u, 0, w, c) for u, w, c in itertools.product([1, 2, 4], [4, 8, 16], [1, 2, 3, 4]) if not (u > 1 and w > 8) and w * c <= 32]
VARS = [int(v) for v in a.variants.split(',')]
configs = [c for c in configs if c[0] in VARS]
for variant, unroll, tile, warps, cps in configs:
    try:
        eng.set_variant(variant, unroll)
        eng.set_tuning(tile, warps, cps)
        for _ in range(3):
            step()
        eng.set_profiling(True)
        for _ in range(a.steps):
            step()
        ms = eng.profile_drain()
        eng.set_profiling(False)
        r = {"variant": variant, "unroll": unroll, "tile_rows": tile, "warps": warps, "ctas_per_sm": cps, "ms": float(np.median(ms)), "gbs": alg / (float(np.median(ms)) * 1e-3) / 1e9}
    except Exception as e:  # config does not fit
        r = {"variant": variant, "unroll": unroll, "tile_rows": tile, "warps": warps, "ctas_per_sm": cps, "error": str(e)[:120]}
    res.append(r)
    print(r, flush=True)
Path(a.out).parent.mkdir(exist_ok=True)
Path(a.out).write_text(json.dumps(res, indent=1))
best = min((r for r in res if "ms" in r), key=lambda r: r["ms"])
print("BEST", best)
for v in VARS:
    print("BEST variant", v, min((r for r in res if "ms" in r and r["variant"] == v), key=lambda r: r["ms"]))
