"""Sweep of the pageable-host staging ring (staging.cu) on the C2 frame: threads x slot size x slots x NT stores.
One subprocess per setting (the ring is configured from the environment when the context is created)."""
import itertools, json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
CHILD = r'''
import sys, time, json, numpy as np
sys.path.insert(0, %r)
import polars_ols_b200 as pls
from polars_ols_b200 import _lib as L
G, N, K = 10000, 1000, 8
rng = np.random.default_rng(0)
x = rng.standard_normal((K, G * N)); y = rng.standard_normal(G * N)
off = np.arange(G + 1, dtype=np.int64) * N
eng = pls.Engine(0)
out = np.empty((G, K))
b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(K)], offsets=off)
step = eng.prepare_least_squares(b, pls.OLSKwargs(alpha=1e-3, l1_ratio=0.0).to_c(), L.COEFFICIENTS, out)
for _ in range(3): step()
ts = []
for _ in range(6):
    t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
key = np.repeat(np.arange(G, dtype=np.int64), N)
tp = []
for _ in range(4):
    t0 = time.perf_counter(); eng.group_plan([key]); tp.append(time.perf_counter() - t0)
print(json.dumps({"ms_min": 1e3 * min(ts), "ms_med": 1e3 * sorted(ts)[len(ts) // 2], "plan_ms_min": 1e3 * min(tp)}))
''' % str(ROOT)
res = []
grid = [(t, kb, sl, nt) for t in (3, 4, 5, 6, 8) for kb in (512, 2048, 4096) for sl in (0,) for nt in (1, 0)]
grid += [(4, 1024, 8, 1), (4, 4096, 32, 1), (6, 1024, 12, 1), (6, 256, 24, 1)]
for t, kb, sl, nt in grid:
    env = dict(os.environ, B200OLS_STAGE_THREADS=str(t), B200OLS_STAGE_SLOT_KB=str(kb), B200OLS_STAGE_NT=str(nt))
    if sl:
        env["B200OLS_STAGE_SLOTS"] = str(sl)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        d = {"error": r.stderr[-300:]}
    d.update(threads=t, slot_kb=kb, slots=sl, nt=nt)
    res.append(d)
    print(d, flush=True)
json.dump(res, open(ROOT / "gpurun_out" / "stage_sweep.json", "w"), indent=1)
