"""Where does the device RLS differ from the oracle?  Prints the normalised error (|d| / (1e-8 + 1e-6 |ref|)) by row range."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import polars_ols_b200 as pls
from polars_ols_b200 import Frame, col
from oracle import semantics as S

def report(tag, got, ref):
    got, ref = np.asarray(got, dtype=float), np.asarray(ref, dtype=float)
    if got.ndim == 1: got, ref = got[:, None], ref[:, None]
    e = np.nanmax(np.abs(got - ref) / (1e-8 + 1e-6 * np.abs(ref)), axis=1)
    nanmis = int((np.isnan(got) != np.isnan(ref)).sum())
    edges = [0, 1, 2, 4, 8, 16, 32, 64, 128, 256, 1024, 4096, len(e)]
    parts = []
    for a, b in zip(edges[:-1], edges[1:]):
        if a < len(e):
            parts.append(f"[{a},{min(b, len(e))}):{np.nanmax(e[a:b]):.2e}")
    print(tag, "nan-mismatch", nanmis, " ".join(parts), flush=True)

rng = np.random.default_rng(0)
for (n, k, p0, hl, mean) in [(6000, 2, 1e6, None, None), (6000, 2, 10.0, 252.0, None), (4000, 3, 10.0, 21.0, [-1.0, -1.0, -1.0]),
                            (4000, 8, 10.0, 500.0, None), (4000, 10, 10.0, 500.0, None), (300000, 6, 10.0, 252.0, None)]:
    x = rng.normal(size=(n, k)); y = x @ np.ones(k) + 0.1 * rng.normal(size=n)
    d = {"y": y, **{f"x{i}": np.ascontiguousarray(x[:, i]) for i in range(k)}}
    names = [f"x{i}" for i in range(k)]
    for mode in ("coefficients", "predictions"):
        kw = dict(half_life=hl, initial_state_covariance=p0, initial_state_mean=mean)
        r = Frame(d).select(col("y").least_squares.rls(*names, mode=mode, **kw))
        r = r["coefficients" if mode == "coefficients" else "y"].to_numpy()
        ref, _ = S.recursive_least_squares(y, *[d[nm] for nm in names], mode=mode, kwargs=S.RLSKwargs(**kw))
        report(f"n={n} k={k} p0={p0} hl={hl} mean={mean is not None} {mode}", r, ref)
