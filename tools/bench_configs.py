"""Times every BASELINE.json config (C1..C5) through the public API with device-resident inputs and checks
a sample of each against the CPU oracle.  Run on the GPU box:  python tools/bench_configs.py [--only C3,C5]
Writes gpurun_out/configs.json.  (bench.py stays the contract benchmark for C2.)"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import polars_ols_b200 as pls  # noqa: E402
from polars_ols_b200 import _lib as L  # noqa: E402
from oracle import semantics as S  # noqa: E402  (checker)

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="C1,C2,F,C3,C4,C5")
ap.add_argument("--out", default="gpurun_out/configs.json")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--scale", type=float, default=1.0, help="shrink the big configs (debug)")
a = ap.parse_args()
only = set(a.only.split(","))
dev = torch.device("cuda", 0)
PEAK = 6569.3
pk = Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"
if pk.exists():
    PEAK = float(json.loads(pk.read_text())["hbm_gbs"])
eng = pls.Engine(0, 1)
results = {}


def timed(fn, reps):
    fn(); fn()
    torch.cuda.synchronize()
    eng.set_profiling(True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        out = fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    k = eng.profile_drain()
    eng.set_profiling(False)
    ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]))  # median: one allocator hiccup must not count
    return out, ms, (float(np.median(k)) if len(k) else None)


def report(name, ms, kms, alg_bytes, units, unit_name, err, extra=None):
    r = {"ms_per_call": ms, "dominant_kernel_ms": kms, "algorithmic_GB": alg_bytes / 1e9,
         "GBps_whole_call": alg_bytes / ms / 1e6, "frac_of_hbm_peak_whole_call": alg_bytes / ms / 1e6 / PEAK,
         f"{unit_name}_per_s": units / (ms * 1e-3), "max_rel_err_vs_oracle_sample": err}
    if extra:
        r.update(extra)
    results[name] = r
    print(name, json.dumps(r), flush=True)


def gen(n, k, G, dtype, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(k, n, dtype=dtype, device=dev, generator=g)
    beta = (1 + 0.25 * torch.randn(G, k, dtype=torch.float64, device=dev, generator=g))
    per = n // G
    y = torch.empty(n, dtype=torch.float64, device=dev)
    step = max(1, G // 16)
    for g0 in range(0, G, step):  # chunked to bound temporary memory
        g1 = min(G, g0 + step)
        xs = x[:, g0 * per:g1 * per].T.reshape(g1 - g0, per, k).to(torch.float64)
        y[g0 * per:g1 * per] = (xs * beta[g0:g1, None, :]).sum(-1).reshape(-1)
    y += 0.1 * torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    return x, y.to(dtype)


def rel_err(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    m = ~np.isnan(ref)
    assert (np.isnan(got) == np.isnan(ref)).all()
    return float(np.max(np.abs(got[m] - ref[m]) / (1e-3 + np.abs(ref[m])))) if m.any() else 0.0


# ------------------------------------------------------------------------------------------------ C1
if "C1" in only:
    n, k = 1000, 3
    x, y = gen(n, k, 1, torch.float64, 1)
    b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)])
    kw = pls.OLSKwargs().to_c()
    out, ms, kms = timed(lambda: eng.least_squares(b, kw, L.PREDICTIONS)[0], 50)
    ref = S.least_squares(y.cpu().numpy(), *[x[i].cpu().numpy() for i in range(k)])[0]
    report("C1 ols predictions 1x1000x3 f64", ms, kms, 40_000, 1, "regressions", rel_err(out.cpu().numpy(), ref),
           {"latency_us": ms * 1e3})

# ------------------------------------------------------------------------------------------------ C2
if "C2" in only:
    G, per, k = 10_000, 1000, 8
    x, y = gen(G * per, k, G, torch.float64, 2)
    offs = np.arange(G + 1, dtype=np.int64) * per
    b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)], offsets=offs)
    kw = pls.OLSKwargs(alpha=1e-3, l1_ratio=0.0).to_c()
    out, ms, kms = timed(lambda: eng.least_squares(b, kw, L.COEFFICIENTS)[0], a.reps)
    sel = np.arange(0, G, G // 32)
    ref = np.stack([S.solve_ridge(y[g * per:(g + 1) * per].cpu().numpy(), np.ascontiguousarray(x[:, g * per:(g + 1) * per].T.cpu().numpy()), 1e-3, None, None) for g in sel])
    report("C2 ridge coefficients 10kx1000x8 f64", ms, kms, G * (per * 9 * 8 + 64), G, "regressions", rel_err(out[sel].cpu().numpy(), ref))
    for mode, nm in ((L.PREDICTIONS, "predictions"), (L.RESIDUALS, "residuals")):
        out, ms, kms = timed(lambda: eng.least_squares(b, kw, mode)[0], a.reps)
        report(f"C2 ridge {nm} 10kx1000x8 f64", ms, kms, G * (per * 9 * 8 + per * 8), G, "regressions", None)
    del x, y

# ------------------------------------------------------------------------------------------------ §8f rows on the C2 shape
if "F" in only:
    G, per, k, m = 10_000, 1000, 8, 4
    x, y = gen(G * per, k, G, torch.float64, 2)
    offs = np.arange(G + 1, dtype=np.int64) * per
    b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)], offsets=offs)
    kw = pls.OLSKwargs(alpha=1e-3, l1_ratio=0.0).to_c()
    out, ms, kms = timed(lambda: eng.least_squares_statistics(b, kw), a.reps)
    g0 = slice(0, per)
    ref = S.least_squares_statistics(y[g0].cpu().numpy(), *[x[i, g0].cpu().numpy() for i in range(k)], kwargs=S.OLSKwargs(alpha=1e-3, l1_ratio=0.0))
    err = max(rel_err(out[nm][0].cpu().numpy(), np.asarray(ref[nm])) for nm in ("r2", "mae", "mse", "coefficients", "standard_errors", "t_values"))
    report("F4 ridge statistics 10kx1000x8 f64", ms, kms, G * (per * 9 * 8 + (3 + 4 * k) * 8), G, "regressions", err)
    g = torch.Generator(device=dev).manual_seed(7)
    ys = [y] + [x[j] - 0.5 * x[j + 1] + 0.1 * torch.randn(G * per, dtype=torch.float64, device=dev, generator=g) for j in range(m - 1)]
    kws = pls.OLSKwargs(alpha=1e-3, solve_method="svd").to_c()
    out, ms, kms = timed(lambda: eng.multi_target_least_squares(b, [pls.Col(t) for t in ys], kws, L.PREDICTIONS)[0], a.reps)
    ref = S.multi_target_least_squares([t[g0].cpu().numpy() for t in ys], *[x[i, g0].cpu().numpy() for i in range(k)],
                                       kwargs=S.OLSKwargs(alpha=1e-3, solve_method="svd"))[0]
    report(f"F2 multi-target ridge predictions 10kx1000x8, {m} targets f64", ms, kms, G * per * ((k + m) * 8 + m * 8), G * m,
           "regressions", rel_err(out[:, g0].cpu().numpy().T, ref))
    del x, y, ys

# ------------------------------------------------------------------------------------------------ C3
if "C3" in only:
    G, per, k = int(100_000 * a.scale), 256, 16
    x, y = gen(G * per, k, G, torch.float32, 3)
    w = torch.rand(G * per, dtype=torch.float32, device=dev) + 0.05
    offs = np.arange(G + 1, dtype=np.int64) * per
    b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)], weights=pls.Col(w), offsets=offs)
    kw = pls.OLSKwargs(alpha=1e-3, l1_ratio=0.5).to_c()
    out, ms, kms = timed(lambda: eng.least_squares(b, kw, L.PREDICTIONS)[0], a.reps)
    sel = np.arange(0, G, G // 16)
    errs = []
    for g in sel:
        sl = slice(g * per, (g + 1) * per)
        ref = S.least_squares(y[sl].cpu().numpy(), *[x[i, sl].cpu().numpy() for i in range(k)], sample_weights=w[sl].cpu().numpy(),
                              kwargs=S.OLSKwargs(alpha=1e-3, l1_ratio=0.5))[0]
        errs.append(rel_err(out[sl].cpu().numpy(), ref))
    report("C3 wls+elastic_net predictions 100kx256x16 f32", ms, kms, G * (per * 18 * 4 + per * 8), G, "regressions", max(errs))
    out, ms, kms = timed(lambda: eng.least_squares(b, kw, L.COEFFICIENTS)[0], a.reps)
    report("C3 wls+elastic_net coefficients 100kx256x16 f32", ms, kms, G * (per * 18 * 4 + 16 * 8), G, "regressions", None)
    del x, y, w

# ------------------------------------------------------------------------------------------------ C4
if "C4" in only:
    n, k = int(50_000_000 * a.scale), 6
    x, y = gen(n, k, 1, torch.float64, 4)
    b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)])
    lo, hi = n // 2, n // 2 + 20_000
    xs = [x[i, lo - 30_000:hi].cpu().numpy() for i in range(k)]
    ys = y[lo - 30_000:hi].cpu().numpy()
    for mode, nm, outb in ((L.PREDICTIONS, "predictions", 8), (L.COEFFICIENTS, "coefficients", 48)):
        kwr = pls.RollingKwargs(window_size=252, min_periods=6, null_policy="drop").to_c()
        out, ms, kms = timed(lambda: eng.rolling_least_squares(b, kwr, mode)[0], 2)
        ref = S.rolling_least_squares(ys, *xs, mode=nm, kwargs=S.RollingKwargs(window_size=252, min_periods=6, null_policy="drop"))[0]
        report(f"C4 rolling_ols(252) {nm} 1x50Mx6 f64", ms, kms, n * (56 + outb), n, "rows", rel_err(out[lo:hi].cpu().numpy(), ref[30_000:]))
        del out
        mean = None
        kwl = L.RLSKwargs(252.0, 10.0, None, L.NULL_POLICY["drop"], 0)
        out, ms, kms = timed(lambda: eng.recursive_least_squares(b, kwl, mode)[0], 2)
        ref = S.recursive_least_squares(ys, *xs, mode=nm, kwargs=S.RLSKwargs(half_life=252.0))[0]
        report(f"C4 rls(half_life=252) {nm} 1x50Mx6 f64", ms, kms, n * (56 + outb), n, "rows", rel_err(out[lo:hi].cpu().numpy(), ref[30_000:]))
        del out
    del x, y

# ------------------------------------------------------------------------------------------------ C5
if "C5" in only:
    G, per, k = int(1000 * a.scale), 10_000, 64
    x, y = gen(G * per, k, G, torch.float64, 5)
    offs = np.arange(G + 1, dtype=np.int64) * per
    b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)], offsets=offs)
    kw = pls.OLSKwargs(alpha=1e-4, l1_ratio=1.0).to_c()
    out, ms, kms = timed(lambda: eng.least_squares(b, kw, L.COEFFICIENTS)[0], a.reps)
    sel = [0, G // 2, G - 1]
    t0 = time.time()
    ref = np.stack([S.solve_elastic_net(y[g * per:(g + 1) * per].cpu().numpy(), np.ascontiguousarray(x[:, g * per:(g + 1) * per].T.cpu().numpy()),
                                        1e-4, 1.0, 1000, 1e-5, False, None) for g in sel])
    report("C5 lasso coefficients 1000x10kx64 f64", ms, kms, G * (per * 65 * 8 + 512), G, "regressions", rel_err(out[sel].cpu().numpy(), ref),
           {"oracle_s_per_group": (time.time() - t0) / len(sel)})

Path(a.out).parent.mkdir(exist_ok=True)
Path(a.out).write_text(json.dumps(results, indent=1))
