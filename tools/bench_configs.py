"""Every BASELINE.json config (C1..C5) through the public C ABI: device-resident `value`, host -> device -> host `e2e`,
roofline of the dominant kernel(s) and the CPU baseline (the oracle port running the reference's algorithm for that
config on the host cores, bounded sample) — the same line shape bench.py prints for C2.

    python tools/bench_configs.py [--only C1,C3,C4,C5] [--out profiles/r02_configs.json]
    python bench.py --config C4          # one config, one JSON line (delegates to run_config below)

The oracle (`oracle/`) is only used as the checker of a sample and as the timed CPU baseline."""
import argparse
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

UNIT = {"C1": "regressions/s", "C2": "regressions/s", "C3": "regressions/s", "C4": "rows/s", "C4rls": "rows/s", "C5": "regressions/s"}
WORKLOAD = {
    "C1": "C1: ols predictions, 1 group x 1000 rows x 3 features, f64 (latency case)",
    "C2": "C2: ridge(alpha=1e-3) coefficients .over(group), 10000 groups x 1000 rows x 8 features, f64",
    "C3": "C3: wls + elastic_net(alpha=1e-3, l1_ratio=0.5) predictions .over(group), 100000 groups x 256 rows x 16 features, f32",
    "C4": "C4: rolling_ols(window_size=252, min_periods=6) predictions, 1 group x 50M rows x 6 features, f64",
    "C4rls": "C4: rls(half_life=252) predictions, 1 group x 50M rows x 6 features, f64",
    "C5": "C5: lasso(alpha=1e-4) coefficients .over(group), 1000 groups x 10000 rows x 64 features, f64",
}


def _peak():
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        return float(json.loads(pk.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _gen(torch, dev, n, k, G, dtype, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(k, n, dtype=dtype, device=dev, generator=g)
    beta = 1 + 0.25 * torch.randn(G, k, dtype=torch.float64, device=dev, generator=g)
    per = n // G
    y = torch.empty(n, dtype=torch.float64, device=dev)
    step = max(1, G // 16)
    for g0 in range(0, G, step):  # chunked to bound temporary memory
        g1 = min(G, g0 + step)
        xs = x[:, g0 * per:g1 * per].T.reshape(g1 - g0, per, k).to(torch.float64)
        y[g0 * per:g1 * per] = (xs * beta[g0:g1, None, :]).sum(-1).reshape(-1)
    y += 0.1 * torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    return x, y.to(dtype)


def _rel_err(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    m = ~np.isnan(ref)
    assert (np.isnan(got) == np.isnan(ref)).all()
    return float(np.max(np.abs(got[m] - ref[m]) / (1e-3 + np.abs(ref[m])))) if m.any() else 0.0


def run_config(cfg: str, steps: int = 10, warmup: int = 3, with_cpu: bool = True, with_e2e: bool = True, scale: float = 1.0):
    """-> dict in bench.py's line shape for one config."""
    import torch
    import polars_ols_b200 as pls
    from polars_ols_b200 import _lib as L
    from oracle import lib as oracle_lib
    from oracle import semantics as S

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    eng = pls.Engine(0, torch.cuda.current_stream(dev).cuda_stream or 1)
    OL = oracle_lib()
    peak, peak_src = _peak()
    threads = os.cpu_count() or 1          # passed explicitly to the OpenMP legs: `cores` is what actually ran, whatever OMP_NUM_THREADS says
    warmup = max(warmup, 3)

    # ---- per config: device batch, call, algorithmic bytes, units, checker, CPU baseline -----------------------------
    if cfg == "C1":
        n, k, G, per, dt = 1000, 3, 1, 1000, torch.float64
        x, y = _gen(torch, dev, n, k, 1, dt, 1)
        batch = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)])
        kw, mode, units, alg = pls.OLSKwargs().to_c(), L.PREDICTIONS, 1, 40_000
        call = lambda b: eng.least_squares(b, kw, mode, want_validity=False)[0]                                # noqa: E731
    elif cfg == "C3":
        G, per, k, dt = int(100_000 * scale), 256, 16, torch.float32
        n = G * per
        x, y = _gen(torch, dev, n, k, G, dt, 3)
        w = torch.rand(n, dtype=dt, device=dev) + 0.05
        offs = np.arange(G + 1, dtype=np.int64) * per
        batch = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)], weights=pls.Col(w), offsets=offs)
        kw, mode, units, alg = pls.OLSKwargs(alpha=1e-3, l1_ratio=0.5).to_c(), L.PREDICTIONS, G, G * (per * 18 * 4 + per * 8)
        call = lambda b: eng.least_squares(b, kw, mode, want_validity=False)[0]                                # noqa: E731
    elif cfg in ("C4", "C4rls"):
        n, k, G, dt = int(50_000_000 * scale), 6, 1, torch.float64
        per = n
        x, y = _gen(torch, dev, n, k, 1, dt, 4)
        batch = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)])
        mode, units, alg = L.PREDICTIONS, n, n * (56 + 8)
        if cfg == "C4":
            kw = pls.RollingKwargs(window_size=252, min_periods=6, null_policy="drop").to_c()
            call = lambda b: eng.rolling_least_squares(b, kw, mode)[0]                                         # noqa: E731
        else:
            kw = L.RLSKwargs(252.0, 10.0, None, L.NULL_POLICY["drop"], 0)
            call = lambda b: eng.recursive_least_squares(b, kw, mode)[0]                                       # noqa: E731
    elif cfg == "C5":
        G, per, k, dt = int(1000 * scale), 10_000, 64, torch.float64
        n = G * per
        x, y = _gen(torch, dev, n, k, G, dt, 5)
        offs = np.arange(G + 1, dtype=np.int64) * per
        batch = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)], offsets=offs)
        kw, mode, units, alg = pls.OLSKwargs(alpha=1e-4, l1_ratio=1.0).to_c(), L.COEFFICIENTS, G, G * (per * 65 * 8 + 512)
        call = lambda b: eng.least_squares(b, kw, mode, want_validity=False)[0]                                # noqa: E731
    else:
        raise ValueError(cfg)

    # ---- value: inputs resident in HBM ----------------------------------------------------------------------------------
    for _ in range(warmup):
        out = call(batch)
    torch.cuda.synchronize()
    l0 = eng.launch_count
    eng.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = call(batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    kms = eng.profile_drain()
    eng.set_profiling(False)
    launches = eng.launch_count - l0
    k_avg = float(np.mean(kms)) if len(kms) else float("nan")
    value = units / (ms * 1e-3)

    # ---- sample check against the oracle ----------------------------------------------------------------------------------
    xh = [x[i].cpu().numpy() for i in range(k)]
    yh = y.cpu().numpy()
    outh = out.cpu().numpy()
    if cfg == "C1":
        err = _rel_err(outh, S.least_squares(yh, *xh)[0])
    elif cfg == "C3":
        wh = w.cpu().numpy()
        errs = []
        for g in range(0, G, max(1, G // 8)):
            sl = slice(g * per, (g + 1) * per)
            ref = S.least_squares(yh[sl], *[c[sl] for c in xh], sample_weights=wh[sl], kwargs=S.OLSKwargs(alpha=1e-3, l1_ratio=0.5))[0]
            errs.append(_rel_err(outh[sl], ref))
        err = max(errs)
    elif cfg in ("C4", "C4rls"):
        lo, hi, back = n // 2, n // 2 + 20_000, 30_000
        sl = slice(lo - back, hi)
        if cfg == "C4":
            ref = S.rolling_least_squares(yh[sl], *[c[sl] for c in xh], kwargs=S.RollingKwargs(window_size=252, min_periods=6, null_policy="drop"))[0]
        else:
            ref = S.recursive_least_squares(yh[sl], *[c[sl] for c in xh], kwargs=S.RLSKwargs(half_life=252.0))[0]
        err = _rel_err(outh[lo:hi], ref[back:])
    else:
        sel = [0, G // 2, G - 1]
        ref = np.stack([S.solve_elastic_net(yh[g * per:(g + 1) * per], np.ascontiguousarray(np.stack([c[g * per:(g + 1) * per] for c in xh], axis=1)),
                                            1e-4, 1.0, 1000, 1e-5, False, None) for g in sel])
        err = _rel_err(outh[sel], ref)

    # ---- e2e: pinned host columns through the C ABI, results back in host memory ----------------------------------------------
    e2e = None
    if with_e2e:
        heng = pls.Engine(0)
        npdt = np.float32 if dt == torch.float32 else np.float64
        hcols = [heng.pinned_empty((n,), npdt) for _ in range(k + 1)]
        for i in range(k):
            hcols[i][:] = xh[i]
        hcols[k][:] = yh
        hw = None
        if cfg == "C3":
            hw = heng.pinned_empty((n,), npdt)
            hw[:] = wh
        hb = pls.Batch(pls.Col(hcols[k]), [pls.Col(hcols[i]) for i in range(k)], weights=pls.Col(hw) if hw is not None else None,
                       offsets=batch.offsets)
        if cfg in ("C4", "C4rls"):
            hcall = (lambda: heng.rolling_least_squares(hb, kw, mode)[0]) if cfg == "C4" else (lambda: heng.recursive_least_squares(hb, kw, mode)[0])
        else:
            hcall = lambda: heng.least_squares(hb, kw, mode, want_validity=False)[0]                          # noqa: E731
        for _ in range(2):
            ho = hcall()
        esteps = max(2, min(steps, 5))
        t0 = time.perf_counter()
        for _ in range(esteps):
            ho = hcall()
        te = (time.perf_counter() - t0) / esteps
        assert np.allclose(np.asarray(ho), outh, rtol=1e-9, atol=1e-11, equal_nan=True)
        h2d = sum(c_.nbytes for c_ in hcols) + (hw.nbytes if hw is not None else 0)
        e2e = {"value": units / te, "unit": UNIT[cfg], "ms_per_step": 1e3 * te, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(np.asarray(ho).nbytes), "steps": esteps,
               "note": "inputs in page-locked host memory; outputs into pageable numpy arrays through the engine's pinned ring"}
        heng.close()

    # ---- CPU baseline: the oracle port, bounded sample ------------------------------------------------------------------------
    cpu = None
    if with_cpu:
        f64 = [np.ascontiguousarray(c, dtype=np.float64) for c in [yh] + xh]
        arr = (C.c_void_p * (k + 1))(*[c.ctypes.data for c in f64])
        if cfg == "C1":
            o1 = np.empty(n)
            offs1 = np.array([0, n], dtype=np.int64)
            reps, t0 = 2000, time.perf_counter()
            for _ in range(reps):
                OL.orc_grouped_least_squares_predictions(arr, k, None, offs1.ctypes.data, 1, 1, 0.0, 0.0, 1000, 1e-5, 0, 1, o1.ctypes.data)
            t = (time.perf_counter() - t0) / reps
            cpu = {"value": 1.0 / t, "unit": UNIT[cfg], "cores": 1, "kind": "port",
                   "sample": f"{reps} x one 1000 x 3 pivoted-QR OLS + predictions (C port, no polars / FFI overhead), {t * 1e6:.1f} us per call"}
        elif cfg == "C3":
            Gs = min(G, 20_000)
            o3 = np.empty(Gs * per)
            w64 = np.ascontiguousarray(wh, dtype=np.float64)
            offs3 = np.arange(Gs + 1, dtype=np.int64) * per
            OL.orc_grouped_least_squares_predictions(arr, k, w64.ctypes.data, offs3.ctypes.data, Gs, 2, 1e-3, 0.5, 1000, 1e-5, 0, threads, o3.ctypes.data)
            t0, reps = time.perf_counter(), 0
            while time.perf_counter() - t0 < 10.0:
                OL.orc_grouped_least_squares_predictions(arr, k, w64.ctypes.data, offs3.ctypes.data, Gs, 2, 1e-3, 0.5, 1000, 1e-5, 0, threads, o3.ctypes.data)
                reps += 1
            t = (time.perf_counter() - t0) / reps
            cpu = {"value": Gs / t, "unit": UNIT[cfg], "cores": threads, "kind": "port",
                   "sample": f"{reps} x the first {Gs} groups (sqrt(w) scaling + row-major copy + residual-form CD + predictions per group), OpenMP over groups"}
        elif cfg in ("C4", "C4rls"):
            ns = min(n, 2_000_000)
            xr = np.ascontiguousarray(np.stack([c[:ns] for c in f64[1:]], axis=1))
            ys = np.ascontiguousarray(f64[0][:ns])
            co = np.empty((ns, k))
            t0 = time.perf_counter()
            if cfg == "C4":
                OL.orc_solve_rolling_ols(ys.ctypes.data, xr.ctypes.data, ns, k, 252, 6, 0, 0.0, None, 0, co.ctypes.data)
            else:
                OL.orc_solve_recursive_least_squares(ys.ctypes.data, xr.ctypes.data, ns, k, 252.0, 10.0, None, None, co.ctypes.data)
            pred = (xr * co).sum(1)
            t = time.perf_counter() - t0
            cpu = {"value": ns / t, "unit": UNIT[cfg], "cores": 1, "kind": "port",
                   "sample": f"first {ns} rows, sequential recurrence + predictions (the reference walks one series on one thread), {t:.1f} s"}
            del xr, co, pred
        else:
            Gs = min(G, 2 * threads)
            o5 = np.empty((Gs, k))
            offs5 = np.arange(Gs + 1, dtype=np.int64) * per
            t0 = time.perf_counter()
            OL.orc_grouped_least_squares_coefficients(arr, k, offs5.ctypes.data, Gs, 2, 1e-4, 1.0, 1000, 1e-5, 0, threads, o5.ctypes.data)
            t = time.perf_counter() - t0
            cpu = {"value": Gs / t, "unit": UNIT[cfg], "cores": threads, "kind": "port",
                   "sample": f"first {Gs} groups (row-major copy + residual-form CD on 10000 x 64), OpenMP over groups, {t:.1f} s"}
        cpu["gpu_over_cpu_device_resident"] = value / cpu["value"]
        if e2e:
            cpu["gpu_over_cpu_e2e"] = e2e["value"] / cpu["value"]

    dominant = {"C1": "gram_cta_kernel + predict_kernel (latency-bound: 40 KB)", "C3": "gram_multi_kernel<float,2> + cd_thread_kernel<16> + predict_kernel (three passes)",
                "C4": "chunk_totals_kernel + rolling_nbr_kernel<double,6> (window-length chunks, neighbour-shared lag rows)",
                "C4rls": "chunk_totals_kernel (information-form summaries) + rls_scan + rls_fast_main_kernel<double,6>",
                "C5": "gram_wide_kernel<double> (DMMA, 72 per 8 rows) + cd_solve_kernel<32,2>"}[cfg]
    line = {
        "metric": f"{UNIT[cfg].split('/')[0]} per second, {WORKLOAD[cfg]}", "value": value, "unit": UNIT[cfg], "n_gpus": 1, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 inputs, f64 arithmetic" if cfg == "C3" else "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD[cfg], "l2": "inputs exceed the 126 MB L2" if alg > 2e8 else "fits L2 (latency case)",
                   "inputs": "resident in HBM (value) / pinned host memory (e2e)"},
        "e2e": e2e, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak,
                     "traffic": None, "kernel": dominant, "kernel_ms_avg": k_avg, "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                     "note": "whole call (all launches of the path) against the HBM roofline; kernel_ms_avg = the engine's own event pair around the dominant launch(es)"},
        "cpu_baseline": cpu, "max_rel_err_vs_oracle_sample": err,
    }
    if cfg == "C1":
        line["latency_us"] = ms * 1e3
    eng.close()
    return line


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="C1,C3,C4,C4rls,C5")
    ap.add_argument("--out", default="gpurun_out/configs.json")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    res = {}
    for cfg in a.only.split(","):
        r = run_config(cfg, steps=a.steps if cfg != "C1" else 200, with_cpu=not a.no_cpu, with_e2e=not a.no_e2e, scale=a.scale)
        res[cfg] = r
        print(cfg, json.dumps(r), flush=True)
    Path(a.out).parent.mkdir(exist_ok=True)
    Path(a.out).write_text(json.dumps(res, indent=1))
