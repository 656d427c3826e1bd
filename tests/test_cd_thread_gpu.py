"""Elastic net / lasso with many short groups (BASELINE config C3): the thread-per-group coordinate descent
(csrc/cd_thread.cu) against
  * the oracle's solve_elastic_net (src/least_squares.rs:386-492 restated) at 1e-6 (f64) / 1e-4 (f32), and
  * the sub-warp kernel (csrc/cd_solve.cuh), which it must reproduce BIT FOR BIT (same arithmetic in the same order,
    same stop rule)."""
import numpy as np
import pytest

from polars_ols_b200 import Frame, col
from oracle import semantics as S

pytestmark = pytest.mark.gpu


def _frame(G, k, seed, dtype=np.float64, lo=20, hi=120, sparsity=0.5, long_group=0):
    rng = np.random.default_rng(seed)
    sizes = rng.integers(lo, hi, size=G)
    sizes[min(5, G - 1)] = 0                                   # an empty group: no key, no rows
    if long_group:
        sizes[G // 2] = long_group                             # split into several segments (group_seg_off path)
    n = int(sizes.sum())
    x = rng.normal(size=(n, k))
    beta = rng.normal(size=k) * (rng.random(k) > sparsity)
    y = x @ beta + 0.1 * rng.normal(size=n)
    d = {f"x{i + 1}": np.ascontiguousarray(x[:, i]).astype(dtype) for i in range(k)}
    d["y"] = y.astype(dtype)
    d["group"] = np.repeat(np.arange(G), sizes)
    return d, [f"x{i + 1}" for i in range(k)]


def _run(d, names, mode, monkeypatch, env, **kw):
    for key in ("B200OLS_CD_THREAD", "B200OLS_CD_THREAD_BLOCKS"):
        monkeypatch.delenv(key, raising=False)
    for key, v in env.items():
        monkeypatch.setenv(key, str(v))
    e = col("y").least_squares.elastic_net(*names, mode=mode, **kw).over("group")
    return Frame(d).select(e)["coefficients" if mode == "coefficients" else "y"].to_numpy()


def _close(got, ref, rtol, atol):
    assert got.shape == ref.shape
    assert (np.isnan(got) == np.isnan(ref)).all()
    m = ~np.isnan(ref)
    assert (np.abs(got[m] - ref[m]) <= atol + rtol * np.abs(ref[m])).all(), float(np.max(np.abs(got[m] - ref[m])))


@pytest.mark.parametrize("k,method,positive,l1_ratio,long_group", [
    (3, None, False, 0.5, 0), (8, "cd", False, 1.0, 0), (8, "cd_active_set", False, 0.5, 0), (10, "cd", True, 0.5, 0),
    (16, None, False, 0.5, 0), (16, "cd_active_set", False, 1.0, 0), (13, "cd", False, 0.3, 40_000)])
def test_thread_per_group_cd_matches_oracle_and_sub_warp_kernel(k, method, positive, l1_ratio, long_group, monkeypatch):
    d, names = _frame(300, k, seed=k, long_group=long_group)
    kw = dict(alpha=1e-3, l1_ratio=l1_ratio, positive=positive, solve_method=method)
    per_thread = _run(d, names, "coefficients", monkeypatch, {"B200OLS_CD_THREAD": 2}, **kw)
    sub_warp = _run(d, names, "coefficients", monkeypatch, {"B200OLS_CD_THREAD": 0}, **kw)
    if long_group:   # Gram summed over segments: the kernels read row j of G from different triangles (equal to ~1 ulp)
        _close(per_thread, sub_warp, rtol=1e-10, atol=1e-13)
    else:
        assert np.array_equal(per_thread.view(np.int64), sub_warp.view(np.int64)), "the two coordinate-descent kernels differ"
    _, c, _ = S.over(S.least_squares, d["group"], d["y"], *[d[n] for n in names], per_group=True, mode="coefficients",
                     kwargs=S.OLSKwargs(alpha=1e-3, l1_ratio=l1_ratio, positive=positive, solve_method=method))
    _close(per_thread, c, rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("dtype,k", [(np.float32, 16), (np.float64, 7), (np.float32, 9)])
def test_many_groups_take_the_thread_per_group_kernel_by_default(dtype, k, monkeypatch):
    """>= 1024 groups: the default route; predictions / residuals identical to the sub-warp kernel's, oracle within tolerance."""
    d, names = _frame(3000, k, seed=100 + k, dtype=dtype, lo=30, hi=90)
    rng = np.random.default_rng(1)
    d["w"] = rng.uniform(0.05, 1.0, size=len(d["y"])).astype(dtype)
    kw = dict(alpha=1e-3, l1_ratio=0.5, sample_weights="w")
    for mode in ("predictions", "residuals"):
        sub_warp = _run(d, names, mode, monkeypatch, {"B200OLS_CD_THREAD": 0}, **kw)
        got = _run(d, names, mode, monkeypatch, {}, **kw)
        assert np.array_equal(got.view(np.int64), sub_warp.view(np.int64)), mode
    ref = S.over(S.least_squares, d["group"], d["y"], *[d[n] for n in names], sample_weights=d["w"], kwargs=S.OLSKwargs(alpha=1e-3, l1_ratio=0.5))
    ref = ref[0] if isinstance(ref, tuple) else ref
    tol = 1e-4 if dtype == np.float32 else 1e-6
    got = _run(d, names, "predictions", monkeypatch, {}, **kw)
    _close(got, np.asarray(ref, dtype=np.float64), rtol=tol, atol=tol)
