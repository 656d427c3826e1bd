"""Elastic-net / lasso predictions written by the coordinate-descent kernel itself (cd_solve.cuh: cd_predict_group) must be
BIT-IDENTICAL to the separate predict_kernel pass they replace, and match the oracle (src/least_squares.rs:386-492 +
src/expressions.rs:175-195 + polars_ols/least_squares.py:234-239) — weights, intercept, nulls, shuffled groups, f32."""
import numpy as np
import pytest

import polars_ols_b200 as pls
from polars_ols_b200 import Frame, col
from oracle import semantics as S

pytestmark = pytest.mark.gpu


def _frame(n, k, G, seed, dtype=np.float64, missing=0.0):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, k))
    y = x[:, : max(1, k // 2)].sum(1) + 0.1 * rng.normal(size=n)
    d = {f"x{i}": np.ascontiguousarray(x[:, i]).astype(dtype) for i in range(k)}
    d["y"] = y.astype(dtype)
    d["w"] = rng.uniform(0.1, 4.0, size=n).astype(dtype)
    d["g"] = rng.integers(G, size=n)
    if missing:
        for c in list(d):
            if c not in ("g", "w"):
                d[c] = (d[c], rng.random(n) >= missing)
    return d


@pytest.mark.parametrize("k,weights,intercept,policy,missing,dtype", [
    (16, True, False, "ignore", 0.0, np.float32), (8, False, True, "ignore", 0.0, np.float64), (5, True, True, "drop", 0.1, np.float64),
    (12, False, False, "zero", 0.05, np.float64), (15, True, True, "drop_y_zero_x", 0.1, np.float32), (3, False, False, "drop_zero", 0.2, np.float64)])
@pytest.mark.parametrize("mode", ["predictions", "residuals"])
def test_cd_kernel_predictions_match_predict_pass_and_oracle(monkeypatch, k, weights, intercept, policy, missing, dtype, mode):
    d = _frame(30_000, k, 37, seed=k, dtype=dtype, missing=missing)
    names = [f"x{i}" for i in range(k)]
    kw = dict(alpha=1e-3, l1_ratio=0.5, add_intercept=intercept, null_policy=policy, mode=mode)
    if weights:
        kw["sample_weights"] = "w"
    e = col("y").least_squares.elastic_net(*names, **kw).over("g")
    fused = Frame(d).select(e, engine=pls.Engine(0))["y"]
    monkeypatch.setenv("B200OLS_CD_PRED", "0")
    two_pass = Frame(d).select(e, engine=pls.Engine(0))["y"]
    a, b = fused.to_numpy(), two_pass.to_numpy()
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])
    assert np.array_equal(fused.is_null(), two_pass.is_null())
    okw = S.OLSKwargs(alpha=1e-3, l1_ratio=0.5, null_policy=policy)
    ref = S.over(S.least_squares, d["g"], d["y"], *[d[n] for n in names], sample_weights=d["w"] if weights else None,
                 add_intercept=intercept, mode=mode, kwargs=okw)
    refv = np.where(ref[1], ref[0], np.nan) if ref[1] is not None else ref[0]
    tol = 1e-4 if dtype == np.float32 else 1e-6
    m = ~np.isnan(refv)
    assert np.array_equal(np.isnan(a), np.isnan(refv))
    assert np.max(np.abs(a[m] - refv[m]) / (1e-2 + np.abs(refv[m]))) < tol
