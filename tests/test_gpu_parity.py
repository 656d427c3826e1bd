"""Parity tests proper (`-m gpu`): the CUDA path, called through the public API -> ctypes -> C ABI of
libb200ols.so, against the CPU oracle on the same seeded inputs and against the reference's README
known-answer frame.  Tolerances: 1e-6 relative for f64, 1e-4 for f32 inputs (BASELINE.json north_star).
The structure follows the reference's tests/test_ols.py (cited per test)."""
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import semantics as S  # noqa: E402  (checker only)

import polars_ols_b200 as pls  # noqa: E402
from polars_ols_b200 import Frame, OLSKwargs, RLSKwargs, RollingKwargs, col  # noqa: E402

GOLD = json.loads((Path(__file__).parent / "golden" / "readme_frame.json").read_text())
RTOL, ATOL = 1e-6, 1e-9


def _make_data(n_samples=5000, n_features=2, n_groups=None, scale=0.1, sparsity=0.0, add_missing=False, seed=0,
               dtype=np.float64):
    """tests/test_ols.py:22-51"""
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n_samples, n_features))
    eps = rng.normal(size=n_samples, scale=scale)
    y = x[:, : int(n_features * (1.0 - sparsity))].sum(1) + eps
    d = {f"x{i + 1}": np.ascontiguousarray(x[:, i]).astype(dtype) for i in range(n_features)}
    d["y"] = y.astype(dtype)
    if n_groups is not None:
        d["group"] = rng.integers(n_groups, size=n_samples)
    if add_missing:
        for c in [c for c in d if c != "group"]:
            d[c] = (d[c], rng.random(n_samples) >= 0.1)
    return d


def _xs(d):
    return [k for k in d if k.startswith("x")]


def _close(got, ref, rtol=RTOL, atol=ATOL):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert (np.isnan(got) == np.isnan(ref)).all(), "null/NaN pattern differs"
    m = ~np.isnan(ref)
    if m.any():
        err = np.abs(got[m] - ref[m]) / (atol / rtol + np.abs(ref[m]))
        assert err.max() <= rtol, f"max rel err {err.max():.3e}"


def _oracle_cols(d, names):
    return [d[n] for n in names]


def _ref(values_mask):
    v, m = values_mask
    return np.where(m, v, np.nan) if m is not None else v


# ----------------------------------------------------------------------------------- golden frame
def test_readme_golden_frame():
    F = Frame({k: np.asarray(v, dtype=np.float64) for k, v in GOLD["frame"].items()})
    tolc, tolr = GOLD["printed_abs_tol_coefficients"], GOLD["printed_abs_tol_round2"]
    r = F.select(col("y").least_squares.ols(col("x1"), col("x2"), add_intercept=True, mode="coefficients"))["coefficients"]
    assert r.fields == ["x1", "x2", "const"]
    assert np.allclose(r.to_numpy()[0], GOLD["coefficients_ols_intercept"], atol=tolc, rtol=0)
    r = F.select(col("y").least_squares.ols("x1", "x2", add_intercept=True, mode="coefficients").over("group"))["coefficients"]
    for i, k in enumerate(r.keys):
        assert np.allclose(r.to_numpy()[i], GOLD["coefficients_ols_intercept_by_group"][str(int(k))], atol=tolc, rtol=0)
    assert r.to_numpy(broadcast=True).shape == (10, 3)               # tests/test_ols.py:404-433 shapes
    r = F.select(col("y").least_squares.rls(col("x1"), col("x2"), mode="coefficients").over("group"))["coefficients"]
    assert np.allclose(r.to_numpy()[:5], GOLD["coefficients_rls_group1"], atol=tolc, rtol=0)
    r = F.select(col("y").least_squares.lasso("x1", "x2", alpha=0.0001, add_intercept=True).over("group"))["y"]
    assert np.allclose(r.to_numpy()[:5], GOLD["predictions_lasso_head5_round2"], atol=tolr, rtol=0)
    r = F.select(col("y").least_squares.wls("x1", "x2", sample_weights="weights"))["y"]
    assert np.allclose(r.to_numpy()[:5], GOLD["predictions_wls_head5_round2"], atol=tolr, rtol=0)


# ----------------------------------------------------------------------------------- static models
@pytest.mark.parametrize("solve_method", ["qr", "chol", "lu", None])
def test_ols(solve_method):                                                # tests/test_ols.py:54-73 (C1 shape)
    d = _make_data(1000, 3)
    F = Frame(d)
    r = F.select(col("y").least_squares.ols(*_xs(d), mode="coefficients", solve_method=solve_method))["coefficients"]
    ref = S.least_squares(d["y"], *_oracle_cols(d, _xs(d)), mode="coefficients", kwargs=S.OLSKwargs(solve_method=solve_method))
    _close(r.to_numpy()[0], _ref(ref))
    p = F.select(col("y").least_squares.ols(*_xs(d), solve_method=solve_method))["y"]
    _close(p.to_numpy(), _ref(S.least_squares(d["y"], *_oracle_cols(d, _xs(d)), kwargs=S.OLSKwargs(solve_method=solve_method))))
    e = F.select(col("y").least_squares.ols(*_xs(d), mode="residuals"))["y"]
    _close(e.to_numpy(), _ref(S.least_squares(d["y"], *_oracle_cols(d, _xs(d)), mode="residuals")))


def test_ols_svd_matches_lapack():                                          # tests/test_ols.py:54-73 [svd]
    d = _make_data(1000, 3, n_groups=4)
    r = Frame(d).select(col("y").least_squares.ols(*_xs(d), mode="coefficients", solve_method="svd").over("group"))["coefficients"]
    keys, c, m = S.over(S.least_squares, d["group"], d["y"], *_oracle_cols(d, _xs(d)), per_group=True, mode="coefficients",
                        kwargs=S.OLSKwargs(solve_method="svd"))
    _close(r.to_numpy(), c)
    assert (pls.get_engine(0).last_group_flags(4) & 32).all()


@pytest.mark.parametrize("k", [2, 10, 40, 64])
def test_fit_wide_takes_the_min_norm_path(k):                              # tests/test_ols.py:272-312 (n = 10 rows)
    d = _make_data(10, k, seed=k)
    r = Frame(d).select(col("y").least_squares.ols(*_xs(d), mode="coefficients"))["coefficients"]
    ref = S.least_squares(d["y"], *_oracle_cols(d, _xs(d)), mode="coefficients")          # lstsq (dgelsd) when n <= k
    _close(r.to_numpy()[0], _ref(ref), rtol=1e-6, atol=1e-9)
    p = Frame(d).select(col("y").least_squares.ols(*_xs(d)))["y"].to_numpy()
    if k >= 10:
        assert np.corrcoef(p, d["y"])[0, 1] > 1 - 1e-5                      # interpolates exactly


def test_fit_multi_collinear_svd_and_qr():                                 # tests/test_ols.py:315-360
    rng = np.random.default_rng(0)
    n = 1000
    x1, x2 = rng.normal(size=n), rng.normal(size=n)
    x3 = x1 + x2                                                            # exactly collinear
    y = x1 + x2 + x3 + 0.01 * rng.normal(size=n)
    d = {"y": y, "x1": x1, "x2": x2, "x3": x3}
    r = Frame(d).select(col("y").least_squares.ols("x1", "x2", "x3", mode="coefficients", solve_method="svd"))["coefficients"]
    ref = np.linalg.lstsq(np.c_[x1, x2, x3], y, rcond=None)[0]              # min-norm
    assert np.allclose(r.to_numpy()[0], ref, rtol=1e-6, atol=1e-8)
    p = Frame(d).select(col("y").least_squares.ols("x1", "x2", "x3", solve_method="svd"))["y"].to_numpy()
    assert np.allclose(p, np.c_[x1, x2, x3] @ ref, rtol=1e-6, atol=1e-8)
    q = Frame(d).select(col("y").least_squares.ols("x1", "x2", "x3", solve_method="qr"))["y"].to_numpy()
    assert np.isfinite(q).all()                                             # reference: qr predictions stay finite


def test_ridge_svd():                                                      # tests/test_ols.py:475-503 [svd]
    d = _make_data(5000, 3, n_groups=3)
    r = Frame(d).select(col("y").least_squares.ridge(*_xs(d), alpha=0.01, solve_method="svd", mode="coefficients").over("group"))["coefficients"]
    keys, c, m = S.over(S.least_squares, d["group"], d["y"], *_oracle_cols(d, _xs(d)), per_group=True, mode="coefficients",
                        kwargs=S.OLSKwargs(alpha=0.01, l1_ratio=0.0, solve_method="svd"))
    _close(r.to_numpy(), c)


@pytest.mark.parametrize("k,n_groups,n", [(8, 300, 60000), (3, 7, 5000), (16, 50, 20000)])
def test_ridge_coefficients_over_random_groups(k, n_groups, n):           # C2 shape, tests/test_ols.py:380-401,475-503
    d = _make_data(n, k, n_groups=n_groups, seed=3)
    F = Frame(d)
    r = F.select(col("y").least_squares.ridge(*_xs(d), alpha=1e-3, mode="coefficients").over("group"))["coefficients"]
    keys, c, m = S.over(S.least_squares, d["group"], d["y"], *_oracle_cols(d, _xs(d)), per_group=True, mode="coefficients",
                        kwargs=S.OLSKwargs(alpha=1e-3, l1_ratio=0.0))
    assert (r.keys == keys).all()
    _close(r.to_numpy(), c)
    for mode in ("predictions", "residuals"):
        p = F.select(col("y").least_squares.ridge(*_xs(d), alpha=1e-3, mode=mode).over("group"))["y"]
        ref = S.over(S.least_squares, d["group"], d["y"], *_oracle_cols(d, _xs(d)), mode=mode,
                     kwargs=S.OLSKwargs(alpha=1e-3, l1_ratio=0.0))
        _close(p.to_numpy(), _ref(ref))


def test_contiguous_equal_groups_c2_small():
    G, n, k = 500, 1000, 8
    rng = np.random.default_rng(0)
    x = rng.normal(size=(G * n, k))
    beta = 1 + 0.25 * rng.normal(size=(G, k))
    y = (x.reshape(G, n, k) * beta[:, None, :]).sum(-1).reshape(-1) + 0.1 * rng.normal(size=G * n)
    d = {f"x{i}": np.ascontiguousarray(x[:, i]) for i in range(k)}
    d["y"] = y
    d["group"] = np.repeat(np.arange(G), n)
    r = Frame(d).select(col("y").least_squares.ridge(*[f"x{i}" for i in range(k)], alpha=1e-3, mode="coefficients").over("group"))["coefficients"]
    xg = x.reshape(G, n, k)
    ref = np.linalg.solve(np.einsum("gni,gnj->gij", xg, xg) + 1e-3 * np.eye(k), np.einsum("gni,gn->gi", xg, y.reshape(G, n))[..., None])[..., 0]
    _close(r.to_numpy(), ref)
    flags = pls.get_engine(0).last_group_flags(G)
    assert (flags == 0).all()


def test_large_group_is_split_into_segments():
    d = _make_data(50_000, 5, seed=7)
    d["group"] = np.concatenate([np.zeros(30_000, int), np.ones(19_999, int), np.full(1, 2)])
    F = Frame(d)
    r = F.select(col("y").least_squares.ridge(*_xs(d), alpha=0.5, mode="coefficients", add_intercept=True).over("group"))["coefficients"]
    keys, c, m = S.over(S.least_squares, d["group"], d["y"], *_oracle_cols(d, _xs(d)), per_group=True, mode="coefficients",
                        add_intercept=True, kwargs=S.OLSKwargs(alpha=0.5, l1_ratio=0.0))
    _close(r.to_numpy(), c)
    p = F.select(col("y").least_squares.ridge(*_xs(d), alpha=0.5, add_intercept=True).over("group"))["y"]
    _close(p.to_numpy(), _ref(S.over(S.least_squares, d["group"], d["y"], *_oracle_cols(d, _xs(d)), add_intercept=True,
                                     kwargs=S.OLSKwargs(alpha=0.5, l1_ratio=0.0))))


@pytest.mark.parametrize("null_policy", ["drop", "drop_zero", "drop_y_zero_x", "zero", "ignore"])
@pytest.mark.parametrize("mode", ["predictions", "residuals", "coefficients"])
def test_missing_data(null_policy, mode):                                  # tests/test_ols.py:130-249
    d = _make_data(4000, 2, n_groups=5, add_missing=True)
    F = Frame(d)
    r = F.select(col("y").least_squares.ols("x1", "x2", null_policy=null_policy, mode=mode).over("group"))
    r = r["coefficients" if mode == "coefficients" else "y"]
    kw = S.OLSKwargs(null_policy=null_policy)
    if mode == "coefficients":
        keys, c, m = S.over(S.least_squares, d["group"], d["y"], d["x1"], d["x2"], per_group=True, mode=mode, kwargs=kw)
        _close(r.to_numpy(), np.where(m, c, np.nan))
    else:
        ref = S.over(S.least_squares, d["group"], d["y"], d["x1"], d["x2"], mode=mode, kwargs=kw)
        got = r.to_numpy()
        refv = _ref(ref)
        if null_policy == "ignore":   # NaN (not null) propagates: whole groups are NaN
            assert (np.isnan(got) == np.isnan(refv)).all()
        else:
            _close(got, refv)
            assert (r.is_null() == ~ref[1]).all()


def test_all_empty_data():                                                 # tests/test_ols.py:252-269
    F = Frame({"A": (np.array([0.0, 2, 0, 4]), np.array([False, True, False, True])),
               "B": (np.array([1.0, 0, 3, 0]), np.array([True, False, True, False]))})
    r = F.select(col("A").least_squares.ols(col("B"), mode="residuals", null_policy="drop", solve_method="svd"))["A"]
    assert r.is_null().all()


def test_wls_intercept_and_f32():                                          # tests/test_ols.py:506-541; C3 dtype
    d = _make_data(3000, 4, n_groups=6, seed=5)
    rng = np.random.default_rng(1)
    d["w"] = rng.uniform(0.05, 1.0, size=3000)
    d["w"] = (d["w"], rng.random(3000) >= 0.02)                             # null weights -> sqrt_w = 1e-12
    for mode in ("coefficients", "predictions", "residuals"):
        r = Frame(d).select(col("y").least_squares.wls(*_xs(d), sample_weights="w", add_intercept=True, mode=mode).over("group"))
        r = r["coefficients" if mode == "coefficients" else "y"]
        if mode == "coefficients":
            keys, c, m = S.over(S.least_squares, d["group"], d["y"], *_oracle_cols(d, _xs(d)), sample_weights=d["w"],
                                per_group=True, add_intercept=True, mode=mode)
            _close(r.to_numpy(), c)
        else:
            ref = S.over(S.least_squares, d["group"], d["y"], *_oracle_cols(d, _xs(d)), sample_weights=d["w"],
                         add_intercept=True, mode=mode)
            _close(r.to_numpy(), _ref(ref), rtol=1e-6, atol=1e-6)
    # f32 inputs: arithmetic is f64 on f32-rounded inputs; tolerance 1e-4
    d32 = {k: (v.astype(np.float32) if k != "group" and not isinstance(v, tuple) else v) for k, v in d.items()}
    d32["w"] = (d["w"][0].astype(np.float32), d["w"][1])
    r = Frame(d32).select(col("y").least_squares.elastic_net(*_xs(d), alpha=1e-3, l1_ratio=0.5, sample_weights="w").over("group"))["y"]
    ref = S.over(S.least_squares, d32["group"], d32["y"], *_oracle_cols(d32, _xs(d)), sample_weights=d32["w"],
                 kwargs=S.OLSKwargs(alpha=1e-3, l1_ratio=0.5))
    _close(r.to_numpy(), _ref(ref), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("k,sparsity,alpha,l1_ratio,method,positive", [
    (10, 0.5, 0.001, 0.5, "cd", False), (10, 0.5, 0.001, 0.5, "cd_active_set", False), (16, 0.5, 0.001, 0.5, None, False),
    (64, 0.9, 0.0001, 1.0, None, False), (4, 0.0, 0.1, 0.5, None, True), (100, 0.9, 0.01, 0.5, "cd", False)])
def test_elastic_net(k, sparsity, alpha, l1_ratio, method, positive):      # tests/test_ols.py:561-630; C3/C5 shapes
    d = _make_data(6000, k, n_groups=3, sparsity=sparsity, seed=11)
    expr = col("y").least_squares.elastic_net(*_xs(d), alpha=alpha, l1_ratio=l1_ratio, positive=positive,
                                              solve_method=method, mode="coefficients").over("group")
    r = Frame(d).select(expr)["coefficients"]
    keys, c, m = S.over(S.least_squares, d["group"], d["y"], *_oracle_cols(d, _xs(d)), per_group=True, mode="coefficients",
                        kwargs=S.OLSKwargs(alpha=alpha, l1_ratio=l1_ratio, positive=positive, solve_method=method))
    _close(r.to_numpy(), c, rtol=1e-6, atol=1e-8)


def test_ill_conditioned_ols_uses_qr_kernel():                             # tests/test_ols.py:315-360 (qr finite / accurate)
    rng = np.random.default_rng(0)
    n = 2000
    x1 = rng.normal(size=n)
    x2 = x1 + 1e-5 * rng.normal(size=n)      # cond(X) ~ 1e5 -> cond(G) ~ 1e10: normal equations lose 1e-6
    x3 = rng.normal(size=n)
    y = x1 + 2 * x2 - x3 + 0.01 * rng.normal(size=n)
    d = {"y": y, "x1": x1, "x2": x2, "x3": x3}
    r = Frame(d).select(col("y").least_squares.ols("x1", "x2", "x3", mode="coefficients"))["coefficients"]
    ref = np.linalg.lstsq(np.c_[x1, x2, x3], y, rcond=None)[0]
    _close(r.to_numpy()[0], ref, rtol=1e-6, atol=1e-8)
    assert pls.get_engine(0).last_group_flags(1)[0] & 4


# ----------------------------------------------------------------------------------- moving-window models
@pytest.mark.parametrize("half_life,p0,mean,policy", [(None, 10.0, None, "drop"), (252.0, 10.0, None, "drop"),
                                                      (None, 1e6, None, "drop_y_zero_x"), (20.0, 0.01, [0.25, 0.25], "zero")])
@pytest.mark.parametrize("mode", ["coefficients", "predictions", "residuals"])
def test_recursive_least_squares(half_life, p0, mean, policy, mode):       # tests/test_ols.py:633-715
    d = _make_data(6000, 2, n_groups=3, add_missing=True, seed=2)
    kw = dict(half_life=half_life, initial_state_covariance=p0, initial_state_mean=mean, null_policy=policy)
    r = Frame(d).select(col("y").least_squares.rls("x1", "x2", mode=mode, **kw).over("group"))
    r = r["coefficients" if mode == "coefficients" else "y"]
    ref = S.over(S.recursive_least_squares, d["group"], d["y"], d["x1"], d["x2"], mode=mode, kwargs=S.RLSKwargs(**kw))
    got, refv = r.to_numpy(), _ref(ref)
    _close(got, refv, rtol=1e-6, atol=1e-8)          # from row 0: the prior-dominated head runs the reference's literal arithmetic


@pytest.mark.parametrize("window,min_periods,policy,alpha", [(21, None, "drop", None), (252, 5, "drop", None),
                                                             (50, 10, "drop_window", None), (30, 3, "zero", 0.01),
                                                             (100, None, "drop_zero", None)])
@pytest.mark.parametrize("mode", ["coefficients", "predictions", "residuals"])
def test_rolling_least_squares(window, min_periods, policy, alpha, mode):  # tests/test_ols.py:718-841
    d = _make_data(5000, 3, n_groups=2, add_missing=True, seed=4)
    kw = dict(window_size=window, min_periods=min_periods, null_policy=policy, alpha=alpha)
    r = Frame(d).select(col("y").least_squares.rolling_ols("x1", "x2", "x3", mode=mode, **kw).over("group"))
    r = r["coefficients" if mode == "coefficients" else "y"]
    ref = S.over(S.rolling_least_squares, d["group"], d["y"], d["x1"], d["x2"], d["x3"], mode=mode, kwargs=S.RollingKwargs(**kw))
    got, refv = r.to_numpy(), _ref(ref)
    assert (np.isnan(got) == np.isnan(refv)).all()
    ok = ~np.isnan(refv) & (np.abs(refv) < 1e6)       # singular warm-up windows are rounding noise in the reference too
    assert np.allclose(got[ok], refv[ok], rtol=1e-6, atol=1e-7)


def test_rolling_insufficient_data():                                      # tests/test_ols.py:775-806
    d = {"y": np.array([1.0, 2.0, 3.0]), "x": np.array([1.0, 2.0, 4.0])}
    for mp, n_valid in ((2, 2), (3, 1), (4, 0)):
        r = Frame(d).select(col("y").least_squares.rolling_ols("x", window_size=10, min_periods=mp, mode="coefficients"))["coefficients"]
        assert (~np.isnan(r.to_numpy()[:, 0])).sum() == n_valid


def test_single_long_series_rls_and_rolling_c4_small():                    # C4 shape (1 group, k=6) at 300k rows
    d = _make_data(300_000, 6, seed=9)
    names = _xs(d)
    r = Frame(d).select(col("y").least_squares.rolling_ols(*names, window_size=252, min_periods=6, mode="coefficients"))["coefficients"]
    ref = S.rolling_least_squares(d["y"], *_oracle_cols(d, names), mode="coefficients",
                                  kwargs=S.RollingKwargs(window_size=252, min_periods=6, null_policy="drop"))
    _close(r.to_numpy(), _ref(ref), rtol=1e-6, atol=1e-8)
    r = Frame(d).select(col("y").least_squares.rls(*names, half_life=252.0))["y"]
    ref = S.recursive_least_squares(d["y"], *_oracle_cols(d, names), kwargs=S.RLSKwargs(half_life=252.0))
    _close(r.to_numpy(), _ref(ref), rtol=1e-6, atol=1e-8)


# ----------------------------------------------------------------------------------- device-resident frames
def test_device_resident_torch_frame():
    import torch
    d = _make_data(40_000, 8, n_groups=None, seed=6)
    G = 40
    d["group"] = np.repeat(np.arange(G), 1000)
    dev = {k: torch.as_tensor(v, device="cuda") for k, v in d.items() if k != "group"}
    dev["group"] = d["group"]
    r = Frame(dev).select(col("y").least_squares.ridge(*_xs(d), alpha=1e-3, mode="coefficients").over("group"))["coefficients"]
    assert r.values.is_cuda
    keys, c, m = S.over(S.least_squares, d["group"], d["y"], *_oracle_cols(d, _xs(d)), per_group=True, mode="coefficients",
                        kwargs=S.OLSKwargs(alpha=1e-3, l1_ratio=0.0))
    _close(r.to_numpy(), c)
    p = Frame(dev).select(col("y").least_squares.ridge(*_xs(d), alpha=1e-3).over("group"))["y"]
    _close(p.to_numpy(), _ref(S.over(S.least_squares, d["group"], d["y"], *_oracle_cols(d, _xs(d)),
                                     kwargs=S.OLSKwargs(alpha=1e-3, l1_ratio=0.0))))


def test_full_size_c2_normal_equation_property():
    """BASELINE.json configs[1] at full size (10k groups x 1k rows x 8 f64, alpha=1e-3): size-independent
    property instead of the oracle — the ridge normal equations hold per group:
    X^T (y - X b) - alpha b = 0, evaluated in f64 with torch on the device; plus the oracle on 64 groups."""
    import torch
    G, n, k, alpha = 10_000, 1000, 8, 1e-3
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(k, G * n, dtype=torch.float64, device="cuda", generator=g)
    beta = 1 + 0.25 * torch.randn(G, k, dtype=torch.float64, device="cuda", generator=g)
    y = (x.T.reshape(G, n, k) * beta[:, None, :]).sum(-1).reshape(-1) + 0.1 * torch.randn(G * n, dtype=torch.float64, device="cuda", generator=g)
    cols = {f"x{i}": x[i] for i in range(k)}
    cols["y"] = y
    eng = pls.get_engine(0, torch_stream=True)
    b = pls.Batch(pls.as_col(y), [pls.as_col(x[i]) for i in range(k)], offsets=np.arange(G + 1, dtype=np.int64) * n)
    coef, _ = eng.least_squares(b, OLSKwargs(alpha=alpha, l1_ratio=0.0).to_c(), 2)
    torch.cuda.synchronize()
    xg = x.T.reshape(G, n, k)
    resid = y.reshape(G, n) - (xg * coef[:, None, :]).sum(-1)
    grad = torch.einsum("gnk,gn->gk", xg, resid) - alpha * coef
    scale = torch.einsum("gnk,gn->gk", xg.abs(), y.reshape(G, n).abs())
    assert float((grad.abs() / scale).max()) < 1e-12
    sel = np.arange(0, G, G // 64)
    xs = x.T.reshape(G, n, k)[sel].cpu().numpy()
    ys = y.reshape(G, n)[sel].cpu().numpy()
    ref = np.stack([S.solve_ridge(np.ascontiguousarray(ys[i]), np.ascontiguousarray(xs[i]), alpha, None, None) for i in range(len(sel))])
    _close(coef[sel].cpu().numpy(), ref)


# ----------------------------------------------------------------------------------- predict (§8f rank 1)
def test_predict_matches_dot_product():                                    # tests/test_ols.py:903-944
    d = _make_data(5000, 2, n_groups=3)
    F = Frame(d)
    coef = F.select(col("y").least_squares.ols("x1", "x2", mode="coefficients").over("group"))["coefficients"]
    test = _make_data(40, 2, n_groups=3, seed=12)
    cb = coef.to_numpy()[np.searchsorted(coef.keys, test["group"])]         # the join on "group"
    Ft = Frame({"coefficients": cb, "x1": test["x1"], "x2": test["x2"]})
    p = Ft.select(col("coefficients").least_squares.predict(col("x1"), col("x2"), name="predictions", null_policy="zero"))["predictions"]
    expected = (np.c_[test["x1"], test["x2"]] * cb).sum(1)
    _close(p.to_numpy(), expected)


def test_predict_intercept_and_null_policies():                            # tests/test_ols.py:947-966
    d = {"y": np.array([1.0, 2, 3, 4]), "x1": np.array([3.0, 4, 5, 6]), "x2": np.array([4.0, 5, 6, 7]), "x3": np.array([5.0, 6, 7, 8])}
    F = Frame(d)
    c = F.select(col("y").least_squares.ridge("x1", "x2", "x3", alpha=1e-9, add_intercept=True, mode="coefficients"))["coefficients"]
    F["coefficients"] = np.broadcast_to(c.to_numpy()[0], (4, 4)).copy()
    p = F.select(col("coefficients").least_squares.predict("x1", "x2", "x3", add_intercept=True).alias("y_pred"))["y_pred"]
    ref = S.predict([F["coefficients"][:, j] for j in range(4)], [d["x1"], d["x2"], d["x3"]], "zero", True)
    _close(p.to_numpy(), _ref(ref))
    # nulls in the features
    dm = _make_data(3000, 3, add_missing=True, seed=21)
    cb = np.random.default_rng(0).normal(size=(3000, 3))
    Fm = Frame({**dm, "coefficients": cb})
    for pol in ("zero", "ignore", "drop"):
        p = Fm.select(col("coefficients").least_squares.predict("x1", "x2", "x3", null_policy=pol))["predictions"]
        ref = S.predict([cb[:, j] for j in range(3)], [dm["x1"], dm["x2"], dm["x3"]], pol)
        _close(p.to_numpy(), _ref(ref))


# ----------------------------------------------------------------------------------- time-axis shards (§8e)
def _run_time_sharded(expr, frame, world):
    """all ranks of parallel.time_sharded, one after the other in this process (the exchange is the only
    cross-rank step: pass 1 records every rank's shard map, pass 2 hands all of them to every rank)."""
    from polars_ols_b200.parallel import time_sharded
    maps = {}
    if expr.kind == "recursive_least_squares":
        for rank in range(world):
            def record(local, rank=rank):
                maps[rank] = local
                return [local] * world
            time_sharded(expr, frame, rank, world, exchange=record)
    outs = []
    for rank in range(world):
        r, (r0, r1) = time_sharded(expr, frame, rank, world, exchange=lambda local: [maps[q] for q in range(world)])
        v = r.to_numpy()
        assert v.shape[0] == r1 - r0
        outs.append(v)
    return np.concatenate(outs, axis=0)


@pytest.mark.parametrize("half_life,mean,policy", [(None, None, "drop"), (252.0, [0.5, -0.5, 0.25], "drop"), (30.0, None, "zero")])
@pytest.mark.parametrize("mode", ["coefficients", "predictions"])
def test_time_sharded_rls_matches_whole_series(half_life, mean, policy, mode):
    d = _make_data(30_000, 3, add_missing=True, seed=12)
    kw = dict(half_life=half_life, initial_state_mean=mean, null_policy=policy)
    expr = col("y").least_squares.rls("x1", "x2", "x3", mode=mode, **kw)
    got = _run_time_sharded(expr, Frame(d), 3)
    ref = _ref(S.recursive_least_squares(d["y"], d["x1"], d["x2"], d["x3"], mode=mode, kwargs=S.RLSKwargs(**kw)))
    _close(got, ref, rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("window,min_periods,policy,missing", [(252, 6, "drop", False), (50, 10, "drop", True),
                                                               (64, 8, "drop_window", True), (40, None, "zero", True)])
@pytest.mark.parametrize("mode", ["coefficients", "residuals"])
def test_time_sharded_rolling_matches_whole_series(window, min_periods, policy, missing, mode):
    d = _make_data(30_000, 3, add_missing=missing, seed=13)
    kw = dict(window_size=window, min_periods=min_periods, null_policy=policy)
    expr = col("y").least_squares.rolling_ols("x1", "x2", "x3", mode=mode, **kw)
    got = _run_time_sharded(expr, Frame(d), 4)
    ref = _ref(S.rolling_least_squares(d["y"], d["x1"], d["x2"], d["x3"], mode=mode, kwargs=S.RollingKwargs(**kw)))
    assert (np.isnan(got) == np.isnan(ref)).all()
    ok = ~np.isnan(ref) & (np.abs(ref) < 1e6)
    assert np.allclose(got[ok], ref[ok], rtol=1e-6, atol=1e-7)


def test_rls_state_abi_against_information_form():
    """b200ols_recursive_least_squares_state: (A, b, D) leaving each series, prior and continued."""
    from polars_ols_b200 import _lib as L
    from polars_ols_b200.engine import Batch, as_col
    rng = np.random.default_rng(5)
    n, k, lam = 4000, 3, np.exp(np.log(0.5) / 100.0)
    x = rng.standard_normal((n, k))
    y = x @ np.array([1.0, -2.0, 0.5]) + 0.1 * rng.standard_normal(n)
    offs = np.array([0, 1500, 1500, 4000], dtype=np.int64)      # three series, the middle one empty
    b = Batch(as_col(y), [as_col(np.ascontiguousarray(x[:, j])) for j in range(k)], offsets=offs)
    eng = pls.get_engine(0)
    mean = np.array([0.1, 0.2, 0.3])
    for info in (None, rng.standard_normal((3, k * k + k))):
        if info is not None:
            for g in range(3):
                a = rng.standard_normal((k, k))
                info[g, :k * k] = (a @ a.T + np.eye(k)).reshape(-1)
        kw = L.RLSKwargs(100.0, 5.0, mean.ctypes.data, L.NULL_POLICY["drop"], 0, None if info is None else info.ctypes.data)
        st = eng.recursive_least_squares_state(b, kw, (mean, info))
        for g in range(3):
            if info is None:
                A, bb = np.eye(k) / 5.0, mean / 5.0
            else:
                A, bb = info[g, :k * k].reshape(k, k).copy(), info[g, k * k:].copy()
            D = 1.0
            for r in range(offs[g], offs[g + 1]):
                A, bb, D = lam * A + np.outer(x[r], x[r]), lam * bb + x[r] * y[r], D * lam
            np.testing.assert_allclose(st[g, :k * k].reshape(k, k), A, rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(st[g, k * k:k * k + k], bb, rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(st[g, -1], D, rtol=1e-10, atol=1e-300)


def test_moving_models_with_ten_features_and_weights():                    # tests/test_ols.py:969-996 shape (k = 10)
    d = _make_data(20_000, 10, n_groups=5, add_missing=True, seed=21)
    rng = np.random.default_rng(0)
    d["weights"] = rng.uniform(0.1, 10.0, size=20_000)
    names = _xs(d)
    kw = dict(window_size=100, min_periods=12, null_policy="drop")
    r = Frame(d).select(col("y").least_squares.rolling_ols(*names, sample_weights="weights", mode="coefficients", **kw).over("group"))
    ref = S.over(S.rolling_least_squares, d["group"], d["y"], *_oracle_cols(d, names), sample_weights=d["weights"],
                 mode="coefficients", kwargs=S.RollingKwargs(**kw))
    got, refv = r["coefficients"].to_numpy(), _ref(ref)
    assert (np.isnan(got) == np.isnan(refv)).all()
    ok = ~np.isnan(refv) & (np.abs(refv) < 1e6)
    assert np.allclose(got[ok], refv[ok], rtol=1e-6, atol=1e-7)
    assert abs(np.nanmean(got[-1]) - 1.0) < 0.05
    r = Frame(d).select(col("y").least_squares.rls(*names, half_life=500.0, mode="predictions").over("group"))["y"]
    ref = S.over(S.recursive_least_squares, d["group"], d["y"], *_oracle_cols(d, names), kwargs=S.RLSKwargs(half_life=500.0))
    got, refv = r.to_numpy(), _ref(ref)
    assert (np.isnan(got) == np.isnan(refv)).all()
    m = ~np.isnan(refv)
    _close(got[m], refv[m], rtol=1e-6, atol=1e-8)


# ----------------------------------------------------------------------------------- more than 64 coefficients (big.cuh)
@pytest.mark.parametrize("n_features", (100, 1000))
def test_fit_wide_hundreds_of_features(n_features):                        # tests/test_ols.py:272-313
    d = _make_data(10, n_features, scale=1.0e-4, seed=n_features)
    names, F = _xs(d), Frame(d)
    xs = _oracle_cols(d, names)
    r = F.select(col("y").least_squares.ols(*names, mode="coefficients"))["coefficients"]
    ref = _ref(S.least_squares(d["y"], *xs, mode="coefficients"))                                            # dgelsd min-norm
    _close(r.to_numpy()[0], ref, rtol=1e-6, atol=1e-6 * np.abs(ref).max())    # 1e-6 of the coefficient scale
    assert pls.get_engine(0).last_group_flags(1)[0] & 32
    F["coefficients"] = np.broadcast_to(r.to_numpy()[0], (10, n_features)).copy()
    p = F.select(col("coefficients").least_squares.predict(*names))["predictions"].to_numpy()
    assert np.corrcoef(p, d["y"])[0, 1] == pytest.approx(1.0, rel=1e-5, abs=1e-5)
    r = F.select(col("y").least_squares.ridge(*names, mode="coefficients", alpha=1.0e-5))["coefficients"]
    ref = _ref(S.least_squares(d["y"], *xs, mode="coefficients", kwargs=S.OLSKwargs(alpha=1.0e-5, l1_ratio=0.0)))
    # cond(X^T X + alpha I) = (n_features + alpha) / alpha ~ 1e7..1e8 here: a Cholesky answer (the reference's, the oracle's,
    # the device's) carries ~cond * eps * k of rounding, so the primal comparison is held at 1e-5, and both are also
    # checked against the well-conditioned dual form beta = X^T (X X^T + alpha I)^-1 y (cond ~ 1)
    xm = np.column_stack(xs)
    dual = xm.T @ np.linalg.solve(xm @ xm.T + 1.0e-5 * np.eye(10), d["y"])
    tol_r = 1e-6 if n_features <= 100 else 1e-5
    _close(r.to_numpy()[0], ref, rtol=tol_r, atol=1e-9)
    _close(r.to_numpy()[0], dual, rtol=tol_r, atol=1e-9)
    kw = dict(alpha=1.0e-6, tol=1.0e-8, max_iter=3_000)
    r = F.select(col("y").least_squares.lasso(*names, mode="coefficients", **kw))["coefficients"]
    ref = _ref(S.least_squares(d["y"], *xs, mode="coefficients", kwargs=S.OLSKwargs(l1_ratio=1.0, **kw)))
    _close(r.to_numpy()[0], ref, rtol=1e-6, atol=1e-9)
    p = F.select(col("y").least_squares.lasso(*names, **kw))["y"].to_numpy()
    assert np.corrcoef(p, d["y"])[0, 1] == pytest.approx(1.0, rel=1e-5, abs=1e-5)


@pytest.mark.parametrize("model,kw", [("ols", {}), ("ols", {"solve_method": "qr"}), ("ols", {"solve_method": "svd"}),
                                      ("ridge", {"alpha": 0.1}), ("ridge", {"alpha": 0.1, "solve_method": "lu"}),
                                      ("lasso", {"alpha": 1e-3}), ("elastic_net", {"alpha": 1e-3, "l1_ratio": 0.3, "solve_method": "cd_active_set"})])
def test_hundred_features_over_groups(model, kw):                          # k = 100 > 64, tall groups, nulls + weights
    d = _make_data(3000, 100, n_groups=3, sparsity=0.5, add_missing=False, seed=31)
    rng = np.random.default_rng(1)
    d["weights"] = rng.uniform(0.2, 5.0, size=3000)
    d["x7"] = (d["x7"], rng.random(3000) >= 0.05)
    d["y"] = (d["y"], rng.random(3000) >= 0.05)
    names, F = _xs(d), Frame(d)
    xs = _oracle_cols(d, names)
    okw = dict(kw)
    if model == "ridge":
        okw["l1_ratio"] = 0.0
    if model == "lasso":
        okw["l1_ratio"] = 1.0
    for mode in ("coefficients", "residuals"):
        e = getattr(col("y").least_squares, model)(*names, sample_weights="weights", add_intercept=True, mode=mode,
                                                   null_policy="drop", **kw).over("group")
        r = F.select(e)["coefficients" if mode == "coefficients" else "y"]
        if mode == "coefficients":
            keys, c, m = S.over(S.least_squares, d["group"], d["y"], *xs, sample_weights=d["weights"], add_intercept=True,
                                per_group=True, mode=mode, kwargs=S.OLSKwargs(null_policy="drop", **okw))
            _close(r.to_numpy(), c, rtol=1e-6, atol=1e-8)
        else:
            ref = S.over(S.least_squares, d["group"], d["y"], *xs, sample_weights=d["weights"], add_intercept=True, mode=mode,
                         kwargs=S.OLSKwargs(null_policy="drop", **okw))
            _close(r.to_numpy(), _ref(ref), rtol=1e-6, atol=1e-8)


def test_hundred_features_f32_device_frame():
    import torch
    d = _make_data(4000, 80, seed=33, dtype=np.float32)
    names = _xs(d)
    dev = {k: torch.as_tensor(v, device="cuda") for k, v in d.items()}
    r = Frame(dev).select(col("y").least_squares.ridge(*names, alpha=1e-2, mode="coefficients"))["coefficients"]
    ref = S.least_squares(d["y"], *_oracle_cols(d, names), mode="coefficients", kwargs=S.OLSKwargs(alpha=1e-2, l1_ratio=0.0))
    _close(r.to_numpy()[0], _ref(ref), rtol=1e-4, atol=1e-7)
    p = Frame(dev).select(col("y").least_squares.ridge(*names, alpha=1e-2))["y"]
    ref = S.least_squares(d["y"], *_oracle_cols(d, names), kwargs=S.OLSKwargs(alpha=1e-2, l1_ratio=0.0))
    _close(p.to_numpy(), _ref(ref), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("k,n_rows,n_groups", [(3, 40_000, 400), (8, 60_000, 97), (13, 50_000, 300), (16, 30_000, 7)])
@pytest.mark.parametrize("solve_method", [None, "lu"])
def test_fused_and_batched_solve_agree(k, n_rows, n_groups, solve_method, monkeypatch):
    """the streaming kernel's fused solve and batch_solve_kernel are the same ladder: identical to rounding,
    for every team size, with ragged groups (some empty, one long enough to be split)"""
    d = _make_data(n_rows, k, n_groups=n_groups, seed=k)
    d["group"][: n_rows // 3] = 1                      # one long group
    d["group"][d["group"] == 5] = 6                    # an empty key range
    names = _xs(d)
    keys, offsets, row_index, _ = pls.least_squares._group_plan([d["group"]])
    b = lambda: pls.Batch(pls.as_col(d["y"]), [pls.as_col(d[n]) for n in names], offsets=offsets, row_index=row_index)  # noqa: E731
    kw = OLSKwargs(alpha=0.05, l1_ratio=0.0, solve_method=solve_method).to_c()
    from polars_ols_b200 import _lib as L
    res = {}
    for fuse, multi in (("0", "1"), (str(1 << 40), "0"), (str(1 << 40), "1")):
        monkeypatch.setenv("B200OLS_FUSE_MIN_BYTES", fuse)
        monkeypatch.setenv("B200OLS_MULTI", multi)          # gram_multi_kernel (tiles of whole groups) on / off
        eng = pls.Engine(0)
        for team in (0, 1, 2, 4):
            eng.set_tuning(0, 0, team)
            res[(fuse, multi, team)] = eng.least_squares(b(), kw, L.COEFFICIENTS)[0].copy()
        eng.close()
    _, c, _ = S.over(S.least_squares, d["group"], d["y"], *_oracle_cols(d, names), per_group=True, mode="coefficients",
                     kwargs=S.OLSKwargs(alpha=0.05, l1_ratio=0.0, solve_method=solve_method))
    for key, v in res.items():
        _close(v, c, rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(v, res[("0", "1", 0)], rtol=1e-10, atol=1e-12)


# ----------------------------------------------------------------------------------- §8f: mode = "statistics"
STAT_KEYS = ("r2", "mae", "mse", "coefficients", "standard_errors", "t_values", "p_values")


def _stats_oracle_by_group(d, names, group=None, **kw):
    gid = np.zeros(len(d["y"] if not isinstance(d["y"], tuple) else d["y"][0]), dtype=int) if group is None else d[group]
    out = []
    for g in np.unique(gid):
        sel = gid == g
        take = lambda c: (c[0][sel], c[1][sel]) if isinstance(c, tuple) else c[sel]  # noqa: E731
        kk = {k: (take(v) if k == "sample_weights" and v is not None else v) for k, v in kw.items()}
        out.append(S.least_squares_statistics(take(d["y"]), *[take(d[n]) for n in names], **kk))
    return out


def _check_stats(res, refs, rtol=1e-6):
    got = res.to_struct()
    for i, ref in enumerate(refs):
        for k in STAT_KEYS:
            g, r = np.asarray(got[k][i]), np.asarray(ref[k])
            assert (np.isnan(g) == np.isnan(r)).all(), (k, i)
            m = ~np.isnan(r)
            # p-values below ~1e-12 are dominated by the reference's own 1 - (1 - ib) cancellation (multiples of 1.1e-16)
            assert np.allclose(g[m], r[m], rtol=rtol, atol=(3e-16 if k == "p_values" else 1e-12)), (k, i, g, r)


def test_statistics_readme_frame():                                         # README.md:143-165
    F = Frame({k: np.asarray(v, dtype=np.float64) for k, v in GOLD["frame"].items()})
    r = F.select(col("y").least_squares.ols("x1", "x2", mode="statistics", add_intercept=True))["statistics"]
    g, st = GOLD["statistics_ols_intercept"], r.to_struct()
    assert st["feature_names"] == g["feature_names"]
    for k in ("r2", "mae", "mse"):
        assert st[k][0] == pytest.approx(g[k], rel=g["printed_rel_tol"])
    for k in ("standard_errors", "t_values", "p_values"):
        assert np.allclose(st[k][0], g[k], rtol=g["printed_rel_tol"], atol=0)
    assert np.allclose(st["coefficients"][0], g["coefficients"], atol=GOLD["printed_abs_tol_coefficients"], rtol=0)


@pytest.mark.parametrize("model,kw", [("ols", {}), ("ridge", {"alpha": 10.0}), ("ridge", {"alpha": 0.3, "solve_method": "lu"}),
                                      ("lasso", {"alpha": 1e-3}), ("ols", {"solve_method": "svd"})])
@pytest.mark.parametrize("null_policy", ["ignore", "drop", "drop_y_zero_x", "zero"])
def test_statistics_over_groups(model, kw, null_policy):                    # tests/test_ols.py:998-1029, grouped + nulls + WLS
    d = _make_data(6000, 5, n_groups=7, scale=1.0, add_missing=null_policy != "ignore", seed=11)
    rng = np.random.default_rng(2)
    d["w"] = rng.uniform(0.2, 3.0, size=6000)
    names = _xs(d)
    e = getattr(col("y").least_squares, model)(*names, mode="statistics", add_intercept=True, sample_weights="w",
                                                null_policy=null_policy, **kw).over("group")
    r = Frame(d).select(e)["statistics"]
    okw = dict(kw)
    okw.setdefault("alpha", 0.0)
    if model == "ridge":
        okw["l1_ratio"] = 0.0
    if model == "lasso":
        okw["l1_ratio"] = 1.0
    refs = _stats_oracle_by_group(d, names, "group", sample_weights=d["w"], add_intercept=True,
                                  kwargs=S.OLSKwargs(null_policy=null_policy, **okw))
    _check_stats(r, refs)
    assert r.to_struct()["feature_names"] == names + ["const"]


@pytest.mark.parametrize("k,model,kw,null_policy", [(100, "ols", {}, "ignore"), (100, "ridge", {"alpha": 0.7}, "drop"),
                                                    (150, "lasso", {"alpha": 1e-3}, "ignore"), (300, "ridge", {"alpha": 5.0}, "zero"),
                                                    (80, "ridge", {"alpha": 0.3, "solve_method": "lu"}, "drop_y_zero_x")])
def test_statistics_above_64_coefficients(k, model, kw, null_policy):       # src/statistics.rs:76-156 has no bound on k
    """mode="statistics" behind the general path (big_stats.cuh): Cholesky of X^T X + lambda I in global memory, the
    diagonal / trace of the inverse from the columns of L^-1, residual metrics over the materialised fit rows."""
    n = 5 * k + 200
    d = _make_data(n, k, n_groups=3, scale=1.0, seed=k)
    rng = np.random.default_rng(4)
    if null_policy != "ignore":                      # nulls in three columns only: with 10 % per column over 100+ columns no row would survive "drop"
        for c_ in ("y", "x1", "x7"):
            d[c_] = (d[c_], rng.random(n) >= 0.1)
    d["w"] = rng.uniform(0.2, 3.0, size=n)
    names = _xs(d)
    e = getattr(col("y").least_squares, model)(*names, mode="statistics", add_intercept=True, sample_weights="w",
                                                null_policy=null_policy, **kw).over("group")
    r = Frame(d).select(e)["statistics"]
    okw = dict(kw)
    okw.setdefault("alpha", 0.0)
    if model == "ridge":
        okw["l1_ratio"] = 0.0
    if model == "lasso":
        okw["l1_ratio"] = 1.0
    refs = _stats_oracle_by_group(d, names, "group", sample_weights=d["w"], add_intercept=True,
                                  kwargs=S.OLSKwargs(null_policy=null_policy, **okw))
    _check_stats(r, refs)
    assert r.to_struct()["feature_names"] == names + ["const"]


def test_statistics_above_64_coefficients_wide_group():
    """more coefficients than rows: with lambda > 0 the inverse exists (df = n - trace(inv)); with lambda = 0 the Cholesky
    factorisation of the singular X^T X fails and the feature metrics are NaN (src/statistics.rs:101-111)"""
    rng = np.random.default_rng(9)
    n, k = 60, 90
    d = {f"x{i + 1}": rng.normal(size=n) for i in range(k)}
    d["y"] = rng.normal(size=n)
    names = [f"x{i + 1}" for i in range(k)]
    r = Frame(d).select(col("y").least_squares.ridge(*names, alpha=2.0, mode="statistics"))["statistics"]
    _check_stats(r, [S.least_squares_statistics(d["y"], *[d[c] for c in names], kwargs=S.OLSKwargs(alpha=2.0, l1_ratio=0.0))])
    r = Frame(d).select(col("y").least_squares.ols(*names, mode="statistics"))["statistics"].to_struct()
    ref = S.least_squares_statistics(d["y"], *[d[c] for c in names], kwargs=S.OLSKwargs(alpha=0.0))
    for key in ("standard_errors", "t_values", "p_values"):
        assert np.isnan(np.asarray(ref[key])).all() and np.isnan(np.asarray(r[key][0])).all(), key
    assert np.allclose(r["coefficients"][0], ref["coefficients"], rtol=1e-6, atol=1e-9)


def test_statistics_f32_device_frame_and_long_group():
    import torch
    rng = np.random.default_rng(3)
    n, k = 120_000, 12                                                       # split into segments, k > 8 (wide record path)
    x = rng.normal(size=(n, k)).astype(np.float32)
    y = (x @ rng.normal(size=k) + rng.normal(size=n)).astype(np.float32)
    g = np.repeat(np.arange(3), n // 3)
    dev = torch.device("cuda", 0)
    d = {f"x{j}": torch.as_tensor(x[:, j].copy(), device=dev) for j in range(k)}
    d["y"] = torch.as_tensor(y, device=dev)
    d["group"] = g
    names = [f"x{j}" for j in range(k)]
    r = Frame(d).select(col("y").least_squares.ridge(*names, alpha=2.0, mode="statistics").over("group"))["statistics"]
    dn = {f"x{j}": x[:, j] for j in range(k)}
    dn["y"], dn["group"] = y, g
    refs = _stats_oracle_by_group(dn, names, "group", kwargs=S.OLSKwargs(alpha=2.0, l1_ratio=0.0))
    _check_stats(r, refs, rtol=1e-4)                                         # f32 inputs


def test_statistics_failure_modes():
    d = _make_data(50, 3, seed=4)
    F = Frame(d)
    # collinear features, alpha = 0: the Cholesky of X^T X fails -> NaN feature metrics, residual metrics still reported
    d2 = dict(d)
    d2["x4"] = d["x1"] * 2.0
    r = Frame(d2).select(col("y").least_squares.ols("x1", "x2", "x3", "x4", mode="statistics", solve_method="svd"))["statistics"].to_struct()
    ref = S.least_squares_statistics(d2["y"], d2["x1"], d2["x2"], d2["x3"], d2["x4"], kwargs=S.OLSKwargs(alpha=0.0, solve_method="svd"))
    if np.isnan(ref["standard_errors"]).all():                              # (numpy's Cholesky may squeak through on rounding)
        assert np.isnan(r["standard_errors"][0]).all() and np.isnan(r["p_values"][0]).all()
    assert r["r2"][0] == pytest.approx(ref["r2"], rel=1e-6)
    # df <= 0: the reference asserts (src/statistics.rs:131-134)
    tiny = Frame({k: v[:3] for k, v in d.items()})
    with pytest.raises(pls.B200OLSError, match="Degrees of freedom"):
        tiny.select(col("y").least_squares.ols("x1", "x2", "x3", mode="statistics", solve_method="chol"))
    # alpha = None: `kwargs.alpha.unwrap()` panics in the reference
    with pytest.raises(pls.B200OLSError, match="alpha"):
        F.select(pls.compute_least_squares("y", "x1", mode="statistics", ols_kwargs=OLSKwargs(alpha=None)))


# ----------------------------------------------------------------------------------- §8f: multi-target
def _multi_oracle(d, tnames, names, group, mode, kw, **extra):
    n = len(d[group])
    out_v, out_m = np.full((len(tnames), n), np.nan), np.zeros((len(tnames), n), dtype=bool)
    take = lambda c, sel: (c[0][sel], c[1][sel]) if isinstance(c, tuple) else c[sel]  # noqa: E731
    for g in np.unique(d[group]):
        sel = d[group] == g
        ex = {k: (take(v, sel) if k == "sample_weights" else v) for k, v in extra.items()}
        v, m = S.multi_target_least_squares([take(d[t], sel) for t in tnames], *[take(d[x], sel) for x in names], mode=mode,
                                            kwargs=kw, **ex)
        out_v[:, sel], out_m[:, sel] = v.T, m.T
    return out_v, out_m


@pytest.mark.parametrize("alpha,mode,null_policy", [(0.0, "residuals", "ignore"), (0.0, "residuals", "drop"),
                                                    (0.0001, "residuals", "drop_y_zero_x"), (0.01, "residuals", "drop_zero"),
                                                    (0.5, "predictions", "zero"), (0.0, "predictions", "drop")])
def test_multi_target_regression(alpha, mode, null_policy):                # tests/test_ols.py:76-127
    d = _make_data(10_000, 3, n_groups=3, seed=21)
    if null_policy not in ("zero", "ignore"):
        d["x1"] = (d["x1"], np.random.default_rng(5).random(10_000) >= 0.1)    # missing_columns=("x1",)
    x1, x2, x3 = (d["x1"][0] if isinstance(d["x1"], tuple) else d["x1"]), d["x2"], d["x3"]
    m1 = d["x1"][1] if isinstance(d["x1"], tuple) else None
    ys = {"y1": x1 + x2 + x3, "y2": x1 - x2 + x3, "y3": -x1 + x2 - x3}
    ys = {k: (v, m1) if m1 is not None else v for k, v in ys.items()}           # struct fields inherit x1's nulls
    F = Frame(d)
    F["ys"] = ys
    names = _xs(d)
    kw = OLSKwargs(null_policy=null_policy, solve_method="svd", alpha=alpha)
    r = F.select(pls.compute_multi_target_least_squares("ys", *names, mode=mode, ols_kwargs=kw).over("group").alias(mode))[mode]
    assert r.fields == ["y1", "y2", "y3"]
    dd = dict(d)
    dd.update(ys)
    okw = S.OLSKwargs(null_policy=null_policy, solve_method="svd", alpha=alpha)
    ref_v, ref_m = _multi_oracle(dd, list(ys), names, "group", mode, okw)
    _close(r.to_numpy(), np.where(ref_m, ref_v, np.nan), rtol=1e-6, atol=1e-8)
    # the reference's own assertion: equal to independent single-target regressions
    for j, t in enumerate(ys):
        F1 = Frame(dd)
        r1 = F1.select(col(t).least_squares.least_squares(*names, mode=mode, null_policy=null_policy, solve_method="svd",
                                                          alpha=alpha).over("group"))[t]
        a, b = r.to_numpy()[j], r1.to_numpy()
        ok = ~(np.isnan(a) | np.isnan(b))
        assert np.allclose(a[ok], b[ok], atol=1e-8, rtol=1e-6)


def test_multi_target_weights_intercept_rank_deficient_and_wide():
    rng = np.random.default_rng(9)
    n = 900
    d = _make_data(n, 4, n_groups=4, seed=8)
    d["x5"] = d["x1"] - 2.0 * d["x2"]                                           # exactly collinear -> SVD kernel (min-norm)
    d["group"][:3] = 99                                                         # a 3-row group: n <= k -> SVD kernel
    d["w"] = rng.uniform(0.3, 2.0, size=n)
    names = _xs(d)
    ys = {"a": d["y"], "b": d["x1"] * 0.5 - d["x3"] + 0.1 * rng.normal(size=n)}
    F = Frame(d)
    F["ys"] = ys
    dd = dict(d)
    dd.update(ys)
    for alpha in (0.0, 0.25):
        kw = OLSKwargs(solve_method="svd", alpha=alpha)
        r = F.select(col("ys").least_squares.multi_target_ols(*names, sample_weights="w", add_intercept=True, alpha=alpha,
                                                               solve_method="svd").over("group"))["predictions"]
        ref_v, ref_m = _multi_oracle(dd, ["a", "b"], names, "group", "predictions", S.OLSKwargs(solve_method="svd", alpha=alpha),
                                     sample_weights=d["w"], add_intercept=True)
        _close(r.to_numpy(), np.where(ref_m, ref_v, np.nan), rtol=1e-6, atol=1e-7)
        fl = pls.get_engine(0).last_group_flags(5)
        assert (fl & 32).any()                                                  # some groups went through the Jacobi SVD kernel
    with pytest.raises(AssertionError):
        pls.compute_multi_target_least_squares("ys", *names, ols_kwargs=OLSKwargs(l1_ratio=0.5, alpha=0.1))
    with pytest.raises(NotImplementedError):
        pls.compute_multi_target_least_squares("ys", *names, mode="coefficients")


# ----------------------------------------------------------------------------------- fused Gram -> solve -> predict kernel
@pytest.mark.parametrize("dtype,k,n_groups,model,kw,weights,intercept", [
    (np.float64, 8, 40, "ridge", {"alpha": 1e-3}, False, False),                 # C2 shape (predictions / residuals)
    (np.float64, 3, 7, "ols", {}, False, True),                                  # default OLS: QR guard + restricted re-predict
    (np.float64, 13, 25, "ridge", {"alpha": 0.1, "solve_method": "lu"}, True, True),
    (np.float32, 16, 30, "ridge", {"alpha": 0.5}, True, False),                  # C3 dtype, KB = 2
    (np.float64, 5, 300, "ols", {"solve_method": "chol"}, True, False),          # short ragged groups
])
@pytest.mark.parametrize("mode", ["predictions", "residuals"])
def test_fused_predict_kernel(dtype, k, n_groups, model, kw, weights, intercept, mode, monkeypatch):
    """gram_pred.cuh (one pass) against the two-pass route (B200OLS_PRED=0) and the oracle.  B200OLS_FUSE_MIN_BYTES=0
    makes the fused route eligible for these small groups; stage counts 2 and 3+ are both exercised."""
    from polars_ols_b200 import _lib as L
    rng = np.random.default_rng(k)
    sizes = rng.integers(8, 200, size=n_groups) if n_groups >= 300 else rng.integers(200, 1000 if k <= 8 else 400, size=n_groups)
    sizes[1] = 0                                                                # an empty group
    n = int(sizes.sum())                                                        # no group above 1024 rows: the plan never splits one
    d = _make_data(n, k, seed=100 + k, dtype=dtype)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    gid = np.repeat(np.arange(n_groups), sizes)
    if model == "ols" and not kw:                                               # one nearly collinear group -> flagged, re-solved by QR
        sl = slice(offsets[2], offsets[3])
        d["x2"][sl] = d["x1"][sl] * (1.0 + 1e-9 * rng.normal(size=sizes[2]))
    names = _xs(d)
    w = rng.uniform(0.2, 3.0, size=n).astype(dtype) if weights else None
    mk = lambda: pls.Batch(pls.as_col(d["y"]), [pls.as_col(d[nm]) for nm in names], None if w is None else pls.as_col(w),  # noqa: E731
                           add_intercept=intercept, offsets=offsets)
    okw = dict(kw)
    if model == "ridge":
        okw["l1_ratio"] = 0.0
    ckw = OLSKwargs(**okw).to_c()
    monkeypatch.setenv("B200OLS_FUSE_MIN_BYTES", "0")
    res = {}
    for pred, stages in (("1", 0), ("1", 2), ("0", 0)):
        monkeypatch.setenv("B200OLS_PRED", pred)
        eng = pls.Engine(0)
        eng.set_tuning(0, stages, 0)
        n0 = eng.launch_count
        v, m = eng.least_squares(mk(), ckw, L.MODE[mode])
        res[(pred, stages)] = (v.copy(), m.copy(), eng.launch_count - n0)
        eng.close()
    ref = S.over(S.least_squares, gid, d["y"], *_oracle_cols(d, names), sample_weights=w, add_intercept=intercept, mode=mode,
                 kwargs=S.OLSKwargs(**okw))
    tol = dict(rtol=1e-6, atol=1e-8) if dtype == np.float64 else dict(rtol=1e-4, atol=1e-4)
    if model == "ols" and not kw:
        tol = dict(rtol=1e-5, atol=1e-6)                                        # the collinear group: cond ~ 1e9
    for key, (v, m, launches) in res.items():
        assert m.all()
        _close(v, _ref(ref), **tol)
    if model == "ols" and not kw:                                               # QR guard: the flagged groups are re-predicted
        assert res[("1", 0)][2] <= res[("0", 0)][2]
    else:
        assert res[("1", 0)][2] < res[("0", 0)][2]                              # fewer launches: no separate predict pass
    np.testing.assert_allclose(res[("1", 0)][0], res[("0", 0)][0], rtol=1e-9 if dtype == np.float64 else 1e-5, atol=1e-9)
    np.testing.assert_array_equal(res[("1", 0)][0], res[("1", 2)][0])           # stage count does not change the arithmetic


# ----------------------------------------------------------------------------------- formula API
def test_from_formula_matches_explicit_expressions():                      # tests/test_ols.py:362-371,437-449,456-470
    d = _make_data(4000, 4, n_groups=5, seed=31)
    F = Frame(d)
    a = F.select(col("y").least_squares.from_formula("x1 + x2 -1", mode="coefficients").over("group"))["coefficients"]
    b = F.select(col("y").least_squares.ols("x1", "x2", mode="coefficients").over("group"))["coefficients"]
    np.testing.assert_array_equal(a.to_numpy(), b.to_numpy())
    # intercept + interaction term, against the oracle on the explicitly multiplied column
    e = pls.compute_least_squares_from_formula("y ~ x1 + x3:x4", mode="predictions")
    got = F.select(e.over("group"))["y"].to_numpy()
    ref = S.over(S.least_squares, d["group"], d["y"], d["x1"], d["x3"] * d["x4"], add_intercept=True, mode="predictions",
                 kwargs=S.OLSKwargs())
    _close(got, _ref(ref), rtol=1e-6, atol=1e-8)
    # kwargs dispatch: window_size -> rolling, half_life -> rls
    r1 = F.select(col("y").least_squares.from_formula("x1 + x2 -1", window_size=50, mode="coefficients").over("group"))["coefficients"]
    r2 = F.select(col("y").least_squares.rolling_ols("x1", "x2", window_size=50, mode="coefficients").over("group"))["coefficients"]
    np.testing.assert_array_equal(r1.to_numpy(), r2.to_numpy())
    l1 = F.select(pls.compute_least_squares_from_formula("y ~ x1 + x2 -1", half_life=20.0).over("group"))["y"]
    l2 = F.select(col("y").least_squares.rls("x1", "x2", half_life=20.0).over("group"))["y"]
    np.testing.assert_array_equal(l1.to_numpy(), l2.to_numpy())
    # predict_from_formula
    coef = F.select(col("y").least_squares.from_formula("x1 + x2", mode="coefficients"))["coefficients"].to_numpy()[0]
    Fp = Frame({"coefficients": np.broadcast_to(coef, (4000, 3)).copy(), "x1": d["x1"], "x2": d["x2"]})
    p = Fp.select(col("coefficients").least_squares.predict_from_formula("x1 + x2", name="p"))["p"].to_numpy()
    _close(p, d["x1"] * coef[0] + d["x2"] * coef[1] + coef[2], rtol=1e-9, atol=1e-12)


# ----------------------------------------------------------------------------------- C3 shape: medium groups, 4 per tile
@pytest.mark.parametrize("mode", ["predictions", "coefficients"])
def test_c3_shaped_medium_groups_weighted_elastic_net_f32(mode):
    """BASELINE config 3 in small: f32 columns, sample weights, elastic net, ragged groups of ~256 rows x 16 features —
    the shape that takes gram_multi's three-stage tiles of four groups (sqrt(w) converted in place in the stage)."""
    rng = np.random.default_rng(3)
    G, k = 420, 16
    sizes = rng.integers(150, 257, size=G)
    sizes[7] = 0
    n = int(sizes.sum())
    d = _make_data(n, k, seed=77, dtype=np.float32)
    names = _xs(d)
    w = rng.uniform(0.05, 1.0, size=n).astype(np.float32)
    gid = np.repeat(np.arange(G), sizes)
    F = Frame({**d, "w": w, "group": gid})
    e = col("y").least_squares.elastic_net(*names, alpha=1e-3, l1_ratio=0.5, sample_weights="w", mode=mode).over("group")
    r = F.select(e)["coefficients" if mode == "coefficients" else "y"]
    if mode == "coefficients":
        _, c, _ = S.over(S.least_squares, gid, d["y"], *_oracle_cols(d, names), sample_weights=w, per_group=True,
                         mode="coefficients", kwargs=S.OLSKwargs(alpha=1e-3, l1_ratio=0.5))
        got = r.to_numpy()
        assert got.shape == (G - 1, k)                                          # the empty group has no key
        _close(got, c, rtol=1e-4, atol=1e-5)
    else:
        ref = S.over(S.least_squares, gid, d["y"], *_oracle_cols(d, names), sample_weights=w, kwargs=S.OLSKwargs(alpha=1e-3, l1_ratio=0.5))
        _close(r.to_numpy(), _ref(ref), rtol=1e-4, atol=1e-4)
