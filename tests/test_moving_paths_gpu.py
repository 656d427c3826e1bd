"""rls / rolling_ols execution paths added in round 2, each against the sequential oracle at 1e-6 (f64) from row 0:
  * staged thread-per-chunk kernels for null-free frames, k <= 8 (moving_fast.cuh)
  * block-per-chunk kernels for 9 <= k <= 64 (moving_wide.cuh), incl. the reference's Woodbury regime k > 60
    (src/least_squares.rs:737-787,863)
  * warm-up longer than the window in both null branches (src/least_squares.rs:881-921, 987-1029)"""
import numpy as np
import pytest

import polars_ols_b200 as pls
from polars_ols_b200 import Frame, col
from oracle import semantics as S

pytestmark = pytest.mark.gpu


def _data(n, k, n_groups=None, missing=0.0, seed=0, dtype=np.float64, scale=0.1):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, k))
    y = x.sum(1) + scale * rng.normal(size=n)
    d = {f"x{i + 1}": np.ascontiguousarray(x[:, i]).astype(dtype) for i in range(k)}
    d["y"] = y.astype(dtype)
    if n_groups:
        d["group"] = rng.integers(n_groups, size=n)
    if missing:
        for c in [c for c in d if c != "group"]:
            m = rng.random(n) < missing
            d[c] = (d[c], ~m)
    return d


def _plain(v):
    return v[0] if isinstance(v, tuple) else v


def _ocols(d, names):
    return [d[n] for n in names]          # the oracle takes the same (values, valid) pairs as the Frame


def _ref(values_mask):
    v, m = values_mask
    return np.where(m, v, np.nan) if m is not None else v


def _check(got, ref, rtol=1e-6, atol=1e-8, big=1e6):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape
    assert (np.isnan(got) == np.isnan(ref)).all(), "null / NaN pattern differs"
    ok = ~np.isnan(ref) & (np.abs(ref) < big)
    err = np.abs(got[ok] - ref[ok]) / (atol / rtol + np.abs(ref[ok]))
    assert err.max() <= rtol, f"max rel err {err.max():.3e}"


def _oracle(fn, d, names, mode, kwargs, over=None, weights=None):
    y = _ocols(d, ["y"])[0]
    xs = _ocols(d, names)
    kw = dict(mode=mode, kwargs=kwargs)
    if weights is not None:
        kw["sample_weights"] = weights
    if over is not None:
        return _ref(S.over(fn, over, y, *xs, **kw))
    return _ref(fn(y, *xs, **kw))


# ------------------------------------------------------------------------------------------ staged fast path (k <= 8)
@pytest.mark.parametrize("mode", ["coefficients", "predictions", "residuals"])
@pytest.mark.parametrize("k,window,min_periods,alpha,n", [(6, 252, 6, None, 40_000), (3, 21, None, None, 5_000),
                                                          (8, 64, 10, 0.01, 20_000), (1, 5, 1, None, 3_000), (2, 7, 7, None, 4_001)])
def test_rolling_fast_path_null_free(mode, k, window, min_periods, alpha, n):
    d = _data(n, k, seed=k + window)
    names = [f"x{i + 1}" for i in range(k)]
    kw = dict(window_size=window, min_periods=min_periods, alpha=alpha)
    for policy in ("drop", "drop_window"):
        r = Frame(d).select(col("y").least_squares.rolling_ols(*names, mode=mode, null_policy=policy, **kw))
        got = r["coefficients" if mode == "coefficients" else "y"].to_numpy()
        _check(got, _oracle(S.rolling_least_squares, d, names, mode, S.RollingKwargs(null_policy=policy, **kw)))


@pytest.mark.parametrize("mode", ["coefficients", "predictions", "residuals"])
@pytest.mark.parametrize("k,window,min_periods,alpha,n,groups", [(6, 252, 6, None, 60_000, None), (3, 64, None, None, 9_001, 5),
                                                                 (8, 100, 10, 0.01, 20_003, 3), (2, 65, 65, None, 7_000, None)])
def test_rolling_window_chunk_kernel(monkeypatch, mode, k, window, min_periods, alpha, n, groups):
    """rolling_nbr_kernel (chunks of exactly `window` rows, lag rows read from the neighbouring thread's staging slot):
    forced on small frames through the test hook; odd series starts / lengths exercise the shifted unit boundaries"""
    monkeypatch.setenv("B200OLS_MOVING_NBR_MIN_CHUNKS", "1")
    d = _data(n, k, n_groups=groups, seed=k * window)
    names = [f"x{i + 1}" for i in range(k)]
    kw = dict(window_size=window, min_periods=min_periods, alpha=alpha)
    for dt, rtol, atol in ((np.float64, 1e-6, 1e-8), (np.float32, 1e-4, 1e-6)):
        dd = {c: (v.astype(dt) if c != "group" else v) for c, v in d.items()}
        d64 = {c: (v.astype(np.float64) if c != "group" else v) for c, v in dd.items()}
        e = col("y").least_squares.rolling_ols(*names, mode=mode, **kw)
        r = Frame(dd).select(e.over("group") if groups else e)
        got = r["coefficients" if mode == "coefficients" else "y"].to_numpy()
        ref = _oracle(S.rolling_least_squares, d64, names, mode, S.RollingKwargs(null_policy="drop", **kw), over=d["group"] if groups else None)
        _check(got, ref, rtol=rtol, atol=atol)


@pytest.mark.parametrize("mode", ["coefficients", "predictions"])
def test_fast_path_over_groups_weights_and_f32(mode):
    d = _data(30_000, 4, n_groups=7, seed=5)
    rng = np.random.default_rng(1)
    d["w"] = rng.uniform(0.2, 5.0, size=30_000)
    names = ["x1", "x2", "x3", "x4"]
    key = "coefficients" if mode == "coefficients" else "y"
    r = Frame(d).select(col("y").least_squares.rolling_ols(*names, window_size=100, min_periods=8, sample_weights="w", mode=mode).over("group"))
    _check(r[key].to_numpy(), _oracle(S.rolling_least_squares, d, names, mode, S.RollingKwargs(window_size=100, min_periods=8, null_policy="drop"),
                                      over=d["group"], weights=d["w"]))
    r = Frame(d).select(col("y").least_squares.rls(*names, half_life=60.0, sample_weights="w", add_intercept=True, mode=mode).over("group"))
    ref = S.over(S.recursive_least_squares, d["group"], d["y"], *[d[n] for n in names], sample_weights=d["w"], add_intercept=True,
                 mode=mode, kwargs=S.RLSKwargs(half_life=60.0))
    _check(r[key].to_numpy(), _ref(ref))
    # f32 columns: the reference computes in f64 on the f32-rounded inputs (src/expressions.rs:33,47,80) -> 1e-4
    d32 = {k_: (v.astype(np.float32) if k_ != "group" else v) for k_, v in d.items()}
    r = Frame(d32).select(col("y").least_squares.rolling_ols(*names, window_size=100, min_periods=8, mode=mode).over("group"))
    d64 = {k_: (v.astype(np.float64) if k_ != "group" else v) for k_, v in d32.items()}
    _check(r[key].to_numpy(), _oracle(S.rolling_least_squares, d64, names, mode, S.RollingKwargs(window_size=100, min_periods=8, null_policy="drop"),
                                      over=d["group"]), rtol=1e-4, atol=1e-6)
    r = Frame(d32).select(col("y").least_squares.rls(*names, half_life=60.0, mode=mode).over("group"))
    _check(r[key].to_numpy(), _oracle(S.recursive_least_squares, d64, names, mode, S.RLSKwargs(half_life=60.0), over=d["group"]),
           rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("half_life,p0,mean", [(None, 1e6, None), (252.0, 10.0, None), (21.0, 10.0, [-1.0] * 5), (None, 10.0, None)])
@pytest.mark.parametrize("mode", ["coefficients", "predictions", "residuals"])
def test_rls_fast_path_from_row_zero(half_life, p0, mean, mode):
    d = _data(50_000, 5, seed=11)
    names = [f"x{i + 1}" for i in range(5)]
    kw = dict(half_life=half_life, initial_state_covariance=p0, initial_state_mean=mean)
    r = Frame(d).select(col("y").least_squares.rls(*names, mode=mode, **kw))
    got = r["coefficients" if mode == "coefficients" else "y"].to_numpy()
    _check(got, _oracle(S.recursive_least_squares, d, names, mode, S.RLSKwargs(**kw)))


# ------------------------------------------------------------------------------------------ 9 <= k <= 64 (block per chunk)
@pytest.mark.parametrize("k,missing", [(9, 0.0), (12, 0.1), (20, 0.05), (33, 0.0), (64, 0.002)])
@pytest.mark.parametrize("policy", ["drop", "drop_window"])
def test_rolling_wide(k, missing, policy):
    n = 2_500 if k <= 33 else 1_200
    d = _data(n, k, missing=missing, seed=k)
    names = [f"x{i + 1}" for i in range(k)]
    kw = dict(window_size=max(3 * k, 100), min_periods=k + 3, null_policy=policy)
    for mode in ("coefficients", "residuals"):
        r = Frame(d).select(col("y").least_squares.rolling_ols(*names, mode=mode, **kw))
        got = r["coefficients" if mode == "coefficients" else "y"].to_numpy()
        _check(got, _oracle(S.rolling_least_squares, d, names, mode, S.RollingKwargs(**kw)), atol=1e-7)


@pytest.mark.parametrize("use_woodbury", [None, True, False])
def test_rolling_woodbury_regime(use_woodbury):
    """k > 60: the reference switches to a Woodbury update of (X^T X)^-1 (src/least_squares.rs:863); the device solves
    S beta = v directly — same coefficients."""
    k, n = 61, 1_000
    d = _data(n, k, seed=61)
    names = [f"x{i + 1}" for i in range(k)]
    kw = dict(window_size=400, min_periods=150, use_woodbury=use_woodbury, alpha=0.01, null_policy="drop")
    r = Frame(d).select(col("y").least_squares.rolling_ols(*names, mode="coefficients", **kw))["coefficients"].to_numpy()
    _check(r, _oracle(S.rolling_least_squares, d, names, "coefficients", S.RollingKwargs(**kw)), atol=1e-7)


@pytest.mark.parametrize("k,missing,half_life,p0", [(9, 0.0, 100.0, 10.0), (16, 0.1, None, 1e4), (24, 0.05, 500.0, 10.0), (64, 0.0, 300.0, 1.0)])
def test_rls_wide(k, missing, half_life, p0):
    n = 3_000 if k <= 24 else 1_500
    d = _data(n, k, n_groups=2, missing=missing, seed=100 + k)
    names = [f"x{i + 1}" for i in range(k)]
    kw = dict(half_life=half_life, initial_state_covariance=p0)
    for mode in ("coefficients", "predictions"):
        r = Frame(d).select(col("y").least_squares.rls(*names, mode=mode, **kw).over("group"))
        got = r["coefficients" if mode == "coefficients" else "y"].to_numpy()
        _check(got, _oracle(S.recursive_least_squares, d, names, mode, S.RLSKwargs(**kw), over=d["group"]))


def test_more_than_64_moving_coefficients_is_refused_loudly():
    d = _data(300, 65, seed=3)
    names = [f"x{i + 1}" for i in range(65)]
    with pytest.raises(pls.B200OLSError) as ei:
        Frame(d).select(col("y").least_squares.rolling_ols(*names, window_size=100))
    assert ei.value.code == -2


# ------------------------------------------------------------------------------------------ warm-up longer than the window
@pytest.mark.parametrize("policy", ["drop", "drop_zero", "drop_window", "zero"])
@pytest.mark.parametrize("k,window,min_periods,missing", [(2, 10, 25, 0.0), (3, 8, 30, 0.2), (10, 12, 40, 0.1)])
def test_rolling_min_periods_longer_than_window(policy, k, window, min_periods, missing):
    d = _data(3_000, k, n_groups=3, missing=missing, seed=window)
    names = [f"x{i + 1}" for i in range(k)]
    kw = dict(window_size=window, min_periods=min_periods, null_policy=policy)
    r = Frame(d).select(col("y").least_squares.rolling_ols(*names, mode="coefficients", **kw).over("group"))["coefficients"].to_numpy()
    _check(r, _oracle(S.rolling_least_squares, d, names, "coefficients", S.RollingKwargs(**kw), over=d["group"]), atol=1e-7)
