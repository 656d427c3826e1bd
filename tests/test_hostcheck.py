"""Host logic tests: the device algorithm templates (solvers.cuh, moving_core.cuh), compiled for the CPU by
tests/hostcheck, against the oracle.  This validates the chunked (time-parallel) rolling / RLS
restarts and the Gram-form coordinate descent before any GPU time is spent.  The CUDA kernels wrap these
same templates; the `-m gpu` tests check the kernels themselves through the C ABI."""
import numpy as np
import pytest

import hostcheck
from oracle import semantics as S


def _data(n, k, seed=0, null_frac=0.0):
    rng = np.random.default_rng(seed)
    x = np.ascontiguousarray(rng.normal(size=(n, k)))
    y = x @ (1.0 + 0.25 * rng.normal(size=k)) + 0.1 * rng.normal(size=n)
    valid = (rng.random(n) >= null_frac) if null_frac > 0 else np.ones(n, bool)
    x = np.where(valid[:, None], x, 0.0)   # the reference zero-fills invalid rows (NullPolicy::Zero)
    y = np.where(valid, y, 0.0)
    return np.ascontiguousarray(y), np.ascontiguousarray(x), valid


def _rel(a, b):
    m = ~(np.isnan(a) & np.isnan(b))
    assert (np.isnan(a) == np.isnan(b)).all()
    return np.max(np.abs(a[m] - b[m]) / (1e-12 + np.abs(b[m]))) if m.any() else 0.0


def test_normal_equations_chol_and_lu_fallback():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(200, 6)); y = rng.normal(size=200)
    G = np.ascontiguousarray(x.T @ x + 1e-3 * np.eye(6)); c = np.ascontiguousarray(x.T @ y)
    ref = np.linalg.solve(G, c)
    for use_lu in (0, 1):
        g, cc = G.copy(), c.copy()
        fl = hostcheck.lib().hc_normal_equations(g.ctypes.data, 6, cc.ctypes.data, use_lu, 1e7)
        assert fl == 0 and np.allclose(cc, ref, rtol=1e-12)
    # indefinite matrix: Cholesky fails -> LU (flag bit 0)
    A = np.ascontiguousarray(np.array([[1.0, 2.0], [2.0, 1.0]])); b = np.array([1.0, 0.0])
    fl = hostcheck.lib().hc_normal_equations(A.copy().ctypes.data, 2, b.ctypes.data, 0, 1e7)
    assert fl & 1 and np.allclose(b, np.linalg.solve(A, [1.0, 0.0]))


@pytest.mark.parametrize("k,alpha,l1,positive,active", [(8, 1e-3, 0.5, 0, 0), (16, 1e-3, 0.5, 0, 1), (64, 1e-4, 1.0, 0, 0),
                                                         (5, 1e-2, 0.3, 1, 0), (10, 1e-3, 1.0, 1, 1)])
def test_cd_gram_matches_residual_form_oracle(k, alpha, l1, positive, active):
    y, x, _ = _data(1500, k, seed=k)
    if k >= 10:
        y = x[:, : k // 2].sum(1) + 0.1 * np.random.default_rng(1).normal(size=len(y))
    w_ref, sweeps_ref = S.solve_elastic_net(y, x, alpha, l1, 1000, 1e-5, bool(positive),
                                            "cd_active_set" if active else "cd", return_sweeps=True)
    G = np.ascontiguousarray(x.T @ x); c = np.ascontiguousarray(x.T @ y); w = np.empty(k)
    sweeps = hostcheck.lib().hc_cd_gram(G.ctypes.data, k, c.ctypes.data, alpha * len(y), l1, 1000, 1e-5, positive, active,
                                        w.ctypes.data)
    assert sweeps == sweeps_ref
    assert np.max(np.abs(w - w_ref)) < 1e-10


@pytest.mark.parametrize("fixed", [0, 1])
@pytest.mark.parametrize("k,window,min_periods,chunk,null_frac,alpha", [
    (2, 21, None, 64, 0.0, 0.0), (3, 50, 10, 37, 0.1, 0.0), (6, 252, None, 128, 0.0, 0.0), (4, 30, 4, 1000000, 0.2, 0.0),
    (2, 5, 2, 7, 0.3, 0.01), (8, 40, 8, 64, 0.05, 0.0), (1, 10, 1, 16, 0.5, 0.0), (3, 15, 3, 8, 0.6, 0.0),
    # warm-up longer than the window: the reference never subtracts the warm-up rows its deque did not take
    # (src/least_squares.rs:917-919) / the rows a fixed window has already passed (:987-1029)
    (2, 10, 25, 16, 0.0, 0.0), (3, 8, 30, 7, 0.2, 0.0), (2, 4, 9, 1000000, 0.3, 0.001), (3, 12, 40, 64, 0.1, 0.0)])
def test_rolling_chunks_match_sequential_oracle(fixed, k, window, min_periods, chunk, null_frac, alpha):
    n = 700
    y, x, valid = _data(n, k, seed=window + k, null_frac=null_frac)
    policy = "drop_window" if fixed else "drop"
    ref = S.solve_rolling_ols(y, x, window, min_periods, False, alpha or None, valid, policy)
    mp = min(k, window) if min_periods is None else min_periods
    out = np.empty((n, k))
    v8 = np.ascontiguousarray(valid, dtype=np.uint8)
    hostcheck.lib().hc_rolling(y.ctypes.data, x.ctypes.data, v8.ctypes.data, n, k, window, mp, alpha, fixed, chunk,
                               out.ctypes.data)
    assert (np.isnan(out) == np.isnan(ref)).all()
    m = ~np.isnan(ref)
    # windows that hold fewer valid rows than features are singular: the reference's own answer there is
    # rounding noise, so compare only where the sequential result is well determined
    ok = m & (np.abs(ref) < 1e6)
    assert np.allclose(out[ok], ref[ok], rtol=1e-6, atol=1e-7)


def test_rolling_insufficient_data_all_nan():
    y, x, valid = _data(5, 2)
    out = np.zeros((5, 2))
    hostcheck.lib().hc_rolling(y.ctypes.data, x.ctypes.data, None, 5, 2, 100, 6, 0.0, 0, 64, out.ctypes.data)
    ref = S.solve_rolling_ols(y, x, 100, 6, False, None, valid, "drop")
    assert np.isnan(out).all() and np.isnan(ref).all()


@pytest.mark.parametrize("k,half_life,p0,chunk,null_frac,mean", [
    (2, None, 10.0, 64, 0.0, None), (3, 252.0, 10.0, 50, 0.1, None), (6, 252.0, 10.0, 128, 0.0, None),
    (2, None, 1e6, 33, 0.2, None), (4, 20.0, 0.01, 16, 0.0, [0.25, 0.25, 0.25, 0.25]), (8, 100.0, 10.0, 64, 0.05, None)])
def test_rls_chunks_match_sequential_oracle(k, half_life, p0, chunk, null_frac, mean):
    n = 900
    y, x, valid = _data(n, k, seed=k + 3, null_frac=null_frac)
    ref = S.solve_recursive_least_squares(y, x, half_life, p0, mean, valid)
    lam = 1.0 if half_life is None else float(np.exp(np.log(0.5) / half_life))
    out = np.empty((n, k))
    v8 = np.ascontiguousarray(valid, dtype=np.uint8)
    mean_arr = None if mean is None else np.asarray(mean, dtype=np.float64)
    hostcheck.lib().hc_rls(y.ctypes.data, x.ctypes.data, v8.ctypes.data, n, k, lam, p0,
                           None if mean_arr is None else mean_arr.ctypes.data, chunk, out.ctypes.data)
    # 1e-6 from row 0 (north_star: no warm-up exemption): the prior-dominated head of a series runs the reference's
    # literal operation sequence (rls_update_exact), which reproduces the sequential oracle bit for bit there
    assert np.allclose(out, ref, rtol=1e-6, atol=1e-8)
    head = np.flatnonzero(valid)[:min(chunk, 64)]
    head = head[head < chunk]
    assert np.array_equal(out[head], ref[head])


# ------------------------------------------------------------------------------- mode = "statistics" building blocks
def test_students_t_p_value_matches_scipy():
    """stats_math.cuh (the device routine, compiled for the host) against the oracle's scipy restatement of
    statrs StudentsT::cdf, including the reference's 1 - (1 - ib) cancellation to exactly 0 for huge |t|."""
    rng = np.random.default_rng(0)
    for df in (1.0, 2.5, 7.0, 30.0, 197.0, 997.3, 9999.0, 1.0e6, 2.5e7):
        ts = np.concatenate([[0.0, 1e-8, 1e-3, 0.02021, 0.5, 1.0, 2.0, 5.0, 26.212765, 60.0, 1e3, np.inf], rng.normal(size=20) * 3])
        ref = S.students_t_two_sided_p(ts, df)
        got = np.array([hostcheck.lib().hc_students_t_p(float(t), df) for t in ts])
        assert np.all(np.abs(got - ref) <= 1e-7 * np.abs(ref) + 1e-15), (df, ts[np.argmax(np.abs(got - ref))])
    assert np.isnan(hostcheck.lib().hc_students_t_p(float("nan"), 5.0))


def test_cholesky_inverse():
    rng = np.random.default_rng(1)
    for n in (1, 3, 8, 17, 64):
        x = rng.normal(size=(4 * n + 5, n))
        A = np.ascontiguousarray(x.T @ x + 1e-3 * np.eye(n))
        ref = np.linalg.inv(A)
        a = A.copy()
        assert hostcheck.lib().hc_chol_inverse(a.ctypes.data, n) == 0
        assert np.allclose(a, ref, rtol=1e-9, atol=1e-12 * np.abs(ref).max())
    bad = np.ascontiguousarray(np.array([[1.0, 2.0], [2.0, 1.0]]))
    assert hostcheck.lib().hc_chol_inverse(bad.ctypes.data, 2) == 1
