"""Pins the CPU oracle (oracle/) before anything is checked against it.

1. the reference's own published outputs (README known-answer frame, tests/golden/readme_frame.json; the executed
   cells of its demo notebook, tests/golden/notebook_outputs.json — data regenerated from the notebook's seeded recipe);
2. the reference's Rust unit tests (src/lib.rs:47-171) re-derived on the same data recipe;
3. the third-party oracles the reference's tests/test_ols.py uses (numpy lstsq/solve, sklearn).
"""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import lib
from oracle import semantics as S

GOLD = json.loads((Path(__file__).parent / "golden" / "readme_frame.json").read_text())
F = {k: np.asarray(v, dtype=np.float64) for k, v in GOLD["frame"].items()}
TOLC = GOLD["printed_abs_tol_coefficients"]
TOLR = GOLD["printed_abs_tol_round2"]
NB = json.loads((Path(__file__).parent / "golden" / "notebook_outputs.json").read_text())
TOLN = NB["printed_abs_tol"]


def notebook_frame(n_samples=2000, n_features=3, n_groups=5, noise=0.1):
    """notebooks/polars_ols_demo.ipynb cell 1 (`_make_data`), same rng stream."""
    rng = np.random.default_rng(0)
    x = rng.normal(size=(n_samples, n_features))
    eps = rng.normal(size=n_samples, scale=noise)
    return {"x": [np.ascontiguousarray(x[:, j]) for j in range(n_features)], "y": -1 * x.sum(1) + eps,
            "group": rng.integers(0, n_groups, size=n_samples), "w": rng.uniform(0, 1, size=n_samples)}


def _make_data(n_samples=5000, n_features=2, n_groups=None, scale=0.1, sparsity=0.0, add_missing=False):
    """tests/test_ols.py:22-51 (same rng stream)."""
    rng = np.random.default_rng(0)
    x = rng.normal(size=(n_samples, n_features))
    eps = rng.normal(size=n_samples, scale=scale)
    y = x[:, : int(n_features * (1.0 - sparsity))].sum(1) + eps
    out = {"x": x, "y": y}
    if n_groups is not None:
        out["group"] = rng.integers(n_groups, size=n_samples)
    if add_missing:
        cols = [x[:, j].copy() for j in range(n_features)] + [y.copy()]
        masks = []
        for c in cols:
            m = np.array([not (rng.random() < 0.1) for _ in range(n_samples)])
            masks.append(m)
        out["masks"] = masks
    return out


# ----------------------------------------------------------------------------- the reference's executed notebook
def test_notebook_data_recipe_reproduces_the_printed_frame():
    d = notebook_frame()
    assert np.allclose(d["x"][0][-10:], NB["cell7_tail10"]["x1"], atol=TOLN, rtol=0)


def test_notebook_ols_svd_and_wls_predictions():                               # cells 5, 7
    d, g = notebook_frame(), NB["cell7_tail10"]
    kw = S.OLSKwargs(null_policy="drop", solve_method="svd")
    p, _ = S.over(S.least_squares, d["group"], d["y"], *d["x"], kwargs=kw)
    assert np.allclose(np.asarray(p)[-10:], g["predictions_ols_group"], atol=TOLN, rtol=0)
    p, _ = S.least_squares(d["y"], *d["x"], kwargs=kw)
    assert np.allclose(p[-10:], g["predictions_ols"], atol=TOLN, rtol=0)
    p, _ = S.least_squares(d["y"], *d["x"], sample_weights=d["w"])
    assert np.allclose((p * (d["group"] == 2))[-10:], g["predictions_wls_masked"], atol=TOLN, rtol=0)


def test_notebook_grouped_coefficients():                                      # cells 11, 49
    d = notebook_frame()
    keys, c, _ = S.over(S.least_squares, d["group"], d["y"], *d["x"], add_intercept=True, per_group=True, mode="coefficients")
    by_key = {int(k): c[i] for i, k in enumerate(keys)}
    for grp, want in zip(NB["cell11_head5_unnested"]["group"], NB["cell11_head5_unnested"]["coefficients"]):
        assert np.allclose(by_key[grp], want, atol=TOLN, rtol=0)
    keys, c, _ = S.over(S.least_squares, d["group"], d["y"], d["x"][0], d["x"][1], per_group=True, mode="coefficients")
    for i, k in enumerate(keys):
        assert np.allclose(c[i], NB["cell49_by_group"][str(int(k))], atol=TOLN, rtol=0)


def test_notebook_regularised_and_collinear():                                 # cells 26, 30, 36
    d = notebook_frame()
    c, _ = S.least_squares(d["y"], *d["x"], sample_weights=d["w"], mode="coefficients", kwargs=S.OLSKwargs(alpha=100.0, l1_ratio=0.0))
    assert np.allclose(c, NB["cell36"]["coef_ridge_alpha100_weighted"], atol=TOLN, rtol=0)
    c, _ = S.least_squares(d["y"], *d["x"], mode="coefficients", kwargs=S.OLSKwargs(alpha=0.0001, l1_ratio=0.5, positive=True))
    assert np.allclose(c, NB["cell36"]["coef_enet_non_negative"], atol=TOLN, rtol=0)   # all true coefficients are -1: NNLS -> 0
    x1, x2 = d["x"][0], d["x"][1]
    c, m = S.least_squares(x1 + 2 * x2, x1, x2, x2.copy(), mode="coefficients", kwargs=S.OLSKwargs(solve_method="chol"))
    assert NB["cell30_collinear_chol"] == [None, None, None] and not m.any()             # Cholesky and LU fail -> nulls


def test_notebook_rolling_rls_expanding():                                     # cell 47
    d, g = notebook_frame(), NB["cell47"]
    c, _ = S.over(S.rolling_least_squares, d["group"], d["y"], *d["x"], mode="coefficients",
                  kwargs=S.RollingKwargs(window_size=252, min_periods=5, alpha=0.0001, null_policy="drop"))
    c = np.asarray(c)
    assert np.isnan(c[:5]).all() and all(v is None for row in g["rolling_ridge_coef_head5"] for v in row)
    assert np.allclose(c[-5:], g["rolling_ridge_coef_tail5"], atol=TOLN, rtol=0)
    c, _ = S.over(S.recursive_least_squares, d["group"], d["y"], *d["x"], mode="coefficients",
                  kwargs=S.RLSKwargs(half_life=21.0, initial_state_mean=[-1.0, -1.0, -1.0], initial_state_covariance=10.0, null_policy="drop"))
    c = np.asarray(c)
    assert np.allclose(c[:5], g["rls_coef_head5"], atol=TOLN, rtol=0) and np.allclose(c[-5:], g["rls_coef_tail5"], atol=TOLN, rtol=0)
    p, _ = S.recursive_least_squares(d["y"], *d["x"], mode="predictions", kwargs=S.RLSKwargs(half_life=None, null_policy="drop"))
    assert np.allclose(p[:5], g["expanding_ols_pred_head5"], atol=TOLN, rtol=0)
    assert np.allclose(p[-5:], g["expanding_ols_pred_tail5"], atol=TOLN, rtol=0)


def test_notebook_out_of_sample_predict():                                     # cells 49, 50
    d, t = notebook_frame(), notebook_frame(n_features=5, n_groups=1)           # df_test = _make_data(n_groups=1): group == 0
    keys, c, _ = S.over(S.least_squares, d["group"], d["y"], d["x"][0], d["x"][1], per_group=True, mode="coefficients")
    b = c[list(keys).index(0)]
    n = len(t["y"])
    p = S.predict([np.full(n, b[0]), np.full(n, b[1])], [t["x"][0], t["x"][1]], "zero")
    assert np.allclose(_vals(p)[:5], NB["cell50_predictions_test_head5"], atol=TOLN, rtol=0)


def _vals(values_mask):
    return values_mask[0] if isinstance(values_mask, tuple) else values_mask


# ----------------------------------------------------------------------------- README golden frame
def test_readme_ols_coefficients_with_intercept():
    c, m = S.least_squares(F["y"], F["x1"], F["x2"], add_intercept=True, mode="coefficients")
    assert m.all()
    assert np.allclose(c, GOLD["coefficients_ols_intercept"], atol=TOLC, rtol=0)


def test_readme_ols_coefficients_over_group():
    keys, c, m = S.over(S.least_squares, F["group"], F["y"], F["x1"], F["x2"], per_group=True,
                        add_intercept=True, mode="coefficients")
    for i, k in enumerate(keys):
        assert np.allclose(c[i], GOLD["coefficients_ols_intercept_by_group"][str(int(k))], atol=TOLC, rtol=0)


def test_readme_rls_coefficients_over_group():
    c, m = S.over(S.recursive_least_squares, F["group"], F["y"], F["x1"], F["x2"], mode="coefficients",
                  kwargs=S.RLSKwargs())
    assert np.allclose(c[:5], GOLD["coefficients_rls_group1"], atol=TOLC, rtol=0)


def test_readme_lasso_and_wls_predictions():
    p, _ = S.over(S.least_squares, F["group"], F["y"], F["x1"], F["x2"], add_intercept=True,
                  kwargs=S.OLSKwargs(alpha=0.0001, l1_ratio=1.0))
    assert np.allclose(p[:5], GOLD["predictions_lasso_head5_round2"], atol=TOLR, rtol=0)
    p, _ = S.least_squares(F["y"], F["x1"], F["x2"], sample_weights=F["weights"])
    assert np.allclose(p[:5], GOLD["predictions_wls_head5_round2"], atol=TOLR, rtol=0)


# ----------------------------------------------------------------------------- src/lib.rs unit tests
def _rust_data(seed=0):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(10_000, 2))
    return x[:, 0] + x[:, 1], np.ascontiguousarray(x)         # src/lib.rs:28-45


def test_rust_ols_ridge_elastic_net():
    y, x = _rust_data()
    assert np.linalg.norm(S.solve_ols(y, x, "qr") - 1.0) < 1e-3            # :47-55
    assert np.linalg.norm(S.solve_ols(y, x, "svd") - 1.0) < 1e-3
    for m in ("chol", "svd"):                                                 # :57-65
        assert np.linalg.norm(S.solve_ridge(y, x, 10.0, m, None) - 0.999) < 1e-2
    w = S.solve_elastic_net(y, x, 0.001, 0.5, 1000, 1e-4, False, None)        # :67-82
    assert np.linalg.norm(w - 0.999) < 1e-2


def test_rust_rls_and_rolling():
    y, x = _rust_data()
    c = S.solve_recursive_least_squares(y, x, 252.0, 0.01, None, np.ones(len(y), bool))   # :84-101
    assert np.linalg.norm(c[-1] - 1.0) < 1e-4
    c = S.solve_rolling_ols(y, x, 1000, 100, False, None, np.ones(len(y), bool), "drop_window")  # :103-122
    assert np.linalg.norm(c[-1] - 1.0) < 1e-4
    assert np.isnan(c[:99]).all() and not np.isnan(c[99:]).any()


def test_rust_woodbury():
    # src/lib.rs:145-171: rank-2 update of inv(X^T X) equals re-inversion
    rng = np.random.default_rng(1)
    x = rng.normal(size=(252, 5))
    x_new, x_old = rng.normal(size=5), x[0]
    inv = np.linalg.inv(x.T @ x)
    upd = np.ascontiguousarray(np.stack([-x_old, x_new]))
    c = np.array([-1.0, 1.0])
    a = np.ascontiguousarray(inv.copy())
    lib().orc_update_xtx_inv(a.ctypes.data, 5, upd.ctypes.data, c.ctypes.data, 2)
    x2 = np.vstack([x[1:], x_new])
    assert np.allclose(a, np.linalg.inv(x2.T @ x2), atol=1e-5)


# ----------------------------------------------------------------------------- tests/test_ols.py oracles
@pytest.mark.parametrize("solve_method", ["qr", "svd", "chol", "lu", None])
def test_ols_vs_lstsq(solve_method):                                          # tests/test_ols.py:54-73
    d = _make_data(1000, 2)
    c, _ = S.least_squares(d["y"], d["x"][:, 0], d["x"][:, 1], mode="coefficients",
                           kwargs=S.OLSKwargs(solve_method=solve_method))
    ref = np.linalg.lstsq(d["x"], d["y"], rcond=None)[0]
    assert np.allclose(c, ref, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("solve_method", ["svd", "chol"])
def test_ridge_vs_numpy_sklearn(solve_method):                               # tests/test_ols.py:475-503
    from sklearn.linear_model import Ridge
    d = _make_data(5000, 3)
    alpha = 0.01
    c, _ = S.least_squares(d["y"], *d["x"].T, mode="coefficients",
                           kwargs=S.OLSKwargs(alpha=alpha, l1_ratio=0.0, solve_method=solve_method))
    x, y = d["x"], d["y"]
    ref = np.linalg.solve(x.T @ x + alpha * np.eye(3), x.T @ y)
    assert np.allclose(c, ref, rtol=1e-10, atol=1e-12)
    sk = Ridge(alpha=alpha, fit_intercept=False, solver="svd").fit(x, y).coef_
    assert np.allclose(c, sk, rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("k,sparsity,alpha,method", [(10, 0.5, 0.001, "cd"), (10, 0.5, 0.001, "cd_active_set"),
                                                      (100, 0.9, 0.01, "cd"), (4, 0.0, 0.1, "cd")])
def test_elastic_net_vs_sklearn(k, sparsity, alpha, method):                  # tests/test_ols.py:561-599
    from sklearn.linear_model import ElasticNet
    d = _make_data(5000, k, sparsity=sparsity)
    c, _ = S.least_squares(d["y"], *d["x"].T, mode="coefficients",
                           kwargs=S.OLSKwargs(alpha=alpha, l1_ratio=0.5, tol=1e-7, solve_method=method))
    sk = ElasticNet(alpha=alpha, l1_ratio=0.5, fit_intercept=False, max_iter=10000, tol=1e-10).fit(d["x"], d["y"]).coef_
    assert np.allclose(c, sk, rtol=1e-4, atol=1e-4)


def test_elastic_net_positive_vs_sklearn():                                   # tests/test_ols.py:602-630
    from sklearn.linear_model import ElasticNet
    rng = np.random.default_rng(0)
    x = rng.normal(size=(2000, 5))
    y = x @ np.array([1.0, -1.0, 0.5, -0.5, 0.0]) + 0.1 * rng.normal(size=2000)
    c, _ = S.least_squares(y, *x.T, mode="coefficients",
                           kwargs=S.OLSKwargs(alpha=0.001, l1_ratio=0.5, positive=True, tol=1e-8))
    sk = ElasticNet(alpha=0.001, l1_ratio=0.5, fit_intercept=False, positive=True, tol=1e-10, max_iter=10000).fit(x, y).coef_
    assert (c >= 0).all() and np.allclose(c, sk, atol=1e-5)


def test_rls_matches_expanding_ols_and_information_form():                    # tests/test_ols.py:633-681
    d = _make_data(2000, 3)
    y, x = d["y"], np.ascontiguousarray(d["x"])
    c = S.solve_recursive_least_squares(y, x, None, 1e6, None, np.ones(len(y), bool))
    assert np.allclose(c[-1], np.linalg.lstsq(x, y, rcond=None)[0], rtol=1e-4, atol=1e-4)
    # information form with forgetting
    lam, p0 = np.exp(np.log(0.5) / 50.0), 10.0
    c = S.solve_recursive_least_squares(y, x, 50.0, p0, None, np.ones(len(y), bool))
    A, b = np.eye(3) / p0, np.zeros(3)
    for t in range(300):
        A = lam * A + np.outer(x[t], x[t]); b = lam * b + x[t] * y[t]
    assert np.allclose(c[299], np.linalg.solve(A, b), rtol=1e-9)


@pytest.mark.parametrize("window,min_periods,woodbury", [(50, None, False), (50, 10, True), (252, 5, False)])
def test_rolling_vs_direct_windows(window, min_periods, woodbury):            # tests/test_ols.py:718-772
    d = _make_data(600, 3)
    y, x = d["y"], np.ascontiguousarray(d["x"])
    c = S.solve_rolling_ols(y, x, window, min_periods, woodbury, None, np.ones(len(y), bool), "drop")
    mp = min_periods or 3
    for i in (mp - 1, window - 1, window, 400, 599):
        lo = max(0, i - window + 1)
        ref = np.linalg.lstsq(x[lo:i + 1], y[lo:i + 1], rcond=None)[0]
        assert np.allclose(c[i], ref, rtol=1e-6, atol=1e-8), i
    assert np.isnan(c[:mp - 1]).all()


def test_rolling_drop_equals_drop_then_roll():                                # tests/test_ols.py:809-841
    d = _make_data(800, 2, add_missing=True)
    cols = [(d["y"], d["masks"][2]), (d["x"][:, 0], d["masks"][0]), (d["x"][:, 1], d["masks"][1])]
    kw = S.RollingKwargs(window_size=21, min_periods=2, null_policy="drop")
    c, m = S.plugin_rolling_least_squares_coefficients(cols, kw)
    valid = d["masks"][0] & d["masks"][1] & d["masks"][2]
    dropped = [(v[valid], None) for v, _ in cols]
    c2, _ = S.plugin_rolling_least_squares_coefficients(dropped, kw)
    # re-align with forward fill
    idx = np.cumsum(valid) - 1
    ok = idx >= 0
    assert np.allclose(c[ok], c2[idx[ok]], equal_nan=True)


@pytest.mark.parametrize("null_policy", ["drop", "drop_zero", "drop_y_zero_x"])
def test_missing_data_predictions(null_policy):                               # tests/test_ols.py:179-249
    d = _make_data(add_missing=True)
    x = d["x"].copy(); y = d["y"].copy()
    mx0, mx1, my = d["masks"]
    x[~mx0, 0] = np.nan; x[~mx1, 1] = np.nan; y[~my] = np.nan
    if null_policy == "drop_y_zero_x":
        is_valid = ~np.isnan(y)
        coef = np.linalg.lstsq(np.nan_to_num(x[is_valid]), y[is_valid], rcond=None)[0]
    else:
        is_valid = ~np.isnan(x).any(axis=1) & ~np.isnan(y)
        coef = np.linalg.lstsq(x[is_valid], y[is_valid], rcond=None)[0]
    expected = np.nan_to_num(x) @ coef
    if null_policy == "drop":
        expected[~is_valid] = np.nan
    p, m = S.least_squares((d["y"], my), (d["x"][:, 0], mx0), (d["x"][:, 1], mx1),
                           kwargs=S.OLSKwargs(null_policy=null_policy))
    got = np.where(m if m is not None else True, p, np.nan)
    assert np.allclose(got, expected, rtol=1e-8, atol=1e-10, equal_nan=True)
    r, m = S.least_squares((d["y"], my), (d["x"][:, 0], mx0), (d["x"][:, 1], mx1), mode="residuals",
                           kwargs=S.OLSKwargs(null_policy=null_policy))
    got = np.where(m, r, np.nan)
    assert np.allclose(got, y - expected, rtol=1e-8, atol=1e-10, equal_nan=True)


def test_all_empty_data():                                                    # tests/test_ols.py:252-269
    a = (np.array([0.0, 2, 0, 4]), np.array([False, True, False, True]))
    b = (np.array([1.0, 0, 3, 0]), np.array([True, False, True, False]))
    r, m = S.least_squares(a, b, mode="residuals", kwargs=S.OLSKwargs(null_policy="drop", solve_method="svd"))
    assert not m.any()


def test_wls_vs_closed_form():                                                # tests/test_ols.py:506-541
    d = _make_data(2000, 3)
    rng = np.random.default_rng(5)
    w = rng.uniform(0.1, 1.0, size=2000)
    c, _ = S.least_squares(d["y"], *d["x"].T, sample_weights=w, mode="coefficients")
    xw = d["x"] * w[:, None]
    ref = np.linalg.solve(d["x"].T @ xw, xw.T @ d["y"])
    assert np.allclose(c, ref, rtol=1e-9)
    p, _ = S.least_squares(d["y"], *d["x"].T, sample_weights=w)
    assert np.allclose(p, d["x"] @ ref, rtol=1e-8, atol=1e-10)


def test_grouped_driver_matches_per_group():
    d = _make_data(4000, 4, n_groups=None)
    offs = np.array([0, 1000, 1001, 1001, 2500, 4000], dtype=np.int64)
    cols = [np.ascontiguousarray(d["y"])] + [np.ascontiguousarray(d["x"][:, j]) for j in range(4)]
    import ctypes as C
    arr = (C.c_void_p * 5)(*[c.ctypes.data for c in cols])
    out = np.empty((5, 4))
    lib().orc_grouped_least_squares_coefficients(arr, 4, offs.ctypes.data, 5, 0, 1e-3, 0.0, 1000, 1e-5, 0, 0, out.ctypes.data)
    for g in range(5):
        sl = slice(offs[g], offs[g + 1])
        if offs[g + 1] == offs[g]:
            assert (out[g] == 0).all()
            continue
        x = d["x"][sl]
        ref = np.linalg.solve(x.T @ x + 1e-3 * np.eye(4), x.T @ d["y"][sl])
        assert np.allclose(out[g], ref, rtol=1e-9)


def test_predict_semantics():                                                 # tests/test_ols.py:903-966
    rng = np.random.default_rng(3)
    x = rng.normal(size=(50, 3)); c = rng.normal(size=(50, 4))
    mx = rng.random(50) >= 0.2
    p, m = S.predict([c[:, j] for j in range(4)], [(x[:, 0], mx), x[:, 1], x[:, 2]], "zero", add_intercept=True)
    assert m is None and np.allclose(p, np.where(mx, x[:, 0], 0) * c[:, 0] + x[:, 1] * c[:, 1] + x[:, 2] * c[:, 2] + c[:, 3])
    p, m = S.predict([c[:, j] for j in range(3)], [(x[:, 0], mx), x[:, 1], x[:, 2]], "drop")
    assert (m == mx).all()
    p, m = S.predict([c[:, j] for j in range(3)], [(x[:, 0], mx), x[:, 1], x[:, 2]], "ignore")
    assert (np.isnan(p) == ~mx).all()


# ----------------------------------------------------------------------------- §8f rows: statistics, multi-target
def test_readme_statistics_frame():
    """README.md:143-165 (the reference's printed statistics struct) pins mode="statistics" of the oracle."""
    g = GOLD["statistics_ols_intercept"]
    r = S.least_squares_statistics(F["y"], F["x1"], F["x2"], add_intercept=True, kwargs=S.OLSKwargs(alpha=0.0))
    rt = g["printed_rel_tol"]
    for k in ("r2", "mae", "mse"):
        assert r[k] == pytest.approx(g[k], rel=rt)
    for k in ("standard_errors", "t_values", "p_values"):
        assert np.allclose(r[k], g[k], rtol=rt, atol=0)
    assert np.allclose(r["coefficients"], g["coefficients"], atol=TOLC, rtol=0)


def test_statistics_against_textbook_ols_and_scipy():
    """tests/test_ols.py:998-1029 compares against statsmodels (absent here): the same closed forms via scipy."""
    from scipy import stats
    d = _make_data(n_samples=2000, n_features=4, scale=1.0)
    x, y = d["x"], d["y"]
    r = S.least_squares_statistics(y, *x.T, add_intercept=True, kwargs=S.OLSKwargs(alpha=0.0))
    X = np.column_stack([x, np.ones(len(y))])
    b = np.linalg.lstsq(X, y, rcond=None)[0]
    res = y - X @ b
    dof = len(y) - X.shape[1]
    se = np.sqrt(res @ res / dof * np.diag(np.linalg.inv(X.T @ X)))
    assert np.allclose(r["coefficients"], b, rtol=1e-9)
    assert np.allclose(r["standard_errors"], se, rtol=1e-9)
    assert np.allclose(r["t_values"], b / se, rtol=1e-8)
    pv = 2 * stats.t.sf(np.abs(b / se), dof)
    big = pv > 1e-12                                     # below that the reference's 1 - (1 - ib) has cancelled
    assert np.allclose(r["p_values"][big], pv[big], rtol=1e-6)
    assert r["r2"] == pytest.approx(1 - res @ res / ((y - y.mean()) ** 2).sum(), rel=1e-12)
    assert r["mae"] == pytest.approx(np.abs(res).mean(), rel=1e-12)
    # ridge: df = n - trace((X^T X + lambda I)^-1)  (src/statistics.rs:125-129)
    rr = S.least_squares_statistics(y, *x.T, kwargs=S.OLSKwargs(alpha=10.0, l1_ratio=0.0))
    inv = np.linalg.inv(x.T @ x + 10.0 * np.eye(4))
    br = inv @ x.T @ y
    rs = y - x @ br
    assert np.allclose(rr["standard_errors"], np.sqrt(rs @ rs / (len(y) - np.trace(inv)) * np.diag(inv)), rtol=1e-9)


@pytest.mark.parametrize("alpha,policy", [(0.0, "ignore"), (0.0, "drop"), (1e-4, "drop_y_zero_x"), (0.01, "drop_zero"), (0.5, "zero")])
def test_multi_target_equals_independent_regressions(alpha, policy):
    """tests/test_ols.py:76-127: a multi-target fit equals one single-target (svd) fit per target."""
    d = _make_data(n_samples=3000, n_features=3, add_missing=policy not in ("zero", "ignore"))
    x = d["x"]
    xs = [(x[:, j], d["masks"][j] if (j == 0 and "masks" in d) else None) for j in range(3)]
    ys = [x[:, 0] + x[:, 1] + x[:, 2], x[:, 0] - x[:, 1] + x[:, 2], -x[:, 0] + x[:, 1] - x[:, 2]]
    ys = [(v, xs[0][1]) for v in ys]                                   # y_t inherits x1's nulls, as in the reference test
    kw = S.OLSKwargs(null_policy=policy, solve_method="svd", alpha=alpha)
    v, m = S.multi_target_least_squares(ys, *xs, mode="residuals", kwargs=kw)
    for t in range(3):
        v1, m1 = S.least_squares(ys[t], *xs, mode="residuals", kwargs=kw)
        m1 = np.ones(len(v1), bool) if m1 is None else m1
        both = m[:, t] & m1 & ~np.isnan(v1)
        assert (m[:, t] == (m1 & ~np.isnan(v1))).all() or policy == "ignore"
        assert np.allclose(v[both, t], v1[both], atol=1e-9)
