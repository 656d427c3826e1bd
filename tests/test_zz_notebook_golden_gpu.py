"""The CUDA path against outputs the REFERENCE ITSELF printed: the executed cells of its demo notebook
(tests/golden/notebook_outputs.json, transcribed by tests/golden/make_golden_notebook.py; the notebook's seeded data
are regenerated bit for bit).  Printed precision is 6 decimals: tolerance 2e-6 absolute on O(1) values.
The oracle is pinned to the same numbers in tests/test_oracle.py (CPU)."""
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from polars_ols_b200 import Frame, col  # noqa: E402

NB = json.loads((Path(__file__).parent / "golden" / "notebook_outputs.json").read_text())
TOL = 2.0e-6


def _notebook_frame(n_samples=2000, n_features=3, n_groups=5, noise=0.1):
    """notebooks/polars_ols_demo.ipynb cell 1 (`_make_data`), same rng stream."""
    rng = np.random.default_rng(0)
    x = rng.normal(size=(n_samples, n_features))
    eps = rng.normal(size=n_samples, scale=noise)
    d = {f"x{j + 1}": np.ascontiguousarray(x[:, j]) for j in range(n_features)}
    d["y"] = -1 * x.sum(1) + eps
    d["group"] = rng.integers(0, n_groups, size=n_samples)
    d["sample_weights"] = rng.uniform(0, 1, size=n_samples)
    return d


def test_notebook_ols_svd_and_wls_predictions():                               # cells 5, 7
    d, g = _notebook_frame(), NB["cell7_tail10"]
    F = Frame(d)
    assert np.allclose(d["x1"][-10:], g["x1"], atol=1e-6, rtol=0)               # the regenerated frame is the printed one
    # the notebook passes null_policy="drop" as well; the frame has no nulls, so the policy does not enter the numbers
    p = F.select(col("y").least_squares.ols("x1", "x2", "x3", solve_method="svd").over("group"))["y"].to_numpy()
    assert np.allclose(p[-10:], g["predictions_ols_group"], atol=TOL, rtol=0)
    p = F.select(col("y").least_squares.ols("x1", "x2", "x3", solve_method="svd"))["y"].to_numpy()
    assert np.allclose(p[-10:], g["predictions_ols"], atol=TOL, rtol=0)
    p = F.select(col("y").least_squares.wls("x1", "x2", "x3", sample_weights="sample_weights"))["y"].to_numpy()
    assert np.allclose((p * (d["group"] == 2))[-10:], g["predictions_wls_masked"], atol=TOL, rtol=0)


def test_notebook_grouped_coefficients():                                      # cells 11, 49
    d = _notebook_frame()
    F = Frame(d)
    r = F.select(col("y").least_squares.ols("x1", "x2", "x3", add_intercept=True, mode="coefficients").over("group"))["coefficients"]
    by_key = {int(k): r.to_numpy()[i] for i, k in enumerate(r.keys)}
    for grp, want in zip(NB["cell11_head5_unnested"]["group"], NB["cell11_head5_unnested"]["coefficients"]):
        assert np.allclose(by_key[grp], want, atol=TOL, rtol=0)
    r = F.select(col("y").least_squares.ols("x1", "x2", mode="coefficients").over("group"))["coefficients"]
    for i, k in enumerate(r.keys):
        assert np.allclose(r.to_numpy()[i], NB["cell49_by_group"][str(int(k))], atol=TOL, rtol=0)


def test_notebook_regularised():                                               # cell 36
    d = _notebook_frame()
    F = Frame(d)
    r = F.select(col("y").least_squares.ridge("x1", "x2", "x3", alpha=100.0, sample_weights="sample_weights",
                                              mode="coefficients"))["coefficients"]
    assert np.allclose(r.to_numpy()[0], NB["cell36"]["coef_ridge_alpha100_weighted"], atol=TOL, rtol=0)
    r = F.select(col("y").least_squares.elastic_net("x1", "x2", "x3", alpha=0.0001, l1_ratio=0.5, positive=True,
                                                    mode="coefficients"))["coefficients"]
    assert np.allclose(r.to_numpy()[0], NB["cell36"]["coef_enet_non_negative"], atol=TOL, rtol=0)
    # (cell 30, Cholesky on an exactly collinear frame -> nulls, is pinned on the oracle only: whether the last pivot
    # rounds to a tiny positive or a non-positive number is not a property any two implementations share)


def test_notebook_rolling_rls_expanding():                                     # cell 47
    d, g = _notebook_frame(), NB["cell47"]
    F = Frame(d)
    c = F.select(col("y").least_squares.rolling_ols("x1", "x2", "x3", window_size=252, min_periods=5, alpha=0.0001,
                                                    mode="coefficients").over("group"))["coefficients"].to_numpy()
    assert np.isnan(c[:5]).all()
    assert np.allclose(c[-5:], g["rolling_ridge_coef_tail5"], atol=TOL, rtol=0)
    c = F.select(col("y").least_squares.rls("x1", "x2", "x3", half_life=21.0, initial_state_mean=[-1.0, -1.0, -1.0],
                                            initial_state_covariance=10.0, mode="coefficients").over("group"))["coefficients"].to_numpy()
    assert np.allclose(c[:5], g["rls_coef_head5"], atol=TOL, rtol=0)
    assert np.allclose(c[-5:], g["rls_coef_tail5"], atol=TOL, rtol=0)
    p = F.select(col("y").least_squares.expanding_ols("x1", "x2", "x3", mode="predictions"))["y"].to_numpy()
    assert np.allclose(p[:5], g["expanding_ols_pred_head5"], atol=TOL, rtol=0)
    assert np.allclose(p[-5:], g["expanding_ols_pred_tail5"], atol=TOL, rtol=0)


def test_notebook_out_of_sample_predict():                                     # cells 49, 50
    d, t = _notebook_frame(), _notebook_frame(n_features=5, n_groups=1)          # df_test = _make_data(n_groups=1): group == 0
    coef = Frame(d).select(col("y").least_squares.ols("x1", "x2", mode="coefficients").over("group"))["coefficients"]
    cb = coef.to_numpy()[np.searchsorted(coef.keys, t["group"])]                 # the join on "group"
    Ft = Frame({"coefficients": cb, "x1": t["x1"], "x2": t["x2"]})
    p = Ft.select(col("coefficients").least_squares.predict(col("x1"), col("x2"), name="predictions_test"))["predictions_test"]
    assert np.allclose(p.to_numpy()[:5], NB["cell50_predictions_test_head5"], atol=TOL, rtol=0)
