"""numpy restatement of the reference's expression-level semantics (test infrastructure only).

Follows, line by line:
  * ``polars_ols/least_squares.py:163-239,372-409``  — ``_pre_process_data`` (intercept, WLS sqrt-w
    scaling), plugin selection, ``predictions *= 1/sqrt_w``, ``residuals = target - predictions``,
    rolling ``fill_nan(None)``;
  * ``src/expressions.rs:22-103,145-296``            — null policies, validity masks, f64 casts;
  * ``src/expressions.rs:351-446,594-701``           — solver dispatch, fit/predict asymmetry,
    RLS / rolling entry points;
  * solvers: ``oracle/ols_oracle.c`` (restating ``src/least_squares.rs``) and, for the LAPACK
    ``dgelsd`` paths (``src/least_squares.rs:183-191``), ``numpy.linalg.lstsq`` — the same routine.

A column is a pair ``(values, valid)``: ``values`` a float ndarray (f32 or f64), ``valid`` a bool
ndarray or ``None`` (no nulls).  NaN is a VALUE, not a null (polars semantics).
Outputs use the same convention; ``valid=False`` marks a polars null.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .loader import lib

Col = Tuple[np.ndarray, Optional[np.ndarray]]

_EPSILON = 1.0e-12  # polars_ols/least_squares.py:63


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def as_col(c) -> Col:
    if isinstance(c, tuple):
        v, m = c
        v = np.asarray(v)
        if v.dtype not in (np.float32, np.float64):
            v = v.astype(np.float64)
        return v, (None if m is None else np.asarray(m, dtype=bool))
    v = np.asarray(c)
    if v.dtype not in (np.float32, np.float64):
        v = v.astype(np.float64)
    return v, None


def _is_valid(c: Col) -> np.ndarray:
    return np.ones(len(c[0]), dtype=bool) if c[1] is None else c[1]


def _mul(a: Col, b: Col) -> Col:
    """polars `a * b`: dtype promotion like numpy, null if either side is null."""
    v = a[0] * b[0]
    if a[1] is None and b[1] is None:
        return v, None
    return v, _is_valid(a) & _is_valid(b)


# ---------------------------------------------------------------------------------------------
# kwargs (polars_ols/least_squares.py:66-160  <->  src/expressions.rs:298-330)
# ---------------------------------------------------------------------------------------------
@dataclass
class OLSKwargs:
    null_policy: str = "ignore"
    alpha: Optional[float] = 0.0
    l1_ratio: Optional[float] = None
    max_iter: Optional[int] = 1000
    tol: Optional[float] = 1.0e-5
    positive: Optional[bool] = False
    solve_method: Optional[str] = None
    rcond: Optional[float] = None


@dataclass
class RLSKwargs:
    null_policy: str = "drop"
    half_life: Optional[float] = None
    initial_state_covariance: Optional[float] = 10.0
    initial_state_mean: Optional[Sequence[float]] = None


@dataclass
class RollingKwargs:
    null_policy: str = "drop_window"
    window_size: int = 1_000_000
    min_periods: Optional[int] = None
    use_woodbury: Optional[bool] = None
    alpha: Optional[float] = None


# ---------------------------------------------------------------------------------------------
# L4: _pre_process_data (polars_ols/least_squares.py:163-196)
# ---------------------------------------------------------------------------------------------
def pre_process(target: Col, features: List[Col], sample_weights: Optional[Col], add_intercept: bool):
    features = list(features)
    if add_intercept:
        # target.fill_null(0.0).mul(0.0).add(1.0).alias("const")   (:188) — appended LAST
        tv = np.where(_is_valid(target), target[0], 0.0).astype(target[0].dtype)
        features.append((tv * 0.0 + 1.0, None))
    sqrt_w = None
    if sample_weights is not None:
        # sqrt_w = w.sqrt().fill_null(1e-12)  (:193)
        wv = np.sqrt(sample_weights[0])
        if sample_weights[1] is not None:
            wv = np.where(sample_weights[1], wv, np.asarray(_EPSILON, dtype=wv.dtype))
        sqrt_w = (wv, None)
        target = _mul(target, sqrt_w)                     # :194
        features = [_mul(f, sqrt_w) for f in features]    # :195
    return target, features, sqrt_w


# ---------------------------------------------------------------------------------------------
# L2: null handling (src/expressions.rs:201-296) and conversion (src/expressions.rs:22-103)
# ---------------------------------------------------------------------------------------------
def compute_is_valid_mask(inputs: List[Col], null_policy: str) -> Optional[np.ndarray]:
    if null_policy in ("drop", "drop_zero", "drop_window"):           # :209-216
        m = _is_valid(inputs[0]).copy()
        for c in inputs[1:]:
            m &= _is_valid(c)
        return m
    if null_policy == "drop_y_zero_x":                                  # :218-219
        return _is_valid(inputs[0]).copy()
    return None                                                         # zero / ignore (:226)


def _to_f64(c: Col, fill: float) -> np.ndarray:
    v = c[0].astype(np.float64)
    if c[1] is not None:
        v = np.where(c[1], v, fill)
    return v


def convert_to_ndarray(inputs: List[Col], null_policy: str, is_valid: Optional[np.ndarray]):
    """convert_polars_to_ndarray + handle_nulls (src/expressions.rs:66-103, 257-296):
    returns (y [n], X [n,k] row-major) in f64; remaining nulls become NaN."""
    if null_policy == "zero":                                           # :264-271
        cols = [(np.where(_is_valid(c), c[0], 0).astype(c[0].dtype), None) for c in inputs]
    elif null_policy == "drop_y_zero_x":                                # :272-282
        cols = [(np.where(_is_valid(c), c[0], 0).astype(c[0].dtype)[is_valid], None) for c in inputs]
    elif null_policy in ("drop", "drop_zero", "drop_window"):           # :283-291
        cols = [(c[0][is_valid], None if c[1] is None else c[1][is_valid]) for c in inputs]
    else:                                                               # ignore (:294)
        cols = inputs
    y = _to_f64(cols[0], np.nan)                                        # :79-91
    if len(cols) > 1:
        x = np.stack([_to_f64(c, np.nan) for c in cols[1:]], axis=1)    # :22-63 (fill_zero=false)
    else:
        x = np.zeros((len(y), 0))
    return np.ascontiguousarray(y), np.ascontiguousarray(x)


def features_zero_filled(features: List[Col]) -> np.ndarray:
    """construct_features_array(.., fill_zero=true) (src/expressions.rs:30-44)."""
    return np.ascontiguousarray(np.stack([_to_f64(c, 0.0) for c in features], axis=1))


# ---------------------------------------------------------------------------------------------
# L1: solvers (src/least_squares.rs) through the C restatement / LAPACK
# ---------------------------------------------------------------------------------------------
def solve_ols(y: np.ndarray, x: np.ndarray, solve_method: Optional[str]) -> np.ndarray:
    """solve_ols (src/least_squares.rs:211-240): QR if n > k else SVD (LAPACK dgelsd)."""
    n, k = x.shape
    if solve_method is None:
        solve_method = "qr" if n > k else "svd"
    if solve_method == "qr":
        beta = np.empty(k)
        lib().orc_solve_ols_qr(_ptr(y), _ptr(x), n, k, _ptr(beta))
        return beta
    if solve_method == "svd":
        if np.isnan(x).any() or np.isnan(y).any():
            return np.full(k, np.nan)
        # ndarray-linalg least_squares -> dgelsd with rcond < 0 (machine precision)
        return np.linalg.lstsq(x, y, rcond=-1)[0]
    raise ValueError("Only 'QR' and 'SVD' are currently supported solve methods for OLS.")  # :231


def solve_ridge_svd(y, x, alpha, rcond):
    """solve_ridge_svd (src/least_squares.rs:106-168)."""
    u, s, vt = np.linalg.svd(x, full_matrices=False)
    cutoff = (rcond if rcond is not None else np.finfo(np.float64).eps * max(x.shape)) * s.max()
    s = np.where(s < cutoff, 0.0, s)
    d = s / (s * s + alpha)
    uty = u.T @ y
    return vt.T @ ((d[:, None] * uty) if uty.ndim == 2 else (d * uty))      # multi-target: d scales the rows of U^T Y


def solve_ridge(y, x, alpha, solve_method, rcond) -> np.ndarray:
    """solve_ridge (src/least_squares.rs:342-371)."""
    assert alpha >= 0.0, "alpha must be non-negative"
    n, k = x.shape
    if solve_method in (None, "chol", "lu"):
        beta = np.empty(k)
        lib().orc_solve_ridge(_ptr(y), _ptr(x), n, k, float(alpha), 1 if solve_method == "lu" else 0, _ptr(beta))
        return beta
    if solve_method == "svd":
        return solve_ridge_svd(y, x, alpha, rcond)
    raise ValueError("Only 'Cholesky', 'LU', & 'SVD' are currently supported solver methods for Ridge.")


def solve_elastic_net(y, x, alpha, l1_ratio, max_iter, tol, positive, solve_method,
                      return_sweeps: bool = False):
    """solve_elastic_net (src/least_squares.rs:386-492)."""
    l1_ratio = 0.5 if l1_ratio is None else l1_ratio
    max_iter = 1000 if max_iter is None else max_iter
    tol = 1e-5 if tol is None else tol
    positive = bool(positive)
    solve_method = solve_method or "cd"
    if solve_method not in ("cd", "cd_active_set"):
        raise ValueError("Only solve_method 'CD' (coordinate descent) is currently supported")
    assert alpha > 0.0, "'alpha' must be strictly positive"
    assert 0.0 <= l1_ratio <= 1.0
    n, k = x.shape
    w = np.empty(k)
    sweeps = lib().orc_solve_elastic_net(_ptr(y), _ptr(x), n, k, float(alpha), float(l1_ratio), int(max_iter),
                                         float(tol), int(positive), int(solve_method == "cd_active_set"), _ptr(w))
    return (w, sweeps) if return_sweeps else w


def get_least_squares_coefficients(y: np.ndarray, x: np.ndarray, kw: OLSKwargs) -> np.ndarray:
    """_get_least_squares_coefficients (src/expressions.rs:351-388)."""
    if x.size == 0:                                                      # :357-359
        return np.zeros(x.shape[1])
    alpha = 0.0 if kw.alpha is None else kw.alpha
    positive = bool(kw.positive)
    m = kw.solve_method
    if alpha == 0.0 and not positive and m in (None, "svd", "qr"):       # :366-373
        return solve_ols(y, x, m)
    if alpha >= 0.0 and (0.0 if kw.l1_ratio is None else kw.l1_ratio) == 0.0 and not positive:
        return solve_ridge(y, x, alpha, m, kw.rcond)                     # :374-375
    return solve_elastic_net(y, x, alpha, kw.l1_ratio, kw.max_iter, kw.tol, kw.positive, m)


def solve_recursive_least_squares(y, x, half_life, initial_state_covariance, initial_state_mean, is_valid):
    n, k = x.shape
    out = np.empty((n, k))
    mean = None if initial_state_mean is None else np.ascontiguousarray(initial_state_mean, dtype=np.float64)
    iv = np.ascontiguousarray(is_valid, dtype=np.uint8)
    lib().orc_solve_recursive_least_squares(
        _ptr(y), _ptr(x), n, k, float("nan") if half_life is None else float(half_life),
        10.0 if initial_state_covariance is None else float(initial_state_covariance), _ptr(mean), _ptr(iv), _ptr(out))
    return out


def solve_rolling_ols(y, x, window_size, min_periods, use_woodbury, alpha, is_valid, null_policy):
    n, k = x.shape
    out = np.empty((n, k))
    iv = np.ascontiguousarray(is_valid, dtype=np.uint8)
    branch = 0 if null_policy in ("drop", "drop_zero", "drop_y_zero_x") else 1   # :947-950
    lib().orc_solve_rolling_ols(
        _ptr(y), _ptr(x), n, k, int(window_size), -1 if min_periods is None else int(min_periods),
        -1 if use_woodbury is None else int(bool(use_woodbury)), float("nan") if alpha is None else float(alpha),
        _ptr(iv), branch, _ptr(out))
    return out


# ---------------------------------------------------------------------------------------------
# L3: the six plugin entry points (single group = one plugin call)
# ---------------------------------------------------------------------------------------------
def plugin_least_squares(inputs: List[Col], kw: OLSKwargs) -> Col:
    """least_squares (src/expressions.rs:391-428) -> predictions [n] with nulls."""
    pol = kw.null_policy
    is_valid = compute_is_valid_mask(inputs, pol)
    y_fit, x_fit = convert_to_ndarray(inputs, pol, is_valid)
    coef = get_least_squares_coefficients(y_fit, x_fit, kw)
    if pol in ("ignore", "zero"):
        return x_fit @ coef, None                                        # :398-405 (mask is None)
    x_pred = features_zero_filled(inputs[1:])                            # :408
    pred = x_pred @ coef
    if pol == "drop":
        return pred, is_valid                                            # :409-416
    return pred, None                                                    # :417-426


def plugin_least_squares_coefficients(inputs: List[Col], kw: OLSKwargs) -> Col:
    """least_squares_coefficients (src/expressions.rs:431-446) -> [k], NaN -> null (:137-139)."""
    pol = kw.null_policy
    is_valid = compute_is_valid_mask(inputs, pol)
    y, x = convert_to_ndarray(inputs, pol, is_valid)
    coef = get_least_squares_coefficients(y, x, kw)
    return coef, ~np.isnan(coef)


def _moving_inputs(inputs: List[Col], null_policy: str):
    is_valid = compute_is_valid_mask(inputs, null_policy)
    n = len(inputs[0][0])
    is_valid_vec = np.ones(n, dtype=bool) if is_valid is None else is_valid   # :230-244
    y, x = convert_to_ndarray(inputs, "zero", None)                           # always NullPolicy::Zero
    return is_valid, is_valid_vec, y, x


def plugin_recursive_least_squares_coefficients(inputs: List[Col], kw: RLSKwargs) -> Col:
    """recursive_least_squares_coefficients (src/expressions.rs:594-622) -> [n, k]."""
    _, iv, y, x = _moving_inputs(inputs, kw.null_policy)
    coef = solve_recursive_least_squares(y, x, kw.half_life, kw.initial_state_covariance, kw.initial_state_mean, iv)
    return coef, ~np.isnan(coef)


def plugin_recursive_least_squares(inputs: List[Col], kw: RLSKwargs) -> Col:
    """recursive_least_squares (src/expressions.rs:625-646): NOTE initial_state_mean is dropped (:636)."""
    is_valid, iv, y, x = _moving_inputs(inputs, kw.null_policy)
    coef = solve_recursive_least_squares(y, x, kw.half_life, kw.initial_state_covariance, None, iv)
    pred = (x * coef).sum(axis=1)                                            # :184
    return pred, is_valid                                                     # :640-645


def plugin_rolling_least_squares_coefficients(inputs: List[Col], kw: RollingKwargs) -> Col:
    """rolling_least_squares_coefficients (src/expressions.rs:649-676)."""
    _, iv, y, x = _moving_inputs(inputs, kw.null_policy)
    coef = solve_rolling_ols(y, x, kw.window_size, kw.min_periods, kw.use_woodbury, kw.alpha, iv, kw.null_policy)
    return coef, ~np.isnan(coef)


def plugin_rolling_least_squares(inputs: List[Col], kw: RollingKwargs) -> Col:
    """rolling_least_squares (src/expressions.rs:679-701)."""
    is_valid, iv, y, x = _moving_inputs(inputs, kw.null_policy)
    coef = solve_rolling_ols(y, x, kw.window_size, kw.min_periods, kw.use_woodbury, kw.alpha, iv, kw.null_policy)
    pred = (x * coef).sum(axis=1)
    return pred, is_valid


# ---------------------------------------------------------------------------------------------
# L4: _register_least_squares_plugin (polars_ols/least_squares.py:199-239) for ONE group
# ---------------------------------------------------------------------------------------------
def _finish_predictions(pred: Col, target: Col, sqrt_w: Optional[Col], mode: str, fill_nan: bool) -> Col:
    v, m = pred
    if sqrt_w is not None:
        inv = (np.asarray(1.0, dtype=sqrt_w[0].dtype) / sqrt_w[0])            # 1.0 / sqrt_w  (:235)
        v = v * inv.astype(np.float64)
    if mode == "residuals":
        v = target[0].astype(np.float64) - v                                  # target - predictions (:239)
        if target[1] is not None:
            m = target[1] if m is None else (m & target[1])
    if fill_nan:                                                               # fill_nan(None) (:407-408)
        nn = ~np.isnan(v)
        m = nn if m is None else (m & nn)
    return v, m


def least_squares(target, *features, sample_weights=None, add_intercept=False, mode="predictions",
                  kwargs: Optional[OLSKwargs] = None) -> Col:
    """compute_least_squares (polars_ols/least_squares.py:242-279) on a single group."""
    kw = kwargs or OLSKwargs()
    target = as_col(target)
    feats = [as_col(f) for f in features]
    w = None if sample_weights is None else as_col(sample_weights)
    t_fit, f_fit, sqrt_w = pre_process(target, feats, w, add_intercept)
    if mode == "coefficients":
        return plugin_least_squares_coefficients([t_fit, *f_fit], kw)
    return _finish_predictions(plugin_least_squares([t_fit, *f_fit], kw), target, sqrt_w, mode, False)


def recursive_least_squares(target, *features, sample_weights=None, add_intercept=False, mode="predictions",
                            kwargs: Optional[RLSKwargs] = None) -> Col:
    """compute_recursive_least_squares (polars_ols/least_squares.py:332-369) on a single group."""
    kw = kwargs or RLSKwargs()
    target = as_col(target)
    feats = [as_col(f) for f in features]
    w = None if sample_weights is None else as_col(sample_weights)
    t_fit, f_fit, sqrt_w = pre_process(target, feats, w, add_intercept)
    if mode == "coefficients":
        return plugin_recursive_least_squares_coefficients([t_fit, *f_fit], kw)
    return _finish_predictions(plugin_recursive_least_squares([t_fit, *f_fit], kw), target, sqrt_w, mode, False)


def rolling_least_squares(target, *features, sample_weights=None, add_intercept=False, mode="predictions",
                          kwargs: Optional[RollingKwargs] = None) -> Col:
    """compute_rolling_least_squares (polars_ols/least_squares.py:372-409) on a single group."""
    kw = kwargs or RollingKwargs()
    target = as_col(target)
    feats = [as_col(f) for f in features]
    w = None if sample_weights is None else as_col(sample_weights)
    t_fit, f_fit, sqrt_w = pre_process(target, feats, w, add_intercept)
    if mode == "coefficients":
        return plugin_rolling_least_squares_coefficients([t_fit, *f_fit], kw)
    return _finish_predictions(plugin_rolling_least_squares([t_fit, *f_fit], kw), target, sqrt_w, mode, True)


# ---------------------------------------------------------------------------------------------
# SURVEY §8f "next" rows: mode="statistics" (src/statistics.rs, src/expressions.rs:448-509) and
# multi_target_least_squares (src/expressions.rs:511-591, src/least_squares.rs:243-260)
# ---------------------------------------------------------------------------------------------
def students_t_two_sided_p(t: np.ndarray, df: float) -> np.ndarray:
    """t_value_to_p_value (src/statistics.rs:44-48) over statrs 0.17.1 StudentsT::cdf (third-party, not under
    /root/reference; its published formula): k = t, h = df / (df + k^2), ib = 0.5 * I_h(df/2, 1/2),
    cdf(x) = ib if x <= 0 else 1 - ib;  p = 2 * (1 - cdf(|t|)) — the 1 - (1 - ib) cancellation is kept."""
    from scipy.special import betainc
    t = np.abs(np.asarray(t, dtype=np.float64))
    with np.errstate(invalid="ignore", divide="ignore"):
        h = df / (df + t * t)
        ib = 0.5 * betainc(df / 2.0, 0.5, h)
        cdf = np.where(t <= 0.0, ib, 1.0 - ib)
        p = 2.0 * (1.0 - cdf)
    return np.where(np.isnan(t), np.nan, p)


def compute_residual_metrics(targets: np.ndarray, predicted: np.ndarray) -> dict:
    """compute_residual_metrics (src/statistics.rs:15-36)."""
    n = len(targets)
    mean = targets.mean() if n else 0.0
    err = targets - predicted
    with np.errstate(invalid="ignore", divide="ignore"):
        sse, sae, sst = float((err ** 2).sum()), float(np.abs(err).sum()), float(((targets - mean) ** 2).sum())
        return {"mse": np.float64(sse) / n, "mae": np.float64(sae) / n, "r2": 1.0 - np.float64(sse) / np.float64(sst)}


def compute_feature_metrics(features: np.ndarray, targets: np.ndarray, lam: float) -> dict:
    """compute_feature_metrics (src/statistics.rs:77-156): explicit Cholesky inverse of X^T X + lambda I."""
    n, p = features.shape
    xtx_reg = features.T @ features + lam * np.eye(p)
    nans = np.full(p, np.nan)
    try:
        if not np.isfinite(xtx_reg).all():
            raise np.linalg.LinAlgError
        L = np.linalg.cholesky(xtx_reg)                                   # faer cholesky(Lower), Err -> NaNs (:101-111)
    except np.linalg.LinAlgError:
        return {"standard_errors": nans, "t_values": nans.copy(), "p_values": nans.copy()}
    Li = np.linalg.inv(L)
    xtx_inv = Li.T @ Li
    coef = xtx_inv @ (features.T @ targets)                               # :116
    resid = targets - features @ coef
    rss = float((resid ** 2).sum())
    df = n - np.trace(xtx_inv) if lam > 0.0 else float(n - p)             # :125-129
    assert df > 0.0, "Degrees of freedom <= 0. Cannot compute standard errors."
    sigma2 = rss / df
    se = np.sqrt(sigma2 * np.abs(np.diag(xtx_inv)))
    with np.errstate(invalid="ignore", divide="ignore"):
        tv = coef / se
    return {"standard_errors": se, "t_values": tv, "p_values": students_t_two_sided_p(tv, df)}


def plugin_least_squares_statistics(inputs: List[Col], kw: OLSKwargs) -> dict:
    """least_squares_statistics (src/expressions.rs:469-509) for ONE group."""
    pol = kw.null_policy
    is_valid = compute_is_valid_mask(inputs, pol)
    y, x = convert_to_ndarray(inputs, pol, is_valid)
    lam = kw.alpha                                                        # kwargs.alpha.unwrap() (:475)
    assert lam is not None
    coef = get_least_squares_coefficients(y, x, kw)
    out = compute_residual_metrics(y, x @ coef)
    out.update(compute_feature_metrics(x, y, float(lam)))
    out["coefficients"] = coef
    return out


def least_squares_statistics(target, *features, sample_weights=None, add_intercept=False,
                             kwargs: Optional[OLSKwargs] = None) -> dict:
    """compute_least_squares(mode="statistics") (polars_ols/least_squares.py:213-224) on a single group."""
    kw = kwargs or OLSKwargs(alpha=0.0)
    target = as_col(target)
    feats = [as_col(f) for f in features]
    w = None if sample_weights is None else as_col(sample_weights)
    t_fit, f_fit, _ = pre_process(target, feats, w, add_intercept)
    return plugin_least_squares_statistics([t_fit, *f_fit], kw)


def solve_multi_target(y: np.ndarray, x: np.ndarray, alpha, rcond) -> np.ndarray:
    """solve_multi_target (src/least_squares.rs:243-260): always SVD; ridge-SVD when alpha > 0 (rcond honoured),
    else LAPACK dgelsd (rcond ignored)."""
    if x.size == 0:
        return np.zeros((x.shape[1], y.shape[1]))
    alpha = 0.0 if alpha is None else alpha
    if alpha > 0.0:
        return solve_ridge_svd(y, x, alpha, rcond)
    return np.linalg.lstsq(x, y, rcond=None)[0]


def plugin_multi_target_least_squares(targets: List[Col], features: List[Col], kw: OLSKwargs):
    """multi_target_least_squares (src/expressions.rs:521-591) for ONE group -> (pred [n, m], mask [n] or None).
    Every array is built with fill_zero = true (:549-550, :569): nulls become 0 even under 'ignore'."""
    pol = kw.null_policy
    m = len(targets)
    series = list(targets) + list(features)
    if pol in ("drop", "drop_zero", "drop_window"):                       # compute_is_valid_mask(.., Some(m)) (:201-228)
        is_valid = np.ones(len(series[0][0]), dtype=bool)
        for c in series:
            is_valid &= _is_valid(c)
    elif pol == "drop_y_zero_x":
        is_valid = np.ones(len(series[0][0]), dtype=bool)
        for c in targets:
            is_valid &= _is_valid(c)
    else:
        is_valid = None
    sel = slice(None) if is_valid is None else is_valid                    # handle_nulls (:257-296) then fill_zero
    x_fit = np.ascontiguousarray(np.stack([_to_f64(c, 0.0)[sel] for c in features], axis=1))
    y_fit = np.ascontiguousarray(np.stack([_to_f64(c, 0.0)[sel] for c in targets], axis=1))
    coef = solve_multi_target(y_fit, x_fit, kw.alpha, kw.rcond)           # [k, m]
    if pol in ("ignore", "zero"):
        return x_fit @ coef, None
    pred = features_zero_filled(features) @ coef
    return pred, (is_valid if pol == "drop" else None)


def multi_target_least_squares(targets: Sequence, *features, sample_weights=None, add_intercept=False,
                               mode="predictions", kwargs: Optional[OLSKwargs] = None):
    """compute_multi_target_least_squares (polars_ols/least_squares.py:282-329) on a single group.
    Returns (values [n, m], validity [n, m]): NaN -> null (convert_array_to_struct_series :137-139)."""
    kw = kwargs or OLSKwargs()
    assert not kw.positive and (kw.l1_ratio is None or kw.l1_ratio == 0.0)
    assert kw.solve_method in ("svd", None)
    if mode == "coefficients":
        raise NotImplementedError("Only mode={'predictions', 'residuals'} is currently supported.")
    tg = [as_col(t) for t in targets]
    feats = [as_col(f) for f in features]
    w = None if sample_weights is None else as_col(sample_weights)
    if add_intercept:
        feats = feats + [(np.ones(len(tg[0][0]), dtype=tg[0][0].dtype), None)]
    sqrt_w = None
    t_fit, f_fit = tg, feats
    if w is not None:
        _, f_fit, sqrt_w = pre_process(tg[0], feats, w, False)
        t_fit = [_mul(t, sqrt_w) for t in tg]
    pred, row_mask = plugin_multi_target_least_squares(t_fit, f_fit, kw)
    vals, masks = [], []
    for j, t in enumerate(tg):
        pm = ~np.isnan(pred[:, j])
        if row_mask is not None:
            pm &= row_mask
        v, mk = _finish_predictions((pred[:, j], pm), t, sqrt_w, mode, False)
        vals.append(v)
        masks.append(np.ones(len(v), dtype=bool) if mk is None else mk)
    return np.stack(vals, axis=1), np.stack(masks, axis=1)


def predict(coefficients: Sequence, features: Sequence, null_policy: str = "zero", add_intercept: bool = False) -> Col:
    """predict (src/expressions.rs:706-741 + polars_ols/least_squares.py:455-491): rowwise
    (features * coefficients).sum(1); features zero-filled unless 'ignore'; 'drop' masks rows with any null input."""
    coefs = [as_col(c) for c in coefficients]
    feats = [as_col(f) for f in features]
    if add_intercept:
        feats = feats + [(np.ones(len(coefs[0][0])), None)]
    assert len(coefs) == len(feats), "number of coefficients must match number of features!"
    fill = np.nan if null_policy == "ignore" else 0.0
    x = np.stack([_to_f64(f, fill) for f in feats], axis=1)
    c = np.stack([_to_f64(cc, np.nan) for cc in coefs], axis=1)
    pred = (x * c).sum(axis=1)
    if null_policy == "drop":
        m = np.ones(len(pred), dtype=bool)
        for col_ in coefs + feats:
            m &= _is_valid(col_)
        return pred, m
    return pred, None


# ---------------------------------------------------------------------------------------------
# polars `.over(group)`: split rows by key (order preserved inside a group), one call per group,
# scatter back.  Static coefficients (returns_scalar) are broadcast to the group's rows.
# ---------------------------------------------------------------------------------------------
def over(fn, group_ids: np.ndarray, target, *features, sample_weights=None, per_group: bool = False, **kw):
    """Apply one of the three functions above per group.

    per_group=False -> polars `.over()` shape: [N] (or [N, k] for coefficients).
    per_group=True  -> `group_by(..).agg(..)` shape for static coefficients: (keys, [G, k]).
    """
    group_ids = np.asarray(group_ids)
    target = as_col(target)
    feats = [as_col(f) for f in features]
    w = None if sample_weights is None else as_col(sample_weights)
    keys, inv = np.unique(group_ids, return_inverse=True)
    order = np.argsort(inv, kind="stable")
    counts = np.bincount(inv, minlength=len(keys))
    offs = np.concatenate([[0], np.cumsum(counts)])
    n = len(group_ids)
    out_v = out_m = None
    per_v, per_m = [], []

    def take(c: Col, idx):
        return c[0][idx], (None if c[1] is None else c[1][idx])

    for g in range(len(keys)):
        idx = order[offs[g]:offs[g + 1]]
        v, m = fn(take(target, idx), *[take(f, idx) for f in feats],
                  sample_weights=None if w is None else take(w, idx), **kw)
        if m is None:
            m = np.ones(v.shape, dtype=bool)
        if per_group:
            per_v.append(v)
            per_m.append(m)
            continue
        if v.ndim == 1 and len(v) != len(idx):       # static coefficients: broadcast to rows
            v = np.broadcast_to(v, (len(idx), len(v)))
            m = np.broadcast_to(m, v.shape)
        if out_v is None:
            out_v = np.full((n,) + v.shape[1:], np.nan)
            out_m = np.zeros((n,) + v.shape[1:], dtype=bool)
        out_v[idx] = v
        out_m[idx] = m
    if per_group:
        return keys, np.stack(per_v), np.stack(per_m)
    return out_v, out_m
