"""Host-side mirror of ``polars_ols/least_squares.py`` + ``polars_ols/__init__.py`` of the reference.

Same names, argument meaning, defaults and error behaviour:
``OLSKwargs / RLSKwargs / RollingKwargs`` (reference ``polars_ols/least_squares.py:66-160``),
``compute_least_squares / compute_recursive_least_squares / compute_rolling_least_squares`` (``:242-409``)
and the ``least_squares`` expression namespace ``LeastSquares`` (``polars_ols/__init__.py:35-295``:
``ols, wls, ridge, lasso, elastic_net, rls, rolling_ols, expanding_ols, least_squares``).

polars is not available in this image, so expressions are evaluated against a minimal ``Frame``
(dict of numpy arrays / pyarrow arrays / CUDA torch tensors) instead of a ``pl.DataFrame``:

    df = Frame({"y": y, "x1": x1, "x2": x2, "group": g})
    out = df.select(col("y").least_squares.ridge(col("x1"), col("x2"), alpha=1e-3, mode="coefficients").over("group"))

What differs from the reference by design: `.over()` is evaluated as ONE batched call into the CUDA
engine (all groups at once) instead of one plugin call per group, and the WLS pre/post-processing
(``_pre_process_data`` / ``predictions *= 1/sqrt_w`` / ``target - predictions``, reference
``polars_ols/least_squares.py:163-239``) is fused into the kernels rather than run as extra column passes.
There is no CPU path: without a CUDA device evaluation raises.
"""
from __future__ import annotations

import copy
import ctypes as C
import logging
import math
from dataclasses import asdict, dataclass
from typing import Any, Dict, List, Literal, Optional, Sequence, Set, Tuple, Union, get_args

import numpy as np

from . import _lib as L
from .engine import Batch, Col, Engine, as_col, get_engine, nan_or, _is_torch


def torch_as_device(a: np.ndarray, device):
    import torch
    return torch.as_tensor(a, device=device)


def torch_index(idx: np.ndarray, like):
    import torch
    return torch.as_tensor(idx, device=like.device)

logger = logging.getLogger(__name__)

__all__ = [
    "compute_least_squares", "compute_recursive_least_squares", "compute_rolling_least_squares",
    "compute_multi_target_least_squares", "OLSKwargs", "RLSKwargs", "RollingKwargs", "NullPolicy", "OutputMode", "SolveMethod",
    "LeastSquares", "Frame", "col", "Expr", "Result", "predict", "PredictExpr",
]

NullPolicy = Literal["zero", "drop", "ignore", "drop_zero", "drop_y_zero_x", "drop_window"]
OutputMode = Literal["predictions", "residuals", "coefficients", "statistics"]
SolveMethod = Literal["qr", "svd", "chol", "lu", "cd", "cd_active_set"]

_VALID_NULL_POLICIES: Set[str] = set(get_args(NullPolicy))
_VALID_OUTPUT_MODES: Set[str] = set(get_args(OutputMode))
_VALID_SOLVE_METHODS: Set[Optional[str]] = set(get_args(SolveMethod)).union({None})


# ------------------------------------------------------------------------------------------------
# kwargs dataclasses — field for field the reference's (polars_ols/least_squares.py:66-160)
# ------------------------------------------------------------------------------------------------
@dataclass
class Kwargs:
    null_policy: NullPolicy = "ignore"

    def to_dict(self) -> Dict[str, Any]:
        return asdict(self)

    def __post_init__(self):
        assert (
            self.null_policy in _VALID_NULL_POLICIES
        ), f"'null_policy' must be one of {_VALID_NULL_POLICIES}. You passed: {self.null_policy}"


@dataclass
class OLSKwargs(Kwargs):
    alpha: Optional[float] = 0.0
    l1_ratio: Optional[float] = None
    max_iter: Optional[int] = 1_000
    tol: Optional[float] = 1.0e-5
    positive: Optional[bool] = False
    solve_method: Optional[SolveMethod] = None
    rcond: Optional[float] = None

    def __post_init__(self):
        valid_ols_policies = _VALID_NULL_POLICIES - {"drop_window"}
        assert (
            self.null_policy in valid_ols_policies
        ), f"'null_policy' must be one of {valid_ols_policies}. You passed: {self.null_policy}"
        assert (
            self.solve_method in _VALID_SOLVE_METHODS
        ), f"'solve_method' must be one of {_VALID_SOLVE_METHODS}. You passed: {self.solve_method}"

    def to_c(self) -> L.OLSKwargs:
        return L.OLSKwargs(nan_or(self.alpha), nan_or(self.l1_ratio), -1 if self.max_iter is None else int(self.max_iter),
                           nan_or(self.tol), 1 if self.positive else 0, L.SOLVE_METHOD[self.solve_method],
                           L.NULL_POLICY[self.null_policy], 0, nan_or(self.rcond))


@dataclass
class RLSKwargs(Kwargs):
    half_life: Optional[float] = None
    initial_state_covariance: Optional[float] = 10.0
    initial_state_mean: Union[Optional[List[float]], float] = None
    null_policy: NullPolicy = "drop"


@dataclass
class RollingKwargs(Kwargs):
    window_size: int = 1_000_000
    min_periods: Optional[int] = None
    use_woodbury: Optional[bool] = None
    alpha: Optional[float] = None
    null_policy: NullPolicy = "drop_window"

    def to_c(self) -> L.RollingKwargs:
        return L.RollingKwargs(int(self.window_size), -1 if self.min_periods is None else int(self.min_periods),
                               -1 if self.use_woodbury is None else int(bool(self.use_woodbury)),
                               L.NULL_POLICY[self.null_policy], nan_or(self.alpha))


# ------------------------------------------------------------------------------------------------
# results
# ------------------------------------------------------------------------------------------------
@dataclass
class Result:
    """Output of one expression.  ``values`` is f64; ``valid`` (uint8, same shape, or None = all valid)
    marks polars nulls.  Static coefficients carry one row per group (``keys``) plus the row->group map
    needed to broadcast them the way ``.over()`` does."""
    name: str
    values: Any
    valid: Any = None
    fields: Optional[List[str]] = None      # coefficient struct field names
    keys: Optional[np.ndarray] = None       # group keys (per-group results)
    group_of_row: Any = None                # [N] group id of every row, or a zero-argument callable producing it

    def to_numpy(self, broadcast: bool = False) -> np.ndarray:
        v = self.values.cpu().numpy() if _is_torch(self.values) else np.asarray(self.values)
        v = v.copy()
        if self.valid is not None:
            m = self.valid.cpu().numpy() if _is_torch(self.valid) else np.asarray(self.valid)
            v[m == 0] = np.nan
        elif self.fields is not None:
            pass  # coefficient NaN <=> null already
        if broadcast and self.group_of_row is not None:
            if callable(self.group_of_row):          # device plan: the row -> group map is only built when asked for
                self.group_of_row = self.group_of_row()
            v = v[self.group_of_row]
        return v

    def to_struct(self) -> Dict[str, np.ndarray]:
        """struct-valued results as {field: array}: multi-target predictions / residuals ({target: [n]}, nulls as NaN)
        and mode="statistics" ({r2, mae, mse: [G]; feature_names: list; coefficients, ...: [G, k]})."""
        if isinstance(self.values, dict):
            return {k: (v.cpu().numpy() if _is_torch(v) else v) for k, v in self.values.items()}
        v = self.to_numpy()
        return {f: v[j] for j, f in enumerate(self.fields)}

    def is_null(self) -> np.ndarray:
        if self.valid is not None:
            m = self.valid.cpu().numpy() if _is_torch(self.valid) else np.asarray(self.valid)
            return m == 0
        v = self.to_numpy()
        return np.isnan(v) if self.fields is not None else np.zeros(v.shape, dtype=bool)


# ------------------------------------------------------------------------------------------------
# grouping: polars' `.over()` / group_by split (reference: 3rd-party polars engine)
# ------------------------------------------------------------------------------------------------
def _group_plan(keys: Sequence[np.ndarray]):
    """HOST statement of the group plan: keys -> (unique keys, offsets [G+1], row_index [N] or None if groups are
    contiguous slices, group_of_row [N]).  The product plans on the device (`Engine.group_plan`, group_plan.cuh); this
    numpy version is the specification the device plan is tested against (bit-identical offsets and permutation)."""
    if len(keys) == 1:
        k = np.asarray(keys[0])
        uniq, inv = np.unique(k, return_inverse=True)
    else:
        stacked = np.rec.fromarrays([np.asarray(k) for k in keys])
        uniq, inv = np.unique(stacked, return_inverse=True)
    inv = inv.reshape(-1)
    counts = np.bincount(inv, minlength=len(uniq))
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    if len(inv) == 0 or np.all(inv[1:] >= inv[:-1]):
        return uniq, offsets, None, inv           # already sorted & contiguous: GroupsSlice
    order = np.argsort(inv, kind="stable").astype(np.int64)  # row order inside a group is preserved
    return uniq, offsets, order, inv


# ------------------------------------------------------------------------------------------------
# expressions
# ------------------------------------------------------------------------------------------------
class Expr:
    """A column reference (``col("x")``) or a literal buffer."""

    def __init__(self, name: Optional[str] = None, data=None):
        self._name = name
        self._data = data

    @property
    def output_name(self) -> str:
        return self._name or ""

    def resolve(self, frame: "Frame") -> Col:
        if self._data is not None:
            return as_col(self._data)
        return as_col(frame[self._name])

    @property
    def least_squares(self) -> "LeastSquares":
        return LeastSquares(self)


def col(name: str) -> Expr:
    return Expr(name)


class ProductExpr(Expr):
    """Interaction term of a formula (``x3:x4``): the element-wise product of columns, named ``"x3:x4"`` as the
    reference's ``reduce(lambda x, y: x * pl.col(y), factors, pl.lit(1)).alias(":".join(factors))``
    (polars_ols/utils.py:108-110).  A null in any factor is a null in the product."""

    def __init__(self, factors: Sequence[str]):
        super().__init__(":".join(factors))
        self._factors = list(factors)

    def resolve(self, frame: "Frame") -> Col:
        cols = [as_col(frame[f]) for f in self._factors]
        values = cols[0].values
        for c in cols[1:]:
            values = values * c.values
        validity = None
        for c in cols:
            if c.validity is not None:
                validity = c.validity if validity is None else (validity & c.validity)
        return Col(values, validity)


def build_expressions_from_patsy_formula(formula: str, include_dependent_variable: bool = False) -> Tuple[List[Expr], bool]:
    """Subset of patsy formulas the reference supports (polars_ols/utils.py:62-111): ``y ~ x1 + x2:x3 - 1`` — plain
    columns, ``:`` interactions and the intercept switch.  patsy is not required: the grammar is parsed here.  As in
    the reference, functions of columns (``log(x1)``) and categoricals (``C(group)``) are not supported.  The intercept
    switch is the reference's literal test ``"-1" not in formula`` on the RAW text (utils.py:99): ``"y ~ x1 -1"`` drops the
    ``const`` column, ``"y ~ x1 - 1"`` (with a space) keeps it — a quirk, copied, not fixed.  Repeated terms are kept
    once, as patsy's ``ModelDesc`` does."""
    text = formula.replace(" ", "")
    if "~" in text:
        lhs, rhs = text.split("~", 1)
    else:
        lhs, rhs = "", text
    if include_dependent_variable:
        assert lhs and "+" not in lhs and "-" not in lhs, "must provide exactly one LHS variable"
    else:
        assert not lhs, "can not provide LHS variables in this context"
    add_intercept = "-1" not in formula
    exprs: List[Expr] = []
    seen: Set[str] = set()
    if include_dependent_variable:
        exprs.append(Expr(lhs))
    for term in rhs.replace("-", "+-").split("+"):
        if term in ("", "1", "0", "-1", "-0"):
            continue
        if "C(" in term:
            raise NotImplementedError("building patsy categories into polars expressions is not supported")
        if any(ch in term for ch in "()*/^") or term.startswith("-"):
            raise NotImplementedError(f"formula term {term!r}: only columns, ':' interactions and the intercept are supported")
        factors = term.split(":")
        key = ":".join(sorted(factors))          # patsy: a term is a set of factors, listed once
        if key in seen:
            continue
        seen.add(key)
        exprs.append(Expr(factors[0]) if len(factors) == 1 else ProductExpr(factors))
    return exprs, add_intercept


ExprOrStr = Union[Expr, str, Any]


def parse_into_expr(e: ExprOrStr) -> Expr:
    """reference polars_ols/utils.py:21-58: strings are column names."""
    if isinstance(e, Expr):
        return e
    if isinstance(e, str):
        return Expr(e)
    return Expr(None, e)


class LsExpr:
    """A pending least-squares expression (what ``register_plugin_function`` returns in the reference)."""

    def __init__(self, kind: str, target: Expr, features: List[Expr], sample_weights: Optional[Expr],
                 add_intercept: bool, mode: str, kwargs: Kwargs):
        self.kind = kind
        self.target, self.features, self.sample_weights = target, features, sample_weights
        self.add_intercept, self.mode, self.kwargs = add_intercept, mode, kwargs
        self._over: List[ExprOrStr] = []
        self._alias: Optional[str] = None
        self._per_group = False

    def over(self, *keys: ExprOrStr) -> "LsExpr":
        new = copy.copy(self)                  # expressions are immutable, as in polars: `e.over(..)` leaves `e` ungrouped
        new._over = list(keys)
        return new

    def alias(self, name: str) -> "LsExpr":
        new = copy.copy(self)
        new._alias = name
        return new

    @property
    def output_name(self) -> str:
        if self._alias:
            return self._alias
        # reference: coefficients / statistics are aliased to the mode (least_squares.py:224);
        # predictions keep the target's name (src/expressions.rs:404)
        if self.kind == "multi_target_least_squares":     # Field::new("predictions", ..) src/expressions.rs:518
            return "predictions" if self.mode == "predictions" else (self.target.output_name or self.mode)
        return self.mode if self.mode in ("coefficients", "statistics") else self.target.output_name or self.mode

    # -- evaluation ----------------------------------------------------------------------------------
    def rls_c_kwargs(self, n_coef: int, initial_information: Optional[np.ndarray] = None):
        """(ctypes RLSKwargs, buffers to keep alive).  `initial_information` [G, n_coef^2 + n_coef] continues each
        series from the information state of an earlier time shard (parallel.time_sharded)."""
        kw: RLSKwargs = self.kwargs
        mean_arr = None
        if kw.initial_state_mean is not None:
            mean_arr = np.ascontiguousarray(np.broadcast_to(np.asarray(kw.initial_state_mean, dtype=np.float64), (n_coef,)))
        info = None if initial_information is None else np.ascontiguousarray(initial_information, dtype=np.float64)
        ckw = L.RLSKwargs(nan_or(kw.half_life), nan_or(kw.initial_state_covariance),
                          None if mean_arr is None else mean_arr.ctypes.data, L.NULL_POLICY[kw.null_policy], 0,
                          None if info is None else info.ctypes.data)
        return ckw, (mean_arr, info)

    def target_struct(self, frame: "Frame"):
        """multi-target: the target is a struct column, here {field: column} -> [(field, Col)]"""
        t = self.target._data if self.target._data is not None else frame[self.target._name]
        assert isinstance(t, dict) and len(t) > 0, (
            "the first series in a multi-target regression must be of polars struct dtype with each field "
            "corresponding to an output")                 # src/expressions.rs:513-517
        return [(str(k), as_col(v)) for k, v in t.items()]

    def batch(self, frame: "Frame"):
        """(Batch of the resolved input columns, coefficient field names) — no grouping yet."""
        target = self.target_struct(frame)[0][1] if self.kind == "multi_target_least_squares" else self.target.resolve(frame)
        feats = [f.resolve(frame) for f in self.features]
        names = [f.output_name or str(i) for i, f in enumerate(self.features)]  # src/expressions.rs:126-130
        add_intercept = self.add_intercept
        if add_intercept and any(n == "const" for n in names):
            logger.info("feature named 'const' already detected, assuming it is an intercept")  # least_squares.py:185-186
            add_intercept = False
        if add_intercept:
            names = names + ["const"]
        weights = self.sample_weights.resolve(frame) if self.sample_weights is not None else None
        return Batch(target, feats, weights, add_intercept), names

    def evaluate(self, frame: "Frame", engine: Optional[Engine] = None,
                 initial_information: Optional[np.ndarray] = None) -> Result:
        b, names = self.batch(frame)
        target = b.target
        keys = group_of_row = None
        if engine is None:
            dev = target.values.device.index if target.is_device else 0
            engine = get_engine(dev or 0, torch_stream=target.is_device)
        if self._over:
            # `.over()`: keys -> CSR groups on the device (b200ols_group_plan_build); only the [G+1] offsets come back
            key_arrays = []
            for k in self._over:
                kv = frame[k] if isinstance(k, str) else parse_into_expr(k).resolve(frame).values
                if _is_torch(kv) and kv.is_cuda != target.is_device:
                    kv = kv.cpu().numpy() if kv.is_cuda else kv.to(target.values.device)
                elif not _is_torch(kv) and target.is_device:
                    kv = torch_as_device(np.asarray(kv), target.values.device)
                key_arrays.append(kv)
            plan = engine.group_plan(key_arrays)
            b.offsets, b.row_index = plan.offsets, plan.row_index
            def group_of_row(plan=plan, key_arrays=key_arrays, engine=engine):
                # built on demand (the `.over()` broadcast); a later plan on the engine replaced the device tables -> re-plan
                cur = plan if engine._plan_serial == plan.serial else engine.group_plan(key_arrays)
                return cur.group_of_row()
            if self.mode in ("coefficients", "statistics") and self.kind == "least_squares":
                first = plan.first_row        # one representative row per group -> its key values
                ks = [(kv[torch_index(first, kv)].cpu().numpy() if _is_torch(kv) else np.asarray(kv)[first]) for kv in key_arrays]
                keys = ks[0] if len(ks) == 1 else np.rec.fromarrays(ks)
        if self.kind == "multi_target_least_squares":
            tcols = self.target_struct(frame)
            b.target = tcols[0][1]
            v, m = engine.multi_target_least_squares(b, [c_ for _, c_ in tcols], self.kwargs.to_c(), L.MODE[self.mode])
            return Result(self.output_name, v, m, [n_ for n_, _ in tcols])
        if self.mode == "statistics":
            st = engine.least_squares_statistics(b, self.kwargs.to_c())
            st["feature_names"] = list(names)            # src/expressions.rs:483-484
            order = ("r2", "mae", "mse", "feature_names", "coefficients", "standard_errors", "t_values", "p_values")
            return Result(self.output_name, {k_: st[k_] for k_ in order}, None, list(order), keys, group_of_row)
        mode = L.MODE[self.mode]
        if self.kind == "least_squares":
            v, m = engine.least_squares(b, self.kwargs.to_c(), mode)
            if self.mode == "coefficients":
                return Result(self.output_name, v, None, names, keys, group_of_row)
            return Result(self.output_name, v, m)
        if self.kind == "recursive_least_squares":
            ckw, keep = self.rls_c_kwargs(len(names), initial_information)
            v, m = engine.recursive_least_squares(b, ckw, mode, keep)
        else:
            v, m = engine.rolling_least_squares(b, self.kwargs.to_c(), mode)
        if self.mode == "coefficients":
            return Result(self.output_name, v, m, names)
        return Result(self.output_name, v, m)


class PredictExpr:
    """Pending `predict` expression (reference polars_ols/least_squares.py:455-491)."""

    def __init__(self, coefficients: Expr, features: List[Expr], null_policy: str, add_intercept: bool, name: Optional[str]):
        self.coefficients, self.features = coefficients, features
        self.null_policy, self.add_intercept, self.output_name = null_policy, add_intercept, name or "predictions"

    def alias(self, name: str) -> "PredictExpr":
        new = copy.copy(self)
        new.output_name = name
        return new

    def evaluate(self, frame: "Frame", engine: Optional[Engine] = None) -> Result:
        coef = self.coefficients._data if self.coefficients._data is not None else frame[self.coefficients._name]
        if isinstance(coef, Result):
            coef = coef.to_numpy(broadcast=True)
        if isinstance(coef, dict):                      # struct column as {field: array}
            cols = [as_col(v) for v in coef.values()]
        elif _is_torch(coef):
            cols = [as_col(coef[:, j].contiguous()) for j in range(coef.shape[1])]
        else:
            a = np.asarray(coef, dtype=np.float64)      # [N, k]; NaN <=> null field (fill_nan(NULL) at src/expressions.rs:137-139)
            cols = [Col(np.ascontiguousarray(a[:, j])) for j in range(a.shape[1])]
            cols = [Col(c.values, None if not np.isnan(c.values).any() else np.packbits(~np.isnan(c.values), bitorder="little")) for c in cols]
        feats = [f.resolve(frame) for f in self.features]
        add_intercept = self.add_intercept
        if add_intercept and any(f.output_name == "const" for f in self.features):
            logger.warning("feature named 'const' already detected, assuming it is the intercept")
            add_intercept = False
        assert len(cols) == len(feats) + (1 if add_intercept else 0), "number of coefficients must match number of features!"
        if engine is None:
            dev = cols[0].values.device.index if cols[0].is_device else 0
            engine = get_engine(dev or 0, torch_stream=cols[0].is_device)
        v, m = engine.predict(cols, feats, add_intercept, L.NULL_POLICY[self.null_policy])
        return Result(self.output_name, v, m)


def predict(coefficients: ExprOrStr, *features: ExprOrStr, null_policy: NullPolicy = "zero", add_intercept: bool = False,
            name: Optional[str] = None) -> PredictExpr:
    """reference polars_ols/least_squares.py:455-491"""
    assert null_policy in _VALID_NULL_POLICIES, "'null_policy' must be one of {drop, ignore, zero}"
    return PredictExpr(parse_into_expr(coefficients), [parse_into_expr(f) for f in features], null_policy, add_intercept, name)


class Frame(dict):
    """Minimal stand-in for ``pl.DataFrame``: a dict of equal-length columns (numpy arrays, pyarrow arrays,
    ``(values, valid_mask)`` pairs, numpy masked arrays or CUDA torch tensors)."""

    def select(self, *exprs: LsExpr, engine: Optional[Engine] = None) -> Dict[str, Result]:
        out: Dict[str, Result] = {}
        for e in exprs:
            out[e.output_name] = e.evaluate(self, engine)
        return out

    def with_columns(self, *exprs: LsExpr, engine: Optional[Engine] = None) -> "Frame":
        new = Frame(self)
        for e in exprs:
            r = e.evaluate(self, engine)
            new[e.output_name] = r.to_numpy(broadcast=True)
        return new


# ------------------------------------------------------------------------------------------------
# compute_* functions (reference polars_ols/least_squares.py:242-409)
# ------------------------------------------------------------------------------------------------
def _make(kind, target, features, sample_weights, add_intercept, mode, kwargs) -> LsExpr:
    return LsExpr(kind, parse_into_expr(target), [parse_into_expr(f) for f in features],
                  None if sample_weights is None else parse_into_expr(sample_weights), add_intercept, mode, kwargs)


def compute_least_squares(target: ExprOrStr, *features: ExprOrStr, sample_weights: Optional[ExprOrStr] = None,
                          add_intercept: bool = False, mode: OutputMode = "predictions",
                          ols_kwargs: Optional[OLSKwargs] = None) -> LsExpr:
    assert mode in _VALID_OUTPUT_MODES, f"'mode' must be one of {_VALID_OUTPUT_MODES}"
    return _make("least_squares", target, features, sample_weights, add_intercept, mode, ols_kwargs or OLSKwargs())


def compute_multi_target_least_squares(targets: ExprOrStr, *features: ExprOrStr, sample_weights: Optional[ExprOrStr] = None,
                                       add_intercept: bool = False, mode: OutputMode = "predictions",
                                       ols_kwargs: Optional[OLSKwargs] = None) -> LsExpr:
    """reference polars_ols/least_squares.py:282-329: `targets` is a struct column ({field: array} in a Frame)."""
    ols_kwargs = ols_kwargs or OLSKwargs()
    multi_target_conditions = not ols_kwargs.positive and (ols_kwargs.l1_ratio is None or ols_kwargs.l1_ratio == 0.0)
    msg = "Consider running multiple independent regressions on a multi-expression target!"
    assert multi_target_conditions, (
        "Multi-target regression is only supported for unconstrained OLS & Ridge problems." + msg)
    assert ols_kwargs.solve_method in {"svd", None}, "only solve_method='svd' is supported for multi-target regressions"
    if mode == "coefficients":
        raise NotImplementedError("Only mode={'predictions', 'residuals'} is currently supported. " + msg)
    assert mode in ("predictions", "residuals"), f"'mode' must be one of {_VALID_OUTPUT_MODES}"
    return _make("multi_target_least_squares", targets, features, sample_weights, add_intercept, mode, ols_kwargs)


def compute_recursive_least_squares(target: ExprOrStr, *features: ExprOrStr, sample_weights: Optional[ExprOrStr] = None,
                                    add_intercept: bool = False, mode: OutputMode = "predictions",
                                    rls_kwargs: Optional[RLSKwargs] = None) -> LsExpr:
    valid_output_modes = _VALID_OUTPUT_MODES - {"statistics"}
    assert mode in valid_output_modes, f"'mode' must be one of {valid_output_modes}"
    return _make("recursive_least_squares", target, features, sample_weights, add_intercept, mode, rls_kwargs or RLSKwargs())


def compute_rolling_least_squares(target: ExprOrStr, *features: ExprOrStr, sample_weights: Optional[ExprOrStr] = None,
                                  add_intercept: bool = False, mode: OutputMode = "predictions",
                                  rolling_kwargs: Optional[RollingKwargs] = None) -> LsExpr:
    valid_output_modes = _VALID_OUTPUT_MODES - {"statistics"}
    assert mode in valid_output_modes, f"'mode' must be one of {valid_output_modes}"
    return _make("rolling_least_squares", target, features, sample_weights, add_intercept, mode,
                 rolling_kwargs or RollingKwargs())


def compute_least_squares_from_formula(formula: str, sample_weights: Optional[ExprOrStr] = None,
                                       mode: OutputMode = "predictions", **kwargs) -> LsExpr:
    """reference polars_ols/least_squares.py:412-452: `half_life` -> rls, `window_size` -> rolling, else static."""
    expressions, add_intercept = build_expressions_from_patsy_formula(formula, include_dependent_variable=True)
    if kwargs.get("half_life"):
        return compute_recursive_least_squares(expressions[0], *expressions[1:], add_intercept=add_intercept,
                                               sample_weights=sample_weights, mode=mode, rls_kwargs=RLSKwargs(**kwargs))
    if kwargs.get("window_size"):
        return compute_rolling_least_squares(expressions[0], *expressions[1:], add_intercept=add_intercept,
                                             sample_weights=sample_weights, mode=mode, rolling_kwargs=RollingKwargs(**kwargs))
    return compute_least_squares(expressions[0], *expressions[1:], add_intercept=add_intercept,
                                 sample_weights=sample_weights, mode=mode, ols_kwargs=OLSKwargs(**kwargs))


# ------------------------------------------------------------------------------------------------
# the `least_squares` namespace (reference polars_ols/__init__.py:35-295)
# ------------------------------------------------------------------------------------------------
class LeastSquares:
    def __init__(self, expr: Expr):
        self._expr = expr

    def least_squares(self, *features: ExprOrStr, sample_weights: Optional[ExprOrStr] = None, add_intercept: bool = False,
                      mode: OutputMode = "predictions", null_policy: NullPolicy = "ignore",
                      solve_method: Optional[SolveMethod] = None, multi_target: bool = False, **ols_kwargs) -> LsExpr:
        ols_func = compute_least_squares if not multi_target else compute_multi_target_least_squares  # __init__.py:90
        return ols_func(self._expr, *features, sample_weights=sample_weights, add_intercept=add_intercept,
                                     mode=mode,
                                     ols_kwargs=OLSKwargs(null_policy=null_policy, solve_method=solve_method, **ols_kwargs))

    def ols(self, *features: ExprOrStr, **kwargs) -> LsExpr:
        return self.least_squares(*features, **kwargs)

    def multi_target_ols(self, *features: ExprOrStr, **kwargs) -> LsExpr:   # polars_ols/__init__.py:104-105
        return self.least_squares(*features, multi_target=True, **kwargs)

    def wls(self, *features: ExprOrStr, sample_weights: ExprOrStr, **kwargs) -> LsExpr:
        return self.least_squares(*features, sample_weights=sample_weights, **kwargs)

    def ridge(self, *features: ExprOrStr, alpha: float, **kwargs) -> LsExpr:
        return self.least_squares(*features, alpha=alpha, l1_ratio=0.0, **kwargs)

    def lasso(self, *features: ExprOrStr, alpha: float, **kwargs) -> LsExpr:
        return self.least_squares(*features, alpha=alpha, l1_ratio=1.0, **kwargs)

    def elastic_net(self, *features: ExprOrStr, alpha: float, l1_ratio: float = 0.5, positive: bool = False, **kwargs) -> LsExpr:
        return self.least_squares(*features, alpha=alpha, l1_ratio=l1_ratio, positive=positive, **kwargs)

    def rls(self, *features: ExprOrStr, sample_weights: Optional[ExprOrStr] = None, add_intercept: bool = False,
            mode: OutputMode = "predictions", null_policy: NullPolicy = "drop", half_life: Optional[float] = None,
            initial_state_covariance: Optional[float] = 10.0,
            initial_state_mean: Union[Optional[List[float]], float] = None) -> LsExpr:
        return compute_recursive_least_squares(
            self._expr, *features, sample_weights=sample_weights, add_intercept=add_intercept, mode=mode,
            rls_kwargs=RLSKwargs(null_policy=null_policy, half_life=half_life, initial_state_mean=initial_state_mean,
                                 initial_state_covariance=initial_state_covariance))

    def rolling_ols(self, *features: ExprOrStr, window_size: int, sample_weights: Optional[ExprOrStr] = None,
                    add_intercept: bool = False, mode: OutputMode = "predictions", null_policy: NullPolicy = "drop",
                    min_periods: Optional[int] = None, use_woodbury: Optional[bool] = None,
                    alpha: Optional[float] = None) -> LsExpr:
        return compute_rolling_least_squares(
            self._expr, *features, sample_weights=sample_weights, add_intercept=add_intercept, mode=mode,
            rolling_kwargs=RollingKwargs(window_size=window_size, min_periods=min_periods, use_woodbury=use_woodbury,
                                         alpha=alpha, null_policy=null_policy))

    def expanding_ols(self, *features: ExprOrStr, **kwargs) -> LsExpr:
        return self.rls(*features, half_life=None, **kwargs)

    def from_formula(self, formula: str, **kwargs) -> LsExpr:
        """reference polars_ols/__init__.py:263-272"""
        features, add_intercept = build_expressions_from_patsy_formula(formula, include_dependent_variable=False)
        if kwargs.get("half_life"):
            return self.rls(*features, add_intercept=add_intercept, **kwargs)
        if kwargs.get("window_size"):
            return self.rolling_ols(*features, add_intercept=add_intercept, **kwargs)
        return self.least_squares(*features, add_intercept=add_intercept, **kwargs)

    def predict(self, *features: ExprOrStr, name: Optional[str] = None, add_intercept: bool = False,
                null_policy: NullPolicy = "zero") -> "PredictExpr":
        return predict(self._expr, *features, add_intercept=add_intercept, name=name, null_policy=null_policy)

    def predict_from_formula(self, formula: str, name: Optional[str] = None) -> "PredictExpr":
        """reference polars_ols/__init__.py:289-295"""
        features, add_intercept = build_expressions_from_patsy_formula(formula, include_dependent_variable=False)
        has_const = any(f.output_name == "const" for f in features)
        add_intercept &= not has_const
        return self.predict(*features, name=name, add_intercept=add_intercept)
