"""polars_ols_b200 — B200-native batched least squares behind the polars_ols `least_squares` API.

One hot path: per-group ols / wls / ridge / lasso / elastic_net / rls / rolling_ols with
mode = predictions | residuals | coefficients, evaluated for all groups of an `.over()` at once by
hand-written sm_100a CUDA kernels behind the C ABI of ``include/b200ols.h`` (``libb200ols.so``).
There is NO CPU fallback: importing works anywhere, evaluating needs a CUDA device.
"""
from ._lib import B200OLSError, SO_PATH  # noqa: F401
from .engine import Batch, Col, Engine, as_col, get_engine  # noqa: F401
from .least_squares import (  # noqa: F401
    Expr,
    Frame,
    LeastSquares,
    LsExpr,
    NullPolicy,
    OLSKwargs,
    OutputMode,
    PredictExpr,
    Result,
    RLSKwargs,
    RollingKwargs,
    SolveMethod,
    build_expressions_from_patsy_formula,
    col,
    compute_least_squares,
    compute_least_squares_from_formula,
    compute_multi_target_least_squares,
    compute_recursive_least_squares,
    compute_rolling_least_squares,
    predict,
)

__version__ = "0.1.0"
