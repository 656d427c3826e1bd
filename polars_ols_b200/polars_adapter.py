"""Batched `.over()` route for REAL polars frames (guarded: polars is not installed in this image, so this module is
exercised only for its import guard).

The reference registers one plugin expression and lets polars call it once per group
(``polars_ols/least_squares.py:199-239``, ``register_plugin_function(..., is_elementwise=False, returns_scalar=...)``
under ``.over()``).  The batched route keeps the user-facing call identical but evaluates the whole window in ONE call
into the engine: the key column(s) go to ``b200ols_group_plan_build`` (device-side group packing), the value columns
to the batched ``b200ols_*`` entry point, and the result comes back as a polars Series (struct of Float64 for
coefficients, broadcast to the frame's rows exactly as ``.over()`` does; Float64 for predictions / residuals).

    import polars as pl
    from polars_ols_b200 import col
    from polars_ols_b200.polars_adapter import over_batched
    expr = col("y").least_squares.ridge("x1", "x2", alpha=1e-3, mode="coefficients").over("group")
    df = df.with_columns(over_batched(df, expr))
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .least_squares import Frame, LsExpr, Result

try:  # pragma: no cover - polars is absent in the build image
    import polars as pl
except Exception:  # noqa: BLE001
    pl = None


def available() -> bool:
    return pl is not None


def _column(series):
    """polars Series -> what Frame accepts: (values, valid) with f32 / f64 values kept, everything else cast to f64
    (src/expressions.rs:33,47,80); zero-copy where Arrow allows it."""
    if series.dtype not in (pl.Float32, pl.Float64):
        if series.dtype.is_numeric() and not series.dtype.is_float():
            if series.null_count() == 0:
                return series.to_numpy()            # integer key / value column
        series = series.cast(pl.Float64)
    arr = series.to_arrow()
    return arr                                       # engine.as_col understands pyarrow arrays (values + validity bitmap)


def to_frame(df, names) -> Frame:
    return Frame({n: _column(df.get_column(n)) for n in names})


def _needed_columns(expr: LsExpr):
    names = []
    for e in [expr.target, *expr.features, expr.sample_weights]:
        if e is not None and e._name is not None and e._data is None:
            names += getattr(e, "_factors", [e._name])
    for k in expr._over:
        if isinstance(k, str):
            names.append(k)
    return list(dict.fromkeys(names))


def over_batched(df, expr: LsExpr, engine=None):
    """Evaluate `expr` (built with this package's `col(...).least_squares.*(...).over(keys)`) on a polars DataFrame in one
    batched engine call and return a polars Series aligned with `df`'s rows, named as the reference names it."""
    if pl is None:
        raise ImportError("polars is not installed: over_batched needs a polars DataFrame")
    fr = to_frame(df, _needed_columns(expr))
    res: Result = expr.evaluate(fr, engine)
    if res.fields is not None and not isinstance(res.values, dict):         # coefficient struct
        v = res.to_numpy(broadcast=True)
        cols = {f: pl.Series(f, v[:, j]).fill_nan(None) for j, f in enumerate(res.fields)}   # src/expressions.rs:137-139
        return pl.DataFrame(cols).to_struct(res.name)
    v = res.to_numpy()
    s = pl.Series(res.name, v)
    null = res.is_null()
    if null.any():
        s = s.scatter(np.flatnonzero(null), None)
    return s


def register_namespace(name: str = "least_squares_b200") -> Optional[type]:
    """`df.least_squares_b200.over(expr)`: a DataFrame namespace for the batched route (polars only)."""
    if pl is None:
        return None

    @pl.api.register_dataframe_namespace(name)
    class _B200Frame:          # pragma: no cover
        def __init__(self, df):
            self._df = df

        def over(self, expr: LsExpr, engine=None):
            return self._df.with_columns(over_batched(self._df, expr, engine))

    return _B200Frame
