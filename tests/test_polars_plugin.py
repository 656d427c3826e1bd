"""The `_polars_plugin_*` adapter of libb200ols.so, driven through
tests/polars_ffi_harness.py.  CPU part: kwargs pickle reader, schema functions, exported symbols, error path without
a device.  GPU part (`-m gpu`): every plugin function against the oracle, incl. multi-chunk / integer / null inputs."""
import ctypes as C
import dataclasses
import pickle

import numpy as np
import pyarrow as pa
import pytest

import polars_ffi_harness as H
from polars_ols_b200 import OLSKwargs, RLSKwargs, RollingKwargs
from polars_ols_b200 import _lib as L

PLUGIN_FUNCTIONS = ["least_squares", "least_squares_coefficients", "least_squares_statistics", "multi_target_least_squares",
                    "recursive_least_squares", "recursive_least_squares_coefficients", "rolling_least_squares",
                    "rolling_least_squares_coefficients", "predict"]     # src/expressions.rs:390,430,468,521,593,624,648,678,706


def _describe(blob: bytes):
    lib = L.load()
    lib.b200ols_plugin_describe_kwargs.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    buf = C.create_string_buffer(4096)
    n = lib.b200ols_plugin_describe_kwargs(blob, len(blob), buf, 4096)
    return None if n < 0 else buf.value.decode()


def test_plugin_symbols_are_exported():
    lib = L.load()
    for f in PLUGIN_FUNCTIONS:
        assert hasattr(lib, f"_polars_plugin_{f}") and hasattr(lib, f"_polars_plugin_field_{f}")
    lib._polars_plugin_get_version.restype = C.c_uint32
    assert lib._polars_plugin_get_version() == 1
    assert hasattr(lib, "_polars_plugin_get_last_error_message")


@pytest.mark.parametrize("protocol", [2, 3, 4, 5])
def test_kwargs_pickle_reader(protocol):
    """kwargs travel as pickle.dumps(dataclasses.asdict(kwargs)) (polars_ols/least_squares.py:70-71,218,230)"""
    kw = dataclasses.asdict(OLSKwargs(alpha=0.25, l1_ratio=None, max_iter=70000, tol=1e-7, positive=True, solve_method="svd",
                                      null_policy="drop_y_zero_x", rcond=None))
    s = _describe(pickle.dumps(kw, protocol=protocol))
    assert s == "null_policy=drop_y_zero_x;alpha=0.25;l1_ratio=None;max_iter=70000;tol=9.9999999999999995e-08;positive=True;solve_method=svd;rcond=None;"
    kw = dataclasses.asdict(RLSKwargs(half_life=252.0, initial_state_mean=[0.5, -1.0, 2], null_policy="drop"))
    s = _describe(pickle.dumps(kw, protocol=protocol))
    assert s == "null_policy=drop;half_life=252;initial_state_covariance=10;initial_state_mean=[0.5,-1,2];"
    kw = dataclasses.asdict(RollingKwargs(window_size=1_000_000, min_periods=-3, use_woodbury=False, alpha=None, null_policy="drop"))
    s = _describe(pickle.dumps(kw, protocol=protocol))
    assert s == "null_policy=drop;window_size=1000000;min_periods=-3;use_woodbury=False;alpha=None;"
    # repeated strings are memo references (BINGET)
    assert _describe(pickle.dumps({"a": "drop", "b": "drop", "c": 2 ** 40, "d": -(2 ** 33)}, protocol=protocol)) == \
        "a=drop;b=drop;c=1099511627776;d=-8589934592;"
    assert _describe(b"not a pickle") is None
    assert _describe(pickle.dumps([1, 2, 3], protocol=protocol)) is None   # not a dict


def test_field_functions():
    fields = [pa.field("y", pa.float32()), pa.field("x1", pa.float64()), pa.field("", pa.int64())]
    for f in ("least_squares", "recursive_least_squares", "rolling_least_squares", "predict"):
        out = H.call_field(f, fields)
        assert out.name == "y" and out.type == pa.float64()              # output_type=Float64
    for f in ("least_squares_coefficients", "recursive_least_squares_coefficients", "rolling_least_squares_coefficients"):
        out = H.call_field(f, fields)                                      # coefficients_struct_dtype (src/expressions.rs:105-111)
        assert out.name == "coefficients" and pa.types.is_struct(out.type)
        assert [out.type.field(i).name for i in range(out.type.num_fields)] == ["x1", ""]
        assert all(out.type.field(i).type == pa.float64() for i in range(2))
    out = H.call_field("least_squares_statistics", fields)                 # statistics_struct_dtype (:448-466)
    assert out.name == "statistics"
    assert [out.type.field(i).name for i in range(out.type.num_fields)] == [
        "r2", "mae", "mse", "feature_names", "coefficients", "standard_errors", "t_values", "p_values"]
    assert out.type.field(3).type == pa.large_list(pa.large_string()) and out.type.field(4).type == pa.large_list(pa.float64())
    st = pa.field("y", pa.struct([("a", pa.float64()), ("b", pa.float64())]))
    out = H.call_field("multi_target_least_squares", [st, fields[1]])      # multi_target_struct_dtype (:511-519)
    assert out.name == "predictions" and [out.type.field(i).name for i in range(2)] == ["a", "b"]
    with pytest.raises(RuntimeError, match="struct"):
        H.call_field("multi_target_least_squares", fields)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="error path of a box without a CUDA device")
def test_no_device_is_an_error_not_a_fallback():
    y = pa.array(np.arange(8.0))
    x = pa.array(np.arange(8.0) ** 2)
    with pytest.raises(RuntimeError, match="(?i)cuda|device"):
        H.call("least_squares", [("y", y), ("x", x)], dataclasses.asdict(OLSKwargs()))     # inputs still released (harness checks)
    with pytest.raises(RuntimeError, match="at least 2 series"):
        H.call("least_squares", [("y", y)], dataclasses.asdict(OLSKwargs()))
    with pytest.raises(RuntimeError, match="kwargs"):
        lib = L.load()
        ex = H.Exported([("y", y), ("x", x)])
        ret = H.SeriesExport()
        fn = lib._polars_plugin_least_squares
        fn.restype = None
        fn.argtypes = [C.POINTER(H.SeriesExport), C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(H.SeriesExport)]
        fn(ex.exports, 2, b"garbage", 7, C.byref(ret))
        ex.check_consumed()
        assert not ret.release
        lib._polars_plugin_get_last_error_message.restype = C.c_char_p
        raise RuntimeError(lib._polars_plugin_get_last_error_message().decode())


# --------------------------------------------------------------------------------------------------- GPU: parity
def _data(n=600, k=3, seed=0, nulls=False):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, k))
    y = x @ (1.0 + np.arange(k)) + 0.1 * rng.normal(size=n)
    masks = [(rng.random(n) >= 0.1) if nulls else None for _ in range(k + 1)]
    return x, y, masks


def _pa(v, m=None, chunks=1, dtype=None):
    a = pa.array(v if dtype is None else v.astype(dtype), mask=None if m is None else ~m)
    if chunks == 1:
        return a
    cut = len(v) // 3
    return pa.chunked_array([a.slice(0, cut), a.slice(cut)])


def _np(arr):
    return np.array([np.nan if v is None else v for v in arr.to_pylist()], dtype=np.float64)


@pytest.mark.gpu
@pytest.mark.parametrize("policy", ["ignore", "drop", "drop_zero", "zero"])
def test_plugin_least_squares_and_coefficients(policy):
    from oracle import semantics as S
    x, y, masks = _data(nulls=policy != "ignore")
    series = [("y", _pa(y, masks[3], chunks=2))] + [(f"x{j}", _pa(x[:, j], masks[j], chunks=2 if j == 1 else 1)) for j in range(3)]
    kw = OLSKwargs(alpha=0.1, l1_ratio=0.0, null_policy=policy)
    name, out = H.call("least_squares", series, dataclasses.asdict(kw))
    cols = [(y, masks[3])] + [(x[:, j], masks[j]) for j in range(3)]
    ref_v, ref_m = S.plugin_least_squares(cols, S.OLSKwargs(alpha=0.1, l1_ratio=0.0, null_policy=policy))
    assert name == "y" and out.type == pa.float64()
    got = _np(out)
    ref = np.where(ref_m, ref_v, np.nan) if ref_m is not None else ref_v
    assert (np.isnan(got) == np.isnan(ref)).all() and np.allclose(got[~np.isnan(ref)], ref[~np.isnan(ref)], rtol=1e-6, atol=1e-9)
    name, out = H.call("least_squares_coefficients", series, dataclasses.asdict(kw))
    c_ref, _ = S.plugin_least_squares_coefficients(cols, S.OLSKwargs(alpha=0.1, l1_ratio=0.0, null_policy=policy))
    assert name == "coefficients" and len(out) == 1 and [f.name for f in out.type] == ["x0", "x1", "x2"]
    assert np.allclose([out.field(j)[0].as_py() for j in range(3)], c_ref, rtol=1e-6)


@pytest.mark.gpu
def test_plugin_casts_integer_and_f32_inputs():
    from oracle import semantics as S
    rng = np.random.default_rng(2)
    xi = rng.integers(-5, 6, size=400)
    xf = rng.normal(size=400).astype(np.float32)
    y = 2.0 * xi - xf + 0.05 * rng.normal(size=400)
    name, out = H.call("least_squares", [("y", pa.array(y)), ("", pa.array(xi)), ("xf", pa.array(xf))], dataclasses.asdict(OLSKwargs()))
    ref, _ = S.plugin_least_squares([(y, None), (xi.astype(np.float64), None), (xf.astype(np.float64), None)], S.OLSKwargs())
    assert np.allclose(_np(out), ref, rtol=1e-6, atol=1e-9)
    _, out = H.call("least_squares_coefficients", [("y", pa.array(y)), ("", pa.array(xi)), ("xf", pa.array(xf))],
                    dataclasses.asdict(OLSKwargs()))
    assert [f.name for f in out.type] == ["0", "xf"]                        # empty name -> index (src/expressions.rs:126-130)


@pytest.mark.gpu
def test_plugin_moving_models():
    from oracle import semantics as S
    x, y, masks = _data(n=800, nulls=True, seed=3)
    series = [("y", _pa(y, masks[3]))] + [(f"x{j}", _pa(x[:, j], masks[j])) for j in range(3)]
    cols = [(y, masks[3])] + [(x[:, j], masks[j]) for j in range(3)]
    rk = RLSKwargs(half_life=50.0, initial_state_mean=[0.1, 0.2, 0.3], null_policy="drop")
    srk = S.RLSKwargs(half_life=50.0, initial_state_mean=[0.1, 0.2, 0.3], null_policy="drop")
    _, out = H.call("recursive_least_squares_coefficients", series, dataclasses.asdict(rk))
    ref, _ = S.plugin_recursive_least_squares_coefficients(cols, srk)
    got = np.stack([_np(out.field(j)) for j in range(3)], axis=1)
    assert np.allclose(got, ref, rtol=1e-6, atol=1e-8)
    name, out = H.call("recursive_least_squares", series, dataclasses.asdict(rk))
    ref_v, ref_m = S.plugin_recursive_least_squares(cols, srk)
    got = _np(out)
    assert name == "y" and (np.isnan(got) == ~ref_m).all() and np.allclose(got[ref_m], ref_v[ref_m], rtol=1e-6, atol=1e-8)
    wk = RollingKwargs(window_size=40, min_periods=5, null_policy="drop")
    swk = S.RollingKwargs(window_size=40, min_periods=5, null_policy="drop")
    _, out = H.call("rolling_least_squares_coefficients", series, dataclasses.asdict(wk))
    ref, _ = S.plugin_rolling_least_squares_coefficients(cols, swk)
    got = np.stack([_np(out.field(j)) for j in range(3)], axis=1)
    assert (np.isnan(got) == np.isnan(ref)).all() and np.allclose(got[~np.isnan(ref)], ref[~np.isnan(ref)], rtol=1e-6, atol=1e-8)
    _, out = H.call("rolling_least_squares", series, dataclasses.asdict(wk))
    ref_v, ref_m = S.plugin_rolling_least_squares(cols, swk)
    got = _np(out)
    ok = ref_m & ~np.isnan(ref_v)
    assert np.allclose(got[ok], ref_v[ok], rtol=1e-6, atol=1e-8) and np.isnan(got[~ref_m]).all()


@pytest.mark.gpu
def test_plugin_statistics_multi_target_and_predict():
    from oracle import semantics as S
    x, y, _ = _data(n=500, seed=4)
    series = [("y", pa.array(y))] + [(f"x{j}", pa.array(x[:, j])) for j in range(3)]
    cols = [(y, None)] + [(x[:, j], None) for j in range(3)]
    name, out = H.call("least_squares_statistics", series, dataclasses.asdict(OLSKwargs(alpha=0.0)))
    ref = S.plugin_least_squares_statistics(cols, S.OLSKwargs(alpha=0.0))
    row = out[0].as_py()
    assert name == "statistics" and row["feature_names"] == ["x0", "x1", "x2"]
    for k in ("r2", "mae", "mse"):
        assert row[k] == pytest.approx(float(ref[k]), rel=1e-6)
    for k in ("coefficients", "standard_errors", "t_values", "p_values"):
        assert np.allclose(row[k], ref[k], rtol=1e-6, atol=3e-16)
    # multi-target: struct of targets in, struct of predictions out
    y2 = x[:, 0] - x[:, 2]
    tgt = pa.StructArray.from_arrays([pa.array(y), pa.array(y2)], names=["a", "b"])
    name, out = H.call("multi_target_least_squares", [("ys", tgt)] + series[1:], dataclasses.asdict(OLSKwargs(alpha=0.5, solve_method="svd")))
    ref_v, _ = S.plugin_multi_target_least_squares([(y, None), (y2, None)], cols[1:], S.OLSKwargs(alpha=0.5, solve_method="svd"))
    assert name == "predictions" and [f.name for f in out.type] == ["a", "b"]
    assert np.allclose(np.stack([_np(out.field(0)), _np(out.field(1))], 1), ref_v, rtol=1e-6, atol=1e-9)
    # predict: coefficient struct (one row per sample) x features
    coef = np.tile([1.0, -2.0, 0.5], (500, 1))
    cs = pa.StructArray.from_arrays([pa.array(coef[:, j]) for j in range(3)], names=["x0", "x1", "x2"])
    name, out = H.call("predict", [("coefficients", cs)] + series[1:], {"null_policy": "zero"})
    assert name == "coefficients" and np.allclose(_np(out), x @ [1.0, -2.0, 0.5], rtol=1e-12, atol=1e-12)
