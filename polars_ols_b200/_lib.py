"""ctypes binding of ``libb200ols.so`` (the C ABI declared in ``include/b200ols.h``).

This is the stub a maintainer of the reference would write in Python; the Rust equivalent
(``extern "C"`` block for a ``#[polars_expr]`` wrapper) is shown in ``INTEGRATION.md``.
The library is CUDA-only: there is no CPU fallback.  Loading works without a GPU (so that symbols can
be checked on a build box); creating a context does not.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
import os as _os

SO_PATH = Path(_os.environ["B200OLS_LIBRARY"]).resolve() if _os.environ.get("B200OLS_LIBRARY") else _HERE / "libb200ols.so"

# enums of include/b200ols.h
F64, F32 = 0, 1
HOST, DEVICE = 0, 1
PREDICTIONS, RESIDUALS, COEFFICIENTS = 0, 1, 2
NULL_POLICY = {"ignore": 0, "zero": 1, "drop": 2, "drop_zero": 3, "drop_y_zero_x": 4, "drop_window": 5}
SOLVE_METHOD = {None: 0, "qr": 1, "svd": 2, "chol": 3, "lu": 4, "cd": 5, "cd_active_set": 6}
MODE = {"predictions": 0, "residuals": 1, "coefficients": 2}
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE = 0, -1, -2, -3, -4

EXPORTS = [
    "b200ols_create", "b200ols_create_on_stream", "b200ols_destroy", "b200ols_synchronize",
    "b200ols_last_error", "b200ols_version", "b200ols_launch_count", "b200ols_host_alloc",
    "b200ols_host_free", "b200ols_set_tuning", "b200ols_set_variant", "b200ols_set_profiling", "b200ols_profile_drain", "b200ols_least_squares",
    "b200ols_least_squares_coefficients", "b200ols_recursive_least_squares",
    "b200ols_recursive_least_squares_coefficients", "b200ols_rolling_least_squares",
    "b200ols_rolling_least_squares_coefficients", "b200ols_last_group_flags", "b200ols_predict", "b200ols_device_alloc", "b200ols_device_free", "b200ols_ipc_export",
    "b200ols_ipc_open", "b200ols_ipc_close", "b200ols_copy_to_host", "b200ols_set_peer_gather",
    "b200ols_recursive_least_squares_state", "b200ols_least_squares_statistics", "b200ols_multi_target_least_squares",
    "b200ols_group_plan_build", "b200ols_group_plan_row_index", "b200ols_group_plan_group_of_row",
    "b200ols_device_memset", "b200ols_set_peer_flags", "b200ols_peer_step_complete", "b200ols_peer_step_signal", "b200ols_peer_step_wait", "b200ols_peer_step_signal_wait", "b200ols_peer_arm_step", "b200ols_peer_timed_out",
]


class Column(C.Structure):
    _fields_ = [("values", C.c_void_p), ("validity", C.c_void_p)]


class Frame(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int64), ("n_features", C.c_int32), ("dtype", C.c_int32), ("memspace", C.c_int32),
        ("add_intercept", C.c_int32), ("target", Column), ("features", C.POINTER(Column)),
        ("sample_weights", C.POINTER(Column)), ("n_groups", C.c_int64), ("group_offsets", C.c_void_p),
        ("row_index", C.c_void_p), ("row_index_on_device", C.c_int32), ("_reserved", C.c_int32),
    ]


class KeyColumn(C.Structure):
    _fields_ = [("values", C.c_void_p), ("dtype", C.c_int32), ("_reserved", C.c_int32)]


class GroupPlan(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("n_groups", C.c_int64), ("group_offsets", C.c_void_p),
                ("group_first_row", C.c_void_p), ("row_index", C.c_void_p), ("device_ms", C.c_double)]


KEY_I64, KEY_I32, KEY_U64, KEY_U32, KEY_F64, KEY_F32 = range(6)


class OLSKwargs(C.Structure):
    _fields_ = [
        ("alpha", C.c_double), ("l1_ratio", C.c_double), ("max_iter", C.c_int64), ("tol", C.c_double),
        ("positive", C.c_int32), ("solve_method", C.c_int32), ("null_policy", C.c_int32),
        ("_reserved", C.c_int32), ("rcond", C.c_double),
    ]


class RLSKwargs(C.Structure):
    _fields_ = [
        ("half_life", C.c_double), ("initial_state_covariance", C.c_double), ("initial_state_mean", C.c_void_p),
        ("null_policy", C.c_int32), ("_reserved", C.c_int32), ("initial_information", C.c_void_p),
    ]


class RollingKwargs(C.Structure):
    _fields_ = [
        ("window_size", C.c_int64), ("min_periods", C.c_int64), ("use_woodbury", C.c_int32),
        ("null_policy", C.c_int32), ("alpha", C.c_double),
    ]


class Output(C.Structure):
    _fields_ = [("values", C.c_void_p), ("validity", C.c_void_p)]


class StatisticsOutput(C.Structure):
    _fields_ = [("r2", C.c_void_p), ("mae", C.c_void_p), ("mse", C.c_void_p), ("coefficients", C.c_void_p),
                ("standard_errors", C.c_void_p), ("t_values", C.c_void_p), ("p_values", C.c_void_p)]


class B200OLSError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libb200ols error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library; fails loudly if it has not been built (``python -m polars_ols_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not SO_PATH.exists():
        raise ImportError(
            f"{SO_PATH} is missing: build it with `python polars_ols_b200/build.py` (nvcc, sm_100a). "
            "polars_ols_b200 has no CPU fallback.")
    L = C.CDLL(str(SO_PATH))
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.b200ols_create.argtypes = [i32, C.POINTER(vp)]
    L.b200ols_create_on_stream.argtypes = [i32, vp, C.POINTER(vp)]
    L.b200ols_destroy.argtypes = [vp]
    L.b200ols_destroy.restype = None
    L.b200ols_synchronize.argtypes = [vp]
    L.b200ols_last_error.restype = C.c_char_p
    L.b200ols_version.restype = i32
    L.b200ols_launch_count.argtypes = [vp]
    L.b200ols_launch_count.restype = i64
    L.b200ols_host_alloc.argtypes = [C.c_size_t]
    L.b200ols_host_alloc.restype = vp
    L.b200ols_host_free.argtypes = [vp]
    L.b200ols_host_free.restype = None
    L.b200ols_set_tuning.argtypes = [vp, i32, i32, i32]
    L.b200ols_set_variant.argtypes = [vp, i32, i32]
    L.b200ols_set_profiling.argtypes = [vp, i32]
    L.b200ols_profile_drain.argtypes = [vp, vp, i32]
    L.b200ols_least_squares.argtypes = [vp, C.POINTER(Frame), C.POINTER(OLSKwargs), i32, C.POINTER(Output)]
    L.b200ols_least_squares_coefficients.argtypes = [vp, C.POINTER(Frame), C.POINTER(OLSKwargs), C.POINTER(Output)]
    L.b200ols_recursive_least_squares.argtypes = [vp, C.POINTER(Frame), C.POINTER(RLSKwargs), i32, C.POINTER(Output)]
    L.b200ols_recursive_least_squares_coefficients.argtypes = [vp, C.POINTER(Frame), C.POINTER(RLSKwargs), C.POINTER(Output)]
    L.b200ols_recursive_least_squares_state.argtypes = [vp, C.POINTER(Frame), C.POINTER(RLSKwargs), vp]
    L.b200ols_rolling_least_squares.argtypes = [vp, C.POINTER(Frame), C.POINTER(RollingKwargs), i32, C.POINTER(Output)]
    L.b200ols_rolling_least_squares_coefficients.argtypes = [vp, C.POINTER(Frame), C.POINTER(RollingKwargs), C.POINTER(Output)]
    L.b200ols_last_group_flags.argtypes = [vp, vp, i64]
    L.b200ols_device_alloc.argtypes = [vp, C.c_size_t]
    L.b200ols_device_alloc.restype = vp
    L.b200ols_device_free.argtypes = [vp, vp]
    L.b200ols_device_free.restype = None
    L.b200ols_ipc_export.argtypes = [vp, vp, vp]
    L.b200ols_ipc_open.argtypes = [vp, vp, C.POINTER(vp)]
    L.b200ols_ipc_close.argtypes = [vp, vp]
    L.b200ols_copy_to_host.argtypes = [vp, vp, vp, C.c_size_t]
    L.b200ols_set_peer_gather.argtypes = [vp, i32, C.POINTER(vp), i64, i64]
    L.b200ols_predict.argtypes = [vp, i64, i32, i32, i32, C.POINTER(Column), C.POINTER(Column), i32, i32, C.POINTER(Output)]
    L.b200ols_least_squares_statistics.argtypes = [vp, C.POINTER(Frame), C.POINTER(OLSKwargs), C.POINTER(StatisticsOutput)]
    L.b200ols_multi_target_least_squares.argtypes = [vp, C.POINTER(Frame), i32, C.POINTER(Column), C.POINTER(OLSKwargs), i32,
                                                     C.POINTER(Output)]
    L.b200ols_device_memset.argtypes = [vp, vp, i32, C.c_size_t]
    L.b200ols_set_peer_flags.argtypes = [vp, i32, C.POINTER(vp), i32]
    L.b200ols_peer_step_complete.argtypes = [vp, C.c_uint64]
    L.b200ols_peer_step_signal.argtypes = [vp, C.c_uint64]
    L.b200ols_peer_step_wait.argtypes = [vp, C.c_uint64]
    L.b200ols_peer_step_signal_wait.argtypes = [vp, C.c_uint64, C.c_uint64]
    L.b200ols_peer_arm_step.argtypes = [vp, C.c_uint64, C.c_uint64]
    L.b200ols_peer_timed_out.argtypes = [vp]
    L.b200ols_group_plan_build.argtypes = [vp, C.POINTER(KeyColumn), i32, i64, i32, C.POINTER(GroupPlan)]
    L.b200ols_group_plan_row_index.argtypes = [vp, vp, i32]
    L.b200ols_group_plan_group_of_row.argtypes = [vp, vp, i32]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise B200OLSError(rc, load().b200ols_last_error().decode("utf-8", "replace"))
