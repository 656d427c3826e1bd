"""Build + ctypes-load ``oracle/ols_oracle.c`` (test infrastructure, see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_HERE = Path(__file__).resolve().parent
_SRC = _HERE / "ols_oracle.c"
_SO = _HERE / "libols_oracle.so"

_lib = None


def build_oracle(force: bool = False) -> Path:
    """gcc -O3 -fopenmp -shared oracle/ols_oracle.c -> oracle/libols_oracle.so (git-ignored)."""
    if force or not _SO.exists() or _SO.stat().st_mtime < _SRC.stat().st_mtime:
        tmp = _SO.with_suffix(f".{os.getpid()}.tmp.so")
        cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared", "-fvisibility=hidden",
               "-ffp-contract=off", "-o", str(tmp), str(_SRC), "-lm"]
        subprocess.run(cmd, check=True, capture_output=True, text=True)
        os.replace(tmp, _SO)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    so = build_oracle()
    L = C.CDLL(str(so))
    d, i64, i32, vp = C.c_double, C.c_int64, C.c_int, C.c_void_p
    L.orc_solve_ols_qr.argtypes = [vp, vp, i64, i32, vp]
    L.orc_solve_ols_qr.restype = None
    L.orc_solve_normal_equations.argtypes = [vp, vp, i32, i32]
    L.orc_solve_normal_equations.restype = i32
    L.orc_solve_ridge.argtypes = [vp, vp, i64, i32, d, i32, vp]
    L.orc_solve_ridge.restype = i32
    L.orc_solve_elastic_net.argtypes = [vp, vp, i64, i32, d, d, i64, d, i32, i32, vp]
    L.orc_solve_elastic_net.restype = i32
    L.orc_solve_recursive_least_squares.argtypes = [vp, vp, i64, i32, d, d, vp, vp, vp]
    L.orc_solve_recursive_least_squares.restype = None
    L.orc_solve_rolling_ols.argtypes = [vp, vp, i64, i32, i64, i64, i32, d, vp, i32, vp]
    L.orc_solve_rolling_ols.restype = None
    L.orc_update_xtx_inv.argtypes = [vp, i32, vp, vp, i32]
    L.orc_update_xtx_inv.restype = None
    L.orc_inv_lu.argtypes = [vp, i32, vp]
    L.orc_inv_lu.restype = None
    L.orc_grouped_least_squares_coefficients.argtypes = [vp, i32, vp, i64, i32, d, d, i64, d, i32, i32, vp]
    L.orc_grouped_least_squares_coefficients.restype = None
    L.orc_grouped_least_squares_predictions.argtypes = [vp, i32, vp, vp, i64, i32, d, d, i64, d, i32, i32, vp]
    L.orc_grouped_least_squares_predictions.restype = None
    L.orc_max_threads.argtypes = []
    L.orc_max_threads.restype = i32
    _lib = L
    return L
