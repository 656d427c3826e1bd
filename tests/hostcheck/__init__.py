"""Host-compiled copy of the device algorithm templates (tests only; see hostcheck.cpp)."""
import ctypes as C
import os
import subprocess
from pathlib import Path

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "libhostcheck.so"
_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    src = _HERE / "hostcheck.cpp"
    deps = [src] + list((_HERE.parent.parent / "polars_ols_b200" / "csrc").glob("*.cuh"))
    if not _SO.exists() or _SO.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        tmp = _SO.with_suffix(f".{os.getpid()}.tmp.so")
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-ffp-contract=off",
                        "-Wno-unknown-pragmas", "-o", str(tmp), str(src)], check=True, capture_output=True, text=True)
        os.replace(tmp, _SO)
    L = C.CDLL(str(_SO))
    d, i64, i32, vp = C.c_double, C.c_int64, C.c_int, C.c_void_p
    L.hc_normal_equations.argtypes = [vp, i32, vp, i32, d]
    L.hc_cd_gram.argtypes = [vp, i32, vp, d, d, i64, d, i32, i32, vp]
    L.hc_rolling.argtypes = [vp, vp, vp, i64, i32, i64, i64, d, i32, i64, vp]
    L.hc_rolling.restype = None
    L.hc_rls.argtypes = [vp, vp, vp, i64, i32, d, d, vp, i64, vp]
    L.hc_rls.restype = None
    L.hc_students_t_p.argtypes = [d, d]
    L.hc_students_t_p.restype = d
    L.hc_chol_inverse.argtypes = [vp, i32]
    _lib = L
    return L
