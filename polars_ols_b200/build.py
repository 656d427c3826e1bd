"""Builds polars_ols_b200/libb200ols.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).
Every .cu is compiled to its own object (in parallel) and re-compiled only when it or a header it includes
(recursively) changed; objects live in polars_ols_b200/build/ (git-ignored)."""
from __future__ import annotations

import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
# A/B builds (tools only): B200OLS_BUILD_TAG=x B200OLS_BUILD_FLAGS="-DFOO=1" -> build_x/ + libb200ols_x.so, loaded
# with B200OLS_LIBRARY=polars_ols_b200/libb200ols_x.so (see _lib.py)
_TAG = os.environ.get("B200OLS_BUILD_TAG", "")
OBJ = HERE / ("build" + (f"_{_TAG}" if _TAG else ""))
SO = HERE / ("libb200ols" + (f"_{_TAG}" if _TAG else "") + ".so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    # FP contraction stays on in device code: explicit fma() is used where order matters
] + os.environ.get("B200OLS_BUILD_FLAGS", "").split()
LINK = ["--shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]
_INC = re.compile(r'^\s*#include\s+"([^"]+)"', re.M)


def sources():
    return sorted(CSRC.glob("*.cu"))


def _deps(path: Path, seen=None):
    seen = seen if seen is not None else set()
    if path in seen or not path.exists():
        return seen
    seen.add(path)
    for inc in _INC.findall(path.read_text()):
        _deps((path.parent / inc).resolve(), seen)
    return seen


def deps():
    return list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "b200ols.h"]


def _compile(src: Path, verbose: bool):
    obj = OBJ / (src.stem + ".o")
    newest = max(p.stat().st_mtime for p in _deps(src.resolve()))
    newest = max(newest, Path(__file__).stat().st_mtime)
    if obj.exists() and obj.stat().st_mtime >= newest:
        return obj, ""
    cmd = [NVCC, *FLAGS, "-c", "-o", str(obj), str(src)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src.name}:\n{r.stdout}{r.stderr}")
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    newest = max(p.stat().st_mtime for p in deps())
    if not force and SO.exists() and SO.stat().st_mtime >= newest:
        return SO
    OBJ.mkdir(exist_ok=True)
    if force:
        for o in OBJ.glob("*.o"):
            o.unlink()
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), sources()))
    if verbose:
        sys.stderr.write("".join(e for _, e in res))
    tmp = SO.with_suffix(f".{os.getpid()}.tmp.so")
    r = subprocess.run([NVCC, *LINK, "-o", str(tmp), *[str(o) for o, _ in res]], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libb200ols.so")
    os.replace(tmp, SO)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
