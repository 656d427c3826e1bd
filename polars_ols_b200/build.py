"""Builds polars_ols_b200/libb200ols.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
SO = HERE / "libb200ols.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-cudart", "static", "-t", "0",
    # FP contraction stays on in device code: explicit fma() is used where order matters
]


def sources():
    return sorted(CSRC.glob("*.cu"))


def deps():
    return list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "b200ols.h"]


def build(force: bool = False, verbose: bool = False) -> Path:
    newest = max(p.stat().st_mtime for p in deps())
    if not force and SO.exists() and SO.stat().st_mtime >= newest:
        return SO
    tmp = SO.with_suffix(f".{os.getpid()}.tmp.so")
    cmd = [NVCC, *FLAGS, "-o", str(tmp), *[str(s) for s in sources()]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libb200ols.so")
    if verbose:
        sys.stderr.write(r.stderr)
    os.replace(tmp, SO)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
