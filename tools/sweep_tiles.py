"""gram_multi tile size: the default rule (28 KB tiles, or the largest three-stage tile when that is what it takes to
hold four groups) against forced tile sizes, over short / medium group shapes.  Ridge coefficients (batch solve),
device-resident inputs; prints the streaming kernel's own time.  Writes gpurun_out/sweep_tiles.json."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import polars_ols_b200 as pls  # noqa: E402
from polars_ols_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda", 0)
BUDGET = 232448 - 1024


def stride(R, esz):
    raw = (R + 8) * esz
    return ((raw + 127) // 128) * 128 + (64 if esz == 8 else 32)


def r3(nc, esz):
    R = BUDGET // 3 // nc // esz // 8 * 8
    while R > 64 and 3 * nc * stride(R, esz) > BUDGET:
        R -= 8
    return R


cases = [("f64", 8, 64, 160_000, False), ("f64", 8, 128, 80_000, False), ("f64", 8, 256, 40_000, False),
         ("f64", 8, 500, 20_000, False), ("f32", 16, 64, 200_000, True), ("f32", 16, 128, 100_000, True),
         ("f32", 16, 256, 100_000, True), ("f64", 3, 100, 200_000, False)]
out = {}
eng = pls.Engine(0, 1)
for dt, k, n, G, weighted in cases:
    tdt = torch.float32 if dt == "f32" else torch.float64
    esz = 4 if dt == "f32" else 8
    g = torch.Generator(device=dev).manual_seed(1)
    N = n * G
    x = torch.randn(k, N, dtype=tdt, device=dev, generator=g)
    y = x.sum(0) + 0.1 * torch.randn(N, dtype=tdt, device=dev, generator=g)
    w = torch.rand(N, dtype=tdt, device=dev, generator=g) + 0.1 if weighted else None
    b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)], None if w is None else pls.Col(w),
                  offsets=np.arange(G + 1, dtype=np.int64) * n)
    kw = pls.OLSKwargs(alpha=1e-3, l1_ratio=0.0).to_c()
    nc = k + 1 + (1 if weighted else 0)
    coef = torch.empty((G, k), dtype=torch.float64, device=dev)
    call = eng.prepare_least_squares(b, kw, L.COEFFICIENTS, coef)
    R3 = r3(nc, esz)
    ref = None
    for tile in (0, R3, (R3 // 2) // 8 * 8, (R3 * 3 // 4) // 8 * 8):
        eng.set_tuning(tile, 0, 0)
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        eng.set_profiling(True)
        for _ in range(8):
            call()
        torch.cuda.synchronize()
        ms = float(np.median(eng.profile_drain()))
        eng.set_profiling(False)
        cc = coef.cpu().numpy().copy()
        ref = cc if ref is None else ref
        key = f"{dt} k={k} n={n} G={G} tile_rows={tile}"
        out[key] = {"gram_ms": round(ms, 4), "GBps": round(N * nc * esz / ms / 1e6, 1), "max_abs_diff": float(np.abs(cc - ref).max())}
        print(key, out[key], flush=True)
    eng.set_tuning(0, 0, 0)
    del call, coef, x, y, w, b
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/sweep_tiles.json").write_text(json.dumps(out, indent=1))
