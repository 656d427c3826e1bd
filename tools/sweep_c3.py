"""C3-shaped Gram (100k groups x 256 rows x 16 features f32, weights): gram_cta team mode (default) against gram_multi
with tiles of several whole groups (set_tuning tile_rows > 0 forces gram_multi).  Prints the streaming kernel's time."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import polars_ols_b200 as pls  # noqa: E402
from polars_ols_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda", 0)
k, n, G = 16, 256, 100_000
g = torch.Generator(device=dev).manual_seed(1)
N = n * G
x = torch.randn(k, N, dtype=torch.float32, device=dev, generator=g)
y = x.sum(0) + 0.1 * torch.randn(N, dtype=torch.float32, device=dev, generator=g)
w = torch.rand(N, dtype=torch.float32, device=dev, generator=g) + 0.1
b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)], pls.Col(w), offsets=np.arange(G + 1, dtype=np.int64) * n)
kw = pls.OLSKwargs(alpha=1e-3, l1_ratio=0.0).to_c()
eng = pls.Engine(0, 1)
coef = torch.empty((G, k), dtype=torch.float64, device=dev)
call = eng.prepare_least_squares(b, kw, L.COEFFICIENTS, coef)
out, ref = {}, None
configs = [(0, 0, 0), (0, 0, 1), (0, 0, 2), (0, 0, 4)] + [(r, s, 0) for r in (512, 768, 1024, 1536, 2048) for s in (0,)]
for tile, stages, team in configs:
    eng.set_tuning(tile, stages, team)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    eng.set_profiling(True)
    for _ in range(10):
        call()
    torch.cuda.synchronize()
    ms = float(np.median(eng.profile_drain()))
    eng.set_profiling(False)
    cc = coef.cpu().numpy().copy()
    ref = cc if ref is None else ref
    key = f"tile_rows={tile} stages={stages} team={team}"
    out[key] = {"gram_ms": round(ms, 4), "GBps": round(N * 18 * 4 / ms / 1e6, 1), "max_abs_diff": float(np.abs(cc - ref).max())}
    print(key, out[key], flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/sweep_c3.json").write_text(json.dumps(out, indent=1))
