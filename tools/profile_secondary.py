"""One call of each secondary configuration (C3, C4 reduced, C5 reduced) for
  ncu --set full --clock-control none --import-source on \
      -k regex:'gram_cta|gram_wide|cd_solve|predict_kernel|rolling_main|rls_main|rls_summary|rls_scan|chunk_transpose' \
      -o gpurun_out/secondary python tools/profile_secondary.py
(the numbers a run under ncu prints are never bench values)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import polars_ols_b200 as pls  # noqa: E402
from polars_ols_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda", 0)
eng = pls.Engine(0, 1)
g = torch.Generator(device=dev).manual_seed(0)


def data(k, N, dt):
    x = torch.randn(k, N, dtype=dt, device=dev, generator=g)
    y = x.sum(0) + 0.1 * torch.randn(N, dtype=dt, device=dev, generator=g)
    return x, y


# C3: WLS + elastic net predictions, 100k groups x 256 rows x 16 features, f32
G, n, k = 100_000, 256, 16
x, y = data(k, G * n, torch.float32)
w = torch.rand(G * n, dtype=torch.float32, device=dev, generator=g) + 0.1
b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)], pls.Col(w), offsets=np.arange(G + 1, dtype=np.int64) * n)
eng.least_squares(b, pls.OLSKwargs(alpha=1e-3, l1_ratio=0.5).to_c(), L.PREDICTIONS)
torch.cuda.synchronize()
del x, y, w, b

# C4 (reduced to 10M rows): rolling_ols(252) and rls(half_life=252), k = 6, f64, predictions
x, y = data(6, 10_000_000, torch.float64)
b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(6)])
eng.rolling_least_squares(b, pls.RollingKwargs(window_size=252, min_periods=6, null_policy="drop").to_c(), L.PREDICTIONS)
eng.recursive_least_squares(b, L.RLSKwargs(252.0, 10.0, None, L.NULL_POLICY["drop"], 0, None), L.PREDICTIONS)
torch.cuda.synchronize()
del x, y, b

# C5 (reduced to 200 groups): lasso, 10k rows x 64 features, f64, coefficients
G, n, k = 200, 10_000, 64
x, y = data(k, G * n, torch.float64)
b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)], offsets=np.arange(G + 1, dtype=np.int64) * n)
eng.least_squares(b, pls.OLSKwargs(alpha=1e-4, l1_ratio=1.0).to_c(), L.COEFFICIENTS)
torch.cuda.synchronize()
