"""rolling + rls on a C4-shaped series (reduced), for ncu."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import polars_ols_b200 as pls
from polars_ols_b200 import _lib as L
n, k = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000, 6
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(k, n, dtype=torch.float64, device=dev, generator=g)
y = x.sum(0) + 0.1 * torch.randn(n, dtype=torch.float64, device=dev, generator=g)
eng = pls.Engine(0, 1)
b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)])
kwr = pls.RollingKwargs(window_size=252, min_periods=6, null_policy="drop").to_c()
kwl = L.RLSKwargs(252.0, 10.0, None, L.NULL_POLICY["drop"], 0)
for _ in range(2):
    eng.rolling_least_squares(b, kwr, L.PREDICTIONS)
    eng.recursive_least_squares(b, kwl, L.PREDICTIONS)
torch.cuda.synchronize()
