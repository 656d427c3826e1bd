"""One C3 call (100k groups x 256 rows x 16 f32, wls + elastic_net predictions) for ncu captures.  Usage: python tools/c3_once.py [calls]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import polars_ols_b200 as pls
from polars_ols_b200 import _lib as L
from tools.bench_configs import _gen

dev = torch.device("cuda", 0)
eng = pls.Engine(0, torch.cuda.current_stream(dev).cuda_stream or 1)
G, per, k = 100_000, 256, 16
x, y = _gen(torch, dev, G * per, k, G, torch.float32, 3)
w = torch.rand(G * per, dtype=torch.float32, device=dev) + 0.05
batch = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)], weights=pls.Col(w), offsets=np.arange(G + 1, dtype=np.int64) * per)
kw = pls.OLSKwargs(alpha=1e-3, l1_ratio=0.5).to_c()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    eng.least_squares(batch, kw, L.PREDICTIONS, want_validity=False)
torch.cuda.synchronize()
