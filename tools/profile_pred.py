"""A few launches of the fused Gram -> solve -> predict kernel (gram_pred.cuh) on the C2 workload, for
`ncu --set full --import-source on -k regex:gram_pred -s 3 -c 1`."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import polars_ols_b200 as pls
from polars_ols_b200 import _lib as L
import bench

x, y, offsets = bench.make_data(0)
dev = torch.device("cuda", 0)
xd, yd = torch.as_tensor(x, device=dev), torch.as_tensor(y, device=dev)
kw = pls.OLSKwargs(alpha=bench.ALPHA, l1_ratio=0.0).to_c()
eng = pls.Engine(0, 1)
batch = pls.Batch(pls.Col(yd), [pls.Col(xd[i]) for i in range(bench.K)], offsets=offsets)
mode = L.RESIDUALS if "residuals" in sys.argv else L.PREDICTIONS
for _ in range(5):
    eng.least_squares(batch, kw, mode)
torch.cuda.synchronize()
