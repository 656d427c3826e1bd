"""One launch of each Gram kernel variant (best sweep config) on the C2 workload, for
`ncu --set full -k regex:gram_ -s 9 -c 3`."""
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
coef = torch.empty((bench.G, bench.K), dtype=torch.float64, device=dev)
kw = pls.OLSKwargs(alpha=bench.ALPHA, l1_ratio=0.0).to_c()
eng = pls.Engine(0, 1)
batch = pls.Batch(pls.Col(yd), [pls.Col(xd[i]) for i in range(bench.K)], offsets=offsets)
step = eng.prepare_least_squares(batch, kw, L.COEFFICIENTS, coef)
CONFIGS = [(0, 0, 256, 12, 2), (1, 1, 0, 8, 3), (2, 4, 0, 4, 3)]   # (variant, unroll, tile_rows, warps, ctas_per_sm)
if len(sys.argv) > 1:
    CONFIGS = [tuple(int(t) for t in a.split(",")) for a in sys.argv[1:]]
for variant, unroll, tile, warps, cps in CONFIGS:
    eng.set_variant(variant, unroll)
    eng.set_tuning(tile, warps, cps)
    for _ in range(3):
        step()
for variant, unroll, tile, warps, cps in CONFIGS:
    eng.set_variant(variant, unroll)
    eng.set_tuning(tile, warps, cps)
    step()
torch.cuda.synchronize()
