"""C2 predictions / residuals through the fused kernel (gram_pred.cuh) for several producer lags, and the two-pass
route (B200OLS_PRED=0) beside them.  Writes gpurun_out/sweep_pred.json."""
import json
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import polars_ols_b200 as pls
from polars_ols_b200 import _lib as L
import bench

x, y, offsets = bench.make_data(0)
dev = torch.device("cuda", 0)
xd, yd = torch.as_tensor(x, device=dev), torch.as_tensor(y, device=dev)
kw = pls.OLSKwargs(alpha=bench.ALPHA, l1_ratio=0.0).to_c()
batch = pls.Batch(pls.Col(yd), [pls.Col(xd[i]) for i in range(bench.K)], offsets=offsets)
res = {}
ref = {}
lags = [int(t) for t in sys.argv[1:]] or [1, 2, 3, 4, 6, 8, 12]
for lag in [0] + lags:
    os.environ["B200OLS_PRED"] = "0" if lag == 0 else "1"
    os.environ["B200OLS_PRED_LAG"] = str(max(lag, 1))
    eng = pls.Engine(0, 1)
    for mode, nm in ((L.PREDICTIONS, "predictions"), (L.RESIDUALS, "residuals")):
        for _ in range(3):
            out = eng.least_squares(batch, kw, mode)[0]
        torch.cuda.synchronize()
        if lag == 0:
            ref[nm] = out.clone()
        err = float((out - ref[nm]).abs().max())
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
        ev[0].record()
        for i in range(10):
            eng.least_squares(batch, kw, mode)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(10)]))
        res[f"lag{lag}_{nm}"] = ms
        print("two-pass" if lag == 0 else f"lag {lag}", nm, f"{ms:.4f} ms  {0.8 / ms * 1e3:.0f} GB/s  max |diff to two-pass| {err:.1e}", flush=True)
    eng.close()
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/sweep_pred.json").write_text(json.dumps(res, indent=1))
