#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log; tail -8 gpurun_out/h_pytest.log
timeout 1200 python tools/bench_configs.py --out gpurun_out/h_configs.json > gpurun_out/h_configs.log 2>&1; tail -3 gpurun_out/h_configs.log | cut -c1-300
B200OLS_CD_PRED=0 timeout 300 python tools/bench_configs.py --only C3 --no-cpu --no-e2e --out gpurun_out/h_c3_twopass.json > gpurun_out/h_c3_twopass.log 2>&1; tail -1 gpurun_out/h_c3_twopass.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; tail -2 gpurun_out/h_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/h_bench_ref.json 2>&1
cat > /tmp/prof_c5.py <<'PY'
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
import torch, numpy as np, polars_ols_b200 as pls
from polars_ols_b200 import _lib as L
from bench_configs import _gen
dev = torch.device('cuda', 0)
which = sys.argv[1]
eng = pls.Engine(0, 1)
if which == 'C5':
    G, per, k = 200, 10000, 64
    x, y = _gen(torch, dev, G * per, k, G, torch.float64, 5)
    b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)], offsets=np.arange(G + 1, dtype=np.int64) * per)
    kw = pls.OLSKwargs(alpha=1e-4, l1_ratio=1.0).to_c()
    for _ in range(2): eng.least_squares(b, kw, L.COEFFICIENTS)
else:
    G, per, k = 20000, 256, 16
    x, y = _gen(torch, dev, G * per, k, G, torch.float32, 3)
    w = torch.rand(G * per, dtype=torch.float32, device=dev) + 0.05
    b = pls.Batch(pls.Col(y), [pls.Col(x[i]) for i in range(k)], weights=pls.Col(w), offsets=np.arange(G + 1, dtype=np.int64) * per)
    kw = pls.OLSKwargs(alpha=1e-3, l1_ratio=0.5).to_c()
    for _ in range(2): eng.least_squares(b, kw, L.PREDICTIONS)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gram_wide|cd_solve' -c 2 -o gpurun_out/h_c5 -f python /tmp/prof_c5.py C5 > gpurun_out/h_ncu_c5.log 2>&1; tail -2 gpurun_out/h_ncu_c5.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gram_multi|cd_solve' -c 2 -o gpurun_out/h_c3 -f python /tmp/prof_c5.py C3 > gpurun_out/h_ncu_c3.log 2>&1; tail -2 gpurun_out/h_ncu_c3.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/h_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/h_launches_bench.log 2>&1
