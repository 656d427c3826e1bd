#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/bench_configs.py --only C3 --no-cpu --no-e2e --out gpurun_out/i_c3_fused.json > gpurun_out/i_c3_fused.log 2>&1; tail -1 gpurun_out/i_c3_fused.log | cut -c1-250
B200OLS_CD_PRED=0 timeout 300 python tools/bench_configs.py --only C3 --no-cpu --no-e2e --out gpurun_out/i_c3_twopass.json > gpurun_out/i_c3_twopass.log 2>&1; tail -1 gpurun_out/i_c3_twopass.log | cut -c1-250
timeout 300 python -m pytest tests/test_cd_predict_gpu.py -x -q -m gpu > gpurun_out/i_cdtest.log 2>&1; tail -2 gpurun_out/i_cdtest.log
bash tools/gpu_r2_multi.sh 2
