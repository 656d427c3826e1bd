#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_moving_paths_gpu.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g_tests.log; tail -4 gpurun_out/g_tests.log
timeout 600 python tools/bench_configs.py --only C4,C5 --out gpurun_out/g_c4c5.json > gpurun_out/g_c4c5.log 2>&1; grep "^C" gpurun_out/g_c4c5.log | cut -c1-200
bash tools/gpu_r2_multi.sh 2
