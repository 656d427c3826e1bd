#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_moving_paths_gpu.py -x -q -m gpu > gpurun_out/c_moving.log 2>&1; echo "moving rc=$?" >> gpurun_out/c_moving.log; tail -25 gpurun_out/c_moving.log
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log; tail -25 gpurun_out/c_pytest.log
timeout 600 python tools/bench_configs.py --only C4 --out gpurun_out/c_c4_fast.json > gpurun_out/c_c4_fast.log 2>&1; cat gpurun_out/c_c4_fast.log
B200OLS_MOVING_FAST=0 timeout 600 python tools/bench_configs.py --only C4 --out gpurun_out/c_c4_legacy.json > gpurun_out/c_c4_legacy.log 2>&1; cat gpurun_out/c_c4_legacy.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fast' -c 4 -o gpurun_out/c_moving_fast -f python tools/profile_moving.py 10000000 > gpurun_out/c_ncu.log 2>&1; tail -5 gpurun_out/c_ncu.log
ls -la gpurun_out/*.ncu-rep
