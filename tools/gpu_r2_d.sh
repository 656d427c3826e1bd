#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_moving_paths_gpu.py -x -q -m gpu > gpurun_out/d_moving.log 2>&1; echo "moving rc=$?" >> gpurun_out/d_moving.log; tail -25 gpurun_out/d_moving.log
timeout 600 python tools/bench_configs.py --only C4 --out gpurun_out/d_c4_nbr.json > gpurun_out/d_c4_nbr.log 2>&1; cat gpurun_out/d_c4_nbr.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'nbr|totals' -c 4 -o gpurun_out/d_moving_nbr -f python tools/profile_moving.py 10000000 > gpurun_out/d_ncu.log 2>&1; tail -5 gpurun_out/d_ncu.log
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log; tail -8 gpurun_out/d_pytest.log
