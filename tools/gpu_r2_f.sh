#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_moving_paths_gpu.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/f_tests.log; tail -6 gpurun_out/f_tests.log
timeout 600 python tools/bench_configs.py --only C4,C5 --out gpurun_out/f_c4c5.json > gpurun_out/f_c4c5.log 2>&1; grep "^C" gpurun_out/f_c4c5.log
B200OLS_LIBRARY=polars_ols_b200/libb200ols_libsqrt.so timeout 600 python tools/bench_configs.py --only C4 --out gpurun_out/f_c4_libsqrt.json > gpurun_out/f_c4_libsqrt.log 2>&1; grep "^C" gpurun_out/f_c4_libsqrt.log
timeout 600 python tools/bench_configs.py --only C4 --out gpurun_out/f_c4_again.json > gpurun_out/f_c4_again.log 2>&1; grep "^C" gpurun_out/f_c4_again.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'nbr|rls_fast_main' -c 2 -o gpurun_out/f_moving -f python tools/profile_moving.py 50000000 > gpurun_out/f_ncu.log 2>&1; tail -3 gpurun_out/f_ncu.log
