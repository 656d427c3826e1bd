#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_moving_paths_gpu.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/e_tests.log; tail -6 gpurun_out/e_tests.log
timeout 600 python tools/bench_configs.py --only C4 --out gpurun_out/e_c4.json > gpurun_out/e_c4.log 2>&1; cat gpurun_out/e_c4.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; tail -3 gpurun_out/e_bench.err; cat gpurun_out/e_bench.json
# hardening: racecheck / synccheck over the kernels with shared-memory hand-offs (small cases only: the tools slow kernels 10-100x)
export B200OLS_MOVING_NBR_MIN_CHUNKS=1
for tool in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest -x -q -m gpu \
     "tests/test_moving_paths_gpu.py::test_rolling_window_chunk_kernel" "tests/test_moving_paths_gpu.py::test_rolling_wide" "tests/test_moving_paths_gpu.py::test_rls_wide" \
     "tests/test_group_plan_gpu.py::test_random_int64_keys" "tests/test_group_plan_gpu.py::test_multi_key" \
     "tests/test_gpu_parity.py::test_ols" "tests/test_gpu_parity.py::test_wls_intercept_and_f32" "tests/test_gpu_parity.py::test_missing_data" \
     > gpurun_out/e_$tool.log 2>&1; echo "$tool rc=$?" >> gpurun_out/e_$tool.log; grep -c "ERROR SUMMARY" gpurun_out/e_$tool.log; tail -4 gpurun_out/e_$tool.log
done
