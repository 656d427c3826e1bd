#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_group_plan_gpu.py -x -q -m gpu > gpurun_out/b_plan.log 2>&1; echo "plan rc=$?" >> gpurun_out/b_plan.log; tail -5 gpurun_out/b_plan.log
timeout 300 python tools/rls_diag.py > gpurun_out/b_rls_diag.log 2>&1; cat gpurun_out/b_rls_diag.log
timeout 900 python tools/stage_sweep.py > gpurun_out/b_stage_sweep.log 2>&1; cat gpurun_out/b_stage_sweep.log
