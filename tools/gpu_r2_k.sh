#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_full_size_gpu.py -x -q -m gpu > gpurun_out/k_fullsize.log 2>&1; echo "rc=$?" >> gpurun_out/k_fullsize.log; tail -6 gpurun_out/k_fullsize.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/k_m2_bench.json 2> gpurun_out/k_m2_bench.err; tail -3 gpurun_out/k_m2_bench.err | cut -c1-300; tail -1 gpurun_out/k_m2_bench.json | cut -c1-400
timeout 600 $TR --master-port 29524 tools/time_shard_check.py --rows 50000000 > gpurun_out/k_m2_c4_shards.log 2>&1; tail -1 gpurun_out/k_m2_c4_shards.log
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_full_size_gpu.py > gpurun_out/k_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/k_pytest.log; tail -3 gpurun_out/k_pytest.log
