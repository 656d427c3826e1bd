#!/bin/bash
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/l_m8_bench.json 2> gpurun_out/l_m8_bench.err; tail -2 gpurun_out/l_m8_bench.err | cut -c1-300; tail -1 gpurun_out/l_m8_bench.json | cut -c1-300
timeout 600 $TR --master-port 29534 tools/time_shard_check.py --rows 50000000 > gpurun_out/l_m8_c4_shards.log 2>&1; tail -1 gpurun_out/l_m8_c4_shards.log
timeout 900 python -m pytest tests/test_full_size_gpu.py -x -q -m gpu > gpurun_out/l_fullsize.log 2>&1; echo "rc=$?" >> gpurun_out/l_fullsize.log; tail -4 gpurun_out/l_fullsize.log | cut -c1-200
