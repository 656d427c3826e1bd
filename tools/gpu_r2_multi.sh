#!/bin/bash
# usage: bash tools/gpu_r2_multi.sh N   (under gpurun --gpus N)
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/m${N}_topo.txt 2>&1; nproc >> gpurun_out/m${N}_topo.txt; numactl -H >> gpurun_out/m${N}_topo.txt 2>&1
nvidia-smi nvlink -gt d > gpurun_out/m${N}_nvlink_before.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err; tail -5 gpurun_out/m${N}_bench.err; tail -1 gpurun_out/m${N}_bench.json
nvidia-smi nvlink -gt d > gpurun_out/m${N}_nvlink_after.txt 2>&1
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --gather nccl > gpurun_out/m${N}_bench_nccl.json 2> gpurun_out/m${N}_bench_nccl.err; tail -1 gpurun_out/m${N}_bench_nccl.json
timeout 600 $TR --master-port 29513 tools/multi_gpu_configs.py --out gpurun_out/m${N}_c3c5.json > gpurun_out/m${N}_c3c5.log 2>&1; tail -2 gpurun_out/m${N}_c3c5.log
timeout 600 $TR --master-port 29514 tools/time_shard_check.py --rows 50000000 > gpurun_out/m${N}_c4_shards.log 2>&1; tail -2 gpurun_out/m${N}_c4_shards.log
timeout 300 python bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/m${N}_bench_ref.json 2>&1; tail -1 gpurun_out/m${N}_bench_ref.json
