#!/bin/bash
# Multi-GPU measurements of one box:   gpurun --gpus N -- 'bash tools/gpu_multi.sh N'
# bench.py under torchrun (fused peer gather, then the NCCL gather for comparison), C3 / C5 group shards, C4 time shards,
# the reference arm, topology.  Everything lands in gpurun_out/m<N>_*.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/m${N}_topo.txt 2>&1; nproc >> gpurun_out/m${N}_topo.txt; numactl -H >> gpurun_out/m${N}_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err; tail -3 gpurun_out/m${N}_bench.err; tail -1 gpurun_out/m${N}_bench.json | cut -c1-300
timeout 900 $TR --master-port 29515 bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/m${N}_bench200.json 2> gpurun_out/m${N}_bench200.err; tail -1 gpurun_out/m${N}_bench200.json | cut -c1-200
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --gather nccl > gpurun_out/m${N}_bench_nccl.json 2> gpurun_out/m${N}_bench_nccl.err; tail -1 gpurun_out/m${N}_bench_nccl.json | cut -c1-200
timeout 600 $TR --master-port 29513 tools/multi_gpu_configs.py --out gpurun_out/m${N}_c3c5.json > gpurun_out/m${N}_c3c5.log 2>&1; tail -2 gpurun_out/m${N}_c3c5.log | cut -c1-400
timeout 600 $TR --master-port 29514 tools/time_shard_check.py --rows 50000000 > gpurun_out/m${N}_c4_shards.log 2>&1; tail -2 gpurun_out/m${N}_c4_shards.log | cut -c1-400
timeout 300 python bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/m${N}_bench_ref.json 2>&1; tail -1 gpurun_out/m${N}_bench_ref.json | cut -c1-300
