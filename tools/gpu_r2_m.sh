#!/bin/bash
N=${1:-2}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/m_m${N}_bench.json 2> gpurun_out/m_m${N}_bench.err; tail -3 gpurun_out/m_m${N}_bench.err | cut -c1-300; tail -1 gpurun_out/m_m${N}_bench.json | cut -c1-300
timeout 900 $TR --master-port 29542 bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/m_m${N}_bench200.json 2> gpurun_out/m_m${N}_bench200.err; tail -1 gpurun_out/m_m${N}_bench200.json | cut -c1-200
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_moving_paths_gpu.py -x -q -m gpu > gpurun_out/m_tests.log 2>&1; tail -2 gpurun_out/m_tests.log; fi
