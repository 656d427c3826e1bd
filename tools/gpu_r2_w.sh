#!/bin/bash
# usage: bash tools/gpu_r2_w.sh N   (under gpurun --gpus N): default bench twice (device-side rendezvous before the start event) + C3 / C5
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for i in 1 2; do
  timeout 900 $TR --master-port 2955$i bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/w_m${N}_bench$i.json 2> gpurun_out/w_m${N}_bench$i.err
  tail -1 gpurun_out/w_m${N}_bench$i.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=$N run $i', round(d['value']/1e6,1), 'M', d['ms_per_step'], 'strong', round(d['strong']['value']/1e6,1), d['strong']['ms_per_step'])" || tail -5 gpurun_out/w_m${N}_bench$i.err
done
timeout 600 $TR --master-port 29557 tools/multi_gpu_configs.py --out gpurun_out/w_m${N}_c3c5.json > gpurun_out/w_m${N}_c3c5.log 2>&1; tail -3 gpurun_out/w_m${N}_c3c5.log | cut -c1-400
