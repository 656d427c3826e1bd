#!/bin/bash
set -x
mkdir -p gpurun_out
for k in gram_multi cd_thread predict_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/s_$k python tools/c3_once.py 2 > gpurun_out/s_ncu_$k.log 2>&1; tail -2 gpurun_out/s_ncu_$k.log
  ncu -i gpurun_out/s_$k.ncu-rep --page details --csv > gpurun_out/s_${k}_details.csv 2>/dev/null
  ncu -i gpurun_out/s_$k.ncu-rep --page raw --csv > gpurun_out/s_${k}_raw.csv 2>/dev/null
done
python bench.py --config C3 > gpurun_out/s_c3_config.json 2> gpurun_out/s_c3_config.err; tail -1 gpurun_out/s_c3_config.json | cut -c1-600
