#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 60 tools/tf32_gram_probe > gpurun_out/n_tf32_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/n_tf32_probe.log; cat gpurun_out/n_tf32_probe.log
nvidia-smi --query-gpu=name,memory.used --format=csv | head -3
bash tools/gpu_r2_m.sh 8
