#!/bin/bash
# round-2 GPU call A: new group-plan tests, full GPU suite, bench (default + staging thread sweep)
set -x
mkdir -p gpurun_out
nproc > gpurun_out/a_nproc.txt; lscpu | head -30 >> gpurun_out/a_nproc.txt; nvidia-smi topo -m >> gpurun_out/a_nproc.txt 2>&1
timeout 600 python -m pytest tests/test_group_plan_gpu.py -x -q -m gpu -s > gpurun_out/a_plan.log 2>&1; echo "plan rc=$?" >> gpurun_out/a_plan.log
tail -15 gpurun_out/a_plan.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -15 gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; tail -3 gpurun_out/a_bench.err; cat gpurun_out/a_bench.json
for t in 2 4 12 16; do
  B200OLS_STAGE_THREADS=$t timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_t$t.json 2> gpurun_out/a_bench_t$t.err
done
B200OLS_STAGE_THREADS=8 B200OLS_STAGE_SLOT_MB=2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_t8_s2.json 2>&1
B200OLS_STAGE_THREADS=8 B200OLS_STAGE_SLOT_MB=8 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_t8_s8.json 2>&1
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/a_bench_ref.json 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/a_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value',d.get('value'),'e2e',d.get('e2e',{}).get('value'),'api',d.get('e2e_api',{}).get('value'), d.get('e2e_api',{}).get('ms_per_step'), d.get('e2e_api',{}).get('plan_device_ms'), d.get('e2e_api',{}).get('shuffled_rows'))
    except Exception as e:
        print(f,'ERR',e)
PY
