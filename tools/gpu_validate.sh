#!/bin/bash
# One-GPU validation of a build:   gpurun -- 'bash tools/gpu_validate.sh'
# GPU test suite, smoke(), both bench arms, every BASELINE.json config, the launch list of the bench command and the
# sanitizer passes over the newest kernels.  Everything lands in gpurun_out/v_*.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/v_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/v_pytest.log; tail -3 gpurun_out/v_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/v_smoke.log 2>&1; tail -1 gpurun_out/v_smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/v_bench_ref.json 2> gpurun_out/v_bench_ref.err; tail -1 gpurun_out/v_bench_ref.json | cut -c1-250
timeout 900 python bench.py > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; tail -1 gpurun_out/v_bench.json | cut -c1-400
for c in C1 C3 C4 C4rls C5; do
  timeout 900 python bench.py --config $c > gpurun_out/v_config_$c.json 2> gpurun_out/v_config_$c.err
  tail -1 gpurun_out/v_config_$c.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$c', d['value'], d['unit'], 'ms', d['ms_per_step'], 'frac', d.get('roofline',{}).get('frac'), 'e2e', d.get('e2e',{}).get('value'), 'cpu', d.get('cpu_baseline',{}).get('value'), 'err', d.get('max_rel_err'))" || tail -3 gpurun_out/v_config_$c.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; grep -c . gpurun_out/v_launches_bench.csv
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool python -m pytest tests/test_cd_thread_gpu.py -x -q -m gpu -k "thread_per_group" > gpurun_out/v_sanitizer_$tool.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/v_sanitizer_$tool.log | tail -3
done
