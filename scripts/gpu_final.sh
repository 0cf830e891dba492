#!/bin/bash
# round-2 final evidence at N=1: tests, smoke, default bench line, reference arm, configs C1-C4, drivers, sanitizers
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_n1.json 2>gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.err
python scripts/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; tail -2 gpurun_out/configs.err
python scripts/bench_drivers.py > gpurun_out/drivers.jsonl 2> gpurun_out/drivers.err; tail -2 gpurun_out/drivers.err; cut -c1-400 gpurun_out/drivers.jsonl
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize run ok" gpurun_out/sanitizer_$tool.log | head -4
done
python -c "
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value'] if 'cpu_baseline' in d else None)
d=json.loads(open('gpurun_out/bench_ref.json').read().strip().splitlines()[-1]); print(d['value'], d['steps'], d['config']['workload'])"
