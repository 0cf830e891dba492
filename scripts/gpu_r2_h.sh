#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; tail -c 400 gpurun_out/bench_full.err
python - <<'PY'
import json
p=json.loads(open('gpurun_out/bench_full.log').read().strip().splitlines()[-1])
print(p['ms_per_step'], p['e2e']['ms_per_step'], json.dumps(p.get('parity')), json.dumps(p.get('cpu_baseline')))
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.err; cut -c1-1500 gpurun_out/bench_ref.log
python scripts/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; tail -c 500 gpurun_out/configs.err; cut -c1-400 gpurun_out/configs.jsonl
nproc; lscpu | head -20
