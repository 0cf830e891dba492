#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_new.log 2> gpurun_out/bench_new.err; tail -c 300 gpurun_out/bench_new.err
python - <<'PY'
import json
p=json.loads(open('gpurun_out/bench_new.log').read().strip().splitlines()[-1])
print(p['ms_per_step'], p['e2e']['ms_per_step'], json.dumps(p['stages_ms']))
PY
