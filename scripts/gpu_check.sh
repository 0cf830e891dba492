#!/bin/bash
# parity tests + C5 bench + launch list (no full ncu)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench1415.log 2>gpurun_out/bench1415.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench1415.log').read().strip().splitlines()[-1])
    print(d['ms_per_step'], d['e2e']['ms_per_step'], {k: round(v,3) for k,v in d['stages_ms'].items()}, d['counts']['hits'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/bench1415.err').read()[-2000:])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c5.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_c5.csv 70 2>/dev/null | head -30
