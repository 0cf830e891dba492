#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
CCD_NP_TRACE=1 CCD_BP_TRACE=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_trace.log 2> gpurun_out/bench_trace.err; tail -n 5 gpurun_out/bench_trace.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_new.log 2> gpurun_out/bench_new.err; tail -c 300 gpurun_out/bench_new.err
python - <<'PY'
import json
p=json.loads(open('gpurun_out/bench_new.log').read().strip().splitlines()[-1])
print(p['ms_per_step'], p['e2e']['ms_per_step'], json.dumps(p['stages_ms']), p['counts'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pair|cluster|kdop|heap|leaf_rec|face_aabb|uncertain|morton|centroid|Radix|adjacency|emit|active_list|Scan|decide|general" -c 700 --csv --log-file gpurun_out/launches_bp.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_bp.csv 70 2>/dev/null | head -24
