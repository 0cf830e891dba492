#!/bin/bash
# round-2 iteration check: parity tests, default bench line, per-kernel-group trace of the narrowphase
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2>gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['gpu_launches'], json.dumps(d.get('stages_ms')), json.dumps(d.get('parity')))"
CCD_NP_TRACE=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_trace.log 2> gpurun_out/bench_trace.err
grep "np trace" gpurun_out/bench_trace.err | tail -2
grep "np counts" gpurun_out/bench_trace.err | tail -2
