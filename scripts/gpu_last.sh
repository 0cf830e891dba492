#!/bin/bash
# last measurements of the round on one GPU: default bench line (with parity block and CPU baseline), narrowphase trace
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1.json 2>gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err
CCD_NP_TRACE=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_trace.log 2> gpurun_out/bench_trace.err
grep "np trace\|np counts" gpurun_out/bench_trace.err | tail -4 > gpurun_out/np_trace_final.txt
python scripts/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; tail -2 gpurun_out/configs.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['parity']['unexplained'], d['parity']['flags_differ'])"
