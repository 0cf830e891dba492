#!/bin/bash
# round 2, call A: parity tests, then C5 bench with the new warp-cooperative traversal and (A/B) the old one
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
CCD_NP_TRACE=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_trace.log 2> gpurun_out/bench_trace.err; tail -c 600 gpurun_out/bench_trace.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_new.log 2> gpurun_out/bench_new.err; tail -c 3000 gpurun_out/bench_new.log
CCD_TRAVERSE_OLD=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_old.log 2> gpurun_out/bench_old.err; tail -c 1200 gpurun_out/bench_old.log
