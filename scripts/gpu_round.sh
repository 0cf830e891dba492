#!/bin/bash
# one GPU call: parity tests, default bench, launch list and a full ncu capture of the narrowphase kernels (C5)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench1415.log 2>gpurun_out/bench1415.err; tail -c 1500 gpurun_out/bench1415.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c5.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'np_|roots_kernel|stencil_resume|bucket' --launch-skip 60 -c 24 -o gpurun_out/np_c5b -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_np.log 2>&1
ls -la gpurun_out | tail -5
