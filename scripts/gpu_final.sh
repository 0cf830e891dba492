#!/bin/bash
# round-end check: parity tests, smoke, default bench line, launch list, ncu of the largest kernels (text summaries only)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_n1.json 2>gpurun_out/bench_n1.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['gpu_launches'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_c5.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none -k regex:"solve_kernel|traverse_kernel|np_ve_kernel|emit_sort_kernel|np_combine_kernel|exact_pairs" --launch-skip 72 -c 24 -o /tmp/top_c5 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_np.log 2>&1
python scripts/ncu_summary.py /tmp/top_c5.ncu-rep > gpurun_out/ncu_top_summary.txt 2>&1
grep -c "^==" gpurun_out/ncu_top_summary.txt
