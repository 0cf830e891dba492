#!/bin/bash
# launch list + full ncu capture of the narrowphase kernels at C5; only small text summaries are kept
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c5.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"${1:-np_|roots_kernel|stencil_resume|bucket}" --launch-skip ${2:-60} -c ${3:-20} -o /tmp/np_c5 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_np.log 2>&1
python scripts/ncu_summary.py /tmp/np_c5.ncu-rep > gpurun_out/ncu_np_summary.txt 2>&1
ls -la /tmp/np_c5.ncu-rep; tail -3 gpurun_out/ncu_np_summary.txt
