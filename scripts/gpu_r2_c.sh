#!/bin/bash
# launch list of one C5 bench run (per-kernel device time), summarised
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_r2.csv 70 | head -70
