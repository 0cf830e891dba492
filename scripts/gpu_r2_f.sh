#!/bin/bash
# launch list of the broadphase kernels only
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pair|cluster|kdop|heap|leaf_rec|face_aabb|uncertain|morton|centroid|Radix|adjacency|emit|active_list|Scan" -c 600 --csv --log-file gpurun_out/launches_bp.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_bp.csv 70 2>/dev/null | head -40
