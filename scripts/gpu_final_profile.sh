#!/bin/bash
# round-2 final profile: launch list + ncu --set full of every kernel of one whole C5 step (text summaries only)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_c5.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_l.log 2>&1
read SKIP COUNT < <(python - <<'PY'
import csv, io
rows = [l for l in open('gpurun_out/launches_c5.csv') if l.startswith('"')]
r = list(csv.DictReader(io.StringIO(''.join(rows))))
starts = [i for i, x in enumerate(r) if 'hash_ints_kernel' in x['Kernel Name']]
print(starts[3], starts[4] - starts[3])
PY
)
echo "step 4 = launches $SKIP .. +$COUNT"
timeout 1500 ncu --set full --clock-control none --import-source on --launch-skip $SKIP -c $COUNT -o /tmp/step_c5 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_np.log 2>&1
python scripts/ncu_summary.py /tmp/step_c5.ncu-rep > gpurun_out/ncu_step_summary.txt 2>&1
python scripts/ncu_table.py gpurun_out/ncu_step_summary.txt > gpurun_out/ncu_step_table.txt 2>&1
tail -5 gpurun_out/ncu_step_table.txt
