#!/bin/bash
# launch list of the bench + one full ncu capture of every kernel of ONE C5 step (the 4th); only text summaries come back
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_c5.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
read SKIP COUNT < <(python - <<'PY'
import csv, io
rows = [l for l in open('gpurun_out/launches_c5.csv') if l.startswith('"')]
r = list(csv.DictReader(io.StringIO(''.join(rows))))
starts = [i for i, x in enumerate(r) if 'hash_ints_kernel' in x['Kernel Name']]
print(starts[3], starts[4] - starts[3])
PY
)
echo "step 4 = launches $SKIP .. +$COUNT"
ncu --set full --clock-control none --import-source on --launch-skip $SKIP -c $COUNT -o /tmp/step_c5 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_np.log 2>&1
python scripts/ncu_summary.py /tmp/step_c5.ncu-rep > gpurun_out/ncu_step_summary.txt 2>&1
ls -la /tmp/step_c5.ncu-rep; grep -c "^==" gpurun_out/ncu_step_summary.txt
