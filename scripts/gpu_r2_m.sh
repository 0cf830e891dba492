#!/bin/bash
# round-2 checkpoint: parity tests, smoke, default bench line, launch list, full ncu of one whole C5 step (text summaries only)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_n1.json 2>gpurun_out/bench_n1.err; tail -c 400 gpurun_out/bench_n1.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['gpu_launches'], json.dumps(d.get('stages_ms')))"
CCD_NP_TRACE=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_trace.log 2> gpurun_out/bench_trace.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c5.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_l.log 2>&1
read SKIP COUNT < <(python - <<'PY'
import csv, io
rows = [l for l in open('gpurun_out/launches_c5.csv') if l.startswith('"')]
r = list(csv.DictReader(io.StringIO(''.join(rows))))
starts = [i for i, x in enumerate(r) if 'hash_ints_kernel' in x['Kernel Name']]
print(starts[3], starts[4] - starts[3])
PY
)
echo "step 4 = launches $SKIP .. +$COUNT"
timeout 1200 ncu --set full --clock-control none --import-source on --launch-skip $SKIP -c $COUNT -o /tmp/step_c5 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_np.log 2>&1
python scripts/ncu_summary.py /tmp/step_c5.ncu-rep > gpurun_out/ncu_step_summary.txt 2>&1
ls -la /tmp/step_c5.ncu-rep; grep -c "^==" gpurun_out/ncu_step_summary.txt
