#!/bin/bash
# N = 2: parity tests on one GPU, then the 2-rank bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; tail -c 600 gpurun_out/bench_n2.err
python - <<'PY'
import json
p=json.loads([l for l in open('gpurun_out/bench_n2.log').read().strip().splitlines() if l.startswith('{')][-1])
print(p['n_gpus'], p['ms_per_step'], p['e2e']['ms_per_step'], json.dumps(p['stages_ms_per_rank']), p['stencils_per_rank'])
PY
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_new.log 2> gpurun_out/bench_new.err; tail -c 300 gpurun_out/bench_new.err
python - <<'PY'
import json
p=json.loads(open('gpurun_out/bench_new.log').read().strip().splitlines()[-1])
print(p['ms_per_step'], p['e2e']['ms_per_step'], json.dumps(p['stages_ms']))
PY
