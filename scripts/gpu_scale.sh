#!/bin/bash
# scaling check: bench.py at N GPUs (N = $1), JSON line to gpurun_out/bench_n$1.json
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 800 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['e2e']['ms_per_step'], d['value'])
print(json.dumps(d.get('stages_ms_per_rank')))
print(json.dumps({k:v for k,v in d.items() if k in ('stencils_per_rank','vf_rank0','ee_rank0','host_ms','exchange_ms')}))
PY
