#!/bin/bash
cp collisiondetection_b200/libccd_b200.so /tmp/orig.so
for v in var_tmp/lib_*.so; do
  cp "$v" collisiondetection_b200/libccd_b200.so
  echo "== $v"
  CCD_NP_TRACE=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity 2>&1 | grep "np trace" | tail -2 | cut -c1-330
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step', round(d['ms_per_step'],3))"
done
cp /tmp/orig.so collisiondetection_b200/libccd_b200.so
