#!/bin/bash
cp collisiondetection_b200/libccd_b200.so /tmp/orig.so
cp "$1" collisiondetection_b200/libccd_b200.so
python bench.py --steps 2 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], {k: round(v,3) for k,v in d['stages_ms'].items() if k.startswith('np')}, d['counts']['hits'])"
cp /tmp/orig.so collisiondetection_b200/libccd_b200.so
