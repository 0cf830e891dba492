#!/bin/bash
# usage: try_lib.sh <variant.so> : run the cloth355 bench with an alternative build of the library
cp collisiondetection_b200/libccd_b200.so /tmp/orig.so
cp "$1" collisiondetection_b200/libccd_b200.so
python bench.py --workload cloth355 --steps 3 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], {k: round(v,3) for k,v in d['stages_ms'].items() if k.startswith('np')})"
cp /tmp/orig.so collisiondetection_b200/libccd_b200.so
