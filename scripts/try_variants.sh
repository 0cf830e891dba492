#!/bin/bash
# usage: try_variants.sh <so> ... : narrowphase trace of the C5 bench with alternative builds of the library
cp collisiondetection_b200/libccd_b200.so /tmp/orig.so
for v in "$@"; do
  cp "$v" collisiondetection_b200/libccd_b200.so
  echo "== $v"
  CCD_NP_TRACE=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep "np trace" | tail -2 | cut -c1-260
done
cp /tmp/orig.so collisiondetection_b200/libccd_b200.so
