#!/bin/bash
mkdir -p gpurun_out
for lib in $(ls eirgrid_b200/libeg_*.so | xargs -n1 basename); do
  EIRGRID_LIB_NAME=$lib python scripts/ab_time.py 65536 7 2>&1 | tail -1
done | tee gpurun_out/ab.log
