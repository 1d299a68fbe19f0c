#!/bin/bash
# A/B timing of every eirgrid_b200/libeg_*.so variant (and the default library)
mkdir -p gpurun_out
EIRGRID_LIB_NAME=libeirgrid_b200.so python scripts/ab_time.py 65536 7 2>&1 | tail -2 | tee gpurun_out/ab.log
for lib in $(ls eirgrid_b200/libeg_*.so | xargs -n1 basename); do
  EIRGRID_LIB_NAME=$lib python scripts/ab_time.py 65536 7 2>&1 | tail -2
done | tee -a gpurun_out/ab.log
