#!/bin/bash
# parity tests with the default library, then A/B timing of every eirgrid_b200/libeg_*.so variant
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
EIRGRID_LIB_NAME=libeirgrid_b200.so python scripts/ab_time.py 65536 7 2>&1 | tail -1 | tee gpurun_out/ab.log
for lib in $(ls eirgrid_b200/libeg_*.so | xargs -n1 basename); do
  EIRGRID_LIB_NAME=$lib python scripts/ab_time.py 65536 7 2>&1 | tail -1
done | tee -a gpurun_out/ab.log
