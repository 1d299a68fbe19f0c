#!/bin/bash
# first GPU contact: parity tests + a tiny timing
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
