#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 600 python scripts/update_device_time.py 65536 2>&1 | tail -8 | tee gpurun_out/update_device_time.log
