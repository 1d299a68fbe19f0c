#!/bin/bash
# in-order device update: parity tests, then wall times host vs device
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_update_inorder.py -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_update_inorder.log
timeout 600 python scripts/update_device_time.py 65536 2>&1 | tail -8 | tee gpurun_out/update_device_time.log
