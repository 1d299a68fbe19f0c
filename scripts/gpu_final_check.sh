#!/bin/bash
# what the driver runs at round end, in one go: GPU tests, smoke(), both arms of the bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_err.log | tee gpurun_out/bench_latest.json | cut -c1-300
tail -2 gpurun_out/bench_err.log
python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tee gpurun_out/bench_reference.json | cut -c1-200
