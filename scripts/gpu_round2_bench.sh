#!/bin/bash
# GPU tests, headline bench, suitability bench (1 GPU)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_latest.json
tail -3 gpurun_out/bench_err.log
python bench.py --workload suitability --steps 5 --warmup 3 2>gpurun_out/bench_suit_err.log | tee gpurun_out/bench_suitability.json
tail -3 gpurun_out/bench_suit_err.log
