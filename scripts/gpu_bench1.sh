#!/bin/bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench_r01_first.json
tail -5 gpurun_out/bench_err.log
