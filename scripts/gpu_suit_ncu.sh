#!/bin/bash
# suitability kernel: launch list + one full capture at the bench's size
mkdir -p gpurun_out
TAG=${1:-latest}
python bench.py --workload suitability --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/suit_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eg_suitability_kernel -s 3 -c 1 -o gpurun_out/suitability_$TAG \
    python bench.py --workload suitability --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/suit_ncu_full.log 2>&1
ls -la gpurun_out | tail -3
