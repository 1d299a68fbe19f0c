#!/bin/bash
# ncu launch list + full capture of the rollout kernel at the bench's own size (after a plain run of the same command exits 0)
set -x
mkdir -p gpurun_out
TAG=${1:-latest}
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eg_episode_kernel -s 3 -c 1 -o gpurun_out/rollout_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
