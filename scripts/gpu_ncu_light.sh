#!/bin/bash
set -x
TAG=${1:-latest}
mkdir -p gpurun_out
python scripts/profile_rollout.py 16384 3 > gpurun_out/plain_light.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eg_episode_kernel -s 1 -c 1 -o gpurun_out/rollout_$TAG \
    python scripts/profile_rollout.py 16384 3 > gpurun_out/ncu_light.log 2>&1
tail -2 gpurun_out/ncu_light.log
