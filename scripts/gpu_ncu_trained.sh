#!/bin/bash
set -x
TAG=${1:-trained}
mkdir -p gpurun_out
python scripts/profile_rollout_trained.py 65536 3 > gpurun_out/plain_trained.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eg_episode_kernel -s 3 -c 1 -o gpurun_out/rollout_$TAG \
    python scripts/profile_rollout_trained.py 65536 3 > gpurun_out/ncu_trained.log 2>&1
tail -n 2 gpurun_out/plain_trained.log; tail -n 2 gpurun_out/ncu_trained.log
