#!/bin/bash
# r02 (second half): GPU tests, headline bench, launch list, full ncu capture of the rollout kernel (initial and trained table),
# 10x-grid bench lines (65,536 in flight and fixed total of 1,000,000) on one GPU
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/r02b_bench_1gpu.json; tail -2 gpurun_out/bench_err.log; cut -c1-300 gpurun_out/r02b_bench_1gpu.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02b_rollout_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python scripts/profile_rollout.py 65536 2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eg_episode_kernel -s 1 -c 1 -o gpurun_out/rollout_r02c \
    python scripts/profile_rollout.py 65536 2 > gpurun_out/ncu_full.log 2>&1
python scripts/profile_rollout_trained.py 65536 3 > gpurun_out/plain_trained.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eg_episode_kernel -s 3 -c 1 -o gpurun_out/rollout_r02c_trained \
    python scripts/profile_rollout_trained.py 65536 3 > gpurun_out/ncu_trained.log 2>&1
python bench.py --workload scaled10 --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02b_bench_scaled10_1gpu.json; cut -c1-200 gpurun_out/r02b_bench_scaled10_1gpu.json
python bench.py --workload scaled10 --total-episodes 1000000 --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02b_bench_scaled10_total1m_1gpu.json; cut -c1-200 gpurun_out/r02b_bench_scaled10_total1m_1gpu.json
ls -la gpurun_out | tail -12
