#!/bin/bash
# 8-GPU box: config 4 (fixed total of 1,000,000 episodes on the 10x grid) and config 5 (suitability, sharded by site) at N = 2, 4, 8,
# the headline workload at N = 8, and the 2-GPU host-driver test
mkdir -p gpurun_out
run() {  # run <n> <tag> <bench args...>
  local n=$1 tag=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" \
    2>gpurun_out/${tag}_${n}gpu.err | tail -1 | tee gpurun_out/${tag}_${n}gpu.json | cut -c1-400
}
nvidia-smi -L | head -8
python -m pytest tests/test_host_driver.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_host_multi.log
for n in 2 4 8; do run $n r02_bench_scaled10_total1m --workload scaled10 --total-episodes 1000000 --steps 8 --warmup 3 --no-cpu-baseline; done
for n in 2 4 8; do run $n r02_bench_suitability --workload suitability --steps 10 --warmup 3 --no-cpu-baseline; done
run 8 r02_bench --steps 20 --warmup 3 --no-cpu-baseline
