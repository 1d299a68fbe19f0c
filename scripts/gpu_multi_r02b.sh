#!/bin/bash
# r02 (second half): headline and configs[3] (fixed total of 1,000,000 episodes on the 10x grid) under torchrun at N = $1 GPUs
N=${1:-2}
mkdir -p gpurun_out
run() {  # run <tag> <bench args...>
  local tag=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N "$@" \
    2>gpurun_out/${tag}_${N}gpu.err | tail -1 | tee gpurun_out/${tag}_${N}gpu.json | cut -c1-330
}
nvidia-smi -L | head -8
run r02b_bench --steps 20 --warmup 5 --no-cpu-baseline
run r02b_bench_scaled10_total1m --workload scaled10 --total-episodes 1000000 --steps 8 --warmup 3 --no-cpu-baseline
