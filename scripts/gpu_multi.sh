#!/bin/bash
# multi-GPU bench (weak scaling) through torchrun, as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/bench_${N}gpu_err.log | tee gpurun_out/bench_${N}gpu.json
tail -5 gpurun_out/bench_${N}gpu_err.log
