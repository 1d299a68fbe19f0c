#!/bin/bash
# multi-GPU bench (weak scaling) through torchrun, as the driver launches it, plus the single-process C++ host on N GPUs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 50 --warmup 3 2>gpurun_out/bench_${N}gpu_err.log | tee gpurun_out/bench_${N}gpu.json
tail -3 gpurun_out/bench_${N}gpu_err.log
make -C host >/dev/null && mkdir -p /tmp/ck /tmp/cache && echo '{}' > /tmp/cache/location_analysis.json
DEV=$(seq -s, 0 $((N-1)))
./host/_build/eirgrid_host --assets tests/golden/ireland_map -c /tmp/ck -C /tmp/cache --no-continue --master-seed 5 -i 100000000 \
    -n $((65536*N*40)) --devices $DEV | tail -2 | tee gpurun_out/host_${N}gpu.log
