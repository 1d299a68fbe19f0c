#!/bin/bash
# exchange step over NVLink peer memory vs NCCL all-gather: same weights after the same steps, timing of both (N GPUs)
N=${1:-2}
mkdir -p gpurun_out
for kind in peer nccl; do
  EIRGRID_EXCHANGE=$kind timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
    bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/exchange_${kind}_${N}gpu.err | tail -1 > gpurun_out/exchange_${kind}_${N}gpu.json
  python - "$kind" "$N" <<'PY'
import json, sys
kind, n = sys.argv[1], sys.argv[2]
try:
    d = json.load(open("gpurun_out/exchange_%s_%sgpu.json" % (kind, n)))
    c = d["config"]
    print(kind, "N=%s value %.3f M  ms/step %.4f  e2e %.3f M  exchange %s us  how: %s  sha %s" % (
        n, d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, round(c["collective"]["us_per_step"], 1), c["collective"].get("how"), c["weights_sha256_16_after_e2e"]))
except Exception as e:
    print(kind, "FAILED", e)
    print(open("gpurun_out/exchange_%s_%sgpu.err" % (kind, n)).read()[-3000:])
PY
done
