#!/bin/bash
# BASELINE configs[2]: 10 seeds x {16, 256, 4096, 65536} in flight x 100,000 iterations under both rules; then the C++ driver's
# 100K / 1M runs (device batch rule + replay tail under the per-episode rule on the GPU)
mkdir -p gpurun_out /tmp/egcache; echo "{}" > /tmp/egcache/location_analysis.json
timeout 1500 python scripts/config3_distribution.py ${1:-100000} ${2:-10} gpurun_out/config3_distribution.json 2>&1 | tail -30 | tee gpurun_out/config3_distribution.log
for n in 100000 1000000; do
  rm -rf /tmp/ck_$n
  ( time host/_build/eirgrid_host --assets tests/golden/ireland_map --no-continue --master-seed 20250101 -n $n -c /tmp/ck_$n -C /tmp/egcache ) 2>&1 | tail -8 | tee gpurun_out/host_run_$n.log
done
rm -rf /tmp/ck_seq
( time host/_build/eirgrid_host --assets tests/golden/ireland_map --no-continue --master-seed 20250101 -n 1000000 -c /tmp/ck_seq -C /tmp/egcache --update-mode sequential ) 2>&1 | tail -8 | tee gpurun_out/host_run_1000000_sequential.log
