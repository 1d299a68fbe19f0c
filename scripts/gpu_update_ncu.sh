#!/bin/bash
# per-kernel times of the in-order device update (launch list) + one full capture of the table walk
mkdir -p gpurun_out
TAG=${1:-latest}
python scripts/update_device_time.py 65536 > gpurun_out/update_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:upd_ -c 120 --csv --log-file gpurun_out/update_launches_$TAG.csv \
    python scripts/update_device_time.py 65536 > gpurun_out/update_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:upd_walk -s 20 -c 1 -o gpurun_out/update_walk_$TAG \
    python scripts/update_device_time.py 65536 > gpurun_out/update_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
