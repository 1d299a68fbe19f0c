#!/bin/bash
# final r02 captures: table walk (update), suitability kernel; launch lists of the update pipeline and of the suitability bench
mkdir -p gpurun_out
python scripts/update_device_time.py 65536 > gpurun_out/update_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:upd_ -c 64 --csv --log-file gpurun_out/update_launches_final.csv \
    python scripts/update_device_time.py 65536 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:upd_walk -s 20 -c 1 -o gpurun_out/update_walk_final \
    python scripts/update_device_time.py 65536 > /dev/null 2>&1
python bench.py --workload suitability --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/suit_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:eg_suit -c 20 --csv --log-file gpurun_out/suitability_launches_final.csv \
    python bench.py --workload suitability --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:eg_suitability_kernel -s 3 -c 1 -o gpurun_out/suitability_final \
    python bench.py --workload suitability --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -6
