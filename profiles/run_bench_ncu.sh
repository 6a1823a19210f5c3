#!/bin/bash
# round-1 profile set for the projected-map path: launch list of bench.py + one full capture of field_bin_kernel
set -e
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-render > gpurun_out/plain_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-render > gpurun_out/ncu_launches.log 2>&1
python profiles/run_bin.py > gpurun_out/plain_run_bin.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:field_bin -s 2 -c 1 -o gpurun_out/prof_r01b \
    python profiles/run_bin.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
