#!/bin/bash
# ncu launch list (durations) of the final bench command, headline config only
mkdir -p gpurun_out
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/r2_launches_final.csv \
    python bench.py --steps 2 --warmup 1 --extra none --no-cpu-baseline > gpurun_out/r2_bench_under_ncu_final.log 2>&1
echo "rc=$?"; wc -l gpurun_out/r2_launches_final.csv
