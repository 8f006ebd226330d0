#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests6.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests6.log
tail -4 gpurun_out/r2_tests6.log | cut -c1-300
ETGPU_TIMING=1 timeout 600 python scripts/one_build.py reg 64 2 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_full2.log 2> gpurun_out/r2_bench_full2.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_full2.err
python scripts/bench_summary.py gpurun_out/r2_bench_full2.log
