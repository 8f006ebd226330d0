#!/bin/bash
# first GPU pass of the round: parity suite, then timing of the three table shapes
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests.log
tail -5 gpurun_out/r2_tests.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_mnist.log 2>&1; tail -c 3000 gpurun_out/r2_bench_mnist.log
ETGPU_WIDE_MIN=2147483647 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_mnist_onecta.log 2>&1; tail -c 600 gpurun_out/r2_bench_mnist_onecta.log
ETGPU_TIMING=2 timeout 300 python scripts/one_build.py mnist 500 2 > gpurun_out/r2_timing_mnist.log 2>&1; tail -3 gpurun_out/r2_timing_mnist.log
ETGPU_TIMING=2 timeout 600 python scripts/one_build.py reg 64 2 > gpurun_out/r2_timing_reg.log 2>&1; tail -3 gpurun_out/r2_timing_reg.log
timeout 900 python scripts/large_table.py 10000000 8 > gpurun_out/r2_large.log 2>&1; tail -4 gpurun_out/r2_large.log
