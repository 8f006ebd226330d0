#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests4.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests4.log
tail -6 gpurun_out/r2_tests4.log | cut -c1-300
for v in "0 0" "128 2048" "64 2048" "256 2048" "512 2048" "128 512"; do
  set -- $v
  echo "== stream ($1, $2]"
  ETGPU_STREAM_MIN=$1 ETGPU_STREAM_MAX=$2 timeout 300 python scripts/one_build.py mnist 500 4 2>&1 | tail -3
done > gpurun_out/r2_stream_sweep.log 2>&1
cat gpurun_out/r2_stream_sweep.log
ETGPU_TIMING=2 timeout 300 python scripts/one_build.py mnist 500 2 > gpurun_out/r2_timing_mnist_stream.log 2>&1; tail -2 gpurun_out/r2_timing_mnist_stream.log
