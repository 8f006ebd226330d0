#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests3.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests3.log
tail -8 gpurun_out/r2_tests3.log | cut -c1-400
for v in default 2147483647; do
  if [ $v = default ]; then unset ETGPU_WIDE_MIN; else export ETGPU_WIDE_MIN=$v; fi
  timeout 300 python scripts/one_build.py mnist 500 4 > gpurun_out/r2_mnist_wide_$v.log 2>&1; tail -3 gpurun_out/r2_mnist_wide_$v.log
  ETGPU_LEVEL_MS=1 timeout 300 python scripts/one_build.py mnist 500 2 > gpurun_out/r2_mnist_levelms_$v.log 2>&1
done
unset ETGPU_WIDE_MIN
