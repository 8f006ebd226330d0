#!/bin/bash
# lane classes with few nodes on CTA teams (ETGPU_TEAM_MAX): parity, then a sweep of the bound on the MNIST-shaped
# build (500 trees = one GPU's forest, 62 trees = one of eight shards) and the 1M x 100 regression build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane_classes or replay_mnist or wide_and_one or resident_target or free_running or shard" > gpurun_out/r2_tests10.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests10.log
tail -4 gpurun_out/r2_tests10.log | cut -c1-300
S=gpurun_out/r2_team_sweep.txt
: > $S
for tm in 0 296 1184 4736 1000000000; do
  echo "== mnist 500 trees ETGPU_TEAM_MAX=$tm" >> $S
  ETGPU_TEAM_MAX=$tm timeout 300 python scripts/one_build.py mnist 500 4 2>&1 | tail -3 | cut -c1-140 >> $S
done
for tm in 0 296 1184 1000000000; do
  echo "== mnist 62 trees ETGPU_TEAM_MAX=$tm" >> $S
  ETGPU_TEAM_MAX=$tm timeout 300 python scripts/one_build.py mnist 62 5 2>&1 | tail -3 | cut -c1-140 >> $S
done
for tm in 0 296 1184; do
  echo "== reg 64 trees ETGPU_TEAM_MAX=$tm" >> $S
  ETGPU_TEAM_MAX=$tm timeout 300 python scripts/one_build.py reg 64 3 2>&1 | tail -2 | cut -c1-140 >> $S
done
cat $S
ETGPU_LEVEL_MS=1 timeout 300 python scripts/one_build.py mnist 500 2 2> gpurun_out/r2_level_ms_mnist.log | tail -1
ETGPU_LEVEL_MS=1 ETGPU_TEAM_MAX=0 timeout 300 python scripts/one_build.py mnist 500 2 2> gpurun_out/r2_level_ms_mnist_team0.log | tail -1
ETGPU_TIMING=1 timeout 300 python tests/pmlb_sweep.py --tables cleve,dermatology,ecoli,collins --no-oracle --seeds 4 2>&1 | tail -24 | cut -c1-200
