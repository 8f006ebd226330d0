#!/bin/bash
# needs 2+ GPUs: gpurun --gpus 2 -- bash scripts/gpu_multi.sh
set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_multi.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests_multi.log
tail -4 gpurun_out/r2_tests_multi.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/r2_dist_check.log 2>&1; echo "dist_check rc=$?"; tail -6 gpurun_out/r2_dist_check.log | cut -c1-400
timeout 300 python scripts/multi_check.py $N > gpurun_out/r2_multi_check.log 2>&1; echo "multi_check rc=$?"; tail -5 gpurun_out/r2_multi_check.log | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --extra none --no-cpu-baseline > gpurun_out/r2_bench_n1.log 2>&1; tail -c 400 gpurun_out/r2_bench_n1.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.log 2>&1; echo "bench rc=$?"; tail -c 2500 gpurun_out/r2_bench_n$N.log
