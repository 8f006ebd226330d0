#!/bin/bash
# 8 GPUs of one box: strong scaling of the headline workload, the large-table config sharded over the GPUs, and the
# one-process group check
set -x
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.log 2>&1; echo "bench mnist rc=$?"
python scripts/bench_summary.py gpurun_out/r2_bench_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --config large --trees 128 --steps 1 --warmup 1 > gpurun_out/r2_bench_large_n$N.log 2>&1; echo "bench large rc=$?"
python scripts/bench_summary.py gpurun_out/r2_bench_large_n$N.log || tail -20 gpurun_out/r2_bench_large_n$N.log
timeout 300 python scripts/multi_check.py $N > gpurun_out/r2_multi_check_n$N.log 2>&1; echo "multi_check rc=$?"; tail -4 gpurun_out/r2_multi_check_n$N.log
