#!/bin/bash
# 8 GPUs: strong scaling of the headline workload (after the gather fix) and BASELINE configs[4] at its full size:
# 10M x 256, 2000 trees sharded over the 8 GPUs (250 each, the forest stays sharded), predict on 10M held-out rows
set -x
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.log 2>&1; echo "bench mnist rc=$?"
python scripts/bench_summary.py gpurun_out/r2_bench_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --config large --steps 1 --warmup 0 > gpurun_out/r2_bench_large_full_n$N.log 2>&1; echo "bench large rc=$?"
python scripts/bench_summary.py gpurun_out/r2_bench_large_full_n$N.log || tail -20 gpurun_out/r2_bench_large_full_n$N.log | cut -c1-400
