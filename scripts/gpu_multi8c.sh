#!/bin/bash
# BASELINE configs[4] at its full size on 8 GPUs: 10M x 256, 2000 trees (250 per GPU, the forest stays sharded),
# predict on 10M held-out rows (tree-sharded, all-reduced)
set -x
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --config large --steps 1 --warmup 0 > gpurun_out/r2_bench_large_full_n$N.log 2>&1; echo "bench large rc=$?"
python scripts/bench_summary.py gpurun_out/r2_bench_large_full_n$N.log || grep -v "^\[W\|^W1" gpurun_out/r2_bench_large_full_n$N.log | grep -m3 "Error"
