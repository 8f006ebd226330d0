#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests5.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests5.log
tail -4 gpurun_out/r2_tests5.log | cut -c1-300
{
for v in 0 1; do echo "== reg 64 trees, ETGPU_NO_PARK=$v"; ETGPU_NO_PARK=$v ETGPU_TIMING=1 timeout 600 python scripts/one_build.py reg 64 2 2>&1 | tail -2; done
for v in 0 1; do echo "== mnist 500 trees, chunked path for every node > 2048 rows, ETGPU_NO_BULK=$v"; ETGPU_WIDE_MIN=2048 ETGPU_NO_BULK=$v ETGPU_TIMING=1 timeout 300 python scripts/one_build.py mnist 500 3 2>&1 | tail -2; done
for v in 0 1; do echo "== predict mnist, ETGPU_NO_BULK=$v"; ETGPU_NO_BULK=$v timeout 300 python scripts/predict_once.py mnist 500 10 2>&1 | tail -1; done
} > gpurun_out/r2_ab.log 2>&1
cat gpurun_out/r2_ab.log
for v in 0 1; do ETGPU_NO_STAGER=$v timeout 600 python bench.py --steps 5 --warmup 3 --extra none --no-cpu-baseline > gpurun_out/r2_bench_stager$v.log 2>&1; python - <<PY
import json
l = json.loads(open('gpurun_out/r2_bench_stager$v.log').read().strip().splitlines()[-1])
print('NO_STAGER=$v value', round(l['value']), 'e2e', l['e2e'])
PY
done
