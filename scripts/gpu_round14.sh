#!/bin/bash
# sparse: inverse index map for the membership test of the chunk walk: parity (fresh process), multi-GPU-front test, timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -x -q -k "csc or sparse" 2>&1 | tail -3 | cut -c1-200
S=gpurun_out/r2_sparse_full_size_inv.txt
: > $S
for m in 2 16; do
  echo "== $m trees (defaults, inverse map)" >> $S
  timeout 200 python scripts/sparse_full.py $m 2>&1 | tail -4 | cut -c1-250 >> $S
done
cat $S
