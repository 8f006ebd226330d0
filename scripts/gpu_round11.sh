#!/bin/bash
# sparse full size: chunk size of the walk x bound above which a node takes the chunked path
mkdir -p gpurun_out
S=gpurun_out/r2_sparse_chunk_sweep.txt
: > $S
for wm in 2048 512; do
for ch in 256 512 1024 2048; do
  echo "== 2 trees ETGPU_WIDE_CHUNK=$ch ETGPU_SPARSE_WIDE_MIN=$wm" >> $S
  ETGPU_WIDE_CHUNK=$ch ETGPU_SPARSE_WIDE_MIN=$wm timeout 300 python scripts/sparse_full.py 2 2>&1 | grep "trees built" | cut -c1-200 >> $S
done
done
for wm in 2048 512; do
  echo "== 16 trees ETGPU_WIDE_CHUNK=512 ETGPU_SPARSE_WIDE_MIN=$wm" >> $S
  ETGPU_WIDE_CHUNK=512 ETGPU_SPARSE_WIDE_MIN=$wm timeout 300 python scripts/sparse_full.py 16 2>&1 | grep "trees built" | cut -c1-200 >> $S
done
cat $S
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "csc or sparse" > gpurun_out/r2_tests13.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests13.log
tail -3 gpurun_out/r2_tests13.log | cut -c1-200
