#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "csc or sparse" 2>&1 | tail -2 | cut -c1-200; done
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "csc_unsorted or csc_input_builds" > gpurun_out/r2_san_sparse.log 2>&1
grep -a -A14 "Invalid\|========= ERROR\|out of bounds" gpurun_out/r2_san_sparse.log | head -80 | cut -c1-220
tail -5 gpurun_out/r2_san_sparse.log | cut -c1-200
