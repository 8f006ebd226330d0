#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests8.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests8.log
tail -4 gpurun_out/r2_tests8.log | cut -c1-300
timeout 900 python scripts/sparse_full.py 2 > gpurun_out/r2_sparse_full2.log 2>&1; echo "sparse rc=$?"; tail -5 gpurun_out/r2_sparse_full2.log | cut -c1-300
timeout 900 python scripts/sparse_full.py 16 > gpurun_out/r2_sparse_full16.log 2>&1; echo "sparse rc=$?"; tail -5 gpurun_out/r2_sparse_full16.log | cut -c1-300
