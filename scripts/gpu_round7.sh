#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python scripts/sparse_full.py 2 > gpurun_out/r2_sparse_full.log 2>&1; echo "sparse rc=$?"; tail -6 gpurun_out/r2_sparse_full.log | cut -c1-300
timeout 900 python tests/pmlb_sweep.py > gpurun_out/r2_pmlb_sweep.txt 2>&1; echo "pmlb rc=$?"; tail -8 gpurun_out/r2_pmlb_sweep.txt | cut -c1-200
