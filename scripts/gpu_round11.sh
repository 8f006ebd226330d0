#!/bin/bash
# tuning sweep (rebuilds on the box): lane class variants and the residency of the coded 128-thread team
mkdir -p gpurun_out
{
for v in 2 4 8; do echo "== ETGPU_LANE_SMALL_NW=$v"; ETGPU_LANE_SMALL_NW=$v timeout 300 python scripts/one_build.py mnist 500 4 2>&1 | tail -3 | cut -c1-120; done
for c in 5 6 3; do
  touch lamp_b200/csrc/node_inst.cu
  make -C lamp_b200/csrc -j16 -s EXTRA="-DMID_CODED_CTAS=$c" > /dev/null 2>&1
  echo "== MID_CODED_CTAS=$c"; timeout 300 python scripts/one_build.py mnist 500 4 2>&1 | tail -3 | cut -c1-120
done
for c in 10; do
  touch lamp_b200/csrc/node_inst.cu
  make -C lamp_b200/csrc -j16 -s EXTRA="-DLANE_SMALL_CTAS=$c" > /dev/null 2>&1
  echo "== LANE_SMALL_CTAS=$c"; timeout 300 python scripts/one_build.py mnist 500 4 2>&1 | tail -3 | cut -c1-120
done
} > gpurun_out/r2_tuning_sweep.log 2>&1
cat gpurun_out/r2_tuning_sweep.log
