#!/bin/bash
# full GPU suite after the allocation / clean-up changes; sparse full size with and without the team hand-off;
# ncu capture of the FP64 small-node kernels at a deep level of the 1M x 100 regression build
mkdir -p gpurun_out /tmp/rep
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests11.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests11.log
tail -4 gpurun_out/r2_tests11.log | cut -c1-300
S=gpurun_out/r2_sparse_full_size.txt
: > $S
for tm in 296 0; do
  echo "== 2 trees, ETGPU_TEAM_MAX=$tm" >> $S
  ETGPU_TEAM_MAX=$tm timeout 300 python scripts/sparse_full.py 2 2>&1 | tail -4 | cut -c1-250 >> $S
done
cat $S
name=lane_reg
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_lane|k_node' -s 110 -c 4 -o /tmp/rep/$name -f python scripts/one_build.py reg 16 1 > gpurun_out/r2_ncu_$name.log 2>&1
python scripts/ncu_summary.py /tmp/rep/$name.ncu-rep > gpurun_out/r2_ncu_${name}_summary.txt 2>&1
python scripts/ncu_lines.py /tmp/rep/$name.ncu-rep 40 > gpurun_out/r2_ncu_${name}_lines.txt 2>&1
rm -f /tmp/rep/$name.ncu-rep
grep -a "^==\|duration\|warps_active\|stall" gpurun_out/r2_ncu_${name}_summary.txt | head -60
