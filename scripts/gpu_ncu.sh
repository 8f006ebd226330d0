#!/bin/bash
# ncu evidence of the round: launch list of the bench command + full captures of the dominant kernels.
# The .ncu-rep files stay on the box (too large to bring back); their summaries come back as text.
set -x
mkdir -p gpurun_out /tmp/rep
for mb in 4 12 24 48 100000; do ETGPU_PREDICT_BLOCK_MB=$mb python scripts/predict_once.py mnist 500 5 2>&1 | tail -1; done > gpurun_out/r2_predict_blocks.log
cat gpurun_out/r2_predict_blocks.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --extra none --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
cap() {  # name, kernel regex, skip, count, command...
  name=$1; shift; rx=$1; shift; sk=$1; shift; ct=$1; shift
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $sk -c $ct -o /tmp/rep/$name -f "$@" > gpurun_out/r2_ncu_$name.log 2>&1
  python scripts/ncu_summary.py /tmp/rep/$name.ncu-rep > gpurun_out/r2_ncu_${name}_summary.txt 2>&1
  python scripts/ncu_lines.py /tmp/rep/$name.ncu-rep 40 > gpurun_out/r2_ncu_${name}_lines.txt 2>&1
  rm -f /tmp/rep/$name.ncu-rep
}
cap wide_mnist 'k_wide' 0 9 python scripts/one_build.py mnist 500 1
cap level_mnist 'k_lane|k_node' 64 8 python scripts/one_build.py mnist 500 1
cap level_reg 'k_wide|k_lane|k_node' 300 16 python scripts/one_build.py reg 32 1
cap predict_mnist 'k_predict' 0 6 python scripts/predict_once.py mnist 500 1
du -sh gpurun_out
