#!/bin/bash
mkdir -p gpurun_out
for wm in 2048 256; do
  ETGPU_SPARSE_WIDE_MIN=$wm ETGPU_TIMING=2 timeout 900 python scripts/sparse_full.py 2 > gpurun_out/r2_sparse_t_$wm.out 2> gpurun_out/r2_sparse_t_$wm.err
  echo "== SPARSE_WIDE_MIN=$wm"; tail -3 gpurun_out/r2_sparse_t_$wm.out | head -1; grep "timing ms" gpurun_out/r2_sparse_t_$wm.err
  grep "etgpu level" gpurun_out/r2_sparse_t_$wm.err | awk '{t=0; for(i=1;i<=NF;i++){ if($i ~ /^(lane|mid|cta|wide)=/){split($i,a,"="); t+=a[2]} } print t, $0}' | sort -n -r | head -6 | cut -c1-260
  grep "etgpu level" gpurun_out/r2_sparse_t_$wm.err | awk 'NR%200==100' | cut -c1-230
  gzip -f gpurun_out/r2_sparse_t_$wm.err
done
