#!/bin/bash
# round-2 final pass on one B200: sparse tests in a fresh process (scratch sizing regression), the whole GPU suite,
# smoke, the full-size sparse build, the full bench line (all configs)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "csc or sparse" > gpurun_out/r2_tests14.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests14.log
tail -3 gpurun_out/r2_tests14.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_final.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests_final.log
tail -3 gpurun_out/r2_tests_final.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke_final.log | cut -c1-200
S=gpurun_out/r2_sparse_full_size_final.txt
: > $S
for m in 2 16; do
  echo "== $m trees (defaults)" >> $S
  timeout 300 python scripts/sparse_full.py $m 2>&1 | tail -4 | cut -c1-250 >> $S
done
cat $S
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_final.log 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2_bench_final.err
python scripts/bench_summary.py gpurun_out/r2_bench_final.log
