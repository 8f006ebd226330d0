#!/bin/bash
# last check of the round: whole GPU suite at HEAD, bench line with the sparse extra (16-tree full-size entry)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_final2.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests_final2.log
tail -3 gpurun_out/r2_tests_final2.log | cut -c1-200
timeout 400 python bench.py --steps 5 --warmup 3 --extra sparse > gpurun_out/r2_bench_final2.log 2> gpurun_out/r2_bench_final2.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2_bench_final2.err
python scripts/bench_summary.py gpurun_out/r2_bench_final2.log | cut -c1-600
