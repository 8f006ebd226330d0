#!/bin/bash
# sparse: lockstep searches (col_at4) + warp-cooperative range search: parity, full-size time, launch window
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -x -q -k "csc or sparse or lane_classes or replay_reg or regression" > gpurun_out/r2_tests12.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests12.log
tail -4 gpurun_out/r2_tests12.log | cut -c1-300
S=gpurun_out/r2_sparse_full_size_b.txt
: > $S
for m in 2 16; do
  echo "== $m trees" >> $S
  timeout 600 python scripts/sparse_full.py $m 2>&1 | tail -4 | cut -c1-250 >> $S
done
cat $S
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 500 --csv --log-file gpurun_out/r2_sparse_launches_b.csv python scripts/sparse_full.py 2 > gpurun_out/r2_sparse_ncu_b.log 2>&1
python - <<'PY'
import csv, collections
f = "gpurun_out/r2_sparse_launches_b.csv"
rows = [r for r in csv.reader(open(f, errors="ignore")) if len(r) > 10 and r[0].isdigit()]
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows:
    name = r[4].split("(")[0][-60:]
    tot[name] += float(r[-1].replace(",", "")) / 1e6; cnt[name] += 1
print(f, "launches", len(rows))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:12]:
    print("  %-62s n=%4d total %.3f ms  avg %.4f ms" % (k, cnt[k], v, v / cnt[k]))
PY
