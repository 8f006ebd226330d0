#!/bin/bash
# the fresh-context regression test; launch window of the 16-tree full-size sparse build
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fresh_context" 2>&1 | tail -3 | cut -c1-200
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 400 --csv --log-file gpurun_out/r2_sparse_launches_16.csv python scripts/sparse_full.py 16 > gpurun_out/r2_sparse_ncu_16.log 2>&1
python - <<'PY'
import csv, collections
f = "gpurun_out/r2_sparse_launches_16.csv"
rows = [r for r in csv.reader(open(f, errors="ignore")) if len(r) > 10 and r[0].isdigit()]
tot = collections.defaultdict(float); cnt = collections.Counter(); grid = collections.defaultdict(list)
for r in rows:
    name = r[4].split("(")[0][-60:]
    tot[name] += float(r[-1].replace(",", "")) / 1e6; cnt[name] += 1
    grid[name].append(r[8] if len(r) > 8 else "")
print(f, "launches", len(rows))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:12]:
    print("  %-62s n=%4d total %.3f ms  avg %.4f ms  grids %s" % (k, cnt[k], v, v / cnt[k], grid[k][:4]))
PY
head -2 gpurun_out/r2_sparse_launches_16.csv | cut -c1-400
