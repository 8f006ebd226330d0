#!/bin/bash
# launch list (durations) of a window of levels of the full-size sparse build: which kernels a level's time goes to
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 500 --csv --log-file gpurun_out/r2_sparse_launches_a.csv python scripts/sparse_full.py 2 > gpurun_out/r2_sparse_ncu_a.log 2>&1
python - <<'PY'
import csv, collections
for f in ["gpurun_out/r2_sparse_launches_a.csv"]:
    rows = [r for r in csv.reader(open(f, errors="ignore")) if len(r) > 10 and r[0].isdigit()]
    tot = collections.defaultdict(float); cnt = collections.Counter()
    for r in rows:
        name = r[4].split("(")[0][-60:]
        ms = float(r[-1].replace(",", "")) / 1e6 if "nsecond" in r[-2] or True else 0
        tot[name] += ms; cnt[name] += 1
    print(f, "launches", len(rows))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:25]:
        print("  %-62s n=%4d total %.3f ms  avg %.4f ms" % (k, cnt[k], v, v / cnt[k]))
    print("  units:", rows[0][-2] if rows else None)
PY
head -3 gpurun_out/r2_sparse_launches_a.csv | cut -c1-300
