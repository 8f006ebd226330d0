"""Readable digest of a bench.py JSON line: python scripts/bench_summary.py LOG"""
import json
import sys

l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("headline: %.0f trees/s (%.2f ms/step), e2e %s, roofline frac %.3f (kernel %.1f ms/step, share %.2f)" %
      (l["value"], l["ms_per_step"], l["e2e"], l["roofline"]["frac"], l["roofline"]["kernel_ms_per_step"],
       l["roofline"]["share_of_step"]))
p = l["predict"]
print("predict: %.2f Mrows/s (%.2f ms), e2e %s, frac %.3f, cpu %s" %
      (p["value"] / 1e6, p["ms_per_step"], p["e2e"], p["roofline"]["frac"], p.get("cpu_baseline")))
print("cpu:", l.get("cpu_baseline"))
print("collectives:", l.get("collectives"))
for k, v in l.get("configs", {}).items():
    if "error" in v:
        print(k, v)
        continue
    print("%s: build %.2f trees/s (%.0f ms, %d trees), e2e %s, frac %.3f, predict %.2f Mrows/s (frac %.3f) on %d rows, wall %.0f s" %
          (k, v["build"]["value"], v["build"]["ms_per_step"], v["trees"], v["e2e"].get("value"), v["roofline"]["frac"],
           v["predict"]["value"] / 1e6, v["predict"]["roofline"]["frac"], v["predict"]["rows"], v["wall_s"]))
    print("    cpu build:", v["cpu_baseline"])
    print("    cpu predict:", v["predict"].get("cpu_baseline"))
    if "full_size" in v:
        print("    full size:", {a: v["full_size"].get(a) for a in ("rows", "features", "stored_entries", "trees", "build", "hbm_in_use_after_build_bytes", "error")})
    print("    stats:", {a: v["stats_per_step"][a] for a in ("nodes", "levels", "parallel_sum_nodes", "ambiguous_splits")})
