"""One resident build of a bench workload (for ncu captures): python scripts/one_build.py [config] [trees] [builds]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import lamp_b200 as et
cfg = dict(bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "mnist"])
m = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["trees"]
builds = int(sys.argv[3]) if len(sys.argv) > 3 else 1
x, y = bench.make_host_data(cfg)
ctx = et.Context(0)
dd = et.DeviceData.from_rowmajor(x, ctx)
if cfg["task"] == "cls":
    dd.set_target_classification(y, cfg["C"])
else:
    dd.set_target_regression(y)
for b in range(builds):
    if cfg["task"] == "cls":
        f = et.buildForestClassification(dd, None, None, cfg["C"], cfg["n_min"], cfg["k"], m, 8, seed=7 + b, ctx=ctx)
    else:
        f = et.buildForestRegression(dd, None, cfg["n_min"], cfg["k"], m, 8, seed=7 + b, ctx=ctx)
    print({k: f.stats[k] for k in ("gpu_ms", "gpu_ms_split", "gpu_ms_partition", "nodes", "levels", "launches")})
