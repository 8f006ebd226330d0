"""One forest + a few predict launches (for ncu): python scripts/predict_once.py [config] [trees] [repeats]"""
import ctypes as CT
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import lamp_b200 as et
from lamp_b200 import _capi as capi

cfg = dict(bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "mnist"])
m = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["trees"]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
x, y = bench.make_host_data(cfg)
ctx = et.Context(0)
dd = et.DeviceData.from_rowmajor(x, ctx)
dd.set_target_classification(y, cfg["C"])
f = et.buildForestClassification(dd, None, None, cfg["C"], cfg["n_min"], cfg["k"], m, 8, seed=7, ctx=ctx)
xt = torch.from_numpy(x).cuda()
out = torch.empty((len(x), cfg["C"]), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st = torch.cuda.Stream()
ctx.set_stream(st.cuda_stream)
for r in range(reps + 1):
    if r == 1:
        e0.record(st)
    capi.check(capi.lib().et_predict_classification_device(ctx.h, f.h, CT.c_void_p(xt.data_ptr()), len(x), x.shape[1],
                                                           CT.c_void_p(out.data_ptr()), 0))
e1.record(st)
torch.cuda.synchronize()
print("predict %d rows x %d trees: %.3f ms per call (block budget %s MB)" %
      (len(x), m, e0.elapsed_time(e1) / reps, os.environ.get("ETGPU_PREDICT_BLOCK_MB", "24")))
