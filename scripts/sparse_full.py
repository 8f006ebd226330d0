"""BASELINE configs[3] at full size on one GPU, bounded tree count: 1M x 10000 CSC at 1 % density kept SPARSE in HBM
(et_data_csc: 12 bytes per stored entry; the dense form would be 80 GB).  python scripts/sparse_full.py [trees] [rows] [cols]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import lamp_b200 as et

m = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
t0 = time.perf_counter()
colptr, rowidx, vals, y = bench.gen_sparse_gpu(torch, n, d, 0.01, 4)  # (drawn on the GPU: 100M entries in < 1 s)
t1 = time.perf_counter()
free0, total = torch.cuda.mem_get_info()
ctx = et.Context(0)
dd = et.DeviceData.from_csc(colptr, rowidx, vals, n, d, ctx)
dd.set_target_classification(y, 2)
t2 = time.perf_counter()
f = et.buildForestClassification(dd, None, None, 2, 2, 100, m, 8, seed=1, ctx=ctx)
t3 = time.perf_counter()
free1, _ = torch.cuda.mem_get_info()
print("table %d x %d, %d stored entries (%.2f GB as CSC, %.1f GB dense) generated in %.1f s, uploaded in %.1f s" %
      (n, d, len(vals), (len(vals) * 12 + len(colptr) * 8) / 1e9, n * d * 8 / 1e9, t1 - t0, t2 - t1))
print("%d trees built in %.1f s (%.3f trees/s): %d nodes, %d levels, gpu %.0f ms; HBM in use after the build %.1f GB" %
      (m, t3 - t2, m / (t3 - t2), f.stats["nodes"], f.stats["levels"], f.stats["gpu_ms"], (free0 - free1) / 1e9))
# every fully grown tree reproduces the labels of its training rows (checked on a sample of rows, dense form)
rows = np.sort(np.random.default_rng(1).choice(n, size=2000, replace=False))
keep = np.isin(rowidx, rows)
cols = np.repeat(np.arange(d), np.diff(colptr))[keep]
xs = np.zeros((len(rows), d))
xs[np.searchsorted(rows, rowidx[keep]), cols] = vals[keep]
one = et.predictClassification(et.Forest.from_trees([f.flat(0)], ctx=ctx), xs, ctx=ctx)
acc = (one.argmax(1) == y[rows]).mean()
print("first tree reproduces the training labels of %d sampled rows: %.4f" % (len(rows), acc))
assert acc == 1.0
print("sparse_full ok")
