"""BASELINE.json configs[4] shape on ONE GPU, bounded: an n x 256 FP64 table (default 10M rows = 20.5 GB, more than
2^31 elements) is uploaded through et_data_dense_alloc + et_data_dense_colblock in blocks of whole columns
generated on the fly (the host never holds the table), a few trees are built, and the forest is checked on a
sample of rows regenerated from the same seeds: every fully grown tree (nMin=2, continuous features) must
reproduce the training label of its own rows.

    python scripts/large_table.py [rows] [trees]
"""
import os, sys, time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lamp_b200 as et

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 4
d, blk, seed = 256, 16, 5


def block(b):  # columns [b*blk, (b+1)*blk) as [blk][n], iid N(0,1) (float32 draws widened to FP64), reproducible per block
    return np.random.default_rng([seed, b]).standard_normal(size=(blk, n), dtype=np.float32).astype(np.float64)


def labels(x):  # 3-level planted tree on 8 features of the first block + 5 % flips
    y = np.where(x[0] > 0, np.where(x[1] > 0.3, x[2] > -0.2, x[3] > 0.1), np.where(x[4] > -0.4, x[5] > 0.2, x[6] > 0.0))
    flip = np.random.default_rng([seed, 999]).random(n) < 0.05
    return (y ^ flip).astype(np.int32)


t0 = time.perf_counter()
ctx = et.Context(0)
rows = np.sort(np.random.default_rng(7).choice(n, size=20000, replace=False))  # rows kept on the host for the check
xs = np.empty((len(rows), d))
state = {}


def blocks():
    for b in range(d // blk):
        x = block(b)
        if b == 0:
            state["y"] = labels(x)
        xs[:, b * blk:(b + 1) * blk] = x[:, rows].T
        yield b * blk, x


dd = et.DeviceData.from_column_blocks(n, d, blocks(), ctx)
y = state["y"]
dd.set_target_classification(y, 2)
t1 = time.perf_counter()
f = et.buildForestClassification(dd, None, None, 2, 2, 16, m, 8, seed=1, ctx=ctx)
t2 = time.perf_counter()
print("table %d x %d (%.1f GB, %.2e elements) generated + uploaded in %.1f s; %d trees built in %.1f s (%.2f trees/s), "
      "%d nodes, %d levels, gpu %.0f ms" % (n, d, n * d * 8 / 1e9, float(n) * d, t1 - t0, m, t2 - t1, m / (t2 - t1),
                                            f.stats["nodes"], f.stats["levels"], f.stats["gpu_ms"]))
pred = et.predictClassification(f, xs, ctx=ctx)
acc = (pred.argmax(1) == y[rows]).mean()
one = et.predictClassification(et.Forest.from_trees([f.flat(0)], ctx=ctx), xs, ctx=ctx)
acc1 = (one.argmax(1) == y[rows]).mean()
print("training rows reproduced: forest %.4f, first tree alone %.4f" % (acc, acc1))
assert acc1 == 1.0 and acc == 1.0
print("large_table ok")
