"""One process, several GPUs (et_init_multi): python scripts/multi_check.py [n_gpus]
Builds a forest on the group and on one GPU (identical trees), predicts with both (within 1e-12), prints timings."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lamp_b200 as et

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "mnist_test_u8.npz"))
x, y = z["pixels"].astype(np.float64), z["label"].astype(np.int32)
group, one = et.Context.multi(list(range(G))), et.Context(0)
m = 64
for ctx, name in ((one, "1 GPU"), (group, "%d GPUs" % G)):
    dd = et.DeviceData.from_rowmajor(x, ctx)
    dd.set_target_classification(y, 10)
    et.buildForestClassification(dd, None, None, 10, 2, 28, m, 8, seed=1, ctx=ctx)
    t0 = time.perf_counter()
    f = et.buildForestClassification(dd, None, None, 10, 2, 28, m, 8, seed=2, ctx=ctx)
    t1 = time.perf_counter()
    p = et.predictClassification(f, x, ctx=ctx)
    t2 = time.perf_counter()
    print("%s: build %d trees %.1f ms (gather %.2f ms), predict %d rows %.1f ms" %
          (name, m, 1e3 * (t1 - t0), ctx.comm_last_ms(), len(x), 1e3 * (t2 - t1)))
    if ctx is one:
        f1, p1 = f, p
for t in range(m):
    assert np.array_equal(f.flat(t).feature, f1.flat(t).feature) and np.array_equal(f.flat(t).leaf, f1.flat(t).leaf)
np.testing.assert_allclose(p, p1, rtol=1e-12, atol=1e-300)
print("multi_check ok")
