"""Multi-GPU check on real GPUs, one process per GPU (NCCL inside libetgpu.so):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py
Every rank builds its shard of a forest, the serialized trees are all-gathered in the library, and the result is
compared with the forest ONE GPU builds (identical) and with the single-GPU predict (within 1e-12)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lamp_b200 as et
from lamp_b200 import dist as D

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")  # the process group only hands the NCCL unique id around
ctx = et.Context(local)
D.init_comm(ctx)
z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "mnist_test_u8.npz"))
x, y = z["pixels"].astype(np.float64)[:6000], z["label"].astype(np.int32)[:6000]
m = 13
# the table travels once over PCIe (rank 0) and then over NVLink
dd0 = et.DeviceData.from_rowmajor(x, ctx) if rank == 0 else None
dd = D.broadcast_data(ctx, dd0, 0)
dd.set_target_classification(y, 10)
local_f, ids = D.buildForestClassificationSharded(dd, None, None, 10, 2, 28, m, 8, seed=77, ctx=ctx)
full = D.gather_forest(ctx, local_f)
ref = et.buildForestClassification(x, y, None, 10, 2, 28, m, 8, seed=77, ctx=ctx)
assert len(full) == m
for t in range(m):
    a, b = full.flat(t), ref.flat(t)
    assert np.array_equal(a.feature, b.feature) and np.array_equal(a.cut.view(np.int64), b.cut.view(np.int64)), t
    assert np.array_equal(a.leaf, b.leaf), t
p_ref = et.predictClassification(ref, x, ctx=ctx)
assert np.array_equal(et.predictClassification(full, x, ctx=ctx), p_ref)  # gathered forest: bit-exact (tree order)
p_sh = D.predictClassificationSharded(ctx, local_f, x, m)
np.testing.assert_allclose(p_sh, p_ref, rtol=1e-12, atol=1e-300)
print("rank %d/%d: shard %s, gathered forest identical to the single-GPU forest (%d nodes), sharded predict within "
      "1e-12 (max abs diff %.2e), gather %.3f ms" % (rank, world, ids.tolist(), full.total_nodes,
                                                      float(np.abs(p_sh - p_ref).max()), full.stats["gather_ms"]))
dist.barrier()
if rank == 0:
    # the same through ONE process driving all GPUs (et_init_multi)
    pass
dist.destroy_process_group()
print("dist_check ok")
