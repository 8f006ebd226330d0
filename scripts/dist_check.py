"""Multi-GPU check on real devices: torchrun --nproc-per-node G scripts/dist_check.py
Every rank builds its shard of trees (tree t -> rank t mod G) from its own replica of the table, the serialized
trees are all-gathered over NCCL/gloo into tree order and compared with the forest one GPU builds alone (must be
identical: a tree depends only on (seed, global tree id)); sharded predict all-reduces the per-rank vote sums on
the device (NCCL) and must match the single-GPU prediction within 1e-12 relative."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lamp_b200 as et
from lamp_b200 import dist as D


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = et.Context(local)
    z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "mnist_test_u8.npz"))
    x, y = z["pixels"].astype(np.float64), z["label"].astype(np.int32)
    m = 37
    # ---- classification
    local_f, ids = D.buildForestClassificationSharded(x, y, None, 10, 2, 28, m, 8, seed=11, ctx=ctx)
    assert ids.tolist() == list(range(rank, m, world))
    merged = D.gather_forest(local_f.export_all(), ids)
    full = et.buildForestClassification(x, y, None, 10, 2, 28, m, 8, seed=11, ctx=ctx).export_all()
    for k in ("tree_sizes", "feature", "mil", "left", "right"):
        assert np.array_equal(merged[k], full[k]), k
    assert np.array_equal(merged["cut"].view(np.int64), full["cut"].view(np.int64))
    assert np.array_equal(merged["leaf"], full["leaf"])
    pred = D.predictClassificationSharded(local_f, x[:4000], m, device=torch.device("cuda", local))
    ref = et.predictClassification(et.Forest.import_arrays(full, ctx=ctx), x[:4000])
    np.testing.assert_allclose(pred, ref, rtol=1e-12, atol=1e-15)
    # ---- regression
    yr = y.astype(np.float64) + 0.01 * x[:, 300]
    local_r, ids = D.buildForestRegressionSharded(x[:3000], yr[:3000], 2, 28, 9, 8, seed=3, ctx=ctx)
    predr = D.predictRegressionSharded(local_r, x[:3000], 9, device=torch.device("cuda", local))
    fullr = et.buildForestRegression(x[:3000], yr[:3000], 2, 28, 9, 8, seed=3, ctx=ctx)
    np.testing.assert_allclose(predr, et.predictRegression(fullr, x[:3000]), rtol=1e-12, atol=0)
    dist.barrier()
    if rank == 0:
        print("dist_check ok: world=%d, %d trees gathered in order, sharded predict within 1e-12" % (world, m))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
