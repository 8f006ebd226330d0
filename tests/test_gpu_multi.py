"""Multi-GPU front context (et_init_multi, dist.cu) through the public facade.  The GPU test box has ONE device: a
group of one runs every code path (replication, tree sharding, all-gather of the serialized trees, all-reduced
predict); with two or more devices the same checks run on two (scripts/dist_check.py covers one process per GPU)."""
import ctypes as C

import numpy as np
import pytest

import lamp_b200 as et
from lamp_b200 import _capi as capi
from tests.helpers import synth_regression

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module", params=[1, 2])
def group(request):
    if request.param > _n_gpus():
        pytest.skip("needs %d GPUs" % request.param)
    ctx = et.Context.multi(list(range(request.param)))
    yield ctx
    ctx.close()


def test_multi_classification_equals_single_gpu(mnist, group):
    x, y = mnist
    x, y = x[:4000], y[:4000]
    single = et.buildForestClassification(x, y, None, 10, 2, 28, 7, 4, seed=31)
    multi = et.buildForestClassification(x, y, None, 10, 2, 28, 7, 4, seed=31, ctx=group)
    assert len(multi) == 7 and multi.total_nodes == single.total_nodes
    for t in range(7):  # a tree's stream depends only on (seed, tree id): the gathered forest is the same forest
        a, b = single.flat(t), multi.flat(t)
        assert np.array_equal(a.feature, b.feature) and np.array_equal(a.cut.view(np.int64), b.cut.view(np.int64))
        assert np.array_equal(a.left, b.left) and np.array_equal(a.right, b.right) and np.array_equal(a.leaf, b.leaf)
    ps, pm = et.predictClassification(single, x[:1000]), et.predictClassification(multi, x[:1000], ctx=group)
    np.testing.assert_allclose(pm, ps, rtol=1e-12, atol=1e-300)  # the all-reduce re-associates the sum over trees
    # the packed serialization of the gathered forest round-trips through a single-GPU context
    g = et.Forest.import_packed(multi.export_packed())
    assert np.array_equal(et.predictClassification(g, x[:1000]), ps)
    for key in ("nodes", "v_mm", "s_rows", "p_rows"):
        assert multi.stats[key] == single.stats[key], key


def test_multi_regression_and_resident_data(group):
    x, y = synth_regression(5000, 12, 9)
    dd = et.DeviceData.from_rowmajor(x, group)
    dd.set_target_regression(y)
    multi = et.buildForestRegression(dd, None, 5, 4, 5, 4, seed=3, ctx=group)
    single = et.buildForestRegression(x, y, 5, 4, 5, 4, seed=3)
    for t in range(5):
        a, b = single.flat(t), multi.flat(t)
        assert np.array_equal(a.feature, b.feature) and np.array_equal(a.cut.view(np.int64), b.cut.view(np.int64))
        assert np.array_equal(a.leaf, b.leaf)
    np.testing.assert_allclose(et.predictRegression(multi, x, ctx=group), et.predictRegression(single, x), rtol=1e-12)
    with pytest.raises(capi.EtError):  # the replay hook is single-GPU only
        et.buildForestRegression(dd, None, 5, 4, 1, 1, seed=3, ctx=group,
                                 replay=[dict(left=np.array([-1], np.int32), right=np.array([-1], np.int32),
                                              cand_begin=np.zeros(1, np.int64), cand_count=np.zeros(1, np.int32),
                                              cand_feature=np.zeros(0, np.int32), cand_u=np.zeros(0),
                                              cand_flag=np.zeros(0, np.uint8))])


def test_multi_more_gpus_than_trees_and_imported_forest(mnist, group):
    x, y = mnist
    x, y = x[:1000], y[:1000]
    one = et.buildForestClassification(x, y, None, 10, 2, 16, 1, 4, seed=5, ctx=group)  # empty shards take part
    ref = et.buildForestClassification(x, y, None, 10, 2, 16, 1, 4, seed=5)
    assert np.array_equal(one.flat(0).feature, ref.flat(0).feature)
    imp = et.Forest.from_trees([ref.flat(0)], ctx=group)  # host-held trees live on the first GPU
    assert np.array_equal(et.predictClassification(imp, x, ctx=group), et.predictClassification(ref, x))


def test_comm_rank_collectives_on_a_group_of_one(mnist):
    """The one-process-per-GPU entry points (et_comm_init_rank, et_forest_allgather, et_predict_*_allreduce) with a
    world of one rank: the gathered forest equals the shard, the all-reduced predict equals the plain one."""
    import torch
    from lamp_b200 import dist as D
    x, y = mnist
    x, y = x[:2000], y[:2000]
    ctx = et.Context(0)
    ctx.comm_init_rank(1, 0, D.make_unique_id())
    shard = et.buildForestClassification(x, y, None, 10, 2, 28, 4, 4, seed=8, ctx=ctx, tree_ids=[6, 2, 4, 0])
    full = D.gather_forest(ctx, shard)
    order = [3, 1, 2, 0]  # trees come back sorted by their global ids 0, 2, 4, 6
    for j, t in enumerate(order):
        assert np.array_equal(full.flat(j).feature, shard.flat(t).feature)
        assert np.array_equal(full.flat(j).leaf, shard.flat(t).leaf)
    xt = torch.from_numpy(x).cuda()
    out = torch.empty((len(x), 10), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    D.predict_sharded_device(ctx, shard, xt.data_ptr(), len(x), x.shape[1], out.data_ptr(), 4)
    assert np.array_equal(out.cpu().numpy(), et.predictClassification(shard, x, ctx=ctx))
    assert ctx.comm_last_ms() >= 0.0
    ctx.close()


def test_multi_sparse_resident_table(group, monkeypatch):
    """A CSC table that stays sparse in HBM replicates over the group (the three CSC arrays travel, the row-major index
    is rebuilt per GPU) and builds the forest a single GPU builds."""
    monkeypatch.setenv("ETGPU_CSC_DENSE_MAX", "0")
    rng = np.random.default_rng(12)
    n, d = 6000, 120
    dense = np.where(rng.random((n, d)) < 0.03, np.abs(rng.normal(size=(n, d))) + 0.1, 0.0)
    y = (dense[:, :20].sum(axis=1) > np.median(dense[:, :20].sum(axis=1))).astype(np.int32)
    colptr = np.concatenate([[0], np.cumsum((dense != 0).sum(axis=0))]).astype(np.int64)
    rowidx = np.concatenate([np.flatnonzero(dense[:, c]) for c in range(d)]).astype(np.int32)
    vals = np.concatenate([dense[dense[:, c] != 0, c] for c in range(d)])
    one = et.DeviceData.from_csc(colptr, rowidx, vals, n, d)
    one.set_target_classification(y, 2)
    many = et.DeviceData.from_csc(colptr, rowidx, vals, n, d, group)
    many.set_target_classification(y, 2)
    f1 = et.buildForestClassification(one, None, None, 2, 2, 10, 5, 4, seed=2)
    f2 = et.buildForestClassification(many, None, None, 2, 2, 10, 5, 4, seed=2, ctx=group)
    for t in range(5):
        a, b = f1.flat(t), f2.flat(t)
        assert np.array_equal(a.feature, b.feature) and np.array_equal(a.cut.view(np.int64), b.cut.view(np.int64))
        assert np.array_equal(a.leaf, b.leaf)
    one.free()
    many.free()
