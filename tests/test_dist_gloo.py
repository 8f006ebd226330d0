"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo run of tree sharding, the all-gather
of serialized trees and the vote all-reduce (lamp_b200/dist.py).  Compute is injected from the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lamp_b200 import dist as D
from oracle import oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _serialize(trees, lw, regression):
    return dict(tree_sizes=np.array([t.n_nodes for t in trees], np.int32),
                feature=np.concatenate([t.feature for t in trees]), cut=np.concatenate([t.cut for t in trees]),
                mil=np.concatenate([t.mil for t in trees]), left=np.concatenate([t.left for t in trees]),
                right=np.concatenate([t.right for t in trees]),
                leaf=np.concatenate([t.leaf for t in trees]).reshape(-1, lw), leaf_width=lw, regression=regression)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        x = rng.normal(size=(400, 6))
        y = (x[:, 0] + x[:, 1] > 0).astype(np.int32)
        m = 7
        full = O.build_forest_classification(x, y, None, 2, 2, 3, m, 4, seed=5)  # identical on every rank
        trees = full.trees()

        def build(ids):  # stands in for the GPU build of this rank's shard
            return [trees[t] for t in ids]

        local, ids = D.build_forest_sharded(build, m)
        assert ids.tolist() == list(range(rank, m, world))
        merged = D.gather_forest(_serialize(local, 2, False), ids)
        ref = _serialize(trees, 2, False)
        for k in ("tree_sizes", "feature", "mil", "left", "right"):
            assert np.array_equal(merged[k], ref[k]), k
        assert np.array_equal(merged["cut"].view(np.int64), ref["cut"].view(np.int64))
        assert np.array_equal(merged["leaf"], ref["leaf"])
        # sharded predict: per-rank partial sums -> all-reduce -> / m
        part = np.zeros((len(x), 2))
        for t in local:
            part += O.import_forest([t], False).predict(x)
        pred = D.predict_sharded(part, m)
        np.testing.assert_allclose(pred, full.predict(x), rtol=1e-12, atol=0)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_shard_ids():
    assert D.shard_tree_ids(10, 1, 4).tolist() == [1, 5, 9]
    assert D.shard_tree_ids(2, 3, 4).tolist() == []
    ids = np.concatenate([D.shard_tree_ids(11, r, 3) for r in range(3)])
    assert sorted(ids.tolist()) == list(range(11))
    with pytest.raises(ValueError):
        D.shard_tree_ids(4, 4, 4)


def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
