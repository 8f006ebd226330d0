"""Host-side logic of the one-process-per-GPU path on CPU: a world_size-2 gloo run of the tree sharding and of the
NCCL unique-id hand-off (lamp_b200/dist.py).  The collectives themselves run inside libetgpu.so on device buffers
(dist.cu) and are covered by the GPU tests (single-device group) and scripts/dist_check.py (2+ GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from lamp_b200 import dist as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = 7
        # every rank builds its own shard: the shards partition 0..m-1 and keep tree-id order inside a rank
        local, ids = D.build_forest_sharded(lambda tid: [int(t) * 10 for t in tid], m)
        assert ids.tolist() == list(range(rank, m, world)) and local == [t * 10 for t in ids]
        gathered = [None] * world
        dist.all_gather_object(gathered, ids.tolist())
        assert sorted(sum(gathered, [])) == list(range(m))
        # the NCCL unique id is made on rank 0 only and arrives unchanged everywhere
        made = []

        def make_id():
            made.append(1)
            return bytes((7 * i + 3) % 256 for i in range(128))

        uid = D.exchange_unique_id(make_id)
        assert uid == bytes((7 * i + 3) % 256 for i in range(128))
        assert len(made) == (1 if rank == 0 else 0)
        # a malformed id is refused on every rank
        try:
            D.exchange_unique_id(lambda: b"short")
            raise AssertionError("short id accepted")
        except RuntimeError:
            pass
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
