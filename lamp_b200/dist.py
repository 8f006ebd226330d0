"""Tree sharding across the GPUs of one box with ONE PROCESS PER GPU (torchrun): the launcher-side plumbing over the
library's own collectives.  Replaces the reference's JVM thread-pool fan-out (`parTraverseN(parallelism)`,
pkg:653-675).  (One process driving several GPUs needs none of this: `Context.multi(devices)` / et_init_multi.)

Trees are independent (per-tree random stream, pkg:654-655), so the build has NO data-path collective: rank r
builds trees r, r+G, r+2G, ... (round-robin balances the depth lottery) from its own replica of the table.
NCCL runs INSIDE libetgpu.so (dist.cu), on device buffers, only where the path has a real exchange:
  * gather_forest  -> et_forest_allgather: ncclAllGather of the sizes, then of the packed 16-byte
    nodes and leaf tables (slots padded to the largest shard), trees put back in tree-id order by a kernel; every rank holds the whole forest;
  * predict_*Sharded -> et_predict_*_allreduce: every rank traverses all rows for ITS trees, the per-row partial sums
    are ncclAllReduce'd in row chunks overlapped with the traversal, one division by the total tree count.  The
    reference sums leaf values in tree order (pkg:549,584); the sharded sum re-associates, so outputs agree to
    ~1e-15 relative (inside the 1e-12 bar), not bit for bit -- predict on the gathered forest when bit-exactness
    matters.
`torch.distributed` only carries the 128-byte NCCL unique id from rank 0 to the other ranks (any launcher channel
would do); that hand-off and the shard arithmetic are what the CPU (gloo) tests cover.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import numpy as np


def shard_tree_ids(m: int, rank: int, world: int) -> np.ndarray:
    """Global ids of the trees rank `rank` builds: r, r+G, r+2G, ..."""
    if m < 0 or world <= 0 or not 0 <= rank < world:
        raise ValueError("bad shard arguments")
    return np.arange(rank, m, world, dtype=np.int32)


def _dist():
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
    return dist


def make_unique_id() -> bytes:
    """A fresh NCCL unique id (128 bytes) from the library."""
    from . import _capi as capi
    buf = (C.c_uint8 * capi.COMM_ID_BYTES)()
    capi.check(capi.lib().et_comm_unique_id(buf))
    return bytes(buf)


def exchange_unique_id(make_id: Callable[[], bytes] = make_unique_id) -> bytes:
    """Rank 0 creates the id, every rank returns the same 128 bytes (broadcast over the launcher's process group,
    whatever its backend)."""
    dist = _dist()
    box = [make_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise RuntimeError("bad NCCL unique id received")
    return bytes(uid)


def init_comm(ctx, rank: Optional[int] = None, world: Optional[int] = None) -> None:
    """Attaches an NCCL communicator over all ranks of the process group to this rank's context."""
    dist = _dist()
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    ctx.comm_init_rank(world, rank, exchange_unique_id())


def build_forest_sharded(build_fn: Callable[[np.ndarray], object], m: int, rank: Optional[int] = None,
                         world: Optional[int] = None):
    """Runs `build_fn(tree_ids)` for this rank's shard; returns (local_forest, tree_ids)."""
    if rank is None or world is None:
        dist = _dist()
        rank, world = dist.get_rank(), dist.get_world_size()
    ids = shard_tree_ids(m, rank, world)
    return build_fn(ids), ids


def gather_forest(ctx, local_forest):
    """All-gather of the serialized trees (in the library, device to device): the whole forest, in tree-id order,
    resident on this rank's GPU."""
    from . import _capi as capi
    from .extratrees import Forest
    h = C.c_void_p()
    capi.check(capi.lib().et_forest_allgather(ctx.h, local_forest.h, C.byref(h)))
    return Forest(ctx, h, dict(local_forest.stats, gather_ms=ctx.comm_last_ms()))


def broadcast_data(ctx, data, root: int = 0):
    """Replicates a resident table from `root` to every rank over NVLink; `data` is None on the other ranks."""
    from . import _capi as capi
    from .extratrees import DeviceData
    h = C.c_void_p()
    capi.check(capi.lib().et_data_broadcast(ctx.h, None if data is None else data.h, root, C.byref(h)))
    return data if data is not None else DeviceData(ctx, h)


def predict_sharded_device(ctx, local_forest, x_dev_ptr: int, n: int, d: int, out_dev_ptr: int, m_total: int) -> None:
    """Tree-sharded predict on device buffers (x row-major [n, d], out [n, leaf_width]); collective."""
    from . import _capi as capi
    fn = capi.lib().et_predict_regression_allreduce if local_forest.regression else \
        capi.lib().et_predict_classification_allreduce
    capi.check(fn(ctx.h, local_forest.h, C.c_void_p(x_dev_ptr), n, d, C.c_void_p(out_dev_ptr), m_total))


# ---- convenience wrappers over the GPU facade (one process per GPU) ------------------------------


def buildForestClassificationSharded(data, target, sampleWeights, numClasses, nMin, k, m, parallelism,
                                     bestSplit=False, maxDepth=2**31 - 1, seed=0, ctx=None):
    from . import extratrees as et
    return build_forest_sharded(
        lambda ids: et.buildForestClassification(data, target, sampleWeights, numClasses, nMin, k, len(ids),
                                                 parallelism, bestSplit, maxDepth, seed, ctx=ctx, tree_ids=ids), m)


def buildForestRegressionSharded(data, target, nMin, k, m, parallelism, bestSplit=False, maxDepth=2**31 - 1,
                                 seed=0, ctx=None):
    from . import extratrees as et
    return build_forest_sharded(
        lambda ids: et.buildForestRegression(data, target, nMin, k, len(ids), parallelism, bestSplit, maxDepth,
                                             seed, ctx=ctx, tree_ids=ids), m)


def _predict_sharded_host(ctx, local_forest, samples, m_total: int):
    import torch
    x = torch.from_numpy(np.ascontiguousarray(samples, dtype=np.float64)).cuda(ctx.device)
    out = torch.empty((x.shape[0], local_forest.leaf_width), dtype=torch.float64, device=x.device)
    torch.cuda.synchronize(x.device)
    predict_sharded_device(ctx, local_forest, x.data_ptr(), x.shape[0], x.shape[1], out.data_ptr(), m_total)
    return out.cpu().numpy()


def predictClassificationSharded(ctx, local_forest, samples, m_total: int):
    return _predict_sharded_host(ctx, local_forest, samples, m_total)


def predictRegressionSharded(ctx, local_forest, samples, m_total: int):
    return _predict_sharded_host(ctx, local_forest, samples, m_total)[:, 0]
