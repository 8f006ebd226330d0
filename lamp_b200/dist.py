"""Tree sharding across the GPUs of one box: one process per GPU (torchrun), `torch.distributed` for the
plumbing.  Replaces the reference's JVM thread-pool fan-out (`parTraverseN(parallelism)`, pkg:653-675).

Trees are independent (per-tree random stream, pkg:654-655), so the build has NO data-path collective:
rank r builds trees r, r+G, r+2G, ... (round-robin balances the depth lottery) from its own replica of
the table.  Collectives are used only where the path has a real exchange:
  * gather_forest: all-gather of the serialized trees so that every rank (and the host) holds the whole
    forest in tree order;
  * predict_*: trees stay sharded, every rank traverses all rows for ITS trees and the per-row partial sums
    are all-reduced (sum), then divided by the total tree count.  The reference sums leaf values in tree
    order (pkg:549,584); the sharded sum re-associates, so outputs agree to ~1e-15 relative (inside the
    1e-12 bar), not bit for bit -- use gather_forest + single-GPU predict when bit-exactness matters.

A tree's stream depends only on (seed, global tree id): a forest sharded over G GPUs equals the forest
built on one.  The same code runs on CPU with the gloo backend for the host-side tests (build/predict
callables are injected there; there is no CPU compute path in this package).
"""
from __future__ import annotations

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


_SER_KEYS = ("tree_sizes", "feature", "cut", "mil", "left", "right", "leaf")


def merge_serialized(parts: list[dict], ids: list[np.ndarray]) -> dict:
    """Interleaves per-rank serialized forests (Forest.export_all dicts) back into global tree order."""
    lw = parts[0]["leaf_width"]
    m = int(sum(len(i) for i in ids))
    owner = np.empty(m, np.int64)
    local = np.empty(m, np.int64)
    for r, tid in enumerate(ids):
        owner[tid] = r
        local[tid] = np.arange(len(tid))
    offs = [np.concatenate([[0], np.cumsum(p["tree_sizes"], dtype=np.int64)]) for p in parts]
    out = {k: [] for k in _SER_KEYS}
    for t in range(m):
        r, j = owner[t], local[t]
        a, b = offs[r][j], offs[r][j + 1]
        out["tree_sizes"].append(parts[r]["tree_sizes"][j:j + 1])
        for k in ("feature", "cut", "mil", "left", "right", "leaf"):
            out[k].append(parts[r][k][a:b])
    res = {k: np.concatenate(v) if v else np.zeros(0) for k, v in out.items()}
    res["leaf"] = res["leaf"].reshape(-1, lw)
    res["leaf_width"] = lw
    res["regression"] = parts[0]["regression"]
    return res


def gather_forest(local_serialized: dict, local_ids: np.ndarray) -> dict:
    """All-gather of the serialized trees (variable size per rank) -> whole forest in tree order."""
    dist = _dist()
    world = dist.get_world_size()
    parts: list = [None] * world
    dist.all_gather_object(parts, (local_serialized, np.asarray(local_ids)))
    return merge_serialized([p[0] for p in parts], [p[1] for p in parts])


def build_forest_sharded(build_fn: Callable[[np.ndarray], object], m: int, rank: Optional[int] = None,
                         world: Optional[int] = None):
    """Runs `build_fn(tree_ids)` for this rank's shard; returns (local_forest, tree_ids)."""
    if rank is None or world is None:
        dist = _dist()
        rank, world = dist.get_rank(), dist.get_world_size()
    ids = shard_tree_ids(m, rank, world)
    return build_fn(ids), ids


def predict_sharded(partial_sum: np.ndarray, m_total: int, device=None) -> np.ndarray:
    """All-reduce (sum) of the per-rank partial vote / mean sums, then the division by the tree count."""
    import torch
    dist = _dist()
    t = torch.from_numpy(np.ascontiguousarray(partial_sum, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return (t.cpu().numpy() if device is not None else t.numpy()) / float(m_total)


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


def predictClassificationSharded(local_forest, samples, m_total: int, device=None):
    from . import extratrees as et
    part = et.predictClassification(local_forest, samples, sum_only=True) if len(local_forest) else \
        np.zeros((np.asarray(samples).shape[0], local_forest.leaf_width))
    return predict_sharded(part, m_total, device)


def predictRegressionSharded(local_forest, samples, m_total: int, device=None):
    from . import extratrees as et
    part = et.predictRegression(local_forest, samples, sum_only=True) if len(local_forest) else \
        np.zeros(np.asarray(samples).shape[0])
    return predict_sharded(part, m_total, device)
