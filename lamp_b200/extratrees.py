"""Host-side mirror of the reference's public extratrees API (package lamp.extratrees,
extratrees/src/main/scala/lamp/forest/package.scala, cited pkg:LINE) over the C ABI of
include/etgpu.h.  Same function names, argument meaning and error behaviour:

    buildForestClassification   pkg:611-681        predictClassification   pkg:542-551
    buildForestRegression       pkg:704-764        predictRegression       pkg:577-586
    ClassificationLeaf / ClassificationNonLeaf / RegressionLeaf / RegressionNonLeaf   extratrees.scala:3-63

numpy arrays stand in for saddle's Mat[Double] (row-major n x d) / Vec[Int] / Vec[Double];
`require` failures raise ValueError (IllegalArgumentException in the reference).  All compute runs
in hand-written sm_100a CUDA kernels; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import time
from collections.abc import Sequence
from dataclasses import dataclass
from typing import Optional, Union

import numpy as np

from . import _capi as capi
from ._capi import check, ptr, dp, ip, lp, bp

INT_MAX = 2**31 - 1

# ---- tree ADTs (extratrees.scala:3-63) ----------------------------------------------------------


@dataclass(frozen=True)
class ClassificationLeaf:
    targetDistribution: tuple


@dataclass(frozen=True)
class ClassificationNonLeaf:
    left: "ClassificationTree"
    right: "ClassificationTree"
    splitFeature: int
    cutpoint: float
    splitMissingIsLess: bool


@dataclass(frozen=True)
class RegressionLeaf:
    targetMean: float


@dataclass(frozen=True)
class RegressionNonLeaf:
    left: "RegressionTree"
    right: "RegressionTree"
    splitFeature: int
    cutpoint: float
    splitMissingIsLess: bool


ClassificationTree = Union[ClassificationLeaf, ClassificationNonLeaf]
RegressionTree = Union[RegressionLeaf, RegressionNonLeaf]


@dataclass
class FlatTree:
    """One tree in pre-order -- the wire format of et_forest_export."""
    feature: np.ndarray  # int32, -1 = leaf
    cut: np.ndarray      # float64
    mil: np.ndarray      # uint8  splitMissingIsLess
    left: np.ndarray     # int32 (-1 for leaves)
    right: np.ndarray    # int32
    leaf: np.ndarray     # float64 [n_nodes, leaf_width]

    @property
    def n_nodes(self):
        return len(self.feature)


def flat_to_adt(t: FlatTree, regression: bool):
    """Pre-order arrays -> nested case-class objects (children are built before parents)."""
    n = t.n_nodes
    built = [None] * n
    for i in range(n - 1, -1, -1):  # pre-order: children have larger ids than their parent
        if t.feature[i] < 0:
            built[i] = RegressionLeaf(float(t.leaf[i, 0])) if regression else \
                ClassificationLeaf(tuple(float(v) for v in t.leaf[i]))
        else:
            cls = RegressionNonLeaf if regression else ClassificationNonLeaf
            built[i] = cls(built[t.left[i]], built[t.right[i]], int(t.feature[i]), float(t.cut[i]), bool(t.mil[i]))
    return built[0]


def adt_to_flat(root, leaf_width: int) -> FlatTree:
    feature, cut, mil, left, right, leaf = [], [], [], [], [], []
    stack = [(root, -1, False)]
    while stack:
        node, parent, is_right = stack.pop()
        me = len(feature)
        if parent >= 0:
            (right if is_right else left)[parent] = me
        if isinstance(node, (ClassificationLeaf, RegressionLeaf)):
            feature.append(-1)
            cut.append(float("nan"))
            mil.append(0)
            left.append(-1)
            right.append(-1)
            leaf.append([node.targetMean] if isinstance(node, RegressionLeaf) else list(node.targetDistribution))
        else:
            feature.append(node.splitFeature)
            cut.append(node.cutpoint)
            mil.append(1 if node.splitMissingIsLess else 0)
            left.append(-1)
            right.append(-1)
            leaf.append([0.0] * leaf_width)
            stack.append((node.right, me, True))
            stack.append((node.left, me, False))
    return FlatTree(np.array(feature, np.int32), np.array(cut, np.float64), np.array(mil, np.uint8),
                    np.array(left, np.int32), np.array(right, np.int32),
                    np.array(leaf, np.float64).reshape(len(feature), leaf_width))


# ---- context / resident data --------------------------------------------------------------------


class Context:
    """One GPU, or -- Context.multi(devices) -- several GPUs of one box behind one handle (et_init_multi: trees
    sharded by tree id, NCCL inside the library).  The reference hides its thread pool behind the call
    (pkg:653-675); so does this (default context)."""

    def __init__(self, device: int = 0, _handle=None, _devices=None):
        if _handle is None:
            h = C.c_void_p()
            check(capi.lib().et_init(device, C.byref(h)))
            _handle, _devices = h, [device]
        self.h = _handle
        self.device = _devices[0]
        self.devices = list(_devices)

    @classmethod
    def multi(cls, devices) -> "Context":
        devs = np.ascontiguousarray(list(devices), dtype=np.int32)
        h = C.c_void_p()
        check(capi.lib().et_init_multi(ptr(devs, ip), len(devs), C.byref(h)))
        return cls(_handle=h, _devices=devs.tolist())

    @property
    def n_devices(self) -> int:
        return len(self.devices)

    def comm_init_rank(self, world: int, rank: int, unique_id: bytes):
        """One process per GPU: attach an NCCL communicator (the launcher distributes rank 0's unique id)."""
        buf = (C.c_uint8 * capi.COMM_ID_BYTES).from_buffer_copy(unique_id)
        check(capi.lib().et_comm_init_rank(self.h, world, rank, buf))

    def comm_last_ms(self) -> float:
        return float(capi.lib().et_comm_last_ms(self.h))

    def set_stream(self, cuda_stream: Optional[int]):
        check(capi.lib().et_set_stream(self.h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def synchronize(self):
        check(capi.lib().et_synchronize(self.h))

    def close(self):
        if self.h:
            capi.lib().et_shutdown(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: dict[int, Context] = {}


def default_context(device: Optional[int] = None) -> Context:
    if device is None:
        import os
        device = int(os.environ.get("LOCAL_RANK", "0")) if "ETGPU_DEVICE" not in os.environ else \
            int(os.environ["ETGPU_DEVICE"])
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def _f64_2d(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim != 2:
        raise ValueError("data must be a 2-d row-major matrix (saddle Mat[Double])")
    return a


class DeviceData:
    """A table resident in HBM (column-major FP64), optionally with targets/weights attached."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self.h = handle

    @staticmethod
    def from_rowmajor(data, ctx: Optional[Context] = None) -> "DeviceData":
        ctx = ctx or default_context()
        x = _f64_2d(data)
        h = C.c_void_p()
        check(capi.lib().et_data_dense_rowmajor(ctx.h, ptr(x, dp), x.shape[0], x.shape[1], C.byref(h)))
        return DeviceData(ctx, h)

    @staticmethod
    def from_device_rowmajor(dev_ptr: int, n: int, d: int, ctx: Optional[Context] = None) -> "DeviceData":
        ctx = ctx or default_context()
        h = C.c_void_p()
        check(capi.lib().et_data_dense_rowmajor_device(ctx.h, C.c_void_p(dev_ptr), n, d, C.byref(h)))
        return DeviceData(ctx, h)

    @staticmethod
    def from_column_blocks(n: int, d: int, blocks, ctx: Optional[Context] = None) -> "DeviceData":
        """blocks: iterable of (first_col, array [n_cols, n] -- whole columns, column-major)."""
        ctx = ctx or default_context()
        h = C.c_void_p()
        check(capi.lib().et_data_dense_alloc(ctx.h, n, d, C.byref(h)))
        out = DeviceData(ctx, h)
        for first, blk in blocks:
            blk = np.ascontiguousarray(blk, dtype=np.float64)
            assert blk.ndim == 2 and blk.shape[1] == n
            check(capi.lib().et_data_dense_colblock(ctx.h, h, ptr(blk, dp), first, blk.shape[0]))
        return out

    @staticmethod
    def from_csc(colptr, rowidx, values, n: int, d: int, ctx: Optional[Context] = None) -> "DeviceData":
        """Compressed-sparse-column input (entries not listed are 0.0, dense semantics).  Also accepts the
        attributes of a scipy.sparse.csc_matrix: from_csc(m.indptr, m.indices, m.data, *m.shape)."""
        ctx = ctx or default_context()
        cp = np.ascontiguousarray(colptr, dtype=np.int64)
        ri = np.ascontiguousarray(rowidx, dtype=np.int32)
        va = np.ascontiguousarray(values, dtype=np.float64)
        if len(cp) != d + 1 or len(ri) != len(va) or (len(cp) and cp[-1] != len(va)):
            raise ValueError("from_csc: colptr must have d + 1 entries ending at the number of stored values")
        h = C.c_void_p()
        check(capi.lib().et_data_csc(ctx.h, cp.ctypes.data_as(C.POINTER(C.c_int64)), ptr(ri, ip), ptr(va, dp), n, d,
                                     C.byref(h)))
        return DeviceData(ctx, h)

    @property
    def shape(self):
        n, d = C.c_int64(), C.c_int32()
        check(capi.lib().et_data_dims(self.h, C.byref(n), C.byref(d)))
        return n.value, d.value

    def set_target_classification(self, target, num_classes: int):
        y = np.ascontiguousarray(target, dtype=np.int32)
        check(capi.lib().et_data_set_target_classification(self.ctx.h, self.h, ptr(y, ip), len(y), num_classes))

    def set_target_regression(self, target):
        y = np.ascontiguousarray(target, dtype=np.float64)
        check(capi.lib().et_data_set_target_regression(self.ctx.h, self.h, ptr(y, dp), len(y)))

    def set_weights(self, weights):
        if weights is None:
            check(capi.lib().et_data_set_weights(self.ctx.h, self.h, None, 0))
        else:
            w = np.ascontiguousarray(weights, dtype=np.float64)
            check(capi.lib().et_data_set_weights(self.ctx.h, self.h, ptr(w, dp), len(w)))

    def free(self):
        if self.h:
            capi.lib().et_data_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---- forest --------------------------------------------------------------------------------------


class Forest(Sequence):
    """Seq[ClassificationTree] / Seq[RegressionTree] backed by the library's flat forest.
    Indexing materialises the reference's nested case classes lazily; `flat(t)` gives the arrays."""

    def __init__(self, ctx: Context, handle, stats: Optional[dict] = None):
        self.ctx = ctx
        self.h = handle
        self.stats = stats or {}
        m, lw, reg, tot = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        check(capi.lib().et_forest_dims(handle, C.byref(m), C.byref(lw), C.byref(reg), C.byref(tot)))
        self.m, self.leaf_width, self.regression, self.total_nodes = m.value, lw.value, bool(reg.value), tot.value

    def __len__(self):
        return self.m

    def __getitem__(self, t):
        if isinstance(t, slice):
            return [self[i] for i in range(*t.indices(self.m))]
        if t < 0:
            t += self.m
        if not 0 <= t < self.m:
            raise IndexError(t)
        return flat_to_adt(self.flat(t), self.regression)

    def flat(self, t: int) -> FlatTree:
        n = C.c_int32()
        check(capi.lib().et_forest_tree_size(self.h, t, C.byref(n)))
        n = n.value
        ft = FlatTree(np.empty(n, np.int32), np.empty(n, np.float64), np.empty(n, np.uint8), np.empty(n, np.int32),
                      np.empty(n, np.int32), np.empty((n, self.leaf_width), np.float64))
        check(capi.lib().et_forest_export(self.h, t, ptr(ft.feature, ip), ptr(ft.cut, dp), ptr(ft.mil, bp),
                                          ptr(ft.left, ip), ptr(ft.right, ip), ptr(ft.leaf, dp)))
        return ft

    def export_all(self) -> dict:
        """The serialized forest: arrays concatenated over trees (what is gathered across GPUs)."""
        tot, lw = self.total_nodes, self.leaf_width
        out = dict(tree_sizes=np.empty(self.m, np.int32), feature=np.empty(tot, np.int32),
                   cut=np.empty(tot, np.float64), mil=np.empty(tot, np.uint8), left=np.empty(tot, np.int32),
                   right=np.empty(tot, np.int32), leaf=np.empty((tot, lw), np.float64))
        check(capi.lib().et_forest_export_all(self.h, ptr(out["tree_sizes"], ip), ptr(out["feature"], ip),
                                              ptr(out["cut"], dp), ptr(out["mil"], bp), ptr(out["left"], ip),
                                              ptr(out["right"], ip), ptr(out["leaf"], dp)))
        out["leaf_width"] = lw
        out["regression"] = self.regression
        return out

    PACKED_NODE = np.dtype([("cut", np.float64), ("feat", np.int32), ("right_or_leaf", np.int32)])

    def export_packed(self, nodes_out=None, leaves_out=None) -> dict:
        """The forest in the device layout (et_forest_export_packed): 16-byte pre-order node records, the
        compact leaf table and the tree offsets, copied device->host as they are.  `nodes_out` / `leaves_out`
        may be preallocated (e.g. pinned) buffers at least as large as needed."""
        tn, tl = C.c_int64(), C.c_int64()
        check(capi.lib().et_forest_packed_dims(self.h, C.byref(tn), C.byref(tl)))
        tn, tl = tn.value, tl.value
        nodes = np.empty(tn, self.PACKED_NODE) if nodes_out is None else nodes_out[:tn]
        leaves = np.empty((tl, self.leaf_width), np.float64) if leaves_out is None else \
            leaves_out.reshape(-1)[:tl * self.leaf_width].reshape(tl, self.leaf_width)
        if nodes.dtype != self.PACKED_NODE or len(nodes) != tn or not nodes.flags.c_contiguous:
            raise ValueError("nodes_out must be a contiguous PACKED_NODE array with room for the forest")
        if leaves.dtype != np.float64 or leaves.size != tl * self.leaf_width or not leaves.flags.c_contiguous:
            raise ValueError("leaves_out must be a contiguous float64 array with room for the leaf table")
        off = np.empty(self.m + 1, np.int64)
        check(capi.lib().et_forest_export_packed(self.h, C.c_void_p(nodes.ctypes.data), C.c_void_p(leaves.ctypes.data),
                                                 C.c_void_p(off.ctypes.data)))
        return dict(nodes=nodes, leaves=leaves, tree_off=off, leaf_width=self.leaf_width, regression=self.regression)

    @staticmethod
    def import_packed(ser: dict, ctx: Optional[Context] = None) -> "Forest":
        ctx = ctx or default_context()
        nodes = np.ascontiguousarray(ser["nodes"], dtype=Forest.PACKED_NODE)
        leaves = np.ascontiguousarray(ser["leaves"], dtype=np.float64)
        off = np.ascontiguousarray(ser["tree_off"], dtype=np.int64)
        lw = int(ser["leaf_width"])
        h = C.c_void_p()
        check(capi.lib().et_forest_import_packed(ctx.h, len(off) - 1, lw, int(ser["regression"]), len(nodes),
                                                 leaves.size // lw, C.c_void_p(nodes.ctypes.data),
                                                 C.c_void_p(leaves.ctypes.data), C.c_void_p(off.ctypes.data),
                                                 C.byref(h)))
        return Forest(ctx, h)

    @staticmethod
    def import_arrays(ser: dict, ctx: Optional[Context] = None) -> "Forest":
        ctx = ctx or default_context()
        h = C.c_void_p()
        a = {k: np.ascontiguousarray(ser[k]) for k in ("tree_sizes", "feature", "cut", "mil", "left", "right", "leaf")}
        check(capi.lib().et_forest_import(ctx.h, len(a["tree_sizes"]), int(ser["leaf_width"]), int(ser["regression"]),
                                          ptr(a["tree_sizes"].astype(np.int32), ip), ptr(a["feature"].astype(np.int32), ip),
                                          ptr(a["cut"].astype(np.float64), dp), ptr(a["mil"].astype(np.uint8), bp),
                                          ptr(a["left"].astype(np.int32), ip), ptr(a["right"].astype(np.int32), ip),
                                          ptr(a["leaf"].astype(np.float64), dp), C.byref(h)))
        return Forest(ctx, h)

    @staticmethod
    def from_trees(trees, ctx: Optional[Context] = None, regression_hint: bool = False) -> "Forest":
        """Seq of nested ADT trees or FlatTrees -> device forest (the reference's predict input)."""
        trees = list(trees)
        if not trees:
            raise ValueError("empty forest")
        first = trees[0]
        if hasattr(first, "feature") and hasattr(first, "leaf"):  # FlatTree-like (pre-order arrays)
            flats = trees
            lw = first.leaf.shape[1]
            regression = regression_hint
        else:
            regression = isinstance(first, (RegressionLeaf, RegressionNonLeaf))
            lw = 1 if regression else len(_first_leaf(first).targetDistribution)
            flats = [adt_to_flat(t, lw) for t in trees]
        ser = dict(tree_sizes=np.array([t.n_nodes for t in flats], np.int32),
                   feature=np.concatenate([t.feature for t in flats]), cut=np.concatenate([t.cut for t in flats]),
                   mil=np.concatenate([t.mil for t in flats]), left=np.concatenate([t.left for t in flats]),
                   right=np.concatenate([t.right for t in flats]),
                   leaf=np.concatenate([t.leaf for t in flats]).reshape(-1, lw), leaf_width=lw,
                   regression=bool(regression))
        return Forest.import_arrays(ser, ctx)

    def free(self):
        if self.h:
            capi.lib().et_forest_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _first_leaf(t):
    while not isinstance(t, (ClassificationLeaf, RegressionLeaf)):
        t = t.left
    return t


def make_replay(per_tree):
    """Builds the et_replay test hook from per-tree dicts/objects holding pre-order `left`, `right`
    and the candidate trace (`cand_begin`, `cand_count`, `cand_feature`, `cand_u`, `cand_flag`).
    Returns (EtReplay, keepalive)."""
    node_offset = np.zeros(len(per_tree) + 1, np.int64)
    cand_off = 0
    cb, cc, le, ri, cf, cu, cfl = [], [], [], [], [], [], []
    for t, tr in enumerate(per_tree):
        g = (lambda k: tr[k]) if isinstance(tr, dict) else (lambda k: getattr(tr, k))
        n = len(g("left"))
        node_offset[t + 1] = node_offset[t] + n
        cb.append(np.asarray(g("cand_begin"), np.int64) + cand_off)
        cc.append(np.asarray(g("cand_count"), np.int32))
        le.append(np.asarray(g("left"), np.int32))
        ri.append(np.asarray(g("right"), np.int32))
        cf.append(np.asarray(g("cand_feature"), np.int32))
        cu.append(np.asarray(g("cand_u"), np.float64))
        cfl.append(np.asarray(g("cand_flag"), np.uint8))
        cand_off += len(cf[-1])
    cat = lambda xs, dt: np.ascontiguousarray(np.concatenate(xs) if xs else np.zeros(0, dt), dtype=dt)
    keep = dict(node_offset=node_offset, cand_begin=cat(cb, np.int64), cand_count=cat(cc, np.int32),
                left=cat(le, np.int32), right=cat(ri, np.int32), cand_feature=cat(cf, np.int32),
                cand_u=cat(cu, np.float64), cand_flag=cat(cfl, np.uint8))
    r = capi.EtReplay(len(per_tree), ptr(keep["node_offset"], lp), ptr(keep["cand_begin"], lp),
                      ptr(keep["cand_count"], ip), ptr(keep["left"], ip), ptr(keep["right"], ip), cand_off,
                      ptr(keep["cand_feature"], ip), ptr(keep["cand_u"], dp), ptr(keep["cand_flag"], bp))
    return r, keep


# ---- the four public functions ---------------------------------------------------------------------


def _now_ms() -> int:
    return int(time.time() * 1000)  # java.time.Instant.now.toEpochMilli (pkg:622,713)


def _wrap64(v: int) -> int:
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def _resolve(data, ctx):
    if isinstance(data, DeviceData):
        return data, False
    if hasattr(data, "indptr") and hasattr(data, "indices") and hasattr(data, "tocsc"):  # a scipy.sparse matrix
        m = data.tocsc()
        m.sum_duplicates()
        return DeviceData.from_csc(m.indptr, m.indices, m.data, m.shape[0], m.shape[1], ctx), True
    return DeviceData.from_rowmajor(data, ctx), True


def buildForestClassification(data, target, sampleWeights, numClasses: int, nMin: int, k: int, m: int,
                              parallelism: int, bestSplit: bool = False, maxDepth: int = INT_MAX,
                              seed: Optional[int] = None, *, ctx: Optional[Context] = None, replay=None,
                              tree_ids=None, allow_replay_mismatch: bool = False) -> Forest:
    """pkg:611-681.  `data`: row-major [n, d] float64 array (or a resident DeviceData, in which case
    target/sampleWeights may be None to use the attached ones).  Returns the trees in index order."""
    ctx = ctx or (data.ctx if isinstance(data, DeviceData) else default_context())
    dd, own = _resolve(data, ctx)
    try:
        y = None if target is None else np.ascontiguousarray(target, dtype=np.int32)
        w = None if sampleWeights is None else np.ascontiguousarray(sampleWeights, dtype=np.float64)
        if y is not None and w is not None and len(w) != len(y):
            raise ValueError("sampleWeights.length != target.length")
        ids = None if tree_ids is None else np.ascontiguousarray(tree_ids, dtype=np.int32)
        if ids is not None and len(ids) != m:
            raise ValueError("tree_ids must have m entries")
        rp, keep = (None, None) if replay is None else (replay if isinstance(replay, tuple) else make_replay(replay))
        h, st = C.c_void_p(), capi.EtStats()
        rc = capi.lib().et_build_classification(
            ctx.h, dd.h, ptr(y, ip), 0 if y is None else len(y), ptr(w, dp), numClasses, nMin, k, m, parallelism,
            int(bestSplit), maxDepth, C.c_int64(_wrap64(_now_ms() if seed is None else seed)), ptr(ids, ip),
            None if rp is None else C.byref(rp), C.byref(h), C.byref(st))
        check(rc)
        f = Forest(ctx, h, st.as_dict())
        if f.stats["replay_mismatches"] and not allow_replay_mismatch:
            raise capi.EtError(capi.ET_EREPLAY, f"{f.stats['replay_mismatches']} decisions contradict the trace")
        return f
    finally:
        if own:
            dd.free()


def buildForestRegression(data, target, nMin: int, k: int, m: int, parallelism: int, bestSplit: bool = False,
                          maxDepth: int = INT_MAX, seed: Optional[int] = None, *, ctx: Optional[Context] = None,
                          replay=None, tree_ids=None, allow_replay_mismatch: bool = False) -> Forest:
    """pkg:704-764."""
    ctx = ctx or (data.ctx if isinstance(data, DeviceData) else default_context())
    dd, own = _resolve(data, ctx)
    try:
        y = None if target is None else np.ascontiguousarray(target, dtype=np.float64)
        ids = None if tree_ids is None else np.ascontiguousarray(tree_ids, dtype=np.int32)
        if ids is not None and len(ids) != m:
            raise ValueError("tree_ids must have m entries")
        rp, keep = (None, None) if replay is None else (replay if isinstance(replay, tuple) else make_replay(replay))
        h, st = C.c_void_p(), capi.EtStats()
        rc = capi.lib().et_build_regression(
            ctx.h, dd.h, ptr(y, dp), 0 if y is None else len(y), nMin, k, m, parallelism, int(bestSplit), maxDepth,
            C.c_int64(_wrap64(_now_ms() if seed is None else seed)), ptr(ids, ip),
            None if rp is None else C.byref(rp), C.byref(h), C.byref(st))
        check(rc)
        f = Forest(ctx, h, st.as_dict())
        if f.stats["replay_mismatches"] and not allow_replay_mismatch:
            raise capi.EtError(capi.ET_EREPLAY, f"{f.stats['replay_mismatches']} decisions contradict the trace")
        return f
    finally:
        if own:
            dd.free()


def _as_forest(trees, ctx) -> tuple[Forest, bool]:
    if isinstance(trees, Forest):
        return trees, False
    return Forest.from_trees(trees, ctx), True


def predictClassification(trees, samples, *, ctx: Optional[Context] = None, sum_only: bool = False) -> np.ndarray:
    """pkg:542-551.  Returns n x numClasses (column c = class c), mean over trees of the leaf
    distributions."""
    f, own = _as_forest(trees, ctx)
    try:
        if f.regression:
            raise ValueError("predictClassification needs classification trees")
        x = _f64_2d(samples)
        out = np.empty((x.shape[0], f.leaf_width), np.float64)
        check(capi.lib().et_predict_classification(f.ctx.h, f.h, ptr(x, dp), x.shape[0], x.shape[1], ptr(out, dp),
                                                   int(sum_only)))
        return out
    finally:
        if own:
            f.free()


def predictRegression(trees, samples, *, ctx: Optional[Context] = None, sum_only: bool = False) -> np.ndarray:
    """pkg:577-586.  Returns a length-n vector, mean over trees of the leaf means."""
    f, own = _as_forest(trees, ctx)
    try:
        if not f.regression:
            raise ValueError("predictRegression needs regression trees")
        x = _f64_2d(samples)
        out = np.empty(x.shape[0], np.float64)
        check(capi.lib().et_predict_regression(f.ctx.h, f.h, ptr(x, dp), x.shape[0], x.shape[1], ptr(out, dp),
                                               int(sum_only)))
        return out
    finally:
        if own:
            f.free()
