"""ctypes wrapper around the CPU oracle (oracle/et_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (lamp_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libetoracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "et_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return _LIB_PATH


_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_bp = C.POINTER(C.c_uint8)


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(t)


class Cmwc5State(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in "xyzwv"]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.eo_cmwc5_from_time.argtypes = [C.POINTER(Cmwc5State), C.c_int64]
        L.eo_cmwc5_next_long.argtypes = [C.POINTER(Cmwc5State)]
        L.eo_cmwc5_next_long.restype = C.c_int64
        L.eo_cmwc5_next_int.argtypes = [C.POINTER(Cmwc5State)]
        L.eo_cmwc5_next_int.restype = C.c_int32
        L.eo_cmwc5_next_double.argtypes = [C.POINTER(Cmwc5State)]
        L.eo_cmwc5_next_double.restype = C.c_double
        L.eo_cmwc5_next_int_range.argtypes = [C.POINTER(Cmwc5State), C.c_int32, C.c_int32]
        L.eo_cmwc5_next_int_range.restype = C.c_int32
        L.eo_cmwc5_next_double_range.argtypes = [C.POINTER(Cmwc5State), C.c_double, C.c_double]
        L.eo_cmwc5_next_double_range.restype = C.c_double
        L.eo_sample_variance.argtypes = [_dp, C.c_int64]
        L.eo_sample_variance.restype = C.c_double
        L.eo_pop_variance.argtypes = [_dp, C.c_int64]
        L.eo_pop_variance.restype = C.c_double
        L.eo_gini_impurity.argtypes = [_ip, _dp, C.c_int64, C.c_int32]
        L.eo_gini_impurity.restype = C.c_double
        L.eo_gini_score.argtypes = [_ip, _dp, _bp, C.c_int64, C.c_double, C.c_int32]
        L.eo_gini_score.restype = C.c_double
        L.eo_variance_reduction.argtypes = [_dp, _bp, C.c_int64, C.c_double]
        L.eo_variance_reduction.restype = C.c_double
        L.eo_split_kat.argtypes = [_dp, C.c_int64, C.c_int32, _ip, C.c_int64, _ip, C.c_int32, C.c_int32,
                                   _ip, _dp, _dp, C.c_int32, C.c_int32, C.c_int64, _ip, _dp, _ip, _ip]
        L.eo_build_classification.argtypes = [_dp, C.c_int64, C.c_int32, _ip, C.c_int64, _dp, C.c_int32,
                                              C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                              C.c_int32, C.c_int64, C.c_int32, C.c_int32]
        L.eo_build_classification.restype = C.c_void_p
        L.eo_build_regression.argtypes = [_dp, C.c_int64, C.c_int32, _dp, C.c_int64, C.c_int32, C.c_int32,
                                          C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                          C.c_int32, C.c_int32]
        L.eo_build_regression.restype = C.c_void_p
        L.eo_forest_free.argtypes = [C.c_void_p]
        L.eo_forest_num_trees.argtypes = [C.c_void_p]
        L.eo_forest_leaf_width.argtypes = [C.c_void_p]
        L.eo_forest_next_long_after.argtypes = [C.c_void_p]
        L.eo_forest_next_long_after.restype = C.c_int64
        L.eo_tree_size.argtypes = [C.c_void_p, C.c_int32]
        L.eo_tree_trace_size.argtypes = [C.c_void_p, C.c_int32]
        L.eo_tree_trace_size.restype = C.c_int64
        L.eo_tree_export.argtypes = [C.c_void_p, C.c_int32, _ip, _dp, _bp, _ip, _ip, _dp]
        L.eo_tree_trace_export.argtypes = [C.c_void_p, C.c_int32, _lp, _ip, _ip, _dp, _bp]
        L.eo_forest_stats.argtypes = [C.c_void_p, _lp]
        L.eo_forest_import.argtypes = [C.c_int32, C.c_int32, C.c_int32, _ip, _ip, _dp, _bp, _ip, _ip, _dp]
        L.eo_forest_import.restype = C.c_void_p
        L.eo_predict_classification.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int32, _dp]
        L.eo_predict_regression.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int32, _dp]
        _lib = L
    return _lib


class Cmwc5:
    """org.saddle.spire.random.rng.Cmwc5 (restated)."""

    def __init__(self, seed: int):
        self.s = Cmwc5State()
        lib().eo_cmwc5_from_time(C.byref(self.s), C.c_int64(_wrap64(seed)))

    def next_long(self) -> int:
        return lib().eo_cmwc5_next_long(C.byref(self.s))

    def next_int(self) -> int:
        return lib().eo_cmwc5_next_int(C.byref(self.s))

    def next_double(self) -> float:
        return lib().eo_cmwc5_next_double(C.byref(self.s))

    def next_int_range(self, lo: int, hi: int) -> int:
        return lib().eo_cmwc5_next_int_range(C.byref(self.s), lo, hi)

    def next_double_range(self, a: float, b: float) -> float:
        return lib().eo_cmwc5_next_double_range(C.byref(self.s), a, b)


def _wrap64(v: int) -> int:
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _u8(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint8)


def sample_variance(v) -> float:
    v = _f64(v)
    return lib().eo_sample_variance(_ptr(v, _dp), len(v))


def pop_variance(v) -> float:
    v = _f64(v)
    return lib().eo_pop_variance(_ptr(v, _dp), len(v))


def gini_impurity(target, weights, num_classes) -> float:
    t, w = _i32(target), _f64(weights)
    return lib().eo_gini_impurity(_ptr(t, _ip), _ptr(w, _dp), len(t), num_classes)


def gini_score(target, weights, mask, g_nosplit, num_classes) -> float:
    t, w, m = _i32(target), _f64(weights), _u8(mask)
    return lib().eo_gini_score(_ptr(t, _ip), _ptr(w, _dp), _ptr(m, _bp), len(t), g_nosplit, num_classes)


def variance_reduction(target, mask, var_nosplit) -> float:
    t, m = _f64(target), _u8(mask)
    return lib().eo_variance_reduction(_ptr(t, _dp), _ptr(m, _bp), len(t), var_nosplit)


def split(data_rowmajor, subset, attributes, num_constant, k, target_at_subset, weights_at_subset=None,
          num_classes=0, regression=False, best=False, rng_seed=0):
    """Direct call of splitClassification / splitRegression / splitBest* (pkg:56-511).
    `attributes` (np.int32 array) is mutated in place like the reference's Array[Int]."""
    x = _f64(data_rowmajor)
    n_rows, d = x.shape
    sub = _i32(subset)
    assert attributes.dtype == np.int32 and attributes.flags.c_contiguous
    tc = None if regression else _i32(target_at_subset)
    tr = _f64(target_at_subset) if regression else None
    w = _f64(weights_at_subset)
    f, cut, nc, mil = C.c_int32(), C.c_double(), C.c_int32(), C.c_int32()
    lib().eo_split_kat(_ptr(x, _dp), n_rows, d, _ptr(sub, _ip), len(sub), _ptr(attributes, _ip),
                       num_constant, k, _ptr(tc, _ip), _ptr(tr, _dp), _ptr(w, _dp), num_classes,
                       int(best), C.c_int64(_wrap64(rng_seed)), C.byref(f), C.byref(cut), C.byref(nc),
                       C.byref(mil))
    return f.value, cut.value, nc.value, bool(mil.value)


@dataclass
class FlatTree:
    """One tree as pre-order arrays (the wire format shared with the C ABI, include/etgpu.h)."""
    feature: np.ndarray  # int32, -1 = leaf
    cut: np.ndarray      # float64
    mil: np.ndarray      # uint8 splitMissingIsLess
    left: np.ndarray     # int32
    right: np.ndarray    # int32
    leaf: np.ndarray     # float64 [n_nodes, leaf_width]

    @property
    def n_nodes(self):
        return len(self.feature)


@dataclass
class Trace:
    """Replay trace of one tree: per pre-order node the ordered candidate draws."""
    cand_begin: np.ndarray    # int64 [n_nodes]
    cand_count: np.ndarray    # int32 [n_nodes]
    cand_feature: np.ndarray  # int32
    cand_u: np.ndarray        # float64 raw nextDouble(); NaN when none drawn
    cand_flag: np.ndarray     # uint8 0 const / 1 scored / 2 scored-NaN


STAT_NAMES = ("v_mm", "v_sc", "s_rows", "p_rows", "draws", "const_hits", "scored", "nodes")


class Forest:
    def __init__(self, handle, regression):
        if not handle:
            raise ValueError("requirement failed")  # the reference's require(...) -> IllegalArgumentException
        self.h = handle
        self.regression = regression
        self.m = lib().eo_forest_num_trees(handle)
        self.leaf_width = lib().eo_forest_leaf_width(handle)

    def __del__(self):
        try:
            if self.h:
                lib().eo_forest_free(self.h)
                self.h = None
        except Exception:
            pass

    def tree(self, t) -> FlatTree:
        n = lib().eo_tree_size(self.h, t)
        ft = FlatTree(np.empty(n, np.int32), np.empty(n, np.float64), np.empty(n, np.uint8),
                      np.empty(n, np.int32), np.empty(n, np.int32),
                      np.empty((n, self.leaf_width), np.float64))
        lib().eo_tree_export(self.h, t, _ptr(ft.feature, _ip), _ptr(ft.cut, _dp), _ptr(ft.mil, _bp),
                             _ptr(ft.left, _ip), _ptr(ft.right, _ip), _ptr(ft.leaf, _dp))
        return ft

    def trees(self):
        return [self.tree(t) for t in range(self.m)]

    def trace(self, t) -> Trace:
        n = lib().eo_tree_size(self.h, t)
        c = lib().eo_tree_trace_size(self.h, t)
        tr = Trace(np.empty(n, np.int64), np.empty(n, np.int32), np.empty(c, np.int32),
                   np.empty(c, np.float64), np.empty(c, np.uint8))
        lib().eo_tree_trace_export(self.h, t, _ptr(tr.cand_begin, _lp), _ptr(tr.cand_count, _ip),
                                   _ptr(tr.cand_feature, _ip), _ptr(tr.cand_u, _dp), _ptr(tr.cand_flag, _bp))
        return tr

    def stats(self) -> dict:
        out = np.zeros(8, np.int64)
        lib().eo_forest_stats(self.h, _ptr(out, _lp))
        return dict(zip(STAT_NAMES, (int(v) for v in out)))

    @property
    def next_long_after(self) -> int:
        return lib().eo_forest_next_long_after(self.h)

    def predict(self, x_rowmajor):
        x = _f64(x_rowmajor)
        n, d = x.shape
        if self.regression:
            out = np.empty(n, np.float64)
            lib().eo_predict_regression(self.h, _ptr(x, _dp), n, d, _ptr(out, _dp))
        else:
            out = np.empty((n, self.leaf_width), np.float64)
            lib().eo_predict_classification(self.h, _ptr(x, _dp), n, d, _ptr(out, _dp))
        return out


INT_MAX = 2**31 - 1


def build_forest_classification(data, target, sample_weights, num_classes, n_min, k, m, parallelism,
                                best_split=False, max_depth=INT_MAX, seed=0, record_trace=False,
                                n_threads=0) -> Forest:
    x, y, w = _f64(data), _i32(target), _f64(sample_weights)
    n, d = x.shape
    h = lib().eo_build_classification(_ptr(x, _dp), n, d, _ptr(y, _ip), len(y), _ptr(w, _dp), num_classes,
                                      n_min, k, m, parallelism, int(best_split), max_depth,
                                      C.c_int64(_wrap64(seed)), int(record_trace), n_threads)
    return Forest(h, False)


def build_forest_regression(data, target, n_min, k, m, parallelism, best_split=False, max_depth=INT_MAX,
                            seed=0, record_trace=False, n_threads=0) -> Forest:
    x, y = _f64(data), _f64(target)
    n, d = x.shape
    h = lib().eo_build_regression(_ptr(x, _dp), n, d, _ptr(y, _dp), len(y), n_min, k, m, parallelism,
                                  int(best_split), max_depth, C.c_int64(_wrap64(seed)), int(record_trace),
                                  n_threads)
    return Forest(h, True)


def import_forest(trees: list[FlatTree], regression: bool) -> Forest:
    sizes = np.array([t.n_nodes for t in trees], np.int32)
    cat = lambda f, dt: np.ascontiguousarray(np.concatenate([getattr(t, f) for t in trees]), dtype=dt)
    feature, cut, mil = cat("feature", np.int32), cat("cut", np.float64), cat("mil", np.uint8)
    left, right, leaf = cat("left", np.int32), cat("right", np.int32), cat("leaf", np.float64)
    lw = trees[0].leaf.shape[1]
    h = lib().eo_forest_import(len(trees), lw, int(regression), _ptr(sizes, _ip), _ptr(feature, _ip),
                               _ptr(cut, _dp), _ptr(mil, _bp), _ptr(left, _ip), _ptr(right, _ip),
                               _ptr(leaf, _dp))
    return Forest(h, regression)


def tree_checksum(t: FlatTree) -> str:
    """SURVEY Appendix B checksum: sha256 over the pre-order walk."""
    import hashlib
    import struct
    h = hashlib.sha256()
    # arrays are already pre-order (left subtree directly follows its parent)
    stack = [0]
    while stack:
        i = stack.pop()
        if t.feature[i] >= 0:
            h.update(b"N" + struct.pack("<idB", int(t.feature[i]), float(t.cut[i]), int(t.mil[i])))
            stack.append(int(t.right[i]))
            stack.append(int(t.left[i]))
        else:
            h.update(b"L" + struct.pack("<%dd" % t.leaf.shape[1], *t.leaf[i]))
    return h.hexdigest()[:16]


def tree_depth(t: FlatTree) -> int:
    depth = np.zeros(t.n_nodes, np.int64)
    md = 0
    for i in range(t.n_nodes):  # pre-order: parent precedes children
        if t.feature[i] >= 0:
            depth[t.left[i]] = depth[i] + 1
            depth[t.right[i]] = depth[i] + 1
        md = max(md, int(depth[i]))
    return md
