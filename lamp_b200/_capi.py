"""ctypes binding of include/etgpu.h (libetgpu.so).  This is the binding a JNI/Panama shim mirrors
(INTEGRATION.md).  There is no fallback: if the CUDA library is missing, importing fails loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libetgpu.so")

ET_OK, ET_EINVAL, ET_ECUDA, ET_ENOMEM, ET_EREPLAY, ET_EUNSUPPORTED, ET_ENCCL = 0, -1, -2, -3, -4, -5, -6
ABI_VERSION = 2
COMM_ID_BYTES = 128

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)
lp = C.POINTER(C.c_int64)
bp = C.POINTER(C.c_uint8)
vp = C.c_void_p


class EtReplay(C.Structure):
    _fields_ = [("n_trees", C.c_int32), ("node_offset", lp), ("cand_begin", lp), ("cand_count", ip),
                ("left", ip), ("right", ip), ("n_cand", C.c_int64), ("cand_feature", ip), ("cand_u", dp),
                ("cand_flag", bp)]


class EtStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("v_mm", "v_sc", "s_rows", "p_rows", "draws", "const_hits", "scored",
                                         "nodes", "levels", "rounds", "launches", "replay_mismatches")] + \
               [(n, C.c_double) for n in ("gpu_ms", "gpu_ms_split", "gpu_ms_partition")] + \
               [(n, C.c_int64) for n in ("parallel_sum_nodes", "ambiguous_splits")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class EtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libetgpu error {code}: {msg}")
        self.code = code


# every symbol include/etgpu.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "et_abi_version": (C.c_int32, []),
    "et_last_error": (C.c_char_p, []),
    "et_init": (C.c_int, [C.c_int32, C.POINTER(vp)]),
    "et_init_multi": (C.c_int, [ip, C.c_int32, C.POINTER(vp)]),
    "et_device_count": (C.c_int32, [vp]),
    "et_shutdown": (None, [vp]),
    "et_set_stream": (C.c_int, [vp, vp]),
    "et_synchronize": (C.c_int, [vp]),
    "et_data_dense_rowmajor": (C.c_int, [vp, dp, C.c_int64, C.c_int32, C.POINTER(vp)]),
    "et_data_dense_alloc": (C.c_int, [vp, C.c_int64, C.c_int32, C.POINTER(vp)]),
    "et_data_dense_colblock": (C.c_int, [vp, vp, dp, C.c_int32, C.c_int32]),
    "et_data_dense_rowmajor_device": (C.c_int, [vp, vp, C.c_int64, C.c_int32, C.POINTER(vp)]),
    "et_data_csc": (C.c_int, [vp, C.POINTER(C.c_int64), ip, dp, C.c_int64, C.c_int32, C.POINTER(vp)]),
    "et_data_set_target_classification": (C.c_int, [vp, vp, ip, C.c_int64, C.c_int32]),
    "et_data_set_target_regression": (C.c_int, [vp, vp, dp, C.c_int64]),
    "et_data_set_weights": (C.c_int, [vp, vp, dp, C.c_int64]),
    "et_data_dims": (C.c_int, [vp, lp, ip]),
    "et_data_free": (None, [vp]),
    "et_build_classification": (C.c_int, [vp, vp, ip, C.c_int64, dp, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_int32, C.c_int32, C.c_int32, C.c_int64, ip, C.POINTER(EtReplay),
                                          C.POINTER(vp), C.POINTER(EtStats)]),
    "et_build_regression": (C.c_int, [vp, vp, dp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_int64, ip, C.POINTER(EtReplay), C.POINTER(vp),
                                      C.POINTER(EtStats)]),
    "et_forest_dims": (C.c_int, [vp, ip, ip, ip, lp]),
    "et_forest_tree_size": (C.c_int, [vp, C.c_int32, ip]),
    "et_forest_export": (C.c_int, [vp, C.c_int32, ip, dp, bp, ip, ip, dp]),
    "et_forest_export_all": (C.c_int, [vp, ip, ip, dp, bp, ip, ip, dp]),
    "et_forest_import": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, ip, ip, dp, bp, ip, ip, dp, C.POINTER(vp)]),
    "et_forest_packed_dims": (C.c_int, [vp, lp, lp]),
    "et_forest_export_packed": (C.c_int, [vp, vp, vp, vp]),
    "et_forest_import_packed": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, vp, vp, vp,
                                          C.POINTER(vp)]),
    "et_forest_free": (None, [vp]),
    "et_predict_classification": (C.c_int, [vp, vp, dp, C.c_int64, C.c_int32, dp, C.c_int32]),
    "et_predict_regression": (C.c_int, [vp, vp, dp, C.c_int64, C.c_int32, dp, C.c_int32]),
    "et_predict_classification_device": (C.c_int, [vp, vp, vp, C.c_int64, C.c_int32, vp, C.c_int32]),
    "et_predict_regression_device": (C.c_int, [vp, vp, vp, C.c_int64, C.c_int32, vp, C.c_int32]),
    "et_comm_unique_id": (C.c_int, [bp]),
    "et_comm_init_rank": (C.c_int, [vp, C.c_int32, C.c_int32, bp]),
    "et_data_broadcast": (C.c_int, [vp, vp, C.c_int32, C.POINTER(vp)]),
    "et_forest_allgather": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "et_predict_classification_allreduce": (C.c_int, [vp, vp, vp, C.c_int64, C.c_int32, vp, C.c_int32]),
    "et_predict_regression_allreduce": (C.c_int, [vp, vp, vp, C.c_int64, C.c_int32, vp, C.c_int32]),
    "et_comm_last_ms": (C.c_double, [vp]),
    "et_debug_repeat_add": (C.c_double, [C.c_double, C.c_int64]),
}

_lib = None


def lib():
    """Loads libetgpu.so (built in-tree by __graft_entry__.build()).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C lamp_b200/csrc).  lamp_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        if L.et_abi_version() != ABI_VERSION:
            raise ImportError(f"libetgpu ABI {L.et_abi_version()} != binding ABI {ABI_VERSION}")
        _lib = L
    return _lib


def check(rc):
    if rc != ET_OK:
        msg = lib().et_last_error().decode("utf-8", "replace")
        if rc == ET_EINVAL:
            # the reference's require(...) failures are IllegalArgumentException
            raise ValueError(msg)
        raise EtError(rc, msg)


def ptr(a, t):
    return None if a is None else a.ctypes.data_as(t)
