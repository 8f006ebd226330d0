"""lamp_b200 -- B200-native (sm_100a) implementation of the extratrees hot path of pityka/lamp.

Only what the path needs lives here: csrc/ (CUDA kernels + the C ABI of include/etgpu.h), the ctypes
binding, the host-side mirror of the reference's public API, and tree sharding across GPUs."""
from .extratrees import (ClassificationLeaf, ClassificationNonLeaf, Context, DeviceData, FlatTree, Forest,  # noqa: F401
                         RegressionLeaf, RegressionNonLeaf, buildForestClassification, buildForestRegression,
                         default_context, make_replay, predictClassification, predictRegression)
from ._capi import EtError  # noqa: F401
