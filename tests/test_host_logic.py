"""CPU-side checks of the product's host logic: the C ABI loads and exports every symbol of
include/etgpu.h, argument errors mirror the reference's require(...), ADT <-> flat conversions, and
the closed-form repeated addition the kernels use for pkg:905-911."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import lamp_b200 as et
from lamp_b200 import _capi as capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "etgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(et_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = capi.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"libetgpu.so does not export {name}"
    assert declared == set(capi.SIGNATURES), declared ^ set(capi.SIGNATURES)
    assert L.et_abi_version() == capi.ABI_VERSION


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.EtError) as e:
        et.Context(0)
    assert e.value.code == capi.ET_ECUDA


def naive_repeat_add(c, h):
    acc = np.float64(0.0)
    c = np.float64(c)
    for _ in range(h):
        acc = acc + c
    return float(acc)


@pytest.mark.parametrize("s", [1, 2, 3, 7, 10, 49, 100, 1000, 4096, 10007, 60000, 65536, 99991])
def test_repeat_add_matches_sequential_loop(s):
    L = capi.lib()
    c = 1.0 / s
    hs = sorted(set([0, 1, 2, 3, 23, 24, 25, 26, 50, s // 3, s // 2, s - 1, s]))
    for h in hs:
        if h < 0 or h > s:
            continue
        assert L.et_debug_repeat_add(c, h) == naive_repeat_add(c, h), (s, h)


def test_repeat_add_random():
    L = capi.lib()
    rng = np.random.default_rng(0)
    for _ in range(300):
        s = int(rng.integers(1, 200000))
        h = int(rng.integers(0, s + 1))
        assert L.et_debug_repeat_add(1.0 / s, h) == naive_repeat_add(1.0 / s, h), (s, h)


def test_repeat_add_large_counts():
    # 10M-row tables: compare against numpy's sequential cumsum (exact same additions)
    L = capi.lib()
    for s in (10_000_000, 7_654_321, 2**23):
        c = np.float64(1.0 / s)
        cs = np.cumsum(np.full(s, c))  # cumsum accumulates left to right in float64
        for h in (s, s // 2, s // 7, 123457):
            assert L.et_debug_repeat_add(float(c), h) == float(cs[h - 1]), (s, h)


def test_adt_roundtrip():
    t = et.ClassificationNonLeaf(
        et.ClassificationLeaf((1.0, 0.0)),
        et.ClassificationNonLeaf(et.ClassificationLeaf((0.25, 0.75)), et.ClassificationLeaf((0.0, 1.0)), 3, 0.5, True),
        1, -2.0, False)
    from lamp_b200.extratrees import adt_to_flat, flat_to_adt
    f = adt_to_flat(t, 2)
    assert f.feature.tolist() == [1, -1, 3, -1, -1]
    assert f.left.tolist() == [1, -1, 3, -1, -1] and f.right.tolist() == [2, -1, 4, -1, -1]
    assert flat_to_adt(f, False) == t
    r = et.RegressionNonLeaf(et.RegressionLeaf(1.5), et.RegressionLeaf(-1.0), 0, 0.0, False)
    assert flat_to_adt(adt_to_flat(r, 1), True) == r


def test_make_replay_offsets():
    a = dict(left=[1, -1, -1], right=[2, -1, -1], cand_begin=[0, 2, 2], cand_count=[2, 0, 0],
             cand_feature=[4, 5], cand_u=[0.5, np.nan], cand_flag=[1, 0])
    r, keep = et.make_replay([a, a])
    assert r.n_trees == 2 and r.n_cand == 4
    assert keep["node_offset"].tolist() == [0, 3, 6]
    assert keep["cand_begin"].tolist() == [0, 2, 2, 2, 4, 4]


# ---- JSON wire format of the tree ADTs (lamp_b200/wire.py; upickle macroRW shape, adt:10-63) -----------------
def test_wire_format_shape_and_roundtrip():
    import json
    import math
    from lamp_b200 import wire
    import lamp_b200 as et
    t = et.ClassificationNonLeaf(et.ClassificationLeaf((1.0, 0.0)),
                                 et.ClassificationNonLeaf(et.ClassificationLeaf((float("nan"), 0.25)),
                                                          et.ClassificationLeaf((0.0, 1.0)), 3, -0.0, False),
                                 0, 97.54668482609304, True)
    s = wire.tree_to_json(t)
    o = json.loads(s)
    assert list(o)[0] == "$type" and o["$type"] == "lamp.extratrees.ClassificationNonLeaf"
    assert list(o) == ["$type", "left", "right", "splitFeature", "cutpoint", "splitMissingIsLess"]  # constructor order
    assert o["left"] == {"$type": "lamp.extratrees.ClassificationLeaf", "targetDistribution": [1.0, 0.0]}
    assert o["right"]["left"]["targetDistribution"][0] == "NaN"  # upickle writes non-finite doubles as strings
    back = wire.tree_from_json(s)
    assert back.cutpoint == t.cutpoint and back.splitMissingIsLess and back.right.splitFeature == 3
    assert math.copysign(1.0, back.right.cutpoint) == -1.0  # -0.0 survives
    assert math.isnan(back.right.left.targetDistribution[0]) and back.right.right == t.right.right
    # short type names (upickle 4 default) and whole numbers without a fraction are accepted on input
    r = wire.tree_from_json('{"$type":"RegressionNonLeaf","left":{"$type":"RegressionLeaf","targetMean":1},'
                            '"right":{"$type":"RegressionLeaf","targetMean":"-Infinity"},"splitFeature":2,'
                            '"cutpoint":0.5,"splitMissingIsLess":false}')
    assert r == et.RegressionNonLeaf(et.RegressionLeaf(1.0), et.RegressionLeaf(float("-inf")), 2, 0.5, False)


def test_wire_format_forest_roundtrip_is_bit_exact():
    from lamp_b200 import wire
    from lamp_b200.extratrees import FlatTree, flat_to_adt
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    x = rng.normal(size=(300, 5))
    y = x[:, 0] * 2 + np.sin(x[:, 1])
    of = O.build_forest_regression(x, y, 2, 3, 4, 2, seed=7)
    trees = [flat_to_adt(FlatTree(t.feature, t.cut, t.mil, t.left, t.right, t.leaf), True) for t in of.trees()]
    back = wire.forest_from_json(wire.forest_to_json(trees))
    assert back == trees  # dataclass equality: every cutpoint and leaf mean identical as floats


def test_jni_shim_compiles_against_the_header_and_matches_the_scala_natives():
    """jni/etgpu_jni.c is syntax-checked with gcc against include/etgpu.h and a declaration-only jni.h (this image
    has no JDK); its exported names are exactly the @native methods of scala/lamp/extratrees/gpu/EtGpu.scala."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "jni", "etgpu_jni.c")
    subprocess.check_call(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(root, "jni", "stub"),
                           "-I" + os.path.join(root, "include"), src])
    exported = set(re.findall(r"Java_lamp_extratrees_gpu_Native_(\w+)\(", open(src).read()))
    scala = open(os.path.join(root, "scala", "lamp", "extratrees", "gpu", "EtGpu.scala")).read()
    natives = set(re.findall(r"@native def (\w+)\(", scala))
    assert exported == natives and len(natives) >= 9, (exported ^ natives)
    header = open(os.path.join(root, "include", "etgpu.h")).read()
    for fn in set(re.findall(r"\b(et_\w+)\(", open(src).read())):
        assert re.search(r"\b%s\(" % fn, header), fn
