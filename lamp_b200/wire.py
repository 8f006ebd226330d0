"""JSON wire format of the tree ADTs: the reference derives upickle `ReadWriter`s for its case
classes with `macroRW` (extratrees/src/main/scala/lamp/forest/extratrees.scala:10-63, upickle 3.1.4 per build.sbt:84).
upickle 3.x writes a member of a sealed hierarchy as a JSON object whose first key `$type` holds the fully
qualified class name, followed by the constructor fields by name; `Seq[Double]` is an array; non-finite doubles are
the strings "NaN" / "Infinity" / "-Infinity"; whole doubles may appear without a fraction ("1").

PARITY UNPINNED: no JVM exists in this image, so the shape below follows upickle's documented conventions and the
field names of the case classes; it has not been checked against bytes produced by the reference.  Reading accepts
everything upickle 3.x or 4.x writes for these classes (full or short `$type`, strings or numbers for doubles).

Pure host code: no GPU, no dependency on the C library."""
from __future__ import annotations

import json
import math

from .extratrees import ClassificationLeaf, ClassificationNonLeaf, RegressionLeaf, RegressionNonLeaf

PKG = "lamp.extratrees."


def _num(x: float):
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    return float(x)


def _dbl(v) -> float:
    return float(v)  # float("NaN"), float("Infinity"), float("-Infinity") parse upickle's strings


def to_obj(t):
    """ADT tree -> plain dict / list structure (json.dumps-able).  Iterative: trees can be hundreds of levels deep."""
    out = {}
    stack = [(t, out)]
    while stack:
        node, dst = stack.pop()
        if isinstance(node, ClassificationLeaf):
            dst["$type"] = PKG + "ClassificationLeaf"
            dst["targetDistribution"] = [_num(v) for v in node.targetDistribution]
        elif isinstance(node, RegressionLeaf):
            dst["$type"] = PKG + "RegressionLeaf"
            dst["targetMean"] = _num(node.targetMean)
        elif isinstance(node, (ClassificationNonLeaf, RegressionNonLeaf)):
            dst["$type"] = PKG + type(node).__name__
            dst["left"], dst["right"] = {}, {}
            dst["splitFeature"] = int(node.splitFeature)
            dst["cutpoint"] = _num(node.cutpoint)
            dst["splitMissingIsLess"] = bool(node.splitMissingIsLess)
            stack.append((node.left, dst["left"]))
            stack.append((node.right, dst["right"]))
        else:
            raise TypeError(f"not a tree node: {type(node).__name__}")
    return out


def from_obj(o):
    """Inverse of to_obj (also accepts short `$type` names).  Children are built bottom-up without recursion."""
    order, stack = [], [o]
    while stack:  # pre-order list of the dict nodes
        d = stack.pop()
        order.append(d)
        kind = d["$type"].rsplit(".", 1)[-1]
        if kind.endswith("NonLeaf"):
            stack.append(d["right"])
            stack.append(d["left"])
    built = {}
    for d in reversed(order):
        kind = d["$type"].rsplit(".", 1)[-1]
        if kind == "ClassificationLeaf":
            node = ClassificationLeaf(tuple(_dbl(v) for v in d["targetDistribution"]))
        elif kind == "RegressionLeaf":
            node = RegressionLeaf(_dbl(d["targetMean"]))
        elif kind in ("ClassificationNonLeaf", "RegressionNonLeaf"):
            cls = ClassificationNonLeaf if kind[0] == "C" else RegressionNonLeaf
            node = cls(built.pop(id(d["left"])), built.pop(id(d["right"])), int(d["splitFeature"]), _dbl(d["cutpoint"]),
                       bool(d["splitMissingIsLess"]))
        else:
            raise ValueError(f"unknown $type {d['$type']!r}")
        built[id(d)] = node
    return built[id(o)]


def tree_to_json(t) -> str:
    return json.dumps(to_obj(t), separators=(",", ":"))


def tree_from_json(s: str):
    return from_obj(json.loads(s))


def forest_to_json(trees) -> str:
    """`Seq[ClassificationTree]` / `Seq[RegressionTree]` -> JSON array (upickle's encoding of a Seq)."""
    return json.dumps([to_obj(t) for t in trees], separators=(",", ":"))


def forest_from_json(s: str):
    return [from_obj(o) for o in json.loads(s)]
