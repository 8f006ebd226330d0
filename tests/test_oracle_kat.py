"""Pins the CPU oracle against every deterministic known-answer test the reference holds for the
extratrees path (extratrees/src/test/scala/lamp/forest/extratree.test.scala, cited tst:LINE), plus the
session-derived self-consistency vectors of SURVEY.md Appendix B."""
import math

import numpy as np
import pytest

from oracle import oracle as O

NAN = float("nan")


def cols(*c):
    """saddle Mat(Vec, Vec, ...) builds a matrix from COLUMNS; return row-major [n, d]."""
    return np.ascontiguousarray(np.stack([np.asarray(x, np.float64) for x in c], axis=1))


X2 = cols([0, 2, 3, 4, 5], [100, 99, 98, 97, 96])
X2b = cols([0, 0, 3, 3, 3], [100, 99, 98, 97, 96])
X3 = cols([0, 0, 3, 3, 3], [0, 0, 3, 3, 3], [100, 99, 98, 97, 96])


def test_variance_reduction():  # tst:7-15
    t = np.array([0.0, 0.01, 100.0, 100.1])
    v = O.sample_variance(t) * (len(t) - 1.0) / len(t)
    assert O.variance_reduction(t, [1, 1, 0, 0], v) == 0.999999495454448


def test_gini_impurity():  # tst:16-30
    t = [1, 1, 0, 0]
    gt = O.gini_impurity(t, None, 2)
    assert O.gini_score(t, None, [1, 1, 0, 0], gt, 2) == 0.5


def test_gini_impurity_weighted():  # tst:31-45
    t = [1, 1, 0, 0]
    gt = O.gini_impurity(t, np.ones(4), 2)
    assert O.gini_score(t, None, [1, 1, 0, 0], gt, 2) == 0.5


def test_gini_zero_weighted():  # tst:46-60
    t = [1, 1, 0, 0]
    gt = O.gini_impurity(t, np.zeros(4), 2)
    assert math.isnan(O.gini_score(t, np.zeros(4), [1, 1, 0, 0], gt, 2))


def test_gini_zero_weighted_2():  # tst:61-75
    t = [1, 1, 0, 0]
    gt = O.gini_impurity(t, np.ones(4), 2)
    assert math.isnan(O.gini_score(t, [1.0, 1.0, 0.0, 0.0], [1, 1, 0, 0], gt, 2))


def test_split_regression_1():  # tst:76-88
    attr = np.array([0, 1], np.int32)
    r = O.split(X2, [0, 1, 2, 3, 4], attr, 0, 2, [0.0, 0.1, 100.0, 100.1, 100.2], regression=True, rng_seed=0)
    assert r == (0, 3.424021023861243, 0, False)


def test_split_best_regression_1():  # tst:89-101
    attr = np.array([0, 1], np.int32)
    r = O.split(X2, [0, 1, 2, 3, 4], attr, 0, 2, [0.0, 0.1, 0.1, 100.1, 100.2], regression=True, best=True,
                rng_seed=0)
    assert r == (0, 4.0, 0, False)


def test_split_classification_1():  # tst:102-117
    attr = np.array([0, 1], np.int32)
    r = O.split(X2, [0, 1, 2, 3, 4], attr, 0, 2, [1, 1, 0, 0, 0], None, 2, rng_seed=0)
    assert attr.tolist() == [1, 0]
    assert r == (0, 3.424021023861243, 0, False)


def test_split_best_classification():  # tst:118-133
    attr = np.array([0, 1], np.int32)
    r = O.split(X2, [0, 1, 2, 3, 4], attr, 0, 2, [1, 1, 1, 0, 0], None, 2, best=True, rng_seed=0)
    assert attr.tolist() == [1, 0]
    assert r == (0, 4.0, 0, False)


def test_split_classification_1_weighted():  # tst:134-149
    attr = np.array([0, 1], np.int32)
    r = O.split(X2, [0, 1, 2, 3, 4], attr, 0, 2, [1, 1, 0, 0, 0], np.ones(5), 2, rng_seed=0)
    assert attr.tolist() == [1, 0]
    assert r == (0, 3.424021023861243, 0, False)


def test_split_classification_1_zero_weighted():  # tst:150-165
    attr = np.array([0, 1], np.int32)
    r = O.split(X2, [0, 1, 2, 3, 4], attr, 0, 2, [1, 1, 0, 0, 0], [1.0, 1.0, 0.0, 0.0, 0.0], 2, rng_seed=0)
    assert attr.tolist() == [0, 1]
    assert r[0] == -1


def test_split_classification_2():  # tst:166-181
    attr = np.array([0, 1], np.int32)
    r = O.split(X2b, [2, 3, 4], attr, 0, 2, [1, 1, 0], None, 2, rng_seed=0)
    assert r == (1, 97.54668482609304, 1, False)
    assert attr.tolist() == [0, 1]


@pytest.mark.parametrize("attr0,nc,k,seed,expect,attr1", [
    ([2, 1, 0], 0, 2, 0, (2, 97.54668482609304, 2, False), [0, 1, 2]),    # tst:182-201
    ([2, 0, 1], 0, 1, 0, (2, 97.54668482609304, 1, False), [1, 0, 2]),    # tst:202-221
    ([0, 2, 1], 1, 1, 1, (2, 97.84900936098786, 1, False), [0, 1, 2]),    # tst:222-241
    ([0, 2, 1], 1, 1, 123, (2, 96.07259095141863, 2, False), [0, 1, 2]),  # tst:242-261
    ([1, 2, 0], 1, 1, 123, (2, 96.07259095141863, 2, False), [1, 0, 2]),  # tst:262-281
])
def test_split_classification_3_to_7(attr0, nc, k, seed, expect, attr1):
    attr = np.array(attr0, np.int32)
    r = O.split(X3, [2, 3, 4], attr, nc, k, [1, 1, 0], None, 2, rng_seed=seed)
    assert r == expect
    assert attr.tolist() == attr1


def test_missing_regression():  # tst:515-533
    x = np.array([[1.0], [1.0], [1.0], [1.0], [NAN], [NAN], [NAN]])
    y = np.array([1.0, 1, 1, 1, 0, 0, 0])
    f = O.build_forest_regression(x, y, n_min=1, k=1, m=100, parallelism=1, max_depth=200, seed=1234567)
    assert np.array_equal(f.predict(x), y)
    t = f.tree(0)
    assert t.feature.tolist() == [0, -1, -1] and t.cut[0] == 1.0 and t.mil[0] == 1
    assert t.leaf[1, 0] == 0.0 and t.leaf[2, 0] == 1.0  # left = NaN rows, right = 1-rows


def test_missing_classification():  # tst:534-554
    x = np.array([[1.0], [1.0], [1.0], [1.0], [NAN], [NAN], [NAN]])
    y = np.array([1, 1, 1, 1, 0, 0, 0], np.int32)
    f = O.build_forest_classification(x, y, None, 2, n_min=1, k=1, m=100, parallelism=1, max_depth=200,
                                      seed=987)
    assert np.array_equal(f.predict(x)[:, 1], y.astype(np.float64))


def test_require_failures():  # pkg:624-633, 715-718
    x = np.zeros((4, 2))
    with pytest.raises(ValueError):
        O.build_forest_classification(x, np.zeros(3, np.int32), None, 2, 2, 1, 1, 1)
    with pytest.raises(ValueError):
        O.build_forest_classification(x, np.zeros(4, np.int32), [1.0, -1.0, 1.0, 1.0], 2, 2, 1, 1, 1)
    with pytest.raises(ValueError):
        O.build_forest_regression(x, np.zeros(5), 2, 1, 1, 1)


# ---- SURVEY.md Appendix B (second, independently written restatement; not reference-pinned) ----
def test_rng_vectors():
    g = O.Cmwc5(0)
    assert [g.next_long() for _ in range(3)] == [4193861061333511706, -5814336167675693050, -4181094498895388101]
    g = O.Cmwc5(42)
    assert [g.next_long() for _ in range(3)] == [-6400204466741627489, 3288247062548297731, 8440593475565983106]
    g = O.Cmwc5(0)
    assert [g.next_int_range(0, 783) for _ in range(5)] == [698, 383, 84, 421, 776]
    assert [g.next_double() for _ in range(2)] == [0.8301927100331622, 0.4096298110685359]


APPENDIX_B = [
    ("cls", 2000, 0, (434, 60.40319931377828), 1063, 532, 17, "3760c9fbe55dd6f2", -6773734371773005106),
    ("reg", 2000, 0, (716, 22.718987746596184), 1349, 675, 23, "027212123b5280c4", -388517389295671167),
    ("cls", 2000, 42, (350, 77.03409179780358), 1051, 526, 20, "553df54ada497da3", 6114440105907361848),
    ("reg", 2000, 42, (236, 56.6647598119146), 1345, 673, 23, "8fc28ae0a5354011", 1533558803129324859),
    ("cls", 10000, 0, (484, 47.43724019171581), 3679, 1840, 27, "77469fd974dc2812", 7960988403183244520),
    ("reg", 10000, 0, (382, 22.718987746596184), 4501, 2251, 31, "3b147f6e7e5c0c5d", 5247006346217893389),
    ("cls", 10000, 42, (429, 206.97786138744425), 3635, 1818, 24, "661a51133d5bd825", -401303695535431327),
    ("reg", 10000, 42, (236, 56.6647598119146), 4767, 2384, 28, "0a846825b2f60ce2", 3674479768458645999),
]


@pytest.mark.parametrize("task,rows,seed,root,nodes,leaves,depth,checksum,next_long", APPENDIX_B)
def test_appendix_b(mnist, task, rows, seed, root, nodes, leaves, depth, checksum, next_long):
    x, y = mnist
    x, y = x[:rows], y[:rows]
    if task == "cls":
        f = O.build_forest_classification(x, y, None, 10, n_min=2, k=32, m=1, parallelism=1, max_depth=200,
                                          seed=seed)
    else:
        f = O.build_forest_regression(x, y.astype(np.float64), n_min=2, k=32, m=1, parallelism=1,
                                      max_depth=200, seed=seed)
    t = f.tree(0)
    assert (int(t.feature[0]), float(t.cut[0]), int(t.mil[0])) == (root[0], root[1], 0)
    assert t.n_nodes == nodes
    assert int((t.feature < 0).sum()) == leaves
    assert O.tree_depth(t) == depth
    assert O.tree_checksum(t) == checksum
    assert f.next_long_after == next_long
    # reference property tests tst:283-319 / 442-477: one tree fits its training set exactly
    p = f.predict(x)
    if task == "cls":
        assert np.array_equal(p.argmax(axis=1), y)
    else:
        assert np.array_equal(p.astype(np.int64), y)
