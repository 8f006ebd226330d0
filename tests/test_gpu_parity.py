"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Replay mode (candidate features and threshold uniforms replayed from the reference's Cmwc5 stream):
tree structure, cutpoints and leaf values must be BIT-EXACT (the north star allows 1e-12 relative on
leaf values; these tests hold them to 0).  Free-running mode: accuracy/RMSE properties."""
import numpy as np
import pytest

import lamp_b200 as et
from oracle import oracle as O
from tests.helpers import (assert_trees_bit_exact, oracle_replay, synth_classification, synth_regression)

pytestmark = pytest.mark.gpu
NAN = float("nan")


@pytest.fixture(autouse=True, params=["codes", "fp64", "codes-onecta", "fp64-onecta"])
def table_coding(request, monkeypatch):
    """Every test runs with the byte-coded copy of the table (encode.cu; used whenever all columns have <= 255
    distinct values) and with FP64 gathers only (ETGPU_NO_CODES=1), each with nodes of more than 2048 rows cut into
    chunks of rows over several CTAs (wide.cu, the default) and with one CTA per node whatever its size
    (ETGPU_WIDE_MIN beyond any table)."""
    monkeypatch.setenv("ETGPU_NO_CODES", "1" if request.param.startswith("fp64") else "0")
    monkeypatch.setenv("ETGPU_WIDE_MIN", "2147483647" if request.param.endswith("onecta") else "2048")
    return request.param


# ---- predict ----------------------------------------------------------------------------------------
def test_predict_classification_matches_oracle(mnist):
    x, y = mnist
    of = O.build_forest_classification(x[:3000], y[:3000], None, 10, 2, 28, 7, 4, seed=3)
    gf = et.Forest.from_trees(of.trees())
    p_gpu = et.predictClassification(gf, x[3000:5000])
    p_ora = of.predict(x[3000:5000])
    assert np.array_equal(p_gpu, p_ora)  # same additions in tree order, same division


def test_predict_regression_matches_oracle():
    x, y = synth_regression(4000, 12, 1, nan_frac=0.02)
    of = O.build_forest_regression(x, y, 5, 4, 9, 4, seed=5)
    gf = et.Forest.from_trees(of.trees(), regression_hint=True)
    assert np.array_equal(et.predictRegression(gf, x), of.predict(x))


def test_predict_accepts_adt_trees():
    leaf = et.ClassificationLeaf
    t = et.ClassificationNonLeaf(leaf((1.0, 0.0)), leaf((0.0, 1.0)), 0, 1.0, True)
    x = np.array([[0.5], [1.5], [NAN]])
    assert et.predictClassification([t, t], x).tolist() == [[1.0, 0.0], [0.0, 1.0], [1.0, 0.0]]


# ---- replay: classification ----------------------------------------------------------------------------
@pytest.mark.parametrize("rows,seed,k", [(2000, 0, 32), (2000, 42, 32), (10000, 0, 32), (10000, 7, 28), (3000, 5, 1)])
def test_replay_mnist_classification(mnist, rows, seed, k):
    x, y = mnist
    x, y = x[:rows], y[:rows]
    of = O.build_forest_classification(x, y, None, 10, 2, k, 1, 1, max_depth=200, seed=seed, record_trace=True)
    gf = et.buildForestClassification(x, y, None, 10, 2, k, 1, 1, maxDepth=200, seed=seed, replay=oracle_replay(of))
    assert_trees_bit_exact(gf, of)
    so, sg = of.stats(), gf.stats
    for key in ("v_mm", "v_sc", "s_rows", "p_rows", "draws", "const_hits", "scored", "nodes"):
        assert so[key] == sg[key], (key, so[key], sg[key])


def test_replay_forest_parallel_seeding(mnist):
    x, y = mnist
    x, y = x[:1500], y[:1500]
    of = O.build_forest_classification(x, y, None, 10, 2, 28, 12, 4, seed=99, record_trace=True)
    gf = et.buildForestClassification(x, y, None, 10, 2, 28, 12, 4, seed=99, replay=oracle_replay(of))
    assert_trees_bit_exact(gf, of)
    assert np.array_equal(et.predictClassification(gf, x), of.predict(x))


@pytest.mark.parametrize("nan_frac,max_depth,n_min", [(0.0, 2**31 - 1, 2), (0.05, 2**31 - 1, 2), (0.3, 6, 2)])
def test_replay_synthetic_classification(nan_frac, max_depth, n_min):
    x, y = synth_classification(5000, 20, 4, 11, nan_frac=nan_frac, const_cols=3, quantize=4)
    of = O.build_forest_classification(x, y, None, 4, n_min, 5, 6, 2, max_depth=max_depth, seed=1, record_trace=True)
    gf = et.buildForestClassification(x, y, None, 4, n_min, 5, 6, 2, maxDepth=max_depth, seed=1,
                                      replay=oracle_replay(of))
    assert_trees_bit_exact(gf, of)


def test_replay_all_nan_column_and_nmin_quirk():
    # all-NaN column: min=MaxValue, max=MinValue, cut=-inf/NaN -> NaN score -> "constant" (pkg:283-285)
    x, y = synth_classification(400, 5, 2, 3)
    x[:, 2] = NAN
    of = O.build_forest_classification(x, y, None, 2, 2, 3, 5, 2, seed=4, record_trace=True)
    gf = et.buildForestClassification(x, y, None, 2, 2, 3, 5, 2, seed=4, replay=oracle_replay(of))
    assert_trees_bit_exact(gf, of)
    # pkg:993: nMin compares the WHOLE TABLE's rows: nMin > n makes every root a leaf
    of = O.build_forest_classification(x, y, None, 2, 401, 3, 2, 2, seed=4, record_trace=True)
    gf = et.buildForestClassification(x, y, None, 2, 401, 3, 2, 2, seed=4, replay=oracle_replay(of))
    assert_trees_bit_exact(gf, of)
    assert gf.flat(0).n_nodes == 1


# ---- replay: weighted classification ---------------------------------------------------------------
@pytest.mark.parametrize("kind", ["half_zero", "real"])
def test_replay_weighted_classification(mnist, kind):
    x, y = mnist
    x, y = x[:2000], y[:2000]
    if kind == "half_zero":  # tst:357-395
        w = np.concatenate([np.ones(1000), np.zeros(1000)])
    else:
        w = np.random.default_rng(0).gamma(2.0, size=2000)
    of = O.build_forest_classification(x, y, w, 10, 2, 32, 2, 8, max_depth=200, seed=13, record_trace=True)
    gf = et.buildForestClassification(x, y, w, 10, 2, 32, 2, 8, maxDepth=200, seed=13, replay=oracle_replay(of))
    assert_trees_bit_exact(gf, of)


# ---- replay: regression ---------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,seed", [(2000, 0), (2000, 42), (10000, 0)])
def test_replay_mnist_regression(mnist, rows, seed):
    x, y = mnist
    x, y = x[:rows], y[:rows].astype(np.float64)
    of = O.build_forest_regression(x, y, 2, 32, 1, 1, max_depth=200, seed=seed, record_trace=True)
    gf = et.buildForestRegression(x, y, 2, 32, 1, 1, maxDepth=200, seed=seed, replay=oracle_replay(of))
    assert_trees_bit_exact(gf, of)


@pytest.mark.parametrize("nan_frac,max_depth,n_min", [(0.0, 2**31 - 1, 5), (0.1, 2**31 - 1, 2), (0.0, 4, 2)])
def test_replay_synthetic_regression(nan_frac, max_depth, n_min):
    x, y = synth_regression(6000, 15, 2, nan_frac=nan_frac)
    of = O.build_forest_regression(x, y, n_min, 4, 5, 3, max_depth=max_depth, seed=8, record_trace=True)
    gf = et.buildForestRegression(x, y, n_min, 4, 5, 3, maxDepth=max_depth, seed=8, replay=oracle_replay(of))
    assert_trees_bit_exact(gf, of)
    assert np.array_equal(et.predictRegression(gf, x), of.predict(x))
    # nodes above 2048 samples are scored from fixed-shape parallel sums (not the reference's sequential order):
    # the structure is still the reference's, and no split was decided by less than 1e-9 (relative)
    assert gf.stats["parallel_sum_nodes"] > 0
    assert gf.stats["ambiguous_splits"] == 0


# ---- free-running: the reference's own property tests --------------------------------------------------
def test_free_mnist_one_tree_fits_training_set(mnist):  # tst:283-319
    x, y = mnist
    f = et.buildForestClassification(x, y, None, 10, 2, 32, 1, 1, maxDepth=200, seed=1)
    assert (et.predictClassification(f, x).argmax(1) == y).mean() == 1.0


def test_free_mnist_regression_fits_training_set(mnist):  # tst:442-477
    x, y = mnist
    f = et.buildForestRegression(x, y.astype(np.float64), 2, 32, 1, 8, maxDepth=200, seed=2)
    assert np.array_equal(et.predictRegression(f, x).astype(np.int64), y)


def test_free_mnist_weighted(mnist):  # tst:357-395
    x, y = mnist
    w = np.concatenate([np.ones(5000), np.zeros(5000)])
    f = et.buildForestClassification(x, y, w, 10, 2, 32, 1, 8, maxDepth=200, seed=3)
    acc = (et.predictClassification(f, x).argmax(1) == y).mean()
    assert 0.5 < acc <= 0.9
    assert (et.predictClassification(f, x[:5000]).argmax(1) == y[:5000]).mean() == 1.0


def test_free_missing_regression():  # tst:515-533
    x = np.array([[1.0], [1.0], [1.0], [1.0], [NAN], [NAN], [NAN]])
    y = np.array([1.0, 1, 1, 1, 0, 0, 0])
    f = et.buildForestRegression(x, y, 1, 1, 100, 1, maxDepth=200, seed=1)
    assert np.array_equal(et.predictRegression(f, x), y)
    t = f[0]
    assert isinstance(t, et.RegressionNonLeaf) and t.splitFeature == 0 and t.cutpoint == 1.0 and t.splitMissingIsLess
    assert t.left == et.RegressionLeaf(0.0) and t.right == et.RegressionLeaf(1.0)


def test_free_missing_classification():  # tst:534-554
    x = np.array([[1.0], [1.0], [1.0], [1.0], [NAN], [NAN], [NAN]])
    y = np.array([1, 1, 1, 1, 0, 0, 0], np.int32)
    f = et.buildForestClassification(x, y, None, 2, 1, 1, 100, 1, maxDepth=200, seed=1)
    assert np.array_equal(et.predictClassification(f, x)[:, 1], y.astype(np.float64))


def test_free_accuracy_close_to_oracle(mnist):
    """Free-running tolerance: held-out accuracy of a 30-tree GPU forest within 2 points of the
    oracle's (reference algorithm, Cmwc5 stream) on the same split."""
    x, y = mnist
    xtr, ytr, xte, yte = x[:6000], y[:6000], x[6000:], y[6000:]
    of = O.build_forest_classification(xtr, ytr, None, 10, 2, 28, 30, 8, seed=5)
    gf = et.buildForestClassification(xtr, ytr, None, 10, 2, 28, 30, 8, seed=5)
    acc_o = (of.predict(xte).argmax(1) == yte).mean()
    acc_g = (et.predictClassification(gf, xte).argmax(1) == yte).mean()
    assert abs(acc_o - acc_g) < 0.02, (acc_o, acc_g)
    assert acc_g > 0.9
    # same amount of work: nodes per tree within 5%
    assert abs(gf.stats["nodes"] / of.stats()["nodes"] - 1) < 0.05


def test_free_rmse_close_to_oracle():
    x, y = synth_regression(12000, 20, 5)
    xtr, ytr, xte, yte = x[:8000], y[:8000], x[8000:], y[8000:]
    of = O.build_forest_regression(xtr, ytr, 5, 5, 30, 8, seed=6)
    gf = et.buildForestRegression(xtr, ytr, 5, 5, 30, 8, seed=6)
    rm_o = np.sqrt(np.mean((of.predict(xte) - yte) ** 2))
    rm_g = np.sqrt(np.mean((et.predictRegression(gf, xte) - yte) ** 2))
    assert abs(rm_o - rm_g) / rm_o < 0.05, (rm_o, rm_g)


def test_free_determinism_and_tree_ids(mnist):
    """A tree's stream depends only on (seed, tree id): shards rebuild the same trees."""
    x, y = mnist
    x, y = x[:1500], y[:1500]
    full = et.buildForestClassification(x, y, None, 10, 2, 16, 6, 4, seed=21)
    odd = et.buildForestClassification(x, y, None, 10, 2, 16, 3, 4, seed=21, tree_ids=[1, 3, 5])
    for j, t in enumerate([1, 3, 5]):
        a, b = full.flat(t), odd.flat(j)
        assert np.array_equal(a.feature, b.feature) and np.array_equal(a.cut.view(np.int64), b.cut.view(np.int64))
        assert np.array_equal(a.leaf, b.leaf)


# ---- argument errors mirror the reference's require(...) -------------------------------------------------
def test_require_failures():
    x = np.zeros((4, 2))
    with pytest.raises(ValueError):  # pkg:624-627
        et.buildForestClassification(x, np.zeros(3, np.int32), None, 2, 2, 1, 1, 1)
    with pytest.raises(ValueError):  # pkg:631-633
        et.buildForestClassification(x, np.zeros(4, np.int32), [1.0, -1.0, 1.0, 1.0], 2, 2, 1, 1, 1)
    with pytest.raises(ValueError):  # pkg:715-718
        et.buildForestRegression(x, np.zeros(5), 2, 1, 1, 1)


def test_packed_export_import_roundtrip(mnist):
    """The packed (device-layout) serialization: export -> import gives the same predictions, and the node
    records agree with the per-tree export."""
    x, y = mnist
    x, y = x[:2000], y[:2000]
    f = et.buildForestClassification(x, y, None, 10, 2, 28, 6, 4, seed=3)
    ser = f.export_packed()
    assert ser["tree_off"][0] == 0 and ser["tree_off"][-1] == len(ser["nodes"]) == f.total_nodes
    g = et.Forest.import_packed(ser)
    assert np.array_equal(et.predictClassification(f, x), et.predictClassification(g, x))
    for t in (0, 5):
        ft = f.flat(t)
        a, b = ser["tree_off"][t], ser["tree_off"][t + 1]
        nodes = ser["nodes"][a:b]
        split = nodes["feat"] >= 0
        assert np.array_equal(np.where(split, nodes["feat"] & 0x3FFFFFFF, -1), ft.feature)
        assert np.array_equal(nodes["cut"][split].view(np.int64), ft.cut[split].view(np.int64))
        assert np.array_equal(nodes["right_or_leaf"][split], ft.right[split])
        assert np.array_equal(ser["leaves"][nodes["right_or_leaf"][~split]], ft.leaf[~split])
    with pytest.raises(ValueError):  # ET_EINVAL surfaces like the reference's IllegalArgumentException
        bad = dict(ser)
        bad["tree_off"] = ser["tree_off"].copy()
        bad["tree_off"][-1] += 1
        et.Forest.import_packed(bad)


# ---- byte-coded table: edge cases of the dictionary coding (encode.cu) ------------------------------------
def _coding_edge_table(with_nan):
    rng = np.random.default_rng(17)
    n = 3000
    x = np.empty((n, 6))
    x[:, 0] = rng.integers(0, 256, size=n)                      # 256 distinct values: the whole byte range
    x[:, 1] = rng.choice([-np.inf, -1.5, -0.0, 0.0, 2.5, np.inf], size=n)  # infinities, both zeros
    x[:, 2] = rng.integers(0, 255, size=n) * 1e-300             # tiny magnitudes (subnormal products)
    x[:, 3] = 7.0                                               # constant
    x[:, 4] = rng.integers(0, 3, size=n)                        # three values
    x[:, 5] = rng.normal(size=n).round(1)
    if with_nan:
        x[rng.random(n) < 0.2, 2] = np.nan                      # 255 values + NaN = 256 codes
        x[rng.random(n) < 0.5, 4] = np.nan
    y = ((x[:, 0] > 100).astype(np.int32) + (np.nan_to_num(x[:, 4]) > 0)).astype(np.int32)
    return x, y


@pytest.mark.parametrize("with_nan", [False, True])
def test_replay_coding_edge_cases(with_nan):
    x, y = _coding_edge_table(with_nan)
    of = O.build_forest_classification(x, y, None, 3, 2, 4, 6, 2, seed=31, record_trace=True)
    gf = et.buildForestClassification(x, y, None, 3, 2, 4, 6, 2, seed=31, replay=oracle_replay(of))
    assert_trees_bit_exact(gf, of)
    assert np.array_equal(et.predictClassification(gf, x), of.predict(x))
    yr = y + 0.25 * x[:, 5]
    of = O.build_forest_regression(x, yr, 3, 4, 4, 2, seed=32, record_trace=True)
    gf = et.buildForestRegression(x, yr, 3, 4, 4, 2, seed=32, replay=oracle_replay(of))
    assert_trees_bit_exact(gf, of)


# ---- bestSplit = true (pkg:56-202, 298-426): every sample value tried as the cutpoint ------------------------------
@pytest.mark.parametrize("rows,k,max_depth", [(600, 28, 4), (1500, 5, 200)])
def test_replay_best_split_classification(mnist, rows, k, max_depth):
    x, y = mnist
    x, y = x[:rows], y[:rows]
    of = O.build_forest_classification(x, y, None, 10, 2, k, 2, 2, best_split=True, max_depth=max_depth, seed=6,
                                       record_trace=True)
    gf = et.buildForestClassification(x, y, None, 10, 2, k, 2, 2, bestSplit=True, maxDepth=max_depth, seed=6,
                                      replay=oracle_replay(of))
    assert gf.stats["replay_mismatches"] == 0
    assert_trees_bit_exact(gf, of)
    so, sg = of.stats(), gf.stats
    for key in ("v_mm", "v_sc", "s_rows", "p_rows", "draws", "const_hits", "scored", "nodes"):
        assert so[key] == sg[key], (key, so[key], sg[key])


def test_replay_best_split_weighted_and_missing():
    x, y = synth_classification(700, 9, 4, 21, nan_frac=0.05, const_cols=1, quantize=4)
    w = np.random.default_rng(3).uniform(0.0, 2.0, size=len(y))
    w[::7] = 0.0
    of = O.build_forest_classification(x, y, w, 4, 2, 3, 3, 2, best_split=True, seed=8, record_trace=True)
    gf = et.buildForestClassification(x, y, w, 4, 2, 3, 3, 2, bestSplit=True, seed=8, replay=oracle_replay(of))
    assert gf.stats["replay_mismatches"] == 0
    assert_trees_bit_exact(gf, of)
    assert np.array_equal(et.predictClassification(gf, x), of.predict(x))


@pytest.mark.parametrize("nan_frac", [0.0, 0.04])
def test_replay_best_split_regression(nan_frac):
    x, y = synth_regression(800, 8, 13, nan_frac=nan_frac)
    of = O.build_forest_regression(x, y, 3, 3, 2, 2, best_split=True, max_depth=12, seed=4, record_trace=True)
    gf = et.buildForestRegression(x, y, 3, 3, 2, 2, bestSplit=True, maxDepth=12, seed=4, replay=oracle_replay(of))
    assert gf.stats["replay_mismatches"] == 0
    assert_trees_bit_exact(gf, of)
    assert np.array_equal(et.predictRegression(gf, x), of.predict(x))


def test_free_best_split_mnist(mnist):  # tst:320-356 and tst:478-513
    x, y = mnist
    f = et.buildForestClassification(x, y, None, 10, 2, 32, 1, 1, bestSplit=True, maxDepth=1, seed=1)
    assert (et.predictClassification(f, x).argmax(1) == y).mean() > 0.15
    t = f[0]
    assert isinstance(t, et.ClassificationNonLeaf) and isinstance(t.left, et.ClassificationLeaf)
    fr = et.buildForestRegression(x[:3000], y[:3000].astype(np.float64), 2, 32, 1, 8, bestSplit=True, maxDepth=3, seed=2)
    assert (np.round(et.predictRegression(fr, x[:3000])).astype(np.int64) == y[:3000]).mean() > 0.15


# ---- CSC input (BASELINE.json configs[3], down-scaled): dense semantics ----------------------------------------
def _sparse_table(n, d, density, seed):
    """per column Binomial(n, density) rows, values abs(N(0,1)) + 0.1"""
    rng = np.random.default_rng(seed)
    colptr, rows, vals = [0], [], []
    for _ in range(d):
        r = np.flatnonzero(rng.random(n) < density).astype(np.int32)
        rows.append(r)
        vals.append(np.abs(rng.normal(size=len(r))) + 0.1)
        colptr.append(colptr[-1] + len(r))
    return np.array(colptr, np.int64), np.concatenate(rows), np.concatenate(vals)


@pytest.mark.parametrize("resident", ["sparse", "expanded"])
def test_csc_input_builds_the_forest_of_its_dense_expansion(resident, monkeypatch):
    """`sparse`: the table stays CSC in HBM (values found by binary search among a column's stored rows, zeros
    implicit -- what a table too large for its dense form gets); `expanded`: small tables are expanded into the
    resident dense matrix.  Either way the forest is the dense oracle's, bit for bit."""
    monkeypatch.setenv("ETGPU_CSC_DENSE_MAX", "0" if resident == "sparse" else str(1 << 40))
    n, d = 20000, 500
    colptr, rowidx, vals = _sparse_table(n, d, 0.01, 4)
    dense = np.zeros((n, d))
    for c in range(d):
        dense[rowidx[colptr[c]:colptr[c + 1]], c] = vals[colptr[c]:colptr[c + 1]]
    rng = np.random.default_rng(5)
    y = ((dense[:, :50].sum(axis=1) + 0.2 * rng.normal(size=n)) > np.median(dense[:, :50].sum(axis=1))).astype(np.int32)
    of = O.build_forest_classification(dense, y, None, 2, 2, 100, 2, 2, seed=4, record_trace=True)
    dd = et.DeviceData.from_csc(colptr, rowidx, vals, n, d)
    assert dd.shape == (n, d)
    dd.set_target_classification(y, 2)
    gf = et.buildForestClassification(dd, None, None, 2, 2, 100, 2, 2, seed=4, replay=oracle_replay(of))
    assert_trees_bit_exact(gf, of)
    assert np.array_equal(et.predictClassification(gf, dense[:3000]), of.predict(dense[:3000]))
    # free-running: the expanded CSC table and the dense table give the same forest
    f1 = et.buildForestClassification(dd, None, None, 2, 2, 100, 3, 2, seed=9)
    f2 = et.buildForestClassification(dense, y, None, 2, 2, 100, 3, 2, seed=9)
    if resident == "sparse":
        # a sparse-resident table marks the features none of a node's rows stores as constant up front instead of
        # finding them by drawing them: other draws, same distribution of scored candidates -- the trees differ but
        # do the same job (fit the training rows, same size within a few percent)
        p1, p2 = et.predictClassification(f1, dense), et.predictClassification(f2, dense)
        a1, a2 = (p1.argmax(1) == y).mean(), (p2.argmax(1) == y).mean()  # (all-zero rows with both labels exist)
        assert a1 > 0.99 and abs(a1 - a2) < 0.005, (a1, a2)
        assert abs(f1.total_nodes / f2.total_nodes - 1) < 0.1
        assert f1.stats["const_hits"] < 0.25 * f2.stats["const_hits"]  # (what the prefilter is for)
        dd.free()
        return
    a, b = f1.export_packed(), f2.export_packed()
    for field in ("feat", "right_or_leaf"):
        assert np.array_equal(a["nodes"][field], b["nodes"][field])
    assert np.array_equal(a["nodes"]["cut"].view(np.int64), b["nodes"]["cut"].view(np.int64))  # (leaves hold NaN)
    assert np.array_equal(a["leaves"], b["leaves"])
    dd.free()


def test_scipy_sparse_matrix_is_accepted():
    sp = pytest.importorskip("scipy.sparse")
    rng = np.random.default_rng(8)
    dense = np.where(rng.random((3000, 40)) < 0.05, np.round(rng.normal(size=(3000, 40)), 1), 0.0)
    y = (dense[:, :10].sum(axis=1) > 0).astype(np.int32)
    f1 = et.buildForestClassification(sp.csc_matrix(dense), y, None, 2, 2, 6, 3, 2, seed=2)
    f2 = et.buildForestClassification(dense, y, None, 2, 2, 6, 3, 2, seed=2)
    a, b = f1.export_all(), f2.export_all()
    for key in ("tree_sizes", "feature", "left", "right", "mil"):
        assert np.array_equal(a[key], b[key]), key
    assert np.array_equal(a["cut"].view(np.int64), b["cut"].view(np.int64)) and np.array_equal(a["leaf"], b["leaf"])


def test_csc_unsorted_rows_duplicates_and_regression(monkeypatch):
    """Unsorted columns are sorted on the host and a row listed twice keeps the LATER entry (deterministically), in
    both resident forms; regression and weighted classification read a sparse-resident table through the same
    loader."""
    rng = np.random.default_rng(3)
    n, d = 4000, 30
    dense = np.where(rng.random((n, d)) < 0.08, np.round(rng.normal(size=(n, d)), 2), 0.0)
    dense[rng.random((n, d)) < 0.01] = np.nan  # stored NaNs are missing values like anywhere else
    colptr, rows, vals = [0], [], []
    for c in range(d):
        r = np.flatnonzero((dense[:, c] != 0) | np.isnan(dense[:, c])).astype(np.int32)
        v = dense[r, c]
        dup = r[: min(5, len(r))]  # the first rows again, EARLIER in the list, with junk values: the later entry wins
        r2, v2 = np.concatenate([dup, r]), np.concatenate([np.full(len(dup), 99.0), v])
        perm = np.concatenate([np.arange(len(dup)), len(dup) + rng.permutation(len(r))])  # unsorted, duplicates first
        rows.append(r2[perm])
        vals.append(v2[perm])
        colptr.append(colptr[-1] + len(r2))
    colptr, rows, vals = np.array(colptr, np.int64), np.concatenate(rows), np.concatenate(vals)
    yr = np.nan_to_num(dense[:, :5]).sum(axis=1) + 0.1 * rng.normal(size=n)
    yc = (yr > np.median(yr)).astype(np.int32)
    w = rng.gamma(2.0, size=n)
    ofr = O.build_forest_regression(dense, yr, 3, 5, 3, 2, seed=6, record_trace=True)
    ofw = O.build_forest_classification(dense, yc, w, 2, 2, 5, 3, 2, seed=7, record_trace=True)
    for dense_max in ("0", str(1 << 40)):
        monkeypatch.setenv("ETGPU_CSC_DENSE_MAX", dense_max)
        dd = et.DeviceData.from_csc(colptr, rows, vals, n, d)
        gfr = et.buildForestRegression(dd, yr, 3, 5, 3, 2, seed=6, replay=oracle_replay(ofr))
        assert_trees_bit_exact(gfr, ofr)
        gfw = et.buildForestClassification(dd, yc, w, 2, 2, 5, 3, 2, seed=7, replay=oracle_replay(ofw))
        assert_trees_bit_exact(gfw, ofw)
        dd.free()


def test_csc_input_argument_errors():
    with pytest.raises(ValueError):  # row index outside the table
        et.DeviceData.from_csc([0, 1], [7], [1.0], 5, 1)
    with pytest.raises(ValueError):  # colptr decreases
        et.DeviceData.from_csc([0, 2, 1], [0, 1], [1.0, 2.0], 5, 2)
    empty = et.DeviceData.from_csc([0, 0, 0], [], [], 6, 2)  # an all-zero table is legal
    empty.set_target_classification(np.array([0, 1, 0, 1, 0, 1], np.int32), 2)
    f = et.buildForestClassification(empty, None, None, 2, 2, 2, 3, 1, seed=1)
    assert all(isinstance(t, et.ClassificationLeaf) for t in f)  # every feature is constant: roots are leaves


# ---- BASELINE.json sizes: size-independent properties --------------------------------------------------------
def test_full_size_mnist_shaped_properties(table_coding):
    """configs[1] at full size (60000 x 784, 10 classes; the bench generator): every fully grown tree (nMin=2)
    reproduces its training labels, its leaves are one-hot, it is a proper pre-order binary tree, and the forest
    vote is the mean of the per-tree votes."""
    import bench
    cfg = bench.CONFIGS["mnist"]
    x, y = bench.make_host_data(cfg)
    m = 8
    f = et.buildForestClassification(x, y, None, cfg["C"], cfg["n_min"], cfg["k"], m, 8, seed=77)
    ser = f.export_packed()
    nodes, leaves, off = ser["nodes"], ser["leaves"], ser["tree_off"]
    split = nodes["feat"] >= 0
    assert split.sum() + 1 * m == (~split).sum()              # a binary tree has one more leaf than splits
    assert len(leaves) == (~split).sum()
    # (a pure leaf of s samples holds 1/s added s times, pkg:905-911: one-hot up to the last bits)
    assert np.all((leaves == 0.0) | (np.abs(leaves - 1.0) < 1e-12)) and np.all(np.abs(leaves.sum(axis=1) - 1.0) < 1e-12)
    assert np.array_equal(np.sort(nodes["right_or_leaf"][~split]), np.arange(len(leaves)))
    for t in range(m):
        nt = nodes[off[t]:off[t + 1]]
        s = nt["feat"] >= 0
        idx = np.nonzero(s)[0]
        assert np.all(nt["right_or_leaf"][s] > idx + 1) and np.all(nt["right_or_leaf"][s] < len(nt))
        assert np.all((nt["feat"][s] & 0x3FFFFFFF) < cfg["d"])
    votes = et.predictClassification(f, x)
    assert np.array_equal(votes.argmax(1), y)                  # (no two identical rows carry different labels)
    one = et.Forest.import_packed(dict(ser, nodes=nodes[off[0]:off[1]], tree_off=off[:2] - off[0]))
    p0 = et.predictClassification(one, x[:2000])
    assert np.all(np.abs(p0.sum(axis=1) - 1.0) < 1e-12) and np.array_equal(p0.argmax(1), y[:2000])
    assert np.allclose(votes.sum(axis=1), 1.0, atol=1e-12)
    if table_coding == "codes":
        assert f.stats["launches"] > 0 and f.stats["v_mm"] >= f.stats["v_sc"] > 0


def test_full_size_regression_properties():
    """configs[2] shape at 1/4 of the rows (250000 x 100, continuous features): trees are proper, leaf means lie
    inside the target range, nodes of more than 2048 samples took the parallel-sum path, and the fit beats the
    constant predictor by a wide margin."""
    import bench
    x, y = bench.gen_regression(250_000, 100, 3)
    f = et.buildForestRegression(x, y, 5, 10, 4, 8, seed=9)
    ser = f.export_packed()
    nodes, leaves = ser["nodes"], ser["leaves"]
    split = nodes["feat"] >= 0
    assert split.sum() + 4 == (~split).sum()
    assert leaves.min() >= y.min() and leaves.max() <= y.max()
    assert f.stats["parallel_sum_nodes"] > 0
    pred = et.predictRegression(f, x[:50_000])
    assert np.mean((pred - y[:50_000]) ** 2) < 0.2 * np.var(y)


# ---- nodes cut into row chunks over several CTAs (wide.cu) -----------------------------------------------------
@pytest.mark.parametrize("chunk", [1024, 2048, 8192])
def test_wide_nodes_replay_small_chunks(mnist, chunk, monkeypatch, table_coding):
    """Small chunks force many chunks per node on a small table: replay stays bit-exact (classification: integer
    histograms merged with atomics; the stable partition places a chunk after the chunks before it)."""
    if table_coding.endswith("onecta"):
        pytest.skip("chunked path switched off in this variant")
    monkeypatch.setenv("ETGPU_WIDE_CHUNK", str(chunk))
    x, y = mnist
    of = O.build_forest_classification(x, y, None, 10, 2, 28, 2, 2, seed=17, record_trace=True)
    gf = et.buildForestClassification(x, y, None, 10, 2, 28, 2, 2, seed=17, replay=oracle_replay(of))
    assert_trees_bit_exact(gf, of)
    assert gf.stats["replay_mismatches"] == 0
    xs, ys = synth_classification(9000, 12, 3, 5, nan_frac=0.1, const_cols=2, quantize=2)
    of = O.build_forest_classification(xs, ys, None, 3, 2, 4, 3, 2, seed=2, record_trace=True)
    gf = et.buildForestClassification(xs, ys, None, 3, 2, 4, 3, 2, seed=2, replay=oracle_replay(of))
    assert_trees_bit_exact(gf, of)
    xr, yr = synth_regression(9000, 10, 4, nan_frac=0.05)
    of = O.build_forest_regression(xr, yr, 5, 3, 3, 2, seed=3, record_trace=True)
    gf = et.buildForestRegression(xr, yr, 5, 3, 3, 2, seed=3, replay=oracle_replay(of))
    assert_trees_bit_exact(gf, of)
    assert gf.stats["ambiguous_splits"] == 0


def test_wide_and_one_cta_paths_build_the_same_free_running_forest(mnist, monkeypatch):
    """The chunked path draws the same candidates as the one-CTA path, and unweighted classification scores are
    exact in both: a free-running forest does not depend on which path built a node."""
    x, y = mnist
    monkeypatch.setenv("ETGPU_WIDE_MIN", "2048")
    a = et.buildForestClassification(x, y, None, 10, 2, 28, 3, 4, seed=77)
    monkeypatch.setenv("ETGPU_WIDE_MIN", "2147483647")
    b = et.buildForestClassification(x, y, None, 10, 2, 28, 3, 4, seed=77)
    for t in range(3):
        fa, fb = a.flat(t), b.flat(t)
        assert np.array_equal(fa.feature, fb.feature) and np.array_equal(fa.cut.view(np.int64), fb.cut.view(np.int64))
        assert np.array_equal(fa.leaf, fb.leaf)
    for key in ("v_mm", "v_sc", "s_rows", "p_rows", "draws", "const_hits", "scored", "nodes"):
        assert a.stats[key] == b.stats[key], key


def _same_forest(a, b, m):
    for t in range(m):
        fa, fb = a.flat(t), b.flat(t)
        assert np.array_equal(fa.feature, fb.feature) and np.array_equal(fa.cut.view(np.int64), fb.cut.view(np.int64))
        assert np.array_equal(fa.leaf.view(np.int64), fb.leaf.view(np.int64))
    for key in ("v_mm", "v_sc", "s_rows", "p_rows", "draws", "const_hits", "scored", "nodes"):
        assert a.stats[key] == b.stats[key], key


def test_lane_classes_on_cta_teams_build_the_same_forest(mnist, monkeypatch):
    """A level with few nodes of a lane class (129..512 rows) hands them to 128-thread CTA teams (launch_level,
    ETGPU_TEAM_MAX): same draws, same scores, same tree -- free-running forests must not depend on how many nodes a
    level happens to hold (shard independence), and replay stays bit-exact on the team path."""
    x, y = mnist
    x, y = x[:6000], y[:6000]
    w = 0.25 + (np.arange(6000) % 7) / 4.0
    xr, yr = synth_regression(6000, 10, 4, nan_frac=0.05)
    xs, ys = synth_classification(5000, 12, 3, 5, nan_frac=0.1, const_cols=2, quantize=2)
    built = {}
    for team_max in ("0", "1000000000"):
        monkeypatch.setenv("ETGPU_TEAM_MAX", team_max)
        built[team_max] = (
            et.buildForestClassification(x, y, None, 10, 2, 28, 3, 4, seed=31),
            et.buildForestClassification(x, y, w, 10, 2, 28, 2, 4, seed=32),
            et.buildForestRegression(xr, yr, 5, 3, 3, 4, seed=33),
            et.buildForestClassification(xs, ys, None, 3, 2, 4, 3, 2, seed=34),
        )
    for (a, b, m) in zip(built["0"], built["1000000000"], (3, 2, 3, 3)):
        _same_forest(a, b, m)
    monkeypatch.setenv("ETGPU_TEAM_MAX", "1000000000")
    of = O.build_forest_classification(x, y, None, 10, 2, 28, 2, 2, seed=17, record_trace=True)
    gf = et.buildForestClassification(x, y, None, 10, 2, 28, 2, 2, seed=17, replay=oracle_replay(of))
    assert_trees_bit_exact(gf, of)
    assert gf.stats["replay_mismatches"] == 0
    of = O.build_forest_classification(x, y, w, 10, 2, 28, 1, 1, seed=18, record_trace=True)
    gf = et.buildForestClassification(x, y, w, 10, 2, 28, 1, 1, seed=18, replay=oracle_replay(of))
    assert_trees_bit_exact(gf, of)
    of = O.build_forest_regression(xr, yr, 5, 3, 2, 2, seed=3, record_trace=True)
    gf = et.buildForestRegression(xr, yr, 5, 3, 2, 2, seed=3, replay=oracle_replay(of))
    assert_trees_bit_exact(gf, of)


def test_small_regression_and_weighted_tables_in_a_fresh_context():
    """Tables whose every node sits in a lane class (<= 512 rows), built first thing in a new context: the CTA teams
    these classes are handed to keep their side bits in a scratch buffer that must be sized for them too (it used
    to be sized for the 513+ row classes only, and was a null pointer here)."""
    xr, yr = synth_regression(400, 6, 11, nan_frac=0.05)
    xs, ys = synth_classification(450, 8, 3, 12, nan_frac=0.05)
    w = 0.5 + (np.arange(450) % 5) / 3.0
    ofr = O.build_forest_regression(xr, yr, 3, 3, 4, 2, seed=4, record_trace=True)
    ofw = O.build_forest_classification(xs, ys, w, 3, 2, 3, 4, 2, seed=5, record_trace=True)
    for first in ("reg", "clsw"):
        ctx = et.Context(0)
        try:
            order = [("reg", ofr), ("clsw", ofw)] if first == "reg" else [("clsw", ofw), ("reg", ofr)]
            for kind, of in order:
                if kind == "reg":
                    gf = et.buildForestRegression(xr, yr, 3, 3, 4, 2, seed=4, replay=oracle_replay(of), ctx=ctx)
                else:
                    gf = et.buildForestClassification(xs, ys, w, 3, 2, 3, 4, 2, seed=5, replay=oracle_replay(of), ctx=ctx)
                assert_trees_bit_exact(gf, of)
                del gf
        finally:
            ctx.close()


# ---- argument handling (advisor findings, round 1) ---------------------------------------------------------------
def test_predict_rejects_samples_narrower_than_the_split_features(mnist):
    x, y = mnist
    x, y = x[:800], y[:800]
    f = et.buildForestClassification(x, y, None, 10, 2, 28, 2, 2, seed=1)
    used = max(int(f.flat(t).feature.max()) for t in range(2))
    with pytest.raises(ValueError):  # the reference fails with ArrayIndexOutOfBounds (pkg:517)
        et.predictClassification(f, x[:10, :used])
    assert et.predictClassification(f, x[:10, : used + 1]).shape == (10, 10)
    g = et.Forest.import_packed(f.export_packed())
    with pytest.raises(ValueError):
        et.predictClassification(g, x[:10, :used])


def test_resident_target_with_per_call_weights(mnist):
    """weights passed with a resident target are used (not silently dropped), and a call without weights keeps the
    attached ones."""
    x, y = mnist
    x, y = x[:1500], y[:1500]
    w = np.concatenate([np.ones(750), np.zeros(750)])
    ref = et.buildForestClassification(x, y, w, 10, 2, 16, 2, 4, seed=9)
    dd = et.DeviceData.from_rowmajor(x)
    dd.set_target_classification(y, 10)
    got = et.buildForestClassification(dd, None, w, 10, 2, 16, 2, 4, seed=9)
    again = et.buildForestClassification(dd, None, None, 10, 2, 16, 2, 4, seed=9)  # attached weights stay
    for t in range(2):
        for other in (got, again):
            assert np.array_equal(ref.flat(t).feature, other.flat(t).feature)
            assert np.array_equal(ref.flat(t).leaf, other.flat(t).leaf)


# ---- replay at BASELINE sizes (one tree each; the oracle runs once per module) -----------------------------------
@pytest.fixture(scope="module")
def mnist_shaped_60k():
    """BASELINE configs[1]'s table from bench.py's generator (60000 x 784, 10 classes) and one oracle tree."""
    import bench
    x, y = bench.gen_mnist_like(60000, 784, 10, 20260201)
    of = O.build_forest_classification(x, y, None, 10, 2, 28, 1, 1, seed=3, record_trace=True)
    return x, y, of


def test_replay_at_size_mnist_shaped_60000x784(mnist_shaped_60k):
    x, y, of = mnist_shaped_60k
    gf = et.buildForestClassification(x, y, None, 10, 2, 28, 1, 1, seed=3, replay=oracle_replay(of))
    assert_trees_bit_exact(gf, of)
    so, sg = of.stats(), gf.stats
    for key in ("v_mm", "v_sc", "s_rows", "p_rows", "draws", "const_hits", "scored", "nodes"):
        assert so[key] == sg[key], (key, so[key], sg[key])
    assert np.array_equal(et.predictClassification(gf, x[:5000]), of.predict(x[:5000]))


@pytest.fixture(scope="module")
def regression_1m():
    """BASELINE configs[2]'s table from bench.py's generator (1M x 100, nMin = 5, k = 10) and one oracle tree."""
    import bench
    x, y = bench.gen_regression(1_000_000, 100, 3)
    of = O.build_forest_regression(x, y, 5, 10, 1, 1, seed=4, record_trace=True)
    return x, y, of


def test_replay_at_size_regression_1Mx100(regression_1m, table_coding):
    if table_coding.startswith("codes"):
        pytest.skip("continuous table: never byte-coded (same run as the fp64 variant)")
    x, y, of = regression_1m
    gf = et.buildForestRegression(x, y, 5, 10, 1, 1, seed=4, replay=oracle_replay(of))
    # nodes of more than 2048 rows are scored from fixed-shape parallel sums: the structure is the reference's as
    # long as no such split was decided by less than 1e-9 (relative); the count is part of the result
    assert gf.stats["parallel_sum_nodes"] > 0
    assert gf.stats["ambiguous_splits"] == 0
    assert_trees_bit_exact(gf, of)
    assert np.array_equal(et.predictRegression(gf, x[:20000]), of.predict(x[:20000]))
