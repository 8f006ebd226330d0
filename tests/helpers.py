"""Shared helpers of the parity tests: the oracle is the checker, the C ABI is the thing checked."""
import numpy as np

from oracle import oracle as O


def oracle_replay(of):
    """Per-tree replay dicts (the et_replay test hook) from an oracle forest built with record_trace."""
    out = []
    for i in range(of.m):
        t, tr = of.tree(i), of.trace(i)
        out.append(dict(left=t.left, right=t.right, cand_begin=tr.cand_begin, cand_count=tr.cand_count,
                        cand_feature=tr.cand_feature, cand_u=tr.cand_u, cand_flag=tr.cand_flag))
    return out


def assert_trees_bit_exact(gpu_forest, oracle_forest, leaf_rtol=0.0):
    assert len(gpu_forest) == oracle_forest.m
    for i in range(oracle_forest.m):
        g, o = gpu_forest.flat(i), oracle_forest.tree(i)
        assert g.n_nodes == o.n_nodes, (i, g.n_nodes, o.n_nodes)
        assert np.array_equal(g.feature, o.feature), i
        assert np.array_equal(g.left, o.left) and np.array_equal(g.right, o.right), i
        assert np.array_equal(g.mil, o.mil), i
        split = o.feature >= 0
        assert np.array_equal(g.cut[split].view(np.int64), o.cut[split].view(np.int64)), i
        if leaf_rtol == 0.0:
            assert np.array_equal(g.leaf[~split].view(np.int64), o.leaf[~split].view(np.int64)), i
        else:
            np.testing.assert_allclose(g.leaf[~split], o.leaf[~split], rtol=leaf_rtol, atol=0)


def synth_classification(n, d, C, seed, nan_frac=0.0, const_cols=0, quantize=None):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, d))
    if quantize:
        x = np.round(x * quantize) / quantize
    score = x[:, : min(d, 4)].sum(axis=1) + 0.3 * rng.normal(size=n)
    y = np.digitize(score, np.quantile(score, np.linspace(0, 1, C + 1)[1:-1])).astype(np.int32)
    for j in range(const_cols):
        x[:, d - 1 - j] = float(j)
    if nan_frac > 0:
        x[rng.random((n, d)) < nan_frac] = np.nan
    return x, y


def synth_regression(n, d, seed, nan_frac=0.0):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, d))
    y = x[:, : min(d, 5)] @ rng.normal(size=min(d, 5)) + np.sin(3 * x[:, 0]) * x[:, 1] + 0.1 * rng.normal(size=n)
    if nan_frac > 0:
        x[rng.random((n, d)) < nan_frac] = np.nan
    return x, y
