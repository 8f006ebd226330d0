"""BASELINE.json configs[0]: the reference's end-to-end extratrees check (e2e.test.scala:156-244) on its bundled
penn-ml-benchmarks tables (tests/golden/pmlb_classification.npz, made by scripts/make_pmlb_fixture.py).
CPU part: the fixture and the oracle.  GPU part: the CUDA builder against the oracle on the same tables."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.pmlb_sweep import load_tables, split
from tests.helpers import assert_trees_bit_exact, oracle_replay

SMALL = ["vehicle", "led7", "cleve", "optdigits", "dermatology", "segmentation"]


def test_fixture_holds_the_reference_tables():
    tabs = {name: (x, y) for name, x, y in load_tables()}
    assert len(tabs) == 50
    for name, (x, y) in tabs.items():  # the reference's own filter, e2e.test.scala:196-200
        n, d = x.shape
        assert 300 < n < 20000 and 5 < d < 1000 and y.min() >= 0, name
        assert np.bincount(y).max() / n < 0.6, name
    assert tabs["nursery"][0].shape == (12958, 8) and tabs["dna"][0].shape == (3186, 180)


@pytest.mark.parametrize("name", ["vehicle", "led7", "cleve"])
def test_oracle_beats_the_majority_class(name):
    (_, x, y), = load_tables({name})
    xtr, ytr, xte, yte = split(x, y)
    C, k = int(y.max()) + 1, int(np.sqrt(x.shape[1]))
    of = O.build_forest_classification(xtr, ytr, None, C, 2, k, 30, 4, seed=0)
    acc = (of.predict(xte).argmax(1) == yte).mean()
    assert acc > np.bincount(yte).max() / len(yte) + 0.05, acc


@pytest.mark.gpu
@pytest.mark.parametrize("name", SMALL)
def test_gpu_accuracy_matches_the_oracle(name):
    """free-running mode: |mean held-out accuracy (GPU) - (oracle)| <= max(0.02, 3 SE) over 5 seeds, 100 trees"""
    import lamp_b200 as et
    (_, x, y), = load_tables({name})
    xtr, ytr, xte, yte = split(x, y)
    C, k = int(y.max()) + 1, int(np.sqrt(x.shape[1]))
    ag, ao = [], []
    for seed in range(5):
        f = et.buildForestClassification(xtr, ytr, None, C, 2, k, 100, 8, seed=seed)
        ag.append((et.predictClassification(f, xte).argmax(1) == yte).mean())
        of = O.build_forest_classification(xtr, ytr, None, C, 2, k, 100, 8, seed=seed)
        ao.append((of.predict(xte).argmax(1) == yte).mean())
    se = np.sqrt((np.var(ag, ddof=1) + np.var(ao, ddof=1)) / 5)
    assert abs(np.mean(ag) - np.mean(ao)) <= max(0.02, 3 * se), (np.mean(ag), np.mean(ao), se)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["vehicle", "dna"])
def test_gpu_replay_is_bit_exact_on_pmlb(name):
    import lamp_b200 as et
    (_, x, y), = load_tables({name})
    xtr, ytr, xte, _ = split(x, y)
    C, k = int(y.max()) + 1, int(np.sqrt(x.shape[1]))
    of = O.build_forest_classification(xtr, ytr, None, C, 2, k, 6, 4, seed=3, record_trace=True)
    gf = et.buildForestClassification(xtr, ytr, None, C, 2, k, 6, 4, seed=3, replay=oracle_replay(of))
    assert_trees_bit_exact(gf, of)
    assert np.array_equal(et.predictClassification(gf, xte), of.predict(xte))
