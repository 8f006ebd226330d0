"""CPU checks of bench.py: the workload table against BASELINE.json, the synthetic generators, and the JSON line of the
reference arm (`--impl reference` runs the oracle port on the host cores: the one bench leg that needs no GPU)."""
import json
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_workloads_follow_baseline_json():
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    cfgs = base["configs"]
    assert len(cfgs) == 5
    shapes = {"mnist": (cfgs[1], 60000, 784, 500), "reg": (cfgs[2], 1_000_000, 100, 1000),
              "sparse": (cfgs[3], 1_000_000, 10_000, 500), "large": (cfgs[4], 10_000_000, 256, 2000)}
    for key, (text, n, d, trees) in shapes.items():
        c = bench.CONFIGS[key]
        assert (c["n"], c["d"], c["trees"]) == (n, d, trees), key
        assert re.search(r"%d trees" % trees, text), (key, text)  # the tree count BASELINE.json quotes
    assert bench.CONFIGS["reg"]["task"] == "reg" and bench.CONFIGS["sparse"]["density"] == 0.01
    # the bounded instances inside the default line keep the table shape except where they say so
    assert bench.EXTRA["reg"] == dict(trees=100) and bench.EXTRA["large"] == dict(trees=32)


def test_sparse_generator_is_sorted_csc_at_the_stated_density():
    n, d = 5000, 200
    colptr, rowidx, vals, y = bench.gen_sparse(n, d, 0.01, 4)
    assert colptr[0] == 0 and colptr[-1] == len(rowidx) == len(vals) and len(colptr) == d + 1
    assert abs(len(vals) / (n * d) - 0.01) < 0.002
    for c in range(0, d, 17):
        r = rowidx[colptr[c]:colptr[c + 1]]
        assert np.all(np.diff(r) > 0) and (len(r) == 0 or (r[0] >= 0 and r[-1] < n))
    assert set(np.unique(y)) <= {0, 1} and 0.2 < y.mean() < 0.8
    dense = bench.csc_to_dense(colptr, rowidx, vals, n, d)
    assert np.count_nonzero(dense) == np.count_nonzero(vals)


def test_mnist_like_generator_statistics():
    x, y = bench.gen_mnist_like(3000, 784, 10, 20260201)
    assert x.shape == (3000, 784) and x.dtype == np.float64 and set(np.unique(y)) == set(range(10))
    assert 0.7 < (x == 0).mean() < 0.9  # the real fixture has 80.7 % zeros
    assert np.all(x == np.round(x)) and x.min() >= 0 and x.max() <= 255  # byte-codable like MNIST


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "small",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1  # ONE JSON line
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "trees built/sec" and j["unit"] == "trees/s"
    assert j["higher_is_better"] is True and j["n_gpus"] == 1 and j["steps"] == 1 and j["warmup"] == 0
    assert j["value"] > 0 and j["ms_per_step"] > 0 and j["vs_baseline"] is None
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    e = j["e2e"]
    assert e["value"] == j["value"] and e["unit"] == j["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in j["config"] and "model" not in j["config"]
