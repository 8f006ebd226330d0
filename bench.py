#!/usr/bin/env python
"""bench.py -- extratrees build / predict throughput on B200 (BASELINE.json's metric and configs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config mnist|reg|sparse|large|small]
                    [--extra all|none|reg,sparse,large]

A "step" is one pass of the hot path over one batch of synthetic input: building the forest of the workload on the
HBM-resident table (`value`, trees/s), and the same through the public API with host buffers (`e2e`: host->device
copy of the table and device->host read of the serialized forest inside the timed region).  Prediction throughput
(rows/s) is measured in the same run and reported under "predict" with its own roofline and CPU baseline.

The headline line is BASELINE.json configs[1] (MNIST-shaped 60000x784, 500 trees).  At N = 1 the same line carries,
under "configs", bounded runs of the other BASELINE configs (reg: configs[2], sparse: configs[3], large: configs[4]),
each with build, predict, roofline and cpu_baseline (sizes stated per entry).

N > 1 (torchrun, one process per GPU): STRONG scaling -- the workload's forest is fixed and its trees are sharded by
tree id (rank r builds trees r, r+N, ...); no collective during the build; the serialized trees are all-gathered
INSIDE the library (NCCL, dist.cu) and the gather is part of both timed regions; predict is tree-sharded with an
in-library all-reduce.  The time of the collectives is itemised under "collectives".

`--impl reference` times the reference algorithm on the host cores: the C oracle (a faithful port of the JVM code
incl. its row-major strided column walk -- no JVM exists in this image), all host threads, on a bounded sample of
the same workload (a few trees per step on the full table; the line says how many).
"""
import argparse
import ctypes as CT
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[1]: MNIST-shaped synthetic dense classification, 60000x784 FP64, 10 classes, 500 trees
    "mnist": dict(task="cls", gen="mnist", n=60000, d=784, C=10, trees=500, k=28, n_min=2, seed=20260201,
                  name="MNIST-shaped synthetic dense classification 60000x784 f64, 10 classes, 500 trees, k=28, nMin=2"),
    # configs[2]: synthetic dense regression 1Mx100, variance criterion, 1000 trees
    "reg": dict(task="reg", gen="reg", n=1_000_000, d=100, C=1, trees=1000, k=10, n_min=5, seed=3,
                name="synthetic dense regression 1Mx100 f64, variance criterion, 1000 trees, k=10, nMin=5"),
    # configs[3]: synthetic sparse classification 1Mx10000 at 1% density (CSC), 500 trees
    "sparse": dict(task="cls", gen="sparse", n=1_000_000, d=10_000, C=2, trees=500, k=100, n_min=2, seed=4, density=0.01,
                   name="synthetic sparse classification 1Mx10000 at 1% density (CSC), 2 classes, 500 trees, k=100, nMin=2"),
    # configs[4]: large forest 10Mx256 dense FP64, 2000 trees (sharded over 8 GPUs), predict on 10M held-out rows
    "large": dict(task="cls", gen="large", n=10_000_000, d=256, C=2, trees=2000, k=16, n_min=2, seed=5,
                  name="large forest 10Mx256 dense f64, 2 classes, 2000 trees, k=16, nMin=2, predict on 10M held-out rows"),
    "small": dict(task="cls", gen="mnist", n=6000, d=784, C=10, trees=50, k=28, n_min=2, seed=20260201,
                  name="MNIST-shaped synthetic 6000x784 (development size)"),
}
# bounded instances of the non-headline configs run inside the default N = 1 line (what is cut is said in the entry)
EXTRA = {
    "reg": dict(trees=100),
    "sparse": dict(n=200_000, d=2_000, trees=50, k=44),
    "large": dict(trees=32),
}


def gen_mnist_like(n, d, C, seed):
    """MNIST-shaped table: C classes x 3 blob prototypes on a 28x28 grid, blended, jittered by +-2 pixels,
    integer values 0..255.  Tuned to the statistics of the reference's mnist_test fixture (80.7 % zeros, 116
    constant columns; one k=32 tree on 10k rows: 3679 nodes, depth 27, 52 % constant hits): this generator
    gives 81 % zeros, 108 always-zero border columns, ~3100 nodes, depth 27, 48 % constant hits."""
    assert d == 784
    rng = np.random.default_rng(seed)
    nsub, jit, thr = 3, 2, 0.10
    yy, xx = np.mgrid[0:28, 0:28]
    protos = np.zeros((C * nsub, 28, 28))
    for c in range(C * nsub):
        for _ in range(5):
            cy, cx = rng.uniform(6, 22, size=2)
            sy, sx = rng.uniform(1.0, 2.2, size=2)
            protos[c] += np.exp(-(((yy - cy) / sy) ** 2 + ((xx - cx) / sx) ** 2) / 2)
        protos[c] /= protos[c].max()
    y = rng.integers(0, C, size=n).astype(np.int32)
    sub = rng.integers(0, nsub, size=n)
    x = np.empty((n, 784), np.float64)
    chunk = 10000
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        a = rng.uniform(0, 1, size=(e - s, 1, 1))
        sub2 = rng.integers(0, nsub, size=e - s)
        img = a * protos[y[s:e] * nsub + sub[s:e]] + (1 - a) * protos[y[s:e] * nsub + sub2]
        dy, dx = rng.integers(-jit, jit + 1, size=(2, e - s))
        for sh in range(-jit, jit + 1):  # per-sample integer jitter
            m = dy == sh
            img[m] = np.roll(img[m], sh, axis=1)
            m = dx == sh
            img[m] = np.roll(img[m], sh, axis=2)
        img = img * rng.uniform(0.7, 1.3, size=(e - s, 1, 1)) * (1 + rng.normal(0, 0.3, size=img.shape)) \
            + rng.normal(0, 0.03, size=img.shape)
        img[:, :1, :] = 0
        img[:, -1:, :] = 0
        img[:, :, :1] = 0
        img[:, :, -1:] = 0
        v = np.clip(np.floor((img - thr) / (1 - thr) * 255.0), 0, 255)
        x[s:e] = v.reshape(e - s, 784)
    return x, y


def gen_regression(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d))
    a = rng.standard_normal(10)
    y = x[:, :10] @ a + np.sin(3 * x[:, 0]) * x[:, 1] + 0.1 * rng.standard_normal(n)
    return x, y


def gen_sparse(n, d, density, seed):
    """CSC table: per column Binomial(n, density) stored rows, values |N(0,1)| + 0.1; label = sign of the sum of 50
    informative columns + noise.  Returns (colptr int64, rowidx int32, values f64, y int32)."""
    rng = np.random.default_rng(seed)
    counts = rng.binomial(n, density, size=d).astype(np.int64)
    cols = np.repeat(np.arange(d, dtype=np.int64), counts)
    rows = rng.integers(0, n, size=len(cols), dtype=np.int64)
    key = np.unique(cols * n + rows)  # sorted by (column, row), duplicates dropped
    cols, rows = key // n, (key % n).astype(np.int32)
    colptr = np.zeros(d + 1, np.int64)
    np.add.at(colptr, cols + 1, 1)
    colptr = np.cumsum(colptr)
    vals = np.abs(rng.standard_normal(len(rows))) + 0.1
    info = rng.choice(d, size=min(50, d), replace=False)
    sgn = rng.choice([-1.0, 1.0], size=len(info))
    score = np.zeros(n)
    for c, s in zip(info, sgn):
        a, b = colptr[c], colptr[c + 1]
        score[rows[a:b]] += s * vals[a:b]
    score += 0.05 * rng.standard_normal(n)
    y = (score > np.median(score)).astype(np.int32)
    return colptr, rows, vals, y


def gen_sparse_gpu(torch, n, d, density, seed):
    """The same law as gen_sparse, drawn on the GPU (100M entries in a fraction of a second); returns host arrays."""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    nnz0 = int(n * d * density)
    key = torch.randint(0, n * d, (nnz0,), device="cuda", generator=g, dtype=torch.int64)  # column * n + row
    key = torch.unique(key)  # sorted, duplicates dropped (a Binomial(n, density) count per column in the limit)
    cols = torch.div(key, n, rounding_mode="floor")
    rows = (key - cols * n).to(torch.int32)
    colptr = torch.zeros(d + 1, dtype=torch.int64, device="cuda")
    colptr[1:] = torch.cumsum(torch.bincount(cols, minlength=d), 0)
    vals = torch.randn(len(key), device="cuda", generator=g, dtype=torch.float64).abs() + 0.1
    info = torch.randperm(d, device="cuda", generator=g)[:50]
    sgn = (torch.randint(0, 2, (50,), device="cuda", generator=g) * 2 - 1).to(torch.float64)
    w = torch.zeros(d, dtype=torch.float64, device="cuda")
    w[info] = sgn
    score = torch.zeros(n, dtype=torch.float64, device="cuda")
    score.index_add_(0, rows.to(torch.int64), vals * w[cols])
    score += 0.05 * torch.randn(n, device="cuda", generator=g, dtype=torch.float64)
    y = (score > score.median()).to(torch.int32)
    out = tuple(t.cpu().numpy() for t in (colptr, rows, vals, y))
    del key, cols, rows, vals, score, w
    torch.cuda.empty_cache()
    return out


def sparse_full_size(et, torch, ctx, trees=16):
    """BASELINE configs[3] at its full size, kept sparse in HBM (et_data_csc: 12 bytes per stored entry; the dense
    form would be 80 GB): a bounded number of trees, GPU only (the CPU port needs the dense matrix)."""
    cfg = CONFIGS["sparse"]
    n, d = cfg["n"], cfg["d"]
    colptr, rowidx, vals, y = gen_sparse_gpu(torch, n, d, cfg["density"], cfg["seed"])
    free0, _ = torch.cuda.mem_get_info()
    t0 = time.perf_counter()
    dd = et.DeviceData.from_csc(colptr, rowidx, vals, n, d, ctx)
    dd.set_target_classification(y, cfg["C"])
    t1 = time.perf_counter()
    f = et.buildForestClassification(dd, None, None, cfg["C"], cfg["n_min"], cfg["k"], trees, 8, seed=1, ctx=ctx)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    free1, _ = torch.cuda.mem_get_info()
    ent = {"workload": cfg["name"], "rows": n, "features": d, "stored_entries": int(len(vals)), "trees": trees,
           "k": cfg["k"], "resident": "CSC (sparse): one-team kernels search a column's stored rows (lockstep searches), chunks of a large node walk the stored entries of their row range (membership through an inverse index map); implicit zeros by count",
           "table_bytes_hbm": int(len(vals)) * 12 + (d + 1) * 8, "dense_bytes": n * d * 8,
           "hbm_in_use_after_build_bytes": int(free0 - free1),
           "upload_s": t1 - t0, "build": {"value": trees / (t2 - t1), "unit": "trees/s", "ms_per_step": 1e3 * (t2 - t1),
                                          "steps": 1, "warmup": 0, "note": "first build of the context (allocations inside)"},
           "stats_per_step": {k: f.stats[k] for k in ("nodes", "levels", "v_mm", "draws", "const_hits")}}
    f.free()
    dd.free()
    torch.cuda.empty_cache()
    return ent


def csc_to_dense(colptr, rowidx, vals, n, d):
    x = np.zeros((n, d))
    cols = np.repeat(np.arange(d), np.diff(colptr))
    x[rowidx, cols] = vals
    return x


def large_labels_np(x):
    """3-level planted tree on 8 features; the 5 % label flips are applied by the caller."""
    return np.where(x[:, 0] > 0, np.where(x[:, 1] > 0.3, x[:, 2] > -0.2, x[:, 3] > 0.1),
                    np.where(x[:, 4] > -0.4, x[:, 5] > 0.2, x[:, 6] > 0.0))


def make_host_data(cfg):
    """(x row-major f64, y) on the host for the workloads a host can hold."""
    if cfg["gen"] == "mnist":
        return gen_mnist_like(cfg["n"], cfg["d"], cfg["C"], cfg["seed"])
    if cfg["gen"] == "reg":
        return gen_regression(cfg["n"], cfg["d"], cfg["seed"])
    if cfg["gen"] == "large":
        rng = np.random.default_rng(cfg["seed"])
        x = rng.standard_normal((cfg["n"], cfg["d"]), dtype=np.float32).astype(np.float64)
        y = large_labels_np(x) ^ (rng.random(cfg["n"]) < 0.05)
        return x, y.astype(np.int32)
    raise ValueError(cfg["gen"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(device)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [s.strip() for s in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out



def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def oracle_build(cfg, x, y, trees, cores, seed):
    from oracle import oracle as O
    if cfg["task"] == "cls":
        return O.build_forest_classification(x, y, None, cfg["C"], cfg["n_min"], cfg["k"], trees, cores, seed=seed,
                                             n_threads=cores)
    return O.build_forest_regression(x, y, cfg["n_min"], cfg["k"], trees, cores, seed=seed, n_threads=cores)


def cpu_build_baseline(cfg, x, y, seconds_target=12.0, note=""):
    """Reference algorithm (oracle port) on the host cores, bounded sample: `cores` trees per round."""
    cores = os.cpu_count() or 1
    trees = max(cores, 4)
    t0 = time.perf_counter()
    built = 0
    while True:
        oracle_build(cfg, x, y, trees, cores, 1234 + built)
        built += trees
        el = time.perf_counter() - t0
        if el > seconds_target or el * 2 > seconds_target * 1.5:
            break
    return {"value": built / el, "unit": "trees/s", "cores": cores, "kind": "port",
            "sample": f"{built} trees on a {x.shape[0]}x{x.shape[1]} table{note}, oracle C port of the reference "
                      f"algorithm (JVM unavailable), {cores} threads, {el:.1f} s"}


def cpu_predict_baseline(forest, x_rows, regression, max_trees=8, seconds_target=4.0):
    """predictClassification / predictRegression of the reference (single thread, row-outer / tree-inner, pkg:546-551)
    through the oracle port, on a bounded sample: the first trees of the GPU-built forest, a slice of the rows."""
    from oracle import oracle as O
    nt = min(max_trees, len(forest))
    of = O.import_forest([forest.flat(t) for t in range(nt)], regression)
    rows = min(len(x_rows), 2000)
    t0 = time.perf_counter()
    of.predict(np.ascontiguousarray(x_rows[:rows]))
    el = time.perf_counter() - t0
    rows = int(min(len(x_rows), max(rows, rows * seconds_target / max(el, 1e-6))))
    t0 = time.perf_counter()
    of.predict(np.ascontiguousarray(x_rows[:rows]))
    el = time.perf_counter() - t0
    v = rows / el
    return {"value": v, "unit": "rows/s", "trees": nt, "cores": 1, "kind": "port",
            "value_scaled_to_forest": v * nt / max(len(forest), 1),
            "sample": f"{rows} rows through the first {nt} of {len(forest)} trees, oracle C port (the reference predicts on "
                      f"one thread), {el:.1f} s; value_scaled_to_forest = value x {nt}/{len(forest)} (cost is linear in trees)"}


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    x, y = make_host_data(cfg)
    cores = os.cpu_count() or 1
    trees = max(cores, 4)

    def step(seed):
        t0 = time.perf_counter()
        oracle_build(cfg, x, y, trees, cores, seed)
        return time.perf_counter() - t0

    for w in range(args.warmup):
        step(100 + w)
    el = sum(step(200 + s) for s in range(args.steps))
    value = trees * args.steps / el
    sample = (f"{trees} trees per step (not the workload's {cfg['trees']}: throughput is per tree) on the full "
              f"{cfg['n']}x{cfg['d']} table, oracle C port (JVM unavailable), {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": "trees built/sec", "value": value, "unit": "trees/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * el / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["name"], "trees_per_step": trees,
                   "note": f"the reference arm builds {trees} trees per step on the same table; trees/s is per tree"},
        "cpu_baseline": {"value": value, "unit": "trees/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "trees/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


class Bench:
    """One workload on one rank: resident table, build / e2e / predict timings."""

    def __init__(self, cfg, args, et, torch, ctx, stream, rank, world, dist_mod=None):
        self.cfg, self.args, self.et, self.torch, self.ctx, self.stream = cfg, args, et, torch, ctx, stream
        self.rank, self.world, self.D = rank, world, dist_mod
        self.m = cfg["trees"]
        from lamp_b200 import dist as D
        self.ids = D.shard_tree_ids(self.m, rank, world)  # strong scaling: the forest is fixed, its trees are sharded
        self.x_host = self.y = self.csc = self.x_dev = self.x_pred_dev = self.x_pageable = None
        self.pinned = {}
        self.gathered = world > 1
        self._make_data()

    # ---- data -------------------------------------------------------------------------------------------------
    def _make_data(self):
        cfg, et, torch = self.cfg, self.et, self.torch
        n, d = cfg["n"], cfg["d"]
        if cfg["gen"] == "sparse":
            colptr, rowidx, vals, y = gen_sparse(n, d, cfg["density"], cfg["seed"])
            self.csc, self.y = (colptr, rowidx, vals), y
            self.dd = et.DeviceData.from_csc(colptr, rowidx, vals, n, d, self.ctx)
            rows = min(n, 20000)  # rows predicted (dense rows: the reference's predict takes a dense Mat)
            self.x_pred_host = csc_to_dense(*self._csc_rows(rows), rows, d)
            self.x_pred_dev = torch.from_numpy(self.x_pred_host).cuda()
        elif cfg["gen"] == "large":
            g = torch.Generator(device="cuda")
            g.manual_seed(cfg["seed"])
            xd = torch.randn((n, d), dtype=torch.float64, device="cuda", generator=g)
            yt = torch.where(xd[:, 0] > 0, torch.where(xd[:, 1] > 0.3, xd[:, 2] > -0.2, xd[:, 3] > 0.1),
                             torch.where(xd[:, 4] > -0.4, xd[:, 5] > 0.2, xd[:, 6] > 0.0))
            yt = yt ^ (torch.rand(n, device="cuda", generator=g) < 0.05)
            self.y = yt.to(torch.int32).cpu().numpy()
            torch.cuda.synchronize()
            self.dd = et.DeviceData.from_device_rowmajor(xd.data_ptr(), n, d, self.ctx)
            del xd, yt
            torch.cuda.empty_cache()
            self.x_pred_dev = torch.randn((n, d), dtype=torch.float64, device="cuda", generator=g)  # held-out draw
            self.x_pred_host = None
        else:
            x, self.y = make_host_data(cfg)
            self.x_pin = torch.from_numpy(x).pin_memory()
            self.x_host = self.x_pin.numpy()
            self.dd = et.DeviceData.from_rowmajor(self.x_host, self.ctx)
            self.x_pred_host = self.x_host
            self.x_pred_dev = torch.from_numpy(self.x_host).cuda()
        if cfg["task"] == "cls":
            self.dd.set_target_classification(self.y, cfg["C"])
        else:
            self.dd.set_target_regression(self.y)

    def _csc_rows(self, rows):
        colptr, rowidx, vals = self.csc
        keep = rowidx < rows
        cols = np.repeat(np.arange(self.cfg["d"]), np.diff(colptr))[keep]
        cp = np.zeros(self.cfg["d"] + 1, np.int64)
        np.add.at(cp, cols + 1, 1)
        return np.cumsum(cp), rowidx[keep], vals[keep]

    # ---- builds -----------------------------------------------------------------------------------------------
    def build(self, data, target, seed, ids=None):
        cfg, et = self.cfg, self.et
        ids = self.ids if ids is None else ids
        if cfg["task"] == "cls":
            return et.buildForestClassification(data, target, None, cfg["C"], cfg["n_min"], cfg["k"], len(ids), 8,
                                                seed=seed, tree_ids=ids, ctx=self.ctx)
        return et.buildForestRegression(data, target, cfg["n_min"], cfg["k"], len(ids), 8, seed=seed, tree_ids=ids,
                                        ctx=self.ctx)

    def build_resident(self, seed):
        """Table resident in HBM; N > 1: + the in-library all-gather of the serialized trees."""
        f = self.build(self.dd, None, seed)
        if self.world > 1:
            # (a forest whose gathered form would not fit next to the tables stays sharded: configs[4] at its full
            # 2000 trees is ~180 GB of nodes; predict is tree-sharded either way)
            self.gathered = f.total_nodes * self.world * 24 < 40e9 and self.args.gather != "off"
            if self.gathered:
                full = self.D.gather_forest(self.ctx, f)
                return f, full, self.ctx.comm_last_ms()
        return f, f, 0.0

    def export_pinned(self, f):
        et, torch = self.et, self.torch
        if self.pinned.get("cap", 0) < f.total_nodes:
            cap = int(f.total_nodes * 1.25) + 1024
            self.pinned = {"cap": cap,
                           "nodes": torch.empty(cap * 16, dtype=torch.uint8).pin_memory().numpy().view(et.Forest.PACKED_NODE),
                           "leaves": torch.empty(cap * f.leaf_width, dtype=torch.float64).pin_memory().numpy()}
        ser = f.export_packed(self.pinned["nodes"], self.pinned["leaves"])
        return ser["nodes"].nbytes + ser["leaves"].nbytes + ser["tree_off"].nbytes

    def build_e2e(self, seed, pageable=False):
        """The public API with HOST buffers: H2D of the table (pinned), transpose + coding, build, [gather], and the
        device -> host read of the step's result, the serialized forest (packed device layout).  N > 1: rank 0
        uploads, the replicas travel over NVLink (et_data_broadcast), every rank builds its shard, the trees are
        all-gathered and rank 0 reads the forest."""
        cfg, et = self.cfg, self.et
        if self.world == 1:
            if self.csc is not None:
                dd = et.DeviceData.from_csc(*self.csc, cfg["n"], cfg["d"], self.ctx)
                f = self.build(dd, self.y, seed)
                dd.free()
            else:
                if pageable and self.x_pageable is None:
                    self.x_pageable = np.array(self.x_host)  # an ordinary (pageable) host array, like a JVM heap array
                f = self.build(self.x_pageable if pageable else self.x_host, self.y, seed)
            return self.export_pinned(f)
        dd0 = et.DeviceData.from_rowmajor(self.x_host, self.ctx) if self.rank == 0 else None
        dd = self.D.broadcast_data(self.ctx, dd0, 0)
        f = self.build(dd, self.y, seed)
        full = self.D.gather_forest(self.ctx, f)
        nb = self.export_pinned(full) if self.rank == 0 else 0
        dd.free()
        return nb

    def h2d_bytes(self):
        if self.csc is not None:
            return sum(a.nbytes for a in self.csc) + self.y.nbytes
        return (self.x_host.nbytes if self.x_host is not None else 0) + self.y.nbytes

    # ---- predict ----------------------------------------------------------------------------------------------
    def predict_resident(self, shard, out_t):
        from lamp_b200 import _capi as capi
        n, d = self.x_pred_dev.shape
        if self.world > 1:
            self.D.predict_sharded_device(self.ctx, shard, self.x_pred_dev.data_ptr(), n, d, out_t.data_ptr(), self.m)
            return self.ctx.comm_last_ms()
        fn = capi.lib().et_predict_regression_device if self.cfg["task"] == "reg" else capi.lib().et_predict_classification_device
        capi.check(fn(self.ctx.h, shard.h, CT.c_void_p(self.x_pred_dev.data_ptr()), n, d, CT.c_void_p(out_t.data_ptr()), 0))
        return 0.0


def measure(b, steps, warmup, want_e2e=True, want_cpu=True, clocks_device=None, cpu_table=None):
    """Runs one workload: returns the pieces of the JSON line."""
    torch, cfg, args = b.torch, b.cfg, b.args
    world, rank = b.world, b.rank
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(v, op):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=op)
            return float(t.item())
        return v

    # ---- value: build with inputs resident in HBM -----------------------------------------------------------
    shard = full = None
    for w in range(warmup):
        shard, full, _ = b.build_resident(1000 + w)
    barrier()
    sampler = ClockSampler(clocks_device) if (clocks_device is not None and rank == 0) else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    keys = ("v_mm", "s_rows", "p_rows", "v_sc", "launches", "gpu_ms_split", "gpu_ms_partition", "nodes", "levels",
            "rounds", "parallel_sum_nodes", "ambiguous_splits", "draws", "const_hits")
    agg = {k: 0 for k in keys}
    gather_ms = 0.0
    e0.record(b.stream)
    for s in range(steps):
        shard, full, gms = b.build_resident(2000 + s)
        gather_ms += gms
        for kk in agg:
            agg[kk] += shard.stats[kk]
    e1.record(b.stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_build = reduce(e0.elapsed_time(e1), dist.ReduceOp.MAX if world > 1 else None)
    value = b.m * steps / (ms_build / 1e3)
    for kk in agg:  # whole-job counters: summed over the ranks (every rank built its shard)
        agg[kk] = reduce(float(agg[kk]), dist.ReduceOp.SUM if world > 1 else None)
    gather_ms = reduce(gather_ms, dist.ReduceOp.MAX if world > 1 else None)

    # ---- e2e -----------------------------------------------------------------------------------------------
    e2e = None
    if want_e2e:
        for w in range(1 if steps <= 2 else 2):
            b.build_e2e(3000 + w)
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        for s in range(steps):
            d2h += b.build_e2e(4000 + s)
        barrier()
        el = reduce(time.perf_counter() - t0, dist.ReduceOp.MAX if world > 1 else None)
        d2h = reduce(float(d2h), dist.ReduceOp.SUM if world > 1 else None)
        e2e = {"value": b.m * steps / el, "unit": "trees/s", "h2d_bytes_per_step": b.h2d_bytes(),
               "d2h_bytes_per_step": int(d2h) // max(steps, 1), "host_input": "pinned"}
        if world == 1 and b.x_host is not None and cfg.get("_main"):
            # the same from an ordinary pageable array (what a JVM / numpy caller holds): the library stages it
            # through its own pinned bounce buffers (api.cu et_h2d)
            b.build_e2e(5000, pageable=True)
            t0 = time.perf_counter()
            for s in range(steps):
                b.build_e2e(6000 + s, pageable=True)
            torch.cuda.synchronize()
            e2e["pageable_input_value"] = b.m * steps / (time.perf_counter() - t0)

    # ---- predict ---------------------------------------------------------------------------------------------
    n_pred, d_pred = b.x_pred_dev.shape
    lw = shard.leaf_width
    out_t = torch.empty((n_pred, lw), dtype=torch.float64, device="cuda")
    for w in range(max(1, min(warmup, 2))):
        b.predict_resident(shard, out_t)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ar_ms = 0.0
    p0.record(b.stream)
    for s in range(steps):
        ar_ms += b.predict_resident(shard, out_t)
    p1.record(b.stream)
    barrier()
    ms_pred = reduce(p0.elapsed_time(p1), dist.ReduceOp.MAX if world > 1 else None)
    pred_rows = n_pred * steps / (ms_pred / 1e3)
    ar_ms = reduce(ar_ms, dist.ReduceOp.MAX if world > 1 else None)
    pred_e2e = None
    if b.x_pred_host is not None and world == 1:
        fnh = b.et.predictRegression if cfg["task"] == "reg" else b.et.predictClassification
        rows_h = min(len(b.x_pred_host), 200_000)
        fnh(full, b.x_pred_host[:rows_h], ctx=b.ctx)
        t0 = time.perf_counter()
        for s in range(steps):
            fnh(full, b.x_pred_host[:rows_h], ctx=b.ctx)
        pred_e2e = rows_h * steps / (time.perf_counter() - t0)

    # ---- rooflines -------------------------------------------------------------------------------------------
    # build: achieved = algorithmic bytes of the step (8*V_mm + (4+L)*S + 16*P, counters from et_stats)
    #        / time inside the node kernels (CUDA events around every level's node-kernel launches, et_stats)
    peak, peak_kind = measured_peak_hbm()
    L = 8 if cfg["task"] == "reg" else 4
    alg_bytes = 8 * agg["v_mm"] + (4 + L) * agg["s_rows"] + 16 * agg["p_rows"]
    k_ms = agg["gpu_ms_split"] / world  # per rank (ranks run side by side)
    achieved = alg_bytes / world / (k_ms / 1e3) / 1e9 if k_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per step from the last ncu --set full capture
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.config if cfg is b.cfg and cfg.get("_main") else cfg.get("_key"))
        except Exception:
            traffic = None
    roofline = {"bound": "hbm",
                "kernel": "node kernels: k_lane (warp per node, lane per candidate) + k_node (CTA per node) + k_wide_* "
                          "(chunks of a large node over several CTAs): split search + stable partition, all size classes "
                          "of a level on concurrent streams",
                "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "per_gpu": True,
                "algorithmic_bytes_per_step": alg_bytes / steps, "kernel_ms_per_step": k_ms / steps,
                "share_of_step": k_ms / ms_build,
                "levels_with_cta_nodes_ms_per_step": agg["gpu_ms_partition"] / world / steps,
                "whole_build_achieved": alg_bytes / world / (ms_build / 1e3) / 1e9}
    # predict: algorithmic bytes = the rows read once + the outputs written once + the forest read once (SURVEY 8d:
    # 8 n d + 8 n C + 24 nodes; this library's node is 16 bytes + a leaf table: 16 nodes + 8 C leaves); the
    # traversal itself (n x m x depth node visits) is served from L2 / shared memory
    nodes_f = full.total_nodes
    pred_bytes = 8.0 * n_pred * d_pred + 8.0 * n_pred * lw + 16.0 * nodes_f + 8.0 * lw * (nodes_f + len(full)) / 2
    pred_ach = pred_bytes / (ms_pred / steps / 1e3) / 1e9
    predict = {"value": pred_rows, "unit": "rows/s", "e2e": pred_e2e, "trees": b.m, "rows": n_pred,
               "ms_per_step": ms_pred / steps, "row_trees_per_s": pred_rows * b.m,
               "roofline": {"bound": "hbm", "kernel": "k_predict (batched traversal of the flattened pre-order nodes)",
                            "achieved": pred_ach, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                            "frac": pred_ach / peak, "traffic": None, "algorithmic_bytes_per_step": pred_bytes,
                            "note": "pointer chasing through ~depth dependent L2 reads per (row, tree): latency bound, "
                                    "not HBM bound"}}
    out = {"value": value, "ms_build": ms_build, "e2e": e2e, "clocks": clocks, "roofline": roofline, "predict": predict,
           "agg": agg, "gather_ms": gather_ms, "allreduce_ms": ar_ms, "full": full, "shard": shard}
    if want_cpu and rank == 0 and world == 1:
        xs, ys, note = cpu_table if cpu_table is not None else (b.x_host, b.y, "")
        if xs is not None:
            out["cpu_baseline"] = cpu_build_baseline(cfg, xs, ys, note=note)
        if b.x_pred_host is not None:
            predict["cpu_baseline"] = cpu_predict_baseline(full, b.x_pred_host, cfg["task"] == "reg")
    return out


def stats_block(r, steps):
    a = r["agg"]
    return {kk: a[kk] / steps for kk in ("nodes", "levels", "rounds", "v_mm", "v_sc", "s_rows", "p_rows", "draws",
                                         "const_hits", "parallel_sum_nodes", "ambiguous_splits")}


def run_extra(key, args, et, torch, ctx, stream):
    """A bounded instance of a non-headline BASELINE config (N = 1 only)."""
    cfg = dict(CONFIGS[key])
    full_cfg = dict(cfg)
    cfg.update(EXTRA[key])
    cfg["_key"] = key
    cut = {k: (full_cfg[k], cfg[k]) for k in EXTRA[key] if full_cfg.get(k) != cfg[k]}
    t0 = time.perf_counter()
    b = Bench(cfg, args, et, torch, ctx, stream, 0, 1)
    cpu_table = None
    if key == "large":  # the 10M-row table never leaves the device: the CPU port gets a 1M-row table of the same law
        c2 = dict(cfg, n=1_000_000)
        cpu_table = make_host_data(c2) + (" of the same distribution (the %d-row table is generated on the device)" % cfg["n"],)
        b.x_pred_host = cpu_table[0][:20000]
    if key == "sparse":
        # the CPU port needs the dense form; one tree of the 200000-row table takes a host thread a minute, so the
        # sample is the first 50000 rows (the line says so)
        nc = min(cfg["n"], 50_000)
        cpu_table = (csc_to_dense(*b.csc, cfg["n"], cfg["d"])[:nc].copy(), np.ascontiguousarray(b.y[:nc]),
                     " (the first %d rows of the dense expansion of the %d-row CSC table)" % (nc, cfg["n"])) \
            if cfg["n"] * cfg["d"] <= 1_000_000_000 else None
    steps = 2 if key == "reg" else 1
    r = measure(b, steps, 1, want_e2e=(key != "large"), want_cpu=True, cpu_table=cpu_table)
    ent = {"workload": full_cfg["name"],
           "bounded": {k: {"config": v[0], "run": v[1]} for k, v in cut.items()} or None,
           "rows": cfg["n"], "features": cfg["d"], "trees": cfg["trees"], "k": cfg["k"], "n_min": cfg["n_min"],
           "build": {"value": r["value"], "unit": "trees/s", "ms_per_step": r["ms_build"] / steps, "steps": steps, "warmup": 1},
           "e2e": r["e2e"] if r["e2e"] else {"value": None, "note": "table generated on the device (20.5 GB): no host copy to upload"},
           "roofline": r["roofline"], "predict": r["predict"], "cpu_baseline": r.get("cpu_baseline"),
           "stats_per_step": stats_block(r, steps), "wall_s": None}
    for o in (r["full"], r["shard"]):
        o.free()
    b.dd.free()
    del b
    torch.cuda.empty_cache()
    if key == "sparse":
        try:
            ent["full_size"] = sparse_full_size(et, torch, ctx)
        except Exception as e:
            ent["full_size"] = {"error": repr(e)}
    ent["wall_s"] = time.perf_counter() - t0
    return ent


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="mnist", choices=sorted(CONFIGS))
    ap.add_argument("--trees", type=int, default=0, help="override the workload's tree count (development)")
    ap.add_argument("--extra", default="all", help="bounded runs of the other BASELINE configs in the N=1 line: "
                                                   "all | none | comma list of reg,sparse,large")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="auto", choices=["auto", "off"],
                    help="N > 1: all-gather the serialized trees inside the timed region (auto: unless the gathered "
                         "forest would not fit in HBM)")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    cfg["_main"] = True
    cfg["_key"] = args.config
    if args.trees:
        cfg["trees"] = args.trees
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    import torch
    import torch.distributed as dist

    import lamp_b200 as et
    from lamp_b200 import dist as D

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: lamp_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = et.Context(local)
    stream = torch.cuda.Stream()  # a real (non-null) stream shared by torch's events and the library's kernels
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        D.init_comm(ctx, rank, world)  # NCCL communicator INSIDE the library (gather of trees, predict all-reduce)

    b = Bench(cfg, args, et, torch, ctx, stream, rank, world, D)
    has_host_table = b.x_host is not None or b.csc is not None
    r = measure(b, args.steps, args.warmup, want_e2e=has_host_table, want_cpu=not args.no_cpu_baseline, clocks_device=local)
    n, d, m = cfg["n"], cfg["d"], cfg["trees"]
    launches = r["agg"]["launches"]
    line = {
        "metric": "trees built/sec", "value": r["value"], "unit": "trees/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_build"] / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["name"], "trees": m, "trees_per_gpu": len(b.ids), "rows": n, "features": d,
                   "k": cfg["k"], "n_min": cfg["n_min"],
                   "parallelism": f"forest of {m} trees sharded by tree id over {world} GPU(s); table replicated; "
                                  "no collective during the build" + ("; in-library NCCL all-gather of the serialized "
                                                                      "trees inside the timed region" if world > 1 else ""),
                   "l2": "no flush: the working set of a step (sample-index / label ping-pong buffers %.0f MB + "
                         "byte-coded table 2 x %.0f MB + FP64 table %.0f MB) is larger than the 126 MB L2"
                         % (2 * 8 * len(b.ids) * n / 1e6, n * d / 1e6, 8.0 * n * d / 1e6)},
        "e2e": r["e2e"] if r["e2e"] else {"value": None, "unit": "trees/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                                          "note": "table generated on the device: no host copy to upload"},
        "gpu_launches": int(launches),
        "clocks": r["clocks"],
        "roofline": r["roofline"],
        "predict": r["predict"],
        "stats_per_step": stats_block(r, args.steps),
    }
    if world > 1:
        line["collectives"] = {"forest_gathered": bool(b.gathered),
                               "forest_allgather_ms_per_step": r["gather_ms"] / args.steps,
                               "predict_allreduce_stream_ms_per_step": r["allreduce_ms"] / args.steps,
                               "note": "device time, max over ranks; the all-reduce runs in row chunks on its own stream "
                                       "overlapped with the traversal (the figure is that stream's span)"}
    if "cpu_baseline" in r:
        line["cpu_baseline"] = r["cpu_baseline"]
    if rank == 0 and world == 1 and args.extra != "none" and args.config == "mnist":
        for o in (r["full"],):
            o.free()
        b.dd.free()
        del b
        torch.cuda.empty_cache()
        want = ["reg", "sparse", "large"] if args.extra == "all" else [s for s in args.extra.split(",") if s]
        line["configs"] = {}
        for key in want:
            try:
                line["configs"][key] = run_extra(key, args, et, torch, ctx, stream)
            except Exception as e:  # a failing extra must not take the headline line with it
                line["configs"][key] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
