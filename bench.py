#!/usr/bin/env python
"""bench.py -- extratrees build / predict throughput on B200 (BASELINE.json's metric and configs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config mnist|reg|small]

A "step" is one pass of the hot path over one batch of synthetic input: building the forest of the
workload on the HBM-resident table (`value`, trees/s), and the same through the public API with host
buffers (`e2e`).  Prediction throughput (rows/s) is measured in the same run and reported under
"predict".  N>1: one process per GPU (torchrun), trees sharded by tree id, no data-path collective
during the build ("weak": every GPU builds the workload's tree count).

`--impl reference` times the reference algorithm on the host cores: the C oracle (a faithful port of
the JVM code incl. its row-major strided column walk -- no JVM exists in this image), all host
threads, on a bounded sample of the same workload.
"""
import argparse
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
    "mnist": dict(task="cls", n=60000, d=784, C=10, trees=500, k=28, n_min=2, seed=20260201,
                  name="MNIST-shaped synthetic dense classification 60000x784 f64, 10 classes, 500 trees, k=28, nMin=2"),
    # configs[2]: synthetic dense regression 1Mx100, variance criterion, 1000 trees
    "reg": dict(task="reg", n=1_000_000, d=100, C=1, trees=1000, k=10, n_min=5, seed=3,
                name="synthetic dense regression 1Mx100 f64, variance criterion, 1000 trees, k=10, nMin=5"),
    "small": dict(task="cls", n=6000, d=784, C=10, trees=50, k=28, n_min=2, seed=20260201,
                  name="MNIST-shaped synthetic 6000x784 (development size)"),
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


def make_data(cfg):
    if cfg["task"] == "cls":
        return gen_mnist_like(cfg["n"], cfg["d"], cfg["C"], cfg["seed"])
    return gen_regression(cfg["n"], cfg["d"], cfg["seed"])


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


def cpu_baseline(cfg, x, y, seconds_target=15.0):
    """Reference algorithm (oracle port) on the host cores, bounded sample: `cores` trees per round."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    trees = max(cores, 4)
    t0 = time.perf_counter()
    built = 0
    while True:
        if cfg["task"] == "cls":
            O.build_forest_classification(x, y, None, cfg["C"], cfg["n_min"], cfg["k"], trees, cores,
                                          seed=1234 + built, n_threads=cores)
        else:
            O.build_forest_regression(x, y, cfg["n_min"], cfg["k"], trees, cores, seed=1234 + built, n_threads=cores)
        built += trees
        el = time.perf_counter() - t0
        if el > seconds_target or el * 2 > seconds_target * 1.5:
            break
    return {"value": built / el, "unit": "trees/s", "cores": cores, "kind": "port",
            "sample": f"{built} trees of the same workload ({cfg['n']}x{cfg['d']}), oracle C port of the reference "
                      f"algorithm, {cores} threads, {el:.1f} s"}


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    from oracle import oracle as O
    x, y = make_data(cfg)
    cores = os.cpu_count() or 1
    trees = max(cores, 4)

    def step(seed):
        t0 = time.perf_counter()
        if cfg["task"] == "cls":
            O.build_forest_classification(x, y, None, cfg["C"], cfg["n_min"], cfg["k"], trees, cores, seed=seed,
                                          n_threads=cores)
        else:
            O.build_forest_regression(x, y, cfg["n_min"], cfg["k"], trees, cores, seed=seed, n_threads=cores)
        return time.perf_counter() - t0

    for w in range(args.warmup):
        step(100 + w)
    el = sum(step(200 + s) for s in range(args.steps))
    value = trees * args.steps / el
    sample = f"{trees} trees per step of the same workload, oracle C port (JVM unavailable), {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "trees built/sec", "value": value, "unit": "trees/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * el / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["name"], "trees_per_step": trees},
        "cpu_baseline": {"value": value, "unit": "trees/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "trees/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="mnist", choices=sorted(CONFIGS))
    ap.add_argument("--trees", type=int, default=0, help="override the workload's tree count (development)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
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

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: lamp_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = et.Context(local)
    stream = torch.cuda.Stream()  # a real (non-null) stream shared by torch's events and the library's kernels
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    x, y = make_data(cfg)  # identical on every rank (same seed): the table is replicated per GPU
    n, d, m, C = cfg["n"], cfg["d"], cfg["trees"], cfg["C"]
    tree_ids = np.arange(rank * m, (rank + 1) * m, dtype=np.int32)  # weak scaling: m trees per GPU
    x_pin = torch.from_numpy(x).pin_memory()
    xh = x_pin.numpy()

    dd = et.DeviceData.from_rowmajor(xh, ctx)
    if cfg["task"] == "cls":
        dd.set_target_classification(y, C)
    else:
        dd.set_target_regression(y)

    def build_resident(seed):
        if cfg["task"] == "cls":
            return et.buildForestClassification(dd, None, None, C, cfg["n_min"], cfg["k"], m, 8, seed=seed,
                                                tree_ids=tree_ids, ctx=ctx)
        return et.buildForestRegression(dd, None, cfg["n_min"], cfg["k"], m, 8, seed=seed, tree_ids=tree_ids, ctx=ctx)

    pinned_out = {}  # pinned host buffers the serialized forest is read into (sized after the first build)

    def build_e2e(seed):
        # the public API with HOST buffers: H2D of the table (pinned), transpose + coding, build, and the
        # device -> host read of the step's result, the serialized forest (packed device layout)
        if cfg["task"] == "cls":
            f = et.buildForestClassification(xh, y, None, C, cfg["n_min"], cfg["k"], m, 8, seed=seed,
                                             tree_ids=tree_ids, ctx=ctx)
        else:
            f = et.buildForestRegression(xh, y, cfg["n_min"], cfg["k"], m, 8, seed=seed, tree_ids=tree_ids, ctx=ctx)
        need_nodes = f.total_nodes
        if pinned_out.get("cap", 0) < need_nodes:
            cap = int(need_nodes * 1.25) + 1024
            pinned_out["cap"] = cap
            pinned_out["nodes"] = torch.empty(cap * 16, dtype=torch.uint8).pin_memory().numpy().view(et.Forest.PACKED_NODE)
            pinned_out["leaves"] = torch.empty(cap * f.leaf_width, dtype=torch.float64).pin_memory().numpy()
        ser = f.export_packed(pinned_out["nodes"], pinned_out["leaves"])
        return f, ser["nodes"].nbytes + ser["leaves"].nbytes + ser["tree_off"].nbytes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    def sum_over_ranks(v):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return float(t.item())
        return v

    # ---- value: build with inputs resident in HBM ------------------------------------------------
    forest = None
    for w in range(args.warmup):
        forest = build_resident(1000 + w)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    agg = dict(v_mm=0, s_rows=0, p_rows=0, v_sc=0, launches=0, gpu_ms_split=0.0, gpu_ms_partition=0.0, nodes=0,
               levels=0, rounds=0)
    e0.record(stream)
    for s in range(args.steps):
        forest = build_resident(2000 + s)
        for kk in agg:
            agg[kk] += forest.stats[kk]
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_build = max_over_ranks(e0.elapsed_time(e1))
    trees_total = m * world * args.steps
    value = trees_total / (ms_build / 1e3)

    # ---- e2e: public API with host buffers (H2D of the table + D2H of the forest inside) -----------
    for w in range(max(1, min(args.warmup, 2))):
        build_e2e(3000 + w)
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for s in range(args.steps):
        _, nb = build_e2e(4000 + s)
        d2h += nb
    barrier()
    el_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_value = trees_total / el_e2e
    h2d_per_step = x.nbytes + y.nbytes

    # ---- predict: batched traversal of the forest over the table's rows ----------------------------
    xt = torch.from_numpy(x).cuda()
    lw = forest.leaf_width
    out_t = torch.empty((n, lw), dtype=torch.float64, device="cuda")
    from lamp_b200 import _capi as capi
    import ctypes as CT
    pfn = capi.lib().et_predict_regression_device if cfg["task"] == "reg" else capi.lib().et_predict_classification_device

    def predict_resident():
        capi.check(pfn(ctx.h, forest.h, CT.c_void_p(xt.data_ptr()), n, d, CT.c_void_p(out_t.data_ptr()), 0))

    for w in range(args.warmup):
        predict_resident()
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    for s in range(args.steps):
        predict_resident()
    p1.record(stream)
    barrier()
    ms_pred = max_over_ranks(p0.elapsed_time(p1))
    pred_rows = n * world * args.steps / (ms_pred / 1e3)
    pfn_h = et.predictRegression if cfg["task"] == "reg" else et.predictClassification
    pfn_h(forest, xh)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        pfn_h(forest, xh)
    barrier()
    pred_e2e = n * world * args.steps / max_over_ranks(time.perf_counter() - t0)

    # ---- roofline of the dominant kernels: the node kernels (split search + stable partition fused) --------
    # achieved = algorithmic bytes of the step (8*V_mm + (4+L)*S + 16*P, counters from et_stats)
    #            / time inside the node kernels (CUDA events recorded around every node-kernel launch, et_stats)
    peak, peak_kind = measured_peak_hbm()
    L = 8 if cfg["task"] == "reg" else 4
    alg_bytes = 8 * agg["v_mm"] + (4 + L) * agg["s_rows"] + 16 * agg["p_rows"]
    k_ms = agg["gpu_ms_split"]
    achieved = alg_bytes / (k_ms / 1e3) / 1e9 if k_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per step from the last ncu --set full capture
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.config)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "node kernels k_lane (one warp per node, one lane per candidate, row-major byte-code "
                                          "gathers) + k_node CTA teams: split search + stable partition fused, all size "
                                          "classes of a level on concurrent streams",
                "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "algorithmic_bytes_per_step": alg_bytes / args.steps, "kernel_ms_per_step": k_ms / args.steps,
                "share_of_step": k_ms / (ms_build * (1 if world == 1 else 1)),
                "levels_with_cta_nodes_ms_per_step": agg["gpu_ms_partition"] / args.steps,
                "whole_build_achieved": alg_bytes / (ms_build / 1e3) / 1e9}

    launches = sum_over_ranks(agg["launches"])
    line = {
        "metric": "trees built/sec", "value": value, "unit": "trees/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_build / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["name"], "trees_per_gpu": m, "rows": n, "features": d, "k": cfg["k"],
                   "n_min": cfg["n_min"], "parallelism": f"tree-sharded x{world}",
                   "l2": "no flush: the working set of a step (sample-index / label ping-pong buffers %.0f MB + "
                         "byte-coded table 2 x %.0f MB + FP64 table %.0f MB) is larger than the 126 MB L2"
                         % (2 * 8 * m * n / 1e6, n * d / 1e6, x.nbytes / 1e6)},
        "e2e": {"value": e2e_value, "unit": "trees/s", "h2d_bytes_per_step": h2d_per_step,
                "d2h_bytes_per_step": d2h // max(args.steps, 1)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "predict": {"value": pred_rows, "unit": "rows/s", "e2e": pred_e2e, "trees": m,
                    "ms_per_step": ms_pred / args.steps},
        "stats_per_step": {kk: agg[kk] / args.steps for kk in ("nodes", "levels", "rounds", "v_mm", "v_sc", "s_rows",
                                                               "p_rows")},
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(cfg, x, y)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
