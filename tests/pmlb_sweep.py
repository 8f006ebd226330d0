"""BASELINE.json configs[0] / SURVEY section 8(f) rank 1: the reference's end-to-end extratrees check
(endtoendtest/src/test/scala/lamp/endtoend/e2e.test.scala:156-244) over its 50 bundled penn-ml-benchmarks
classification tables -- the GPU builder (free-running RNG) next to the CPU oracle (the reference's algorithm
and RNG), same split (test = rows 0..N/3, train = rows N/3+1.., e2e.test.scala:162-165), 100 trees,
k = floor(sqrt(d)), nMin = 2, a few seeds each.

    python tests/pmlb_sweep.py [--seeds 5] [--trees 100] [--tables a,b,c] [--out profiles/r1_pmlb_sweep.json]

Prints one line per table: held-out accuracy (mean over seeds) of both, their difference against the tolerance
max(0.01, 3 SE), and build time per forest.  The oracle is the checker here, never the product path (which is why this harness lives under tests/)."""
import argparse, json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load_tables(which=None):
    z = np.load(os.path.join(ROOT, "tests", "golden", "pmlb_classification.npz"))
    for name in z["names"]:
        name = str(name)
        if which and name not in which:
            continue
        yield name, z[name + "__x"].astype(np.float64), z[name + "__y"].astype(np.int32)


def split(x, y):
    n = len(y)
    return x[n // 3 + 1:], y[n // 3 + 1:], x[:n // 3 + 1], y[:n // 3 + 1]  # train, test (saddle slices are inclusive)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=5)
    ap.add_argument("--trees", type=int, default=100)
    ap.add_argument("--tables", default="")
    ap.add_argument("--out", default="")
    ap.add_argument("--no-oracle", action="store_true")
    a = ap.parse_args()
    import lamp_b200 as et
    from oracle import oracle as O
    threads = os.cpu_count() or 1
    rows = []
    print("%-28s %6s %4s %3s | %8s %8s %8s %6s | %9s %9s" % ("table", "rows", "d", "C", "acc_gpu", "acc_cpu", "diff", "tol",
                                                              "gpu_ms", "cpu_ms"))
    for name, x, y in load_tables(set(a.tables.split(",")) if a.tables else None):
        xtr, ytr, xte, yte = split(x, y)
        C = int(y.max()) + 1
        k = int(np.sqrt(x.shape[1]))
        acc_g, acc_o, t_g, t_o = [], [], [], []
        dd = et.DeviceData.from_rowmajor(xtr)
        dd.set_target_classification(ytr, C)
        for seed in range(a.seeds):
            t0 = time.perf_counter()
            f = et.buildForestClassification(dd, None, None, C, 2, k, a.trees, 8, seed=seed)
            p = et.predictClassification(f, xte)
            t_g.append(time.perf_counter() - t0)
            acc_g.append(float((p.argmax(1) == yte).mean()))
            if not a.no_oracle:
                t0 = time.perf_counter()
                of = O.build_forest_classification(xtr, ytr, None, C, 2, k, a.trees, threads, seed=seed)
                po = of.predict(xte)
                t_o.append(time.perf_counter() - t0)
                acc_o.append(float((po.argmax(1) == yte).mean()))
        dd.free()
        mg, mo = float(np.mean(acc_g)), float(np.mean(acc_o)) if acc_o else float("nan")
        se = float(np.sqrt((np.var(acc_g, ddof=1) + np.var(acc_o, ddof=1)) / a.seeds)) if a.seeds > 1 and acc_o else 0.0
        tol = max(0.01, 3 * se)
        rows.append(dict(table=name, rows=int(len(y)), d=int(x.shape[1]), C=C, k=k, acc_gpu=mg, acc_cpu=mo, diff=mg - mo,
                         tol=tol, gpu_ms=1e3 * float(np.median(t_g)), cpu_ms=1e3 * float(np.median(t_o)) if t_o else None))
        r = rows[-1]
        print("%-28s %6d %4d %3d | %8.4f %8.4f %+8.4f %6.4f | %9.1f %9s" % (name, r["rows"], r["d"], C, mg, mo, mg - mo, tol,
                                                                          r["gpu_ms"], "%.1f" % r["cpu_ms"] if t_o else "-"))
    if rows and not a.no_oracle:
        bad = [r["table"] for r in rows if abs(r["diff"]) > r["tol"]]
        print("tables: %d, mean acc gpu %.4f cpu %.4f, outside tolerance: %s" % (
            len(rows), np.mean([r["acc_gpu"] for r in rows]), np.mean([r["acc_cpu"] for r in rows]), bad or "none"))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(dict(seeds=a.seeds, trees=a.trees, cpu_threads=threads, tables=rows), f, indent=1)


if __name__ == "__main__":
    main()
