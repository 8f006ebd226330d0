"""Converts the reference's bundled penn-ml-benchmarks classification tables (datasets/penn-ml-benchmarks/
classification/*/*.tsv.gz: tab separated, header with a `target` column -- parsed by the reference at
endtoendtest/src/test/scala/lamp/endtoend/e2e.test.scala:27-44) into ONE compact .npz under tests/golden/, so
that the GPU box (which has no /root/reference) can run BASELINE.json configs[0].  Only the tables that pass
the reference's own filter (e2e.test.scala:196-200: majority class < 0.6, 300 < rows < 20000, 5 < features <
1000, targets >= 0) are kept.  Values are stored in the narrowest dtype that reproduces the FP64 table exactly.
Run once, here."""
import glob, gzip, os
import numpy as np

root = "/root/reference/datasets/penn-ml-benchmarks/classification"
out = {}
names = []
for path in sorted(glob.glob(os.path.join(root, "*", "*.tsv.gz"))):
    name = os.path.basename(path)[:-len(".tsv.gz")]
    with gzip.open(path, "rt") as f:
        header = f.readline().rstrip("\n").split("\t")
        a = np.loadtxt(f, delimiter="\t", dtype=np.float64, ndmin=2)
    ti = header.index("target")
    y = a[:, ti]
    x = np.delete(a, ti, axis=1)
    n, d = x.shape
    _, counts = np.unique(y, return_counts=True)
    if not (counts.max() / n < 0.6 and 300 < n < 20000 and 5 < d < 1000 and y.min() >= 0):
        continue
    assert np.all(y == np.floor(y))
    for dt in (np.int8, np.int16, np.float32, np.float64):
        if np.array_equal(x.astype(dt).astype(np.float64), x):
            break
    out[name + "__x"] = x.astype(dt)
    out[name + "__y"] = y.astype(np.int16)
    names.append(name)
    print("%-28s %6d x %4d  classes %3d  stored as %s" % (name, n, d, int(y.max()) + 1, np.dtype(dt).name))
out["names"] = np.array(names)
np.savez_compressed("tests/golden/pmlb_classification.npz", **out)
print(len(names), "tables,", os.path.getsize("tests/golden/pmlb_classification.npz") // 1024, "KiB")
