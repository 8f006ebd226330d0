"""Converts the reference's own test fixture lamp-core/src/test/resources/mnist_test.csv.gz
(10000 x 785, header label,1x1..28x28; used by extratree.test.scala:283-513) into a compact uint8
.npz under tests/golden/ so that GPU-box tests never need /root/reference.  Run once, here."""
import gzip, sys
import numpy as np

src = "/root/reference/lamp-core/src/test/resources/mnist_test.csv.gz"
with gzip.open(src, "rt") as f:
    header = f.readline().strip().split(",")
    a = np.loadtxt(f, delimiter=",", dtype=np.float64)
assert header[0] == "label" and a.shape == (10000, 785), (header[:3], a.shape)
assert np.all(a == np.floor(a)) and a.min() >= 0 and a.max() <= 255
np.savez_compressed("tests/golden/mnist_test_u8.npz", label=a[:, 0].astype(np.uint8),
                    pixels=a[:, 1:].astype(np.uint8))
print("ok", a.shape, (a[:, 1:] == 0).mean())
