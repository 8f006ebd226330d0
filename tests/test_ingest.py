"""CSV ingest (lamp_b200/ingest.py): the host side of SURVEY 8(f) row 4 -- parse by column, the Frame operations the
reference's tests use (tst:284-294), missing fields = NaN."""
import gzip
import io

import numpy as np
import pytest

from lamp_b200 import ingest


def _write_csv(path, names, rows, gz):
    text = ",".join(names) + "\n" + "\n".join(",".join(r) for r in rows) + "\n"
    if gz:
        with gzip.open(path, "wt") as f:
            f.write(text)
    else:
        with open(path, "w") as f:
            f.write(text)


@pytest.mark.parametrize("gz", [False, True])
def test_read_csv_columns_labels_and_missing(tmp_path, gz):
    rng = np.random.default_rng(1)
    n, d = 200, 7
    x = np.round(rng.normal(size=(n, d)), 3)
    y = rng.integers(0, 3, size=n)
    x[rng.random((n, d)) < 0.05] = np.nan
    names = ["label"] + ["p%d" % j for j in range(d)]
    rows = [[str(int(y[i]))] + ["" if np.isnan(v) else repr(float(v)) for v in x[i]] for i in range(n)]
    rows[3][2] = "NA"  # not a number -> missing, like saddle's Double parser
    x[3, 1] = np.nan
    p = str(tmp_path / ("t.csv.gz" if gz else "t.csv"))
    _write_csv(p, names, rows, gz)
    fr = ingest.read_csv(p)
    assert fr.names == names and fr.n_rows == n and fr.columns.shape == (d + 1, n)
    assert np.array_equal(fr.first_col("label"), y.astype(np.float64))
    feats = fr.filter_ix(lambda nm: nm != "label")
    assert feats.names == names[1:]
    m = feats.to_mat()
    assert m.shape == (n, d) and m.flags["C_CONTIGUOUS"]
    assert np.array_equal(np.isnan(m), np.isnan(x)) and np.array_equal(np.nan_to_num(m), np.nan_to_num(x))


def test_read_csv_from_a_stream_without_header():
    fr = ingest.read_csv(io.StringIO("1,2.5,\n4,,6\n"), header=False)
    assert fr.columns.shape == (3, 2)
    assert np.array_equal(np.nan_to_num(fr.columns, nan=-1.0), np.array([[1, 4], [2.5, -1], [-1, 6]], dtype=np.float64))


def test_column_blocks_cover_the_table_in_order():
    cols = np.arange(10 * 4, dtype=np.float64).reshape(10, 4)
    got = list(ingest.column_blocks(cols, block_cols=3))
    assert [f for f, _ in got] == [0, 3, 6, 9] and [b.shape[0] for _, b in got] == [3, 3, 3, 1]
    assert np.array_equal(np.concatenate([b for _, b in got]), cols)


def test_mnist_fixture_round_trips_through_csv(tmp_path):
    """The reference's own fixture shape (label + 784 pixel columns): written as CSV, read back by column."""
    z = np.load("tests/golden/mnist_test_u8.npz")
    keys = list(z.keys())
    x = z[[k for k in keys if z[k].ndim == 2][0]][:50].astype(np.float64)
    y = z[[k for k in keys if z[k].ndim == 1][0]][:50]
    names = ["label"] + ["pixel%d" % j for j in range(x.shape[1])]
    rows = [[str(int(y[i]))] + [str(int(v)) for v in x[i]] for i in range(len(x))]
    p = str(tmp_path / "mnist.csv.gz")
    _write_csv(p, names, rows, True)
    fr = ingest.read_csv(p)
    assert np.array_equal(fr.filter_ix(lambda nm: nm != "label").to_mat(), x)
    assert np.array_equal(fr.first_col("label").astype(np.int64), y.astype(np.int64))
