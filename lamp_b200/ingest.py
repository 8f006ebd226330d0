"""CSV ingest for the extratrees path: file -> columns -> resident table, without a row-major intermediate.

The reference reads its tables with saddle's CsvParser into a Frame and takes `toMat` (row-major) of the feature
columns (tst:284-294, endtoendtest e2e.test.scala:31-40; lamp-saddle SaddleTensorHelpers.scala:156-176 is its
row-major -> native copy).  The device table is column-major, so a CSV is parsed by COLUMN here (pandas' C parser) and
whole columns go to the device in blocks (et_data_dense_colblock): no transpose, no n x d host copy in row order.
Fields that are empty or not numbers are NaN = missing values, like saddle's Double parser.
"""
from __future__ import annotations

import gzip
import io
from dataclasses import dataclass
from typing import Iterable, Iterator, Optional, Sequence, Union

import numpy as np

PathOrStream = Union[str, "io.IOBase"]


@dataclass
class Frame:
    """Named FP64 columns (the part of saddle's Frame this path uses)."""
    names: list
    columns: np.ndarray  # [d, n] column-major: columns[j] is column names[j]

    @property
    def n_rows(self) -> int:
        return int(self.columns.shape[1])

    def first_col(self, name: str) -> np.ndarray:
        """saddle `firstCol(name).toVec` (tst:291)."""
        return self.columns[self.names.index(name)]

    def filter_ix(self, keep) -> "Frame":
        """saddle `filterIx(pred)` (tst:293): the columns whose name satisfies `keep`, in file order."""
        idx = [j for j, nm in enumerate(self.names) if keep(nm)]
        return Frame([self.names[j] for j in idx], self.columns[idx])

    def to_mat(self) -> np.ndarray:
        """saddle `toMat`: row-major [n, d] (what the reference's entry points take)."""
        return np.ascontiguousarray(self.columns.T)


def _open(src: PathOrStream):
    if isinstance(src, str):
        with open(src, "rb") as f:
            magic = f.read(2)
        return gzip.open(src, "rt") if magic == b"\x1f\x8b" else open(src, "rt")
    return src


def read_csv(src: PathOrStream, header: bool = True, delimiter: str = ",") -> Frame:
    """Parses a (gzip-compressed or plain) delimited file of numbers into named FP64 columns."""
    import pandas as pd
    f = _open(src)
    try:
        df = pd.read_csv(f, sep=delimiter, header=0 if header else None, dtype=str, keep_default_na=False,
                         skip_blank_lines=True)
    finally:
        if isinstance(src, str):
            f.close()
    names = [str(c) for c in df.columns]
    cols = np.empty((len(names), len(df)), dtype=np.float64)
    for j, c in enumerate(df.columns):
        cols[j] = pd.to_numeric(df[c], errors="coerce").to_numpy(dtype=np.float64)  # not a number -> NaN (missing)
    return Frame(names, cols)


def column_blocks(columns: np.ndarray, block_cols: int = 64) -> Iterator[tuple]:
    """(first_col, [b, n] block) pieces of a column-major table, for DeviceData.from_column_blocks."""
    d = columns.shape[0]
    for first in range(0, d, block_cols):
        yield first, columns[first:first + block_cols]


def device_data_from_frame(features: Frame, ctx=None, block_cols: int = 64):
    """The feature columns of a Frame as a resident table (column blocks: no host transpose)."""
    from .extratrees import DeviceData
    d, n = features.columns.shape
    return DeviceData.from_column_blocks(n, d, column_blocks(features.columns, block_cols), ctx)


def device_data_from_csv(src: PathOrStream, label: Optional[str] = None, num_classes: Optional[int] = None,
                         regression: bool = False, drop: Sequence[str] = (), ctx=None):
    """CSV -> resident table with its target attached.  `label` names the target column (classification: integer
    class ids; regression: FP64); every other column not in `drop` is a feature, in file order.
    Returns (DeviceData, feature names, target or None)."""
    fr = read_csv(src)
    skip = set(drop) | ({label} if label is not None else set())
    feats = fr.filter_ix(lambda nm: nm not in skip)
    dd = device_data_from_frame(feats, ctx)
    target = None
    if label is not None:
        col = fr.first_col(label)
        if regression:
            target = np.ascontiguousarray(col)
            dd.set_target_regression(target)
        else:
            if np.isnan(col).any():
                raise ValueError("classification labels must be numbers")
            target = col.astype(np.int32)  # the reference: `.map(_.toLong)` (tst:291)
            dd.set_target_classification(target, int(num_classes if num_classes is not None else target.max() + 1))
    return dd, feats.names, target
