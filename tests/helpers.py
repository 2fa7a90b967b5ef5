"""Shared test helpers: golden fixture loading, synthetic tables, result comparison."""
import math
import os
from typing import List, Sequence

import numpy as np

from snappydata_b200.column_format import ColumnBatch, SqlType, stats_row

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NUM_LINEITEM_COLS = 16


def load_tpch_golden():
    """-> (batches, snappy_1_out lines, snappy_6_out value).  The fixture holds the 7 columns Q1/Q6
    read at their table ordinals 4..10 (tests/golden/make_tpch_golden.py)."""
    z = np.load(os.path.join(GOLDEN, "tpch_lineitem.npz"))
    batches = []
    for i in range(int(z["num_batches"][0])):
        nrows, bucket = (int(x) for x in z[f"b{i}_meta"])
        cols: List = [None] * NUM_LINEITEM_COLS
        for c in range(7):
            cols[4 + c] = z[f"b{i}_c{c}"].tobytes()
        batches.append(ColumnBatch(num_rows=nrows, columns=cols, stats=None, batch_id=i, bucket_id=bucket))
    q1 = bytes(z["snappy_1_out"]).decode().strip().splitlines()
    q6 = bytes(z["snappy_6_out"]).decode().strip()
    return batches, q1, q6


def format_q1(final_rows: Sequence[Sequence[object]]) -> List[str]:
    """Format like the reference's QueryExecutor (%18.4f per double, then trimmed;
    cluster/src/test/scala/io/snappydata/benchmark/snappy/tpch/QueryExecutor.scala:152-157) and sort
    like TPCHDUnitTest does before comparing."""
    out = []
    for r in final_rows:
        parts = []
        for v in r:
            if isinstance(v, bytes):
                parts.append(v.decode())
            elif isinstance(v, float):
                parts.append(("%18.4f" % v).strip())
            else:
                parts.append(str(v))
        out.append(",".join(parts))
    return sorted(out)


def rows_close(a, b, rel=1e-6) -> bool:
    """COUNT / integer fields bit-exact, DOUBLE within `rel` relative (BASELINE.json north_star)."""
    if len(a) != len(b):
        return False
    for x, y in zip(a, b):
        if isinstance(x, float) or isinstance(y, float):
            if x is None or y is None:
                if x is not y:
                    return False
                continue
            if math.isnan(x) or math.isnan(y):
                if not (math.isnan(x) and math.isnan(y)):
                    return False
                continue
            if x == y:
                continue
            if abs(x - y) > rel * max(abs(x), abs(y)):
                return False
        elif x != y:
            return False
    return True


def _key(r):
    return tuple((0, b"") if v is None else (1, v) if isinstance(v, bytes) else (2, float(v)) if not isinstance(v, float) or not math.isnan(v) else (3, 0.0) for v in r)


def assert_rowsets_match(got, want, nkeys: int, rel=1e-6):
    """Compare two partial/final row sets irrespective of group order (the reference's tests sort rows)."""
    def k(r):
        return _key(r[:nkeys])
    g, w = sorted(got, key=k), sorted(want, key=k)
    assert len(g) == len(w), f"row count {len(g)} != {len(w)}\n got={g}\nwant={w}"
    for x, y in zip(g, w):
        assert rows_close(x, y, rel), f"row mismatch\n got={x}\nwant={y}"
