"""Synthetic TPC-H lineitem-shaped column tables.

The generator is *counter based*: every value is a pure function of (seed, global row index, stream),
so the numpy implementation here and the CUDA implementation in csrc/sd_gen.cu
(`sdx_store_gen_lineitem`) produce byte-identical ColumnBatch buffers, any shard of the table can be
generated independently on any rank, and the CPU oracle can be handed exactly the bytes the GPU scanned.

  value streams (h = mix(seed ^ mix(row * 16 + stream)), mix = splitmix64 finaliser)
    0 l_quantity       1 + h % 50                         DOUBLE
    1 l_extendedprice  (90000 + h % 10410000) / 100.0     DOUBLE   round(U[900, 105000), 2)
    2 l_discount       (h % 11) / 100.0                   DOUBLE
    3 l_tax            (h % 9) / 100.0                    DOUBLE
    4 l_shipdate       8036 + h % 2526                    DATE     1992-01-02 .. 1998-12-01
    5 flag choice      shipdate > 9298 (1995-06-17) -> ('N','O'); else h % 100: <2 ('N','F'), <51 ('R','F'), else ('A','F')

  l_returnflag / l_linestatus are dictionary encoded with per-batch dictionaries in first-seen order
  (int16 indexes), everything else Uncompressed, all NOT NULL -- the reference's default encoders for
  the schema of TPCHTableSchema.scala:122-143.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .column_format import ColumnBatch, SqlType, encode_dictionary, encode_uncompressed
from .plan import L_DISCOUNT, L_EXTENDEDPRICE, L_LINESTATUS, L_QUANTITY, L_RETURNFLAG, L_SHIPDATE, L_TAX

NUM_TABLE_COLS = 16
LINEITEM_SCHEMA = [(SqlType.LONG, False), (SqlType.LONG, False), (SqlType.LONG, False), (SqlType.INT, False),
                   (SqlType.DOUBLE, False), (SqlType.DOUBLE, False), (SqlType.DOUBLE, False), (SqlType.DOUBLE, False),
                   (SqlType.STRING, False), (SqlType.STRING, False), (SqlType.DATE, False), (SqlType.DATE, False),
                   (SqlType.DATE, False), (SqlType.STRING, False), (SqlType.STRING, False), (SqlType.STRING, False)]
Q1_COLUMN_MASK = sum(1 << c for c in (L_QUANTITY, L_EXTENDEDPRICE, L_DISCOUNT, L_TAX, L_RETURNFLAG, L_LINESTATUS, L_SHIPDATE))
Q6_COLUMN_MASK = sum(1 << c for c in (L_QUANTITY, L_EXTENDEDPRICE, L_DISCOUNT, L_SHIPDATE))

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def _h(rows: np.ndarray, stream: int, seed: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        return _mix(np.uint64(seed) ^ _mix(rows * np.uint64(16) + np.uint64(stream)))


def lineitem_values(first_row: int, n: int, seed: int):
    """Raw column values of global rows [first_row, first_row + n)."""
    rows = np.arange(first_row, first_row + n, dtype=np.uint64)
    qty = (np.uint64(1) + _h(rows, 0, seed) % np.uint64(50)).astype(np.float64)
    price = (np.uint64(90000) + _h(rows, 1, seed) % np.uint64(10410000)).astype(np.float64) / 100.0
    disc = (_h(rows, 2, seed) % np.uint64(11)).astype(np.float64) / 100.0
    tax = (_h(rows, 3, seed) % np.uint64(9)).astype(np.float64) / 100.0
    ship = (np.uint64(8036) + _h(rows, 4, seed) % np.uint64(2526)).astype(np.int32)
    r = (_h(rows, 5, seed) % np.uint64(100)).astype(np.int32)
    late = ship > 9298
    rf = np.where(late, b"N", np.where(r < 2, b"N", np.where(r < 51, b"R", b"A"))).astype("S1")
    ls = np.where(late, b"O", b"F").astype("S1")
    return {"l_quantity": qty, "l_extendedprice": price, "l_discount": disc, "l_tax": tax,
            "l_shipdate": ship, "l_returnflag": rf, "l_linestatus": ls}


def gen_batch(batch_index: int, rows_per_batch: int, total_rows: int, seed: int, nbuckets: int = 8,
              column_mask: int = Q1_COLUMN_MASK) -> ColumnBatch:
    """Global batch `batch_index` of a table of `total_rows` rows cut into `rows_per_batch` batches."""
    first = batch_index * rows_per_batch
    n = min(rows_per_batch, total_rows - first)
    v = lineitem_values(first, n, seed)
    cols: List[Optional[bytes]] = [None] * NUM_TABLE_COLS
    for name, ordinal, t in (("l_quantity", L_QUANTITY, SqlType.DOUBLE), ("l_extendedprice", L_EXTENDEDPRICE, SqlType.DOUBLE),
                             ("l_discount", L_DISCOUNT, SqlType.DOUBLE), ("l_tax", L_TAX, SqlType.DOUBLE),
                             ("l_shipdate", L_SHIPDATE, SqlType.DATE)):
        if column_mask & (1 << ordinal):
            cols[ordinal] = encode_uncompressed(v[name], t)
    if column_mask & (1 << L_RETURNFLAG):
        cols[L_RETURNFLAG] = encode_dictionary(v["l_returnflag"], SqlType.STRING)
    if column_mask & (1 << L_LINESTATUS):
        cols[L_LINESTATUS] = encode_dictionary(v["l_linestatus"], SqlType.STRING)
    return ColumnBatch(num_rows=n, columns=cols, stats=None, batch_id=batch_index, bucket_id=batch_index % nbuckets)


def gen_table(total_rows: int, rows_per_batch: int, seed: int, nbuckets: int = 8, column_mask: int = Q1_COLUMN_MASK,
              batches: Optional[Sequence[int]] = None) -> List[ColumnBatch]:
    nb = (total_rows + rows_per_batch - 1) // rows_per_batch
    idx = range(nb) if batches is None else batches
    return [gen_batch(b, rows_per_batch, total_rows, seed, nbuckets, column_mask) for b in idx]


def gen_c1_table(total_rows: int = 1_000_000, rows_per_batch: int = 200_000, seed: int = 1, sorted_values: bool = False):
    """BASELINE.json configs[0]: one INT NOT NULL column c1 ~ U[0, 1e6), with stats rows
    so that the sorted variant exercises batch skipping."""
    from .column_format import column_stats, stats_row
    rows = np.arange(total_rows, dtype=np.uint64)
    vals = (_h(rows, 0, seed) % np.uint64(1_000_000)).astype(np.int32)
    if sorted_values:
        vals = np.sort(vals)
    out = []
    for b, s in enumerate(range(0, total_rows, rows_per_batch)):
        v = vals[s: s + rows_per_batch]
        out.append(ColumnBatch(num_rows=v.shape[0], columns=[encode_uncompressed(v, SqlType.INT)],
                               stats=stats_row(v.shape[0], [column_stats(v, SqlType.INT)]), batch_id=b, bucket_id=b % 8))
    return out, vals
