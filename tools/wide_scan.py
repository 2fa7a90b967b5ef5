"""BASELINE.json configs[3] at reduced scale: wide table c0..c127 cycling (INT, DOUBLE, dictionary
STRING with 1000 distinct 8-12 byte values), every 4th column nullable (10 % NULLs);
  SELECT c0..c7 WHERE c0 BETWEEN a AND b AND c2 = 'lit'   (~1 % combined selectivity)
Only the 8 scanned columns are materialised.  Checks a sample against the oracle, then times the resident scan.
usage: python tools/wide_scan.py [batches of 200k rows]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import assert_rowsets_match  # noqa: E402
from oracle import oracle  # noqa: E402
from snappydata_b200 import capi  # noqa: E402
from snappydata_b200.column_format import ColumnBatch, SqlType as T, encode_dictionary, encode_uncompressed  # noqa: E402
from snappydata_b200.plan import PlanBuilder  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 50
n = 200_000
r = np.random.default_rng(4)
words = np.array([b"str%05d" % i + b"x" * (1 + i % 5) for i in range(1000)], dtype="S12")
types = [(T.INT, T.DOUBLE, T.STRING)[i % 3] for i in range(128)]
schema = [(types[i], i % 4 == 0) for i in range(128)]
batches = []
t0 = time.perf_counter()
for b in range(nb):
    cols = [None] * 128
    for i in range(8):
        nulls = (r.random(n) < 0.1) if i % 4 == 0 else None
        if types[i] == T.INT:
            cols[i] = encode_uncompressed(r.integers(0, 1000, n).astype(np.int32), T.INT, nulls)
        elif types[i] == T.DOUBLE:
            cols[i] = encode_uncompressed(r.random(n) * 100.0, T.DOUBLE, nulls)
        else:
            cols[i] = encode_dictionary(words[r.integers(0, 1000, n)], T.STRING, nulls)
    batches.append(ColumnBatch(num_rows=n, columns=cols, batch_id=b, bucket_id=b % 8))
print(f"generated {nb} batches in {time.perf_counter() - t0:.1f} s")

pb = PlanBuilder()
c = [pb.col(types[i], i, i % 4 == 0) for i in range(8)]
pb.filter((c[0] >= pb.lit(T.INT)) & (c[0] <= pb.lit(T.INT)) & c[2].eq(pb.lit(T.STRING)))
pb.project(*c)
desc = pb.build()
lits = [0, 999, bytes(words[7])]     # c0 is NOT NULL-filtered by the BETWEEN (10 % NULL) and c2 = one of 1000 values: ~0.09 %
lits_1pct = None

api = capi.product_api()
api.check(api.init(0))
store = capi.Store(api, schema)
for b in batches:
    store.put(b)
gp = capi.Plan(api, desc)
# parity on the first two batches
gp.reset().set_literals(lits)
op = oracle.plan(desc).set_literals(lits)
for b in batches[:2]:
    gp.submit(b)
    op.submit(b)
want = op.finish()
assert_rowsets_match(gp.finish(), want, len(want[0]) if want else 0)
for _ in range(4):
    gp.reset().set_literals(lits)
    gp.scan_store(store)
    raw = gp.finish_raw()
    m = gp.metrics()
print(f"wide table: {nb * n} rows x 8 scanned columns, {m['numOutputRows']} rows out ({100.0 * m['numOutputRows'] / (nb * n):.3f} %): parity ok")
print(f"  kernel {m['aggTimeNs'] / 1e6:.3f} ms, {m['algorithmicBytes'] / max(1, m['aggTimeNs']):.1f} GB/s algorithmic read "
      f"({m['algorithmicBytes'] / (nb * n):.1f} B/row), {nb * n / (m['aggTimeNs'] / 1e9) / 1e9:.1f} G rows/s, kernel {gp.kernel_name()}")
