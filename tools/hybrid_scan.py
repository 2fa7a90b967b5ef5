"""BASELINE.json configs[4] at reduced scale: Q6 over lineitem batches that carry update deltas
(depth 0 and 1, a few positions in both), a delete mask and row-buffer rows -- the general decode path.
Prints GPU device time / algorithmic GB/s and the oracle's CPU time on the same bytes, after checking parity.
usage: python tools/hybrid_scan.py [batches] [rows_per_batch]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import assert_rowsets_match  # noqa: E402
from oracle import oracle  # noqa: E402
from snappydata_b200 import capi, lineitem, plan as P  # noqa: E402
from snappydata_b200.column_format import SqlType, encode_delete, encode_delta, unsafe_row  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rpb = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
r = np.random.default_rng(5)
batches = lineitem.gen_table(nb * rpb, rpb, seed=6, column_mask=lineitem.Q6_COLUMN_MASK)
for b in batches:
    n = b.num_rows
    upd = np.sort(r.choice(n, size=n // 200, replace=False)).astype(np.int32)          # 0.5 % updated
    d0 = upd[:100]
    d1 = np.sort(np.unique(np.concatenate([upd[100:], d0[:5]])))                        # a few positions in both
    for col, vals in ((P.L_DISCOUNT, lambda m: r.integers(0, 11, m) / 100.0), (P.L_QUANTITY, lambda m: r.integers(1, 51, m).astype(np.float64))):
        b.delta0[col] = encode_delta(n, d0, vals(len(d0)), SqlType.DOUBLE)
        b.delta1[col] = encode_delta(n, d1, vals(len(d1)), SqlType.DOUBLE)
    b.delete_mask = encode_delete(n, np.sort(r.choice(n, size=n // 200, replace=False)))  # 0.5 % deleted
rows = b""
nrb = 10_000
for i in range(nrb):
    row = unsafe_row([(SqlType.DATE, int(8036 + r.integers(0, 2526))), (SqlType.DOUBLE, float(r.integers(0, 11) / 100.0)),
                      (SqlType.DOUBLE, float(r.integers(1, 51))), (SqlType.DOUBLE, float(r.integers(90000, 10500000) / 100.0))])
    rows += len(row).to_bytes(8, "little") + row

api = capi.product_api()
api.check(api.init(0))
desc = P.q6_plan()
store = capi.Store(api, lineitem.LINEITEM_SCHEMA)
for b in batches:
    store.put(b)
gp = capi.Plan(api, desc)
for _ in range(4):
    gp.reset().set_literals(P.Q6_LITERALS)
    gp.submit_rows(rows, nrb)
    gp.scan_store(store)
    got = gp.finish()
m = gp.metrics()
op = oracle.plan(desc).set_literals(P.Q6_LITERALS)
t0 = time.perf_counter()
op.submit_rows(rows, nrb)
for b in batches:
    op.submit(b)
want = op.finish()
cpu_s = time.perf_counter() - t0
assert_rowsets_match(got, want, 0)
om = op.metrics()
assert m["rowsScanned"] == om["rowsScanned"] and m["numRowsBuffer"] == om["numRowsBuffer"] == nrb
print(f"hybrid Q6: {nb} batches x {rpb} rows (+{nrb} row-buffer rows), 0.5% updated in 2 columns (2 delta levels), 0.5% deleted: parity ok")
print(f"  GPU general path: kernel {m['aggTimeNs'] / 1e6:.3f} ms, {m['algorithmicBytes'] / max(1, m['aggTimeNs']):.1f} GB/s algorithmic, "
      f"{m['rowsScanned'] / (m['aggTimeNs'] / 1e9) / 1e9:.2f} G rows/s; launches {m['kernelLaunches']}")
print(f"  oracle (generic interpreter, 1 thread): {cpu_s:.2f} s = {om['rowsScanned'] / cpu_s / 1e6:.1f} M rows/s")
