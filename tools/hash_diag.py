"""Diagnostic (round 2): MODE_HASH group-by l_shipdate over generated lineitem, EVERY group compared with numpy.
usage: python tools/hash_diag.py <nbatches> <repeats>      env: SD_TUNE_NSTAGES, SD_JIT_DEFINES (-DSD_EXP_VERIFY=1 with SD_DEBUG_VERIFY=1:
staged tiles against global memory; -DSD_EXP_NO_PROXY_FENCE: the release without the generic->async proxy fence) ..."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snappydata_b200 import capi, lineitem, plan as P  # noqa: E402
from snappydata_b200.column_format import SqlType as T  # noqa: E402
from snappydata_b200.plan import PlanBuilder  # noqa: E402

NB = int(sys.argv[1]) if len(sys.argv) > 1 else 300
REP = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = NB * 200_000
cnt = np.zeros(2526, np.int64); sq = np.zeros(2526, np.float64); sp = np.zeros(2526, np.float64)
for f in range(0, N, 10_000_000):
    v = lineitem.lineitem_values(f, min(10_000_000, N - f), 1)
    g = v["l_shipdate"] - 8036
    cnt += np.bincount(g, minlength=2526)
    sq += np.bincount(g, weights=v["l_quantity"], minlength=2526)
    sp += np.bincount(g, weights=np.round(v["l_extendedprice"] * 100), minlength=2526)   # cents: exact

api = capi.product_api()
api.check(api.init(0))
store = capi.Store(api, lineitem.LINEITEM_SCHEMA, 0)
store.gen_lineitem(0, N, 200_000, 8, 1, lineitem.Q1_COLUMN_MASK)
b = PlanBuilder()
ship, qty, price = b.col(T.DATE, P.L_SHIPDATE), b.col(T.DOUBLE, P.L_QUANTITY), b.col(T.DOUBLE, P.L_EXTENDEDPRICE)
b.group_by(ship); b.count().sum(qty).sum(price)
gp = capi.Plan(api, b.build())
tag = "defines=[%s] nstages=%s NB=%d" % (os.environ.get("SD_JIT_DEFINES", ""), os.environ.get("SD_TUNE_NSTAGES", "-"), NB)
for it in range(REP):
    gp.reset().set_literals([])
    gp.scan_store(store)
    rows = gp.finish()
    m = gp.metrics()
    bad = []
    for r in rows:
        g = r[0] - 8036
        dc, dq, dp = r[1] - cnt[g], r[2] - sq[g], round(r[3] * 100) - sp[g]
        if dc or dq or abs(dp) > 2:
            bad.append((r[0], int(dc), float(dq), float(dp)))
    print("%s run %d: %d groups, %d launches, %.3f ms, groups that differ: %d, sum of diffs: count %d qty %.0f price(cents) %.0f" % (
        tag, it, len(rows), m["kernelLaunches"], m["aggTimeNs"] / 1e6, len(bad), sum(x[1] for x in bad), sum(x[2] for x in bad), sum(x[3] for x in bad)), flush=True)
    for x in bad[:12]:
        print("    shipdate %d: count %+d  sum_qty %+.0f  sum_price_cents %+.0f" % x)
