"""MODE_HASH kernel time at three cardinalities, with the per-CTA shared-memory front table and without it (round 2).
usage: python tools/hash_bench.py [nbatches]    (200 000-row lineitem batches generated on the device)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snappydata_b200 import capi, lineitem, plan as P  # noqa: E402
from snappydata_b200.column_format import SqlType as T  # noqa: E402
from snappydata_b200.plan import PlanBuilder  # noqa: E402

api = capi.product_api()
api.check(api.init(0))
NB = int(sys.argv[1]) if len(sys.argv) > 1 else 300
store = capi.Store(api, lineitem.LINEITEM_SCHEMA, 0)
store.gen_lineitem(0, NB * 200_000, 200_000, 8, 1, lineitem.Q1_COLUMN_MASK)


def build(which):
    b = PlanBuilder()
    ship, qty, price = b.col(T.DATE, P.L_SHIPDATE), b.col(T.DOUBLE, P.L_QUANTITY), b.col(T.DOUBLE, P.L_EXTENDEDPRICE)
    disc = b.col(T.DOUBLE, P.L_DISCOUNT)
    keys = {"shipdate": (ship,), "shipdate,quantity": (ship, qty), "shipdate,quantity,discount": (ship, qty, disc)}[which]
    b.group_by(*keys)
    b.count().sum(qty).sum(price)
    return b.build(), len(keys)


for which in ("shipdate", "shipdate,quantity", "shipdate,quantity,discount"):
    res = {}
    for front in (1, 0):
        if front:
            os.environ.pop("SD_TUNE_NO_FRONT_TABLE", None)
        else:
            os.environ["SD_TUNE_NO_FRONT_TABLE"] = "1"
        desc, nk = build(which)
        gp = capi.Plan(api, desc)
        best = None
        for it in range(4):     # the first passes grow the global table and replay; the last ones are steady state
            gp.reset().set_literals([])
            gp.scan_store(store)
            rows = gp.finish()
            m = gp.metrics()
            t = m["aggTimeNs"]
            best = t if best is None or t < best else best
        tot = sum(r[nk] for r in rows)
        chk = sum(r[nk + 1] for r in rows)
        res[front] = (len(rows), tot, chk)
        print("group by %-28s front=%d %8d groups  %8.3f ms  %7.1f GB/s  %6.2f G rows/s  [%s]" % (
            which, front, len(rows), best / 1e6, m["algorithmicBytes"] / best, NB * 200_000 / best, gp.kernel_name()), flush=True)
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1] == NB * 200_000, res
    assert abs(res[0][2] - res[1][2]) <= 1e-9 * abs(res[0][2]), res
print("OK")
