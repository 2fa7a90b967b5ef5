"""Small driver for ncu: resident scans of Q1 / Q6 over a synthetic lineitem store.
usage: python tools/profile_scan.py [q1|q6] [rows] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snappydata_b200 import capi, lineitem, plan as P  # noqa: E402

q = sys.argv[1] if len(sys.argv) > 1 else "q1"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
api = capi.product_api()
api.check(api.init(0))
if q == "rd":   # read ceiling of the kernel structure: sum of the four DOUBLE columns, no filter (32 B/row)
    from snappydata_b200.column_format import SqlType
    b = P.PlanBuilder()
    cs = [b.col(SqlType.DOUBLE, o) for o in (P.L_QUANTITY, P.L_EXTENDEDPRICE, P.L_DISCOUNT, P.L_TAX)]
    b.sum(cs[0]).sum(cs[1]).sum(cs[2]).sum(cs[3])
    desc, lits, mask = b.build(), [], lineitem.Q1_COLUMN_MASK
else:
    desc, lits, mask = (P.q1_plan(), P.Q1_LITERALS, lineitem.Q1_COLUMN_MASK) if q == "q1" else (P.q6_plan(), P.Q6_LITERALS, lineitem.Q6_COLUMN_MASK)
store = capi.Store(api, lineitem.LINEITEM_SCHEMA)
store.gen_lineitem(0, rows, 200_000, 128, 1, mask)
plan = capi.Plan(api, desc)
for _ in range(steps):
    plan.reset().set_literals(lits)
    plan.scan_store(store)
    out = plan.finish()
    m = plan.metrics()
    print(q, "rows", rows, "kernel_ms", m["aggTimeNs"] / 1e6, "GB/s", m["algorithmicBytes"] / max(1, m["aggTimeNs"]), "groups", len(out))
