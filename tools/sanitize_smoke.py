"""Small end-to-end run of every kernel mode, meant to be run under compute-sanitizer (memcheck / racecheck):
  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Sizes are tiny; every result is still checked against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import assert_rowsets_match  # noqa: E402
from oracle import oracle  # noqa: E402
from snappydata_b200 import capi, lineitem, plan as P  # noqa: E402
from snappydata_b200.column_format import SqlType as T, compress_lz4  # noqa: E402
from snappydata_b200.plan import PlanBuilder  # noqa: E402
import test_gpu_general as G  # noqa: E402

api = capi.product_api()
api.check(api.init(0))


def check(desc, lits, batches, nk, rows=None, sort_all=False):
    gp = capi.Plan(api, desc).set_literals(lits)
    op = oracle.plan(desc).set_literals(lits)
    for x in batches:
        gp.submit(x)
        op.submit(x)
    got, want = gp.finish(), op.finish()
    if sort_all:
        assert_rowsets_match(got, want, len(want[0]) if want else 0)   # whole row as the sort key (NaN-safe compare)
    else:
        assert_rowsets_match(got, want, nk)
    print("ok", gp.kernel_name(), len(got), "rows")


li = lineitem.gen_table(5000, 2048, seed=3)
check(P.q6_plan(), P.Q6_LITERALS, li, 0)
check(P.q1_plan(), P.Q1_LITERALS, li, 2)
gen = [G.make_batch(n, seed=10 + i, batch_id=i)[0] for i, n in enumerate((1500, 1, 77))]
b = PlanBuilder(); c = G.cols(b)
b.filter(c["c0"].is_null() | (c["c0"] > b.lit(T.INT))); b.group_by(c["c3"]); b.count().sum(c["c2"]).min(c["c7"]).max(c["c9"])
check(b.build(), [-500], gen, 1)
check(b.build(), [-500], [G.with_deltas_and_deletes(1200, 100)], 1)            # general path with deltas/deletes
b = PlanBuilder(); c = G.cols(b)
b.group_by(c["c0"], c["c4"]); b.count().sum(c["c1"])
check(b.build(), [], gen, 2)                                                   # hash table
b = PlanBuilder(); c = G.cols(b)
b.filter(c["c5"] >= b.lit(T.DATE)); b.project(c["c0"], c["c3"], c["c2"])
check(b.build(), [9010], gen, 0, sort_all=True)                                # projection
import copy
lz = []
for x in li:
    y = copy.copy(x); y.columns = [None if v is None else compress_lz4(v, force=True) for v in x.columns]; lz.append(y)
gp = capi.Plan(api, P.q1_plan()).set_literals(P.Q1_LITERALS)
op = oracle.plan(P.q1_plan()).set_literals(P.Q1_LITERALS)
for x, y in zip(li, lz):
    op.submit(x); gp.submit(y)
assert_rowsets_match(gp.finish(), op.finish(), 2)
print("ok lz4")
# overlay path: lineitem with deltas + deletes over a resident store
from snappydata_b200.column_format import encode_delta, encode_delete
r = np.random.default_rng(1)
hy = lineitem.gen_table(6000, 3000, seed=8, column_mask=lineitem.Q6_COLUMN_MASK)
for x in hy:
    n = x.num_rows
    p0 = np.sort(r.choice(n, 30, replace=False)).astype(np.int32)
    p1 = np.sort(np.unique(np.concatenate([r.choice(n, 200, replace=False), p0[:5]]))).astype(np.int32)
    x.delta0[P.L_DISCOUNT] = encode_delta(n, p0, r.integers(0, 11, len(p0)) / 100.0, T.DOUBLE)
    x.delta1[P.L_DISCOUNT] = encode_delta(n, p1, r.integers(0, 11, len(p1)) / 100.0, T.DOUBLE)
    x.delete_mask = encode_delete(n, np.sort(r.choice(n, 40, replace=False)))
check(P.q6_plan(), P.Q6_LITERALS, hy, 0)
print("ALL OK")
