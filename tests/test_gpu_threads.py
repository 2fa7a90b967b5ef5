"""Re-entrancy: many task threads drive their own plan handles concurrently against one resident store
(SURVEY.md 8b threading: one Spark task thread per partition, many tasks per executor JVM)."""
import threading

import pytest

from oracle import oracle
from snappydata_b200 import capi, lineitem
from snappydata_b200 import plan as P
from snappydata_b200.plan import PlanBuilder
from snappydata_b200.column_format import SqlType as T

from helpers import assert_rowsets_match

pytestmark = pytest.mark.gpu


def test_concurrent_plans_over_one_store(gpu_api):
    batches = lineitem.gen_table(160_000, 20_000, seed=13, nbuckets=8)
    store = capi.Store(gpu_api, lineitem.LINEITEM_SCHEMA)
    for b in batches:
        store.put(b)

    def jit_plan(i):   # distinct plans -> concurrent NVRTC compilations
        b = PlanBuilder()
        q = b.col(T.DOUBLE, P.L_QUANTITY)
        d = b.col(T.DATE, P.L_SHIPDATE)
        b.filter(d > b.lit(T.DATE))
        b.sum(q * b.lit(T.DOUBLE)).count().max(q + b.lit(T.DOUBLE))
        return b.build(), [9000 + 50 * i, float(i + 1), float(i)], 0

    work = [(P.q6_plan(), P.Q6_LITERALS, 0), (P.q1_plan(), P.Q1_LITERALS, 2)] + [jit_plan(i) for i in range(2)]
    results, errors = {}, []

    def task(tid, desc, lits, buckets):
        try:
            gpu_api.check(gpu_api.init(0))
            pl = capi.Plan(gpu_api, desc)
            for _ in range(3):
                pl.reset().set_literals(lits)
                pl.scan_store(store, buckets)
                results[tid] = pl.finish()
        except Exception as e:   # pragma: no cover
            errors.append((tid, repr(e)))

    threads, expect = [], {}
    for w, (desc, lits, nk) in enumerate(work):
        for part in range(2):
            buckets = [b for b in range(8) if b % 2 == part]
            tid = (w, part)
            op = oracle.plan(desc).set_literals(lits)
            for b in batches:
                if b.bucket_id in buckets:
                    op.submit(b)
            expect[tid] = (op.finish(), nk)
            threads.append(threading.Thread(target=task, args=(tid, desc, lits, buckets)))
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    for tid, (want, nk) in expect.items():
        assert_rowsets_match(results[tid], want, nk)
