"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the reference's golden
results, on identical ColumnBatch bytes.  COUNT / integer results bit-exact, DOUBLE within 1e-6 relative
(BASELINE.json north_star)."""
import numpy as np
import pytest

from oracle import oracle
from snappydata_b200 import capi, lineitem
from snappydata_b200 import plan as P
from snappydata_b200.capi import final_merge

from helpers import assert_rowsets_match, format_q1, load_tpch_golden

pytestmark = pytest.mark.gpu


def run_both(gpu_api, desc, lits, batches, nkeys):
    gp = capi.Plan(gpu_api, desc).set_literals(lits)
    op = oracle.plan(desc).set_literals(lits)
    for b in batches:
        gp.submit(b)
        op.submit(b)
    got_raw, want_raw = gp.finish_raw(), op.finish_raw()
    got = capi.parse_row_stream(got_raw, desc.partial_schema())
    want = capi.parse_row_stream(want_raw, desc.partial_schema())
    assert_rowsets_match(got, want, nkeys)
    return gp, op, got_raw, want_raw


def test_tpch_golden_q1_q6_on_gpu(gpu_api):
    """The reference's own known-answer test (Snappy_1.out / Snappy_6.out) through the CUDA path."""
    batches, want_q1, want_q6 = load_tpch_golden()
    d1, d6 = P.q1_plan(), P.q6_plan()
    gp, _, raw, _ = run_both(gpu_api, d1, P.Q1_LITERALS, batches, 2)
    assert gp.kernel_name().startswith("aot:")
    assert format_q1(final_merge(gpu_api, d1, raw)) == sorted(want_q1)
    gp, _, raw, _ = run_both(gpu_api, d6, P.Q6_LITERALS, batches, 0)
    (row,) = final_merge(gpu_api, d6, raw)
    assert ("%18.4f" % row[0]).strip() == want_q6
    m = gp.metrics()
    assert m["rowsScanned"] == 30201 and m["columnBatchesSeen"] == len(batches) and m["kernelLaunches"] >= 1


@pytest.mark.parametrize("rows,per_batch", [(1, 100), (2, 100), (1023, 1024), (1025, 1024), (50_001, 12_000), (200_000, 200_000)])
def test_synthetic_lineitem_q1_q6(gpu_api, rows, per_batch):
    batches = lineitem.gen_table(rows, per_batch, seed=3)
    run_both(gpu_api, P.q6_plan(), P.Q6_LITERALS, batches, 0)
    run_both(gpu_api, P.q1_plan(), P.Q1_LITERALS, batches, 2)


def test_q6_no_matching_rows_gives_null_sum(gpu_api):
    """No-key SUM starts NULL (doProduceWithoutKeys, SnappyHashAggregateExec.scala:337-346)."""
    batches = lineitem.gen_table(5000, 2048, seed=3)
    gp, _, raw, _ = run_both(gpu_api, P.q6_plan(), [0, 1, 0.05, 0.07, 24.0], batches, 0)
    assert capi.parse_row_stream(raw, P.q6_plan().partial_schema()) == [[None]]


def test_empty_plan_outputs_one_row_without_keys_and_none_with_keys(gpu_api):
    gp = capi.Plan(gpu_api, P.q6_plan()).set_literals(P.Q6_LITERALS)
    assert gp.finish() == [[None]]
    gp = capi.Plan(gpu_api, P.q1_plan()).set_literals(P.Q1_LITERALS)
    assert gp.finish() == []


def test_c1_count_with_stats_skipping(gpu_api):
    """BASELINE.json configs[0]: COUNT(*) WHERE c1 > k; the sorted variant must skip batches through
    the stats row exactly as the reference does (ColumnTableScan.scala:532-543)."""
    desc = P.c1_plan()
    for sorted_values in (False, True):
        batches, vals = lineitem.gen_c1_table(1_000_000, 200_000, seed=1, sorted_values=sorted_values)
        for k in (0, 500_000, 999_999):
            gp, op, raw, _ = run_both(gpu_api, desc, [k], batches, 0)
            assert capi.parse_row_stream(raw, desc.partial_schema()) == [[int((vals > k).sum())]]
            gm, om = gp.metrics(), op.metrics()
            assert gm["columnBatchesSkipped"] == om["columnBatchesSkipped"]
            assert gm["columnBatchesSeen"] == om["columnBatchesSeen"] == 5
            if sorted_values and k == 999_999:
                assert gm["columnBatchesSkipped"] >= 4


def test_plan_reuse_with_new_literals(gpu_api):
    """One compiled plan serves every literal value (ParamLiteral tokenisation)."""
    batches = lineitem.gen_table(30_000, 8192, seed=11)
    desc = P.q6_plan()
    gp = capi.Plan(gpu_api, desc)
    for lits in (P.Q6_LITERALS, [8766, 9131, 0.02, 0.04, 30.0], [9000, 9500, 0.0, 0.1, 51.0]):
        gp.reset().set_literals(lits)
        op = oracle.plan(desc).set_literals(lits)
        for b in batches:
            gp.submit(b)
            op.submit(b)
        assert_rowsets_match(gp.finish(), op.finish(), 0)


def test_resident_store_scan_matches_submit_path(gpu_api):
    batches = lineitem.gen_table(70_000, 16_384, seed=5, nbuckets=4)
    store = capi.Store(gpu_api, lineitem.LINEITEM_SCHEMA)
    for b in batches:
        store.put(b)
    assert store.num_batches() == len(batches)
    for desc, lits, nk in ((P.q6_plan(), P.Q6_LITERALS, 0), (P.q1_plan(), P.Q1_LITERALS, 2)):
        op = oracle.plan(desc).set_literals(lits)
        for b in batches:
            op.submit(b)
        want = op.finish()
        gp = capi.Plan(gpu_api, desc).set_literals(lits)
        for _ in range(3):                      # repeated executions of the cached plan
            gp.reset().set_literals(lits)
            gp.scan_store(store)
            assert_rowsets_match(gp.finish(), want, nk)
        # bucket subset
        op = oracle.plan(desc).set_literals(lits)
        for b in batches:
            if b.bucket_id in (1, 3):
                op.submit(b)
        gp.reset().set_literals(lits)
        gp.scan_store(store, [1, 3])
        assert_rowsets_match(gp.finish(), op.finish(), nk)


def test_device_generator_is_byte_identical_to_numpy(gpu_api):
    rows, per_batch = 45_001, 10_000
    store = capi.Store(gpu_api, lineitem.LINEITEM_SCHEMA)
    store.gen_lineitem(0, rows, per_batch, 8, 42, lineitem.Q1_COLUMN_MASK)
    host = lineitem.gen_table(rows, per_batch, seed=42)
    assert store.num_batches() == len(host)
    for i, hb in enumerate(host):
        n, bucket, bid = store.batch_info(i)
        assert (n, bucket, bid) == (hb.num_rows, hb.bucket_id, hb.batch_id)
        for c in range(4, 11):
            assert store.get_buffer(i, c) == hb.columns[c], f"batch {i} column {c}"
    # and scanning the generated store agrees with the oracle over the numpy bytes
    for desc, lits, nk in ((P.q6_plan(), P.Q6_LITERALS, 0), (P.q1_plan(), P.Q1_LITERALS, 2)):
        op = oracle.plan(desc).set_literals(lits)
        for b in host:
            op.submit(b)
        gp = capi.Plan(gpu_api, desc).set_literals(lits)
        gp.scan_store(store)
        assert_rowsets_match(gp.finish(), op.finish(), nk)


def test_sharded_generation_covers_the_same_table(gpu_api):
    """Two 'ranks' generating disjoint batch ranges produce the bytes of the whole table."""
    rows, per_batch = 40_000, 10_000
    whole = lineitem.gen_table(rows, per_batch, seed=9)
    store = capi.Store(gpu_api, lineitem.LINEITEM_SCHEMA)
    store.gen_lineitem(20_000, 20_000, per_batch, 8, 9, lineitem.Q6_COLUMN_MASK)
    for i in range(2):
        for c in (4, 5, 6, 10):
            assert store.get_buffer(i, c) == whole[2 + i].columns[c]
