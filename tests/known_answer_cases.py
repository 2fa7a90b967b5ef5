"""Closed-form known answers of the reference's own aggregate / update / delete tests
(SURVEY.md 8c G4, G6), restated over ColumnBatch bytes written by the fixture writer:

  * cluster/src/test/scala/org/apache/spark/sql/store/SHAByteBufferTest.scala:225-267 -- one nullable STRING key, 200
    rows, sums checked against an arithmetic-series formula and read as LONG (which pins Sum(INT) -> LONG);
  * ...:281-331 -- two nullable STRING keys, 50 x 10 rows, 100 groups;
  * core/src/test/scala/io/snappydata/ColumnUpdateDeleteTests.scala:205-250 (testDeltaStats) -- a 2-row batch whose
    column is updated: the point filter on the NEW value must find the row (the merged stats row may not skip the batch)
    and the old value must be gone; depth-0 delta wins over depth-1;
  * ...:283-332 (testBasicDeleteIter) -- 50,000 rows, every 10th deleted, two UPDATE passes on the survivors:
    counts are exact, id 73 is untouched and found by a point filter.

The cases are engine-agnostic: tests/test_oracle_known_answers.py runs them on the CPU oracle (pinning it),
tests/test_gpu_known_answers.py (-m gpu) runs the SAME cases through the CUDA path.
"""
import numpy as np

from oracle import oracle
from snappydata_b200.capi import final_merge
from snappydata_b200.column_format import (ColumnBatch, SqlType as T, build_batch, column_stats, encode_delete, encode_delta,
                                           encode_uncompressed, stats_row)
from snappydata_b200.plan import PlanBuilder


class OracleEngine:
    """Partial stage = the CPU oracle; its partial rows also go through the product's HOST-side final merge
    (sd_final_merge in libsnappygpu.so makes no CUDA call) and both final results must agree."""

    def __init__(self, oracle_api):
        self.api = oracle_api

    def __call__(self, desc, lits, batches):
        pl = oracle.plan(desc).set_literals(lits)
        for b in batches:
            pl.submit(b)
        raw = pl.finish_raw()
        final = final_merge(self.api, desc, raw)
        from snappydata_b200 import capi
        product = final_merge(capi.product_api(), desc, raw)
        assert sorted(map(repr, product)) == sorted(map(repr, final))
        return final, pl


class GpuEngine:
    """Partial stage = the CUDA path through the C ABI (sd_plan_create / sd_batch_submit / sd_plan_finish), final merge =
    sd_final_merge.  The closed forms below are then checked on what the GPU produced -- not GPU vs oracle."""

    def __init__(self, gpu_api):
        self.api = gpu_api

    def __call__(self, desc, lits, batches):
        from snappydata_b200 import capi
        pl = capi.Plan(self.api, desc).set_literals(lits)
        for b in batches:
            pl.submit(b)
        raw = pl.finish_raw()
        m = pl.metrics()
        assert m["kernelLaunches"] >= 1 or m["columnBatchesSkipped"] == len(batches)   # (every batch skipped by its stats row: nothing to launch)
        return final_merge(self.api, desc, raw), pl


def case_sha_one_nullable_string_key_closed_form(_run):
    rng, div = 200, 10
    i = np.arange(rng)
    schema = [("col1", T.INT, True), ("col2", T.INT, True), ("col3", T.INT, True), ("col4", T.STRING, True)]
    data = {"col1": i.astype(np.int32), "col2": (2 * i).astype(np.int32), "col3": (3 * i).astype(np.int32),
            "col4": np.array([b"test%d" % (x % div) for x in i], dtype=object)}
    nulls = {"col4": (i % div) == 0}
    batch = build_batch(rng, schema, data, nulls)
    b = PlanBuilder()
    c1, c2, c4 = b.col(T.INT, 0, True), b.col(T.INT, 1, True), b.col(T.STRING, 3, True)
    b.group_by(c4)
    b.sum(c1).sum(c2)
    rows, _ = _run(b.build(), [], [batch])
    assert len(rows) == div                                     # 9 strings + the NULL key
    n = rng // div
    for key, s1, s2 in rows:
        k = 0 if key is None else int(key[len(b"test"):])
        assert isinstance(s1, int) and isinstance(s2, int)      # Sum(INT) -> LONG
        assert s1 == int((n / 2.0) * (2 * k + (n - 1) * div))
        assert s2 == int((n / 2.0) * (2 * 2 * k + (n - 1) * 2 * div))


def case_sha_two_nullable_string_keys_closed_form(_run):
    rng, d1, d2 = 50, 10, 10
    i = np.repeat(np.arange(rng), d2)
    j = np.tile(np.arange(d2), rng)
    n_rows = rng * d2
    schema = [("col1", T.INT, True), ("col2", T.INT, True), ("col3", T.INT, True), ("col4", T.STRING, True), ("col5", T.STRING, True)]
    data = {"col1": i.astype(np.int32), "col2": (2 * i).astype(np.int32), "col3": (3 * i).astype(np.int32),
            "col4": np.array([b"test%d" % (x % d1) for x in i], dtype=object),
            "col5": np.array([b"test%d" % (x % d2) for x in j], dtype=object)}
    nulls = {"col4": (i % d1) == 0, "col5": (j % d2) == 0}
    # two batches so that per-batch dictionaries differ in order
    half = n_rows // 2
    batches = []
    for lo, hi, bid in ((0, half, 0), (half, n_rows, 1)):
        batches.append(build_batch(hi - lo, schema, {k: v[lo:hi] for k, v in data.items()}, {k: v[lo:hi] for k, v in nulls.items()}, batch_id=bid))
    b = PlanBuilder()
    c1, c2, c4, c5 = b.col(T.INT, 0, True), b.col(T.INT, 1, True), b.col(T.STRING, 3, True), b.col(T.STRING, 4, True)
    b.group_by(c4, c5)
    b.sum(c1).sum(c2)
    rows, _ = _run(b.build(), [], batches)
    assert len(rows) == d1 * d2
    n = rng // d1
    for k4, k5, s1, s2 in rows:
        k = 0 if k4 is None else int(k4[len(b"test"):])
        assert s1 == int((n / 2.0) * (2 * k + (n - 1) * d1))
        assert s2 == int((n / 2.0) * (2 * 2 * k + (n - 1) * 2 * d1))


def _point_query(col_ordinal, other_ordinal):
    b = PlanBuilder()
    c = [b.col(T.LONG, 0, False), b.col(T.LONG, 1, False)]
    b.filter(c[col_ordinal].eq(b.lit(T.LONG)))
    b.count().sum(c[0]).sum(c[1])
    return b.build()


def case_delta_stats_point_filters_after_updates(_run):
    """testDeltaStats: rows (10,100),(20,200); `update col1 = 100 where col2 = 100` then `update col1 = 200 where col1 = 20`
    then `update col1 = col1 * 10 ...`: the later update of a position lives in the depth-0 delta and wins."""
    col1 = np.array([10, 20], dtype=np.int64)
    col2 = np.array([100, 200], dtype=np.int64)
    base = [encode_uncompressed(col1, T.LONG, None), encode_uncompressed(col2, T.LONG, None)]

    def batch(d0=None, d1=None, eff=None):
        # the writer merges the deltas' bounds into the batch's stats row and flips the sign of the count
        # (encoders/.../impl/ColumnDelta.scala:134-222); `eff` = effective col1 values after the updates
        st = stats_row(2, [column_stats(np.concatenate([col1, eff]) if eff is not None else col1, T.LONG), column_stats(col2, T.LONG)],
                       has_deltas=eff is not None)
        return ColumnBatch(num_rows=2, columns=list(base), stats=st, delta0=({0: d0} if d0 else {}), delta1=({0: d1} if d1 else {}))

    q_col1, q_col2 = _point_query(0, 1), _point_query(1, 0)
    # after the first update: (100,100),(20,200)
    b1 = batch(d0=encode_delta(2, [0], np.array([100], dtype=np.int64), T.LONG), eff=np.array([100, 20]))
    assert _run(q_col1, [100], [b1])[0] == [[1, 100, 100]]
    assert _run(q_col2, [100], [b1])[0] == [[1, 100, 100]]
    assert _run(q_col1, [10], [b1])[0][0][0] == 0          # the old value is gone
    # second update (col1 = 200 where col1 = 20) merged into the same delta: (100,100),(200,200)
    b2 = batch(d0=encode_delta(2, [0, 1], np.array([100, 200], dtype=np.int64), T.LONG), eff=np.array([100, 200]))
    assert _run(q_col1, [200], [b2])[0] == [[1, 200, 200]]
    assert _run(q_col2, [200], [b2])[0] == [[1, 200, 200]]
    # third update (x10): the previous delta has moved to depth 1, the new values sit at depth 0 and win
    b3 = batch(d0=encode_delta(2, [0, 1], np.array([1000, 2000], dtype=np.int64), T.LONG),
               d1=encode_delta(2, [0, 1], np.array([100, 200], dtype=np.int64), T.LONG), eff=np.array([100, 200, 1000, 2000]))
    assert _run(q_col1, [1000], [b3])[0] == [[1, 1000, 100]]
    assert _run(q_col1, [2000], [b3])[0] == [[1, 2000, 200]]
    assert _run(q_col1, [100], [b3])[0][0][0] == 0
    assert _run(q_col2, [100], [b3])[0] == [[1, 1000, 100]]


def case_basic_delete_and_update_counts(_run):
    """testBasicDeleteIter: 50,000 rows (id, status); delete where id % 10 = 0; two passes of
    `update id = id + 25000 where id <> 73` on the survivors."""
    n, per = 50_000, 10_000
    batches, batches_upd = [], []
    for bid in range(n // per):
        ids = np.arange(bid * per, (bid + 1) * per, dtype=np.int32)
        status = (ids % 2) == 0
        cols = [encode_uncompressed(ids, T.INT, None), encode_uncompressed(status, T.BOOLEAN, None)]
        dele = encode_delete(per, np.nonzero(ids % 10 == 0)[0])
        batches.append(ColumnBatch(num_rows=per, columns=cols, delete_mask=dele, batch_id=bid))
        pos = np.nonzero((ids % 10 != 0) & (ids != 73))[0]          # survivors except id 73
        d1 = encode_delta(per, pos, (ids[pos] + n // 2).astype(np.int32), T.INT)      # first pass (older: depth 1)
        d0 = encode_delta(per, pos, (ids[pos] + n).astype(np.int32), T.INT)           # second pass on top of it (depth 0)
        batches_upd.append(ColumnBatch(num_rows=per, columns=cols, delete_mask=dele, delta0={0: d0}, delta1={0: d1}, batch_id=bid))
    b = PlanBuilder()
    idc = b.col(T.INT, 0, False)
    b.count().sum(idc)
    count_sum = b.build()
    survivors = np.array([x for x in range(n) if x % 10 != 0], dtype=np.int64)
    (cnt, total), = _run(count_sum, [], batches)[0]
    assert cnt == (n * 9) // 10 and total == int(survivors.sum())
    (cnt, total), = _run(count_sum, [], batches_upd)[0]
    assert cnt == (n * 9) // 10
    assert total == int(survivors.sum()) + n * (len(survivors) - 1)     # every survivor but id 73 moved by n
    b = PlanBuilder()
    idc, st = b.col(T.INT, 0, False), b.col(T.BOOLEAN, 1, False)
    b.filter(idc.eq(b.lit(T.INT)))
    b.count().max(idc).count(st)
    (cnt, mx, cs), = _run(b.build(), [73], batches_upd)[0]
    assert [cnt, mx, cs] == [1, 73, 1]
    (cnt, _, _), = _run(b.build(), [74], batches_upd)[0]
    assert cnt == 0                                                        # 74 became 74 + n


def case_sha_sum_of_every_numeric_type_per_string_key(_run):
    """SHAByteBufferTest.scala:534-700 ("aggregate functions & grouping on each of spark data type"): ten rows i = 0..9 with
    the value i in a column of the type under test (every other column NULL) and the key 'col{i/5}':
    sum per key = 10 and 35 -- read with getLong for BYTE/SHORT/INT/LONG, with getDouble for FLOAT/DOUBLE."""
    i = np.arange(10)
    keys = np.array([b"col%d" % (x // 5) for x in i], dtype=object)
    cases = [(T.BYTE, np.int8, int), (T.SHORT, np.int16, int), (T.INT, np.int32, int), (T.LONG, np.int64, int),
             (T.FLOAT, np.float32, float), (T.DOUBLE, np.float64, float)]
    for t, dt, py in cases:
        schema = [("col000", T.INT, True), ("v", t, True), ("k", T.STRING, True)]
        batch = build_batch(10, schema, {"col000": i.astype(np.int32), "v": i.astype(dt), "k": keys}, {})
        b = PlanBuilder()
        v, k = b.col(t, 1, True), b.col(T.STRING, 2, True)
        b.group_by(k)
        b.sum(v)
        rows, _ = _run(b.build(), [], [batch])
        got = {key: s for key, s in rows}
        assert got == {b"col0": py(10), b"col1": py(35)}, (t, got)
        assert all(type(s) is py for s in got.values()), (t, got)


def case_sha_decimal_sum_and_avg_per_string_key(_run):
    """SHAByteBufferTest.scala:710-731 ("Big Decimal with precision < 18 as aggregate column"): decimal(12,5) column holding
    0.3 * i for i = 0..9, key 'col{i/5}': sum = 3.0 and 10.5 (the test allows 0.1; here the unscaled values are exact).
    Spark 2.1.1: Sum(DECIMAL(12,5)) -> DECIMAL(22,5) (wider than 18 digits -> BigInteger bytes in the UnsafeRow),
    Average -> DECIMAL(16,9) = sum / count rounded HALF_UP."""
    i = np.arange(10)
    keys = np.array([b"col%d" % (x // 5) for x in i], dtype=object)
    unscaled = (30000 * i).astype(np.int64)                      # 0.3 * i at scale 5
    schema = [("v", T.DECIMAL, True), ("k", T.STRING, True)]
    batch = build_batch(10, schema, {"v": unscaled, "k": keys}, {})
    b = PlanBuilder()
    v, k = b.col(T.DECIMAL, 0, True, scale=5, precision=12), b.col(T.STRING, 1, True)
    b.group_by(k)
    b.sum(v).avg(v).min(v).max(v).count(v)
    desc = b.build()
    assert desc.final_schema()[1:3] == [(T.DECIMAL, 22, 5), (T.DECIMAL, 16, 9)]
    rows, _ = _run(desc, [], [batch])
    got = {r[0]: r[1:] for r in rows}
    assert got[b"col0"] == [300000, 600000000, 0, 120000, 5]              # 3.00000, 0.600000000, 0, 1.2
    assert got[b"col1"] == [1050000, 2100000000, 150000, 270000, 5]       # 10.50000, 2.100000000, 1.5, 2.7
    # a sum that does not fit 18 digits, and HALF_UP rounding of the average (negative and positive)
    big = np.array([999_999_999_999_999_999, 999_999_999_999_999_999, 1], dtype=np.int64)     # DECIMAL(18,0)
    small = np.array([1, 1, 0], dtype=np.int64)                                               # avg = 2/3 -> 0.6667
    neg = -small
    batch = build_batch(3, [("a", T.DECIMAL, False), ("b", T.DECIMAL, False), ("c", T.DECIMAL, False)], {"a": big, "b": small, "c": neg}, {})
    b = PlanBuilder()
    ca, cb, cc = b.col(T.DECIMAL, 0, False, 0, 18), b.col(T.DECIMAL, 1, False, 0, 18), b.col(T.DECIMAL, 2, False, 0, 18)
    b.sum(ca).avg(cb).avg(cc)
    (s, a1, a2), = _run(b.build(), [], [batch])[0]
    assert s == 1_999_999_999_999_999_999 and a1 == 6667 and a2 == -6667


def case_casts_follow_spark(_run):
    """Spark 2.1.1 Cast restated: to BOOLEAN is `v != 0` (256 is TRUE: no truncation through a byte),
    DECIMAL(p,s) -> DOUBLE divides by 10^s, INT -> DECIMAL(p,s) multiplies by 10^s and yields NULL when p digits do not hold it."""
    ints = np.array([0, 256, -1, 65536, 7], dtype=np.int32)
    dec = np.array([12345, -250, 0, 99999, 100], dtype=np.int64)     # DECIMAL(7,2): 123.45, -2.50, 0, 999.99, 1.00
    batch = build_batch(5, [("i", T.INT, False), ("d", T.DECIMAL, False)], {"i": ints, "d": dec}, {})
    b = PlanBuilder()
    ci, cd = b.col(T.INT, 0, False), b.col(T.DECIMAL, 1, False, scale=2, precision=7)
    b.filter(ci.cast(T.BOOLEAN))
    b.count().sum(cd.cast(T.DOUBLE)).sum(ci.cast(T.DECIMAL, 6, 2)).count(ci.cast(T.DECIMAL, 6, 2))
    (cnt, sd, sdec, cdec), = _run(b.build(), [], [batch])[0]
    assert cnt == 4                                   # 256, -1, 65536, 7 are TRUE
    assert abs(sd - (-2.5 + 0.0 + 999.99 + 1.0)) < 1e-9
    assert cdec == 3 and sdec == (256 - 1 + 7) * 100  # 65536 needs 5 integral digits: DECIMAL(6,2) holds 4 -> NULL
    import pytest
    from snappydata_b200.capi import SdError
    for bad in (lambda b, ci, cd: cd.cast(T.INT), lambda b, ci, cd: ci.cast(T.DATE), lambda b, ci, cd: cd.cast(T.DECIMAL, 7, 1)):
        b = PlanBuilder()
        ci, cd = b.col(T.INT, 0, False), b.col(T.DECIMAL, 1, False, scale=2, precision=7)
        b.count(bad(b, ci, cd))
        with pytest.raises(SdError):
            _run(b.build(), [], [batch])


def case_sha_null_key_bit_masking_1_to_32_key_columns(_run):
    """SHAByteBufferTest.scala:407-528 ("null grouping key bit masking for 1 to 100 columns in group by"), second half, for
    n = 1..32 key columns (this engine's limit is 32 grouping keys): 100 rows (1, 2, 'col-3'..'col-{n+2}') and 100 rows
    (4, 8, the same keys but the n-th NULL) -> exactly two groups, sums 100 / 200 and 400 / 800, the group with the NULL
    last key carrying 400."""
    rows_per = 100
    for nkeys in (1, 2, 3, 4, 5, 8, 17, 32):
        names = ["col%d" % j for j in range(3, nkeys + 3)]
        schema = [("num1", T.INT, True), ("num2", T.INT, True)] + [(nm, T.STRING, True) for nm in names]
        data = {"num1": np.array([1] * rows_per + [4] * rows_per, dtype=np.int32),
                "num2": np.array([2] * rows_per + [8] * rows_per, dtype=np.int32)}
        nulls = {}
        for j, nm in enumerate(names):
            data[nm] = np.array([b"col-%d" % (j + 3)] * (2 * rows_per), dtype=object)
            nulls[nm] = np.array([False] * rows_per + [j == nkeys - 1] * rows_per)
        batch = build_batch(2 * rows_per, schema, data, nulls)
        b = PlanBuilder()
        n1, n2 = b.col(T.INT, 0, True), b.col(T.INT, 1, True)
        keys = [b.col(T.STRING, 2 + j, True) for j in range(nkeys)]
        b.group_by(*keys)
        b.sum(n1).sum(n2)
        rows, _ = _run(b.build(), [], [batch])
        assert len(rows) == 2, (nkeys, rows)
        found_null = False
        for r in rows:
            ks, (s1, s2) = r[:nkeys], r[nkeys:]
            assert ks[:-1] == [b"col-%d" % (j + 3) for j in range(nkeys - 1)]
            if ks[-1] is None:
                found_null = True
                assert (s1, s2) == (rows_per * 4, rows_per * 8)
            else:
                assert ks[-1] == b"col-%d" % (nkeys + 2) and (s1, s2) == (rows_per * 1, rows_per * 2)
        assert found_null


def case_min_max_of_strings(_run):
    """MIN / MAX over a STRING column (the aggregate buffer is not fixed-width: the reference's ObjectHashSet path,
    SnappyHashAggregateExec.scala:82-94): 300 rows i, s = 's%03d' % i (NULL when i % 7 == 0), key = i % 4: per key the smallest /
    largest surviving i; one batch dictionary-encoded, one with the strings as an Uncompressed variable-width body."""
    n = 300
    i = np.arange(n)
    s = np.array([b"s%03d" % x for x in i], dtype=object)
    nulls = {"s": (i % 7) == 0}
    schema = [("k", T.INT, False), ("s", T.STRING, True)]
    half = n // 2
    batches = [build_batch(half, schema, {"k": (i[:half] % 4).astype(np.int32), "s": s[:half]}, {"s": nulls["s"][:half]}, batch_id=0),
               build_batch(n - half, schema, {"k": (i[half:] % 4).astype(np.int32), "s": s[half:]}, {"s": nulls["s"][half:]}, batch_id=1,
                           encoders={"s": "uncompressed"})]
    b = PlanBuilder()
    k, sc = b.col(T.INT, 0, False), b.col(T.STRING, 1, True)
    b.group_by(k)
    b.min(sc).max(sc).count(sc)
    rows, _ = _run(b.build(), [], batches)
    assert len(rows) == 4
    for key, mn, mx, cnt in rows:
        alive = [x for x in range(n) if x % 4 == key and x % 7 != 0]
        assert (mn, mx, cnt) == (b"s%03d" % min(alive), b"s%03d" % max(alive), len(alive))
    b = PlanBuilder()   # no key; and a filter that leaves nothing: NULL results
    k, sc = b.col(T.INT, 0, False), b.col(T.STRING, 1, True)
    b.filter(k > b.lit(T.INT))
    b.min(sc).max(sc)
    assert _run(b.build(), [-1], batches)[0] == [[b"s001", b"s299"]]
    assert _run(b.build(), [99], batches)[0] == [[None, None]]


def case_projection_returns_the_stored_values(_run):
    """Filter + project with no aggregate (ColumnTableScan feeding the row consumer directly): every projected field is the
    stored value -- DECIMAL(18,2) and a nullable DECIMAL(12,3) as their unscaled longs (UnsafeRow.setDecimal for precision <=
    18), FLOAT, BOOLEAN, LONG, a dictionary STRING and an Uncompressed one -- checked against the INPUT arrays, not against
    another engine."""
    from snappydata_b200 import capi
    n = 700
    i = np.arange(n)
    data = {"d": (i * 1234567 - 400_000_000).astype(np.int64), "e": ((i % 97) * 1001 - 5000).astype(np.int64), "l": (i * i - 1000).astype(np.int64),
            "f": (i % 13).astype(np.float32) / 4, "b": (i % 3 == 0), "s": np.array([b"v%d" % (x % 11) for x in i], dtype=object),
            "t": np.array([b"raw-%05d" % x for x in i], dtype=object)}
    nulls = {"e": i % 5 == 0, "b": i % 7 == 0}
    schema = [("d", T.DECIMAL, False), ("e", T.DECIMAL, True), ("l", T.LONG, False), ("f", T.FLOAT, False), ("b", T.BOOLEAN, True),
              ("s", T.STRING, False), ("t", T.STRING, False)]
    batches = [build_batch(n, schema, data, nulls, batch_id=0, encoders={"t": "uncompressed"})]
    b = PlanBuilder()
    c = [b.col(T.DECIMAL, 0, False, scale=2, precision=18), b.col(T.DECIMAL, 1, True, scale=3, precision=12), b.col(T.LONG, 2, False),
         b.col(T.FLOAT, 3, False), b.col(T.BOOLEAN, 4, True), b.col(T.STRING, 5, False), b.col(T.STRING, 6, False)]
    b.filter(c[2] >= b.lit(T.LONG))
    b.project(*c)
    pl = capi.Plan(_run.api, b.build()).set_literals([24])     # l >= 24  <=>  i >= 32
    for x in batches:
        pl.submit(x)
    rows = pl.finish()
    want = [[int(data["d"][x]), None if nulls["e"][x] else int(data["e"][x]), int(data["l"][x]), float(data["f"][x]),
             None if nulls["b"][x] else bool(data["b"][x]), bytes(data["s"][x]), bytes(data["t"][x])] for x in range(32, n)]
    key = lambda r: r[2]
    assert sorted(rows, key=key) == sorted(want, key=key)


CASES = [case_sha_one_nullable_string_key_closed_form, case_sha_two_nullable_string_keys_closed_form,
         case_delta_stats_point_filters_after_updates, case_basic_delete_and_update_counts,
         case_sha_sum_of_every_numeric_type_per_string_key, case_sha_decimal_sum_and_avg_per_string_key, case_casts_follow_spark,
         case_sha_null_key_bit_masking_1_to_32_key_columns, case_min_max_of_strings, case_projection_returns_the_stored_values]
