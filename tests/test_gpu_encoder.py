"""ColumnBatch creation on the device (sd_store_encode_batch): the buffers the GPU encoder leaves in the
store are, byte for byte, the ones the fixture writer (snappydata_b200/column_format.py -- the restatement of the reference's
ColumnEncoders, enc/ColumnEncoding.scala:177-736, Uncompressed.scala:228-448, DictionaryEncoding.scala:168-450,
BooleanBitSetEncoding.scala:62-152) writes for the same values, the stats row likewise (ColumnInsertExec.scala:848-921), and a
scan over device-encoded batches equals the oracle's scan over the fixture writer's bytes."""
import numpy as np
import pytest

from oracle import oracle
from snappydata_b200 import capi
from snappydata_b200.column_format import SqlType as T, build_batch
from snappydata_b200.plan import PlanBuilder

from helpers import assert_rowsets_match

pytestmark = pytest.mark.gpu

SCHEMA = [("i", T.INT, True), ("l", T.LONG, False), ("d", T.DOUBLE, True), ("s", T.STRING, True), ("b", T.BOOLEAN, True),
          ("dt", T.DATE, False), ("h", T.SHORT, False), ("f", T.FLOAT, True), ("y", T.BYTE, False), ("ts", T.TIMESTAMP, True),
          ("k", T.STRING, False), ("dec", T.DECIMAL, False)]


def table(n, seed, distinct_strings=12):
    r = np.random.default_rng(seed)
    data = {"i": r.integers(-1000, 1000, n).astype(np.int32), "l": r.integers(-2**40, 2**40, n).astype(np.int64),
            "d": np.round(r.normal(0, 100, n), 3), "s": np.array([b"v%d" % x for x in r.integers(0, distinct_strings, n)], dtype=object),
            "b": r.integers(0, 2, n).astype(bool), "dt": (9000 + r.integers(0, 40, n)).astype(np.int32),
            "h": r.integers(-5, 5, n).astype(np.int16), "f": r.normal(0, 10, n).astype(np.float32),
            "y": r.integers(-128, 128, n).astype(np.int8), "ts": r.integers(0, 10**12, n).astype(np.int64),
            "k": np.array([b"key-%d" % (x % 7) for x in range(n)], dtype=object), "dec": r.integers(-10**10, 10**10, n).astype(np.int64)}
    nulls = {}
    for name, _, nullable in SCHEMA:
        if nullable:
            m = r.random(n) < 0.1
            if n > 300:
                m[n - 200:] = False      # trailing non-null stretch: null words get trimmed
            nulls[name] = m
    return data, nulls


@pytest.mark.parametrize("n,seed,nd", [(5000, 1, 12), (1, 2, 3), (64, 3, 5), (2049, 4, 12), (70000, 5, 60000)])
def test_device_encoder_is_byte_identical_to_the_fixture_writer(gpu_api, n, seed, nd):
    data, nulls = table(n, seed, nd)
    want = build_batch(n, SCHEMA, data, nulls, batch_id=7, bucket_id=3)
    store = capi.Store(gpu_api, [(t, nl) for _, t, nl in SCHEMA], 0)
    raw = {c: (data[name], nulls.get(name)) for c, (name, _, _) in enumerate(SCHEMA)}
    store.encode_batch(n, raw, bucket_id=3, batch_id=7)
    assert store.num_batches() == 1 and store.batch_info(0) == (n, 3, 7)
    for c, (name, t, _) in enumerate(SCHEMA):
        got = store.get_buffer(0, c)
        assert got == bytes(want.columns[c]), (name, len(got), len(want.columns[c]))
    assert store.get_stats(0) == bytes(want.stats)
    if nd > 32767:   # the dictionary switched to int32 indexes
        assert int.from_bytes(store.get_buffer(0, 3)[:4], "little") == 3


def test_scan_over_device_encoded_batches_equals_oracle(gpu_api):
    store = capi.Store(gpu_api, [(t, nl) for _, t, nl in SCHEMA], 0)
    batches = []
    for bid, (n, seed) in enumerate(((5000, 11), (777, 12), (1, 13), (30000, 14))):
        data, nulls = table(n, seed)
        batches.append(build_batch(n, SCHEMA, data, nulls, batch_id=bid, bucket_id=bid % 2))
        store.encode_batch(n, {c: (data[name], nulls.get(name)) for c, (name, _, _) in enumerate(SCHEMA)}, bucket_id=bid % 2, batch_id=bid)
    b = PlanBuilder()
    c = {name: b.col(t, i, nl) for i, (name, t, nl) in enumerate(SCHEMA)}
    b.filter((c["i"] > b.lit(T.INT)) & c["s"].ne(b.lit(T.STRING)))
    b.group_by(c["k"], c["b"])
    b.count().sum(c["d"]).sum(c["l"]).min(c["f"]).max(c["dec"]).count(c["ts"])
    desc = b.build()
    op = oracle.plan(desc).set_literals([-200, b"v3"])
    for x in batches:
        op.submit(x)
    gp = capi.Plan(gpu_api, desc).set_literals([-200, b"v3"])
    gp.scan_store(store)
    assert_rowsets_match(gp.finish(), op.finish(), 2)
    # the encoder's stats rows drive batch skipping like the writer's do
    b = PlanBuilder()
    c = {name: b.col(t, i, nl) for i, (name, t, nl) in enumerate(SCHEMA)}
    b.filter(c["dt"] > b.lit(T.DATE))
    b.count()
    gp = capi.Plan(gpu_api, b.build()).set_literals([9999])
    gp.scan_store(store)
    assert gp.finish() == [[0]] and gp.metrics()["columnBatchesSkipped"] == 4
