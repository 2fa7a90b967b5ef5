"""GPU parity over the *general* decode path: every registered encoding (Uncompressed, Dictionary,
BigDictionary, BooleanBitSet, RunLength), nullable columns with trimmed null words, update deltas of
depth 0/1 (depth 0 wins), delete masks, row-buffer rows, string predicates, NULL group keys, all
aggregate functions.  Plans here have no ahead-of-time kernel: they go through the NVRTC path.
Mirrors what the reference pins with ColumnEncodersTest / ColumnTablesTestBase.runAllTypesTest
(round trip of 13 types x {nullable, not null}), SHAByteBufferTest (group-by with NULL keys) and
ColumnUpdateDeleteTests (deltas + deletes) -- SURVEY.md 4, 8c."""
import numpy as np
import pytest

from oracle import oracle
from snappydata_b200 import capi
from snappydata_b200.column_format import (ColumnBatch, SqlType, build_batch, encode_delete, encode_delta, unsafe_row)
from snappydata_b200.plan import PlanBuilder

from helpers import assert_rowsets_match

pytestmark = pytest.mark.gpu

T = SqlType
SCHEMA = [("c0", T.INT, True), ("c1", T.LONG, False), ("c2", T.DOUBLE, True), ("c3", T.STRING, True),
          ("c4", T.BOOLEAN, True), ("c5", T.DATE, False), ("c6", T.SHORT, False), ("c7", T.FLOAT, True),
          ("c8", T.BYTE, False), ("c9", T.TIMESTAMP, True), ("c10", T.STRING, False), ("c11", T.DECIMAL, False)]
ENCODERS = {"c5": "dictionary", "c6": "rle", "c9": "rle", "c10": "rle"}


def make_batch(n, seed, batch_id=0, null_frac=0.1, encoders=ENCODERS, big_dict=False):
    r = np.random.default_rng(seed)
    data = {
        "c0": r.integers(-1000, 1000, n).astype(np.int32),
        "c1": r.integers(-2**40, 2**40, n).astype(np.int64),
        "c2": np.round(r.normal(0, 100, n), 3),
        "c3": np.array([b"v%d" % x for x in r.integers(0, 12, n)], dtype=object),
        "c4": r.integers(0, 2, n).astype(bool),
        "c5": (9000 + r.integers(0, 40, n)).astype(np.int32),
        "c6": np.repeat(r.integers(-5, 5, n // 7 + 1), 7)[:n].astype(np.int16),
        "c7": r.normal(0, 10, n).astype(np.float32),
        "c8": r.integers(-128, 128, n).astype(np.int8),
        "c9": np.repeat(r.integers(0, 10**12, n // 5 + 1), 5)[:n].astype(np.int64),
        "c10": np.array([b"k%d" % (x // 9) for x in range(n)], dtype=object),
        "c11": r.integers(-10**10, 10**10, n).astype(np.int64),
    }
    # a few special doubles: NaN, -0.0, +-inf exercise the NaN-safe order
    if n > 50:
        data["c2"][r.integers(0, n, 3)] = np.nan
        data["c2"][r.integers(0, n, 2)] = -0.0
        data["c2"][r.integers(0, n, 1)] = np.inf
    nulls = {}
    for name, _, nullable in SCHEMA:
        if nullable:
            m = r.random(n) < null_frac
            if n > 300:
                m[n - 200:] = False          # trailing non-null stretch -> null words get trimmed
            nulls[name] = m
    enc = dict(encoders)
    if big_dict:
        enc["c3"] = "bigdictionary"
    b = build_batch(n, SCHEMA, data, nulls, batch_id=batch_id, bucket_id=batch_id % 4, encoders=enc)
    return b, data, nulls


def cols(b: PlanBuilder):
    return {name: b.col(t, i, nullable) for i, (name, t, nullable) in enumerate(SCHEMA)}


def both(gpu_api, desc, lits, batches, nkeys, rows=None):
    gp = capi.Plan(gpu_api, desc).set_literals(lits)
    assert gp.kernel_name().startswith(("jit:", "aot:"))
    op = oracle.plan(desc).set_literals(lits)
    if rows is not None:
        gp.submit_rows(*rows)
        op.submit_rows(*rows)
    for x in batches:
        gp.submit(x)
        op.submit(x)
    got, want = gp.finish(), op.finish()
    assert_rowsets_match(got, want, nkeys)
    return gp, op, got


@pytest.fixture(scope="module")
def batches():
    return [make_batch(n, seed=10 + i, batch_id=i, big_dict=(i == 2))[0] for i, n in enumerate((5000, 1, 777, 2049, 64))]


def test_no_key_all_aggregates_all_encodings(gpu_api, batches):
    b = PlanBuilder()
    c = cols(b)
    b.filter(((c["c0"] > b.lit(T.INT)) | c["c2"].is_null()) & (c["c5"] >= b.lit(T.DATE)))
    b.count().count(c["c0"]).sum(c["c0"]).sum(c["c2"]).avg(c["c0"]).avg(c["c7"]).min(c["c2"]).max(c["c2"])
    b.min(c["c0"]).max(c["c1"]).sum(c["c6"]).min(c["c9"]).max(c["c9"]).sum(c["c8"]).count(c["c4"]).min(c["c11"]).max(c["c5"])
    b.sum(c["c7"]).max(c["c7"]).min(c["c8"])
    both(gpu_api, b.build(), [-500, 9005], batches, 0)


def test_boolean_and_arithmetic_predicates(gpu_api, batches):
    b = PlanBuilder()
    c = cols(b)
    expr = (c["c2"] * b.lit(T.DOUBLE) + c["c7"].cast(T.DOUBLE)) / c["c2"]
    b.filter((c["c4"].eq(b.lit(T.BOOLEAN)) & ~(c["c1"] < b.lit(T.LONG))) | (c["c6"].cast(T.INT) + c["c0"] > b.lit(T.INT)))
    b.sum(expr).count(expr).sum((c["c0"] * c["c0"]).cast(T.LONG) - c["c1"]).max(-c["c2"]).count(c["c0"].isin(3))
    both(gpu_api, b.build(), [2.5, True, 0, 900, 7, None, 13], batches, 0)


def test_group_by_nullable_string_key(gpu_api, batches):
    """NULL is a group of its own (SHAByteBufferTest.scala:225-267)."""
    b = PlanBuilder()
    c = cols(b)
    b.filter(c["c0"].is_not_null())
    b.group_by(c["c3"])
    b.sum(c["c1"]).avg(c["c2"]).count(c["c7"]).min(c["c7"]).max(c["c8"]).count().sum(c["c0"])
    _, _, got = both(gpu_api, b.build(), [], batches, 1)
    assert any(r[0] is None for r in got) and len(got) == 13


def test_group_by_two_string_keys_including_rle_string(gpu_api, batches):
    b = PlanBuilder()
    c = cols(b)
    b.group_by(c["c3"], c["c10"])
    b.count().sum(c["c2"]).max(c["c0"])
    both(gpu_api, b.build(), [], batches[2:5], 2)


@pytest.mark.parametrize("which", ["eq", "ne", "lt", "ge", "in", "startswith", "rle_lt", "null_lit"])
def test_string_predicates_via_dictionary_truth_tables(gpu_api, batches, which):
    b = PlanBuilder()
    c = cols(b)
    lits = []
    if which == "eq":
        f = c["c3"].eq(b.lit(T.STRING)); lits = [b"v3"]
    elif which == "ne":
        f = c["c3"].ne(b.lit(T.STRING)); lits = [b"v3"]
    elif which == "lt":
        f = c["c3"] < b.lit(T.STRING); lits = [b"v5"]
    elif which == "ge":
        f = b.lit(T.STRING) <= c["c3"]; lits = [b"v10"]
    elif which == "in":
        f = c["c3"].isin(3); lits = [b"v1", b"v11", b"nope"]
    elif which == "startswith":
        f = c["c3"].startswith(b.lit(T.STRING)); lits = [b"v1"]
    elif which == "rle_lt":
        f = c["c10"] < b.lit(T.STRING); lits = [b"k3"]
    else:
        f = c["c3"].eq(b.lit(T.STRING)); lits = [None]
    b.filter(f)
    b.count().sum(c["c1"])
    both(gpu_api, b.build(), lits, batches, 0)


def with_deltas_and_deletes(n, seed):
    base, data, nulls = make_batch(n, seed, encoders={})
    r = np.random.default_rng(seed + 1)
    for col_name, t in (("c0", T.INT), ("c2", T.DOUBLE), ("c3", T.STRING), ("c4", T.BOOLEAN), ("c1", T.LONG)):
        ci = [s[0] for s in SCHEMA].index(col_name)
        nullable = SCHEMA[ci][2]
        p0 = np.sort(r.choice(n, size=min(n, 40), replace=False)).astype(np.int32)
        p1 = np.sort(np.unique(np.concatenate([r.choice(n, size=min(n, 300), replace=False), p0[:10]]))).astype(np.int32)
        for depth, pos in ((0, p0), (1, p1)):
            m = len(pos)
            if t == T.STRING:
                vals = np.array([b"u%d" % x for x in r.integers(0, 5, m)] , dtype=object)
                vals[: m // 3] = np.array([b"v%d" % x for x in r.integers(0, 12, m // 3)], dtype=object)
            elif t == T.BOOLEAN:
                vals = r.integers(0, 2, m).astype(bool)
            elif t == T.DOUBLE:
                vals = np.round(r.normal(1000, 5, m), 2)
            else:
                vals = r.integers(5000, 6000, m)
            dn = (r.random(m) < 0.2) if nullable else None
            buf = encode_delta(n, pos, vals, t, dn)
            (base.delta0 if depth == 0 else base.delta1)[ci] = buf
    dels = np.sort(r.choice(n, size=max(1, n // 50), replace=False)).astype(np.int32)
    base.delete_mask = encode_delete(n, dels)
    return base


@pytest.mark.parametrize("n", [3000, 70, 1025])
def test_update_deltas_and_delete_mask(gpu_api, n):
    """Effective value: depth-0 delta, else depth-1 delta, else base; deleted ordinals skipped
    (ColumnTableScan.scala:757-786, enc/UpdatedColumnDecoder.scala:85-128, enc/ColumnDeleteDecoder.scala:49-55)."""
    bs = [with_deltas_and_deletes(n, 100 + i) for i in range(2)] + [make_batch(500, 7, encoders={})[0]]
    b = PlanBuilder()
    c = cols(b)
    b.filter(c["c0"].is_null() | (c["c0"] > b.lit(T.INT)))
    b.group_by(c["c3"])
    b.count().sum(c["c0"]).sum(c["c2"]).count(c["c4"]).sum(c["c1"]).min(c["c2"]).max(c["c0"])
    gp, op, _ = both(gpu_api, b.build(), [-900], bs, 1)
    gm, om = gp.metrics(), op.metrics()
    for k in ("rowsScanned", "deletedBatchCount", "updatedColumnCount", "columnBatchesSeen"):
        assert gm[k] == om[k], (k, gm, om)
    # no-key variant over the boolean + string predicate path
    b = PlanBuilder()
    c = cols(b)
    b.filter(c["c3"].startswith(b.lit(T.STRING)) | c["c4"].eq(b.lit(T.BOOLEAN)))
    b.count().sum(c["c2"]).count(c["c3"])
    both(gpu_api, b.build(), [b"u", True], bs, 0)


def test_row_buffer_rows_join_the_same_aggregate(gpu_api, batches):
    """Hybrid scan: row-buffer rows are consumed first through the same loop (ColumnTableScan.scala:572-588)."""
    b = PlanBuilder()
    c0 = b.col(T.INT, 0, True)
    c2 = b.col(T.DOUBLE, 2, True)
    c3 = b.col(T.STRING, 3, True)
    c1 = b.col(T.LONG, 1, False)
    b.filter(c1 > b.lit(T.LONG))
    b.group_by(c3)
    b.count().sum(c0).avg(c2)
    r = np.random.default_rng(5)
    rows = b""
    nrows = 257
    for i in range(nrows):
        row = unsafe_row([(T.INT, None if i % 7 == 0 else int(r.integers(-50, 50))), (T.DOUBLE, None if i % 5 == 0 else float(r.normal())),
                          (T.STRING, None if i % 11 == 0 else b"v%d" % r.integers(0, 15)), (T.LONG, int(r.integers(-100, 100)))])
        rows += len(row).to_bytes(8, "little") + row
    gp, op, _ = both(gpu_api, b.build(), [-20], batches, 1, rows=(rows, nrows))
    assert gp.metrics()["numRowsBuffer"] == nrows == op.metrics()["numRowsBuffer"]


def test_errors_not_fallbacks(gpu_api):
    """A buffer the GPU path cannot scan is an error, never a CPU fallback."""
    from snappydata_b200.column_format import encode_uncompressed, compress_lz4
    # (Uncompressed variable-width STRING columns used to be refused here: they are scanned now, tests/test_gpu_strings.py)
    # a corrupt variable-width body is an error
    b = PlanBuilder()
    s = b.col(T.STRING, 0, False)
    b.filter(s.eq(b.lit(T.STRING)))
    b.count()
    gp = capi.Plan(gpu_api, b.build()).set_literals([b"x"])
    good = encode_uncompressed([b"a", b"b", b"c"], T.STRING)
    bad = ColumnBatch(num_rows=3, columns=[good[:-1] + b"\x7f"[:0]])   # last value truncated by one byte
    bad.columns[0] = good[:-1]
    with pytest.raises(capi.SdError) as e:
        gp.submit(bad)
        gp.finish()
    assert e.value.code == capi.SD_ERR_INVALID
    # NOT NULL column carrying a null bitset is rejected like NotNullDecoder does
    b = PlanBuilder()
    x = b.col(T.INT, 0, False)
    b.sum(x)
    gp = capi.Plan(gpu_api, b.build()).set_literals([])
    bad = ColumnBatch(num_rows=3, columns=[encode_uncompressed(np.array([1, 2, 3]), T.INT, np.array([False, True, False]))])
    with pytest.raises(capi.SdError) as e:
        gp.submit(bad)
    assert e.value.code == capi.SD_ERR_INVALID
    # update-delta / delete positions index shared-memory bitmaps in the kernel: they are validated when the batch is uploaded
    import struct
    okcol = encode_uncompressed(np.arange(10, dtype=np.int32), T.INT)
    bad_del = struct.pack("<iii", 0, 10, 2) + struct.pack("<ii", 7, 3)                     # descending
    with pytest.raises(capi.SdError) as e:
        gp.submit(ColumnBatch(num_rows=10, columns=[okcol], delete_mask=bad_del))
    assert e.value.code == capi.SD_ERR_INVALID
    bad_del = struct.pack("<iii", 0, 10, 1) + struct.pack("<i", 10)                        # beyond the batch
    with pytest.raises(capi.SdError) as e:
        gp.submit(ColumnBatch(num_rows=10, columns=[okcol], delete_mask=bad_del))
    assert e.value.code == capi.SD_ERR_INVALID
    good_delta = encode_delta(10, np.array([2, 5], dtype=np.int32), np.array([1, 2], dtype=np.int32), T.INT)
    bad_delta = bytearray(good_delta)
    bad_delta[16:24] = struct.pack("<ii", 5, 2)                                             # positions swapped
    with pytest.raises(capi.SdError) as e:
        gp.submit(ColumnBatch(num_rows=10, columns=[okcol], delta0={0: bytes(bad_delta)}))
    assert e.value.code == capi.SD_ERR_INVALID
    with pytest.raises(capi.SdError) as e:                                                  # values shorter than the entries
        gp.submit(ColumnBatch(num_rows=10, columns=[okcol], delta0={0: good_delta[:-4]}))
    assert e.value.code == capi.SD_ERR_INVALID
    # a corrupt Snappy envelope (codec 2) is an error (valid ones are decoded: tests/test_gpu_strings.py); unknown codecs are refused
    snappy_env = (-2).to_bytes(4, "little", signed=True) + (20000).to_bytes(4, "little") + b"\0" * 100
    with pytest.raises(capi.SdError) as e:
        gp.submit(ColumnBatch(num_rows=5000, columns=[snappy_env]))
    assert e.value.code == capi.SD_ERR_INVALID
    codec3 = (-3).to_bytes(4, "little", signed=True) + (20000).to_bytes(4, "little") + b"\0" * 100
    with pytest.raises(capi.SdError) as e:
        gp.submit(ColumnBatch(num_rows=5000, columns=[codec3]))
    assert e.value.code == capi.SD_ERR_UNSUPPORTED


def _compress_batch(b, force=True):
    from snappydata_b200.column_format import compress_lz4
    import copy
    c = copy.copy(b)
    c.columns = [None if x is None else compress_lz4(x, force=force) for x in b.columns]
    return c


def test_lz4_compressed_buffers_are_expanded_on_the_device(gpu_api, batches):
    """Stored form [-1][uncompressedLen][LZ4 block] (CompressionUtils.scala:53-61): same results as the
    uncompressed bytes, every encoding, nullable or not; fewer bytes cross PCIe."""
    b = PlanBuilder()
    c = cols(b)
    b.filter(c["c5"] >= b.lit(T.DATE))
    b.group_by(c["c3"])
    b.count().sum(c["c0"]).sum(c["c2"]).count(c["c4"]).sum(c["c1"]).min(c["c7"]).max(c["c8"]).max(c["c11"])
    desc = b.build()
    plain = [make_batch(n, seed=30 + i, batch_id=i, encoders={"c5": "dictionary"})[0] for i, n in enumerate((6000, 300, 2049))]
    op = oracle.plan(desc).set_literals([9001])
    gp = capi.Plan(gpu_api, desc).set_literals([9001])
    gplain = capi.Plan(gpu_api, desc).set_literals([9001])
    for x in plain:
        op.submit(x)
        gp.submit(_compress_batch(x))
        gplain.submit(x)
    want = op.finish()
    assert_rowsets_match(gp.finish(), want, 1)
    assert_rowsets_match(gplain.finish(), want, 1)
    assert gp.metrics()["h2dBytes"] < gplain.metrics()["h2dBytes"]


def test_lz4_device_decoder_is_byte_exact_on_hard_blocks(gpu_api):
    """The device LZ4 decoder against liblz4's own output on blocks that stress every copy path: runs (matches that
    overlap their output at offsets 1..9), long literal runs, long matches, matches farther back than the decoder's
    shared-memory ring, dense short sequences, a tiny buffer, several buffers per launch."""
    import numpy as np
    import struct
    from snappydata_b200.column_format import compress_lz4, decompress
    rng = np.random.default_rng(5)
    a30 = rng.bytes(30_000)
    bodies = [
        bytes(40_000),                                                             # one huge overlapping match (offset 1)
        b"".join(bytes([i + 1]) * (i + 1) * 700 for i in range(9)),                # runs of every small period
        b"".join((b"abcdefghi"[:k] * 3000)[:9000] for k in range(1, 10)),          # periods 1..9
        rng.bytes(50_000),                                                         # incompressible: one long literal run
        a30 + rng.bytes(25_000) + a30 + rng.bytes(100) + a30[5_000:9_000],         # long matches ~55 KB back (beyond the ring)
        rng.integers(1, 51, 60_000).astype(np.float64).tobytes(),                  # dense 8-byte sequences, near sources
        rng.integers(0, 3, 100_000).astype(np.int16).tobytes(),                    # short sequences, deep dependency chains
        rng.integers(8000, 10_500, 50_000).astype(np.int32).tobytes(),
        b"".join(rng.bytes(int(n)) + bytes(int(m)) for n, m in zip(rng.integers(0, 80, 400), rng.integers(4, 300, 400))),
        rng.bytes(13) + bytes(27),                                                  # tiny
    ]
    cols = []
    for body in bodies:
        body = body + bytes(-len(body) % 8)
        plain = struct.pack("<ii", 0, 0) + body                                     # Uncompressed LONG column, no nulls
        env = compress_lz4(plain, force=True)
        assert decompress(env) == plain
        cols.append((plain, env, len(body) // 8))
    schema = [(T.LONG, False)]
    store = capi.Store(gpu_api, schema)
    for i, (plain, env, n) in enumerate(cols):
        store.put(ColumnBatch(num_rows=n, columns=[env], batch_id=i))
    for i, (plain, env, n) in enumerate(cols):
        got = store.get_buffer(i, 0)
        assert got == plain, f"block {i}: first difference at byte {next((k for k in range(len(plain)) if got[k] != plain[k]), -1)}"


def test_store_that_grows_between_executions_is_scanned_incrementally(gpu_api, monkeypatch):
    """Queries over a store that ingest keeps appending to (BASELINE.json's hybrid configuration): the cached scan keeps the
    descriptors it has and adds a segment for the new batches (sd_plan_scan_store), folding small segments together now and
    then.  Every execution against the oracle over exactly the batches present: dense group-by with string keys whose
    dictionaries grow with the new batches, a hash group-by, and a projection (batch ordinals continue across segments)."""
    store = capi.Store(gpu_api, [(t, nl) for _, t, nl in SCHEMA], 0)
    b1 = PlanBuilder(); c = cols(b1)
    b1.filter(c["c5"] >= b1.lit(T.DATE)); b1.group_by(c["c3"], c["c10"]); b1.count().sum(c["c2"]).max(c["c1"])
    b2 = PlanBuilder(); c = cols(b2)
    b2.group_by(c["c0"]); b2.count().sum(c["c1"])
    b3 = PlanBuilder(); c = cols(b3)
    b3.filter(c["c0"] > b3.lit(T.INT)); b3.project(c["c0"], c["c3"], c["c2"], c["c10"])
    plans = [(b1.build(), [9005], 2), (b2.build(), [], 1), (b3.build(), [700], None)]
    gps = [capi.Plan(gpu_api, d) for d, _, _ in plans]
    present = []
    launches = []
    for step in range(12):
        for k in range(1 if step else 3):
            x = make_batch(900 + 37 * step + k, seed=100 + 10 * step + k, batch_id=len(present))[0]
            store.put(x)
            present.append(x)
        for gp, (desc, lits, nk) in zip(gps, plans):
            for rep in range(2):   # the second execution finds the cache current
                gp.reset().set_literals(lits)
                gp.scan_store(store)
                got = gp.finish()
                op = oracle.plan(desc).set_literals(lits)
                for x in present:
                    op.submit(x)
                want = op.finish()
                if nk is None:
                    assert sorted(map(_rowkey2, got)) == sorted(map(_rowkey2, want))
                else:
                    assert_rowsets_match(got, want, nk)
            if nk == 2:
                launches.append(gp.metrics()["kernelLaunches"])
    assert max(launches) >= 3 and launches[-1] <= 6, launches      # segments were added, and folded again
    # the same sequence with the incremental path off gives one launch per execution
    monkeypatch.setenv("SD_TUNE_NO_INCREMENTAL_SCAN", "1")
    store.put(make_batch(500, seed=999, batch_id=len(present))[0])
    gp = gps[0]
    gp.reset().set_literals([9005]); gp.scan_store(store); gp.finish()
    assert gp.metrics()["kernelLaunches"] == 1
    store.close()


def _rowkey2(r):
    import struct
    return tuple((0, b"") if v is None else (1, struct.pack("<d", v)) if isinstance(v, float) else (2, v) if isinstance(v, bytes)
                 else (3, struct.pack("<q", int(v))) for v in r)


@pytest.mark.parametrize("copy_streams", [1, 3])
def test_lz4_lineitem_q1_q6(gpu_api, copy_streams, monkeypatch):
    """(copy_streams > 1: SD_TUNE_COPY_STREAMS -- the batches' compressed spans rotate over several H2D queues and the expansion
    waits for all of them)"""
    from snappydata_b200 import lineitem, plan as P
    if copy_streams > 1:
        monkeypatch.setenv("SD_TUNE_COPY_STREAMS", str(copy_streams))
    else:
        monkeypatch.delenv("SD_TUNE_COPY_STREAMS", raising=False)
    plain = lineitem.gen_table(260_001, 65_000, seed=21)
    for desc, lits, nk in ((P.q6_plan(), P.Q6_LITERALS, 0), (P.q1_plan(), P.Q1_LITERALS, 2)):
        op = oracle.plan(desc).set_literals(lits)
        gp = capi.Plan(gpu_api, desc).set_literals(lits)
        for x in plain:
            op.submit(x)
            gp.submit(_compress_batch(x, force=False))   # the reference's rule: only if it shrinks to <= 75 %
        assert_rowsets_match(gp.finish(), op.finish(), nk)
        m = gp.metrics()
        assert m["h2dBytes"] < 0.8 * m["algorithmicBytes"]


# ---- MODE_HASH: general (non dictionary-string) group keys -------------------------------------------
def test_group_by_nullable_int_key_hash_table(gpu_api, batches):
    """Integer key with NULLs (SHAByteBufferTest.scala:225-331 shapes): the device hash table."""
    b = PlanBuilder()
    c = cols(b)
    b.filter(c["c5"] >= b.lit(T.DATE))
    b.group_by(c["c0"])
    b.count().sum(c["c1"]).avg(c["c2"]).min(c["c7"]).max(c["c9"])
    gp, _, got = both(gpu_api, b.build(), [9003], batches, 1)
    assert any(r[0] is None for r in got) and len(got) > 1500


def test_group_by_composite_keys_int_string_date_bool(gpu_api, batches):
    b = PlanBuilder()
    c = cols(b)
    b.group_by(c["c6"], c["c3"], c["c5"], c["c4"])
    b.count().sum(c["c2"]).max(c["c0"]).sum(c["c8"])
    both(gpu_api, b.build(), [], batches, 4)


def test_group_by_expression_key_and_long_key(gpu_api, batches):
    b = PlanBuilder()
    c = cols(b)
    b.group_by((c["c0"] + c["c6"].cast(T.INT)), c["c11"])
    b.count().sum(c["c1"])
    both(gpu_api, b.build(), [], batches[:1], 2)


def test_high_cardinality_group_by_grows_the_hash_table(gpu_api):
    """More distinct keys than the initial table holds: the engine grows the table and replays the launches."""
    r = np.random.default_rng(3)
    n = 120_000
    schema = [("k", T.LONG, False), ("v", T.DOUBLE, False), ("w", T.INT, True)]
    bs = []
    for i in range(3):
        data = {"k": r.integers(0, 90_000, n).astype(np.int64) * 7919 - 10**9, "v": r.normal(0, 1, n), "w": r.integers(0, 5, n).astype(np.int32)}
        bs.append(build_batch(n, schema, data, {"w": r.random(n) < 0.3}, batch_id=i))
    b = PlanBuilder()
    k = b.col(T.LONG, 0, False)
    v = b.col(T.DOUBLE, 1, False)
    w = b.col(T.INT, 2, True)
    b.group_by(k)
    b.count().sum(v).count(w).min(w)
    gp, op, got = both(gpu_api, b.build(), [], bs, 1)
    assert len(got) > 80_000
    # and the store-resident path re-executes (cached descriptors, table re-initialised) with the same answer
    gp.reset().set_literals([])
    for x in bs:
        gp.submit(x)
    assert_rowsets_match(gp.finish(), got, 1)


# ---- MODE_PROJECT: filter + project, no aggregate (BASELINE.json configs[3]) -----------
def _rowkey(r):
    return tuple((0, 0) if v is None else (1, v) if isinstance(v, bytes) else (2, float(v)) for v in r)


def make_wide_batch(n, seed, ncols=16, batch_id=0):
    """c0..c{ncols-1} cycling (INT, DOUBLE, dictionary STRING with 1000 distinct 8-12 byte values); every 4th nullable."""
    r = np.random.default_rng(seed)
    schema, data, nulls = [], {}, {}
    for i in range(ncols):
        t = (T.INT, T.DOUBLE, T.STRING)[i % 3]
        nullable = i % 4 == 0
        name = f"c{i}"
        schema.append((name, t, nullable))
        if t == T.INT:
            data[name] = r.integers(0, 1000, n).astype(np.int32)
        elif t == T.DOUBLE:
            data[name] = r.random(n) * 100.0
        else:
            data[name] = np.array([b"str%05d_" % x + b"x" * (x % 5) for x in r.integers(0, 1000, n)], dtype=object)
        if nullable:
            nulls[name] = r.random(n) < 0.1
    return build_batch(n, schema, data, nulls, batch_id=batch_id), schema


def test_projection_with_filter_wide_table(gpu_api):
    bs, schema = [], None
    for i, n in enumerate((20_000, 3_001, 1)):
        b, schema = make_wide_batch(n, 50 + i, batch_id=i)
        bs.append(b)
    pb = PlanBuilder()
    c = [pb.col(t, i, nullable) for i, (_, t, nullable) in enumerate(schema[:8])]
    pb.filter((c[0] >= pb.lit(T.INT)) & (c[0] <= pb.lit(T.INT)) & c[2].eq(pb.lit(T.STRING)))
    pb.project(*c)
    desc = pb.build()
    lit = bs[0]  # pick a string literal that exists
    from snappydata_b200.column_format import decode_column
    some = decode_column(bs[0].columns[2], T.STRING, bs[0].num_rows)[0][5]
    for lits in ([0, 999, some], [100, 300, some], [0, 999, b"absent"]):
        gp = capi.Plan(gpu_api, desc).set_literals(lits)
        op = oracle.plan(desc).set_literals(lits)
        for x in bs:
            gp.submit(x)
            op.submit(x)
        got, want = sorted(gp.finish(), key=_rowkey), sorted(op.finish(), key=_rowkey)
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert g == w
    # a looser filter: thousands of rows, NULLs in projected columns, projected expression
    pb = PlanBuilder()
    c = [pb.col(t, i, nullable) for i, (_, t, nullable) in enumerate(schema[:8])]
    pb.filter(c[0].is_null() | (c[3] < pb.lit(T.INT)))
    pb.project(c[0], c[1] * c[1], c[2], c[5], c[4], (c[3] + c[6]).cast(T.LONG))
    desc = pb.build()
    gp = capi.Plan(gpu_api, desc).set_literals([200])
    op = oracle.plan(desc).set_literals([200])
    for x in bs:
        gp.submit(x)
        op.submit(x)
    got, want = sorted(gp.finish(), key=_rowkey), sorted(op.finish(), key=_rowkey)
    assert len(got) == len(want) > 4000
    assert got == want
    assert gp.metrics()["numOutputRows"] == len(want)


def _split_rows(raw):
    out, o = [], 0
    while o < len(raw):
        sz = int.from_bytes(raw[o:o + 8], "little")
        out.append(bytes(raw[o:o + 8 + sz]))
        o += 8 + sz
    assert o == len(raw)
    return sorted(out)


def test_projected_rows_written_on_the_device_equal_the_host_writer(gpu_api, batches, monkeypatch):
    """MODE_PROJECT rows are sized, laid out and written by the GPU (sd_rows.cu) and leave in one copy.  Every field width
    (BOOLEAN, BYTE, SHORT, INT, DATE, FLOAT, DOUBLE, LONG, dictionary STRING), NULLs in the record's null bits and as dictionary
    NULL codes: against the oracle row for row, and byte for byte against the engine's host-side writer (SD_TUNE_HOST_ROWS, the
    path a plan with host-only dictionary entries still takes)."""
    b = PlanBuilder()
    c = cols(b)
    b.filter(c["c0"].is_null() | (c["c0"] > b.lit(T.INT)))
    b.project(*[c[name] for name, _, _ in SCHEMA])
    desc = b.build()
    raws = {}
    for host in (False, True):
        if host:
            monkeypatch.setenv("SD_TUNE_HOST_ROWS", "1")
        else:
            monkeypatch.delenv("SD_TUNE_HOST_ROWS", raising=False)
        gp = capi.Plan(gpu_api, desc).set_literals([-200])
        for x in batches:
            gp.submit(x)
        raws[host] = gp.finish_raw()
        assert gp.finish_raw() == raws[host]          # served again (a caller whose buffer was too small asks twice)
    assert len(raws[False]) == len(raws[True]) > 100_000
    assert _split_rows(raws[False]) == _split_rows(raws[True])
    op = oracle.plan(desc).set_literals([-200])
    for x in batches:
        op.submit(x)
    from snappydata_b200.column_format import parse_row_stream
    import struct

    def canon(r):   # bit-exact and NaN-safe: doubles by their bytes (c2 holds NaN, -0.0, inf)
        return tuple((0, b"") if v is None else (1, struct.pack("<d", v)) if isinstance(v, float) else (2, v) if isinstance(v, bytes)
                     else (3, struct.pack("<q", int(v))) for v in r)
    got = sorted(canon(r) for r in parse_row_stream(raws[False], desc.partial_schema()))
    want = sorted(canon(r) for r in op.finish())
    assert len(got) == len(want) and got == want
    # nothing passes: an empty stream, not an error
    b2 = PlanBuilder()
    c2 = cols(b2)
    b2.filter(c2["c0"] > b2.lit(T.INT))
    b2.project(c2["c0"], c2["c3"])
    gp = capi.Plan(gpu_api, b2.build()).set_literals([10**6])
    for x in batches:
        gp.submit(x)
    assert gp.finish_raw() == b""


def test_projection_output_larger_than_initial_buffer_is_replayed(gpu_api):
    """> 2^20 passing rows: the record buffer overflows, the engine grows it and replays the launches."""
    r = np.random.default_rng(9)
    n = 400_000
    schema = [("a", T.INT, False), ("b", T.DOUBLE, False)]
    bs = [build_batch(n, schema, {"a": np.arange(i * n, (i + 1) * n, dtype=np.int32), "b": r.random(n)}, batch_id=i) for i in range(3)]
    pb = PlanBuilder()
    a, b = pb.col(T.INT, 0), pb.col(T.DOUBLE, 1)
    pb.filter(a >= pb.lit(T.INT))
    pb.project(a, b)
    gp = capi.Plan(gpu_api, pb.build()).set_literals([100])
    for x in bs:
        gp.submit(x)
    got = gp.finish()
    assert len(got) == 3 * n - 100
    ids = np.sort(np.array([g[0] for g in got]))
    assert np.array_equal(ids, np.arange(100, 3 * n))


def test_group_by_double_key_with_nan_and_negative_zero(gpu_api, batches):
    """DOUBLE keys group with the NaN-safe equality (NaN == NaN, -0.0 == 0.0)."""
    b = PlanBuilder()
    c = cols(b)
    b.filter(c["c0"] > b.lit(T.INT))
    b.group_by(c["c2"], c["c4"])
    b.count().sum(c["c1"])
    both(gpu_api, b.build(), [900], batches, 2)


def test_many_string_key_combinations_switch_to_the_hash_table(gpu_api):
    """Dictionary-string keys whose id product exceeds the dense group table: the plan switches to the hash-table
    variant mid-execution and replays the launches already made."""
    r = np.random.default_rng(17)
    n = 150_000
    schema = [("a", T.STRING, False), ("b", T.STRING, True), ("v", T.LONG, False)]
    bs = []
    for i in range(2):
        data = {"a": np.array([b"a%03d" % x for x in r.integers(0, 300 + 100 * i, n)], dtype=object),
                "b": np.array([b"b%03d" % x for x in r.integers(0, 300, n)], dtype=object), "v": r.integers(-100, 100, n).astype(np.int64)}
        bs.append(build_batch(n, schema, data, {"b": r.random(n) < 0.05}, batch_id=i))
    b = PlanBuilder()
    a, bb, v = b.col(T.STRING, 0, False), b.col(T.STRING, 1, True), b.col(T.LONG, 2, False)
    b.group_by(a, bb)
    b.count().sum(v).min(v)
    gp, op, got = both(gpu_api, b.build(), [], bs, 2)
    assert len(got) > 65_536
