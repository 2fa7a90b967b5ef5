"""Variable-width (non-dictionary) STRING columns on the GPU path -- SURVEY.md 8a row a3 (enc/Uncompressed.scala:116-161:
back-to-back [len:int32][bytes] read with a sequential cursor) -- and the stored-compressed forms the reference can hand
over besides LZ4 value buffers (Snappy envelopes, compressed update deltas / delete masks, LZ4 over run-length and
variable-width bodies; encoders/.../store/CompressionUtils.scala:125-168).  Every result is compared with the oracle on the
same bytes.  A column may be dictionary-encoded in one batch and variable-width in the next (the reference's encoder
falls back to Uncompressed when a dictionary grows too large): both forms meet in one execution here."""
import copy

import numpy as np
import pytest

from oracle import oracle
from snappydata_b200 import capi
from snappydata_b200.column_format import (ColumnBatch, SqlType as T, build_batch, compress_lz4, compress_snappy, encode_delete,
                                           encode_delta)
from snappydata_b200.plan import PlanBuilder

from helpers import assert_rowsets_match

pytestmark = pytest.mark.gpu

SCHEMA = [("s", T.STRING, True), ("t", T.STRING, False), ("v", T.INT, False), ("d", T.DOUBLE, True)]
WORDS = [b"", b"a", b"ab", b"abc", b"abd", b"b", b"ba", b"zeta", b"\xc3\xa9t\xc3\xa9", b"a" * 40, b"ab\x00c", b"abcdefghijklmnop"]


def make(n, seed, raw=("s", "t"), batch_id=0, high_card=False):
    r = np.random.default_rng(seed)
    pick = r.integers(0, len(WORDS), n)
    s = np.array([WORDS[i] for i in pick], dtype=object)
    if high_card:
        t = np.array([b"row-%d-%d" % (seed, i) for i in range(n)], dtype=object)
    else:
        t = np.array([b"k%02d" % x for x in r.integers(0, 23, n)], dtype=object)
    data = {"s": s, "t": t, "v": r.integers(-50, 50, n).astype(np.int32), "d": np.round(r.normal(0, 10, n), 2)}
    nulls = {"s": r.random(n) < 0.15, "d": r.random(n) < 0.1}
    enc = {c: "uncompressed" for c in raw}
    return build_batch(n, SCHEMA, data, nulls, batch_id=batch_id, encoders=enc)


def cols(b):
    return {name: b.col(t, i, nullable) for i, (name, t, nullable) in enumerate(SCHEMA)}


def both(gpu_api, desc, lits, batches, nkeys, store=False):
    op = oracle.plan(desc).set_literals(lits)
    for x in batches:
        op.submit(x)
    want = op.finish()
    gp = capi.Plan(gpu_api, desc).set_literals(lits)
    if store:
        st = capi.Store(gpu_api, [(t, n) for _, t, n in SCHEMA], 0)
        for x in batches:
            st.put(x)
        gp.scan_store(st)
    else:
        for x in batches:
            gp.submit(x)
    got = gp.finish()
    assert_rowsets_match(got, want, nkeys)
    return got, gp


@pytest.fixture(scope="module")
def mixed():
    """raw + dictionary batches of the same columns, ragged sizes"""
    return [make(3000, 1, raw=("s", "t"), batch_id=0), make(1, 2, raw=("s",), batch_id=1), make(2049, 3, raw=(), batch_id=2),
            make(777, 4, raw=("t",), batch_id=3)]


@pytest.mark.parametrize("store", [False, True])
def test_predicates_on_raw_and_dictionary_batches(gpu_api, mixed, store):
    b = PlanBuilder()
    c = cols(b)
    L = lambda: b.lit(T.STRING)
    pred = (c["s"] >= L()) & (c["s"] < L()) | c["s"].eq(L()) | c["t"].startswith(L()) | (c["t"] > L())
    b.filter(pred)
    b.count().sum(c["v"]).count(c["s"]).sum(c["d"])
    both(gpu_api, b.build(), [b"ab", b"b", b"zeta", b"k1", b"k20"], mixed, 0, store)
    b = PlanBuilder()   # literal on the left, NE / LE, IN with a NULL literal, NOT over a NULL-yielding compare
    c = cols(b)
    b.filter((b.lit(T.STRING) < c["s"]) & c["s"].ne(b.lit(T.STRING)) & ~(c["t"] <= b.lit(T.STRING)) | c["s"].isin(3))
    b.count().sum(c["v"])
    both(gpu_api, b.build(), [b"a", b"abc", b"k05", b"", None, b"\xc3\xa9t\xc3\xa9"], mixed, 0, store)
    for lit in (b"", b"a", b"ab\x00c", b"abcdefghijklmnop", b"nope", None):   # equality with every kind of literal, incl. NULL
        b = PlanBuilder()
        c = cols(b)
        b.filter(c["s"].eq(b.lit(T.STRING)))
        b.count().count(c["d"])
        both(gpu_api, b.build(), [lit], mixed, 0, store)


def test_group_by_raw_string_keys_switches_to_the_hash_table(gpu_api, mixed):
    b = PlanBuilder()
    c = cols(b)
    b.group_by(c["s"])
    b.count().sum(c["v"]).avg(c["d"])
    got, gp = both(gpu_api, b.build(), [], mixed, 1)
    assert any(r[0] is None for r in got) and len(got) == len(WORDS) + 1
    b = PlanBuilder()   # two string keys (one nullable) + an integer expression key
    c = cols(b)
    b.filter(c["v"] > b.lit(T.INT))
    b.group_by(c["t"], c["s"], c["v"] + c["v"])
    b.count().min(c["d"]).max(c["v"])
    both(gpu_api, b.build(), [40], mixed, 3)
    # dictionary batches first (dense table), then a raw batch: the execution switches and replays what it had launched
    order = [mixed[2], mixed[0], mixed[3], mixed[1]]
    b = PlanBuilder()
    c = cols(b)
    b.group_by(c["t"])
    b.count().sum(c["v"])
    both(gpu_api, b.build(), [], order, 1)
    both(gpu_api, b.build(), [], order, 1, store=True)


def test_high_cardinality_raw_string_key(gpu_api):
    batches = [make(20000, 10 + i, raw=("t",), batch_id=i, high_card=True) for i in range(3)]
    b = PlanBuilder()
    c = cols(b)
    b.group_by(c["t"])
    b.count().sum(c["v"])
    got, _ = both(gpu_api, b.build(), [], batches, 1)
    assert len(got) == 60000


def test_projection_of_raw_strings(gpu_api, mixed):
    b = PlanBuilder()
    c = cols(b)
    b.filter((c["v"] >= b.lit(T.INT)) & c["t"].startswith(b.lit(T.STRING)))
    b.project(c["v"], c["s"], c["t"], c["d"])
    op = oracle.plan(b.build()).set_literals([30, b"k1"])
    gp = capi.Plan(gpu_api, b.build()).set_literals([30, b"k1"])
    for x in mixed:
        op.submit(x)
        gp.submit(x)
    want, got = op.finish(), gp.finish()
    assert len(want) > 50
    assert_rowsets_match(got, want, 4)


def test_min_max_over_string_columns(gpu_api, mixed):
    """MIN / MAX(STRING): slot = address of the winning record, compared by bytes; dense table, hash table and no-key plans,
    dictionary and raw batches mixed."""
    b = PlanBuilder()
    c = cols(b)
    b.group_by(c["t"])                       # dictionary keys -> dense table until the raw batch switches it to the hash table
    b.min(c["s"]).max(c["s"]).count(c["s"]).sum(c["v"])
    both(gpu_api, b.build(), [], mixed, 1)
    both(gpu_api, b.build(), [], [mixed[2]], 1)          # dictionary batches only: stays a dense table
    b = PlanBuilder()
    c = cols(b)
    b.group_by(c["v"])                       # integer key -> hash table
    b.min(c["t"]).max(c["s"])
    both(gpu_api, b.build(), [], mixed, 1)
    b = PlanBuilder()
    c = cols(b)
    b.filter(c["d"] > b.lit(T.DOUBLE))
    b.min(c["s"]).max(c["s"]).min(c["t"]).max(c["t"]).count()
    both(gpu_api, b.build(), [5.0], mixed, 0)
    both(gpu_api, b.build(), [1e9], mixed, 0)            # nothing passes: NULL results
    both(gpu_api, b.build(), [5.0], mixed, 0, store=True)


def _recompress(batch, fn):
    c = copy.copy(batch)
    c.columns = [None if x is None else fn(x) for x in batch.columns]
    return c


@pytest.mark.parametrize("codec", ["snappy", "lz4"])
def test_compressed_raw_string_and_other_columns(gpu_api, mixed, codec):
    fn = compress_snappy if codec == "snappy" else (lambda x: compress_lz4(x, force=True))
    comp = [_recompress(x, fn) for x in mixed]
    b = PlanBuilder()
    c = cols(b)
    b.filter(c["s"].ne(b.lit(T.STRING)))
    b.group_by(c["t"])
    b.count().sum(c["v"]).sum(c["d"])
    desc = b.build()
    op = oracle.plan(desc).set_literals([b"abc"])
    gp = capi.Plan(gpu_api, desc).set_literals([b"abc"])
    for x, y in zip(mixed, comp):
        op.submit(x)
        gp.submit(y)
    assert_rowsets_match(gp.finish(), op.finish(), 1)


def test_compressed_deltas_delete_mask_and_rle(gpu_api):
    """a depth-1 delta of ~1.3k doubles is >= 2048 B and is stored compressed (CompressionUtils.scala:47-49); so may the
    delete mask be; LZ4 over a run-length column needs the host to see every run."""
    r = np.random.default_rng(5)
    n = 6000
    schema = [("a", T.DOUBLE, False), ("k", T.INT, False), ("r", T.LONG, False)]
    data = {"a": np.round(r.normal(0, 5, n), 2), "k": r.integers(0, 7, n).astype(np.int32), "r": np.repeat(r.integers(0, 50, n // 40 + 1), 40)[:n].astype(np.int64)}
    plain = build_batch(n, schema, data, {}, encoders={"r": "rle"})
    p1 = np.sort(r.choice(n, 1300, replace=False)).astype(np.int32)
    p0 = np.sort(r.choice(n, 90, replace=False)).astype(np.int32)
    d1 = encode_delta(n, p1, np.round(r.normal(0, 5, len(p1)), 2), T.DOUBLE)
    d0 = encode_delta(n, p0, np.round(r.normal(0, 5, len(p0)), 2), T.DOUBLE)
    dele = encode_delete(n, np.sort(r.choice(n, 900, replace=False)))
    assert len(d1) >= 2048
    plain.delta0, plain.delta1, plain.delete_mask = {0: d0}, {0: d1}, dele
    for fn in (compress_snappy, lambda x: compress_lz4(x, force=True)):
        comp = copy.copy(plain)
        comp.columns = [plain.columns[0], plain.columns[1], fn(plain.columns[2])]   # compressed run-length column
        comp.delta0, comp.delta1, comp.delete_mask = {0: d0}, {0: fn(d1)}, fn(dele)
        b = PlanBuilder()
        a, k, rr = b.col(T.DOUBLE, 0, False), b.col(T.INT, 1, False), b.col(T.LONG, 2, False)
        b.group_by(k)
        b.count().sum(a).sum(rr)
        desc = b.build()
        op = oracle.plan(desc).set_literals([])
        gp = capi.Plan(gpu_api, desc).set_literals([])
        op.submit(plain)
        gp.submit(comp)
        assert_rowsets_match(gp.finish(), op.finish(), 1)
