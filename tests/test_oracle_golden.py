"""Pins the CPU oracle against the reference's own known answers for this path
so that it can be trusted as the checker of the CUDA path."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle
from snappydata_b200 import plan as P
from snappydata_b200.capi import final_merge

from helpers import format_q1, load_tpch_golden


@pytest.fixture(scope="module")
def tpch():
    return load_tpch_golden()


def test_q1_matches_snappy_1_out(tpch, oracle_api):
    batches, want_q1, _ = tpch
    desc = P.q1_plan()
    pl = oracle.plan(desc).set_literals(P.Q1_LITERALS)
    for b in batches:
        pl.submit(b)
    final = final_merge(oracle_api, desc, pl.finish_raw())
    assert format_q1(final) == sorted(want_q1)


def test_q6_matches_snappy_6_out(tpch, oracle_api):
    batches, _, want_q6 = tpch
    desc = P.q6_plan()
    pl = oracle.plan(desc).set_literals(P.Q6_LITERALS)
    for b in batches:
        pl.submit(b)
    raw = pl.finish_raw()
    (row,) = final_merge(oracle_api, desc, raw)
    assert ("%18.4f" % row[0]).strip() == want_q6
    assert pl.metrics()["rowsScanned"] == 30201


def test_q6_float_folded_bounds_give_the_wrong_answer(tpch, oracle_api):
    """literals must arrive DECIMAL-folded (0.05 / 0.07); re-deriving them in
    binary64 (0.06 + 0.01 = 0.06999...) yields 362966.1362, which pins the oracle's comparison."""
    batches, _, _ = tpch
    desc = P.q6_plan()
    pl = oracle.plan(desc).set_literals([8766, 9131, 0.06 - 0.01, 0.06 + 0.01, 24.0])
    for b in batches:
        pl.submit(b)
    (row,) = final_merge(oracle_api, desc, pl.finish_raw())
    assert ("%18.4f" % row[0]).strip() == "362966.1362"


def test_generated_loop_restatements_agree_with_interpreter(tpch, oracle_api):
    """Layer 2 (the timed CPU baseline) == layer 1 (the generic checker) on the golden data."""
    batches, want_q1, want_q6 = tpch
    d6, d1 = P.q6_plan(), P.q1_plan()
    for threads in (1, 3):
        s, matched = oracle.run_q6(oracle.BatchArray(batches, d6.table_cols), P.Q6_LITERALS, threads)
        assert ("%18.4f" % s).strip() == want_q6 and matched == 594
        rows = oracle.run_q1(oracle.BatchArray(batches, d1.table_cols), P.Q1_LITERALS[0], threads)
        final = [[r[0], r[1], r[2], r[3], r[4], r[5], r[6] / r[7], r[8] / r[9], r[10] / r[11], r[12]] for r in rows]
        assert format_q1(final) == sorted(want_q1)


def _words(bits):
    n = (max(bits) >> 6) + 1 if bits else 1
    w = np.zeros(n, dtype=np.uint64)
    for b in bits:
        w[b >> 6] |= np.uint64(1) << np.uint64(b & 63)
    return w


def test_bitset_known_answers(oracle_api):
    """Known answers in the style of cluster/src/test/scala/org/apache/spark/sql/store/BitSetTest.scala:73-196
    (set bits {0, 9, 1, 10, 90, 96}: isSet, nextSetBit walks, cardinality)."""
    L = oracle_api.lib
    setbits = [0, 9, 1, 10, 90, 96]
    w = _words(setbits)
    ptr, nw = w.ctypes.data, w.shape[0]
    for i in range(100):
        assert bool(L.oracle_bitset_is_set(ptr, i, nw)) == (i in setbits)
    assert L.oracle_bitset_is_set(ptr, 1000, nw) == 0           # beyond the trimmed words: not null
    walk, pos = [], L.oracle_bitset_next_set_bit(ptr, 0, nw)
    while pos != 2**31 - 1:
        walk.append(pos)
        pos = L.oracle_bitset_next_set_bit(ptr, pos + 1, nw)
    assert walk == sorted(setbits)
    assert L.oracle_bitset_next_set_bit(ptr, 97, nw) == 2**31 - 1
    for upto, want in [(0, 0), (1, 1), (2, 2), (10, 3), (11, 4), (64, 4), (91, 5), (97, 6), (128, 6), (4096, 6)]:
        assert L.oracle_bitset_cardinality(ptr, upto, nw) == want
