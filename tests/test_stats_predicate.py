"""Batch skipping is a correctness feature: a batch that is wrongly skipped changes the answer.  The product decides it on
the host (StatEval in csrc/sd_engine.cu, restating ColumnTableScan.scala:820-963) before any byte moves; this test drives
that code through the host-only hook `sdx_stats_pass` and compares it with the oracle's independent restatement on a few
thousand random (filter tree, literal values, stats row) combinations: comparisons in both orientations of the bounds,
IN lists with NULLs, StartsWith, IsNull / IsNotNull, AND / OR / NOT, NULL bounds of all-null columns, NULL literals."""
import ctypes as C

import numpy as np

from oracle import oracle
from snappydata_b200 import capi
from snappydata_b200.column_format import ColumnBatch, SqlType as T, build_batch, stats_row
from snappydata_b200.plan import PlanBuilder

SCHEMA = [("c0", T.INT, True), ("c1", T.DOUBLE, True), ("c2", T.STRING, True), ("c3", T.LONG, False), ("c4", T.DATE, True)]
WORDS = [b"", b"a", b"ab", b"abc", b"abd", b"b", b"ba", b"zz", b"\xff", b"ab\xff"]


def _value(r, t, around=None):
    if t == T.STRING:
        return WORDS[int(r.integers(0, len(WORDS)))]
    if t == T.DOUBLE:
        x = float(r.integers(-5, 6)) + (0.5 if r.random() < 0.3 else 0.0)
        return float("nan") if r.random() < 0.03 else x
    return int(r.integers(-5, 6))


def _tree(r, b, cols, depth, lits):
    """random filter; appends the literal values it needs to `lits` in slot order"""
    if depth > 0 and r.random() < 0.55:
        k = r.random()
        if k < 0.45:
            return _tree(r, b, cols, depth - 1, lits) & _tree(r, b, cols, depth - 1, lits)
        if k < 0.9:
            return _tree(r, b, cols, depth - 1, lits) | _tree(r, b, cols, depth - 1, lits)
        return ~_tree(r, b, cols, depth - 1, lits)
    name, t, _ = SCHEMA[int(r.integers(0, len(SCHEMA)))]
    c = cols[name]
    k = r.random()
    if k < 0.55:
        op = ["eq", "__lt__", "__le__", "__gt__", "__ge__"][int(r.integers(0, 5))]
        lit = b.lit(t)
        lits.append(None if r.random() < 0.05 else _value(r, t))
        return getattr(c, op)(lit)
    if k < 0.7:
        n = int(r.integers(1, 5))
        e = c.isin(n)
        lits.extend(None if r.random() < 0.15 else _value(r, t) for _ in range(n))
        return e
    if k < 0.8 and t == T.STRING:
        lit = b.lit(T.STRING)
        lits.append(None if r.random() < 0.05 else _value(r, T.STRING))
        return c.startswith(lit)
    return c.is_null() if r.random() < 0.5 else c.is_not_null()


def _stats(r, num_rows):
    st = []
    for _, t, nullable in SCHEMA:
        if nullable and r.random() < 0.12:          # all-null column: NULL bounds
            st.append((t, None, None, num_rows))
            continue
        a, c = _value(r, t), _value(r, t)
        if t == T.DOUBLE and (a != a or c != c):
            a, c = 0.0, 1.0
        lo, hi = (a, c) if a <= c else (c, a)
        st.append((t, lo, hi, int(r.integers(0, num_rows)) if nullable and r.random() < 0.5 else 0))
    return st


def test_product_stats_predicate_matches_the_oracle(oracle_api):
    api = capi.product_api()
    L = api.lib
    L.sdx_stats_pass.restype = C.c_int
    L.sdx_stats_pass.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_char_p, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    r = np.random.default_rng(11)
    num_rows = 8
    data = {"c0": np.zeros(num_rows, np.int32), "c1": np.zeros(num_rows), "c2": np.array([b"ab"] * num_rows, dtype=object),
            "c3": np.zeros(num_rows, np.int64), "c4": np.zeros(num_rows, np.int32)}
    base = build_batch(num_rows, SCHEMA, data, {})
    skipped = passed = 0
    for case in range(3000):
        b = PlanBuilder()
        cols = {name: b.col(t, i, nullable) for i, (name, t, nullable) in enumerate(SCHEMA)}
        lits = []
        b.filter(_tree(r, b, cols, 3, lits))
        b.count()
        desc = b.build()
        st = stats_row(num_rows, _stats(r, num_rows))
        # oracle: submit a batch that carries this stats row and see whether it was skipped
        op = oracle.plan(desc).set_literals(lits)
        op.submit(ColumnBatch(num_rows=num_rows, columns=base.columns, stats=st))
        want_pass = op.metrics()["columnBatchesSkipped"] == 0
        op.close()
        # product: the same decision from StatEval
        arr = (capi.sd_literal * max(1, len(lits)))()
        for i, v in enumerate(lits):
            arr[i] = capi.make_literal(desc.literal_types_py[i], v)
        got = C.c_int32(-1)
        api.check(L.sdx_stats_pass(C.byref(desc.c), arr, len(lits), st, len(st), len(SCHEMA), num_rows, C.byref(got)))
        assert bool(got.value) == want_pass, (case, lits, desc.dump() if hasattr(desc, "dump") else None)
        skipped += not want_pass
        passed += want_pass
    assert skipped > 300 and passed > 300, (skipped, passed)   # the generator exercises both outcomes
