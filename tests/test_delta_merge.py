"""ColumnDeltaEncoder.merge (enc/ColumnDeltaEncoder.scala:348-556) in the product (sd_delta_merge, host only) against the spec the
fixture writer gives: decode both inputs, two-way merge of the positions with the NEW delta winning on equal positions, re-encode
with the type's default encoder -- byte for byte; and delta -> full column folding."""
import ctypes as C

import numpy as np
import pytest

from snappydata_b200 import capi
from snappydata_b200.column_format import (SqlType as T, compress_lz4, decode_column, decode_delta, encode_column, encode_delta)


def _merge(t, nullable, new, existing, existing_is_delta, num_rows):
    api = capi.product_api()
    f = api.lib.sd_delta_merge
    f.restype = C.c_int
    f.argtypes = [C.POINTER(capi.sd_column), C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int32, C.c_int32, C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]
    col = capi.sd_column(int(t), int(nullable), 0, 0, 18 if t == T.DECIMAL else 0)
    out = C.create_string_buffer(len(new) + len(existing) + 4096)
    n = C.c_int64()
    api.check(f(C.byref(col), new, len(new), existing, len(existing), int(existing_is_delta), num_rows, out, len(out), C.byref(n)))
    return out.raw[: n.value]


def _values(t, r, m):
    if t == T.STRING:
        return np.array([b"u%d" % x for x in r.integers(0, 9, m)], dtype=object)
    if t == T.BOOLEAN:
        return r.integers(0, 2, m).astype(bool)
    if t in (T.DOUBLE, T.FLOAT):
        return np.round(r.normal(0, 50, m), 2).astype("<f8" if t == T.DOUBLE else "<f4")
    return r.integers(-1000, 1000, m)


@pytest.mark.parametrize("t", [T.INT, T.LONG, T.DOUBLE, T.FLOAT, T.SHORT, T.BYTE, T.DATE, T.TIMESTAMP, T.DECIMAL, T.BOOLEAN, T.STRING])
@pytest.mark.parametrize("nullable", [False, True])
def test_delta_merged_with_delta(t, nullable):
    r = np.random.default_rng(int(t) * 10 + nullable)
    n = 5000
    p_new = np.sort(r.choice(n, 120, replace=False)).astype(np.int32)
    p_old = np.sort(np.unique(np.concatenate([r.choice(n, 700, replace=False), p_new[:30]]))).astype(np.int32)   # 30 positions in both
    v_new, v_old = _values(t, r, len(p_new)), _values(t, r, len(p_old))
    n_new = (r.random(len(p_new)) < 0.2) if nullable else None
    n_old = (r.random(len(p_old)) < 0.2) if nullable else None
    new, old = encode_delta(n, p_new, v_new, t, n_new), encode_delta(n, p_old, v_old, t, n_old)
    # the spec: union, new wins
    merged = {}
    for p, v, isn in zip(p_old, v_old, n_old if nullable else [False] * len(p_old)):
        merged[int(p)] = (v, bool(isn))
    for p, v, isn in zip(p_new, v_new, n_new if nullable else [False] * len(p_new)):
        merged[int(p)] = (v, bool(isn))
    pos = np.array(sorted(merged), dtype=np.int32)
    vals = np.array([merged[int(p)][0] for p in pos], dtype=object if t == T.STRING else None)
    nulls = np.array([merged[int(p)][1] for p in pos]) if nullable else None
    want = encode_delta(n, pos, vals, t, nulls)
    got = _merge(t, nullable, new, old, True, n)
    assert got == want
    nb, gp, gv, gn = decode_delta(got, t)
    assert nb == n and np.array_equal(gp, pos)
    # compressed inputs (a depth-1 delta is stored compressed when it is >= 2048 bytes) give the same result
    assert _merge(t, nullable, compress_lz4(new, force=True), compress_lz4(old, force=True), True, n) == want


@pytest.mark.parametrize("t", [T.INT, T.DOUBLE, T.BOOLEAN, T.STRING])
def test_delta_folded_into_the_full_column(t):
    r = np.random.default_rng(7 + int(t))
    n = 3000
    base_v = _values(t, r, n)
    base_n = r.random(n) < 0.1
    base = encode_column(base_v, t, base_n)
    p = np.sort(r.choice(n, 200, replace=False)).astype(np.int32)
    v, vn = _values(t, r, len(p)), r.random(len(p)) < 0.3
    delta = encode_delta(n, p, v, t, vn)
    eff_v = np.array(base_v, dtype=object if t == T.STRING else None).copy()
    eff_n = base_n.copy()
    for pp, vv, nn in zip(p, v, vn):
        eff_v[pp], eff_n[pp] = vv, nn
    want = encode_column(eff_v, t, eff_n)
    got = _merge(t, True, delta, base, False, n)
    assert got == want
    dv, dn = decode_column(got, t, n)
    assert np.array_equal(np.asarray(dn, dtype=bool), eff_n)


def test_errors():
    with pytest.raises(capi.SdError):
        _merge(T.INT, False, b"\\0" * 8, b"\\0" * 8, True, 10)
    good = encode_delta(10, np.array([1, 2], dtype=np.int32), np.array([5, 6]), T.INT, np.array([True, False]))
    with pytest.raises(capi.SdError):   # NULL entry for a NOT NULL column
        _merge(T.INT, False, good, encode_delta(10, np.array([3], dtype=np.int32), np.array([1]), T.INT), True, 10)


def test_fuzzed_delta_buffers_never_overrun_or_crash():
    """Corrupted inputs (bit flips, truncation, overwritten header words, spliced-in noise) to sd_delta_merge end in an error or in
    an output inside the caller's capacity; positions / counts / lengths read from the buffers are validated before use."""
    from snappydata_b200.column_format import encode_column
    api = capi.product_api()
    f = api.lib.sd_delta_merge
    f.restype = C.c_int
    f.argtypes = [C.POINTER(capi.sd_column), C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int32, C.c_int32, C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]
    r = np.random.default_rng(9)
    n_ok = n_err = 0
    for t in (T.INT, T.DOUBLE, T.STRING, T.BOOLEAN):
        for nullable in (False, True):
            n = 2000
            p_new = np.sort(r.choice(n, 60, replace=False)).astype(np.int32)
            p_old = np.sort(r.choice(n, 300, replace=False)).astype(np.int32)
            v_new, v_old = _values(t, r, len(p_new)), _values(t, r, len(p_old))
            n_new = (r.random(len(p_new)) < 0.2) if nullable else None
            n_old = (r.random(len(p_old)) < 0.2) if nullable else None
            new, old = encode_delta(n, p_new, v_new, t, n_new), encode_delta(n, p_old, v_old, t, n_old)
            full = encode_column(_values(t, r, n), t, (r.random(n) < 0.1) if nullable else None)
            col = capi.sd_column(int(t), int(nullable), 0, 0, 0)
            for trial in range(60):
                a, b, is_delta = bytearray(new), bytearray(old if trial % 3 else full), int(trial % 3 != 0)
                tgt = a if trial % 2 else b
                kind = trial % 5
                if kind == 0:
                    for _ in range(1 + trial % 4):
                        i = int(r.integers(0, len(tgt)))
                        tgt[i] ^= 1 << int(r.integers(0, 8))
                elif kind == 1:
                    del tgt[int(r.integers(0, len(tgt))):]
                elif kind == 2:
                    i = int(r.integers(0, max(1, len(tgt) - 4)))
                    tgt[i:i + 4] = int(r.integers(-2**31, 2**31)).to_bytes(4, "little", signed=True)
                elif kind == 3:
                    i = int(r.integers(0, len(tgt)))
                    k = min(len(tgt) - i, int(r.integers(1, 32)))
                    tgt[i:i + k] = bytes(r.integers(0, 256, k, dtype=np.uint8))
                cap = len(a) + len(b) + 4096
                out = C.create_string_buffer(cap + 64)
                m = C.c_int64()
                rc = f(C.byref(col), bytes(a), len(a), bytes(b), len(b), is_delta, n, out, cap, C.byref(m))
                assert out.raw[cap:] == bytes(64)
                if rc == 0:
                    assert 0 <= m.value <= cap
                    n_ok += 1
                else:
                    n_err += 1
    assert n_ok > 50 and n_err > 50
