"""Loader for oracle/liboracle.so (the C restatement of the reference's scan->filter->aggregate loop).

TEST INFRASTRUCTURE -- see oracle/scan_oracle.c.  Exposes the same handle classes as the product
(snappydata_b200.capi.Plan) bound to the ``oracle_`` symbols, plus the hand-restated generated loops
used as the timed CPU baseline.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

from snappydata_b200 import capi
from snappydata_b200.column_format import ColumnBatch

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_DIR, "liboracle.so")
_api: Optional[capi.Api] = None


def build(force: bool = False) -> str:
    src = os.path.join(_DIR, "scan_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _DIR, "-B", "liboracle.so"])
    return _LIB


def api() -> capi.Api:
    global _api
    if _api is None:
        build()
        _api = capi.Api(_LIB, "oracle_")
        L = _api.lib
        L.oracle_run_partitions.restype = C.c_int
        L.oracle_run_partitions.argtypes = [C.c_int, C.POINTER(capi.sd_batch), C.c_int, C.c_int, C.c_int32, C.c_int32,
                                            C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double),
                                            C.POINTER(C.c_int64), C.c_void_p]
        L.oracle_q1_group_size.restype = C.c_int
        L.oracle_q1_max_groups.restype = C.c_int
        for name in ("oracle_bitset_is_set", "oracle_bitset_next_set_bit", "oracle_bitset_cardinality"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_int, C.c_int]
    return _api


def plan(desc: capi.PlanDesc) -> capi.Plan:
    return capi.Plan(api(), desc)


def final_merge(desc: capi.PlanDesc, partial_raw: bytes):
    return capi.final_merge(api(), desc, partial_raw)


class _Q1Group(C.Structure):
    _fields_ = [("used", C.c_int), ("k0", C.c_uint8 * 8), ("l0", C.c_int), ("k1", C.c_uint8 * 8), ("l1", C.c_int),
                ("sum_qty", C.c_double), ("sum_price", C.c_double), ("sum_disc_price", C.c_double),
                ("sum_charge", C.c_double), ("avg_qty_s", C.c_double), ("avg_price_s", C.c_double),
                ("avg_disc_s", C.c_double), ("avg_qty_c", C.c_int64), ("avg_price_c", C.c_int64),
                ("avg_disc_c", C.c_int64), ("count", C.c_int64)]


class BatchArray:
    """Contiguous sd_batch[] over marshalled batches (kept alive here)."""

    def __init__(self, batches: Sequence[ColumnBatch], table_cols: Sequence[int]):
        self.m = [capi.MarshalledBatch(b, table_cols) for b in batches]
        self.arr = (capi.sd_batch * max(1, len(self.m)))()
        for i, mb in enumerate(self.m):
            self.arr[i] = mb.c
        self.n = len(self.m)


def run_q6(ba: BatchArray, lits: Sequence[float], nthreads: int = 1):
    """Generated-loop restatement of Q6 over partitions = threads -> (sum or None, matched rows)."""
    a = api()
    assert C.sizeof(_Q1Group) == a.lib.oracle_q1_group_size()
    out_f = (C.c_double * 2)()
    out_i = (C.c_int64 * 2)()
    rc = a.lib.oracle_run_partitions(6, ba.arr, ba.n, nthreads, int(lits[0]), int(lits[1]), float(lits[2]),
                                     float(lits[3]), float(lits[4]), out_f, out_i, None)
    a.check(rc)
    return (None if out_i[0] else out_f[0]), out_i[1]


def run_c1(ba: BatchArray, k: int, nthreads: int = 1) -> int:
    a = api()
    out_f = (C.c_double * 2)()
    out_i = (C.c_int64 * 2)()
    a.check(a.lib.oracle_run_partitions(1, ba.arr, ba.n, nthreads, int(k), 0, 0.0, 0.0, 0.0, out_f, out_i, None))
    return out_i[0]


def run_q1(ba: BatchArray, cutoff: int, nthreads: int = 1):
    """-> list of partial rows [rf, ls, 11 buffer fields] in the layout of q1_plan().partial_schema()."""
    a = api()
    n = a.lib.oracle_q1_max_groups()
    groups = (_Q1Group * n)()
    out_f = (C.c_double * 2)()
    out_i = (C.c_int64 * 2)()
    a.check(a.lib.oracle_run_partitions(11, ba.arr, ba.n, nthreads, int(cutoff), 0, 0.0, 0.0, 0.0, out_f, out_i,
                                        C.cast(groups, C.c_void_p)))
    rows = []
    for g in groups:
        if g.used:
            rows.append([bytes(g.k0[: g.l0]), bytes(g.k1[: g.l1]), g.sum_qty, g.sum_price, g.sum_disc_price,
                         g.sum_charge, g.avg_qty_s, g.avg_qty_c, g.avg_price_s, g.avg_price_c, g.avg_disc_s,
                         g.avg_disc_c, g.count])
    return rows
