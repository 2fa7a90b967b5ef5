"""sd_final_merge / sd_partial_merge (host only) parse row streams that arrive from other partitions: corrupted streams -- bit
flips, truncation, overwritten size words, spliced-in noise -- must end in an error or in an output inside the caller's capacity."""
import ctypes as C

import numpy as np

from oracle import oracle
from snappydata_b200 import capi, lineitem
from snappydata_b200 import plan as P
from snappydata_b200.column_format import SqlType as T
from snappydata_b200.plan import PlanBuilder


def test_fuzzed_partial_row_streams_never_overrun_or_crash():
    api = capi.product_api()
    r = np.random.default_rng(3)
    table = lineitem.gen_table(40_000, 10_000, seed=4)
    b = PlanBuilder()
    rf, ship, qty, price = (b.col(T.STRING, P.L_RETURNFLAG), b.col(T.DATE, P.L_SHIPDATE), b.col(T.DOUBLE, P.L_QUANTITY),
                            b.col(T.DOUBLE, P.L_EXTENDEDPRICE))
    b.group_by(rf, ship)
    b.count().sum(qty).avg(price).min(price).max(rf)
    n_ok = n_err = 0
    for desc, lits in ((P.q1_plan(), P.Q1_LITERALS), (P.q6_plan(), P.Q6_LITERALS), (b.build(), [])):
        op = oracle.plan(desc).set_literals(lits)
        for x in table:
            op.submit(x)
        raw = op.finish_raw()[:20000]            # (cut at an arbitrary byte: also a corruption for the big result)
        for fn in (api.lib.sd_final_merge, api.lib.sd_partial_merge):
            for trial in range(80):
                bb = bytearray(raw)
                kind = trial % 4
                if kind == 0:
                    for _ in range(1 + trial % 6):
                        i = int(r.integers(0, len(bb)))
                        bb[i] ^= 1 << int(r.integers(0, 8))
                elif kind == 1:
                    del bb[int(r.integers(0, len(bb))):]
                elif kind == 2:
                    i = int(r.integers(0, max(1, len(bb) - 8)))
                    bb[i:i + 8] = int(r.integers(-2**40, 2**40)).to_bytes(8, "little", signed=True)
                else:
                    i = int(r.integers(0, len(bb)))
                    k = min(len(bb) - i, int(r.integers(1, 48)))
                    bb[i:i + k] = bytes(r.integers(0, 256, k, dtype=np.uint8))
                cap = 4 * len(bb) + 4096
                out = C.create_string_buffer(cap + 64)
                ol, orows = C.c_int64(), C.c_int64()
                rc = fn(C.byref(desc.c), bytes(bb), len(bb), out, cap, C.byref(ol), C.byref(orows))
                assert out.raw[cap:] == bytes(64)
                if rc == 0:
                    assert 0 <= ol.value <= cap
                    n_ok += 1
                else:
                    n_err += 1
    assert n_ok > 20 and n_err > 20
