"""Times the device LZ4 decoder on synthetic column buffers (no engine around it) and checks every byte against the
input: the quick loop for working on snappydata_b200/csrc/sd_lz4.cu.

    python tools/lz4_bench.py [buffers per kind=64] [rows per buffer=200000] [reps=5]

Per kind of column it prints the compressed ratio, the time of one launch over `buffers` identical-shape buffers (a
launch lasts as long as its longest buffer chain: the figure of merit is ms per buffer chain) and the aggregate output
rate of that launch; every kernel variant (default / dense shape, serial / window parse).  (Written at the end of round 1 after the GPU budget was spent:
not yet run on hardware.)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snappydata_b200 import capi  # noqa: E402
from snappydata_b200.column_format import compress_lz4  # noqa: E402


def kinds(rows, rng):
    yield "double, 50 distinct (l_quantity)", rng.integers(1, 51, rows).astype(np.float64).tobytes()
    yield "double, 11 distinct (l_discount)", (rng.integers(0, 11, rows) / 100.0).astype(np.float64).tobytes()
    yield "int16 codes, 3 distinct (l_returnflag)", rng.integers(0, 3, rows).astype(np.int16).tobytes()
    yield "int32 dates, 2526 distinct (l_shipdate)", (8036 + rng.integers(0, 2526, rows)).astype(np.int32).tobytes()
    yield "double, ~unique (l_extendedprice)", np.round(rng.uniform(900, 105000, rows), 2).tobytes()


def main():
    nbuf = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    api = capi.product_api()
    api.check(api.init(0))
    L = api.lib
    L.sdx_lz4_expand.restype = C.c_int
    L.sdx_lz4_expand.argtypes = [C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int32, C.c_int32,
                                 C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_double)]
    want_kinds = [int(x) for x in os.environ.get("LZ4_KINDS", "0,1,2,3,4").split(",")]
    want_variants = [int(x) for x in os.environ.get("LZ4_VARIANTS", "0,1,2,3").split(",")]
    ndistinct = min(nbuf, int(os.environ.get("LZ4_DISTINCT", "256")))   # distinct buffers per kind (the launch cycles through them)
    names = [n for n, _ in kinds(8, np.random.default_rng(0))]
    for ki in want_kinds:
        name = names[ki]
        base = [dict(kinds(rows, np.random.default_rng(100 + i)))[name] for i in range(ndistinct)]
        raws = [base[i % ndistinct] for i in range(nbuf)]
        cblocks = [compress_lz4(b, force=True)[8:] for b in base]
        blocks = [cblocks[i % ndistinct] for i in range(nbuf)]
        ratio = sum(map(len, blocks)) / sum(map(len, raws))
        for variant in want_variants:   # bit 0: dense kernel shape, bit 1: window parse
            for mis in ((0, 8) if variant == 0 else (8,)):
                keep = [C.create_string_buffer(b, len(b)) for b in cblocks]
                keep = [keep[i % ndistinct] for i in range(nbuf)]
                outs = [C.create_string_buffer(len(b)) for b in raws]
                bp = (C.c_void_p * nbuf)(*[C.cast(k, C.c_void_p) for k in keep])
                op = (C.c_void_p * nbuf)(*[C.cast(o, C.c_void_p) for o in outs])
                bl = (C.c_int64 * nbuf)(*[len(b) for b in blocks])
                ol = (C.c_int64 * nbuf)(*[len(b) for b in raws])
                ms = C.c_double()
                api.check(L.sdx_lz4_expand(0, bp, bl, ol, nbuf, mis, variant, reps, op, C.byref(ms)))
                for o, b in zip(outs, raws):
                    assert o.raw == b, "device output differs from the input of the compressor"
                out_bytes = sum(map(len, raws))
                label = ("dense" if variant & 1 else "default") + ("+window-parse" if variant & 2 else "")
                print(f"{name:42s} ratio {ratio:.2f}  {label:20s} misalign {mis:2d}: {ms.value:8.3f} ms per launch of {nbuf} "
                      f"buffers ({len(raws[0]) / 1e6:.2f} MB each) = {out_bytes / ms.value / 1e6:8.2f} GB/s out")


if __name__ == "__main__":
    main()
