"""The host-side LZ4 prefix decoder (used to read the header / null words / dictionary of a compressed column
buffer before the device expands it) against liblz4, on CPU."""
import ctypes as C

import numpy as np

from snappydata_b200 import capi
from snappydata_b200.column_format import SqlType, compress_lz4, encode_dictionary, encode_uncompressed


def test_prefix_decoder_matches_liblz4():
    lib = capi.product_api().lib
    lib.sdx_lz4_decode_prefix.restype = C.c_int64
    lib.sdx_lz4_decode_prefix.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64]
    r = np.random.default_rng(0)
    bufs = [encode_uncompressed((r.integers(0, 11, 50_000) / 100.0), SqlType.DOUBLE),
            encode_uncompressed(r.integers(0, 3, 70_000).astype(np.int32), SqlType.INT, r.random(70_000) < 0.2),
            encode_dictionary(np.array([b"value-%03d" % x for x in r.integers(0, 300, 40_000)], dtype=object), SqlType.STRING),
            bytes(r.integers(0, 256, 5000, dtype=np.uint8)) * 3, b"\x07" * 100_000]
    for raw in bufs:
        env = compress_lz4(raw, force=True)
        assert int.from_bytes(env[:4], "little", signed=True) == -1 and int.from_bytes(env[4:8], "little") == len(raw)
        for want in (8, 9, 100, 4099, len(raw)):
            want = min(want, len(raw))
            out = C.create_string_buffer(want + 16)
            got = lib.sdx_lz4_decode_prefix(env[8:], len(env) - 8, out, want)
            assert got == want
            assert out.raw[:want] == raw[:want]
    # corrupt input is reported, not read past
    assert lib.sdx_lz4_decode_prefix(b"\xf0", 1, C.create_string_buffer(64), 32) == -1
