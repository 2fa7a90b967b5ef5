"""Stored-buffer envelopes [-codecId][uncompressedLen][payload] (encoders/.../store/CompressionUtils.scala:53-61,125-168): the
product's host-side decoder (LZ4 and Snappy) against the fixture writer's encoders; no GPU needed."""
import ctypes as C

import numpy as np
import pytest

from snappydata_b200 import capi
from snappydata_b200.column_format import compress_lz4, compress_snappy


def _decode(buf: bytes):
    api = capi.product_api()
    f = api.lib.sdx_decompress_envelope
    f.restype = C.c_int
    f.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]
    n = C.c_int64()
    out = C.create_string_buffer(max(1, int.from_bytes(buf[4:8], "little", signed=True)))
    rc = f(buf, len(buf), out, len(out), C.byref(n))
    return rc, out.raw[: n.value]


def _samples():
    rng = np.random.default_rng(11)
    yield b"abc"
    yield bytes(rng.integers(0, 4, 5000, dtype=np.uint8))                              # short copies with 1-byte offsets
    yield (rng.integers(0, 11, 30000) / 100.0).astype(np.float64).tobytes()            # 8-byte periodic matches
    yield bytes(rng.integers(0, 256, 70000, dtype=np.uint8))                           # incompressible: long literals
    yield b"x" * 100000                                                                # overlapping copies (offset 1)
    yield np.arange(50000, dtype=np.int32).tobytes()


@pytest.mark.parametrize("codec", ["lz4", "snappy"])
def test_host_decoder_round_trips(codec):
    for raw in _samples():
        env = compress_lz4(raw, force=True) if codec == "lz4" else compress_snappy(raw)
        rc, got = _decode(env)
        assert rc == 0 and got == raw, (codec, len(raw))


def test_corrupt_envelopes_are_errors():
    raw = (np.arange(3000) % 7).astype(np.float64).tobytes()
    for env in (compress_lz4(raw, force=True), compress_snappy(raw)):
        bad = bytes(env[: len(env) // 2])              # truncated payload
        rc, _ = _decode(bad)
        assert rc != 0
    rc, _ = _decode(b"\xfd\xff\xff\xff" + b"\x10\x00\x00\x00" + b"\x00" * 16)   # codec id 3: unknown
    assert rc == capi.SD_ERR_UNSUPPORTED
