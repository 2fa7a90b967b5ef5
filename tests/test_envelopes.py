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


def test_fuzzed_envelopes_never_overrun_and_either_fail_or_fill_the_declared_length():
    """Stored buffers are untrusted bytes: bit flips, truncations, a changed declared length and spliced-in noise must end in an
    error or in exactly the declared number of bytes -- never past the caller's capacity, never a crash."""
    api = capi.product_api()
    f = api.lib.sdx_decompress_envelope
    f.restype = C.c_int
    f.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]
    r = np.random.default_rng(5)
    n_ok = n_err = 0
    for raw in _samples():
        for env in (compress_lz4(raw, force=True), compress_snappy(raw)):
            for trial in range(100):
                b = bytearray(env)
                kind = trial % 4
                if kind == 0:
                    for _ in range(1 + trial % 5):
                        i = int(r.integers(8, len(b)))
                        b[i] ^= 1 << int(r.integers(0, 8))
                elif kind == 1:
                    b = b[: int(r.integers(8, len(b)))]
                elif kind == 2:
                    b[4:8] = int(r.integers(0, 2 * len(raw) + 10)).to_bytes(4, "little")
                else:
                    i = int(r.integers(8, len(b)))
                    k = min(len(b) - i, int(r.integers(1, 64)))
                    b[i:i + k] = bytes(r.integers(0, 256, k, dtype=np.uint8))
                declared = int.from_bytes(b[4:8], "little", signed=True)
                cap = max(1, declared) if 0 <= declared < (1 << 24) else 1
                out = C.create_string_buffer(cap + 64)
                n = C.c_int64()
                rc = f(bytes(b), len(b), out, cap, C.byref(n))
                assert out.raw[cap:] == bytes(64), "decoder wrote past the caller's capacity"
                if rc == 0:
                    assert n.value == declared
                    n_ok += 1
                else:
                    n_err += 1
    assert n_ok > 100 and n_err > 100   # both outcomes occur: the corpus is neither all-valid nor all-rejected
