"""The whole-buffer DEFLATE decoder of libbamcols (alntools_b200/csrc/fast_inflate.h) against zlib: every
stream zlib's deflate produces (all levels and strategies, stored / fixed / dynamic blocks, tiny and empty
inputs, long matches, distance-1 runs) must come back byte for byte, and a damaged stream must never be
accepted - the decoder says no and zlib judges the block, which is what the BGZF reader relies on."""
import zlib

import numpy as np
import pytest

from alntools_b200 import bamcols


def _raw(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=-15, memlevel=8):
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, memlevel, strategy)
    return c.compress(data) + c.flush()


def _payloads():
    rng = np.random.default_rng(5)
    yield b""
    yield b"a"
    yield b"ab" * 40000
    yield b"\x00" * 70000                                       # distance 1, length 258 runs
    yield bytes(rng.integers(0, 256, 65536, dtype=np.uint8))     # incompressible: stored blocks
    yield bytes(rng.integers(0, 4, 65280, dtype=np.uint8))       # short codes only
    yield bytes(rng.integers(65, 91, 3000, dtype=np.uint8)) * 20
    words = [bytes(rng.integers(97, 123, int(rng.integers(2, 12)), dtype=np.uint8)) for _ in range(500)]
    yield b" ".join(words[int(i)] for i in rng.integers(0, 500, 20000))
    # BAM-like records: fixed header fields, names counting up, random tails
    recs = []
    for i in range(1500):
        recs.append(np.array([i % 200, i * 3, 0x1248000a, 0x00100000, 36, -1, -1, 0], dtype="<i4").tobytes()
                    + b"read%09d\x00" % (i // 2) + bytes(rng.integers(0, 256, int(rng.integers(0, 30)), dtype=np.uint8)))
    yield b"".join(recs)
    skew = np.minimum(rng.zipf(1.2, 60000), 255).astype(np.uint8)   # long codes (15 bits) for rare symbols
    yield bytes(skew)


def test_every_zlib_stream_round_trips():
    n = 0
    for data in _payloads():
        for level in (0, 1, 4, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
                comp = _raw(data, level, strategy)
                got = bamcols.inflate_raw(comp, len(data), mode=2)
                assert got == data, (len(data), level, strategy)
                assert bamcols.inflate_raw(comp, len(data), mode=0) == data
                n += 1
    assert n == 250
    # small windows and memory levels change the block structure
    data = b"".join(b"%d," % (i * i) for i in range(30000))
    for wbits in (-9, -12, -15):
        for memlevel in (1, 5, 9):
            assert bamcols.inflate_raw(_raw(data, 6, wbits=wbits, memlevel=memlevel), len(data), mode=2) == data


def test_wrong_sizes_and_damaged_streams_are_never_accepted():
    rng = np.random.default_rng(9)
    data = b" ".join(b"%d" % int(v) for v in rng.integers(0, 5000, 8000))
    comp = _raw(data, 6)
    assert bamcols.inflate_raw(comp, len(data), mode=2) == data
    for wrong in (len(data) - 1, len(data) + 1, 1):          # (size 0 is the BGZF EOF block: nothing is decoded)
        assert bamcols.inflate_raw(comp, wrong, mode=2) is None
        assert bamcols.inflate_raw(comp, wrong, mode=0) is None
    for cut in (1, 2, 5, len(comp) // 2, len(comp) - 1):
        assert bamcols.inflate_raw(comp[:cut], len(data), mode=2) is None
        assert bamcols.inflate_raw(comp[:cut], len(data), mode=0) is None
    # random damage: whatever the whole-buffer decoder accepts must be what zlib produces
    accepted = 0
    for trial in range(400):
        bad = bytearray(comp)
        for _ in range(int(rng.integers(1, 4))):
            bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
        got = bamcols.inflate_raw(bytes(bad), len(data), mode=2)
        want = bamcols.inflate_raw(bytes(bad), len(data), mode=1)
        if got is not None:
            accepted += 1
            assert got == want
        assert bamcols.inflate_raw(bytes(bad), len(data), mode=0) == want
    # pure noise
    for trial in range(200):
        noise = bytes(rng.integers(0, 256, int(rng.integers(1, 300)), dtype=np.uint8))
        got = bamcols.inflate_raw(noise, 1000, mode=2)
        assert got is None or got == bamcols.inflate_raw(noise, 1000, mode=1)


def test_multi_block_streams_with_flush_points():
    """Streams cut into several blocks by Z_SYNC_FLUSH / Z_FULL_FLUSH (each leaves an empty stored block) and
    random parameters; 300 seeded cases of the fuzz loop that ran 30 000 streams without a mismatch."""
    rng = np.random.default_rng(2024)
    for case in range(300):
        size = int(rng.integers(0, 70000))
        kind = int(rng.integers(0, 4))
        if kind == 0:
            data = bytes(rng.integers(0, 256, size, dtype=np.uint8))
        elif kind == 1:
            data = bytes(rng.integers(0, int(rng.integers(1, 20)), size, dtype=np.uint8))
        elif kind == 2:
            data = bytes(np.minimum(rng.zipf(float(rng.uniform(1.05, 2.0)), size), 255).astype(np.uint8))
        else:
            base = bytes(rng.integers(0, 256, int(rng.integers(1, 400)), dtype=np.uint8))
            data = (base * (size // len(base) + 1))[:size]
        c = zlib.compressobj(int(rng.integers(0, 10)), zlib.DEFLATED, -int(rng.integers(9, 16)), int(rng.integers(1, 10)),
                             int(rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE])))
        parts, pos = [], 0
        while pos < len(data):
            step = int(rng.integers(1, len(data) + 1))
            parts.append(c.compress(data[pos:pos + step]))
            pos += step
            if rng.random() < 0.3:
                parts.append(c.flush(int(rng.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH]))))
        parts.append(c.flush())
        assert bamcols.inflate_raw(b"".join(parts), len(data), mode=2) == data, case
