"""The deflate decoder the device runs (methyldackel_b200/csrc/inflate_hd.h, here compiled for the host: one lane) against
zlib on streams that exercise every block type and both copy paths: stored, fixed and dynamic Huffman blocks, several
deflate blocks in one stream, run-length matches (distance < length), matches at the maximum distance (beyond the shared-
memory ring: the far path reads bytes that were already flushed), codes longer than the primary tables, unaligned
output addresses and unaligned input offsets."""
import ctypes as C
import os
import random
import zlib

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(built):
    lib = C.CDLL(os.path.join(ROOT, "tests", "native", "libmdemu.so"))
    lib.emu_inflate_raw.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32]
    lib.emu_inflate_raw.restype = C.c_int
    return lib


def _deflate(data, level, strategy=zlib.Z_DEFAULT_STRATEGY, flush_every=0):
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    out = b""
    if flush_every:
        for k in range(0, len(data), flush_every):
            out += co.compress(data[k:k + flush_every]) + co.flush(zlib.Z_FULL_FLUSH)     # forces several deflate blocks (and stored empties)
    else:
        out += co.compress(data)
    return out + co.flush()


def _payloads():
    rnd = random.Random(7)
    yield "empty", b""
    yield "one_byte", b"A"
    yield "random_incompressible", bytes(rnd.getrandbits(8) for _ in range(65280))
    yield "rle", b"\x00" * 40000 + b"ab" * 10000 + b"xyz" * 1700
    base = bytes(rnd.getrandbits(8) for _ in range(32768 - 300))
    yield "max_distance", (base + bytes(rnd.getrandbits(8) for _ in range(300)) + base)[:65536]      # matches ~32 KB back: beyond the ring
    yield "text", (b"".join(b"f%011d\tchr%d\t%d\t%s\n" % (k, k % 24, k * 37, bytes(rnd.choice(b"ACGT") for _ in range(40))) for k in range(900)))[:65280]
    # many distinct byte values with a skewed distribution -> code lengths beyond 10 bits
    yield "skewed", bytes(min(255, int(rnd.expovariate(0.08))) for _ in range(60000))
    yield "bam_like", (b"".join(bytes([rnd.choice((37, 37, 37, 25, 11, 2)) for _ in range(150)]) + bytes(rnd.getrandbits(8) & 0x99 for _ in range(75)) for _ in range(280)))[:65280]


@pytest.mark.parametrize("name,data", list(_payloads()), ids=[n for n, _ in _payloads()])
@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_matches_zlib(emu, name, data, level):
    for strategy, flush_every in ((zlib.Z_DEFAULT_STRATEGY, 0), (zlib.Z_FIXED, 0), (zlib.Z_DEFAULT_STRATEGY, 9000), (zlib.Z_HUFFMAN_ONLY, 0)):
        comp = _deflate(data, level, strategy, flush_every)
        assert zlib.decompress(comp, -15) == data
        for in_off, out_off in ((0, 0), (3, 5), (1, 15), (2, 16)):
            buf = C.create_string_buffer(len(data) + 64)
            C.memset(buf, 0xEE, len(data) + 64)
            rc = emu.emu_inflate_raw(comp, len(comp), in_off, buf, out_off, len(data))
            assert rc == 0, (name, level, strategy, flush_every, in_off, out_off, rc)
            raw = buf.raw
            assert raw[out_off:out_off + len(data)] == data
            assert raw[:out_off] == b"\xee" * out_off and raw[out_off + len(data):] == b"\xee" * (64 - out_off)   # nothing outside the block is touched


def test_errors_are_reported(emu):
    data = bytes(range(256)) * 40
    comp = _deflate(data, 6)
    buf = C.create_string_buffer(len(data) + 64)
    assert emu.emu_inflate_raw(comp, len(comp), 0, buf, 0, len(data) - 1) != 0          # stream longer than the announced size
    assert emu.emu_inflate_raw(comp, len(comp), 0, buf, 0, len(data) + 1) != 0          # ... shorter
    bad = bytes([comp[0] | 0x06]) + comp[1:]                                            # block type 3
    assert emu.emu_inflate_raw(bad, len(bad), 0, buf, 0, len(data)) != 0
