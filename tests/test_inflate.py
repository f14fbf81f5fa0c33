"""The host batcher's own DEFLATE decoder (grom_b200/host/inflate.c) against zlib: every block type and strategy, sizes around the
fast-loop margins, multi-block streams, wrong sizes, truncated and bit-flipped input (never a write outside the output buffer), and the
batcher's use of it (own decoder == zlib path; a CRC mismatch is refused)."""
import ctypes as C
import os
import random
import struct
import zlib

import numpy as np
import pytest

from grom_b200 import hostlib
from tools import synth


def _lib():
    L = hostlib.lib()
    L.gromhost_inflate_raw.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64]
    return L


def own(comp: bytes, n: int):
    out = C.create_string_buffer(n + 64)
    rc = _lib().gromhost_inflate_raw(comp, len(comp), out, n)
    return rc, out.raw[:n], out.raw[n:]


def payload(kind: int, n: int, rng: random.Random) -> bytes:
    if kind == 0:
        return bytes(rng.getrandbits(8) for _ in range(n))                                  # incompressible: stored blocks
    if kind == 1:
        return bytes(rng.choice(b"ACGT") for _ in range(n))                                  # literals only, short codes
    if kind == 2:
        return (b"ACGTTTGACA" * (n // 10 + 1))[:n]                                           # long matches at distance 10
    if kind == 3:
        return bytes(rng.getrandbits(8) if rng.random() < 0.1 else 70 for _ in range(n))     # runs: distance 1
    if kind == 4:
        return b"\0" * n
    if kind == 5:
        w = [bytes(rng.getrandbits(8) for _ in range(rng.randrange(1, 40))) for _ in range(50)]
        return b"".join(rng.choice(w) for _ in range(n // 10 + 1))[:n]                       # dictionary-like: all distances, long codes
    return bytes((i * 7 + (i >> 3)) & 255 for i in range(n))                                 # 256 distinct literals


def deflate(d: bytes, level: int, strategy: int = zlib.Z_DEFAULT_STRATEGY) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    return co.compress(d) + co.flush()


def test_decoder_equals_zlib_on_every_block_type():
    rng = random.Random(1)
    n_streams = 0
    for kind in range(7):
        for n in (0, 1, 3, 17, 258, 259, 329, 330, 331, 1000, 4096, 65280, 65536):
            d = payload(kind, n, rng)
            for level in (0, 1, 6, 9):
                for strat in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                    comp = deflate(d, level, strat)
                    rc, got, tail = own(comp, n)
                    assert rc == 0 and got == d and tail == b"\0" * 64, (kind, n, level, strat)
                    if n:
                        rc, _, tail = own(comp, n - 1)                       # one byte less room: refused, nothing behind the buffer touched
                        assert rc != 0 and tail[1:] == b"\0" * 63
                    assert own(comp, n + 1)[0] != 0                          # the stream ends early
                    if len(comp) > 2 and n:
                        assert own(comp[:len(comp) // 2], n)[0] != 0         # truncated input
                    n_streams += 1
    assert n_streams > 1000


def test_crc32_equals_zlib():
    L = _lib()
    L.gromhost_crc32.argtypes = [C.c_char_p, C.c_int64]
    L.gromhost_crc32.restype = C.c_uint32
    rng = random.Random(4)
    blob = bytes(rng.getrandbits(8) for _ in range(70_000))
    for n in list(range(0, 200)) + [255, 256, 1023, 4096, 65279, 65280, 65536] + [rng.randrange(70_000) for _ in range(200)]:
        o = rng.randrange(0, 70_000 - n + 1)
        assert L.gromhost_crc32(blob[o:o + n], n) == zlib.crc32(blob[o:o + n]), n


def test_variant_without_bmi2_decodes_the_same(monkeypatch):
    """grom_inflate_raw picks the BMI2 build of the loop where the CPU has it; GROMHOST_INFLATE=generic sends the test entry through the other."""
    rng = random.Random(5)
    monkeypatch.setenv("GROMHOST_INFLATE", "generic")
    for kind in range(7):
        for n in (300, 5000, 65536):
            d = payload(kind, n, rng)
            for level in (1, 6):
                rc, got, tail = own(deflate(d, level), n)
                assert rc == 0 and got == d and tail == b"\0" * 64


def test_multi_block_streams_with_flush_points():
    rng = random.Random(2)
    for kind in range(7):
        d = payload(kind, 60000, rng)
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = b""
        for k, i in enumerate(range(0, len(d), 7000)):
            comp += co.compress(d[i:i + 7000]) + co.flush(zlib.Z_SYNC_FLUSH if k % 2 else zlib.Z_FULL_FLUSH)    # empty stored blocks in between
        comp += co.flush()
        rc, got, _ = own(comp, len(d))
        assert rc == 0 and got == d


def test_bit_flips_never_escape_the_buffers():
    rng = random.Random(3)
    accepted = 0
    for _ in range(1500):
        d = payload(rng.randrange(7), rng.choice([300, 5000, 65536]), rng)
        comp = bytearray(deflate(d, rng.choice([1, 6])))
        for _ in range(rng.randrange(1, 4)):
            comp[rng.randrange(len(comp))] ^= 1 << rng.randrange(8)
        rc, got, tail = own(bytes(comp), len(d))
        assert tail == b"\0" * 64
        if rc == 0:                                                          # a harmless flip (or one the CRC of the BGZF trailer would catch)
            accepted += 1
            try:
                ref = zlib.decompressobj(-15).decompress(bytes(comp))
            except zlib.error:
                ref = None
            assert ref is None or ref[:len(d)] == got
    assert accepted < 1500


def test_batcher_is_the_same_with_zlib_and_refuses_a_crc_mismatch(tmp_path, monkeypatch):
    spec = synth.SynthSpec(contigs=[("c1", 80_000), ("c2", 20_000)], depth=15, seed=8, dup_frac=0.05, clip_frac=0.05, sa_frac=0.8, disc_frac=0.03)
    cs = synth.simulate(spec)
    fa, bam = synth.write_dataset(str(tmp_path / "x"), cs)
    with hostlib.Bam(bam) as b:
        mine = [b.read_target(t, keep_names=True) for t in range(2)]
    monkeypatch.setenv("GROMHOST_INFLATE", "zlib")
    with hostlib.Bam(bam) as b:
        theirs = [b.read_target(t, keep_names=True) for t in range(2)]
    monkeypatch.delenv("GROMHOST_INFLATE")
    for m, z in zip(mine, theirs):
        assert m.n_reads == z.n_reads > 0
        for k in ("pos", "flag", "cigar", "seq4", "qual", "qname_hash", "sa_pos", "seq2", "qual2", "sa_index", "seq_exc_slot"):
            assert np.array_equal(getattr(m, k), getattr(z, k)), k
    # flip one bit of the CRC-32 in the trailer of a block in the middle of the file
    raw = bytearray(open(bam, "rb").read())
    off, blocks = 0, []
    while off + 18 <= len(raw):
        bs = struct.unpack_from("<H", raw, off + 16)[0] + 1
        blocks.append((off, bs)); off += bs
    o, bs = blocks[len(blocks) // 2]
    raw[o + bs - 8] ^= 4
    bad = tmp_path / "bad.bam"
    bad.write_bytes(bytes(raw))
    os.link(bam + ".bai", str(bad) + ".bai")
    with hostlib.Bam(str(bad)) as b:
        with pytest.raises(RuntimeError, match="CRC"):
            for t in range(2):
                b.read_target(t)
