"""Host batcher: BAM/BAI writer -> reader round trip, SA parsing rules, C-ABI surface of libgromhost."""
import os

import pytest

import numpy as np

from util import GOLDEN, golden_batches
from grom_b200 import hostlib
from grom_b200.reads import LAYOUT_CANONICAL_OFFSETS, LAYOUT_QUAL2, LAYOUT_QUAL4, LAYOUT_SEQ2, LAYOUT_SPARSE_SA, SA_FIELDS, _DTYPES, fnv1a64
from tools import synth


def test_roundtrip_all_fields(tmp_path):
    spec = synth.SynthSpec(contigs=[("c1", 60_000), ("c2", 30_000)], depth=12, seed=5, dup_frac=0.05, clip_frac=0.05)
    cs = synth.simulate(spec)
    fa, bam = synth.write_dataset(str(tmp_path / "rt"), cs)
    with hostlib.Bam(bam) as b:
        assert b.names == ["c1", "c2"] and b.lens == [60_000, 30_000] and b.has_index
        for tid, c in enumerate(cs):
            r = b.read_target(tid, keep_names=True)
            o = c.batch
            assert r.n_reads == o.n_reads
            for k in ["pos", "mpos", "tlen", "mtid", "l_qseq", "flag", "n_cigar", "mapq", "qname_len", "qname_hash", "cigar",
                      "sa_pos", "sa_strand", "sa_mapq", "sa_same_chr", "sa_start_adj", "sa_end_adj", "sa_end_adj_indel"]:
                assert np.array_equal(getattr(r, k), getattr(o, k)), k
            for i in range(0, r.n_reads, 211):
                assert np.array_equal(r.bases(i), o.bases(i)) and np.array_equal(r.quals(i), o.quals(i))
                assert r.qname(i) == o.qname(i)
            assert np.all(r.base_off % 32 == 0)
            assert (r.sa_pos >= 0).sum() > 0
            # the batcher also hands over the transport-compact forms (grom_reads.h GROM_LAYOUT_*): they decode to its canonical arrays
            assert r.layout_flags == LAYOUT_CANONICAL_OFFSETS | LAYOUT_QUAL2 | LAYOUT_SPARSE_SA | LAYOUT_SEQ2 and r.has_canonical_offsets()
            slot = np.arange(r.n_base_slots)
            nib = (1 << ((r.seq2[slot >> 2] >> ((~slot & 3) << 1)) & 3)).astype(np.uint8)
            nib[r.seq_exc_slot.astype(np.int64)] = r.seq_exc_code
            dec2 = r.qual_lut[(r.qual2[slot >> 2] >> ((~slot & 3) << 1)) & 3]
            for i in range(0, r.n_reads, 53):
                o = int(r.base_off[i]); assert np.array_equal(nib[o:o + int(r.l_qseq[i])], r.bases(i)) and np.array_equal(dec2[o:o + int(r.l_qseq[i])], r.quals(i))
            # ... and equal what the numpy side derives from the same canonical arrays
            mine = synth.slice_batch(r, 0, r.n_reads).compact()
            assert np.array_equal(mine.seq2, r.seq2) and np.array_equal(mine.seq_exc_slot, r.seq_exc_slot) and np.array_equal(mine.seq_exc_code, r.seq_exc_code)
            assert np.array_equal(mine.qual2, r.qual2) and np.array_equal(mine.qual_lut, r.qual_lut) and np.array_equal(mine.sa_index, r.sa_index)
            assert np.array_equal(r.sa_index, np.flatnonzero(r.sa_pos != -1))
            for k in SA_FIELDS:
                assert np.array_equal(r.sa_sparse[k], getattr(r, k)[r.sa_index]), k
            assert r.transport_bytes() < 0.7 * sum(getattr(r, k).nbytes for k in _DTYPES)


def test_reader_without_index(tmp_path):
    spec = synth.SynthSpec(contigs=[("c1", 30_000), ("c2", 20_000)], depth=8, seed=6)
    cs = synth.simulate(spec)
    fa, bam = synth.write_dataset(str(tmp_path / "ni"), cs)
    os.remove(bam + ".bai")
    with hostlib.Bam(bam) as b:
        assert not b.has_index
        for tid, c in enumerate(cs):
            r = b.read_target(tid)
            assert np.array_equal(r.pos, c.batch.pos) and np.array_equal(r.cigar, c.batch.cigar)


def test_index_beside_the_stem(tmp_path):
    """`x.bai` beside `x.bam` is found like `x.bam.bai` (the second place samtools' bam_index_load looks)."""
    spec = synth.SynthSpec(contigs=[("c1", 30_000), ("c2", 20_000)], depth=8, seed=7)
    cs = synth.simulate(spec)
    fa, bam = synth.write_dataset(str(tmp_path / "ix"), cs)
    os.rename(bam + ".bai", bam[:-1] + "i")
    with hostlib.Bam(bam) as b:
        assert b.has_index
        for tid, c in enumerate(cs):
            assert np.array_equal(b.read_target(tid).pos, c.batch.pos)


def test_bgzf_blocks_with_further_extra_subfields(tmp_path):
    """The gzip extra field of a BGZF block may hold subfields besides 'BC' (htslib writes 'BC' alone): the committed BAM re-packed with one
    before and one after it decodes to the same batches and statistics."""
    import struct
    import zlib
    from util import GOLDEN
    raw = open(os.path.join(GOLDEN, "g1.bam"), "rb").read()
    data, off = b"", 0
    while off + 18 <= len(raw):
        bs = struct.unpack_from("<H", raw, off + 16)[0] + 1
        data += zlib.decompressobj(-15).decompress(raw[off + 18:off + bs - 8]); off += bs

    def block(chunk, k):
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        d = co.compress(chunk) + co.flush()
        before = b"XY" + struct.pack("<H", 3) + b"abc" if k % 2 else b""
        after = b"ZZ" + struct.pack("<H", 0) if k % 3 == 0 else b""
        xlen = len(before) + 6 + len(after)
        bsize = 12 + xlen + len(d) + 8
        extra = before + b"BC" + struct.pack("<HH", 2, bsize - 1) + after
        return b"\x1f\x8b\x08\x04" + b"\0" * 6 + struct.pack("<H", xlen) + extra + d + struct.pack("<II", zlib.crc32(chunk), len(chunk))
    out = b"".join(block(data[i:i + 50_001], k) for k, i in enumerate(range(0, len(data), 50_001))) + block(b"", 7)
    p = tmp_path / "x.bam"
    p.write_bytes(out)                                                    # no index: the offsets of the original no longer apply
    with hostlib.Bam(os.path.join(GOLDEN, "g1.bam")) as a, hostlib.Bam(str(p)) as b:
        assert a.names == b.names and not b.has_index
        assert a.library_stats(20) == b.library_stats(20)
        for t in range(len(a.names)):
            x, y = a.read_target(t, keep_names=True), b.read_target(t, keep_names=True)
            assert x.n_reads == y.n_reads > 0
            for k in ("pos", "flag", "cigar", "seq4", "qual", "qname_hash", "sa_pos", "seq2", "qual2"):
                assert np.array_equal(getattr(x, k), getattr(y, k)), k


def test_record_counts_from_the_index(tmp_path):
    """The metadata pseudo-bin of the .bai gives the records per target without touching the BAM: the load measure of the contig -> GPU
    assignment (grom_b200.pipeline, tools/grom_b200.c).  Absent without an index."""
    from grom_b200.partition import assign_contigs
    spec = synth.SynthSpec(contigs=[("c1", 60_000), ("c2", 30_000), ("c3", 30_000)], depth=6, seed=12)
    cs = synth.simulate(spec)
    fa, bam = synth.write_dataset(str(tmp_path / "n"), cs)
    with hostlib.Bam(bam) as b:
        assert b.read_counts == [c.batch.n_reads for c in cs] and min(b.read_counts) > 0
        assert assign_contigs(b.lens, 2) == [[0], [1, 2]]
        # a thinly covered long contig no longer counts as the heaviest
        assert assign_contigs(b.lens, 2, [100.0, 900.0, 800.0]) == [[1], [2, 0]]
    os.remove(bam + ".bai")
    with hostlib.Bam(bam) as b:
        assert b.read_counts is None


TILAPIA_BAI = "/root/reference/test_data/tilapia_SAMD00023995_GL831235-1.bam.bai"


@pytest.mark.skipif(not os.path.exists(TILAPIA_BAI), reason="the reference's test data (build container only)")
def test_index_written_by_samtools(tmp_path):
    """The only real-world index at hand: the .bai of the reference's tilapia example (its BAM is a missing blob).  5,678 targets, reads on
    one of them; the metadata pseudo-bin, the chunk range and the trailing unplaced-read count parse, and the target's records are the
    130,504 the reference's run of that file saw."""
    import shutil
    import struct
    import zlib

    def bgzf(data):
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        d = c.compress(data) + c.flush()
        return b"\x1f\x8b\x08\x04" + b"\0" * 6 + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, len(d) + 25) + d + struct.pack("<II", zlib.crc32(data), len(data))
    n_ref = struct.unpack_from("<i", open(TILAPIA_BAI, "rb").read(8), 4)[0]
    assert n_ref == 5678
    head = b"BAM\1" + struct.pack("<i", 0) + struct.pack("<i", n_ref)
    for i in range(n_ref):
        name = (b"GL831235-1" if i == 29 else b"scaffold%d" % i) + b"\0"
        head += struct.pack("<i", len(name)) + name + struct.pack("<i", 2653313 if i == 29 else 1000)
    bam = tmp_path / "t.bam"
    bam.write_bytes(b"".join(bgzf(head[i:i + 60000]) for i in range(0, len(head), 60000)) + bgzf(b""))
    shutil.copy(TILAPIA_BAI, str(bam) + ".bai")
    with hostlib.Bam(str(bam)) as b:
        assert b.has_index and len(b.names) == n_ref and b.names[29] == "GL831235-1"
        L = hostlib.lib()
        import ctypes as C
        m, u = C.c_int64(), C.c_int64()
        assert L.gromhost_bam_target_reads(b._h, 29, C.byref(m), C.byref(u)) == 0 and (m.value, u.value) == (129050, 1454)
        assert L.gromhost_bam_target_reads(b._h, 0, C.byref(m), C.byref(u)) == 0 and (m.value, u.value) == (0, 0)     # no bins: no records
        assert b.read_counts is not None and sum(b.read_counts) == 130504 and b.read_counts[29] == 130504
        assert b.read_target(3).n_reads == 0                                               # no chunks: nothing is inflated


def test_golden_bam_decodes_and_hashes():
    names, batches = golden_batches()
    assert names == ["chrG", "chrH", "chrZ"]
    for b in batches:
        assert b.n_reads > 0 and np.all(np.diff(b.pos) >= 0)
        assert np.array_equal(b.qname_hash, fnv1a64(b.qname_pool, b.qname_off))


def test_empty_target(tmp_path):
    spec = synth.SynthSpec(contigs=[("c1", 20_000), ("c2", 20_000)], depth=5, seed=9)
    cs = synth.simulate(spec)
    cs[1].batch = None
    synth.write_fasta(str(tmp_path / "e.fa"), [(c.name, c.chars) for c in cs])
    hostlib.write_bam(str(tmp_path / "e.bam"), ["c1", "c2"], [20_000, 20_000], [cs[0].batch])
    with hostlib.Bam(str(tmp_path / "e.bam")) as b:
        assert b.read_target(1).n_reads == 0
        assert b.read_target(0).n_reads == cs[0].batch.n_reads


def test_corrupt_record_is_rejected(tmp_path):
    """A record whose block_size cannot hold its own name + CIGAR + bases (truncated / corrupt file) fails with a message instead of
    letting the fill pass read past the inflated data."""
    import struct
    import zlib
    from grom_b200 import hostlib

    def bgzf(data):
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        d = c.compress(data) + c.flush()
        hdr = b"\x1f\x8b\x08\x04" + b"\0" * 6 + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, len(d) + 25)
        return hdr + d + struct.pack("<II", zlib.crc32(data), len(data))
    text = b"@SQ\tSN:c\tLN:1000\n"
    head = b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", 1) + struct.pack("<i", 2) + b"c\0" + struct.pack("<i", 1000)
    # one record: block_size 40, but l_seq = 100 bases claimed
    rec = struct.pack("<iiiIIiiii", 40, 0, 10, (4680 << 16) | (60 << 8) | 2, (0 << 16) | 0, 100, -1, -1, 0) + b"r\0" + b"\0" * 6
    p = tmp_path / "bad.bam"
    p.write_bytes(bgzf(head) + bgzf(rec) + bgzf(b""))
    with hostlib.Bam(str(p)) as b:
        with pytest.raises(Exception, match="corrupt BAM record"):
            b.read_target(0)
