"""The batcher lists the records of a target with all threads: entry points into the record chain are guessed and kept only when the walk
of the share before arrives exactly at them (grom_b200/host/bamio.c walk_records_parallel).  The result must not depend on the guesses:
forced on small files here (GROMHOST_WALK_PAR_MIN=1), with and without index, and on a file whose base qualities are built to look like
chains of records (every guess wrong -> the one-thread walk takes over)."""
import os
import struct

import numpy as np
import pytest

from grom_b200 import hostlib
from grom_b200.reads import _DTYPES
from tools import synth

FIELDS = list(_DTYPES) + ["seq2", "qual2", "sa_index", "seq_exc_slot", "seq_exc_code"]


def same(a, b):
    assert a.n_reads == b.n_reads and a.layout_flags == b.layout_flags
    for k in FIELDS:
        x, y = getattr(a, k), getattr(b, k)
        assert (x is None and y is None) or np.array_equal(x, y), k
    assert np.array_equal(a.qname_off, b.qname_off) and np.array_equal(a.qname_pool, b.qname_pool)


def read_all(bam, monkeypatch, par_min, threads):
    monkeypatch.setenv("GROMHOST_WALK_PAR_MIN", str(par_min))
    with hostlib.Bam(bam) as b:
        return [b.read_target(t, keep_names=True, threads=threads) for t in range(len(b.names))]


@pytest.mark.parametrize("indexed", [True, False])
def test_all_thread_walk_equals_the_one_thread_walk(tmp_path, monkeypatch, capfd, indexed):
    spec = synth.SynthSpec(contigs=[("c1", 60_000), ("c2", 150_000), ("c3", 40_000)], depth=15, seed=21, dup_frac=0.05, clip_frac=0.05, sa_frac=0.8, disc_frac=0.03)
    cs = synth.simulate(spec)
    fa, bam = synth.write_dataset(str(tmp_path / "w"), cs)
    if not indexed:
        os.remove(bam + ".bai")          # every target then starts at the first record of the file: foreign records lead, and follow
    one = read_all(bam, monkeypatch, 1 << 60, 4)
    monkeypatch.setenv("GROMHOST_TRACE", "1")
    for threads in (2, 4, 7):
        capfd.readouterr()
        par = read_all(bam, monkeypatch, 1, threads)
        err = capfd.readouterr().err
        assert f"record chain: {threads} shares" in err and "fallback=1" not in err
        for a, b in zip(one, par):
            same(a, b)
    for c, a in zip(cs, one):
        assert a.n_reads == c.batch.n_reads and np.array_equal(a.pos, c.batch.pos)


def test_wrong_guesses_fall_back_to_the_one_thread_walk(tmp_path, monkeypatch, capfd):
    """Base qualities that spell chains of well-formed records: a thread that starts looking inside a read finds the forged chain before the
    next real record; the share before it steps over that offset, and the walk is redone by one thread."""
    spec = synth.SynthSpec(contigs=[("c1", 120_000)], depth=12, seed=22, read_len=400, ins_mean=1000.0, ins_sd=60.0, ins_floor=500, clip_frac=0.0,
                           hardclip_frac=0.0, indel_every=10 ** 9, refskip_frac=0.0, sv_sites_per_mb=0.0, disc_frac=0.0)
    cs = synth.simulate(spec)
    bt = cs[0].batch
    fake = b""
    for j in range(9):                   # nine forged records of 40 bytes: block_size 36, tid 0, pos j, l_read_name 2, no CIGAR, no bases, name "x\0"
        fake += struct.pack("<iiiIIiiii", 36, 0, j, (4680 << 16) | (30 << 8) | 2, 0, 0, -1, -1, 0) + b"x\0\0\0"
    fake = np.frombuffer(fake, dtype=np.uint8)
    n_forged = 0
    for i in range(bt.n_reads):
        lq, o = int(bt.l_qseq[i]), int(bt.base_off[i])
        if lq >= len(fake) + 8:
            bt.qual[o + 4:o + 4 + len(fake)] = fake
            n_forged += 1
    assert n_forged > 0.9 * bt.n_reads
    fa, bam = synth.write_dataset(str(tmp_path / "f"), cs)
    one = read_all(bam, monkeypatch, 1 << 60, 4)
    monkeypatch.setenv("GROMHOST_TRACE", "1")
    capfd.readouterr()
    par = read_all(bam, monkeypatch, 1, 7)           # six guesses, each inside a read with probability ~0.4 of coming before its forged chain
    assert "fallback=1" in capfd.readouterr().err
    same(one[0], par[0])
    assert one[0].n_reads == bt.n_reads and np.array_equal(one[0].qual, bt.qual)


def test_corrupt_record_is_reported_by_the_all_thread_walk(tmp_path, monkeypatch):
    spec = synth.SynthSpec(contigs=[("c1", 100_000)], depth=12, seed=23)
    cs = synth.simulate(spec)
    bt = cs[0].batch
    fa, bam = synth.write_dataset(str(tmp_path / "c"), cs)
    # rewrite the file with l_seq of one record in the middle raised beyond its block_size (level-0 BGZF keeps the offsets simple)
    import zlib
    raw = open(bam, "rb").read()
    off, data = 0, b""
    while off + 18 <= len(raw):
        bs = struct.unpack_from("<H", raw, off + 16)[0] + 1
        data += zlib.decompressobj(-15).decompress(raw[off + 18:off + bs - 8]); off += bs
    l_text = struct.unpack_from("<i", data, 4)[0]
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", data, p)[0]; p += 4
    for _ in range(n_ref):
        ln = struct.unpack_from("<i", data, p)[0]; p += 4 + ln + 4
    recs = []
    while p + 4 <= len(data):
        bl = struct.unpack_from("<i", data, p)[0]; recs.append(p); p += 4 + bl
    victim = recs[len(recs) * 2 // 3]
    data = bytearray(data)
    struct.pack_into("<i", data, victim + 20, 100_000)

    def bgzf(chunk):
        co = zlib.compressobj(1, zlib.DEFLATED, -15)
        d = co.compress(bytes(chunk)) + co.flush()
        return (b"\x1f\x8b\x08\x04" + b"\0" * 6 + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, len(d) + 25) + d
                + struct.pack("<II", zlib.crc32(bytes(chunk)), len(chunk)))
    out = b"".join(bgzf(data[i:i + 60000]) for i in range(0, len(data), 60000)) + bgzf(b"")
    bad = tmp_path / "bad.bam"
    bad.write_bytes(out)
    for par_min in (1, 1 << 60):
        monkeypatch.setenv("GROMHOST_WALK_PAR_MIN", str(par_min))
        with hostlib.Bam(str(bad)) as b:
            with pytest.raises(RuntimeError, match="corrupt BAM record"):
                b.read_target(0, threads=4)


@pytest.mark.parametrize("indexed", [True, False])
def test_windows_do_not_change_the_batch(tmp_path, monkeypatch, indexed):
    """A target is decoded a window of BGZF blocks at a time (one window of inflated data in memory, the batch grows); records that straddle
    windows are carried over.  Windows of 1, 2 and 7 blocks -- most records straddle -- give the batch of one window over everything."""
    spec = synth.SynthSpec(contigs=[("c1", 50_000), ("c2", 120_000), ("c3", 30_000)], depth=14, seed=31, dup_frac=0.05, clip_frac=0.05, sa_frac=0.8, disc_frac=0.03,
                           long_name_frac=0.01)
    cs = synth.simulate(spec)
    fa, bam = synth.write_dataset(str(tmp_path / "w"), cs)
    if not indexed:
        os.remove(bam + ".bai")
    monkeypatch.setenv("GROMHOST_WINDOW_BLOCKS", "1000000")
    whole = read_all(bam, monkeypatch, 1 << 60, 3)
    for wb, par_min in (("1", 1 << 60), ("2", 1), ("7", 1)):
        monkeypatch.setenv("GROMHOST_WINDOW_BLOCKS", wb)
        got = read_all(bam, monkeypatch, par_min, 3)
        for a, b in zip(whole, got):
            same(a, b)
    for c, a in zip(cs, whole):
        assert a.n_reads == c.batch.n_reads and np.array_equal(a.pos, c.batch.pos) and np.array_equal(a.cigar, c.batch.cigar)


@pytest.mark.parametrize("window_blocks", ["1000000", "2"])
def test_target_in_pieces_equals_the_whole_target(tmp_path, monkeypatch, window_blocks):
    """gromhost_bam_iter_*: consecutive batches of at least max_reads records.  Laid end to end they are the batch of the whole target;
    every piece carries its own canonical offsets and transport-compact forms (what consecutive gromgpu_push_reads calls take)."""
    spec = synth.SynthSpec(contigs=[("c1", 40_000), ("c2", 110_000), ("c3", 30_000)], depth=14, seed=33, dup_frac=0.05, clip_frac=0.05, sa_frac=0.8, disc_frac=0.03)
    cs = synth.simulate(spec)
    fa, bam = synth.write_dataset(str(tmp_path / "p"), cs)
    monkeypatch.setenv("GROMHOST_WINDOW_BLOCKS", window_blocks)
    per_read = ["pos", "mpos", "tlen", "mtid", "l_qseq", "flag", "n_cigar", "mapq", "qname_len", "qname_hash", "sa_pos", "sa_strand", "sa_mapq",
                "sa_same_chr", "sa_start_adj", "sa_end_adj", "sa_end_adj_indel"]
    with hostlib.Bam(bam) as b:
        for tid in range(3):
            whole = b.read_target(tid, keep_names=True, threads=3)
            for max_reads in (1, 700, 10 ** 9):
                pieces = list(b.iter_target(tid, max_reads, keep_names=True, threads=3))
                pieces = [p for p in pieces if p.n_reads] or pieces[:1]
                assert sum(p.n_reads for p in pieces) == whole.n_reads
                if window_blocks == "2" and max_reads < 10 ** 9:
                    assert len(pieces) > 3 and all(p.n_reads >= max_reads for p in pieces[:-1])
                if max_reads == 10 ** 9:
                    assert len(pieces) == 1
                for k in per_read:
                    assert np.array_equal(np.concatenate([getattr(p, k) for p in pieces]), getattr(whole, k)), k
                assert np.array_equal(np.concatenate([p.cigar for p in pieces]), whole.cigar)
                i0 = 0
                for p in pieces:
                    assert p.has_canonical_offsets() and p.layout_flags == whole.layout_flags
                    for i in range(0, p.n_reads, 37):
                        assert np.array_equal(p.bases(i), whole.bases(i0 + i)) and np.array_equal(p.quals(i), whole.quals(i0 + i)) and p.qname(i) == whole.qname(i0 + i)
                    mine = synth.slice_batch(p, 0, p.n_reads).compact()
                    assert np.array_equal(mine.seq2, p.seq2) and np.array_equal(mine.seq_exc_slot, p.seq_exc_slot) and np.array_equal(mine.qual2, p.qual2)
                    assert np.array_equal(mine.qual_lut, p.qual_lut) and np.array_equal(mine.sa_index, p.sa_index)
                    i0 += p.n_reads
        # an empty target gives one empty piece
        assert [p.n_reads for p in b.iter_target(0, 5)][-1] >= 0


@pytest.mark.parametrize("case", ["late_value", "nine_values", "many_values", "late_fifth"])
def test_quality_forms_when_the_set_of_values_moves(tmp_path, monkeypatch, case):
    """2-bit qualities are packed window by window under the dictionary of the values seen so far; a value that first shows up in a later
    window (or a fifth one) makes the batcher pack again at the end / fall back to 4 bits or plain bytes.  Whatever happened on the way, the
    forms are the ones the numpy side derives from the finished canonical arrays."""
    from grom_b200.reads import LAYOUT_QUAL2, LAYOUT_QUAL4
    spec = synth.SynthSpec(contigs=[("c1", 90_000)], depth=14, seed=35, clip_frac=0.03)
    cs = synth.simulate(spec)
    bt = cs[0].batch
    real = bt.qual > 0
    if case == "late_value":                         # two values on the first 85 % of the slots, a third one only near the end
        bt.qual[real] = np.where(np.arange(real.sum()) % 3 == 0, 30, 37).astype(np.uint8)
        tail = np.flatnonzero(real)[int(real.sum() * 0.85)::5]
        bt.qual[tail] = 12
    elif case == "late_fifth":                       # four values throughout, a fifth near the end: 4-bit form
        idx = np.flatnonzero(real)
        bt.qual[idx] = np.array([12, 25, 30, 37], dtype=np.uint8)[np.arange(idx.size) % 4]
        bt.qual[idx[int(idx.size * 0.9)::7]] = 40
    elif case == "nine_values":
        idx = np.flatnonzero(real)
        bt.qual[idx] = (14 + np.arange(idx.size) % 9).astype(np.uint8)
    else:                                            # more than 16: the bytes travel
        idx = np.flatnonzero(real)
        bt.qual[idx] = (3 + np.arange(idx.size) % 23).astype(np.uint8)
    fa, bam = synth.write_dataset(str(tmp_path / "q"), cs)
    for wb in ("1", "3", "1000000"):
        monkeypatch.setenv("GROMHOST_WINDOW_BLOCKS", wb)
        with hostlib.Bam(bam) as b:
            r = b.read_target(0, threads=3)
        want = {"late_value": LAYOUT_QUAL2, "late_fifth": LAYOUT_QUAL4, "nine_values": LAYOUT_QUAL4, "many_values": 0}[case]
        assert r.layout_flags & (LAYOUT_QUAL2 | LAYOUT_QUAL4) == want, (case, wb)
        mine = synth.slice_batch(r, 0, r.n_reads).compact()
        assert mine.layout_flags == r.layout_flags
        if want == LAYOUT_QUAL2:
            assert np.array_equal(mine.qual2, r.qual2) and np.array_equal(mine.qual_lut, r.qual_lut) and r.qual4 is None
        elif want == LAYOUT_QUAL4:
            assert np.array_equal(mine.qual4, r.qual4) and np.array_equal(mine.qual_lut, r.qual_lut) and r.qual2 is None
        else:
            assert r.qual2 is None and r.qual4 is None
        for i in range(0, r.n_reads, 101):
            assert np.array_equal(r.quals(i), bt.quals(i))
