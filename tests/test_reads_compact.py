"""Transport-compact forms of the read batch (include/grom_reads.h GROM_LAYOUT_*): the host side is lossless."""
import ctypes as C

import numpy as np

from grom_b200.reads import LAYOUT_CANONICAL_OFFSETS, LAYOUT_QUAL2, LAYOUT_QUAL4, LAYOUT_SEQ2, LAYOUT_SPARSE_SA, SA_FIELDS, CReadBatch
from tools import synth


def _batch(**kw):
    spec = synth.SynthSpec(contigs=[("chrA", 60_000)], depth=12, seed=5, clip_frac=0.05, disc_frac=0.03, sa_frac=0.8, **kw)
    return synth.simulate(spec)[0].batch


ALL = LAYOUT_CANONICAL_OFFSETS | LAYOUT_QUAL2 | LAYOUT_SPARSE_SA | LAYOUT_SEQ2       # four distinct base qualities in the synthetic data


def more_qualities(b, n_values=9):
    """Spread the base qualities over n_values distinct values (4-bit dictionary territory)."""
    idx = np.arange(0, b.qual.size, 3)
    b.qual[idx] = np.where(b.qual[idx] > 0, 14 + ((idx // 3) % n_values).astype(np.uint8), 0)
    return b


def plant_non_acgt(b, seed=3, n_runs=300):
    """Runs of N / IUPAC codes inside reads (neighbouring exceptions share bytes and words of the 4-bit array)."""
    rng = np.random.default_rng(seed)
    for i in rng.integers(0, b.n_reads, n_runs):
        lq = int(b.l_qseq[i])
        if lq < 12:
            continue
        k0 = int(rng.integers(0, lq - 10)); ln = int(rng.integers(1, 10))
        for k in range(k0, k0 + ln):
            sl = int(b.base_off[i]) + k
            code = 15 if rng.random() < 0.7 else int(rng.choice([0, 3, 5, 6, 7, 9, 10, 11, 12, 13, 14]))
            sh = ((~sl) & 1) << 2
            b.seq4[sl >> 1] = (int(b.seq4[sl >> 1]) & ~(15 << sh) & 0xff) | (code << sh)
    return b


def test_compact_forms_decode_to_the_canonical_arrays():
    raw = plant_non_acgt(_batch())
    quals = [raw.quals(i).copy() for i in (0, 7, raw.n_reads - 1)]; bases = [raw.bases(i).copy() for i in (0, 7, raw.n_reads - 1)]
    cig = [raw.cigar_of(i) for i in (0, 7, raw.n_reads - 1)]
    b = raw.repack_canonical().compact()
    for j, i in enumerate((0, 7, b.n_reads - 1)):
        assert np.array_equal(b.quals(i), quals[j]) and np.array_equal(b.bases(i), bases[j]) and b.cigar_of(i) == cig[j]
    assert b.layout_flags == ALL
    slot = np.arange(b.n_base_slots)
    two = (b.seq2[slot >> 2] >> ((~slot & 3) << 1)) & 3
    nib = (1 << two).astype(np.uint8)
    nib[b.seq_exc_slot.astype(np.int64)] = b.seq_exc_code
    assert len(b.seq_exc_slot) > 500 and np.all(np.diff(b.seq_exc_slot.astype(np.int64)) > 0)
    for i in range(0, b.n_reads, 97):
        o = int(b.base_off[i]); assert np.array_equal(nib[o:o + int(b.l_qseq[i])], b.bases(i))
    dec2 = b.qual_lut[(b.qual2[slot >> 2] >> ((~slot & 3) << 1)) & 3]
    for i in range(0, b.n_reads, 97):
        o = int(b.base_off[i]); assert np.array_equal(dec2[o:o + int(b.l_qseq[i])], b.quals(i))
    b4 = more_qualities(_batch().repack_canonical()).compact()
    assert b4.layout_flags == (ALL & ~LAYOUT_QUAL2) | LAYOUT_QUAL4 and len(np.unique(b4.qual)) > 8
    slot4 = np.arange(b4.n_base_slots)
    assert np.array_equal(b4.qual_lut[(b4.qual4[slot4 >> 1] >> ((~slot4 & 1) << 2)) & 15], b4.qual)
    for k in SA_FIELDS:
        dense = np.full(b.n_reads, -1 if k in ("sa_pos", "sa_mapq") else 0, dtype=getattr(b, k).dtype)
        dense[b.sa_index] = b.sa_sparse[k]
        assert np.array_equal(dense, getattr(b, k)), k
    assert np.all(np.diff(b.sa_index) > 0)
    full = sum(getattr(b, k).nbytes for k in ("pos", "mpos", "tlen", "mtid", "l_qseq", "flag", "n_cigar", "mapq", "qname_len", "qname_hash",
                                              "cigar_off", "base_off", "cigar", "seq4", "qual", *SA_FIELDS))
    assert b.transport_bytes() < 0.7 * full


def test_compact_is_refused_where_it_would_lose_information():
    b = _batch().repack_canonical()
    b.qual[:20] = np.arange(20, dtype=np.uint8) + 1          # > 16 distinct qualities
    b.base_off = b.base_off + np.uint64(32)                   # not the canonical running sum
    b.compact()
    assert b.layout_flags == LAYOUT_SPARSE_SA and b.qual4 is None and b.qual2 is None and b.seq2 is None


def test_c_struct_carries_the_compact_forms():
    b = _batch().repack_canonical().compact()
    c = b.as_c()
    assert c.layout_flags == b.layout_flags and c.n_sa == len(b.sa_index) and c.qual2 == b.qual2.ctypes.data
    assert c.seq2 == b.seq2.ctypes.data and c.n_seq_exc == len(b.seq_exc_slot)
    assert bytes(c.qual_lut) == bytes(b.qual_lut) and c.sas_pos == b.sa_sparse["sa_pos"].ctypes.data and c.sa_pos == b.sa_pos.ctypes.data
    # appended behind the canonical fields: a caller built against the older header sets layout_flags = 0 and is unaffected
    assert CReadBatch.layout_flags.offset == 28 and CReadBatch.qual4.offset == CReadBatch.qname_pool.offset + C.sizeof(C.c_void_p)
