"""Transport-compact forms of the read batch (include/grom_reads.h GROM_LAYOUT_*): the host side is lossless."""
import ctypes as C

import numpy as np

from grom_b200.reads import LAYOUT_CANONICAL_OFFSETS, LAYOUT_QUAL4, LAYOUT_SPARSE_SA, SA_FIELDS, CReadBatch
from tools import synth


def _batch(**kw):
    spec = synth.SynthSpec(contigs=[("chrA", 60_000)], depth=12, seed=5, clip_frac=0.05, disc_frac=0.03, sa_frac=0.8, **kw)
    return synth.simulate(spec)[0].batch


def test_compact_forms_decode_to_the_canonical_arrays():
    raw = _batch()
    quals = [raw.quals(i).copy() for i in (0, 7, raw.n_reads - 1)]; bases = [raw.bases(i).copy() for i in (0, 7, raw.n_reads - 1)]
    cig = [raw.cigar_of(i) for i in (0, 7, raw.n_reads - 1)]
    b = raw.repack_canonical().compact()
    for j, i in enumerate((0, 7, b.n_reads - 1)):
        assert np.array_equal(b.quals(i), quals[j]) and np.array_equal(b.bases(i), bases[j]) and b.cigar_of(i) == cig[j]
    assert b.layout_flags == LAYOUT_CANONICAL_OFFSETS | LAYOUT_QUAL4 | LAYOUT_SPARSE_SA
    slot = np.arange(b.n_base_slots)
    code = (b.qual4[slot >> 1] >> ((~slot & 1) << 2)) & 15
    assert np.array_equal(b.qual_lut[code], b.qual)
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
    assert b.layout_flags == LAYOUT_SPARSE_SA and b.qual4 is None


def test_c_struct_carries_the_compact_forms():
    b = _batch().repack_canonical().compact()
    c = b.as_c()
    assert c.layout_flags == b.layout_flags and c.n_sa == len(b.sa_index) and c.qual4 == b.qual4.ctypes.data
    assert bytes(c.qual_lut) == bytes(b.qual_lut) and c.sas_pos == b.sa_sparse["sa_pos"].ctypes.data and c.sa_pos == b.sa_pos.ctypes.data
    # appended behind the canonical fields: a caller built against the older header sets layout_flags = 0 and is unaffected
    assert CReadBatch.layout_flags.offset == 28 and CReadBatch.qual4.offset == CReadBatch.qname_pool.offset + C.sizeof(C.c_void_p)
