"""Packed read-record batch (numpy side of include/grom_reads.h).

`ReadBatch` owns one numpy array per field of the C struct `grom_read_batch` and
can present itself as that struct (ctypes) to the host library
(include/gromhost.h) and to the CUDA library (include/gromgpu.h).  Field meaning
follows the reference's per-record intake, reference src/GROM.c:5743-5824.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

BASE_ALIGN = 32

# BAM CIGAR op codes
CMATCH, CINS, CDEL, CREF_SKIP, CSOFT_CLIP, CHARD_CLIP, CPAD, CEQUAL, CDIFF = range(9)
# BAM 4-bit base codes
NT16 = "=ACMGRSVTWYHKDBN"
CODE_A, CODE_C, CODE_G, CODE_T, CODE_N = 1, 2, 4, 8, 15

FPAIRED, FPROPER, FUNMAP, FMUNMAP, FREVERSE, FMREVERSE, FREAD1, FREAD2, FSECONDARY, FQCFAIL, FDUP = (
    1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024)


class CReadBatch(C.Structure):
    _fields_ = [
        ("n_reads", C.c_int64), ("n_cigar_total", C.c_int64), ("n_base_slots", C.c_int64),
        ("tid", C.c_int32), ("layout_flags", C.c_int32),
        ("pos", C.c_void_p), ("mpos", C.c_void_p), ("tlen", C.c_void_p), ("mtid", C.c_void_p),
        ("l_qseq", C.c_void_p), ("flag", C.c_void_p), ("n_cigar", C.c_void_p), ("mapq", C.c_void_p),
        ("qname_len", C.c_void_p), ("qname_hash", C.c_void_p), ("cigar_off", C.c_void_p),
        ("base_off", C.c_void_p), ("cigar", C.c_void_p), ("seq4", C.c_void_p), ("qual", C.c_void_p),
        ("sa_pos", C.c_void_p), ("sa_start_adj", C.c_void_p), ("sa_end_adj", C.c_void_p),
        ("sa_end_adj_indel", C.c_void_p), ("sa_strand", C.c_void_p), ("sa_mapq", C.c_void_p),
        ("sa_same_chr", C.c_void_p), ("qname_off", C.c_void_p), ("qname_pool", C.c_void_p),
        # transport-compact forms (include/grom_reads.h: GROM_LAYOUT_*)
        ("qual4", C.c_void_p), ("qual_lut", C.c_uint8 * 16), ("n_sa", C.c_int64), ("sa_index", C.c_void_p),
        ("sas_pos", C.c_void_p), ("sas_start_adj", C.c_void_p), ("sas_end_adj", C.c_void_p), ("sas_end_adj_indel", C.c_void_p),
        ("sas_mapq", C.c_void_p), ("sas_strand", C.c_void_p), ("sas_same_chr", C.c_void_p),
        ("seq2", C.c_void_p), ("n_seq_exc", C.c_int64), ("seq_exc_slot", C.c_void_p), ("seq_exc_code", C.c_void_p),
        ("qual2", C.c_void_p),
    ]


LAYOUT_CANONICAL_OFFSETS, LAYOUT_QUAL4, LAYOUT_SPARSE_SA, LAYOUT_SEQ2, LAYOUT_QUAL2 = 1, 2, 4, 8, 16
SA_FIELDS = ["sa_pos", "sa_start_adj", "sa_end_adj", "sa_end_adj_indel", "sa_strand", "sa_mapq", "sa_same_chr"]
SA_NONE = {"sa_pos": -1, "sa_mapq": -1}         # what a read without SA / XP entry carries (everything else 0)


_DTYPES = {
    "pos": np.int32, "mpos": np.int32, "tlen": np.int32, "mtid": np.int32, "l_qseq": np.int32,
    "flag": np.uint16, "n_cigar": np.uint16, "mapq": np.uint8, "qname_len": np.uint8,
    "qname_hash": np.uint64, "cigar_off": np.uint64, "base_off": np.uint64, "cigar": np.uint32,
    "seq4": np.uint8, "qual": np.uint8, "sa_pos": np.int32, "sa_start_adj": np.int32,
    "sa_end_adj": np.int32, "sa_end_adj_indel": np.int32, "sa_strand": np.uint8,
    "sa_mapq": np.int16, "sa_same_chr": np.uint8,
}
PER_READ = ["pos", "mpos", "tlen", "mtid", "l_qseq", "flag", "n_cigar", "mapq", "qname_len", "qname_hash",
            "cigar_off", "base_off", "sa_pos", "sa_start_adj", "sa_end_adj", "sa_end_adj_indel",
            "sa_strand", "sa_mapq", "sa_same_chr"]


def fnv1a64(names_pool: np.ndarray, off: np.ndarray) -> np.ndarray:
    """Vectorised FNV-1a over ragged byte strings (pool + offsets; NUL excluded)."""
    n = len(off) - 1
    h = np.full(n, 1469598103934665603, dtype=np.uint64)
    lens = (off[1:] - off[:-1]).astype(np.int64) - 1       # stored with trailing NUL
    maxlen = int(lens.max()) if n else 0
    prime = np.uint64(1099511628211)
    with np.errstate(over="ignore"):
        for k in range(maxlen):
            m = lens > k
            idx = (off[:-1][m] + np.uint64(k)).astype(np.int64)
            h[m] = (h[m] ^ names_pool[idx].astype(np.uint64)) * prime
    return h


@dataclass
class ReadBatch:
    tid: int
    pos: np.ndarray
    mpos: np.ndarray
    tlen: np.ndarray
    mtid: np.ndarray
    l_qseq: np.ndarray
    flag: np.ndarray
    n_cigar: np.ndarray
    mapq: np.ndarray
    qname_len: np.ndarray
    qname_hash: np.ndarray
    cigar_off: np.ndarray
    base_off: np.ndarray
    cigar: np.ndarray
    seq4: np.ndarray
    qual: np.ndarray
    sa_pos: np.ndarray
    sa_start_adj: np.ndarray
    sa_end_adj: np.ndarray
    sa_end_adj_indel: np.ndarray
    sa_strand: np.ndarray
    sa_mapq: np.ndarray
    sa_same_chr: np.ndarray
    qname_off: Optional[np.ndarray] = None      # uint64 [n+1], names stored NUL-terminated
    qname_pool: Optional[np.ndarray] = None     # uint8
    aux_off: Optional[np.ndarray] = None        # uint64 [n+1]  raw BAM aux bytes (tooling only)
    aux_pool: Optional[np.ndarray] = None       # uint8
    # transport-compact forms (compact()): what the CUDA library uploads instead of the canonical arrays
    layout_flags: int = 0
    qual4: Optional[np.ndarray] = None          # uint8, two base slots per byte (nibble order of seq4), indices into qual_lut
    qual_lut: Optional[np.ndarray] = None       # uint8 [16]
    sa_index: Optional[np.ndarray] = None       # int32, reads that have an SA / XP entry
    sa_sparse: Optional[dict] = None            # SA_FIELDS -> arrays of len(sa_index)
    qual2: Optional[np.ndarray] = None          # uint8, four base slots per byte, indices into qual_lut[0..3] (preferred over qual4)
    seq2: Optional[np.ndarray] = None           # uint8, four base slots per byte (A C G T = 0..3), first slot in the top bits
    seq_exc_slot: Optional[np.ndarray] = None   # uint64 base slots holding something other than A/C/G/T ...
    seq_exc_code: Optional[np.ndarray] = None   # uint8  ... and its BAM 4-bit code
    _keep: list = field(default_factory=list, repr=False)

    @property
    def n_reads(self) -> int:
        return int(self.pos.shape[0])

    @property
    def n_base_slots(self) -> int:
        return int(self.qual.shape[0])

    def normalise(self) -> "ReadBatch":
        for k, dt in _DTYPES.items():
            setattr(self, k, np.ascontiguousarray(getattr(self, k), dtype=dt))
        if self.qname_off is not None:
            self.qname_off = np.ascontiguousarray(self.qname_off, dtype=np.uint64)
            self.qname_pool = np.ascontiguousarray(self.qname_pool, dtype=np.uint8)
        if self.aux_off is not None:
            self.aux_off = np.ascontiguousarray(self.aux_off, dtype=np.uint64)
            self.aux_pool = np.ascontiguousarray(self.aux_pool, dtype=np.uint8)
        n = self.n_reads
        for k in PER_READ:
            assert getattr(self, k).shape[0] == n, (k, getattr(self, k).shape, n)
        assert self.qual.shape[0] % 2 == 0 and self.seq4.shape[0] >= self.qual.shape[0] // 2
        return self

    def as_c(self) -> CReadBatch:
        self.normalise()
        c = CReadBatch()
        c.n_reads = self.n_reads
        c.n_cigar_total = int(self.cigar.shape[0])
        c.n_base_slots = self.n_base_slots
        c.tid = int(self.tid)
        for k in _DTYPES:
            setattr(c, k, getattr(self, k).ctypes.data)
        c.qname_off = self.qname_off.ctypes.data if self.qname_off is not None else None
        c.qname_pool = self.qname_pool.ctypes.data if self.qname_pool is not None else None
        c.layout_flags = int(self.layout_flags)
        if self.layout_flags & LAYOUT_QUAL4:
            c.qual4 = self.qual4.ctypes.data
            for k in range(16):
                c.qual_lut[k] = int(self.qual_lut[k])
        if self.layout_flags & LAYOUT_QUAL2:
            c.qual2 = self.qual2.ctypes.data
            for k in range(16):
                c.qual_lut[k] = int(self.qual_lut[k])
        if self.layout_flags & LAYOUT_SEQ2:
            c.seq2 = self.seq2.ctypes.data
            c.n_seq_exc = int(self.seq_exc_slot.shape[0])
            c.seq_exc_slot = self.seq_exc_slot.ctypes.data
            c.seq_exc_code = self.seq_exc_code.ctypes.data
        if self.layout_flags & LAYOUT_SPARSE_SA:
            c.n_sa = int(self.sa_index.shape[0])
            c.sa_index = self.sa_index.ctypes.data
            for k in SA_FIELDS:
                setattr(c, "sas_" + k[3:], self.sa_sparse[k].ctypes.data)
        return c

    def has_canonical_offsets(self) -> bool:
        """cigar_off / base_off are the running sums the batcher produces (GROM_LAYOUT_CANONICAL_OFFSETS)."""
        n = self.n_reads
        if n == 0:
            return True
        co = np.concatenate([[0], np.cumsum(self.n_cigar.astype(np.uint64))[:-1]]).astype(np.uint64)
        pad = (self.l_qseq.astype(np.int64) + BASE_ALIGN - 1) // BASE_ALIGN * BASE_ALIGN
        bo = np.concatenate([[0], np.cumsum(pad)[:-1]]).astype(np.uint64)
        return bool(np.array_equal(co, self.cigar_off) and np.array_equal(bo, self.base_off))

    def repack_canonical(self) -> "ReadBatch":
        """Same reads with cigar[] / base slots laid out back to back (each read's slots rounded up to BASE_ALIGN): the
        layout the host batcher produces.  Tooling for batches assembled by other means (compiled loops: tools-side only)."""
        self.normalise()
        lq = self.l_qseq.astype(np.int64)
        pad = (lq + BASE_ALIGN - 1) // BASE_ALIGN * BASE_ALIGN
        bo = np.concatenate([[0], np.cumsum(pad)]).astype(np.int64)
        nc = self.n_cigar.astype(np.int64)
        co = np.concatenate([[0], np.cumsum(nc)]).astype(np.int64)
        seq4 = np.zeros(int(bo[-1]) // 2, dtype=np.uint8); qual = np.zeros(int(bo[-1]), dtype=np.uint8)
        _jit()["repack"](self.base_off.astype(np.int64), lq, bo, self.seq4, self.qual, seq4, qual)
        cw = np.arange(int(nc.sum())) - np.repeat(co[:-1], nc)
        self.cigar = self.cigar[np.repeat(self.cigar_off.astype(np.int64), nc) + cw]
        self.seq4 = seq4
        self.qual = qual
        self.cigar_off, self.base_off = co[:-1].astype(np.uint64), bo[:-1].astype(np.uint64)
        self.layout_flags = 0
        return self.normalise()

    def compact(self) -> "ReadBatch":
        """Attach the transport-compact forms the data allow (lossless; the canonical arrays stay in place for host users):
        offsets derived on the device, 2- or 4-bit dictionary-coded qualities when the batch holds <= 4 / <= 16 distinct values,
        2-bit bases with an exception list, and the first-SA-entry fields only for the reads that have one."""
        self.normalise()
        flags = 0
        if self.has_canonical_offsets():
            flags |= LAYOUT_CANONICAL_OFFSETS
        ns = int(self.qual.size)
        canon = bool(flags & LAYOUT_CANONICAL_OFFSETS) and ns > 0 and ns % 4 == 0
        if canon:
            J = _jit()
            lq = self.l_qseq.astype(np.int64)
            bo = self.base_off.astype(np.int64)
            hist, exc_cnt = J["survey"](bo, lq, self.seq4, self.qual)            # qualities on the bases themselves (padding aside), non-ACGT codes per read
            base_vals = np.flatnonzero(hist).astype(np.uint8)
            if 0 < base_vals.size <= 4:
                # <= 4 distinct qualities: 2 bits per slot; the device zeroes the padding again
                lut = np.zeros(16, dtype=np.uint8); lut[:base_vals.size] = base_vals
                inv = np.zeros(256, dtype=np.uint8); inv[base_vals] = np.arange(base_vals.size, dtype=np.uint8)
                self.qual2 = np.zeros(ns // 4, dtype=np.uint8); self.qual_lut = lut
                J["pack_qual2"](bo, lq, self.qual, inv, self.qual2)
                flags |= LAYOUT_QUAL2
            n_exc = int(exc_cnt.sum())
            if n_exc <= ns // 16:
                first = np.concatenate([[0], np.cumsum(exc_cnt)]).astype(np.int64)
                self.seq2 = np.zeros(ns // 4, dtype=np.uint8)
                self.seq_exc_slot = np.zeros(n_exc, dtype=np.uint64); self.seq_exc_code = np.zeros(n_exc, dtype=np.uint8)
                J["pack_seq2"](bo, lq, self.seq4, first, self.seq2, self.seq_exc_slot, self.seq_exc_code)
                flags |= LAYOUT_SEQ2
        if not (flags & LAYOUT_QUAL2) and ns and ns % 2 == 0:
            vals = np.flatnonzero(np.bincount(self.qual, minlength=256)).astype(np.uint8)
            if 0 < vals.size <= 16:
                lut = np.zeros(16, dtype=np.uint8); lut[:vals.size] = vals
                inv = np.zeros(256, dtype=np.uint8); inv[vals] = np.arange(vals.size, dtype=np.uint8)
                self.qual4 = np.zeros(ns // 2, dtype=np.uint8)
                step = 1 << 28
                for s0 in range(0, ns, step):
                    code = inv[self.qual[s0:s0 + step]]
                    self.qual4[s0 // 2:(s0 + code.size) // 2] = (code[0::2] << 4) | code[1::2]
                self.qual_lut = lut
                flags |= LAYOUT_QUAL4
        has = np.zeros(self.n_reads, dtype=bool)
        for k in SA_FIELDS:
            has |= getattr(self, k) != SA_NONE.get(k, 0)
        self.sa_index = np.flatnonzero(has).astype(np.int32)
        self.sa_sparse = {k: np.ascontiguousarray(getattr(self, k)[self.sa_index]) for k in SA_FIELDS}
        flags |= LAYOUT_SPARSE_SA
        self.layout_flags = flags
        return self

    def transport_bytes(self) -> int:
        """Bytes the CUDA library copies host -> device for this batch (gromgpu_push_reads)."""
        f = self.layout_flags
        skip = set()
        if f & LAYOUT_CANONICAL_OFFSETS:
            skip |= {"cigar_off", "base_off"}
        if f & (LAYOUT_QUAL4 | LAYOUT_QUAL2):
            skip.add("qual")
        if f & LAYOUT_SEQ2:
            skip.add("seq4")
        if f & LAYOUT_SPARSE_SA:
            skip |= set(SA_FIELDS)
        n = sum(getattr(self, k).nbytes for k in _DTYPES if k not in skip)
        if f & LAYOUT_QUAL2:
            n += self.qual2.nbytes
        elif f & LAYOUT_QUAL4:
            n += self.qual4.nbytes
        if f & LAYOUT_SEQ2:
            n += self.seq2.nbytes + self.seq_exc_slot.nbytes + self.seq_exc_code.nbytes
        if f & LAYOUT_SPARSE_SA:
            n += self.sa_index.nbytes + sum(v.nbytes for v in self.sa_sparse.values())
        return int(n)

    # ------------------------------------------------------------------ helpers
    def bases(self, i: int) -> np.ndarray:
        """4-bit codes of read i."""
        o = int(self.base_off[i]); n = int(self.l_qseq[i])
        sl = np.arange(o, o + n)
        return (self.seq4[sl >> 1] >> ((~sl & 1) << 2)) & 15

    def quals(self, i: int) -> np.ndarray:
        o = int(self.base_off[i]); return self.qual[o:o + int(self.l_qseq[i])]

    def cigar_of(self, i: int):
        o = int(self.cigar_off[i]); c = self.cigar[o:o + int(self.n_cigar[i])]
        return [(int(x) & 15, int(x) >> 4) for x in c]

    def qname(self, i: int) -> str:
        a, b = int(self.qname_off[i]), int(self.qname_off[i + 1])
        return bytes(self.qname_pool[a:b]).split(b"\0")[0].decode()

    def aligned_bases(self) -> int:
        """Sum of M/=/X lengths of mapped, non-duplicate-flagged reads (BASELINE.md §3 metric unit)."""
        ops = self.cigar & 15
        lens = (self.cigar >> 4).astype(np.int64)
        m = (ops == CMATCH) | (ops == CEQUAL) | (ops == CDIFF)
        per_op = np.where(m, lens, 0)
        cs = np.concatenate([[0], np.cumsum(per_op)])
        a = self.cigar_off.astype(np.int64); b = a + self.n_cigar.astype(np.int64)
        per_read = cs[b] - cs[a]
        ok = (self.flag & (FUNMAP | FDUP)) == 0
        return int(per_read[ok].sum())


def batch_from_c(view: CReadBatch, keep_names: bool) -> ReadBatch:
    """Copy a C-owned batch into numpy arrays."""
    n = view.n_reads

    def arr(ptr, dt, cnt):
        if cnt == 0 or not ptr:
            return np.zeros(0, dtype=dt)
        buf = (C.c_char * (cnt * np.dtype(dt).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dt, count=cnt).copy()

    kw = {}
    for k, dt in _DTYPES.items():
        if k == "cigar":
            cnt = view.n_cigar_total
        elif k == "seq4":
            cnt = view.n_base_slots // 2
        elif k == "qual":
            cnt = view.n_base_slots
        else:
            cnt = n
        kw[k] = arr(getattr(view, k), dt, cnt)
    b = ReadBatch(tid=view.tid, **kw)
    b.layout_flags = int(view.layout_flags)
    if view.layout_flags & LAYOUT_QUAL4:
        b.qual4 = arr(view.qual4, np.uint8, view.n_base_slots // 2)
        b.qual_lut = np.array(list(view.qual_lut), dtype=np.uint8)
    if view.layout_flags & LAYOUT_QUAL2:
        b.qual2 = arr(view.qual2, np.uint8, view.n_base_slots // 4)
        b.qual_lut = np.array(list(view.qual_lut), dtype=np.uint8)
    if view.layout_flags & LAYOUT_SEQ2:
        b.seq2 = arr(view.seq2, np.uint8, view.n_base_slots // 4)
        b.seq_exc_slot = arr(view.seq_exc_slot, np.uint64, view.n_seq_exc)
        b.seq_exc_code = arr(view.seq_exc_code, np.uint8, view.n_seq_exc)
    if view.layout_flags & LAYOUT_SPARSE_SA:
        b.sa_index = arr(view.sa_index, np.int32, view.n_sa)
        b.sa_sparse = {k: arr(getattr(view, "sas_" + k[3:]), _DTYPES[k], view.n_sa) for k in SA_FIELDS}
    if keep_names and view.qname_off:
        b.qname_off = arr(view.qname_off, np.uint64, n + 1)
        b.qname_pool = arr(view.qname_pool, np.uint8, int(b.qname_off[-1]) if n else 0)
    return b.normalise()


# ---- compiled loops of the two tooling helpers above (repack_canonical / compact): per read, no per-base index arrays
_JIT = None


def _jit():
    global _JIT
    if _JIT is not None:
        return _JIT
    from numba import njit, prange

    @njit(parallel=True, cache=True)
    def repack(old_off, lq, bo, seq4_old, qual_old, seq4_new, qual_new):
        for i in prange(lq.shape[0]):                      # a read's slots start on a BASE_ALIGN boundary: no two reads share a byte
            o, d = old_off[i], bo[i]
            for j in range(lq[i]):
                s_, t_ = o + j, d + j
                code = (seq4_old[s_ >> 1] >> ((~s_ & 1) << 2)) & 15
                seq4_new[t_ >> 1] |= np.uint8(code << ((~t_ & 1) << 2))
                qual_new[t_] = qual_old[s_]

    @njit(parallel=True, cache=True)
    def survey(bo, lq, seq4, qual):
        n = lq.shape[0]
        nb = 256
        hist = np.zeros((nb, 256), dtype=np.int64)
        exc = np.zeros(n, dtype=np.int64)
        for b in prange(nb):
            for i in range(b * n // nb, (b + 1) * n // nb):
                d = bo[i]
                k = 0
                for j in range(lq[i]):
                    t_ = d + j
                    hist[b, qual[t_]] += 1
                    code = (seq4[t_ >> 1] >> ((~t_ & 1) << 2)) & 15
                    if code != 1 and code != 2 and code != 4 and code != 8:
                        k += 1
                exc[i] = k
        return hist.sum(axis=0), exc

    @njit(parallel=True, cache=True)
    def pack_qual2(bo, lq, qual, inv, out):
        for i in prange(lq.shape[0]):
            d = bo[i]
            for j in range(lq[i]):
                t_ = d + j
                out[t_ >> 2] |= np.uint8(inv[qual[t_]] << ((3 - (t_ & 3)) << 1))

    @njit(parallel=True, cache=True)
    def pack_seq2(bo, lq, seq4, first, out, exc_slot, exc_code):
        for i in prange(lq.shape[0]):
            d = bo[i]
            k = first[i]
            for j in range(lq[i]):
                t_ = d + j
                code = (seq4[t_ >> 1] >> ((~t_ & 1) << 2)) & 15
                two = 255
                if code == 1:
                    two = 0
                elif code == 2:
                    two = 1
                elif code == 4:
                    two = 2
                elif code == 8:
                    two = 3
                if two == 255:
                    exc_slot[k] = t_; exc_code[k] = code; k += 1
                else:
                    out[t_ >> 2] |= np.uint8(two << ((3 - (t_ & 3)) << 1))

    _JIT = {"repack": repack, "survey": survey, "pack_qual2": pack_qual2, "pack_seq2": pack_seq2}
    return _JIT
