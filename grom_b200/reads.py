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
        ("tid", C.c_int32), ("reserved", C.c_int32),
        ("pos", C.c_void_p), ("mpos", C.c_void_p), ("tlen", C.c_void_p), ("mtid", C.c_void_p),
        ("l_qseq", C.c_void_p), ("flag", C.c_void_p), ("n_cigar", C.c_void_p), ("mapq", C.c_void_p),
        ("qname_len", C.c_void_p), ("qname_hash", C.c_void_p), ("cigar_off", C.c_void_p),
        ("base_off", C.c_void_p), ("cigar", C.c_void_p), ("seq4", C.c_void_p), ("qual", C.c_void_p),
        ("sa_pos", C.c_void_p), ("sa_start_adj", C.c_void_p), ("sa_end_adj", C.c_void_p),
        ("sa_end_adj_indel", C.c_void_p), ("sa_strand", C.c_void_p), ("sa_mapq", C.c_void_p),
        ("sa_same_chr", C.c_void_p), ("qname_off", C.c_void_p), ("qname_pool", C.c_void_p),
    ]


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
        return c

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
    if keep_names and view.qname_off:
        b.qname_off = arr(view.qname_off, np.uint64, n + 1)
        b.qname_pool = arr(view.qname_pool, np.uint8, int(b.qname_off[-1]) if n else 0)
    return b.normalise()
