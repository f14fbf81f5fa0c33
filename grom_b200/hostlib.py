"""ctypes binding of libgromhost.so (include/gromhost.h): BAM -> packed batch, batch -> BAM+BAI."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence

import numpy as np

from .reads import CReadBatch, ReadBatch, batch_from_c

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libgromhost.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(path)
        L.gromhost_last_error.restype = C.c_char_p
        L.gromhost_bam_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.gromhost_bam_close.argtypes = [C.c_void_p]
        L.gromhost_bam_n_targets.argtypes = [C.c_void_p]
        L.gromhost_bam_target_name.argtypes = [C.c_void_p, C.c_int]
        L.gromhost_bam_target_name.restype = C.c_char_p
        L.gromhost_bam_target_len.argtypes = [C.c_void_p, C.c_int]
        L.gromhost_bam_target_len.restype = C.c_int64
        L.gromhost_bam_has_index.argtypes = [C.c_void_p]
        L.gromhost_bam_read_target.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.gromhost_bam_iter_open.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.gromhost_bam_iter_next.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]
        L.gromhost_bam_iter_close.argtypes = [C.c_void_p]
        L.gromhost_batch_view.argtypes = [C.c_void_p, C.POINTER(CReadBatch)]
        L.gromhost_batch_free.argtypes = [C.c_void_p]
        L.gromhost_fasta_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.gromhost_fasta_close.argtypes = [C.c_void_p]
        L.gromhost_fasta_n.argtypes = [C.c_void_p]
        L.gromhost_fasta_name.argtypes = [C.c_void_p, C.c_int]
        L.gromhost_fasta_name.restype = C.c_char_p
        L.gromhost_fasta_find.argtypes = [C.c_void_p, C.c_char_p]
        L.gromhost_fasta_raw_bytes.argtypes = [C.c_void_p, C.c_int]
        L.gromhost_fasta_raw_bytes.restype = C.c_int64
        L.gromhost_fasta_load.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
        L.gromhost_fasta_load.restype = C.c_int64
        L.gromhost_bam_write.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.c_int,
                                         C.POINTER(CReadBatch), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int]
        _LIB = L
    return _LIB


def _check(rc: int):
    if rc != 0:
        raise RuntimeError(lib().gromhost_last_error().decode())


class OwnedBatch:
    """A batch that stays in the batcher's own memory (no numpy copy): what the C host program hands to gromgpu_push_reads.
    `as_c()` is the view the CUDA library takes; `free()` (or leaving the `with` block) returns the memory."""

    def __init__(self, handle: C.c_void_p):
        self._h = handle
        self.view = CReadBatch()
        lib().gromhost_batch_view(self._h, C.byref(self.view))
        self.tid = int(self.view.tid)
        self.n_reads = int(self.view.n_reads)
        self.n_base_slots = int(self.view.n_base_slots)
        self.n_cigar_total = int(self.view.n_cigar_total)
        self.layout_flags = int(self.view.layout_flags)

    def as_c(self) -> CReadBatch:
        if not self._h:
            raise RuntimeError("batch already freed")
        return self.view

    def to_numpy(self, keep_names: bool = False) -> ReadBatch:
        return batch_from_c(self.as_c(), keep_names)

    def free(self):
        if self._h:
            lib().gromhost_batch_free(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.free()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Bam:
    """An open BAM; `read_target(tid)` returns the contig's reads as a ReadBatch (numpy copy), `read_target_owned(tid)` as the
    batcher's own memory (OwnedBatch, what the product path pushes to the GPU)."""

    def __init__(self, path: str):
        self._h = C.c_void_p()
        _check(lib().gromhost_bam_open(path.encode(), C.byref(self._h)))
        n = lib().gromhost_bam_n_targets(self._h)
        self.names = [lib().gromhost_bam_target_name(self._h, i).decode() for i in range(n)]
        self.lens = [int(lib().gromhost_bam_target_len(self._h, i)) for i in range(n)]
        self.has_index = bool(lib().gromhost_bam_has_index(self._h))
        # records per target from the index's metadata pseudo-bin (None when the index has none): load measure for contig -> GPU assignment
        lib().gromhost_bam_target_reads.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        counts = []
        for i in range(n):
            m, u = C.c_int64(), C.c_int64()
            counts.append(m.value + u.value if lib().gromhost_bam_target_reads(self._h, i, C.byref(m), C.byref(u)) == 0 else None)
        self.read_counts = counts if n and all(c is not None for c in counts) else None

    def read_target(self, tid: int, keep_names: bool = False, threads: int = 0) -> ReadBatch:
        bt = C.c_void_p()
        _check(lib().gromhost_bam_read_target(self._h, tid, int(keep_names), threads, C.byref(bt)))
        try:
            v = CReadBatch()
            lib().gromhost_batch_view(bt, C.byref(v))
            return batch_from_c(v, keep_names)
        finally:
            lib().gromhost_batch_free(bt)

    def read_target_owned(self, tid: int, keep_names: bool = False, threads: int = 0) -> OwnedBatch:
        bt = C.c_void_p()
        _check(lib().gromhost_bam_read_target(self._h, tid, int(keep_names), threads, C.byref(bt)))
        return OwnedBatch(bt)

    def iter_target(self, tid: int, max_reads: int, keep_names: bool = False, threads: int = 0, owned: bool = False):
        """The reads of a target in consecutive pieces of at least `max_reads` records (gromhost_bam_iter_*): host memory is bounded by the
        piece, and the pieces are what consecutive gromgpu_push_reads calls take.  Yields ReadBatch copies, or OwnedBatch (free it) with
        owned=True.  A trailing piece may be empty."""
        it = C.c_void_p()
        _check(lib().gromhost_bam_iter_open(self._h, tid, int(keep_names), threads, C.byref(it)))
        try:
            while True:
                bt = C.c_void_p()
                rc = lib().gromhost_bam_iter_next(it, int(max_reads), C.byref(bt))
                if rc == 1:
                    return
                _check(rc)
                ob = OwnedBatch(bt)
                if owned:
                    yield ob
                else:
                    try:
                        yield ob.to_numpy(keep_names)
                    finally:
                        ob.free()
        finally:
            lib().gromhost_bam_iter_close(it)

    def library_stats(self, rd_min_mapq: int = 20, threads: int = 0) -> dict:
        """find_insert_mean (reference src/GROM.c:1205-1318) straight over the file (gromhost_bam_library_stats)."""
        L = lib()
        L.gromhost_bam_library_stats.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5
        v = [C.c_int(), C.c_int(), C.c_int(), C.c_int()]
        m = C.c_int64()
        _check(L.gromhost_bam_library_stats(self._h, rd_min_mapq, threads, C.byref(v[0]), C.byref(v[1]), C.byref(v[2]), C.byref(v[3]), C.byref(m)))
        return dict(insert_mean=v[0].value, lseq=v[1].value, insert_min=v[2].value, insert_max=v[3].value, mapped_reads=m.value)

    def close(self):
        if self._h:
            lib().gromhost_bam_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class Fasta:
    """An indexed reference FASTA (gromhost_fasta_*): `names` are the reference's contig names (first word of the header, lower-cased,
    at most 49 characters); `load(k)` returns the characters of contig k, case preserved, one contig in memory at a time."""

    def __init__(self, path: str):
        self._h = C.c_void_p()
        _check(lib().gromhost_fasta_open(path.encode(), C.byref(self._h)))
        n = lib().gromhost_fasta_n(self._h)
        self.names = [lib().gromhost_fasta_name(self._h, k).decode("latin-1") for k in range(n)]

    def find(self, name: str) -> int:
        return int(lib().gromhost_fasta_find(self._h, name.encode("latin-1")))

    def load(self, k: int) -> np.ndarray:
        cap = int(lib().gromhost_fasta_raw_bytes(self._h, k))
        if cap < 0:
            raise IndexError(k)
        buf = np.empty(max(cap, 1), dtype=np.uint8)
        n = int(lib().gromhost_fasta_load(self._h, k, buf.ctypes.data, cap))
        if n < 0:
            raise RuntimeError(lib().gromhost_last_error().decode())
        return buf[:n] if n == cap else buf[:n].copy()

    def close(self):
        if self._h:
            lib().gromhost_fasta_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def write_bam(path: str, names: Sequence[str], lens: Sequence[int], batches: List[ReadBatch], level: int = 1):
    """Serialise batches (sorted by tid, each with qname_off/qname_pool) as BAM + .bai."""
    batches = sorted(batches, key=lambda b: b.tid)
    n = len(names)
    c_names = (C.c_char_p * n)(*[s.encode() for s in names])
    c_lens = (C.c_int64 * n)(*[int(x) for x in lens])
    cb = (CReadBatch * max(1, len(batches)))()
    aux_off = (C.c_void_p * max(1, len(batches)))()
    aux_pool = (C.c_void_p * max(1, len(batches)))()
    any_aux = False
    for i, b in enumerate(batches):
        cb[i] = b.as_c()
        if b.aux_off is not None:
            any_aux = True
            aux_off[i] = b.aux_off.ctypes.data
            aux_pool[i] = b.aux_pool.ctypes.data if b.aux_pool.size else None
        else:
            aux_off[i] = None
            aux_pool[i] = None
    _check(lib().gromhost_bam_write(path.encode(), n, c_names, c_lens, len(batches), cb,
                                    aux_off if any_aux else None, aux_pool if any_aux else None, level))


def _table_protos():
    L = lib()
    L.gromhost_tables_get.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.gromhost_tables_compute.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    L.gromhost_pval2sd.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    return L


def tables(table_dir=None, min_mapq: int = 20, write_missing: bool = False):
    """(hez, mq) binomial tables, float64 [1001, 1001]; loaded from `table_dir` when the reference's
    text files are there, else computed (reference src/GROM.c:21134-21586)."""
    L = _table_protos()
    hez = np.zeros((1001, 1001)); mq = np.zeros((1001, 1001))
    L.gromhost_tables_get(table_dir.encode() if table_dir else None, min_mapq, int(write_missing),
                          hez.ctypes.data, mq.ctypes.data)
    return hez, mq


def pval2sd():
    L = _table_protos()
    pv = np.zeros(1001); sd = np.zeros(1001)
    n = L.gromhost_pval2sd(pv.ctypes.data, sd.ctypes.data, 1001)
    assert n == 1001
    return pv, sd


def _vcf(fn_name, params, chr_name, fasta, *args, room: int = 0):
    from .params import Params
    L = lib()
    fn = getattr(L, fn_name)
    fn.restype = C.c_int64
    fa = np.ascontiguousarray(fasta, dtype=np.uint8)
    cap = max(1 << 16, int(room))
    while True:
        buf = C.create_string_buffer(cap)
        n = fn(C.byref(params), chr_name.encode(), fa.ctypes.data_as(C.c_char_p), *args, buf, C.c_int64(cap))
        if n >= 0:
            return buf.raw[:n].decode()
        if n != -1 or cap > (1 << 34):                 # -1 = buffer too small; anything else is an error of the writer
            raise RuntimeError(f"{fn_name} failed with code {n}")
        cap *= 4


def vcf_header(fasta_name: str, is_ctx: bool = False) -> str:
    """The reference's header block of <out> (or of <out>.ctx.vcf): gromhost_vcf_header."""
    L = lib()
    L.gromhost_vcf_header.restype = C.c_int64
    L.gromhost_vcf_header.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int64]
    buf = C.create_string_buffer(1 << 15)
    n = L.gromhost_vcf_header(fasta_name.encode(), int(is_ctx), buf, len(buf))
    if n < 0:
        raise RuntimeError("gromhost_vcf_header: buffer too small")
    return buf.raw[:n].decode()


def vcf_snv(params, chr_name: str, fasta: np.ndarray, snv: np.ndarray, ave_rd: float) -> str:
    """SNV records (reference src/GROM.c:15046-15095) from the candidates gromgpu_chr_result returns."""
    a = np.ascontiguousarray(snv)
    return _vcf("gromhost_vcf_snv", params, chr_name, fasta, C.c_void_p(a.ctypes.data), C.c_int64(len(a)), C.c_double(ave_rd))


def vcf_ins(params, chr_name: str, fasta: np.ndarray, ins: np.ndarray) -> str:
    """Small-insertion records (reference src/GROM.c:16253-16340)."""
    a = np.ascontiguousarray(ins)
    return _vcf("gromhost_vcf_ins", params, chr_name, fasta, C.c_int64(len(fasta)), C.c_void_p(a.ctypes.data), C.c_int64(len(a)))


def vcf_smalldel(params, chr_name: str, fasta: np.ndarray, events: np.ndarray) -> str:
    """Small-deletion records: pairing state machine + filter + text (reference src/GROM.c:11475-11745, 16351-16490)."""
    a = np.ascontiguousarray(events)
    return _vcf("gromhost_vcf_smalldel", params, chr_name, fasta, C.c_int64(len(fasta)), C.c_void_p(a.ctypes.data), C.c_int64(len(a)))


def vcf_cnv(params, chr_name: str, calls: np.ndarray) -> str:
    """Read-depth CNV records (-V filter + text, reference src/GROM.c:17197-17500) from gromgpu_chr_cnv's calls."""
    a = np.ascontiguousarray(calls)
    return _vcf("gromhost_vcf_cnv", params, chr_name, np.zeros(1, dtype=np.uint8), C.c_int64(0), C.c_void_p(a.ctypes.data), C.c_int64(len(a)))


class CSvLists(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("n_dup", "n_del", "n_inv_f", "n_inv_r", "n_ins", "n_ctx_f", "n_ctx_r")] + \
               [(n, C.c_void_p) for n in ("dup", "del_", "inv_f", "inv_r", "ins", "ctx_f", "ctx_r")]


def sv_lists(params, events: np.ndarray) -> dict:
    """Candidate lists of the structural-variant scan (reference state at src/GROM.c:15164) from the gate events:
    {"dup","del","inv_f","inv_r","ins": SV_PAIR_DTYPE arrays, "ctx_f","ctx_r": SV_EVENT_DTYPE arrays}."""
    from .params import SV_EVENT_DTYPE, SV_PAIR_DTYPE
    L = lib()
    L.gromhost_sv_lists.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(CSvLists)]
    L.gromhost_sv_lists_free.argtypes = [C.POINTER(CSvLists)]
    L.gromhost_sv_lists_free.restype = None
    ev = np.ascontiguousarray(events, dtype=SV_EVENT_DTYPE)
    out = CSvLists()
    if L.gromhost_sv_lists(C.byref(params), ev.ctypes.data, len(ev), C.byref(out)) != 0:
        raise RuntimeError("gromhost_sv_lists failed")

    def take(ptr, n, dt):
        if not n:
            return np.zeros(0, dtype=dt)
        return np.frombuffer((C.c_char * (n * dt.itemsize)).from_address(ptr), dtype=dt, count=n).copy()
    res = {"dup": take(out.dup, out.n_dup, SV_PAIR_DTYPE), "del": take(out.del_, out.n_del, SV_PAIR_DTYPE),
           "inv_f": take(out.inv_f, out.n_inv_f, SV_PAIR_DTYPE), "inv_r": take(out.inv_r, out.n_inv_r, SV_PAIR_DTYPE),
           "ins": take(out.ins, out.n_ins, SV_PAIR_DTYPE), "ctx_f": take(out.ctx_f, out.n_ctx_f, SV_EVENT_DTYPE),
           "ctx_r": take(out.ctx_r, out.n_ctx_r, SV_EVENT_DTYPE)}
    L.gromhost_sv_lists_free(C.byref(out))
    return res


def vcf_contig(params, chr_name: str, fasta: np.ndarray, snv: np.ndarray, snv_ave_rd: float, ins: np.ndarray, del_ev: np.ndarray,
               sv_ev: np.ndarray, cnv_calls: np.ndarray) -> str:
    """Every record of one contig in the reference's output order (gromhost_vcf_contig)."""
    from .params import CNV_CALL_DTYPE, DEL_EVENT_DTYPE, INS_CAND_DTYPE, SNV_CAND_DTYPE, SV_EVENT_DTYPE
    a = [np.ascontiguousarray(x, dtype=dt) for x, dt in ((snv, SNV_CAND_DTYPE), (ins, INS_CAND_DTYPE), (del_ev, DEL_EVENT_DTYPE),
                                                          (sv_ev, SV_EVENT_DTYPE), (cnv_calls, CNV_CALL_DTYPE))]
    return _vcf("gromhost_vcf_contig", params, chr_name, fasta, C.c_int64(len(fasta)), C.c_void_p(a[0].ctypes.data), C.c_int64(len(a[0])),
                C.c_double(snv_ave_rd), C.c_void_p(a[1].ctypes.data), C.c_int64(len(a[1])), C.c_void_p(a[2].ctypes.data), C.c_int64(len(a[2])),
                C.c_void_p(a[3].ctypes.data), C.c_int64(len(a[3])), C.c_void_p(a[4].ctypes.data), C.c_int64(len(a[4])),
                room=224 * len(a[0]) + 320 * (len(a[1]) + len(a[2]) + len(a[3]) + len(a[4])) + (1 << 16))     # every candidate as a record: no refused first call


def library_stats(batches, rd_min_mapq: int = 20) -> dict:
    """find_insert_mean (reference src/GROM.c:1205-1318) over per-contig batches in contig order."""
    L = lib()
    L.gromhost_libstats_new.restype = C.c_void_p
    L.gromhost_libstats_add.argtypes = [C.c_void_p, C.c_void_p]
    L.gromhost_libstats_finish.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    L.gromhost_libstats_free.argtypes = [C.c_void_p]
    L.gromhost_libstats_free.restype = None
    h = L.gromhost_libstats_new(rd_min_mapq)
    try:
        for b in batches:
            cb = b.as_c()
            if L.gromhost_libstats_add(h, C.byref(cb)):
                break
        v = [C.c_int(), C.c_int(), C.c_int(), C.c_int()]
        m = C.c_int64()
        if L.gromhost_libstats_finish(h, C.byref(v[0]), C.byref(v[1]), C.byref(v[2]), C.byref(v[3]), C.byref(m)) != 0:
            raise RuntimeError("no reads to estimate the insert size from")
        return dict(insert_mean=v[0].value, lseq=v[1].value, insert_min=v[2].value, insert_max=v[3].value, mapped_reads=m.value)
    finally:
        L.gromhost_libstats_free(h)


def ctx_contig(params, tid: int, sv_ev: np.ndarray) -> np.ndarray:
    """Translocation records of one contig (candidate merge + emission filter, reference src/GROM.c:16098-16246) from its gate events."""
    from .params import CTX_RECORD_DTYPE, SV_EVENT_DTYPE
    L = lib()
    L.gromhost_ctx_contig.restype = C.c_int64
    L.gromhost_ctx_contig.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    lists = sv_lists(params, sv_ev)
    f, r = np.ascontiguousarray(lists["ctx_f"], dtype=SV_EVENT_DTYPE), np.ascontiguousarray(lists["ctx_r"], dtype=SV_EVENT_DTYPE)
    out = np.zeros(len(f) + len(r) + 1, dtype=CTX_RECORD_DTYPE)
    n = L.gromhost_ctx_contig(C.byref(params), tid, f.ctypes.data, len(f), r.ctypes.data, len(r), out.ctypes.data, len(out))
    if n < 0:
        raise RuntimeError("gromhost_ctx_contig failed")
    return out[:n].copy()


def ctx_vcf(params, target_names, records: np.ndarray) -> str:
    """Mate pairing across contigs and the .ctx.vcf records (reference src/GROM.c:22470-22745); records = concatenation of
    ctx_contig() results in contig processing order."""
    from .params import CTX_RECORD_DTYPE
    L = lib()
    L.gromhost_ctx_vcf.restype = C.c_int64
    names = [n.lower().encode() for n in target_names]
    arr = (C.c_char_p * len(names))(*names)
    rec = np.ascontiguousarray(records, dtype=CTX_RECORD_DTYPE).copy()
    cap = 1 << 16
    while True:
        buf = C.create_string_buffer(cap)
        n = L.gromhost_ctx_vcf(C.byref(params), arr, len(names), C.c_void_p(rec.ctypes.data), C.c_int64(len(rec)), buf, C.c_int64(cap))
        if n >= 0:
            return buf.raw[:n].decode()
        if n != -1 or cap > (1 << 34):
            raise RuntimeError(f"gromhost_ctx_vcf failed with code {n}")
        cap *= 4
        rec = np.ascontiguousarray(records, dtype=CTX_RECORD_DTYPE).copy()
