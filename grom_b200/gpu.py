"""Host-side mirror of the C ABI in include/gromgpu.h (ctypes over grom_b200/libgromgpu.so).

`Chromosome` plays the role the body of `count_discordant_pairs` plays in the reference
(reference src/GROM.c:1432, call site 21057): one object per chromosome, reads pushed in BAM
order, `finish()` returns what the host needs for candidate post-processing and VCF text.
There is no CPU fallback: a missing library or device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

from .params import CNV_CALL_DTYPE, DEL_EVENT_DTYPE, SV_EVENT_DTYPE, GA, GA_COUNT, INS_CAND_DTYPE, Params, SNV_CAND_DTYPE
from .reads import CReadBatch, ReadBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GROMGPU_LIB") or os.path.join(_HERE, "libgromgpu.so")     # override: kernel tuning builds (tools/)
_LIB = None


class Stats(C.Structure):
    _fields_ = [("ms_total", C.c_float), ("ms_clear", C.c_float), ("ms_dup", C.c_float), ("ms_prep", C.c_float),
                ("ms_index", C.c_float), ("ms_pileup", C.c_float), ("ms_rdscan", C.c_float), ("ms_snvscan", C.c_float),
                ("ms_gc", C.c_float), ("ms_sv", C.c_float),
                ("launches", C.c_int32), ("reserved", C.c_int32), ("n_reads", C.c_int64), ("n_applied", C.c_int64),
                ("n_dups", C.c_int64), ("aligned_bases", C.c_int64), ("bytes_reads", C.c_int64), ("n_sv_items", C.c_int64),
                ("n_other_slabs", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("reserved")}


class CResult(C.Structure):
    _fields_ = [("scan_first", C.c_int32), ("scan_last", C.c_int32), ("n_snv", C.c_int64), ("snv", C.c_void_p),
                ("snv_ave_rd", C.c_double), ("n_ins", C.c_int64), ("ins", C.c_void_p), ("n_del", C.c_int64), ("del_ev", C.c_void_p), ("n_sv", C.c_int64), ("sv_ev", C.c_void_p)]


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.gromgpu_last_error.restype = C.c_char_p
        L.gromgpu_init.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Params)]
        L.gromgpu_set_stream.argtypes = [C.c_void_p]
        L.gromgpu_chr_begin.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int64]
        L.gromgpu_push_reads.argtypes = [C.c_void_p, C.POINTER(CReadBatch)]
        L.gromgpu_chr_run.argtypes = [C.c_void_p]
        L.gromgpu_chr_reset.argtypes = [C.c_void_p, C.c_void_p]
        L.gromgpu_chr_rebind.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
        L.gromgpu_chr_result.argtypes = [C.c_void_p, C.POINTER(CResult)]
        L.gromgpu_chr_finish.argtypes = [C.c_void_p, C.POINTER(CResult)]
        L.gromgpu_chr_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.gromgpu_debug_fetch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64]
        L.gromgpu_debug_fetch_cluster.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64]
        L.gromgpu_fetch_read_state.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        L.gromgpu_chr_cnv.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.gromgpu_cnv_fetch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64]
        L.gromgpu_chr_free.argtypes = [C.c_void_p]
        L.gromgpu_chr_free.restype = None
        L.gromgpu_chr_sync.argtypes = [C.c_void_p]
        L.gromgpu_stream_create.argtypes = [C.POINTER(C.c_void_p)]
        L.gromgpu_stream_destroy.argtypes = [C.c_void_p]
        L.gromgpu_stream_destroy.restype = None
        L.gromgpu_chr_begin_on.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        L.gromgpu_chr_bytes_estimate.argtypes = [C.c_int64, C.c_int64, C.c_int64]
        L.gromgpu_chr_bytes_estimate.restype = C.c_int64
        L.gromgpu_device_free_bytes.restype = C.c_int64
        L.gromgpu_shutdown.restype = None
        _LIB = L
    return _LIB


_PARAMS: Optional[Params] = None
_P2S = None                 # the caller's p-value -> sd table (gromhost_pval2sd), built once


class GromGpuError(RuntimeError):
    pass


def _ck(rc: int):
    if rc != 0:
        raise GromGpuError(lib().gromgpu_last_error().decode())


def init(device: int, hez: np.ndarray, mq: np.ndarray, params: Params):
    hez = np.ascontiguousarray(hez, dtype=np.float64); mq = np.ascontiguousarray(mq, dtype=np.float64)
    assert hez.shape == (1001, 1001) and mq.shape == (1001, 1001)
    _ck(lib().gromgpu_init(device, hez.ctypes.data, mq.ctypes.data, C.byref(params)))
    global _PARAMS
    _PARAMS = params


def set_stream(cuda_stream_ptr: Optional[int]):
    _ck(lib().gromgpu_set_stream(C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))


def shutdown():
    lib().gromgpu_shutdown()


class CCnvResult(C.Structure):
    _fields_ = [("n_calls", C.c_int64), ("calls", C.c_void_p), ("chr_ave", C.c_double), ("chr_sd", C.c_double), ("blk_ave", C.c_double),
                ("biased_repeat", C.c_int32), ("n_sample_blocks", C.c_int32), ("n_repeats", C.c_int64), ("n_samples", C.c_int64),
                ("n_frames", C.c_int64), ("win_sd", C.c_void_p), ("win_cnt", C.c_void_p), ("bin_ave", C.c_void_p), ("bin_sd", C.c_void_p),
                ("bin_del_thr", C.c_void_p), ("bin_dup_thr", C.c_void_p), ("bin_n", C.c_void_p),
                ("ms_device", C.c_float), ("ms_host", C.c_float), ("ms_total", C.c_float), ("launches", C.c_int32), ("reserved", C.c_int32),
                ("d2h_bytes", C.c_int64)]


@dataclass
class CnvResult:
    calls: np.ndarray           # CNV_CALL_DTYPE, deletions then duplications
    chr_ave: float
    chr_sd: float
    blk_ave: float
    biased_repeat: int
    n_sample_blocks: int
    n_repeats: int
    n_samples: int
    n_frames: int
    win_sd: np.ndarray
    win_cnt: np.ndarray
    ave: np.ndarray
    sd: np.ndarray
    del_thr: np.ndarray
    dup_thr: np.ndarray
    n: np.ndarray
    ms_device: float
    ms_host: float
    ms_total: float
    launches: int = 0
    d2h_bytes: int = 0


@dataclass
class ChrResult:
    scan_first: int
    scan_last: int
    snv: np.ndarray           # SNV_CAND_DTYPE, ascending position
    snv_ave_rd: float
    ins: np.ndarray = None    # INS_CAND_DTYPE small-insertion candidates, ascending position
    del_ev: np.ndarray = None  # DEL_EVENT_DTYPE small-deletion scan events (position, start before end)
    sv_ev: np.ndarray = None   # SV_EVENT_DTYPE structural-variant gate events, scan order


def stream_create() -> int:
    s = C.c_void_p()
    _ck(lib().gromgpu_stream_create(C.byref(s)))
    return s.value


def stream_destroy(s: Optional[int]):
    if s:
        lib().gromgpu_stream_destroy(C.c_void_p(s))


def chr_bytes_estimate(length: int, n_reads: int, n_base_slots: int) -> int:
    return int(lib().gromgpu_chr_bytes_estimate(length, n_reads, n_base_slots))


def device_free_bytes() -> int:
    return int(lib().gromgpu_device_free_bytes())


class Chromosome:
    def __init__(self, tid: int, fasta: np.ndarray, stream: Optional[int] = None):
        fa = np.ascontiguousarray(fasta, dtype=np.uint8)
        self.length = int(fa.shape[0])
        self._h = C.c_void_p()
        if stream:
            _ck(lib().gromgpu_chr_begin_on(C.byref(self._h), tid, fa.ctypes.data, self.length, C.c_void_p(stream)))
        else:
            _ck(lib().gromgpu_chr_begin(C.byref(self._h), tid, fa.ctypes.data, self.length))

    def push_reads(self, batch: ReadBatch):
        cb = batch.as_c()
        _ck(lib().gromgpu_push_reads(self._h, C.byref(cb)))

    def push_reads_c(self, cb: CReadBatch):
        _ck(lib().gromgpu_push_reads(self._h, C.byref(cb)))

    def reset(self, fasta: Optional[np.ndarray] = None):
        """Drop the pushed reads, keep the device buffers; optionally upload new FASTA characters (same length)."""
        ptr = None
        if fasta is not None:
            assert fasta.dtype == np.uint8 and fasta.shape[0] == self.length and fasta.flags.c_contiguous
            ptr = fasta.ctypes.data
        _ck(lib().gromgpu_chr_reset(self._h, ptr))

    def rebind(self, tid: int, fasta: np.ndarray) -> bool:
        """Reuse the handle (all device buffers kept) for another chromosome no longer than the one it was created for
        (gromgpu_chr_rebind).  False = it does not fit; close() and create a new Chromosome."""
        fa = np.ascontiguousarray(fasta, dtype=np.uint8)
        rc = lib().gromgpu_chr_rebind(self._h, tid, fa.ctypes.data, int(fa.shape[0]))
        if rc == 1:
            return False
        _ck(rc)
        self.length = int(fa.shape[0])
        return True

    def sync(self):
        _ck(lib().gromgpu_chr_sync(self._h))

    def run(self):
        _ck(lib().gromgpu_chr_run(self._h))

    def result(self) -> ChrResult:
        r = CResult()
        _ck(lib().gromgpu_chr_result(self._h, C.byref(r)))
        if r.n_snv:
            buf = (C.c_char * (r.n_snv * SNV_CAND_DTYPE.itemsize)).from_address(r.snv)
            snv = np.frombuffer(buf, dtype=SNV_CAND_DTYPE, count=r.n_snv).copy()
        else:
            snv = np.zeros(0, dtype=SNV_CAND_DTYPE)
        if r.n_ins:
            buf = (C.c_char * (r.n_ins * INS_CAND_DTYPE.itemsize)).from_address(r.ins)
            ins = np.frombuffer(buf, dtype=INS_CAND_DTYPE, count=r.n_ins).copy()
        else:
            ins = np.zeros(0, dtype=INS_CAND_DTYPE)
        if r.n_del:
            buf = (C.c_char * (r.n_del * DEL_EVENT_DTYPE.itemsize)).from_address(r.del_ev)
            dev = np.frombuffer(buf, dtype=DEL_EVENT_DTYPE, count=r.n_del).copy()
        else:
            dev = np.zeros(0, dtype=DEL_EVENT_DTYPE)
        if r.n_sv:
            buf = (C.c_char * (r.n_sv * SV_EVENT_DTYPE.itemsize)).from_address(r.sv_ev)
            sve = np.frombuffer(buf, dtype=SV_EVENT_DTYPE, count=r.n_sv).copy()
        else:
            sve = np.zeros(0, dtype=SV_EVENT_DTYPE)
        return ChrResult(r.scan_first, r.scan_last, snv, r.snv_ave_rd, ins, dev, sve)

    def finish(self) -> ChrResult:
        self.run()
        return self.result()

    def stats(self) -> Stats:
        s = Stats()
        _ck(lib().gromgpu_chr_stats(self._h, C.byref(s)))
        return s

    def fetch(self, name: str, p0: int = 0, p1: Optional[int] = None) -> np.ndarray:
        p1 = self.length if p1 is None else p1
        out = np.empty(p1 - p0, dtype=np.int32)
        _ck(lib().gromgpu_debug_fetch(self._h, GA[name], out.ctypes.data, p0, p1))
        return out

    def fetch_all(self) -> np.ndarray:
        out = np.empty((GA_COUNT, self.length), dtype=np.int32)
        for k in range(GA_COUNT):
            _ck(lib().gromgpu_debug_fetch(self._h, k, out[k].ctypes.data, 0, self.length))
        return out

    def fetch_clusters(self):
        """(w, rs, re, dist, mchr, other_len): [10,P] int32 x3, [10,P] float64, [2,P] int32, [P] int32."""
        P = self.length
        w = np.empty((10, P), np.int32); rs = np.empty((10, P), np.int32); re = np.empty((10, P), np.int32)
        dist = np.empty((10, P), np.float64); mchr = np.empty((2, P), np.int32); ol = np.empty(P, np.int32)
        for k in range(10):
            for what, arr in ((0, w), (1, rs), (2, re), (3, dist)):
                _ck(lib().gromgpu_debug_fetch_cluster(self._h, what, k, arr[k].ctypes.data, 0, P))
        for k in range(2):
            _ck(lib().gromgpu_debug_fetch_cluster(self._h, 4, k, mchr[k].ctypes.data, 0, P))
        _ck(lib().gromgpu_debug_fetch_cluster(self._h, 5, 0, ol.ctypes.data, 0, P))
        return w, rs, re, dist, mchr, ol

    def cnv(self, ploidy: Optional[int] = None, params=None) -> "CnvResult":
        """Read-depth CNV path (gromgpu_chr_cnv) on the depth arrays of the last run()."""
        global _P2S
        if _P2S is None:
            from . import hostlib
            _P2S = hostlib.pval2sd()
        pv, sd = _P2S
        r = CCnvResult()
        params = params if params is not None else _PARAMS
        pl = ploidy if ploidy is not None else params.ploidy
        _ck(lib().gromgpu_chr_cnv(self._h, pv.ctypes.data, sd.ctypes.data, len(pv), pl, C.byref(r)))
        if r.n_calls:
            buf = (C.c_char * (r.n_calls * CNV_CALL_DTYPE.itemsize)).from_address(r.calls)
            calls = np.frombuffer(buf, dtype=CNV_CALL_DTYPE, count=r.n_calls).copy()
        else:
            calls = np.zeros(0, dtype=CNV_CALL_DTYPE)
        nw = int(params.max_rd_window_len) + 1

        def arr(ptr, n, dt):
            return np.frombuffer((C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr), dtype=dt, count=n).copy()
        return CnvResult(calls=calls, chr_ave=r.chr_ave, chr_sd=r.chr_sd, blk_ave=r.blk_ave, biased_repeat=r.biased_repeat,
                         n_sample_blocks=r.n_sample_blocks, n_repeats=r.n_repeats, n_samples=r.n_samples, n_frames=r.n_frames,
                         win_sd=arr(r.win_sd, nw, np.float64), win_cnt=arr(r.win_cnt, nw, np.int64),
                         ave=arr(r.bin_ave, 202, np.float64).reshape(2, 101), sd=arr(r.bin_sd, 202, np.float64).reshape(2, 101),
                         del_thr=arr(r.bin_del_thr, 202, np.float64).reshape(2, 101), dup_thr=arr(r.bin_dup_thr, 202, np.float64).reshape(2, 101),
                         n=arr(r.bin_n, 202, np.int64).reshape(2, 101), ms_device=r.ms_device, ms_host=r.ms_host, ms_total=r.ms_total, launches=r.launches, d2h_bytes=r.d2h_bytes)

    def cnv_fetch(self, what: str, p0: int = 0, p1: Optional[int] = None) -> np.ndarray:
        sel, dt = {"z": (0, np.float64), "mask": (1, np.uint8), "mq_mean": (2, np.uint8), "depth": (3, np.int32)}[what]
        p1 = self.length if p1 is None else p1
        out = np.empty(p1 - p0, dtype=dt)
        _ck(lib().gromgpu_cnv_fetch(self._h, sel, out.ctypes.data, p0, p1))
        return out

    def read_state(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint8)
        _ck(lib().gromgpu_fetch_read_state(self._h, out.ctypes.data, 0, n))
        return out

    def close(self):
        if self._h:
            lib().gromgpu_chr_free(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
