"""TEST INFRASTRUCTURE ONLY: ctypes access to oracle/_ref/liboracle.so and runners for the
reference binaries in oracle/_ref (white-box GROM_ref with dump hooks, prebuilt GROM_dist)."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import sys
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from grom_b200.params import SV_EVENT_DTYPE, DEL_EVENT_DTYPE, GA, GA_COUNT, GA_NAMES, INS_CAND_DTYPE, Params, SNV_CAND_DTYPE  # noqa: E402
from grom_b200.reads import CReadBatch, ReadBatch  # noqa: E402

REF_DIR = os.path.join(_HERE, "_ref")
GH_NI, GH_ND = 86, 10
SCAN_DTYPE = np.dtype([("pos", np.int32), ("v", np.int32, (GH_NI,)), ("d", np.float64, (GH_ND,))])
READS_DTYPE = np.dtype([("pos", np.int32), ("mpos", np.int32), ("tlen", np.int32), ("flag", np.int32),
                        ("mapq", np.int32), ("keep", np.int32)])


class COut(C.Structure):
    _fields_ = [("chr_len", C.c_int64), ("arrays", C.c_void_p), ("read_state", C.c_void_p),
                ("scan_first", C.c_int32), ("scan_last", C.c_int32), ("lookahead_lseq", C.c_void_p),
                ("snv", C.c_void_p), ("snv_cap", C.c_int64), ("n_snv", C.c_int64), ("snv_ave_rd", C.c_double),
                ("cl_w", C.c_void_p), ("cl_rs", C.c_void_p), ("cl_re", C.c_void_p), ("cl_dist", C.c_void_p),
                ("cl_mchr", C.c_void_p), ("other_len", C.c_void_p), ("ins", C.c_void_p), ("ins_cap", C.c_int64), ("n_ins", C.c_int64),
                ("del_ev", C.c_void_p), ("del_cap", C.c_int64), ("n_del", C.c_int64),
                ("sv_ev", C.c_void_p), ("sv_cap", C.c_int64), ("n_sv", C.c_int64)]


_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(REF_DIR, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle`")
        L = C.CDLL(path)
        L.oracle_run_chr.argtypes = [C.POINTER(Params), C.POINTER(CReadBatch), C.c_char_p, C.c_int64,
                                     C.c_void_p, C.c_void_p, C.POINTER(COut)]
        L.oracle_format_snv_vcf.argtypes = [C.POINTER(Params), C.c_char_p, C.c_char_p, C.c_void_p, C.c_int64,
                                            C.c_double, C.c_char_p, C.c_int64]
        L.oracle_format_snv_vcf.restype = C.c_int64
        L.oracle_gc_prepass.argtypes = [C.POINTER(Params), C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.oracle_gc_prepass.restype = None
        _LIB = L
    return _LIB


@dataclass
class OracleResult:
    arrays: np.ndarray          # [GA_COUNT, P] int32
    read_state: np.ndarray      # [n_reads] uint8
    scan_first: int
    scan_last: int
    lookahead_lseq: np.ndarray
    snv: np.ndarray             # SNV_CAND_DTYPE
    snv_ave_rd: float
    cl_w: np.ndarray = None     # [10, P] cluster weights (del_f del_r dup_f dup_r inv_f1 inv_r1 inv_f2 inv_r2 ctx_f ctx_r)
    cl_rs: np.ndarray = None
    cl_re: np.ndarray = None
    cl_dist: np.ndarray = None  # [10, P] float64
    cl_mchr: np.ndarray = None  # [2, P]
    other_len: np.ndarray = None
    ins: np.ndarray = None      # INS_CAND_DTYPE small-insertion candidates
    del_ev: np.ndarray = None   # DEL_EVENT_DTYPE small-deletion scan events
    sv_ev: np.ndarray = None    # SV_EVENT_DTYPE structural-variant gate events, scan order

    def __getitem__(self, name: str) -> np.ndarray:
        return self.arrays[GA[name]]


def run_chr(params: Params, batch: ReadBatch, fasta: np.ndarray, hez: np.ndarray, mq: np.ndarray,
            snv_cap: int = 1 << 20) -> OracleResult:
    P = int(fasta.shape[0])
    arrays = np.zeros((GA_COUNT, P), dtype=np.int32)
    state = np.zeros(max(1, batch.n_reads), dtype=np.uint8)
    look = np.zeros(P, dtype=np.int32)
    snv = np.zeros(snv_cap, dtype=SNV_CAND_DTYPE)
    cl_w = np.zeros((10, P), dtype=np.int32); cl_rs = np.zeros((10, P), dtype=np.int32); cl_re = np.zeros((10, P), dtype=np.int32)
    cl_dist = np.zeros((10, P), dtype=np.float64); cl_mchr = np.zeros((2, P), dtype=np.int32); other_len = np.zeros(P, dtype=np.int32)
    out = COut(chr_len=P, arrays=arrays.ctypes.data, read_state=state.ctypes.data, lookahead_lseq=look.ctypes.data,
               snv=snv.ctypes.data, snv_cap=snv_cap, cl_w=cl_w.ctypes.data, cl_rs=cl_rs.ctypes.data, cl_re=cl_re.ctypes.data,
               cl_dist=cl_dist.ctypes.data, cl_mchr=cl_mchr.ctypes.data, other_len=other_len.ctypes.data)
    ins = np.zeros(1 << 16, dtype=INS_CAND_DTYPE)
    out.ins = ins.ctypes.data; out.ins_cap = len(ins)
    dev = np.zeros(1 << 17, dtype=DEL_EVENT_DTYPE)
    out.del_ev = dev.ctypes.data; out.del_cap = len(dev)
    sve = np.zeros(1 << 18, dtype=SV_EVENT_DTYPE)
    out.sv_ev = sve.ctypes.data; out.sv_cap = len(sve)
    cb = batch.as_c()
    fa = np.ascontiguousarray(fasta, dtype=np.uint8)
    rc = lib().oracle_run_chr(C.byref(params), C.byref(cb), fa.ctypes.data_as(C.c_char_p), P,
                              hez.ctypes.data, mq.ctypes.data, C.byref(out))
    if rc != 0:
        raise RuntimeError(f"oracle_run_chr failed: {rc}")
    assert out.n_snv <= snv_cap
    return OracleResult(arrays, state[:batch.n_reads], out.scan_first, out.scan_last, look, snv[:out.n_snv].copy(),
                        out.snv_ave_rd, cl_w, cl_rs, cl_re, cl_dist, cl_mchr, other_len, ins[:min(out.n_ins, len(ins))].copy(),
                        dev[:min(out.n_del, len(dev))].copy(), sve[:min(out.n_sv, len(sve))].copy())


def format_snv_vcf(params: Params, chr_name: str, fasta: np.ndarray, snv: np.ndarray, ave_rd: float) -> str:
    cap = 512 * (len(snv) + 1)
    buf = C.create_string_buffer(cap)
    fa = np.ascontiguousarray(fasta, dtype=np.uint8)
    s = np.ascontiguousarray(snv)
    n = lib().oracle_format_snv_vcf(C.byref(params), chr_name.encode(), fa.ctypes.data_as(C.c_char_p), s.ctypes.data,
                                    len(s), ave_rd, buf, cap)
    assert n >= 0
    return buf.raw[:n].decode()


def gc_prepass(params: Params, fasta: np.ndarray):
    """(gc_weighted, acgt_weighted) int32 arrays (reference src/GROM.c:1766-1859)."""
    fa = np.ascontiguousarray(fasta, dtype=np.uint8)
    P = int(fa.shape[0])
    gc = np.zeros(P, dtype=np.int32); acgt = np.zeros(P, dtype=np.int32)
    lib().oracle_gc_prepass(C.byref(params), fa.ctypes.data_as(C.c_char_p), P, gc.ctypes.data, acgt.ctypes.data)
    return gc, acgt


def format_ins_vcf(params: Params, chr_name: str, fasta: np.ndarray, ins: np.ndarray) -> str:
    cap = 512 * (len(ins) + 1)
    buf = C.create_string_buffer(cap)
    fa = np.ascontiguousarray(fasta, dtype=np.uint8)
    a = np.ascontiguousarray(ins)
    L = lib()
    L.oracle_format_ins_vcf.argtypes = [C.POINTER(Params), C.c_char_p, C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.c_char_p, C.c_int64]
    L.oracle_format_ins_vcf.restype = C.c_int64
    n = L.oracle_format_ins_vcf(C.byref(params), chr_name.encode(), fa.ctypes.data_as(C.c_char_p), len(fa), a.ctypes.data, len(a), buf, cap)
    assert n >= 0
    return buf.raw[:n].decode()


# ---------------------------------------------------------------- reference runners

def ref_binary(kind: str = "ref") -> str:
    return os.path.join(REF_DIR, "GROM_ref" if kind == "ref" else "GROM_dist")


def have_reference(kind: str = "ref") -> bool:
    return os.path.exists(ref_binary(kind))


def run_reference(bam: str, fasta: str, out_vcf: str, args=(), dump_dir: Optional[str] = None, kind: str = "ref",
                  seed: int = 1, timeout: int = 3600) -> str:
    """Run the reference on (bam, fasta); returns its stdout.  Stale <bam>.mean/<fasta>.info are removed first."""
    for pth in (bam + ".mean", fasta + ".info"):
        if os.path.exists(pth):
            os.remove(pth)
    env = dict(os.environ)
    env["GROM_SEED"] = str(seed)
    if dump_dir:
        os.makedirs(dump_dir, exist_ok=True)
        env["GROM_DUMP_DIR"] = dump_dir
    else:
        env.pop("GROM_DUMP_DIR", None)
    cmd = [ref_binary(kind), "-i", bam, "-r", fasta, "-o", out_vcf, *[str(a) for a in args]]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"reference failed ({r.returncode}): {r.stdout[-2000:]}")
    return r.stdout


def read_mean_file(bam: str) -> Dict[str, int]:
    """<bam>.mean = insert_mean lseq insert_min insert_max mapped_reads (reference src/GROM.c:994-1026)."""
    v = open(bam + ".mean").read().split()
    return dict(insert_mean=int(v[0]), lseq=int(v[1]), insert_min=int(v[2]), insert_max=int(v[3]), mapped_reads=int(v[4]))


def load_scan_dump(dump_dir: str, chr_name: str) -> np.ndarray:
    return np.fromfile(os.path.join(dump_dir, f"scan_{chr_name}.bin"), dtype=SCAN_DTYPE)


def load_reads_dump(dump_dir: str, chr_name: str) -> np.ndarray:
    return np.fromfile(os.path.join(dump_dir, f"reads_{chr_name}.bin"), dtype=READS_DTYPE)


def load_depth_dump(dump_dir: str, chr_name: str) -> np.ndarray:
    a = np.fromfile(os.path.join(dump_dir, f"depth_{chr_name}.bin"), dtype=np.int32)
    return a.reshape(3, -1)


def load_gc_dump(dump_dir: str, chr_name: str) -> np.ndarray:
    a = np.fromfile(os.path.join(dump_dir, f"gc_{chr_name}.bin"), dtype=np.int32)
    return a.reshape(2, -1)


def reference_tables(min_mapq: int = 20):
    """The two tables as the reference binary in oracle/_ref loads them (text files next to it)."""
    from grom_b200 import hostlib
    return hostlib.tables(REF_DIR, min_mapq, write_missing=True)


def normalise_records(lines):
    """VCF record lines with the fields the reference leaves uninitialised masked out: small-insertion
    records print ECO and EOT from candidate-list slots that are never written for that class (reference
    src/GROM.c:16335), so their values are whatever malloc returned."""
    out = []
    for l in lines:
        f = l.rstrip("\n").split("\t")
        if len(f) >= 10 and f[8] == "SPR:SEV:SRD:SCO:ECO:SOT:EOT:SSC:HP":
            v = f[9].split(":")
            v[4] = "*"; v[6] = "*"          # ECO, EOT: never written for this record class
            f[9] = ":".join(v)
        out.append("\t".join(f) + "\n")
    return out


# ---- read-depth CNV path ------------------------------------------------------------------------------------------------
NB = 101


class CCnvCfg(C.Structure):
    _fields_ = [("insert_mean", C.c_int32), ("rd_min_mapq", C.c_int32), ("ploidy", C.c_int32), ("windows_sampling_factor", C.c_int32),
                ("min_win", C.c_int64), ("max_win", C.c_int64), ("sample_cap", C.c_int64), ("seed", C.c_uint32), ("reserved", C.c_uint32),
                ("rd_pval_threshold", C.c_double)]


class CCnvOut(C.Structure):
    _fields_ = [("n_nblk", C.c_long), ("nb_s", C.POINTER(C.c_long)), ("nb_e", C.POINTER(C.c_long)),
                ("n_rep", C.c_long), ("rep_t", C.POINTER(C.c_long)), ("rep_s", C.POINTER(C.c_long)), ("rep_e", C.POINTER(C.c_long)),
                ("chr_ave", C.c_double), ("chr_sd", C.c_double), ("rep_ave", C.c_double * 10), ("rep_sd", C.c_double * 10),
                ("rep_cnt", C.c_long * 10), ("biased", C.c_long), ("blk_ave", C.c_double),
                ("n_sblk", C.c_long), ("sb_s", C.POINTER(C.c_long)), ("sb_e", C.POINTER(C.c_long)),
                ("mq_mean", C.POINTER(C.c_int)), ("z", C.POINTER(C.c_double)), ("mask", C.POINTER(C.c_ubyte)),
                ("win_sd", C.POINTER(C.c_double)), ("win_cnt", C.POINTER(C.c_long)),
                ("ave", C.c_double * (2 * NB)), ("sd", C.c_double * (2 * NB)), ("del_thr", C.c_double * (2 * NB)), ("dup_thr", C.c_double * (2 * NB)),
                ("windows", C.c_long * (2 * NB)), ("n_high", C.c_long * NB), ("n_low", C.c_long * NB),
                ("n_call", C.c_long * 2), ("call_s", C.POINTER(C.c_long) * 2), ("call_e", C.POINTER(C.c_long) * 2),
                ("call_z", C.POINTER(C.c_double) * 2), ("call_cn", C.POINTER(C.c_double) * 2), ("call_cs", C.POINTER(C.c_double) * 2),
                ("call_p", C.POINTER(C.c_double) * 2)]


CALL_DTYPE = np.dtype([("start", np.int64), ("end", np.int64), ("z", np.float64), ("cn", np.float64), ("cs", np.float64), ("p", np.float64)])


@dataclass
class CnvResult:
    nblocks: np.ndarray         # [n, 2] N-run blocks
    repeats: np.ndarray         # [n, 3] type, start, end
    chr_ave: float
    chr_sd: float
    rep_ave: np.ndarray
    rep_sd: np.ndarray
    rep_cnt: np.ndarray
    biased: int
    blk_ave: float
    sample_blocks: np.ndarray   # [n, 2]
    mq_mean: np.ndarray
    z: np.ndarray
    mask: np.ndarray
    win_sd: np.ndarray
    win_cnt: np.ndarray
    ave: np.ndarray             # [2, 101]
    sd: np.ndarray
    del_thr: np.ndarray
    dup_thr: np.ndarray
    windows: np.ndarray
    n_high: np.ndarray
    n_low: np.ndarray
    dels: np.ndarray            # CALL_DTYPE
    dups: np.ndarray
    vcf: str


def _np(ptr, n, dtype):
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype).copy() if n else np.zeros(0, dtype=dtype)


def cnv_cfg(params: Params, ploidy: Optional[int] = None, seed: int = 1, sample_cap: int = 100000, min_win: int = 100, max_win: int = 10000) -> CCnvCfg:
    return CCnvCfg(params.insert_mean, params.rd_min_mapq, params.ploidy if ploidy is None else ploidy, params.windows_sampling_factor,
                   min_win, max_win, sample_cap, seed, 0, params.rd_pval_threshold)


def cnv_run(params: Params, chr_name: str, fasta: np.ndarray, gc: np.ndarray, acgt: np.ndarray, rd_mq: np.ndarray, rd_rd: np.ndarray,
            rd_low: np.ndarray, ploidy: Optional[int] = None, seed: int = 1, sample_cap: int = 100000, min_win: int = 100,
            max_win: int = 10000) -> CnvResult:
    """CNV oracle on the raw CNV depth arrays (rd_mq = MAPQ sums, before the in-place mean)."""
    from grom_b200 import hostlib
    L = lib()
    L.oracle_cnv_run.argtypes = [C.POINTER(CCnvCfg), C.c_char_p, C.c_long] + [C.c_void_p] * 7 + [C.c_int, C.POINTER(CCnvOut)]
    L.oracle_format_cnv_vcf.argtypes = [C.POINTER(CCnvCfg), C.c_char_p, C.POINTER(CCnvOut)]
    L.oracle_format_cnv_vcf.restype = C.c_void_p
    L.oracle_cnv_free.argtypes = [C.POINTER(CCnvOut)]
    cfg = cnv_cfg(params, ploidy, seed, sample_cap, min_win, max_win)
    pv, sdv = hostlib.pval2sd()
    arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (gc, acgt, rd_mq, rd_rd, rd_low)]
    fa = np.ascontiguousarray(fasta, dtype=np.uint8)
    fa0 = np.concatenate([fa, np.zeros(2, dtype=np.uint8)])
    out = CCnvOut()
    rc = L.oracle_cnv_run(C.byref(cfg), fa0.ctypes.data_as(C.c_char_p), len(fa), *[a.ctypes.data for a in arrs],
                          pv.ctypes.data, sdv.ctypes.data, len(pv), C.byref(out))
    if rc != 0:
        raise RuntimeError("oracle_cnv_run failed (contig too short)")
    P, nw = len(fa), max_win + 1

    def calls(k):
        n = out.n_call[k]
        r = np.zeros(n, dtype=CALL_DTYPE)
        for f, src in (("start", out.call_s), ("end", out.call_e), ("z", out.call_z), ("cn", out.call_cn), ("cs", out.call_cs), ("p", out.call_p)):
            r[f] = _np(src[k], n, r.dtype[f])
        return r
    txt = L.oracle_format_cnv_vcf(C.byref(cfg), chr_name.encode(), C.byref(out))
    vcf = C.string_at(txt).decode()
    res = CnvResult(
        nblocks=np.stack([_np(out.nb_s, out.n_nblk, np.int64), _np(out.nb_e, out.n_nblk, np.int64)], 1),
        repeats=np.stack([_np(out.rep_t, out.n_rep, np.int64), _np(out.rep_s, out.n_rep, np.int64), _np(out.rep_e, out.n_rep, np.int64)], 1),
        chr_ave=out.chr_ave, chr_sd=out.chr_sd, rep_ave=np.array(out.rep_ave), rep_sd=np.array(out.rep_sd), rep_cnt=np.array(out.rep_cnt),
        biased=out.biased, blk_ave=out.blk_ave,
        sample_blocks=np.stack([_np(out.sb_s, out.n_sblk, np.int64), _np(out.sb_e, out.n_sblk, np.int64)], 1),
        mq_mean=_np(out.mq_mean, P, np.int32), z=_np(out.z, P, np.float64), mask=_np(out.mask, P, np.uint8),
        win_sd=_np(out.win_sd, nw, np.float64), win_cnt=_np(out.win_cnt, nw, np.int64),
        ave=np.array(out.ave).reshape(2, NB), sd=np.array(out.sd).reshape(2, NB), del_thr=np.array(out.del_thr).reshape(2, NB),
        dup_thr=np.array(out.dup_thr).reshape(2, NB), windows=np.array(out.windows).reshape(2, NB), n_high=np.array(out.n_high),
        n_low=np.array(out.n_low), dels=calls(0), dups=calls(1), vcf=vcf)
    L.oracle_cnv_free(C.byref(out))
    return res


def load_cnvpre_dump(dump_dir: str, chr_name: str) -> dict:
    a = np.fromfile(os.path.join(dump_dir, f"cnvpre_{chr_name}.bin"), dtype=np.int64)
    f = a.view(np.float64)
    i = 0
    n = int(a[i]); i += 1
    nblocks = a[i:i + 2 * n].reshape(n, 2).copy(); i += 2 * n
    n = int(a[i]); i += 1
    repeats = a[i:i + 3 * n].reshape(n, 3).copy(); i += 3 * n
    d = dict(nblocks=nblocks, repeats=repeats, chr_ave=f[i], chr_sd=f[i + 1]); i += 2
    d["rep_ave"] = f[i:i + 10].copy(); d["rep_sd"] = f[i + 10:i + 20].copy(); d["rep_cnt"] = a[i + 20:i + 30].copy(); i += 30
    d["biased"] = int(a[i]); d["blk_ave"] = f[i + 1]; i += 2
    n = int(a[i]); i += 1
    d["sample_blocks"] = a[i:i + 2 * n].reshape(n, 2).copy()
    return d


def load_cnv_dump(dump_dir: str, chr_name: str) -> dict:
    raw = np.fromfile(os.path.join(dump_dir, f"cnv_{chr_name}.bin"), dtype=np.uint8)
    P, nwin, nb, nd, nu = (int(x) for x in raw[:40].view(np.int64))
    o = 40
    d = {}

    def take(n, dt):
        nonlocal o
        sz = n * np.dtype(dt).itemsize
        v = raw[o:o + sz].view(dt).copy()
        o += sz
        return v
    d["z"] = take(P, np.float64); d["mask"] = take(P, np.uint8)
    d["win_sd"] = take(nwin, np.float64); d["win_cnt"] = take(nwin, np.int64)
    for k in ("ave", "sd", "del_thr", "dup_thr"):
        d[k] = take(2 * nb, np.float64).reshape(2, nb)
    d["windows"] = take(2 * nb, np.int64).reshape(2, nb)
    d["n_high"] = take(nb, np.int64); d["n_low"] = take(nb, np.int64)
    rec = np.dtype([("start", np.int64), ("end", np.int64), ("z", np.float64), ("cn", np.float64), ("cs", np.float64)])
    d["dels"] = take(nd, rec); d["dups"] = take(nu, rec)
    return d


def load_svlist_dump(dump_dir: str, chr_name: str) -> dict:
    """The reference's candidate lists at the end of the per-position scan (hook at src/GROM.c:15164), in the product's record layouts
    (grom_b200.params.SV_PAIR_DTYPE / SV_EVENT_DTYPE).  End sides that were never filled carry uninitialised memory in the reference;
    they are zeroed here (pos == -1 marks them)."""
    from grom_b200.params import SV_PAIR_DTYPE
    out = {}

    def cols(path, spec):
        raw = np.fromfile(path, dtype=np.uint8)
        n = int(raw[:8].view(np.int64)[0]); o = 8; r = {}
        for name, dt in spec:
            sz = n * np.dtype(dt).itemsize
            r[name] = raw[o:o + sz].view(dt).copy(); o += sz
        return n, r
    i4, f8 = np.int32, np.float64
    for name in ("dup", "del", "inv_f", "inv_r"):
        n, c = cols(os.path.join(dump_dir, f"svl_{name}_{chr_name}.bin"),
                    [("start", i4), ("end", i4), ("dist", f8), ("sb", f8), ("sh", f8), ("sc", i4), ("srd", i4), ("sw", i4), ("srs", i4), ("sre", i4),
                     ("sol", i4), ("eb", f8), ("eh", f8), ("ec", i4), ("erd", i4), ("ew", i4), ("ers", i4), ("ere", i4), ("eol", i4)])
        a = np.zeros(n, dtype=SV_PAIR_DTYPE)
        a["dist"] = c["dist"]
        for side, pos, pre in (("start", "start", "s"), ("end", "end", "e")):
            a[side]["pos"] = c[pos]
            for f, k in (("binom", "b"), ("hez", "h"), ("conc", "c"), ("rd", "rd"), ("weight", "w"), ("read_start", "rs"), ("read_end", "re"), ("other_len", "ol")):
                a[side][f] = c[pre + k]
        dead = a["end"]["pos"] == -1
        for f in ("binom", "hez", "conc", "rd", "weight", "read_start", "read_end", "other_len"):
            a["end"][f][dead] = 0
        out[name] = a
    for name in ("ctx_f", "ctx_r"):
        n, c = cols(os.path.join(dump_dir, f"svl_{name}_{chr_name}.bin"),
                    [("pos", i4), ("binom", f8), ("hez", f8), ("mchr", i4), ("mpos", i4), ("conc", i4), ("rd", i4), ("weight", i4), ("read_start", i4),
                     ("read_end", i4), ("other_len", i4)])
        a = np.zeros(n, dtype=SV_EVENT_DTYPE)
        for f in ("pos", "binom", "hez", "mchr", "conc", "rd", "weight", "read_start", "read_end", "other_len"):
            a[f] = c[f]
        a["dist"] = c["mpos"]                       # the reference's list keeps (int) of the running mate position
        a["cls"] = 8 if name == "ctx_f" else 9
        out[name] = a
    n, c = cols(os.path.join(dump_dir, f"svl_ins_{chr_name}.bin"),
                [("start", i4), ("end", i4), ("sb", f8), ("eb", f8), ("si", i4), ("ei", i4), ("srd", i4), ("erd", i4), ("sc", i4), ("ec", i4), ("sol", i4), ("eol", i4)])
    a = np.zeros(n, dtype=SV_PAIR_DTYPE)
    for side, pos, pre in (("start", "start", "s"), ("end", "end", "e")):
        a[side]["pos"] = c[pos]
        for f, k in (("binom", "b"), ("weight", "i"), ("rd", "rd"), ("conc", "c"), ("other_len", "ol")):
            a[side][f] = c[pre + k]
        dead = a[side]["pos"] == -1
        for f in ("binom", "weight", "rd", "conc", "other_len"):
            a[side][f][dead] = 0
    out["ins"] = a
    return out


def normalise_sv_lists(lists: dict) -> dict:
    """Product lists in the comparable form of load_svlist_dump: unfilled sides zeroed, insertion sides carry no hez / read range,
    ctx distance truncated like the reference's int list."""
    out = {}
    for k, a in lists.items():
        a = a.copy()
        if k in ("ctx_f", "ctx_r"):
            a["dist"] = a["dist"].astype(np.int32)
            a["reserved"] = 0
        else:
            for side in ("start", "end"):
                a[side]["reserved"] = 0
                dead = a[side]["pos"] == -1
                for f in ("binom", "hez", "conc", "rd", "weight", "read_start", "read_end", "other_len"):
                    a[side][f][dead] = 0
                if k == "ins":
                    for f in ("hez", "read_start", "read_end"):
                        a[side][f] = 0
            if k == "ins":
                a["dist"] = 0
        out[k] = a
    return out
