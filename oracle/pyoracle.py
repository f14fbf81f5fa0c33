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
from grom_b200.params import DEL_EVENT_DTYPE, GA, GA_COUNT, GA_NAMES, INS_CAND_DTYPE, Params, SNV_CAND_DTYPE  # noqa: E402
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
                ("del_ev", C.c_void_p), ("del_cap", C.c_int64), ("n_del", C.c_int64)]


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
    cb = batch.as_c()
    fa = np.ascontiguousarray(fasta, dtype=np.uint8)
    rc = lib().oracle_run_chr(C.byref(params), C.byref(cb), fa.ctypes.data_as(C.c_char_p), P,
                              hez.ctypes.data, mq.ctypes.data, C.byref(out))
    if rc != 0:
        raise RuntimeError(f"oracle_run_chr failed: {rc}")
    assert out.n_snv <= snv_cap
    return OracleResult(arrays, state[:batch.n_reads], out.scan_first, out.scan_last, look, snv[:out.n_snv].copy(),
                        out.snv_ave_rd, cl_w, cl_rs, cl_re, cl_dist, cl_mchr, other_len, ins[:min(out.n_ins, len(ins))].copy(),
                        dev[:min(out.n_del, len(dev))].copy())


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
