"""ctypes mirror of include/grom_params.h (struct grom_params, GA_* array numbering, grom_snv_cand)."""
from __future__ import annotations

import ctypes as C

import numpy as np


class Params(C.Structure):
    _fields_ = [
        ("insert_mean", C.c_int32), ("insert_min", C.c_int32), ("insert_max", C.c_int32), ("lseq", C.c_int32),
        ("min_mapq", C.c_int32), ("rd_min_mapq", C.c_int32), ("min_base_qual", C.c_int32), ("min_snv", C.c_int32),
        ("min_disc", C.c_int32), ("sc_min", C.c_int32), ("rmdup", C.c_int32), ("rmdup_list_len", C.c_int32),
        ("splitread", C.c_int32), ("max_split_loss", C.c_int32), ("min_sr_len", C.c_int32), ("overlap_mult", C.c_int32),
        ("other_len", C.c_int32), ("read_name_len", C.c_int32), ("indel_i_seq_len", C.c_int32), ("max_cigar_ops", C.c_int32),
        ("ploidy", C.c_int32), ("gender", C.c_int32), ("max_trials", C.c_int32), ("add_factor", C.c_int32),
        ("min_snv_ratio", C.c_double), ("min_ave_bq", C.c_double), ("snv_rd_min_factor", C.c_double),
        ("high_cov_min_snv_ratio", C.c_double), ("pval_threshold1", C.c_double), ("pval_threshold", C.c_double),
        ("pval_insertion1", C.c_double), ("pval_insertion", C.c_double), ("rd_pval_threshold", C.c_double),
        ("max_evidence_ratio", C.c_double), ("min_sv_ratio", C.c_double), ("min_indel_ratio", C.c_double),
        ("windows_sampling_factor", C.c_int32), ("rand_seed", C.c_int32), ("min_rd_window_len", C.c_int32),
        ("max_rd_window_len", C.c_int32), ("sample_lists_len", C.c_int32), ("reserved0", C.c_int32),
    ]

    @classmethod
    def default(cls, **kw) -> "Params":
        p = cls(insert_mean=400, insert_min=300, insert_max=500, lseq=150, min_mapq=20, rd_min_mapq=20, min_base_qual=20,
                min_snv=3, min_disc=3, sc_min=1, rmdup=0, rmdup_list_len=10000, splitread=1, max_split_loss=20, min_sr_len=30,
                overlap_mult=1, other_len=50, read_name_len=50, indel_i_seq_len=50, max_cigar_ops=1000, ploidy=2, gender=0,
                max_trials=1000, add_factor=6, min_snv_ratio=0.2, min_ave_bq=15, snv_rd_min_factor=1.75,
                high_cov_min_snv_ratio=0.4, pval_threshold1=0.001, pval_threshold=0.001, pval_insertion1=0.01,
                pval_insertion=1e-10, rd_pval_threshold=1e-9, max_evidence_ratio=0.25, min_sv_ratio=0.05,
                min_indel_ratio=0.125, windows_sampling_factor=2, rand_seed=1, min_rd_window_len=100,
                max_rd_window_len=10000, sample_lists_len=100000, reserved0=0)
        for k, v in kw.items():
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
        return p

    @property
    def window_len(self) -> int:
        a = self.overlap_mult * 8 * (2 * self.insert_mean - 1)
        b = self.overlap_mult * 8 * (self.insert_max + 1)
        return 2 * max(a, b)

    @property
    def first_pos(self) -> int:
        return self.window_len // 4 + 1


GA_NAMES = [
    "snv_a", "snv_c", "snv_g", "snv_t", "snvlow_a", "snvlow_c", "snvlow_g", "snvlow_t",
    "bq", "bq_all", "mq", "mq_all", "bq_rc", "mq_rc", "rc_all",
    "pir_a", "pir_c", "pir_g", "pir_t", "fs_a", "fs_c", "fs_g", "fs_t",
    "rd", "sc_left", "sc_left_rd", "sc_right", "sc_right_rd", "sc_rd",
    "ctx_sc_left", "ctx_sc_left_rd", "ctx_sc_right", "ctx_sc_right_rd", "ctx_sc_rd",
    "indel_sc_left", "indel_sc_left_rd", "indel_sc_right", "indel_sc_right_rd", "indel_sc_rd",
    "conc", "ins", "munmapped_f", "munmapped_r",
    "indel_i", "indel_idist", "indel_d_f", "indel_d_fdist", "indel_d_f_rd", "indel_d_r", "indel_d_rdist", "indel_d_r_rd",
    "rd_mq", "rd_rd", "rd_low", "gc", "acgt",
]
GA = {n: i for i, n in enumerate(GA_NAMES)}
GA_COUNT = len(GA_NAMES)
GA_PILEUP_COUNT = 23

SNV_CAND_DTYPE = np.dtype([("pos", np.int32), ("base", np.int32), ("ratio", np.float64), ("pr", np.float64),
                           ("hez", np.float64), ("v", np.int32, (GA_PILEUP_COUNT,)), ("reserved", np.int32)], align=True)
assert SNV_CAND_DTYPE.itemsize == 128, SNV_CAND_DTYPE.itemsize

INS_CAND_DTYPE = np.dtype([("pos", np.int32), ("dist", np.int32), ("pr", np.float64), ("hez", np.float64), ("conc", np.int32),
                           ("weight", np.int32), ("rd", np.int32), ("sc", np.int32), ("other_len", np.int32), ("reserved", np.int32),
                           ("seq", "S56")], align=True)
assert INS_CAND_DTYPE.itemsize == 104, INS_CAND_DTYPE.itemsize

DEL_EVENT_DTYPE = np.dtype([("pos", np.int32), ("kind", np.int32), ("pr", np.float64), ("hez", np.float64), ("conc", np.int32),
                            ("weight", np.int32), ("rd", np.int32), ("sc", np.int32), ("other_len", np.int32), ("rdist", np.int32)], align=True)
assert DEL_EVENT_DTYPE.itemsize == 48, DEL_EVENT_DTYPE.itemsize

SV_EVENT_DTYPE = np.dtype([("pos", np.int32), ("cls", np.int32), ("binom", np.float64), ("hez", np.float64), ("dist", np.float64),
                           ("weight", np.int32), ("rd", np.int32), ("conc", np.int32), ("read_start", np.int32), ("read_end", np.int32),
                           ("other_len", np.int32), ("mchr", np.int32), ("reserved", np.int32)], align=True)
assert SV_EVENT_DTYPE.itemsize == 64, SV_EVENT_DTYPE.itemsize
SV_CLASSES = ["del_f", "del_r", "dup_f", "dup_r", "inv_f1", "inv_r1", "inv_f2", "inv_r2", "ctx_f", "ctx_r", "ins_l", "ins_r"]
SV_SIDE_DTYPE = np.dtype([("pos", np.int32), ("weight", np.int32), ("rd", np.int32), ("conc", np.int32), ("read_start", np.int32),
                          ("read_end", np.int32), ("other_len", np.int32), ("reserved", np.int32), ("binom", np.float64), ("hez", np.float64)], align=True)
SV_PAIR_DTYPE = np.dtype([("start", SV_SIDE_DTYPE), ("end", SV_SIDE_DTYPE), ("dist", np.float64)], align=True)
assert SV_SIDE_DTYPE.itemsize == 48 and SV_PAIR_DTYPE.itemsize == 104, (SV_SIDE_DTYPE.itemsize, SV_PAIR_DTYPE.itemsize)

CTX_RECORD_DTYPE = np.dtype([("type", np.int32), ("chr", np.int32), ("pos", np.int32), ("rd", np.int32), ("conc", np.int32), ("other_len", np.int32),
                             ("mchr", np.int32), ("mpos", np.int32), ("read_start", np.int32), ("read_end", np.int32), ("mate_id", np.int32),
                             ("keep", np.int32), ("binom", np.float64), ("evidence", np.float64), ("hez", np.float64)], align=True)
assert CTX_RECORD_DTYPE.itemsize == 72, CTX_RECORD_DTYPE.itemsize

CNV_CALL_DTYPE = np.dtype([("start", np.int64), ("end", np.int64), ("kind", np.int32), ("reserved", np.int32), ("z", np.float64),
                           ("pvalue", np.float64), ("cn", np.float64), ("cn_sd", np.float64)], align=True)
assert CNV_CALL_DTYPE.itemsize == 56, CNV_CALL_DTYPE.itemsize
