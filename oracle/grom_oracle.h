/* oracle/grom_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's per-chromosome hot path
 * (reference src/GROM.c:1432-18228 count_discordant_pairs) over chromosome-length
 * arrays instead of the reference's sliding window.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the
 * product (grom_b200/csrc) never does.
 *
 * Parity pinning: every function here is checked in tests/test_oracle_vs_reference.py
 * against array dumps of the reference itself (oracle/_ref/GROM_ref, the reference's own
 * translation unit built with dump hooks) and its VCF output; the binomial tables are
 * additionally pinned against the reference's golden tilapia VCF (tests/test_tables.py).
 */
#ifndef GROM_ORACLE_H
#define GROM_ORACLE_H
#include <stdint.h>
#include "grom_reads.h"
#include "grom_params.h"

#ifdef __cplusplus
extern "C" {
#endif

/* read_state[i]: 0 = never reaches the evidence code (before W/4+1, UNMAP or DUP flag),
 *                1 = applied, 2 = dropped by the -M duplicate filter */
typedef struct oracle_chr_out {
    int64_t  chr_len;
    int32_t *arrays;            /* [GA_COUNT][chr_len], caller-allocated, zeroed by oracle */
    uint8_t *read_state;        /* [n_reads] */
    int32_t  scan_first;        /* first scanned position (W/4+1), -1 if nothing is scanned */
    int32_t  scan_last;         /* last scanned position (inclusive) */
    int32_t *lookahead_lseq;    /* [chr_len] cdp_lseq seen by the scan at each scanned position (0 elsewhere) */
    grom_snv_cand *snv;         /* caller-allocated [snv_cap] */
    int64_t  snv_cap, n_snv;
    double   snv_ave_rd;        /* mean depth used by the SNV emission filter (src/GROM.c:15035-15043) */
    /* breakpoint clusters, 10 classes in the order del_f del_r dup_f dup_r inv_f1 inv_r1 inv_f2 inv_r2 ctx_f ctx_r
     * (src/GROM.c:7955-10953); all caller-allocated, NULL = not wanted */
    int32_t *cl_w;              /* [10][chr_len] weight */
    int32_t *cl_rs, *cl_re;     /* [10][chr_len] first / last read position */
    double  *cl_dist;           /* [10][chr_len] running-mean distance (ctx: signed mate position) */
    int32_t *cl_mchr;           /* [2][chr_len]  mate contig of ctx_f / ctx_r */
    int32_t *other_len;         /* [chr_len] index of the first empty `other` slot */
    grom_ins_cand *ins;         /* small-insertion candidates (src/GROM.c:11340-11453), caller-allocated [ins_cap] or NULL */
    int64_t  ins_cap, n_ins;
    grom_del_event *del_ev;     /* small-deletion scan events in scan order (src/GROM.c:11454-11745), caller-allocated or NULL */
    int64_t  del_cap, n_del;
    grom_sv_event *sv_ev;       /* structural-variant gate events in scan order (src/GROM.c:11750-13541), caller-allocated or NULL */
    int64_t  sv_cap, n_sv;
} oracle_chr_out;

int oracle_run_chr(const grom_params *p, const grom_read_batch *b, const char *fasta, int64_t chr_len,
                   const double *hez_tbl, const double *mq_tbl, oracle_chr_out *out);

/* GC / ACGT percentages of the triangular window (src/GROM.c:1766-1859) into arrays[GA_GC], arrays[GA_ACGT] */
void oracle_gc_prepass(const grom_params *p, const char *fasta, int64_t chr_len, int32_t *gc, int32_t *acgt);

/* format the SNV VCF lines the reference prints at src/GROM.c:15082-15095 into buf; returns bytes written */
int64_t oracle_format_snv_vcf(const grom_params *p, const char *chr_name, const char *fasta,
                              const grom_snv_cand *snv, int64_t n_snv, double ave_rd, char *buf, int64_t cap);

/* small-insertion VCF records as printed at src/GROM.c:16253-16340 (emission filter, homopolymer rule, text) */
int64_t oracle_format_ins_vcf(const grom_params *p, const char *chr_name, const char *fasta, int64_t chr_len,
                              const grom_ins_cand *ins, int64_t n_ins, char *buf, int64_t cap);

/* ---- read-depth CNV path (grom_oracle_cnv.c; src/GROM.c:1684-1764, 16633-17500, 18228-20355) ---- */
#define ORACLE_CNV_BINS 101
typedef struct oracle_cnv_cfg {
    int32_t insert_mean, rd_min_mapq;
    int32_t ploidy;                     /* effective: halved on chrx / x when -g 1 (src/GROM.c:17024-17035) */
    int32_t windows_sampling_factor;    /* -A */
    int64_t min_win, max_win;           /* g_min_rd_window_len 100, g_max_rd_window_len 10000 */
    int64_t sample_cap;                 /* g_sample_lists_len 100000 */
    uint32_t seed; uint32_t reserved;   /* srand() seed (the white-box build takes it from GROM_SEED) */
    double  rd_pval_threshold;          /* -V */
} oracle_cnv_cfg;

typedef struct oracle_cnv_out {
    long n_nblk, *nb_s, *nb_e;                          /* N-run blocks */
    long n_rep, *rep_t, *rep_s, *rep_e;                 /* dinucleotide-repeat runs */
    double chr_ave, chr_sd, rep_ave[10], rep_sd[10]; long rep_cnt[10], biased; double blk_ave;
    long n_sblk, *sb_s, *sb_e;                          /* sample blocks */
    int *mq_mean;                                       /* [len] rd_mq after the in-place mean */
    double *z; unsigned char *mask;                     /* [len] stdev_list, rd_low_acgt_or_windows_list */
    double *win_sd; long *win_cnt;                      /* [max_win+1] */
    double ave[2][ORACLE_CNV_BINS], sd[2][ORACLE_CNV_BINS], del_thr[2][ORACLE_CNV_BINS], dup_thr[2][ORACLE_CNV_BINS];
    long windows[2][ORACLE_CNV_BINS], n_high[ORACLE_CNV_BINS], n_low[ORACLE_CNV_BINS];
    long n_call[2], *call_s[2], *call_e[2];             /* [0] deletions, [1] duplications */
    double *call_z[2], *call_cn[2], *call_cs[2], *call_p[2];
} oracle_cnv_out;

int oracle_cnv_run(const oracle_cnv_cfg *c, const char *fasta, long len, const int *gc, const int *acgt, const int *rd_mq_sum,
                   const int *rd_rd, const int *rd_low, const double *p2s_p, const double *p2s_sd, int n_p2s, oracle_cnv_out *o);
void oracle_cnv_free(oracle_cnv_out *o);
char *oracle_format_cnv_vcf(const oracle_cnv_cfg *c, const char *chr, const oracle_cnv_out *o);
long oracle_bisect_left(const int *a, int v, long s, long e);
long oracle_bisect_right(const int *a, int v, long s, long e);
long oracle_bisect_left_double(const double *a, double v, long s, long e);
long oracle_bisect_right_double(const double *a, double v, long s, long e);

#ifdef __cplusplus
}
#endif
#endif
