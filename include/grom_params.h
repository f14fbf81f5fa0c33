/* grom_params.h -- every global the reference's hot path reads, as one plain struct.
 *
 * The reference keeps its configuration in ~90 file-scope globals set by getopt
 * (reference src/GROM.c:710-974, 21907-22106) plus the library statistics computed
 * by find_insert_mean (src/GROM.c:1205-1318) and the window sizing of
 * src/GROM.c:22260-22290.  This struct carries the subset that
 * count_discordant_pairs / detect_del_dup consume, with the reference's
 * effective defaults (the initialisers, not the help text -- SURVEY.md §5).
 *
 * Also: the canonical numbering of the per-position count arrays ("GA_*"),
 * shared by the CUDA library, the oracle and the tests.
 */
#ifndef GROM_PARAMS_H
#define GROM_PARAMS_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct grom_params {
    /* library statistics (find_insert_mean, src/GROM.c:1205-1318; cached in <bam>.mean) */
    int32_t insert_mean;        /* g_insert_mean (raised to lseq if smaller, src/GROM.c:22260) */
    int32_t insert_min;         /* g_insert_min_size */
    int32_t insert_max;         /* g_insert_max_size */
    int32_t lseq;               /* g_lseq, median read length */
    /* thresholds */
    int32_t min_mapq;           /* -q  g_min_mapq            20 (src/GROM.c:803) */
    int32_t rd_min_mapq;        /*     g_rd_min_mapq  = -q      (src/GROM.c:22102) */
    int32_t min_base_qual;      /* -b  g_min_base_qual       20 (src/GROM.c:892) */
    int32_t min_snv;            /* -n  g_min_snv              3 */
    int32_t min_disc;           /* -d  g_min_disc             3 */
    int32_t sc_min;             /*     g_sc_min               1 */
    int32_t rmdup;              /* -M  g_rmdup                0 */
    int32_t rmdup_list_len;     /*     g_rmdup_list_len   10000 */
    int32_t splitread;          /*     g_splitread            1 */
    int32_t max_split_loss;     /*     g_max_split_loss      20 */
    int32_t min_sr_len;         /*     g_min_sr_len          30 */
    int32_t overlap_mult;       /*     g_overlap_mult         1 */
    int32_t other_len;          /*     g_other_len           50 */
    int32_t read_name_len;      /*     g_read_name_len       50 */
    int32_t indel_i_seq_len;    /*     g_indel_i_seq_len     50 */
    int32_t max_cigar_ops;      /*     cdp_c_type_len      1000 (src/GROM.c:2901) */
    int32_t ploidy;             /* -p  g_ploidy               2 */
    int32_t gender;             /* -g  g_gender               0 */
    int32_t max_trials;         /*     g_max_trials        1000 */
    int32_t add_factor;         /*     cdp_add_factor         6 (src/GROM.c:1548) */
    double  min_snv_ratio;      /* -a  g_min_snv_ratio      0.2 */
    double  min_ave_bq;         /* -x  g_min_ave_bq          15 */
    double  snv_rd_min_factor;  /*     g_snv_rd_min_factor 1.75 */
    double  high_cov_min_snv_ratio; /* g_high_cov_min_snv_ratio 0.4 */
    double  pval_threshold1;    /*     g_pval_threshold1   declared 0.01 but overwritten with the -v value at start-up (src/GROM.c:22101): 0.001 */
    double  pval_threshold;     /* -v  g_pval_threshold   0.001 */
    double  pval_insertion1;    /*     g_pval_insertion1   0.01 */
    double  pval_insertion;     /* -e  g_pval_insertion   1e-10 */
    double  rd_pval_threshold;  /* -V  g_rd_pval_threshold 1e-9 */
    double  max_evidence_ratio; /*     g_max_evidence_ratio 0.25 */
    double  min_sv_ratio;       /*     g_min_sv_ratio      0.05 */
    double  min_indel_ratio;    /*     g_min_indel_ratio  0.125 */
    int32_t windows_sampling_factor; /* -A g_windows_sampling_factor 2 */
    int32_t rand_seed;          /*     srand() argument; the reference uses time() (src/GROM.c:1584), which only matters once a depth
                                       sample list overflows sample_lists_len */
    int32_t min_rd_window_len;  /*     g_min_rd_window_len  100 */
    int32_t max_rd_window_len;  /*     g_max_rd_window_len  10000 */
    int32_t sample_lists_len;   /*     g_sample_lists_len   100000 */
    int32_t reserved0;
} grom_params;

/* g_one_base_rd_len (src/GROM.c:22282-22290): the reference's sliding-window length */
static inline int32_t grom_window_len(const grom_params *p)
{
    int32_t a = p->overlap_mult * 8 * (2 * p->insert_mean - 1);
    int32_t b = p->overlap_mult * 8 * (p->insert_max + 1);
    return 2 * (b > a ? b : a);
}
/* first position at which reads are consumed and positions scanned: W/4 + 1 (src/GROM.c:2918, 6406) */
static inline int32_t grom_first_pos(const grom_params *p) { return grom_window_len(p) / 4 + 1; }

static inline void grom_params_default(grom_params *p)
{
    p->insert_mean = 400; p->insert_min = 300; p->insert_max = 500; p->lseq = 150;
    p->min_mapq = 20; p->rd_min_mapq = 20; p->min_base_qual = 20; p->min_snv = 3; p->min_disc = 3; p->sc_min = 1;
    p->rmdup = 0; p->rmdup_list_len = 10000; p->splitread = 1; p->max_split_loss = 20; p->min_sr_len = 30;
    p->overlap_mult = 1; p->other_len = 50; p->read_name_len = 50; p->indel_i_seq_len = 50; p->max_cigar_ops = 1000;
    p->ploidy = 2; p->gender = 0; p->max_trials = 1000; p->add_factor = 6;
    p->min_snv_ratio = 0.2; p->min_ave_bq = 15; p->snv_rd_min_factor = 1.75; p->high_cov_min_snv_ratio = 0.4;
    p->pval_threshold1 = 0.001; p->pval_threshold = 0.001; p->pval_insertion1 = 0.01; p->pval_insertion = 1e-10;
    p->rd_pval_threshold = 1e-9; p->max_evidence_ratio = 0.25; p->min_sv_ratio = 0.05; p->min_indel_ratio = 0.125;
    p->windows_sampling_factor = 2; p->rand_seed = 1; p->min_rd_window_len = 100; p->max_rd_window_len = 10000;
    p->sample_lists_len = 100000; p->reserved0 = 0;
}

/* ---- canonical per-position int32 arrays (one value per reference position) ----
 * 0..22  pileup            src/GROM.c:6740-7059
 * 23..38 physical depth + soft-clip classes   src/GROM.c:7067-7181
 * 39..42 pair range-adds   src/GROM.c:8345-8365, 8856-8872, 10908-10951
 * 43..50 small-indel primary slots            src/GROM.c:7187-7423
 * 51..53 CNV depth         src/GROM.c:6605-6671
 * 54..55 GC / ACGT percentages                src/GROM.c:1766-1859
 */
enum {
    GA_SNV_A = 0, GA_SNV_C, GA_SNV_G, GA_SNV_T,
    GA_SNVLOW_A, GA_SNVLOW_C, GA_SNVLOW_G, GA_SNVLOW_T,
    GA_BQ, GA_BQ_ALL, GA_MQ, GA_MQ_ALL, GA_BQ_RC, GA_MQ_RC, GA_RC_ALL,
    GA_PIR_A, GA_PIR_C, GA_PIR_G, GA_PIR_T,
    GA_FS_A, GA_FS_C, GA_FS_G, GA_FS_T,
    GA_RD, GA_SC_LEFT, GA_SC_LEFT_RD, GA_SC_RIGHT, GA_SC_RIGHT_RD, GA_SC_RD,
    GA_CTX_SC_LEFT, GA_CTX_SC_LEFT_RD, GA_CTX_SC_RIGHT, GA_CTX_SC_RIGHT_RD, GA_CTX_SC_RD,
    GA_INDEL_SC_LEFT, GA_INDEL_SC_LEFT_RD, GA_INDEL_SC_RIGHT, GA_INDEL_SC_RIGHT_RD, GA_INDEL_SC_RD,
    GA_CONC, GA_INS, GA_MUNMAPPED_F, GA_MUNMAPPED_R,
    GA_INDEL_I, GA_INDEL_IDIST, GA_INDEL_D_F, GA_INDEL_D_FDIST, GA_INDEL_D_F_RD,
    GA_INDEL_D_R, GA_INDEL_D_RDIST, GA_INDEL_D_R_RD,
    GA_RD_MQ, GA_RD_RD, GA_RD_LOW,
    GA_GC, GA_ACGT,
    GA_COUNT
};
#define GA_PILEUP_COUNT 23

/* one SNV candidate = the record the reference appends at src/GROM.c:11150-11199 */
typedef struct grom_snv_cand {
    int32_t pos;                /* 0-based */
    int32_t base;               /* 0..3 = A,C,G,T (winning alt) */
    double  ratio;              /* (float)snv[b]/(float)total, widened */
    double  pr;                 /* mq table value (PR) */
    double  hez;                /* hez table value */
    int32_t v[GA_PILEUP_COUNT]; /* the 23 pileup ints of the position */
    int32_t reserved;
} grom_snv_cand;

/* one small-insertion candidate = the record the reference appends at src/GROM.c:11400-11443 */
typedef struct grom_ins_cand {
    int32_t pos;                /* 0-based */
    int32_t dist;               /* inserted length of the primary slot (indel_idist) */
    double  pr;                 /* mq table value */
    double  hez;                /* hez table value */
    int32_t conc;               /* concordant-pair count at the position */
    int32_t weight;             /* indel_i clamped to depth * 6 (cdp_indel_i_temp) */
    int32_t rd;                 /* sum of snv + snv_lowmq */
    int32_t sc;                 /* sc_left[pos+1] + sc_right[pos] */
    int32_t other_len;
    int32_t reserved;
    char    seq[56];            /* first-seen inserted bases when dist <= 50 (src/GROM.c:7219-7228), NUL padded */
} grom_ins_cand;

/* one small-deletion scan event: position whose deletion-start slot (kind 0, indel_d_f, src/GROM.c:11454-11563) or
 * deletion-end slot (kind 1, indel_d_r, src/GROM.c:11630-11745) passes the binomial gate.  The host pairs them with the
 * reference's sequential state machine. */
typedef struct grom_del_event {
    int32_t pos;                /* 0-based */
    int32_t kind;               /* 0 = start (indel_d_f), 1 = end (indel_d_r) */
    double  pr;                 /* mq table value */
    double  hez;                /* hez table value */
    int32_t conc;
    int32_t weight;             /* indel_d_f / indel_d_r */
    int32_t rd;                 /* weight/6 + sum of snv + snv_lowmq */
    int32_t sc;                 /* sc_right (start) / sc_left (end) */
    int32_t other_len;
    int32_t rdist;              /* indel_d_rdist (end events) */
} grom_del_event;

/* ---- structural-variant scan (src/GROM.c:11750-13541) ----
 * One gate event = one (position, class) whose cluster passes the binomial gate; classes 0-9 are the breakpoint clusters in the order
 * del_f del_r dup_f dup_r inv_f1 inv_r1 inv_f2 inv_r2 ctx_f ctx_r, 10 / 11 the insertion gates (soft clips + short pairs, left / right). */
enum { GROM_SV_DEL_F = 0, GROM_SV_DEL_R, GROM_SV_DUP_F, GROM_SV_DUP_R, GROM_SV_INV_F1, GROM_SV_INV_R1, GROM_SV_INV_F2, GROM_SV_INV_R2,
       GROM_SV_CTX_F, GROM_SV_CTX_R, GROM_SV_INS_L, GROM_SV_INS_R, GROM_SV_CLASSES };
typedef struct grom_sv_event {
    int32_t pos;                /* 0-based */
    int32_t cls;                /* GROM_SV_* */
    double  binom;              /* mq table value */
    double  hez;                /* hez table value; 2.0 = side evidence ratio above g_max_evidence_ratio (not computed) */
    double  dist;               /* cluster running-mean distance; ctx: signed mate position */
    int32_t weight;             /* cluster weight (ins gates: the `ins` range-add value) */
    int32_t rd, conc;
    int32_t read_start, read_end;
    int32_t other_len;
    int32_t mchr;               /* ctx: mate contig */
    int32_t reserved;           /* inversion classes: sum of the CNV depth (rd_rd + rd_low_mq_rd) over [read_start, read_end + lseq), which the
                                   emission compares between the two breakpoints (src/GROM.c:15921-15934); 0 otherwise */
} grom_sv_event;

/* one side of a breakpoint pair as the reference's *_list_start_* / *_list_end_* arrays hold it */
typedef struct grom_sv_side {
    int32_t pos;                /* -1 = side not found */
    int32_t weight, rd, conc, read_start, read_end, other_len;
    int32_t reserved;           /* the event's depth sum (inversions) */
    double  binom, hez;
} grom_sv_side;
/* entry of cdp_dup_list / cdp_del_list / cdp_inv_f_list / cdp_inv_r_list / cdp_ins_list before the list -> list2 merge */
typedef struct grom_sv_pair {
    grom_sv_side start, end;
    double dist;
} grom_sv_pair;

/* one read-depth CNV call of detect_del_dup (src/GROM.c:19654-19658 / 19988-19992) with its copy number (20071-20224) and the
 * reference's p-value (17163-17190) */
typedef struct grom_cnv_call {
    int64_t start, end;         /* 0-based, as stored by the reference (printed +1) */
    int32_t kind;               /* 0 = deletion, 1 = duplication */
    int32_t reserved;
    double  z;                  /* largest window score in units of the window-length-specific sd */
    double  pvalue;
    double  cn, cn_sd;          /* trimmed-mean copy number and its spread; -1 / 0 when no usable base */
} grom_cnv_call;

#ifdef __cplusplus
}
#endif
#endif
