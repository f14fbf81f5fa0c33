/* gromgpu.h -- C ABI of the B200 (sm_100a) implementation of GROM's per-chromosome hot path.
 *
 * The reference has no plugin/FFI seam: the whole path is the body of
 *
 *     void count_discordant_pairs(samfile_t *bam, char *bam_name, char *chr_fasta, long chr_fasta_len,
 *             char *chr_name, int chr_name_len, FILE *results, ..., char *results_file_name);
 *
 * (reference src/GROM.c:1432, single call site src/GROM.c:21057, one call per
 * chromosome, reads pulled one at a time through my_samread src/GROM.c:981-992,
 * configuration in ~90 globals src/GROM.c:710-974).  This header is the seam a
 * maintainer would cut there: the host keeps BAM decode, FASTA load, candidate
 * post-processing and VCF text; everything between "reads of one chromosome" and
 * "compacted candidate records + count arrays" runs on the GPU behind these
 * entry points.  Plain pointers and sizes only.  INTEGRATION.md shows the call
 * sequence that replaces the body of count_discordant_pairs.
 *
 * Conventions: every function returns 0 on success and a non-zero code on
 * failure, with text in gromgpu_last_error() (the reference's convention is
 * printf + exit, src/GROM.c:1611-1615; the host keeps that policy).  No CPU
 * fallback exists: without a CUDA device gromgpu_init fails.
 * One handle = one chromosome on one device; handles are independent, so one
 * host thread per GPU can run different chromosomes concurrently (the -P
 * process fan-out of src/GROM.c:549-599 becomes chromosome -> GPU assignment).
 */
#ifndef GROMGPU_H
#define GROMGPU_H
#include <stdint.h>
#include "grom_reads.h"
#include "grom_params.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gromgpu_chr gromgpu_chr;

/* device-side time of each stage of the last gromgpu_chr_run(), from CUDA events on the handle's stream */
typedef struct gromgpu_stats {
    float ms_total;             /* first kernel start -> last kernel end */
    float ms_clear;             /* zero-fill of the scatter-target arrays */
    float ms_dup;               /* -M duplicate flags            (replaces src/GROM.c:6432-6588) */
    float ms_prep;              /* per-read CIGAR summary + clip / depth-range scatter (7067-7181) */
    float ms_index;             /* tile -> first-read index */
    float ms_pileup;            /* pileup + CNV depth            (6605-6671, 6740-7059) */
    float ms_rdscan;            /* range-add prefix scan -> rd   (7176-7181) */
    float ms_snvscan;           /* per-position SNV gate + compaction (11096-11199); 0 when fused into the pileup epilogue */
    float ms_gc;                /* GC / ACGT triangular-window percentages (1766-1859) */
    float ms_sv;                /* CIGAR indel slots, split reads, pair ranges and breakpoint clusters (7187-10953) */
    int32_t launches;           /* kernels launched by the run */
    int32_t reserved;
    int64_t n_reads, n_applied, n_dups, aligned_bases;
    int64_t bytes_reads;        /* read-record bytes the pileup consumed (sum rec(r)) */
    int64_t n_sv_items;         /* order-dependent evidence items emitted (indel slots, split reads, discordant ranges) */
    int64_t n_other_slabs;      /* positions that needed the 50 side slots */
} gromgpu_stats;

typedef struct gromgpu_result {
    int32_t scan_first, scan_last;      /* positions the reference would scan, inclusive; -1/-1 if none */
    int64_t n_snv;                      /* SNV candidates, ascending position */
    const grom_snv_cand *snv;           /* host memory owned by the handle, valid until chr_free / next run */
    double  snv_ave_rd;                 /* mean depth for the SNV emission filter (src/GROM.c:15035-15043) */
    int64_t n_ins;                      /* small-insertion candidates (src/GROM.c:11400-11443), ascending position */
    const grom_ins_cand *ins;           /* host memory owned by the handle */
    int64_t n_del;                      /* small-deletion scan events (src/GROM.c:11454-11745), by position, start before end */
    const grom_del_event *del_ev;       /* feed to gromhost_vcf_smalldel() */
    int64_t n_sv;                       /* structural-variant gate events (src/GROM.c:11750-13541) in scan order */
    const grom_sv_event *sv_ev;         /* feed to gromhost_sv_lists() */
} gromgpu_result;

/* Select the device, upload both 1001x1001 tables (row-major double) and the parameters.
 * Replaces: read_binom_tables consumers src/GROM.c:796-799 and the g_* thresholds. */
int gromgpu_init(int device, const double *hez_tbl, const double *mq_tbl, const grom_params *p);
void gromgpu_shutdown(void);
const char *gromgpu_last_error(void);

/* Optional: run every later handle on a caller-owned CUDA stream (cudaStream_t passed as void*);
 * NULL restores the library's own stream. */
int gromgpu_set_stream(void *cuda_stream);

/* Start a chromosome: uploads the FASTA characters (case preserved; the reference compares through
 * toupper, src/GROM.c:6806, and tests 'N'/'n' literally, src/GROM.c:11113).
 * Replaces the per-call allocation + zeroing of src/GROM.c:1884-1908, 2931-5719. */
int gromgpu_chr_begin(gromgpu_chr **h, int tid, const char *fasta, int64_t len);

/* Several chromosomes in flight (one host thread and one stream each): the upload of one overlaps the kernels and the host
 * stages of the others -- the shape of the per-genome driver, where the reference forks one process per chromosome (-P,
 * src/GROM.c:22340-22398).  gromgpu_chr_begin_on binds the handle to the given stream (cudaStream_t as void*, e.g. one made
 * by gromgpu_stream_create); calls on different handles may then be made concurrently from different threads. */
int gromgpu_stream_create(void **cuda_stream);
void gromgpu_stream_destroy(void *cuda_stream);
int gromgpu_chr_begin_on(gromgpu_chr **h, int tid, const char *fasta, int64_t len, void *cuda_stream);
/* Device memory a handle for a chromosome of this length with this many reads / base slots will hold (admission control
 * for the in-flight set), and what is free on the device right now. */
/* Block until everything queued on the handle's stream (uploads of gromgpu_push_reads included) has completed. */
int gromgpu_chr_sync(gromgpu_chr *h);
int64_t gromgpu_chr_bytes_estimate(int64_t len, int64_t n_reads, int64_t n_base_slots);
int64_t gromgpu_device_free_bytes(void);

/* Forget the pushed reads (device buffers are kept) and, if fasta != NULL, upload new characters of the
 * same length: lets one handle be reused for the next chromosome-sized unit without reallocating. */
int gromgpu_chr_reset(gromgpu_chr *h, const char *fasta);

/* Reuse the handle for ANOTHER chromosome that is no longer than the one it was begun for: new tid, new characters, every
 * per-position array zeroed, all device buffers (and the read-depth state of gromgpu_chr_cnv) kept.  The per-genome driver
 * processes its contigs largest first (src/GROM.c:22318-22336), so each lane begins one handle and rebinds it for the rest:
 * no allocation after the first contig (the reference allocates and frees ~2-3 GB per chromosome, src/GROM.c:1884-1908).
 * Returns 0, 1 if `len` exceeds the handle's capacity (nothing changed: free it and begin a new one), < 0 on error. */
int gromgpu_chr_rebind(gromgpu_chr *h, int tid, const char *fasta, int64_t len);

/* Append reads of this chromosome in BAM order (host pointers; may be called repeatedly with
 * consecutive slices).  Replaces the my_samread pulls at src/GROM.c:5740, 10968, 14861. */
int gromgpu_push_reads(gromgpu_chr *h, const grom_read_batch *b);

/* Run the kernels over everything pushed so far (inputs already resident in HBM). */
int gromgpu_chr_run(gromgpu_chr *h);

/* Copy the compacted results to the host (position-sorted). */
int gromgpu_chr_result(gromgpu_chr *h, gromgpu_result *out);

/* Convenience = run + result. */
int gromgpu_chr_finish(gromgpu_chr *h, gromgpu_result *out);

int gromgpu_chr_stats(const gromgpu_chr *h, gromgpu_stats *out);

/* Parity access to the raw per-position arrays: copies array `ga` (GA_* of grom_params.h),
 * positions [p0, p1), to dst (int32, host). */
int gromgpu_debug_fetch(gromgpu_chr *h, int ga, int32_t *dst, int64_t p0, int64_t p1);
/* Parity access to the breakpoint clusters (class order: del_f del_r dup_f dup_r inv_f1 inv_r1 inv_f2 inv_r2 ctx_f ctx_r):
 * what = 0 weight, 1 read_start, 2 read_end (int32), 3 running-mean distance (double), 4 ctx mate contig (cls 0/1 = ctx_f/ctx_r),
 * 5 other_len (cls ignored).  dst receives positions [p0, p1). */
int gromgpu_debug_fetch_cluster(gromgpu_chr *h, int what, int cls, void *dst, int64_t p0, int64_t p1);
/* read_state per read [i0, i1): 0 = not applied (before W/4+1, UNMAP/DUP flag), 1 = applied, 2 = -M duplicate */
int gromgpu_fetch_read_state(gromgpu_chr *h, uint8_t *dst, int64_t i0, int64_t i1);

/* ---- read-depth CNV path: replaces the pre-statistics at src/GROM.c:16633-16990 and the call
 *   detect_del_dup(chr, begin, len, gc_weighted, acgt_weighted, rd_mq, rd_rd, rd_low_mq_rd, sample lists ..., pval2sd_pval, pval2sd_sd,
 *                  pval2sd_len, &del_index, del lists ..., &dup_index, dup lists ..., ploidy, repeat lists ..., file, chr_name)
 * at src/GROM.c:17133 plus the p-values of 17163-17190.  Call after gromgpu_chr_run (it consumes the CNV depth arrays of that
 * run, which stay untouched).  pval2sd_* are the caller's tables exactly as the reference passes them (gromhost_pval2sd()).
 * ploidy is the reference's caf_ploidy.
 * Host stages of the call (sample lists, run heads of the segmentation, copy numbers) use GROMGPU_HOST_THREADS threads when that
 * environment variable is set (the genome drivers set it to cores / (processes on the node x contigs in flight)), otherwise as many
 * as the process's affinity mask allows (at most 16). */
typedef struct gromgpu_cnv_result {
    int64_t n_calls;                    /* deletions in position order, then duplications */
    const grom_cnv_call *calls;         /* host memory owned by the handle; filter with -V and print via gromhost_vcf_cnv() */
    double  chr_ave, chr_sd;            /* contig depth mean / clamped sd over positions with >= 99 % ACGT context */
    double  blk_ave;                    /* mean depth over A/C/G/T reference bases (10 kb block threshold = 2 x) */
    int32_t biased_repeat;              /* g_most_biased_repeat, -1 = none */
    int32_t n_sample_blocks;
    int64_t n_repeats, n_samples, n_frames;
    const double  *win_sd;              /* [max_rd_window_len + 1] null-distribution sd per window length */
    const int64_t *win_cnt;             /* [max_rd_window_len + 1] observations per window length */
    const double  *bin_ave, *bin_sd, *bin_del_thr, *bin_dup_thr;   /* [2][101]: high-MAPQ list, low-MAPQ list per GC bin */
    const int64_t *bin_n;               /* [2][101] */
    float   ms_device, ms_host, ms_total;
    int32_t launches;                   /* kernels launched by this call */
    int32_t reserved;
    int64_t d2h_bytes;                  /* bytes copied device -> host by this call (packed records, seed tables, samples ...) */
} gromgpu_cnv_result;
int gromgpu_chr_cnv(gromgpu_chr *h, const double *pval2sd_pval, const double *pval2sd_sd, int pval2sd_len, int ploidy, gromgpu_cnv_result *out);
/* Parity access after gromgpu_chr_cnv: what = 0 z list (double), 1 mask (uint8), 2 mean MAPQ (uint8), 3 depth (int32); positions [p0, p1) */
int gromgpu_cnv_fetch(gromgpu_chr *h, int what, void *dst, int64_t p0, int64_t p1);

void gromgpu_chr_free(gromgpu_chr *h);

#ifdef __cplusplus
}
#endif
#endif
