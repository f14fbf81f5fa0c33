/* gromhost.h -- host-side read supply for the GROM hot path (C, zlib + OpenMP only).
 *
 * Replaces, for the new host, the reference's read supply layer
 * (reference src/GROM.c:82-324 ring buffer + bam_fetch producer, 981-992
 * my_samread, 5743-5824 per-record field/aux extraction): instead of one
 * bam1_t at a time it returns all records of one contig as a packed
 * structure-of-arrays batch (include/grom_reads.h) ready to be handed to
 * gromgpu_push_reads().  BGZF blocks of the contig are inflated in parallel.
 *
 * Also exports the serialiser used by the test/bench tooling to write a batch
 * back out as BAM + BAI, so that the reference binary can be run on exactly
 * the same reads.
 *
 * All functions return 0 on success, non-zero on error with a message in
 * gromhost_last_error().
 */
#ifndef GROMHOST_H
#define GROMHOST_H
#include <stdint.h>
#include "grom_reads.h"
#include "grom_params.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct grom_bam grom_bam;       /* an open BAM (+ optional BAI) */
typedef struct grom_batch grom_batch;   /* owns the arrays a grom_read_batch points to */

const char *gromhost_last_error(void);

/* open / header access (reference: samopen src/GROM.c:20474, bam_index_load 22129) */
int  gromhost_bam_open(const char *path, grom_bam **out);
void gromhost_bam_close(grom_bam *b);
int  gromhost_bam_n_targets(const grom_bam *b);
const char *gromhost_bam_target_name(const grom_bam *b, int tid);
int64_t gromhost_bam_target_len(const grom_bam *b, int tid);
int  gromhost_bam_has_index(const grom_bam *b);
/* records of target `tid` as counted in the index's metadata pseudo-bin (samtools >= 0.1.8 writes it): 0 and the two counts (both 0 for a
 * target the index lists no bins for), or -1 when the index does not count them.  A load measure for assigning contigs to GPUs that follows coverage, not just length. */
int  gromhost_bam_target_reads(const grom_bam *b, int tid, int64_t *mapped, int64_t *unmapped);

/* decode every record of target `tid` (BAM order) into a new batch.
 * keep_names != 0 also fills qname_off/qname_pool.  n_threads <= 0: OpenMP default.
 * The BGZF blocks of the target are inflated block-parallel by the library's own DEFLATE decoder (grom_b200/host/inflate.c; a block it
 * refuses goes to zlib; GROMHOST_INFLATE=zlib sends every block there) and the CRC-32 of every block is verified.  GROMHOST_TRACE=1
 * prints the phase times.  Replaces the record-at-a-time read supply of the reference (samread over bgzf/zlib, src/GROM.c:981-992). */
int  gromhost_bam_read_target(grom_bam *b, int tid, int keep_names, int n_threads, grom_batch **out);

/* The same decode in pieces, for targets whose batch should not exist in host memory at once: every _next() returns a new batch (BAM order
 * continues where the one before ended; each batch carries its own canonical offsets and transport-compact forms, exactly what consecutive
 * gromgpu_push_reads() calls take) holding at least max_reads records -- it stops at the end of the decode window in which the count is
 * reached -- or what is left of the target; max_reads <= 0: no limit.  Returns 0 and a batch, 1 when the target is exhausted (an empty target
 * still yields one empty batch first), -1 on error.  gromhost_bam_read_target() is _open + one _next without limit + _close. */
typedef struct grom_target_iter grom_target_iter;
int  gromhost_bam_iter_open(grom_bam *b, int tid, int keep_names, int n_threads, grom_target_iter **out);
int  gromhost_bam_iter_next(grom_target_iter *it, int64_t max_reads, grom_batch **out);
void gromhost_bam_iter_close(grom_target_iter *it);

/* One raw DEFLATE stream (RFC 1951, e.g. the payload of a BGZF block) of known output size through the library's own decoder, without
 * the zlib second opinion: 0 = well formed and exactly dst_len bytes produced, -1 otherwise.  Test / tooling entry. */
int  gromhost_inflate_raw(const uint8_t *src, int64_t src_len, uint8_t *dst, int64_t dst_len);
/* CRC-32 (gzip polynomial) as the batcher computes it for the BGZF trailers: carry-less-multiply folding where the CPU has it, zlib's
 * crc32 otherwise; always equal to zlib's.  Test / tooling entry. */
uint32_t gromhost_crc32(const uint8_t *p, int64_t n);

/* view of an owned batch; pointers stay valid until gromhost_batch_free() */
void gromhost_batch_view(const grom_batch *bt, grom_read_batch *view);
void gromhost_batch_free(grom_batch *bt);

/* Serialise batches (one per target that has reads; sorted by tid) as a
 * coordinate-sorted BAM plus its .bai.  Every batch must carry qname_off/qname_pool.
 * aux, if not NULL, is per batch: aux_off[b][i]..aux_off[b][i+1] bytes of raw BAM
 * aux data for read i (so SA tags can be planted). */
int  gromhost_bam_write(const char *path, int n_targets, const char *const *names, const int64_t *lens,
                        int n_batches, const grom_read_batch *batches,
                        const uint64_t *const *aux_off, const uint8_t *const *aux_pool, int level);


/* ---- reference FASTA (reference src/GROM.c:1332-1417 index pass, 21011-21045 load of one contig) ----
 * open() maps the file and lists the contigs: name = first word of the header, lower-cased, at most 49 characters (what the reference
 * keeps and matches BAM target names against).  load() writes the characters of contig k (case preserved, line ends removed by the
 * reference's own rule: every line is cut behind its last alphabetic character, the cut being re-evaluated only when the line length
 * changes) and returns their number, or -1 (cap too small: raw_bytes() is an upper bound of the length). */
typedef struct grom_fasta grom_fasta;
int  gromhost_fasta_open(const char *path, grom_fasta **out);
void gromhost_fasta_close(grom_fasta *fa);
int  gromhost_fasta_n(const grom_fasta *fa);
const char *gromhost_fasta_name(const grom_fasta *fa, int k);
int  gromhost_fasta_find(const grom_fasta *fa, const char *name);          /* case-insensitive; -1 = absent */
int64_t gromhost_fasta_raw_bytes(const grom_fasta *fa, int k);
int64_t gromhost_fasta_load(const grom_fasta *fa, int k, char *dst, int64_t cap);

/* ---- statistics tables (reference src/GROM.c:21134-21626, 20705-20748) ----
 * hez / mq: row-major double[1001*1001].  _get() mirrors the reference: load
 * "<dir>/GROM_hez_binom_table_1000.txt" / "<dir>/GROM_mq_binom_table_<max(q,10)>_1000.txt"
 * when present, else compute (and write the file when write_missing != 0).  dir == NULL: compute. */
#define GROM_TABLE_DIM 1001
void   gromhost_tables_compute(int min_mapq, double *hez, double *mq);
int    gromhost_tables_get(const char *dir, int min_mapq, int write_missing, double *hez, double *mq);
void   gromhost_table_paths(const char *dir, int min_mapq, char *hez_path, char *mq_path, int cap);
double gromhost_mq_prob(int min_mapq);
int    gromhost_pval2sd(double *pval, double *sd, int cap);   /* returns the length (1001) */

/* ---- host stages of the scan: emission filters + VCF record text (reference src/GROM.c:15046-15095, 16253-16340,
 * 11475-11745 + 16351-16490).  Each returns the number of bytes written to buf, or -1 if cap is too small.
 * Candidates / events must be in ascending position (events: start before end at equal position). */
int64_t gromhost_vcf_snv(const grom_params *p, const char *chr_name, const char *fasta,
                         const grom_snv_cand *snv, int64_t n, double ave_rd, char *buf, int64_t cap);
int64_t gromhost_vcf_ins(const grom_params *p, const char *chr_name, const char *fasta, int64_t chr_len,
                         const grom_ins_cand *ins, int64_t n, char *buf, int64_t cap);
int64_t gromhost_vcf_smalldel(const grom_params *p, const char *chr_name, const char *fasta, int64_t chr_len,
                              const grom_del_event *ev, int64_t n, char *buf, int64_t cap);
/* read-depth CNV records: -V filter (p-value < rd_pval_threshold) and text, deletions then duplications (src/GROM.c:17197-17500).
 * fasta / chr_len are unused (kept for a uniform signature). */
int64_t gromhost_vcf_cnv(const grom_params *p, const char *chr_name, const char *fasta, int64_t chr_len,
                         const grom_cnv_call *calls, int64_t n, char *buf, int64_t cap);

/* ---- library statistics (grom_b200/host/libstats.c) = find_insert_mean, src/GROM.c:1205-1318: feed the per-contig batches in contig
 * order; finish returns the values the reference caches in <bam>.mean (insert_mean is NOT yet raised to lseq, src/GROM.c:22260). */
typedef struct gromhost_libstats gromhost_libstats;
gromhost_libstats *gromhost_libstats_new(int rd_min_mapq);
int  gromhost_libstats_add(gromhost_libstats *s, const grom_read_batch *b);        /* 1 = sample full (10,000,000 inserts) */
int  gromhost_libstats_finish(gromhost_libstats *s, int *insert_mean, int *lseq, int *insert_min, int *insert_max, int64_t *mapped_reads);
void gromhost_libstats_free(gromhost_libstats *s);
/* the same statistics straight from an open BAM: records in file order, blocks inflated a window at a time by n_threads (<= 0: OpenMP
 * default), core fields only, reading ends with the window in which the 10,000,000-insert sample fills up.  Equals _new / _add over the
 * per-target batches in target order / _finish.  -1 = error (gromhost_last_error), e.g. no usable read. */
int  gromhost_bam_library_stats(grom_bam *b, int rd_min_mapq, int n_threads, int *insert_mean, int *lseq, int *insert_min, int *insert_max, int64_t *mapped_reads);

/* ---- structural-variant candidate lists (grom_b200/host/svlists.c): the state of cdp_dup_list / cdp_del_list / cdp_inv_f_list /
 * cdp_inv_r_list / cdp_ins_list / cdp_ctx_f_list / cdp_ctx_r_list at src/GROM.c:15164, rebuilt from the gate events of
 * gromgpu_chr_result (any order; sorted into scan order here).  Arrays are malloc'ed, release with gromhost_sv_lists_free. */
typedef struct gromhost_sv_lists_t {
    int64_t n_dup, n_del, n_inv_f, n_inv_r, n_ins, n_ctx_f, n_ctx_r;
    grom_sv_pair *dup, *del, *inv_f, *inv_r, *ins;
    grom_sv_event *ctx_f, *ctx_r;
} gromhost_sv_lists_t;
int  gromhost_sv_lists(const grom_params *p, const grom_sv_event *events, int64_t n_events, gromhost_sv_lists_t *out);
void gromhost_sv_lists_free(gromhost_sv_lists_t *l);

/* ---- translocations (grom_b200/host/ctx.c): per-contig candidate merge + filter (src/GROM.c:16098-16246), then the genome-level mate
 * pairing and the <out>.ctx.vcf records (src/GROM.c:22470-22745).  target_names: BAM target names, lower-cased, indexed by tid. */
typedef struct grom_ctx_record {
    int32_t type;               /* 6 = CTX_F, 7 = CTX_R (the reference's g_sv_types index) */
    int32_t chr, pos;           /* contig (tid) and 0-based position */
    int32_t rd, conc, other_len;
    int32_t mchr, mpos;         /* mate contig and signed mate position (sign = mate strand); pairing overwrites mpos with the mate's position */
    int32_t read_start, read_end;
    int32_t mate_id, keep;      /* filled by gromhost_ctx_vcf */
    double  binom, evidence, hez;   /* rounded through "%e" / "%.1f" text like the reference's intermediate file */
} grom_ctx_record;
int64_t gromhost_ctx_contig(const grom_params *p, int tid, const grom_sv_event *ctx_f, int64_t n_f, const grom_sv_event *ctx_r, int64_t n_r,
                            grom_ctx_record *out, int64_t cap);
int64_t gromhost_ctx_vcf(const grom_params *p, const char *const *target_names, int n_targets, grom_ctx_record *rec, int64_t n, char *buf, int64_t cap);

/* Every record of one contig in the reference's output order (src/GROM.c:15046-17500): SNV, <DUP>, <INV>, <INS>, small insertions,
 * small deletions, <DEL>, read-depth <DEL>/<DUP>; includes the list -> list2 merge of the structural-variant candidates
 * (src/GROM.c:15164-16090) and the mutual suppression of small and paired-end deletions (16351-16560).  Inputs are the pieces of
 * gromgpu_result / gromgpu_cnv_result.  Returns bytes written, -1 if buf is too small. */
int64_t gromhost_vcf_contig(const grom_params *p, const char *chr_name, const char *fasta, int64_t chr_len,
                            const grom_snv_cand *snv, int64_t n_snv, double snv_ave_rd, const grom_ins_cand *ins, int64_t n_ins,
                            const grom_del_event *del_ev, int64_t n_del_ev, const grom_sv_event *sv_ev, int64_t n_sv_ev,
                            const grom_cnv_call *cnv, int64_t n_cnv, char *buf, int64_t cap);

/* the header block of the main output file (is_ctx = 0) or of <out>.ctx.vcf (is_ctx = 1) exactly as the reference prints it
 * (src/GROM.c:20517-20565, 22639-22677; ##fileDate unpadded, the four read-depth FORMAT lines without the closing '>'); returns the number of
 * bytes written (no terminating NUL counted) or -1 when cap is too small */
int64_t gromhost_vcf_header(const char *fasta_name, int is_ctx, char *buf, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif
