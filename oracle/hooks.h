/* oracle/hooks.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Dump hooks for the white-box build of the reference (oracle/Makefile).  The
 * reference translation unit is never copied: the Makefile streams
 * /root/reference/src/GROM.c through `sed`, which inserts one-line macro calls
 * at fixed line numbers, straight into gcc's stdin.  The macros below expand,
 * inside the reference's own function scope, to calls that hand the
 * reference's own local arrays to hooks.c.  With GROM_DUMP_DIR unset every
 * hook is a single predictable branch, so the same binary is also the timed
 * CPU baseline.
 *
 * Record layouts written by hooks.c (all little-endian, no padding):
 *   scan_<chr>.bin  : per scanned position  int32 pos, int32 v[GH_NI], double d[GH_ND]
 *   reads_<chr>.bin : per read that reaches the -M test   int32 pos, mpos, tlen, flag, mapq, keep
 *   depth_<chr>.bin : int32 rd_mq[P], rd_rd[P], rd_low_mq_rd[P]   (before the in-place mean, src/GROM.c:16637)
 *   gc_<chr>.bin    : int32 gc_weighted[P], acgt_weighted[P]
 *   svl_<list>_<chr>.bin: candidate lists at the end of the per-position scan (src/GROM.c:15164), see grom_hook_svpairs / _svctx / _svins
 *   cnvpre_<chr>.bin: state after the CNV pre-statistics (src/GROM.c:16633-16990), see grom_hook_cnvpre
 *   cnv_<chr>.bin   : state at the end of detect_del_dup (src/GROM.c:20348), see grom_hook_cnv
 */
#ifndef GROM_ORACLE_HOOKS_H
#define GROM_ORACLE_HOOKS_H

#define GH_NI 86
#define GH_ND 10

extern int g_hook_on;
void grom_hook_scan(const char *chr, int pos, const int *v, const double *d);
void grom_hook_read(const char *chr, int pos, int mpos, int tlen, int flag, int mapq, int keep);
void grom_hook_depth(const char *chr, long len, const int *mq, const int *rd, const int *low);
void grom_hook_gc(const char *chr, long len, const int *gc, const int *acgt);
void grom_hook_srand(unsigned seed);
void grom_hook_svpairs(const char *chr, const char *name, long n, const int *start, const int *end, const double *dist,
                       const double *sb, const double *sh, const int *sc, const int *srd, const int *sw, const int *srs, const int *sre, const int *sol,
                       const double *eb, const double *eh, const int *ec, const int *erd, const int *ew, const int *ers, const int *ere, const int *eol);
void grom_hook_svctx(const char *chr, const char *name, long n, const int *pos, const double *b, const double *h, const int *mchr, const int *mpos,
                     const int *conc, const int *rd, const int *w, const int *rs, const int *re, const int *ol);
void grom_hook_svins(const char *chr, long n, const int *start, const int *end, const double *sb, const double *eb, const int *si, const int *ei,
                     const int *srd, const int *erd, const int *sc, const int *ec, const int *sol, const int *eol);
void grom_hook_cnvpre(const char *chr, long n_nblk, const long *nb_s, const long *nb_e, long n_rep, const int *rep_t, const long *rep_s,
                      const long *rep_e, double chr_ave, double chr_sd, const double *rep_ave, const double *rep_sd, const long *rep_cnt,
                      int biased, double blk_ave, long n_sblk, const long *sb_s, const long *sb_e);
void grom_hook_cnv(const char *chr, long len, const double *z, const int *mask, long nwin, const double *win_sd, const long *win_cnt,
                   long nbins, const double *ave, const double *sd, const double *del_thr, const double *dup_thr, const long *windows,
                   const long *n_high, const long *n_low, long n_del, const long *del_s, const long *del_e, const double *del_z,
                   const double *del_cn, const double *del_cs, long n_dup, const long *dup_s, const long *dup_e, const double *dup_z,
                   const double *dup_cn, const double *dup_cs);

#define GH_CL(k, w, rs, re) v[51 + 3 * (k)] = (w)[ix]; v[52 + 3 * (k)] = (rs)[ix]; v[53 + 3 * (k)] = (re)[ix];

/* inserted immediately before reference src/GROM.c:11086 (per-position scan gate) */
#define GROM_HOOK_SCAN() do { if (g_hook_on && cdp_pos_in_contig_start > 2 * g_insert_max_size) { \
    int v[GH_NI]; double d[GH_ND]; int ix = cdp_one_base_index; int b_, o_; \
    for (b_ = 0; b_ < 4; b_++) { v[b_] = cdp_one_base_snv[b_][ix]; v[4 + b_] = cdp_one_base_snv_lowmq[b_][ix]; \
        v[15 + b_] = cdp_one_base_pos_in_read[b_][ix]; v[19 + b_] = cdp_one_base_fstrand[b_][ix]; } \
    v[8] = cdp_one_base_bq[ix]; v[9] = cdp_one_base_bq_all[ix]; v[10] = cdp_one_base_mq[ix]; v[11] = cdp_one_base_mq_all[ix]; \
    v[12] = cdp_one_base_bq_read_count[ix]; v[13] = cdp_one_base_mq_read_count[ix]; v[14] = cdp_one_base_read_count_all[ix]; \
    v[23] = cdp_one_base_rd[ix]; v[24] = cdp_one_base_sc_left[ix]; v[25] = cdp_one_base_sc_left_rd[ix]; \
    v[26] = cdp_one_base_sc_right[ix]; v[27] = cdp_one_base_sc_right_rd[ix]; v[28] = cdp_one_base_sc_rd[ix]; \
    v[29] = cdp_one_base_ctx_sc_left[ix]; v[30] = cdp_one_base_ctx_sc_left_rd[ix]; v[31] = cdp_one_base_ctx_sc_right[ix]; \
    v[32] = cdp_one_base_ctx_sc_right_rd[ix]; v[33] = cdp_one_base_ctx_sc_rd[ix]; v[34] = cdp_one_base_indel_sc_left[ix]; \
    v[35] = cdp_one_base_indel_sc_left_rd[ix]; v[36] = cdp_one_base_indel_sc_right[ix]; v[37] = cdp_one_base_indel_sc_right_rd[ix]; \
    v[38] = cdp_one_base_indel_sc_rd[ix]; v[39] = cdp_one_base_conc[ix]; v[40] = cdp_one_base_ins[ix]; \
    v[41] = cdp_one_base_munmapped_f[ix]; v[42] = cdp_one_base_munmapped_r[ix]; \
    v[43] = cdp_one_base_indel_i[ix]; v[44] = cdp_one_base_indel_idist[ix]; v[45] = cdp_one_base_indel_d_f[ix]; \
    v[46] = cdp_one_base_indel_d_fdist[ix]; v[47] = cdp_one_base_indel_d_f_rd[ix]; v[48] = cdp_one_base_indel_d_r[ix]; \
    v[49] = cdp_one_base_indel_d_rdist[ix]; v[50] = cdp_one_base_indel_d_r_rd[ix]; \
    GH_CL(0, cdp_one_base_del_f, cdp_one_base_del_f_read_start, cdp_one_base_del_f_read_end) \
    GH_CL(1, cdp_one_base_del_r, cdp_one_base_del_r_read_start, cdp_one_base_del_r_read_end) \
    GH_CL(2, cdp_one_base_dup_f, cdp_one_base_dup_f_read_start, cdp_one_base_dup_f_read_end) \
    GH_CL(3, cdp_one_base_dup_r, cdp_one_base_dup_r_read_start, cdp_one_base_dup_r_read_end) \
    GH_CL(4, cdp_one_base_inv_f1, cdp_one_base_inv_f1_read_start, cdp_one_base_inv_f1_read_end) \
    GH_CL(5, cdp_one_base_inv_r1, cdp_one_base_inv_r1_read_start, cdp_one_base_inv_r1_read_end) \
    GH_CL(6, cdp_one_base_inv_f2, cdp_one_base_inv_f2_read_start, cdp_one_base_inv_f2_read_end) \
    GH_CL(7, cdp_one_base_inv_r2, cdp_one_base_inv_r2_read_start, cdp_one_base_inv_r2_read_end) \
    GH_CL(8, cdp_one_base_ctx_f, cdp_one_base_ctx_f_read_start, cdp_one_base_ctx_f_read_end) \
    GH_CL(9, cdp_one_base_ctx_r, cdp_one_base_ctx_r_read_start, cdp_one_base_ctx_r_read_end) \
    v[81] = cdp_one_base_ctx_f_mchr[ix]; v[82] = cdp_one_base_ctx_r_mchr[ix]; \
    for (o_ = 0; o_ < g_other_len; o_++) if (cdp_one_base_other_type[o_][ix] == OTHER_EMPTY) break; \
    v[83] = o_; v[84] = cdp_lseq; v[85] = cdp_pos; \
    d[0] = cdp_one_base_del_fdist[ix]; d[1] = cdp_one_base_del_rdist[ix]; d[2] = cdp_one_base_dup_fdist[ix]; \
    d[3] = cdp_one_base_dup_rdist[ix]; d[4] = cdp_one_base_inv_f1dist[ix]; d[5] = cdp_one_base_inv_r1dist[ix]; \
    d[6] = cdp_one_base_inv_f2dist[ix]; d[7] = cdp_one_base_inv_r2dist[ix]; d[8] = cdp_one_base_ctx_f_mpos[ix]; \
    d[9] = cdp_one_base_ctx_r_mpos[ix]; \
    grom_hook_scan(cdp_chr_name, cdp_pos_in_contig_start, v, d); } } while (0)

/* inserted immediately before reference src/GROM.c:6605 (after the -M decision) */
#define GROM_HOOK_READ() do { if (g_hook_on) grom_hook_read(cdp_chr_name, cdp_pos, cdp_mpos, cdp_tlen, cdp_flag, cdp_mq, cdp_add_to_list); } while (0)

/* inserted immediately before reference src/GROM.c:16633 (CNV pre-statistics) */
#define GROM_HOOK_DEPTH() do { if (g_hook_on) grom_hook_depth(cdp_chr_name, caf_chr_fasta_len, caf_rd_mq_list, caf_rd_rd_list, caf_rd_low_mq_rd_list); } while (0)

/* inserted immediately before reference src/GROM.c:1883 (end of the FASTA pre-pass) */
#define GROM_HOOK_GC() do { if (g_hook_on) grom_hook_gc(cdp_chr_name, caf_chr_fasta_len, caf_one_base_rd_gc_weighted, caf_one_base_rd_acgt_weighted); } while (0)


/* inserted immediately before reference src/GROM.c:15164 (per-position scan finished, candidate lists complete, before list -> list2) */
#define GH_PAIRS(NAME, L, SW, EW) grom_hook_svpairs(cdp_chr_name, NAME, L##_index, L##_start, L##_end, L##_dist, L##_start_binom_cdf, \
    L##_start_hez_binom_cdf, L##_start_conc, L##_start_rd, L##_start_##SW, L##_start_read_start, L##_start_read_end, L##_start_other_len, \
    L##_end_binom_cdf, L##_end_hez_binom_cdf, L##_end_conc, L##_end_rd, L##_end_##EW, L##_end_read_start, L##_end_read_end, L##_end_other_len)
#define GH_CTX(NAME, L, W) grom_hook_svctx(cdp_chr_name, NAME, L##_index, L, L##_binom_cdf, L##_hez_binom_cdf, L##_mchr, L##_mpos, L##_conc, \
    L##_rd, L##_##W, L##_read_start, L##_read_end, L##_other_len)
#define GROM_HOOK_SVLISTS() do { if (g_hook_on) { \
    GH_PAIRS("dup", cdp_dup_list, dup_r, dup_f); GH_PAIRS("del", cdp_del_list, del_f, del_r); \
    GH_PAIRS("inv_f", cdp_inv_f_list, inv, inv); GH_PAIRS("inv_r", cdp_inv_r_list, inv, inv); \
    GH_CTX("ctx_f", cdp_ctx_f_list, ctx_f); GH_CTX("ctx_r", cdp_ctx_r_list, ctx_r); \
    grom_hook_svins(cdp_chr_name, cdp_ins_list_index + 1, cdp_ins_list_start, cdp_ins_list_end, cdp_ins_list_start_binom_cdf, cdp_ins_list_end_binom_cdf, \
        cdp_ins_list_start_ins, cdp_ins_list_end_ins, cdp_ins_list_start_rd, cdp_ins_list_end_rd, cdp_ins_list_start_conc, cdp_ins_list_end_conc, \
        cdp_ins_list_start_other_len, cdp_ins_list_end_other_len); } } while (0)

/* inserted immediately before reference src/GROM.c:17016 (after the CNV pre-statistics, same block scope) */
#define GROM_HOOK_CNVPRE() do { if (g_hook_on) grom_hook_cnvpre(cdp_chr_name, caf_n_index, caf_n_blocks_start, caf_n_blocks_end, caf_repeat_index, \
    caf_repeat_type_list, caf_repeat_start_list, caf_repeat_end_list, caf_repeat_chr_rd_ave, caf_repeat_chr_rd_stdev, caf_repeat_rd_average, \
    caf_repeat_rd_stdev, caf_repeat_rd_type_count, g_most_biased_repeat, caf_chr_rd_ave, g_lowvar_block_sample_index, \
    g_lowvar_block_sample_start_list, g_lowvar_block_sample_end_list); } while (0)

/* inserted immediately before reference src/GROM.c:20348 (end of detect_del_dup, all locals alive) */
#define GROM_HOOK_CNV() do { if (g_hook_on) grom_hook_cnv(ddd_chr_name, ddd_chr_fasta_len, ddd_stdev_list, ddd_rd_low_acgt_or_windows_list, \
    g_max_rd_window_len + 1, ddd_rd_windows_low_stdev, ddd_rd_windows_count, g_num_gc_bins, &ddd_rd_ave_by_gc[0][0], &ddd_rd_stdev_by_gc[0][0], \
    &ddd_rd_ave_by_gc_del_threshold[0][0], &ddd_rd_ave_by_gc_dup_threshold[0][0], &ddd_rd_windows_by_gc[0][0], ddd_high_mq_index, ddd_low_mq_index, \
    *ddd_del_list_index, ddd_del_list_start, ddd_del_list_end, ddd_del_list_stdev, ddd_del_list_cn, ddd_del_list_cn_stdev, \
    *ddd_dup_list_index, ddd_dup_list_start, ddd_dup_list_end, ddd_dup_list_stdev, ddd_dup_list_cn, ddd_dup_list_cn_stdev); } while (0)

#endif
