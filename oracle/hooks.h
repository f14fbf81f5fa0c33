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

#endif
