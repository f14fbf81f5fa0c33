/* oracle/hooks.c -- TEST INFRASTRUCTURE ONLY: sinks for the dump hooks of hooks.h.
 * GROM_DUMP_DIR=<dir> switches dumping on; GROM_SEED=<n> pins the reference's
 * srand(time) (reference src/GROM.c:1584) so CNV reservoir sampling is repeatable. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "hooks.h"

int g_hook_on = -1;
static char g_dir[2048];
static FILE *g_scan, *g_reads;
static char g_scan_chr[256], g_reads_chr[256];

static void hook_init(void) __attribute__((constructor));
static void hook_init(void)
{
    const char *d = getenv("GROM_DUMP_DIR");
    g_hook_on = (d && *d) ? 1 : 0;
    if (g_hook_on) snprintf(g_dir, sizeof(g_dir), "%s", d);
}

static FILE *open_for(const char *kind, const char *chr)
{
    char p[4096];
    snprintf(p, sizeof(p), "%s/%s_%s.bin", g_dir, kind, chr);
    FILE *f = fopen(p, "wb");
    if (!f) { fprintf(stderr, "hooks: cannot create %s\n", p); exit(3); }
    setvbuf(f, NULL, _IOFBF, 1 << 22);
    return f;
}

void grom_hook_scan(const char *chr, int pos, const int *v, const double *d)
{
    if (!g_scan || strcmp(chr, g_scan_chr)) {
        if (g_scan) fclose(g_scan);
        snprintf(g_scan_chr, sizeof(g_scan_chr), "%s", chr);
        g_scan = open_for("scan", chr);
    }
    fwrite(&pos, 4, 1, g_scan); fwrite(v, 4, GH_NI, g_scan); fwrite(d, 8, GH_ND, g_scan);
}

void grom_hook_read(const char *chr, int pos, int mpos, int tlen, int flag, int mapq, int keep)
{
    if (!g_reads || strcmp(chr, g_reads_chr)) {
        if (g_reads) fclose(g_reads);
        snprintf(g_reads_chr, sizeof(g_reads_chr), "%s", chr);
        g_reads = open_for("reads", chr);
    }
    int r[6] = { pos, mpos, tlen, flag, mapq, keep };
    fwrite(r, 4, 6, g_reads);
}

void grom_hook_depth(const char *chr, long len, const int *mq, const int *rd, const int *low)
{
    /* the per-position scan of this contig is complete once the depth arrays are final */
    if (g_scan) { fclose(g_scan); g_scan = NULL; }
    if (g_reads) { fclose(g_reads); g_reads = NULL; }
    FILE *f = open_for("depth", chr);
    fwrite(mq, 4, len, f); fwrite(rd, 4, len, f); fwrite(low, 4, len, f);
    fclose(f);
}

void grom_hook_gc(const char *chr, long len, const int *gc, const int *acgt)
{
    FILE *f = open_for("gc", chr);
    fwrite(gc, 4, len, f); fwrite(acgt, 4, len, f);
    fclose(f);
}

void grom_hook_srand(unsigned seed)
{
    const char *s = getenv("GROM_SEED");
    srand(s ? (unsigned)strtoul(s, NULL, 10) : seed);
}

/* cnvpre_<chr>.bin: int64 n_nblk, (int64 start, end)[n_nblk], int64 n_rep, (int64 type, start, end)[n_rep], double chr_ave, chr_sd,
 * double rep_ave[10], rep_sd[10], int64 rep_cnt[10], int64 biased, double blk_ave, int64 n_sblk, (int64 start, end)[n_sblk] */
void grom_hook_cnvpre(const char *chr, long n_nblk, const long *nb_s, const long *nb_e, long n_rep, const int *rep_t, const long *rep_s,
                      const long *rep_e, double chr_ave, double chr_sd, const double *rep_ave, const double *rep_sd, const long *rep_cnt,
                      int biased, double blk_ave, long n_sblk, const long *sb_s, const long *sb_e)
{
    FILE *f = open_for("cnvpre", chr);
    long i, t;
    fwrite(&n_nblk, 8, 1, f);
    for (i = 0; i < n_nblk; i++) { fwrite(&nb_s[i], 8, 1, f); fwrite(&nb_e[i], 8, 1, f); }
    fwrite(&n_rep, 8, 1, f);
    for (i = 0; i < n_rep; i++) { t = rep_t[i]; fwrite(&t, 8, 1, f); fwrite(&rep_s[i], 8, 1, f); fwrite(&rep_e[i], 8, 1, f); }
    fwrite(&chr_ave, 8, 1, f); fwrite(&chr_sd, 8, 1, f);
    fwrite(rep_ave, 8, 10, f); fwrite(rep_sd, 8, 10, f); fwrite(rep_cnt, 8, 10, f);
    t = biased; fwrite(&t, 8, 1, f); fwrite(&blk_ave, 8, 1, f);
    fwrite(&n_sblk, 8, 1, f);
    for (i = 0; i < n_sblk; i++) { fwrite(&sb_s[i], 8, 1, f); fwrite(&sb_e[i], 8, 1, f); }
    fclose(f);
}

/* cnv_<chr>.bin: int64 len, nwin, nbins, n_del, n_dup; double z[len]; uint8 mask[len]; double win_sd[nwin]; int64 win_cnt[nwin];
 * double ave[2][nbins], sd[2][nbins], del_thr[2][nbins], dup_thr[2][nbins]; int64 windows[2][nbins], n_high[nbins], n_low[nbins];
 * per call (del then dup): int64 start, end; double z, cn, cs */
void grom_hook_cnv(const char *chr, long len, const double *z, const int *mask, long nwin, const double *win_sd, const long *win_cnt,
                   long nbins, const double *ave, const double *sd, const double *del_thr, const double *dup_thr, const long *windows,
                   const long *n_high, const long *n_low, long n_del, const long *del_s, const long *del_e, const double *del_z,
                   const double *del_cn, const double *del_cs, long n_dup, const long *dup_s, const long *dup_e, const double *dup_z,
                   const double *dup_cn, const double *dup_cs)
{
    FILE *f = open_for("cnv", chr);
    long i;
    fwrite(&len, 8, 1, f); fwrite(&nwin, 8, 1, f); fwrite(&nbins, 8, 1, f); fwrite(&n_del, 8, 1, f); fwrite(&n_dup, 8, 1, f);
    fwrite(z, 8, len, f);
    for (i = 0; i < len; i++) fputc(mask[i], f);
    fwrite(win_sd, 8, nwin, f); fwrite(win_cnt, 8, nwin, f);
    fwrite(ave, 8, 2 * nbins, f); fwrite(sd, 8, 2 * nbins, f); fwrite(del_thr, 8, 2 * nbins, f); fwrite(dup_thr, 8, 2 * nbins, f);
    fwrite(windows, 8, 2 * nbins, f); fwrite(n_high, 8, nbins, f); fwrite(n_low, 8, nbins, f);
    for (i = 0; i < n_del; i++) { fwrite(&del_s[i], 8, 1, f); fwrite(&del_e[i], 8, 1, f); fwrite(&del_z[i], 8, 1, f); fwrite(&del_cn[i], 8, 1, f); fwrite(&del_cs[i], 8, 1, f); }
    for (i = 0; i < n_dup; i++) { fwrite(&dup_s[i], 8, 1, f); fwrite(&dup_e[i], 8, 1, f); fwrite(&dup_z[i], 8, 1, f); fwrite(&dup_cn[i], 8, 1, f); fwrite(&dup_cs[i], 8, 1, f); }
    fclose(f);
}

/* svl_<name>_<chr>.bin: int64 n, then column by column (n values each): int32 start, end; double dist, start_binom, start_hez; int32
 * start_conc, start_rd, start_weight, start_read_start, start_read_end, start_other_len; double end_binom, end_hez; int32 end_conc,
 * end_rd, end_weight, end_read_start, end_read_end, end_other_len */
static FILE *open_list(const char *name, const char *chr)
{
    char tag[64];
    snprintf(tag, sizeof(tag), "svl_%s", name);
    return open_for(tag, chr);
}
void grom_hook_svpairs(const char *chr, const char *name, long n, const int *start, const int *end, const double *dist,
                       const double *sb, const double *sh, const int *sc, const int *srd, const int *sw, const int *srs, const int *sre, const int *sol,
                       const double *eb, const double *eh, const int *ec, const int *erd, const int *ew, const int *ers, const int *ere, const int *eol)
{
    FILE *f = open_list(name, chr);
    fwrite(&n, 8, 1, f);
    fwrite(start, 4, n, f); fwrite(end, 4, n, f); fwrite(dist, 8, n, f); fwrite(sb, 8, n, f); fwrite(sh, 8, n, f);
    fwrite(sc, 4, n, f); fwrite(srd, 4, n, f); fwrite(sw, 4, n, f); fwrite(srs, 4, n, f); fwrite(sre, 4, n, f); fwrite(sol, 4, n, f);
    fwrite(eb, 8, n, f); fwrite(eh, 8, n, f);
    fwrite(ec, 4, n, f); fwrite(erd, 4, n, f); fwrite(ew, 4, n, f); fwrite(ers, 4, n, f); fwrite(ere, 4, n, f); fwrite(eol, 4, n, f);
    fclose(f);
}
/* svl_ctx_?_<chr>.bin: int64 n; int32 pos; double binom, hez; int32 mchr, mpos, conc, rd, weight, read_start, read_end, other_len */
void grom_hook_svctx(const char *chr, const char *name, long n, const int *pos, const double *b, const double *h, const int *mchr, const int *mpos,
                     const int *conc, const int *rd, const int *w, const int *rs, const int *re, const int *ol)
{
    FILE *f = open_list(name, chr);
    fwrite(&n, 8, 1, f);
    fwrite(pos, 4, n, f); fwrite(b, 8, n, f); fwrite(h, 8, n, f); fwrite(mchr, 4, n, f); fwrite(mpos, 4, n, f); fwrite(conc, 4, n, f);
    fwrite(rd, 4, n, f); fwrite(w, 4, n, f); fwrite(rs, 4, n, f); fwrite(re, 4, n, f); fwrite(ol, 4, n, f);
    fclose(f);
}
/* svl_ins_<chr>.bin: int64 n; int32 start, end; double start_binom, end_binom; int32 start_ins, end_ins, start_rd, end_rd, start_conc,
 * end_conc, start_other_len, end_other_len */
void grom_hook_svins(const char *chr, long n, const int *start, const int *end, const double *sb, const double *eb, const int *si, const int *ei,
                     const int *srd, const int *erd, const int *sc, const int *ec, const int *sol, const int *eol)
{
    FILE *f = open_list("ins", chr);
    fwrite(&n, 8, 1, f);
    fwrite(start, 4, n, f); fwrite(end, 4, n, f); fwrite(sb, 8, n, f); fwrite(eb, 8, n, f); fwrite(si, 4, n, f); fwrite(ei, 4, n, f);
    fwrite(srd, 4, n, f); fwrite(erd, 4, n, f); fwrite(sc, 4, n, f); fwrite(ec, 4, n, f); fwrite(sol, 4, n, f); fwrite(eol, 4, n, f);
    fclose(f);
}
