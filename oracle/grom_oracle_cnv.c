/* oracle/grom_oracle_cnv.c -- TEST INFRASTRUCTURE ONLY (see grom_oracle.h).
 *
 * CPU restatement of the reference's read-depth CNV path:
 *   FASTA pre-pass side lists   N-run blocks, dinucleotide-repeat runs     src/GROM.c:1684-1764, 1870-1871
 *   pre-statistics              mean MAPQ, contig depth mean/sd, repeat     src/GROM.c:16633-16990, 17123-17125
 *                               bias, high-depth 10 kb blocks -> sample blocks
 *   detect_del_dup              sampled GC-stratified depth distributions,   src/GROM.c:18228-20355
 *                               rank -> sd transform, window-length sweep,
 *                               greedy DEL / DUP segmentation, copy number
 *   emission                    bug-compatible erf tail, -V filter, VCF text src/GROM.c:17146-17500
 *   helpers                     bisect_left/right(_double), grom_rand        src/GROM.c:21630-21860, 1185-1203
 *
 * Sequential on purpose: it is the checker.  Pinned against the white-box reference's cnvpre_/cnv_ dumps
 *  (oracle/hooks.h): committed fixture tests/golden/g2_cnv.npz (tests/test_oracle_cnv_golden.py), live runs tests/dev/compare_cnv.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "grom_oracle.h"

/* ---- libc rand() as the reference's binaries see it: glibc TYPE_3 additive feedback generator (r[i] = r[i-3] + r[i-31]),
 * seeded like srandom_r: r[0] = seed, r[i] = 16807 * r[i-1] mod 2^31-1 (Schrage), first 310 outputs discarded. */
typedef struct { int32_t r[34]; int f, b; } glibc_rand;
static void grand_seed(glibc_rand *g, unsigned seed)
{
    int32_t w = (int32_t)seed;
    if (w == 0) w = 1;
    g->r[0] = w;
    for (int i = 1; i < 31; i++) {
        long hi = w / 127773, lo = w % 127773;
        long t = 16807 * lo - 2836 * hi;
        if (t < 0) t += 2147483647;
        w = (int32_t)t;
        g->r[i] = w;
    }
    g->f = 3; g->b = 0;
    for (int i = 0; i < 310; i++) {
        g->r[g->f] = (int32_t)((uint32_t)g->r[g->f] + (uint32_t)g->r[g->b]);
        g->f = (g->f + 1) % 31; g->b = (g->b + 1) % 31;
    }
}
static int grand_next(glibc_rand *g)
{
    uint32_t v = (uint32_t)g->r[g->f] + (uint32_t)g->r[g->b];
    g->r[g->f] = (int32_t)v;
    g->f = (g->f + 1) % 31; g->b = (g->b + 1) % 31;
    return (int)(v >> 1);
}
/* grom_rand, src/GROM.c:1185-1203: decimal digit-by-digit rejection sampler in [0, max) */
static long grom_rand_r(glibc_rand *g, long max)
{
    long val = 0, scale = 1;
    while (scale < max) {
        long t = (grand_next(g) % 10) * scale;
        while (t + val >= max) t = (grand_next(g) % 10) * scale;
        val += t;
        scale *= 10;
    }
    return val;
}

/* ---- the reference's four bisections (src/GROM.c:21630-21860): bracket halving with an early exit when the probe reaches
 * either end of the range, which makes two-element ranges answer from the last element alone.  STRICT selects the "right" flavour. */
#define BISECT_BODY(LESS)                                                                     \
    long lo = s, hi = e, i = s + (e - s) / 2;                                                 \
    for (;;) {                                                                                \
        if (i <= s) return LESS(v, a[s]) ? s : s + 1;                                         \
        if (i >= e - 1) return LESS(v, a[e - 1]) ? e - 1 : e;                                 \
        if (LESS(v, a[i])) { hi = i; i = lo + (i - lo) / 2; if (hi == i) return i + 1; }       \
        else { lo = i; i = i + (hi - i) / 2; if (lo == i) return i + 1; }                     \
    }
#define LESS_EQ(x, y) ((x) <= (y))
#define LESS_ST(x, y) ((x) < (y))
long oracle_bisect_left(const int *a, int v, long s, long e) { BISECT_BODY(LESS_EQ) }
long oracle_bisect_right(const int *a, int v, long s, long e) { BISECT_BODY(LESS_ST) }
long oracle_bisect_left_double(const double *a, double v, long s, long e) { BISECT_BODY(LESS_EQ) }
long oracle_bisect_right_double(const double *a, double v, long s, long e) { BISECT_BODY(LESS_ST) }
/* bisect_left is also called with a double key against an int list (src/GROM.c:18867): the prototype converts it to int */
static long bisect_left_trunc(const int *a, double v, long s, long e) { return oracle_bisect_left(a, (int)v, s, e); }

/* ---- qsort as glibc implements it (top-down merge sort, left run first on ties) with the reference's comparator
 * `*(int*)a - *(int*)b` (src/GROM.c:1105): on int lists that is an ordinary sort, on the double lists of the copy-number step
 * it orders by the LOW 32 bits of each double with wrap-around subtraction, so the exact algorithm matters. */
static int cmp_lowword(const void *a, const void *b) { return (int)((uint32_t)*(const int32_t *)a - (uint32_t)*(const int32_t *)b); }
static void msort_rec(char *b, size_t n, size_t sz, char *tmp)
{
    if (n <= 1) return;
    size_t n1 = n / 2, n2 = n - n1;
    char *b1 = b, *b2 = b + n1 * sz, *t = tmp;
    msort_rec(b1, n1, sz, tmp); msort_rec(b2, n2, sz, tmp);
    while (n1 > 0 && n2 > 0) {
        if (cmp_lowword(b1, b2) <= 0) { memcpy(t, b1, sz); b1 += sz; n1--; }
        else { memcpy(t, b2, sz); b2 += sz; n2--; }
        t += sz;
    }
    if (n1 > 0) memcpy(t, b1, n1 * sz);
    memcpy(b, tmp, (n - n2) * sz);
}
static void ref_qsort(void *base, size_t n, size_t sz)
{
    if (n < 2) return;
    char *tmp = malloc(n * sz);
    msort_rec(base, n, sz, tmp);
    free(tmp);
}

static int in_set(const char *set, char c) { return c != 0 && strchr(set, c) != NULL; }

/* dinucleotide class of (c0, c1): 0..9 = AA AC AG AT CC CG CT GG GT TT unordered, same case only; 10 = none (src/GROM.c:1727-1735) */
static int dinuc_type(char c0, char c1)
{
    static const char U[10][2] = {{'A','A'},{'A','C'},{'A','G'},{'A','T'},{'C','C'},{'C','G'},{'C','T'},{'G','G'},{'G','T'},{'T','T'}};
    for (int k = 0; k < 10; k++) {
        char a = U[k][0], b = U[k][1], la = a | 0x20, lb = b | 0x20;
        if ((a == c0 && b == c1) || (b == c0 && a == c1) || (la == c0 && lb == c1) || (lb == c0 && la == c1)) return k;
    }
    return 10;
}

#define NB ORACLE_CNV_BINS
#define SEG 10          /* g_repeat_segments */

typedef struct { int *v; long n, n_all; } slist;
static void slist_add(slist *l, long cap, int val, glibc_rand *g)
{
    /* reservoir with the reference's two grom_rand draws (src/GROM.c:18385-18398) */
    if (l->n < cap) { l->v[l->n++] = val; l->n_all++; }
    else { if (grom_rand_r(g, l->n_all) == 0) l->v[grom_rand_r(g, l->n)] = val; l->n_all++; }
}
static int cmp_int(const void *a, const void *b) { int x = *(const int *)a, y = *(const int *)b; return (x > y) - (x < y); }

void oracle_cnv_free(oracle_cnv_out *o)
{
    free(o->nb_s); free(o->nb_e); free(o->rep_t); free(o->rep_s); free(o->rep_e); free(o->sb_s); free(o->sb_e);
    free(o->z); free(o->mask); free(o->win_sd); free(o->win_cnt); free(o->mq_mean);
    for (int k = 0; k < 2; k++) { free(o->call_s[k]); free(o->call_e[k]); free(o->call_z[k]); free(o->call_cn[k]); free(o->call_cs[k]); free(o->call_p[k]); }
    memset(o, 0, sizeof(*o));
}

/* greedy segmentation, src/GROM.c:19361-19678 (sign = +1, deletions) and 19702-20010 (sign = -1, duplications) */
static void greedy_scan(int sign, long len, long start, long end_blk, const oracle_cnv_cfg *c, const int *depth, const int *mq, const int *gc,
                        const uint8_t *mask, const double *z, const double *win_sd, double thr[2][NB], long windows[2][NB],
                        long *n_out, long **s_out, long **e_out, double **z_out)
{
    const long Lmin = c->min_win, Lmax = c->max_win, max_gap = Lmax + 500;
    const int q = c->rd_min_mapq;
    long cap = len / Lmin + 1, n = 0;
    long *cs = malloc(cap * sizeof(long)), *ce = malloc(cap * sizeof(long));
    double *cz = malloc(cap * sizeof(double));
    const long end = end_blk - Lmin;
    int mi = 0, last_low = 0, mi_a = 0, mi_b = 0;
#define BEYOND(p, m) (sign > 0 ? (double)depth[p] <= thr[m][gc[p]] : (double)depth[p] >= thr[m][gc[p]])
#define STEP_MI(var, p) do { if (mq[p] >= q) var = 0; else if (depth[p] > 0) var = 1; } while (0)
    int32_t pos = (int32_t)start;
    while (pos < end) {
        int stop = 0;
        if (mq[pos] >= q) { mi = 0; last_low = 0; }
        else if (depth[pos] > 0) { mi = 1; last_low = 1; }
        else mi = last_low;
        if (BEYOND(pos, mi)) {
            long tpos = pos, wlen = 0, cnt = 0, cnt2 = 0, pa;
            double tot = 0;
            int begun = 0;
            long c_start = 0, c_end = 0, last_good = 0;
            double c_z = 0, tz;
            for (pa = pos; pa < pos + Lmin; pa++) {
                wlen++;
                if (mask[pa] == 0) {
                    STEP_MI(mi, pa);
                    if (BEYOND(pa, mi)) cnt2++;
                    else if (2 * cnt2 < wlen) { stop = 1; tpos = pa; break; }
                } else if (2 * cnt2 < wlen) { stop = 1; tpos = pa; break; }
            }
            if (!stop) {
                cnt = Lmin; tot = 0;
                for (long a = pos; a < pos + Lmin; a++) { cnt -= mask[a]; tot += sign * z[a]; }
            }
            if (!stop && cnt > 0 && win_sd[Lmin] > 0 && tot / (cnt * win_sd[Lmin]) >= 3 && (Lmin - cnt) / (double)Lmin <= 2.0) {
                begun = 1; c_start = pos; last_good = pos + Lmin; c_end = pos + Lmin; c_z = tot / (cnt * win_sd[Lmin]);
            }
            if (!stop) {
                for (pa = pos + Lmin; pa < pos + Lmax; pa++) {
                    wlen++;
                    if (pa >= end) { stop = 1; break; }
                    if (mask[pa] == 0) {
                        STEP_MI(mi, pa);
                        tot += sign * z[pa]; cnt++;
                        if (BEYOND(pa, mi)) {
                            cnt2++;
                            if (win_sd[wlen] > 0 && tot / (cnt * win_sd[wlen]) >= 3 && (wlen - cnt) / (double)wlen <= 2.0) {
                                last_good = pa;
                                tz = tot / (cnt * win_sd[wlen]);
                                if (!begun) { begun = 1; c_start = pos; c_end = pa; c_z = tz; }
                                else { c_end = pa; if (tz > c_z) c_z = tz; }
                            }
                        } else if (2 * cnt2 < wlen) { stop = 1; break; }
                    } else if (2 * cnt2 < wlen) { stop = 1; break; }
                }
            }
            if (!stop && begun) {
                pa = pos + Lmax; tot = 0; cnt = 0; mi_b = mi;
                while (pa < len && pa - last_good <= max_gap) {
                    if (pa == pos + Lmax) {
                        for (long pb = pa - Lmax + 1; pb < pa + 1; pb++) {
                            STEP_MI(mi_b, pb);
                            if (mask[pb] == 0 && windows[mi_b][gc[pb]] > 1) { tot += sign * z[pb]; cnt++; }
                        }
                    } else {
                        long pb = pa - Lmax;
                        STEP_MI(mi_b, pb);
                        if (mask[pb] == 0 && windows[mi_b][gc[pb]] > 1) { tot -= sign * z[pb]; cnt--; }
                        STEP_MI(mi, pa);
                        if (mask[pa] == 0 && windows[mi][gc[pa]] > 1) { tot += sign * z[pa]; cnt++; }
                    }
                    if (cnt > 0 && win_sd[Lmax] > 0 && tot / (cnt * win_sd[Lmax]) >= 3 && (Lmax - cnt) / (double)Lmax <= 2.0) {
                        last_good = pa; c_end = pa;
                        tz = tot / (cnt * win_sd[Lmax]);
                        if (tz > c_z) c_z = tz;
                    }
                    pa++;
                }
            }
            if (begun) {
                /* trim the end back to a stretch that is still mostly beyond the threshold */
                pos = (int32_t)c_end;
                while (pos > c_start + Lmin) {
                    STEP_MI(mi, pos);
                    if (!BEYOND(pos, mi)) { pos -= 1; c_end = pos; }
                    else {
                        long c2 = 0, c3 = 0;
                        int halt = 0;
                        pa = c_end; mi_a = mi;
                        while (pa > c_start + Lmin && !halt) {
                            if (mask[pa] == 0) {
                                STEP_MI(mi_a, pa);
                                c3++;
                                if (BEYOND(pa, mi_a)) c2++;
                            }
                            if (c3 == 0 || (c3 > 0 && c2 / (double)c3 < 0.5) || (c_end - pa + 1 - c3) / ((double)c_end - (double)pa + 1.0) > 2.0) {
                                c_end = pa - 1; halt = 1;
                            }
                            pa--;
                        }
                        pos = (int32_t)pa;
                    }
                }
                pos = (int32_t)(c_end + 1);
                if (n < cap) { cs[n] = c_start; ce[n] = c_end; cz[n] = c_z; n++; }
            } else if (stop) pos = (int32_t)tpos;
        }
        pos += 1;
    }
#undef BEYOND
#undef STEP_MI
    *n_out = n; *s_out = cs; *e_out = ce; *z_out = cz;
}

/* copy number of one call, src/GROM.c:20071-20153 */
static void copy_number(long s, long e, const oracle_cnv_cfg *c, const int *depth, const int *mq, const int *gc, const uint8_t *mask,
                        double ave[2][NB], double *buf, double *cn, double *cs)
{
    long n = 0;
    for (long p = s; p < e; p++) {
        if (mask[p]) continue;
        const int m = mq[p] >= c->rd_min_mapq ? 0 : 1;
        if (ave[m][gc[p]] > 0) buf[n++] = (double)depth[p] / ave[m][gc[p]];
    }
    *cn = -1; *cs = 0;
    if (n <= 0) return;
    ref_qsort(buf, n, sizeof(double));
    const long a = (long)(0.1 * n), b = n - a;
    double tot = 0;
    for (long k = a; k < b; k++) tot += buf[k];
    if (b - a <= 0) return;
    *cn = (tot / (b - a)) * c->ploidy;
    double v = 0;
    for (long k = 0; k < n; k++) { const double d = c->ploidy * buf[k] - *cn; v += d * d; }
    *cs = sqrt(v / n);
}

int oracle_cnv_run(const oracle_cnv_cfg *c, const char *fasta, long len, const int *gc, const int *acgt, const int *rd_mq_sum,
                   const int *rd_rd, const int *rd_low, const double *p2s_p, const double *p2s_sd, int n_p2s, oracle_cnv_out *o)
{
    memset(o, 0, sizeof(*o));
    const long M = c->insert_mean, W1 = 2 * M - 1, lo = M - 1, hi = len - W1;
    const long Lmin = c->min_win, Lmax = c->max_win, cap = c->sample_cap;
    const int q = c->rd_min_mapq;
    glibc_rand rng;
    grand_seed(&rng, c->seed);
    if (hi <= lo) return -1;

    /* ---- FASTA side lists, src/GROM.c:1684-1764 */
    o->nb_s = calloc(2 + len / 100, sizeof(long)); o->nb_e = calloc(2 + len / 100, sizeof(long));
    o->rep_t = calloc(2 + len / 20, sizeof(long)); o->rep_s = calloc(2 + len / 20, sizeof(long)); o->rep_e = calloc(2 + len / 20, sizeof(long));
    {
        long run = 1, first_n = 0, ni = 0, r_s = 0, r_e = 0;
        int old_t = 10;
        for (long p = lo; p < hi; p++) {
            const int is_n = in_set("Nn", fasta[p]);
            if (run > 0) {
                if (!is_n) {
                    if (run >= 100) {
                        if (ni == 0 && o->nb_s[0] == 0) o->nb_s[0] = p;
                        else { o->nb_e[ni] = first_n; ni++; o->nb_s[ni] = p; }
                        if (o->nb_s[ni] >= len) o->nb_s[ni] = len - 1;
                    }
                    run = 0;
                } else run++;
            } else if (is_n) { run = 1; first_n = p; }
            const int t = dinuc_type(fasta[p], fasta[p + 1]);
            if (t != old_t || t == 10) {
                if (r_e > 0 && r_e - r_s >= 19) { o->rep_s[o->n_rep] = r_s; o->rep_e[o->n_rep] = r_e + 1; o->rep_t[o->n_rep] = old_t; o->n_rep++; }
                if (t == 10) r_s = r_e = 0; else r_s = r_e = p;
            } else r_e = p;
            old_t = t;
        }
        o->nb_e[ni] = len - 1;
        o->n_nblk = ni + 1;
    }

    /* ---- pre-statistics, src/GROM.c:16637-16990 */
    int *depth = malloc(len * sizeof(int)), *mq = malloc(len * sizeof(int));
    for (long p = 0; p < len; p++) {
        depth[p] = rd_rd[p] + rd_low[p];
        mq[p] = depth[p] > 0 ? rd_mq_sum[p] / depth[p] : rd_mq_sum[p];
    }
    o->mq_mean = mq;
    {
        double s = 0; long n = 0;
        for (long p = lo; p < hi; p++) if (acgt[p] >= 99) { s += depth[p]; n++; }
        if (n > 0) s = s / n;
        o->chr_ave = s;
        double v = 0;
        for (long p = lo; p < hi; p++) if (acgt[p] >= 99) {
            if (depth[p] < 2 * s) v += (depth[p] - s) * (depth[p] - s); else v += s * s;
        }
        o->chr_sd = n > 1 ? sqrt(v / ((double)n - 1.0)) : 0;
    }
    {
        double *rl = malloc((o->n_rep + 1) * sizeof(double));
        for (int k = 0; k < 10; k++) { o->rep_ave[k] = 0; o->rep_sd[k] = 0; o->rep_cnt[k] = 0; }
        for (long i = 0; i < o->n_rep; i++) {
            long s = 0;
            for (long p = o->rep_s[i]; p < o->rep_e[i]; p++) s += depth[p];
            rl[i] = (double)s / (o->rep_e[i] - o->rep_s[i]);
            o->rep_ave[o->rep_t[i]] += rl[i] < 2 * o->chr_ave ? rl[i] : 2 * o->chr_ave;
            o->rep_cnt[o->rep_t[i]]++;
        }
        for (int k = 0; k < 10; k++) o->rep_ave[k] = o->rep_ave[k] / (double)o->rep_cnt[k];
        for (long i = 0; i < o->n_rep; i++) {
            const int t = (int)o->rep_t[i];
            const double x = rl[i] < 2 * o->chr_ave ? rl[i] : 2 * o->chr_ave;
            o->rep_sd[t] += (x - o->rep_ave[t]) * (x - o->rep_ave[t]);
        }
        for (int k = 0; k < 10; k++) o->rep_sd[k] = o->rep_cnt[k] > 1 ? sqrt(o->rep_sd[k] / ((double)o->rep_cnt[k] - 1.0)) : 0;
        free(rl);
        o->biased = -1;
        long best = 0;
        for (int k = 0; k < 10; k++)
            if (o->rep_cnt[k] > 100 && o->rep_ave[k] + 1.5 * o->rep_sd[k] < o->chr_ave && o->chr_ave - 1.5 * o->chr_sd > o->rep_ave[k] && o->rep_cnt[k] > best) {
                o->biased = k; best = o->rep_cnt[k];
            }
    }
    {
        /* 10 kb block means -> blocks above twice the contig mean -> clusters -> complement = sample blocks, src/GROM.c:16784-16990 */
        const long U = 10000, nblk = len / U;
        double *bm = malloc((nblk + 1) * sizeof(double));
        long *over = malloc((nblk + 1) * sizeof(long));
        long tot_acgt = 0, n_acgt = 0, tot = 0, cnt = 0, nb = 0, n_over = 0;
        for (long p = 0; p < len; p++) {
            if (in_set("CGcg", fasta[p]) || in_set("ATat", fasta[p])) { tot_acgt += depth[p]; n_acgt++; }
            tot += depth[p];
            if (++cnt == U) { bm[nb++] = tot / (double)cnt; cnt = 0; tot = 0; }
        }
        o->blk_ave = tot_acgt / (double)n_acgt;
        const double thr = 2 * o->blk_ave;
        for (long k = 0; k < nb; k++) if (bm[k] > thr) over[n_over++] = k;
        long *bs = calloc(10001, sizeof(long)), *be = calloc(10001, sizeof(long));
        long run = 0, r_s = 0, r_e = 0, bi = 0;
        for (long a = 1; a < n_over; a++) {
            if (run == 0) {
                if (run + 1 > (over[a] - over[a - 1]) / 4) { r_e = over[a] + 1; run++; }
                else r_e = over[a - 1] + 1;
                r_s = over[a - 1]; run++;
            } else {
                if (run + 1 > (over[a - 1] - r_s) / 4) { r_e = over[a - 1] + 1; run++; }
                else { if (run >= 4) bi++; r_s = over[a - 1]; r_e = over[a - 1] + 1; run = 1; }
                if (run >= 4 && bi < 10000) { bs[bi] = r_s * U; be[bi] = r_e * U; }
            }
        }
        if (run >= 4) bi++;
        long *ls = calloc(bi + 3, sizeof(long)), *le = calloc(bi + 3, sizeof(long));
        long li = 0;
        for (long a = 0; a < bi && a < 10000; a++) if (be[a] - bs[a] >= 10000) { le[li] = bs[a]; ls[li + 1] = be[a]; li++; }
        li++;
        le[li - 1] = len;
        for (long k = 0; k < li; k++) {
            if (ls[k] < lo) ls[k] = lo; else if (ls[k] >= hi) ls[k] = hi;
            if (le[k] < lo) le[k] = lo; else if (le[k] >= hi) le[k] = hi;
        }
        long w = 0;
        for (long k = 0; k < li; k++) if (le[k] - ls[k] >= Lmin) { ls[w] = ls[k]; le[w] = le[k]; w++; }
        o->n_sblk = w; o->sb_s = ls; o->sb_e = le;
        free(bm); free(over); free(bs); free(be);
    }

    /* ---- detect_del_dup, src/GROM.c:18228-20355 */
    const double del_f = 1.0 - 0.6 / c->ploidy, dup_f = 1.0 + 0.6 / c->ploidy;
    slist high[NB], low[NB], rsl[SEG];
    for (int b = 0; b < NB; b++) { high[b].v = malloc(cap * sizeof(int)); low[b].v = malloc(cap * sizeof(int)); high[b].n = high[b].n_all = low[b].n = low[b].n_all = 0; }
    for (int k = 0; k < SEG; k++) { rsl[k].v = malloc(cap * sizeof(int)); rsl[k].n = rsl[k].n_all = 0; }
    double rs_ave[SEG] = {0}, rs_sd[SEG] = {0};
    const long half = M / 2;
#define REP_SEGMENT(p, i) ((p) < o->rep_s[i] ? (SEG - 1) * ((p) - (o->rep_s[i] - half)) / half : (p) >= o->rep_e[i] ? (SEG - 1) * ((o->rep_e[i] + half) - (p)) / half : SEG - 1)
    if (o->biased != -1) {
        for (long i = 0; i < o->n_rep; i++) {
            if (o->rep_t[i] != o->biased) continue;
            for (int32_t p = (int32_t)(o->rep_s[i] - half); p < o->rep_e[i] + half; p++)
                if (acgt[p] >= 99) slist_add(&rsl[REP_SEGMENT(p, i)], cap, depth[p], &rng);
        }
        for (int k = 0; k < SEG; k++) {
            if (rsl[k].n > 1) qsort(rsl[k].v, rsl[k].n, sizeof(int), cmp_int);
            if (rsl[k].n > 0) {
                const long a = rsl[k].n / 20, b = rsl[k].n - a, n = b - a;
                double s = 0, v = 0;
                for (long j = a; j < b; j++) s += rsl[k].v[j];
                rs_ave[k] = s / n;
                for (long j = a; j < b; j++) v += (rsl[k].v[j] - rs_ave[k]) * (rsl[k].v[j] - rs_ave[k]);
                rs_sd[k] = n > 1 ? sqrt(v / (n - 1)) : v;
            }
        }
    }
    {
        /* sampled depth distributions per GC bin, src/GROM.c:18373-18456: one draw every insert_mean/2 bases of the sample blocks;
         * zero-depth positions go to whichever list the last covered sample went to */
        int last_low = 0;
        for (long k = 0; k < o->n_sblk; k++)
            for (int32_t p = (int32_t)o->sb_s[k]; p < o->sb_e[k]; p += (int32_t)half) {
                if (acgt[p] < 99) continue;
                int to_low;
                if (rd_rd[p] == 0 && rd_low[p] == 0) to_low = last_low;
                else if (mq[p] >= q) to_low = last_low = 0;
                else to_low = last_low = 1;
                slist_add(to_low ? &low[gc[p]] : &high[gc[p]], cap, depth[p], &rng);
            }
    }
    for (int b = 0; b < NB; b++) {
        if (high[b].n > 1) qsort(high[b].v, high[b].n, sizeof(int), cmp_int);
        if (low[b].n > 1) qsort(low[b].v, low[b].n, sizeof(int), cmp_int);
    }
    {
        /* thin bins (20 <= n < 100) borrow the ORIGINAL samples of the two neighbours on each side, src/GROM.c:18481-18548 */
        long nh[NB], nl[NB];
        for (int b = 0; b < NB; b++) { nh[b] = high[b].n; nl[b] = low[b].n; }
        for (int b = 2; b < NB - 2; b++) {
            if (high[b].n >= 20 && high[b].n < 100)
                for (int a = b - 2; a <= b + 2; a++) if (a != b) for (long j = 0; j < high[a].n; j++) if (nh[b] < cap) high[b].v[nh[b]++] = high[a].v[j];
            if (low[b].n >= 20 && low[b].n < 100)
                for (int a = b - 2; a <= b + 2; a++) if (a != b) for (long j = 0; j < low[a].n; j++) if (nl[b] < cap) low[b].v[nl[b]++] = low[a].v[j];
        }
        for (int b = 2; b < NB - 2; b++) {
            if (high[b].n >= 20 && high[b].n < 100) { high[b].n = nh[b]; qsort(high[b].v, high[b].n, sizeof(int), cmp_int); }
            if (low[b].n >= 20 && low[b].n < 100) { low[b].n = nl[b]; qsort(low[b].v, low[b].n, sizeof(int), cmp_int); }
        }
    }
    for (int b = 0; b < NB; b++)
        for (int m = 0; m < 2; m++) {
            const slist *l = m ? &low[b] : &high[b];
            double s = 0, v = 0;
            o->ave[m][b] = o->sd[m][b] = o->del_thr[m][b] = o->dup_thr[m][b] = 0; o->windows[m][b] = 0;
            if (l->n > 0) {
                for (long j = 0; j < l->n; j++) s += l->v[j];
                o->ave[m][b] = s / l->n;
                o->del_thr[m][b] = del_f * o->ave[m][b]; o->dup_thr[m][b] = dup_f * o->ave[m][b];
                o->windows[m][b] = l->n;
                for (long j = 0; j < l->n; j++) v += (l->v[j] - o->ave[m][b]) * (l->v[j] - o->ave[m][b]);
                o->sd[m][b] = l->n > 1 ? sqrt(v / (l->n - 1)) : v;
            }
            if (m) o->n_low[b] = l->n; else o->n_high[b] = l->n;
        }

    /* mask: positions outside the analysed span, with < 99 % ACGT context, or whose bin has < 100 samples, src/GROM.c:18673-18720 */
    uint8_t *mask = malloc(len);
    o->mask = mask;
    memset(mask, 1, len);
    {
        int last_low = 0, mi;
        for (long p = lo; p < hi; p++) {
            if (acgt[p] < 99) continue;
            if (depth[p] == 0) mi = last_low;
            else if (mq[p] >= q) mi = last_low = 0;
            else mi = last_low = 1;
            mask[p] = o->windows[mi][gc[p]] < 100 ? 1 : 0;
        }
    }

    /* rank of the depth within its bin's sample -> probability -> sd units, weighted by mean MAPQ, src/GROM.c:18754-18963 */
    double *z = calloc(len, sizeof(double));
    o->z = z;
#define USABLE(p) (mask[p] == 0 && ((mq[p] >= q && o->windows[0][gc[p]] > 1) || (mq[p] < q && o->windows[1][gc[p]] > 1)))
#define HALF_IF_ZERO(i) ((i) <= 0 ? 0.5 : (double)(i))
    {
        int last_low = 0, mi;
        for (long p = lo; p < hi; p++) {
            if (!USABLE(p)) continue;
            if (mq[p] >= q) mi = last_low = 0;
            else if (depth[p] == 0) mi = last_low;
            else mi = last_low = 1;
            const int b = gc[p];
            const slist *l = mi ? &low[b] : &high[b];
            const long n = l->n;
            if (n <= 0) continue;
            const double w = mq[p] >= q ? 0.5 + (1.0 - 0.5) * (mq[p] - q) / (double)(60 - q) : 0.5;
            long i1, i2;
            double sgn;
            if (depth[p] < o->ave[mi][b]) {
                i1 = oracle_bisect_right(l->v, depth[p], 0, n); i2 = oracle_bisect_left(l->v, depth[p], 0, n); sgn = 1.0;
            } else {
                if (depth[p] > 2 * o->ave[mi][b]) i1 = bisect_left_trunc(l->v, 2 * o->ave[mi][b], 0, n);
                else i1 = oracle_bisect_left(l->v, depth[p], 0, n);
                i2 = oracle_bisect_right(l->v, depth[p], 0, n);
                i1 = n - i1; i2 = n - i2; sgn = -1.0;
            }
            const double prob = (HALF_IF_ZERO(i1) + HALF_IF_ZERO(i2)) / (2 * n);
            long k = oracle_bisect_right_double(p2s_p, prob, 0, n_p2s);
            if (k < 0) k = 0; else if (k >= n_p2s) k = n_p2s - 1;
            z[p] = sgn > 0 ? w * p2s_sd[k] : -w * p2s_sd[k];
        }
    }

    /* window-length sweep, src/GROM.c:18967-19018: the analysed span, re-walked at each -A offset WITHOUT resetting the running
     * frame, is cut into frames of Lmax positions; for each frame and each L >= Lmin the mean z over the usable positions among its
     * first L elements is one observation of the null distribution at length L */
    const long nwin = Lmax + 1;
    long *wcnt = calloc(nwin, sizeof(long));
    double *wsq = calloc(nwin, sizeof(double));      /* running sum of squares in observation order == src/GROM.c:19168-19171 */
    o->win_cnt = wcnt;
    for (long k = 0; k < o->n_sblk; k++) {
        long wlen = 0, n_us = 0, n_mask = 0;
        double tot = 0;
        for (long a = 0; a < c->windows_sampling_factor; a++) {
            const long adj = a * Lmax / c->windows_sampling_factor;
            for (int32_t p = (int32_t)(o->sb_s[k] + adj); p < o->sb_e[k]; p++) {
                if (USABLE(p)) { tot += z[p]; n_us++; }
                n_mask += mask[p];
                wlen++;
                if (wlen >= Lmin && n_mask / (double)wlen < 2.0 && n_us > 0) {
                    const double x = tot / (double)n_us;
                    wsq[wlen] += x * x; wcnt[wlen]++;
                }
                if (wlen == Lmax) { wlen = 0; tot = 0; n_mask = 0; n_us = 0; }
            }
        }
    }

    /* most-biased repeat override, src/GROM.c:19023-19150 */
    if (o->biased != -1) {
        for (long i = 0; i < o->n_rep; i++) {
            if (o->rep_t[i] != o->biased) continue;
            for (int32_t p = (int32_t)(o->rep_s[i] - half); p < o->rep_e[i] + half; p++) {
                if (mask[p] != 0) continue;
                const int sg = (int)REP_SEGMENT(p, i);
                const slist *l = &rsl[sg];
                const long n = l->n;
                long i1, i2;
                double sgn;
                if (depth[p] < rs_ave[sg]) { i1 = oracle_bisect_right(l->v, depth[p], 0, n); i2 = oracle_bisect_left(l->v, depth[p], 0, n); sgn = 1.0; }
                else {
                    if (depth[p] > 2 * rs_ave[sg]) i1 = bisect_left_trunc(l->v, 2 * rs_ave[sg], 0, n);
                    else i1 = oracle_bisect_left(l->v, depth[p], 0, n);
                    i2 = oracle_bisect_right(l->v, depth[p], 0, n);
                    i1 = n - i1; i2 = n - i2; sgn = -1.0;
                }
                const double prob = (HALF_IF_ZERO(i1) + HALF_IF_ZERO(i2)) / (2 * n);
                long k = oracle_bisect_right_double(p2s_p, prob, 0, n_p2s);
                if (k < 0) k = 0; else if (k >= n_p2s) k = n_p2s - 1;
                z[p] = sgn > 0 ? p2s_sd[k] : -p2s_sd[k];
            }
        }
    }

    o->win_sd = calloc(nwin, sizeof(double));
    for (long L = Lmin; L <= Lmax; L++) o->win_sd[L] = wcnt[L] > 1 ? sqrt(wsq[L] / (wcnt[L] - 1)) : 0.0;
    free(wsq);

    /* ---- greedy segmentation over the single analysed block [M-1, len-W1) (src/GROM.c:17123-17125), then copy number */
    greedy_scan(+1, len, lo, hi, c, depth, mq, gc, mask, z, o->win_sd, o->del_thr, o->windows, &o->n_call[0], &o->call_s[0], &o->call_e[0], &o->call_z[0]);
    greedy_scan(-1, len, lo, hi, c, depth, mq, gc, mask, z, o->win_sd, o->dup_thr, o->windows, &o->n_call[1], &o->call_s[1], &o->call_e[1], &o->call_z[1]);
    for (int k = 0; k < 2; k++) {
        long longest = 0;
        for (long i = 0; i < o->n_call[k]; i++) if (o->call_e[k][i] - o->call_s[k][i] > longest) longest = o->call_e[k][i] - o->call_s[k][i];
        double *buf = malloc((longest + 1) * sizeof(double));
        o->call_cn[k] = calloc(o->n_call[k] + 1, sizeof(double)); o->call_cs[k] = calloc(o->n_call[k] + 1, sizeof(double));
        o->call_p[k] = calloc(o->n_call[k] + 1, sizeof(double));
        for (long i = 0; i < o->n_call[k]; i++) {
            copy_number(o->call_s[k][i], o->call_e[k][i], c, depth, mq, gc, mask, o->ave, buf, &o->call_cn[k][i], &o->call_cs[k][i]);
            /* one-sided normal tail with the reference's own erf variant: t = 1/(1 + p + x), src/GROM.c:17163-17172 */
            const double x = fabs(o->call_z[k][i]) / sqrt(2.0), t = 1.0 / (1.0 + 0.3275911 + x);
            const double erf_ = 1.0 - ((0.254829592 * t + -0.284496736 * (t * t) + 1.421413741 * pow(t, 3) + -1.453152027 * pow(t, 4) + 1.061405429 * pow(t, 5)) * exp(-(x * x)));
            o->call_p[k][i] = (1.0 - erf_) / 2.0;
        }
        free(buf);
    }
    for (int b = 0; b < NB; b++) { free(high[b].v); free(low[b].v); }
    for (int k = 0; k < SEG; k++) free(rsl[k].v);
    free(depth);
    return 0;
}

/* VCF text of the calls that pass -V, deletions then duplications, src/GROM.c:17197-17500 */
char *oracle_format_cnv_vcf(const oracle_cnv_cfg *c, const char *chr, const oracle_cnv_out *o)
{
    size_t cap = 256 * (size_t)(o->n_call[0] + o->n_call[1] + 1), n = 0;
    char *s = malloc(cap);
    s[0] = 0;
    for (int k = 0; k < 2; k++)
        for (long i = 0; i < o->n_call[k]; i++) {
            if (!(o->call_p[k][i] < c->rd_pval_threshold)) continue;
            n += snprintf(s + n, cap - n, "%s\t%ld\t.\t.\t<%s>\t.\t.\tEND=%ld\tSD:Z:CN:CS\t%e:%e:%.2f:%e\n", chr, o->call_s[k][i] + 1, k ? "DUP" : "DEL",
                          o->call_e[k][i] + 1, o->call_z[k][i], o->call_p[k][i], o->call_cn[k][i], o->call_cs[k][i]);
        }
    return s;
}
