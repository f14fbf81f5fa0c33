/* libstats.c -- library statistics the hot path is parameterised with: median insert size, its lower / upper bounds, median
 * read length and the number of well-mapped bases (find_insert_mean, src/GROM.c:1205-1318, called at 22253-22262).
 *
 * The reference scans the first insert_sample_size (10,000,000) qualifying records of the BAM in file order.  Here the
 * accumulator is fed the per-contig batches the pipeline reads anyway, in contig order (= file order of a sorted BAM).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "gromhost.h"

#define SAMPLE_CAP 10000000L     /* insert_sample_size */
#define MAX_MULT 5               /* g_insert_max_mult */

struct gromhost_libstats {
    int *ins, *len;
    long n, cap;
    int64_t mapped;
    int rd_min_mapq;
    int full;
};

gromhost_libstats *gromhost_libstats_new(int rd_min_mapq)
{
    gromhost_libstats *s = (gromhost_libstats *)calloc(1, sizeof(*s));
    s->rd_min_mapq = rd_min_mapq;
    return s;
}
void gromhost_libstats_free(gromhost_libstats *s) { if (s) { free(s->ins); free(s->len); free(s); } }

static void push(gromhost_libstats *s, int insert, int lseq)
{
    if (s->n == s->cap) { s->cap = s->cap ? s->cap * 2 : 1 << 16; s->ins = realloc(s->ins, (size_t)s->cap * sizeof(int)); s->len = realloc(s->len, (size_t)s->cap * sizeof(int)); }
    s->ins[s->n] = insert; s->len[s->n] = lseq; s->n++;
}

/* returns 1 once the sample is full (further batches are ignored, like the reference's loop condition) */
int gromhost_libstats_add(gromhost_libstats *s, const grom_read_batch *b)
{
    for (int64_t i = 0; i < b->n_reads && !s->full; i++) {
        const int flag = b->flag[i];
        if ((flag & 0x4) || (flag & 0x400)) continue;                                   /* BAM_FUNMAP, BAM_FDUP */
        if (!(flag & 0x1)) push(s, b->l_qseq[i], b->l_qseq[i]);                     /* unpaired: read length stands in */
        else if (!(flag & 0x8) && b->mtid[i] == b->tid && b->pos[i] < b->mpos[i] && (flag & 0x2) && b->tlen[i] > 0)
            push(s, b->tlen[i], b->l_qseq[i]);                                      /* leftmost read of a proper pair */
        if (b->mapq[i] >= s->rd_min_mapq) s->mapped += b->l_qseq[i];
        if (s->n >= SAMPLE_CAP) s->full = 1;
    }
    return s->full;
}

static int cmp_int(const void *a, const void *b) { const int x = *(const int *)a, y = *(const int *)b; return (x > y) - (x < y); }

/* ascending order of up to 10 M insert sizes / read lengths: small non-negative integers, so a counting sort (one histogram pass, one
 * write pass) where the values allow it; qsort otherwise.  Same result either way. */
static void sort_ints(int *a, long n)
{
    int lo = 0, hi = 0;
    for (long i = 0; i < n; i++) { if (i == 0 || a[i] < lo) lo = a[i]; if (i == 0 || a[i] > hi) hi = a[i]; }
    if (n < 4096 || lo < 0 || hi >= (1 << 22)) { qsort(a, (size_t)n, sizeof(int), cmp_int); return; }
    long *cnt = (long *)calloc((size_t)hi + 1, sizeof(long));
    if (!cnt) { qsort(a, (size_t)n, sizeof(int), cmp_int); return; }
    for (long i = 0; i < n; i++) cnt[a[i]]++;
    long w = 0;
    for (int v = lo; v <= hi; v++) for (long k = cnt[v]; k > 0; k--) a[w++] = v;
    free(cnt);
}

int gromhost_libstats_finish(gromhost_libstats *s, int *insert_mean, int *lseq, int *insert_min, int *insert_max, int64_t *mapped_reads)
{
    if (s->n == 0) return -1;
    sort_ints(s->ins, s->n);
    int mean = s->ins[s->n / 2];
    const int max_insert = mean * MAX_MULT;
    long end = 0;
    for (long a = s->n - 1; a >= 0; a--) if (s->ins[a] <= max_insert) { end = a; break; }
    end += 1;
    mean = s->ins[end / 2];
    /* one-sided tail beyond g_insert_num_st_devs = 3 sigma, Abramowitz-Stegun erf as at src/GROM.c:21598-21604 */
    const double xc = 3.0 / sqrt(2), t = 1.0 / (1.0 + 0.3275911 * xc);
    const double erf_ = 1.0 - (0.254829592 * t + -0.284496736 * pow(t, 2) + 1.421413741 * pow(t, 3) + -1.453152027 * pow(t, 4) + 1.061405429 * pow(t, 5)) * exp(-pow(xc, 2));
    const double prob2 = (1.0 - erf_) / 2.0;
    const long lo = (long)(int)(prob2 * end / 2), hi = end - lo;
    *insert_min = s->ins[lo]; *insert_max = s->ins[hi < s->n ? hi : s->n - 1];
    sort_ints(s->len, s->n);
    *lseq = s->len[s->n / 2];
    *insert_mean = mean;
    *mapped_reads = s->mapped;
    return 0;
}
