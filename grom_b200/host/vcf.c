/* vcf.c -- host side of the scan: emission filters and VCF record text for the candidate records the CUDA library
 * returns.  These stages are sequential over a few thousand candidates per chromosome and stay on the host by design
 * (SURVEY.md 8a, rows a11/a12).
 *
 *   gromhost_vcf_snv        src/GROM.c:15046-15095   (depth filter, genotype string, record)
 *   gromhost_vcf_ins        src/GROM.c:16253-16340   (ratio filter, homopolymer rule, record)
 *   gromhost_vcf_smalldel   src/GROM.c:11475-11745 (start/end pairing state machine) + 16351-16490 (filter, record)
 *
 * Bug-compatible on purpose: the second homopolymer run is measured against the character CODE fasta[x] + 1
 * (src/GROM.c:16284, 16433); the small-deletion loop stops before the last list entry (src/GROM.c:16351); record
 * fields named SRD/ERD print the concordant counts (src/GROM.c:16475).  Not reproduced: small-deletion records are not
 * yet suppressed by overlapping <DEL> calls (src/GROM.c:16356-16393) because the discordant-pair DEL caller is not built.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "gromhost.h"

/* one SNV record (depth filter, genotype string, text); 0 = filtered out.  buf must have 1024 bytes of room */
static int snv_record(const grom_params *p, const char *chr_name, const char *fasta, const grom_snv_cand *c, double ave_rd, char *buf)
{
    char gt[512];
    if (!(c->v[GA_RC_ALL] <= round(p->snv_rd_min_factor * ave_rd) || c->ratio >= p->high_cov_min_snv_ratio)) return 0;
    int cn = (int)round(c->ratio * p->ploidy);
    if (cn == 0) cn = 1;
    for (int k = 0; k < p->ploidy && k < 250; k++) { gt[2 * k] = k < cn ? '1' : '0'; gt[2 * k + 1] = k < p->ploidy - 1 ? '/' : '\0'; }
    const int nb = c->v[GA_SNV_A + c->base];
    return snprintf(buf, 1024,
                    "%s\t%d\t\t%c\t%c\t.\t.\t.\tGT:PR:AF:A:C:G:T:AL:CL:GL:TL:BQ:MQ:PIR:FS\t%s:%e:%e:%d:%d:%d:%d:%d:%d:%d:%d:%.2f:%.2f:%.2f:%.2f\n",
                    chr_name, c->pos + 1, fasta[c->pos], "ACGT"[c->base], gt, c->pr, c->ratio,
                    c->v[GA_SNV_A], c->v[GA_SNV_C], c->v[GA_SNV_G], c->v[GA_SNV_T],
                    c->v[GA_SNVLOW_A], c->v[GA_SNVLOW_C], c->v[GA_SNVLOW_G], c->v[GA_SNVLOW_T],
                    (double)c->v[GA_BQ_ALL] / (double)c->v[GA_RC_ALL], (double)c->v[GA_MQ_ALL] / (double)c->v[GA_RC_ALL],
                    (double)c->v[GA_PIR_A + c->base] / (double)nb, (double)c->v[GA_FS_A + c->base] / (double)nb);
}

#define SNV_PAR_MIN 4096          /* candidates from which the records are formatted by all threads (the text is glibc's printf either way) */

int64_t gromhost_vcf_snv(const grom_params *p, const char *chr_name, const char *fasta,
                         const grom_snv_cand *snv, int64_t n, double ave_rd, char *buf, int64_t cap)
{
    int T = 1;
#ifdef _OPENMP
    int64_t par_min = SNV_PAR_MIN;
    { const char *e = getenv("GROMHOST_SNV_PAR_MIN"); if (e && *e) par_min = atoll(e); }          /* tests: force / forbid the all-thread path */
    if (n >= par_min) { T = omp_get_max_threads(); if (T > 64) T = 64; if (T > n / 64) T = (int)(n / 64); if (T < 1) T = 1; }
#endif
    if (T == 1) {
        int64_t w = 0;
        for (int64_t i = 0; i < n; i++) {
            if (cap - w < 1024) return -1;
            w += snv_record(p, chr_name, fasta, &snv[i], ave_rd, buf + w);
        }
        return w;
    }
    /* a contig of chromosome size has tens of thousands of candidates and printf's %e takes microseconds: contiguous shares of the
     * candidates are formatted into per-thread buffers and laid end to end, which is the same text in the same order */
    char *part[64]; int64_t len[64]; int bad = 0;
    memset(part, 0, sizeof(part)); memset(len, 0, sizeof(len));
#ifdef _OPENMP
    #pragma omp parallel for schedule(static, 1) num_threads(T)
#endif
    for (int k = 0; k < T; k++) {
        const int64_t i0 = n * k / T, i1 = n * (k + 1) / T;
        int64_t c = (i1 - i0) * 192 + 2048, w = 0;
        char *b = (char *)malloc((size_t)c);
        for (int64_t i = i0; i < i1 && b; i++) {
            if (c - w < 1024) { c *= 2; char *b2 = (char *)realloc(b, (size_t)c); if (!b2) { free(b); b = NULL; break; } b = b2; }
            w += snv_record(p, chr_name, fasta, &snv[i], ave_rd, b + w);
        }
        if (!b) {
#ifdef _OPENMP
            #pragma omp atomic write
#endif
            bad = 1;
        }
        part[k] = b; len[k] = w;
    }
    int64_t total = 0;
    for (int k = 0; k < T; k++) total += len[k];
    if (bad || cap - total < 1024) { for (int k = 0; k < T; k++) free(part[k]); return bad ? -2 : -1; }
    int64_t o = 0;
    for (int k = 0; k < T; k++) { memcpy(buf + o, part[k], (size_t)len[k]); o += len[k]; free(part[k]); }
    return total;
}

/* homopolymer length as the reference measures it: run of fasta[left] to the left of `left` (inclusive), and the
 * run of the character code fasta[right] + 1 to the right of right + 1 (sic) */
static int homopolymer(const char *fasta, int64_t len, int64_t left, int64_t right, int del_form)
{
    int hp = 1;
    if (!del_form) {
        char hc = fasta[left];
        for (int k = 1; k < 20; k++) { if (left - k >= 0 && hc == fasta[left - k]) hp++; else break; }
    } else if (fasta[left] - 1 >= 0) {                                   /* src/GROM.c:16397 (always true) */
        char hc = fasta[left - 1];
        for (int k = 1; k < 20; k++) { if (left - k - 1 >= 0 && hc == fasta[left - k - 1]) hp++; else break; }
    }
    int hp2 = 1;
    if (fasta[right] + 1 < len) {
        char hc = (char)(fasta[right] + 1);
        for (int k = 1; k < 20; k++) { if (right + k + 1 < len && hc == fasta[right + k + 1]) hp2++; else break; }
    }
    return hp2 > hp ? hp2 : hp;
}

int64_t gromhost_vcf_ins(const grom_params *p, const char *chr_name, const char *fasta, int64_t chr_len,
                         const grom_ins_cand *ins, int64_t n, char *buf, int64_t cap)
{
    int64_t w = 0;
    for (int64_t i = 0; i < n; i++) {
        const grom_ins_cand *c = &ins[i];
        if (!(c->pr <= p->pval_threshold && (double)c->weight / (double)c->rd > p->min_indel_ratio * (double)p->add_factor)) continue;
        const int hp = homopolymer(fasta, chr_len, c->pos, c->pos, 0);
        if (hp > 10) continue;                                            /* g_max_homopolymer */
        char alt[64];
        if (c->dist <= p->indel_i_seq_len) { memcpy(alt, c->seq, (size_t)c->dist); alt[c->dist] = 0; } else strcpy(alt, "<INS>");
        if (cap - w < 1024) return -1;
        /* END = list_end + 1 with list_end never set (-1); ECO / EOT are never written by the reference (printed as 0) */
        w += snprintf(buf + w, (size_t)(cap - w), "%s\t%d\t.\t.\t%s\t.\t.\tEND=%d\tSPR:SEV:SRD:SCO:ECO:SOT:EOT:SSC:HP\t%e:%.1f:%d:%d:%d:%d:%d:%d:%d\n",
                      chr_name, c->pos + 1, alt, 0, c->pr, (double)c->weight / (double)p->add_factor, c->rd, c->conc, 0, c->other_len, 0, c->sc, hp);
    }
    return w;
}

typedef struct {
    int start, end;
    double s_pr, e_pr;
    int s_conc, e_conc, s_w, e_w, s_rd, e_rd, s_sc, e_sc, s_ol, e_ol;
} delrec;

/* pairing state machine over the small-deletion events in (position, start-before-end) order, src/GROM.c:11475-11745;
 * returns the list (n + 2 entries allocated) and the index the reference's cdp_indel_d_list_index ends on */
static delrec *smalldel_list(const grom_params *p, const grom_del_event *ev, int64_t n, int *idx_out)
{
    delrec *L = (delrec *)calloc((size_t)n + 2, sizeof(delrec));
    for (int64_t i = 0; i < n + 2; i++) L[i].end = -1;
    int idx = -1;
    for (int64_t i = 0; i < n; i++) {
        const grom_del_event *e = &ev[i];
        if (e->kind == 0) {
            int set = 0;
            if (idx == -1) { idx = 0; set = 1; }
            else if (L[idx].start != -1 && L[idx].end != -1) { idx++; set = 1; }
            else if ((e->pos - L[idx].start > p->lseq && L[idx].end == -1) || e->pr < L[idx].s_pr) {
                set = 1;
            }
            if (set) {
                L[idx].start = e->pos; L[idx].s_pr = e->pr; L[idx].s_conc = e->conc;
                if (L[idx].end < L[idx].start) L[idx].end = -1;
                L[idx].s_w = e->weight; L[idx].s_sc = e->sc; L[idx].s_rd = e->rd; L[idx].s_ol = e->other_len;
            }
        } else if (idx >= 0) {
            const int near_ = ((float)e->pos - (float)L[idx].start - e->rdist) < 5;      /* g_indel_d_dist_range */
            if ((near_ && L[idx].start != -1 && L[idx].end != -1) || (near_ && (L[idx].end == -1 || e->pr < L[idx].e_pr))) {
                L[idx].end = e->pos; L[idx].e_pr = e->pr; L[idx].e_conc = e->conc; L[idx].e_w = e->weight; L[idx].e_sc = e->sc;
                L[idx].e_rd = e->rd; L[idx].e_ol = e->other_len;
            }
        }
    }
    *idx_out = idx;
    return L;
}
static int smalldel_passes(const grom_params *p, const delrec *d)
{
    return d->s_pr <= p->pval_threshold && d->e_pr <= p->pval_threshold &&
           (double)d->s_w / (double)d->s_rd > p->min_indel_ratio * (double)p->add_factor &&
           (double)d->e_w / (double)d->e_rd > p->min_indel_ratio * (double)p->add_factor;
}
/* share of either interval covered by the other, as the reference computes it (src/GROM.c:16366-16393, 16517-16544); a = the
 * paired-end deletion [as, ae], b = the small deletion [bs, be]; ae_quirk replaces ae in one numerator (src/GROM.c:16378) */
static void overlap_ratios(int as, int ae, int bs, int be, int ae_quirk, int quirk, double *r_small, double *r_pair)
{
    *r_small = 0; *r_pair = 0;
    if (as >= bs && as <= be) {
        if (ae >= be) { *r_small = (double)(be - as) / (double)(be - bs); *r_pair = (double)(be - as) / (double)(ae - as); }
        else { *r_small = (double)(ae - as) / (double)(be - bs); *r_pair = (double)((quirk ? ae_quirk : ae) - as) / (double)(ae - as); }
    } else if (bs >= as && bs <= ae) {
        if (ae >= be) { *r_small = (double)(be - bs) / (double)(be - bs); *r_pair = (double)(be - bs) / (double)(ae - as); }
        else { *r_small = (double)(ae - bs) / (double)(be - bs); *r_pair = (double)(ae - bs) / (double)(ae - as); }
    }
}
/* emission of the small deletions, src/GROM.c:16351-16490 (the loop stops before the entry the index points at); del2 = the merged
 * paired-end deletion list, which suppresses a small deletion it overlaps by half when its evidence is stronger (NULL: no list) */
static int64_t smalldel_emit(const grom_params *p, const char *chr_name, const char *fasta, int64_t chr_len, const delrec *L, int idx,
                             const grom_sv_pair *del2, int64_t n_del2, char *buf, int64_t cap)
{
    int64_t w = 0;
    const int reach = p->insert_max - 2 * p->lseq;
    for (int a = 0; a < idx; a++) {
        const delrec *d = &L[a];
        if (!smalldel_passes(p, d)) continue;
        int covered = 0;
        for (int64_t b = 0; b < n_del2; b++) {
            const grom_sv_pair *q = &del2[b];
            if (!(abs(q->start.pos - d->start) < reach && abs(q->end.pos - d->end) < reach)) continue;
            double r1, r2;
            overlap_ratios(q->start.pos, q->end.pos, d->start, d->end, a < n_del2 ? del2[a].end.pos : -1, 1, &r1, &r2);
            if (r1 >= 0.5 && r2 >= 0.5 && q->start.binom * q->end.binom < d->s_pr * d->e_pr) covered = 1;
        }
        if (covered) continue;
        if (d->start < 0 || d->end < 0 || d->end >= chr_len) continue;
        const int hp = homopolymer(fasta, chr_len, d->start, d->end, 1);
        if (hp > 10) continue;
        const int cn = d->end - d->start + 1;
        if (cap - w < 2048) return -1;
        char ref[128];
        if (cn > 0 && cn < 99) {
            memcpy(ref, fasta + d->start, (size_t)cn); ref[cn] = 0;
            w += snprintf(buf + w, (size_t)(cap - w), "%s\t%d\t.\t%s\t.\t.\t.\tEND=%d\tSPR:EPR:SEV:EEV:SRD:ERD:SCO:ECO:SOT:EOT:SSC:ESC:HP\t%e:%e:%.1f:%.1f:%d:%d:%d:%d:%d:%d:%d:%d:%d\n",
                          chr_name, d->start + 1, ref, d->end + 1, d->s_pr, d->e_pr, (double)d->s_w / (double)p->add_factor, (double)d->e_w / (double)p->add_factor,
                          d->s_conc, d->e_conc, d->s_ol, d->e_ol, d->s_rd, d->e_rd, d->s_sc, d->e_sc, hp);
        } else {
            w += snprintf(buf + w, (size_t)(cap - w), "%s\t%d\t.\t.\t<DEL>\t.\t.\tEND=%d\tSPR:EPR:SEV:EEV:SRD:ERD:SCO:ECO:SOT:EOT:SSC:ESC:HP\t%e:%e:%.1f:%.1f:%d:%d:%d:%d:%d:%d:%d:%d:%d\n",
                          chr_name, d->start + 1, d->end + 1, d->s_pr, d->e_pr, (double)d->s_w / (double)p->add_factor, (double)d->e_w / (double)p->add_factor,
                          d->s_conc, d->e_conc, d->s_ol, d->e_ol, d->s_rd, d->e_rd, d->s_sc, d->e_sc, hp);
        }
    }
    return w;
}

int64_t gromhost_vcf_smalldel(const grom_params *p, const char *chr_name, const char *fasta, int64_t chr_len,
                              const grom_del_event *ev, int64_t n, char *buf, int64_t cap)
{
    int idx;
    delrec *L = smalldel_list(p, ev, n, &idx);
    const int64_t w = smalldel_emit(p, chr_name, fasta, chr_len, L, idx, NULL, 0, buf, cap);
    free(L);
    return w;
}

/* ---- structural variants: list -> list2 merge and records, src/GROM.c:15164-16570 ---------------------------------------------- */

/* Candidates whose start gates lie within ins_max - 2 lseq of each other describe one event: keep the one whose worse side has the
 * smaller p-value, then the larger weights; exact ties average the positions (src/GROM.c:15172-15329, identical for all four lists) */
static int64_t merge_pairs(const grom_params *p, const grom_sv_pair *l, int64_t n, grom_sv_pair **out)
{
    grom_sv_pair *m = (grom_sv_pair *)calloc((size_t)n + 1, sizeof(grom_sv_pair));
    int64_t n2 = 0;
    int begun = 0, first_s = 0, last_s = 0, first_e = 0, last_e = 0;
    double first_d = 0, last_d = 0;
    const int reach = p->insert_max - 2 * p->lseq;
    for (int64_t a = 0; a < n; a++) {
        const grom_sv_pair *c = &l[a];
        if (begun) {
            grom_sv_pair *k = &m[n2 - 1];
            if (c->start.pos > last_s + reach) { begun = 0; first_s = last_s = first_e = last_e = 0; first_d = last_d = 0; }
            else {
                const double worst_c = c->end.binom > c->start.binom ? c->end.binom : c->start.binom;
                const double worst_k = k->end.binom > k->start.binom ? k->end.binom : k->start.binom;
                if (worst_c <= worst_k && c->start.pos >= 0 && c->end.pos >= 0 && k->start.weight <= c->start.weight && k->end.weight <= c->end.weight) {
                    int take = 1;
                    if (c->start.binom == k->start.binom && c->end.binom == k->end.binom) {
                        take = (k->start.weight < c->start.weight && k->end.weight <= c->end.weight) || (k->start.weight <= c->start.weight && k->end.weight < c->end.weight);
                        if (!take && k->start.weight == c->start.weight && k->end.weight == c->end.weight) {
                            last_s = c->start.pos; last_e = c->end.pos; last_d = c->dist;
                            const int ks = (first_s + last_s) / 2, ke = (first_e + last_e) / 2;
                            *k = *c;
                            k->start.pos = ks; k->end.pos = ke; k->dist = (first_d + last_d) / 2.0;
                        }
                    }
                    if (take) { first_s = last_s = c->start.pos; first_e = last_e = c->end.pos; first_d = last_d = c->dist; *k = *c; }
                }
            }
        }
        if (!begun && c->start.pos >= 0 && c->end.pos >= 0 && n2 < 100000 - 1) {           /* g_sv_list2_len */
            begun = 1; first_s = last_s = c->start.pos; first_e = last_e = c->end.pos; first_d = last_d = c->dist;
            m[n2++] = *c;
        }
    }
    *out = m;
    return n2;
}

static int pair_passes(const grom_params *p, const grom_sv_pair *q, int with_hez)
{
    const double t = p->pval_threshold, r = p->min_sv_ratio * (double)p->add_factor;
    return (q->start.binom <= t || (with_hez && q->start.hez <= t)) && (q->end.binom <= t || (with_hez && q->end.hez <= t)) &&
           (double)q->start.weight / (double)q->start.rd >= r && (double)q->end.weight / (double)q->end.rd >= r;
}
static int64_t pair_record(const grom_params *p, const char *chr_name, const char *alt, const grom_sv_pair *q, char *buf, int64_t cap)
{
    if (cap < 1024) return -1;
    return snprintf(buf, (size_t)cap, "%s\t%d\t.\t.\t<%s>\t.\t.\tEND=%d\tSPR:EPR:SEV:EEV:SRD:ERD:SCO:ECO:SOT:EOT:SFR:SLR:EFR:ELR\t%e:%e:%.1f:%.1f:%d:%d:%d:%d:%d:%d:%d:%d:%d:%d\n",
                    chr_name, q->start.pos + 1, alt, q->end.pos + 1, q->start.binom, q->end.binom, (double)q->start.weight / (double)p->add_factor,
                    (double)q->end.weight / (double)p->add_factor, q->start.rd, q->end.rd, q->start.conc, q->end.conc, q->start.other_len, q->end.other_len,
                    q->start.read_start + 1, q->start.read_end + 1, q->end.read_start + 1, q->end.read_end + 1);
}
/* inversions: a candidate of one orientation yields to an overlapping one of the other orientation with stronger evidence, and the
 * depth around its two breakpoints must agree within g_max_inv_rd_diff (src/GROM.c:15897-16007); side.reserved carries the depth
 * sum over [read_start, read_end + lseq) that the CUDA library attaches to inversion gate events */
static int64_t inv_records(const grom_params *p, const char *chr_name, const grom_sv_pair *mine, int64_t n, const grom_sv_pair *other, int64_t n_other,
                           int other_wins_ties, char *buf, int64_t cap)
{
    int64_t w = 0;
    const int reach = p->insert_max - 2 * p->lseq;
    for (int64_t a = 0; a < n; a++) {
        const grom_sv_pair *q = &mine[a];
        if (!pair_passes(p, q, 0)) continue;
        int covered = 0;
        for (int64_t b = 0; b < n_other && !covered; b++) {
            const grom_sv_pair *o = &other[b];
            if (!(abs(q->start.pos - o->start.pos) < reach && abs(q->end.pos - o->end.pos) < reach)) continue;
            if (!((q->start.pos >= o->start.pos && q->start.pos <= o->end.pos) || (o->start.pos >= q->start.pos && o->start.pos <= q->end.pos))) continue;
            const double po = o->start.binom * o->end.binom, pq = q->start.binom * q->end.binom;
            if (other_wins_ties ? po <= pq : po < pq) covered = 1;
        }
        const double d1 = (double)q->start.reserved / (q->start.read_end + p->lseq - q->start.read_start);
        const double d2 = (double)q->end.reserved / (q->end.read_end + p->lseq - q->end.read_start);
        if (covered || !(d1 / d2 <= 1.75 && d2 / d1 <= 1.75)) continue;                     /* g_max_inv_rd_diff */
        const int64_t k = pair_record(p, chr_name, "INV", q, buf + w, cap - w);
        if (k < 0) return -1;
        w += k;
    }
    return w;
}
/* insertions: groups within ins_max - 2 lseq keep the entry whose both sides are at least as significant (src/GROM.c:16013-16090) */
static int64_t ins_records(const grom_params *p, const char *chr_name, const grom_sv_pair *l, int64_t n, char *buf, int64_t cap)
{
    grom_sv_pair *m = (grom_sv_pair *)calloc((size_t)n + 1, sizeof(grom_sv_pair));
    int64_t n2 = 0, w = 0;
    int begun = 0;
    const int reach = p->insert_max - 2 * p->lseq;
    for (int64_t a = 0; a + 1 < n; a++) {                                                  /* the reference stops before its last entry */
        const grom_sv_pair *c = &l[a];
        if (begun) {
            grom_sv_pair *k = &m[n2 - 1];
            if (c->start.pos > k->start.pos + reach || c->start.pos > k->end.pos + reach || c->end.pos > k->start.pos + reach || c->end.pos > k->end.pos + reach) begun = 0;
            else if (c->start.binom <= k->start.binom && c->start.pos >= 0 && c->end.binom <= k->end.binom && c->end.pos >= 0) *k = *c;
        }
        if (!begun && c->start.pos >= 0 && c->end.pos >= 0 && n - 1 < 100000 - 1) { begun = 1; m[n2++] = *c; }
    }
    for (int64_t a = 0; a < n2; a++) {
        const grom_sv_pair *q = &m[a];
        if (!(q->start.binom <= p->pval_insertion && q->end.binom <= p->pval_insertion && abs(q->end.pos - q->start.pos) <= 10)) continue;   /* g_max_ins_range */
        if (cap - w < 1024) { free(m); return -1; }
        w += snprintf(buf + w, (size_t)(cap - w), "%s\t%d\t.\t.\t<INS>\t.\t.\tEND=%d\tSPR:EPR:SEV:EEV:SRD:ERD:SCO:ECO:SOT:EOT\t%e:%e:%.1f:%.1f:%d:%d:%d:%d:%d:%d\n", chr_name,
                      q->start.pos + 1, q->start.pos + 1, q->start.binom, q->end.binom, (double)q->start.weight / (double)p->add_factor,
                      (double)q->end.weight / (double)p->add_factor, q->start.rd, q->end.rd, q->start.conc, q->end.conc, q->start.other_len, q->end.other_len);
    }
    free(m);
    return w;
}

/* All records of one contig in the reference's order (src/GROM.c:15046-17500): SNV, <DUP>, <INV> (forward then reverse orientation),
 * <INS>, small insertions, small deletions (minus those a stronger paired-end deletion covers), <DEL>, read-depth <DEL> / <DUP>.
 * Translocation candidates go to the .ctx.vcf path, which is outside this library. */
int64_t gromhost_vcf_contig(const grom_params *p, const char *chr_name, const char *fasta, int64_t chr_len,
                            const grom_snv_cand *snv, int64_t n_snv, double snv_ave_rd, const grom_ins_cand *ins, int64_t n_ins,
                            const grom_del_event *del_ev, int64_t n_del_ev, const grom_sv_event *sv_ev, int64_t n_sv_ev,
                            const grom_cnv_call *cnv, int64_t n_cnv, char *buf, int64_t cap)
{
    int64_t w = 0, k;
#define ADD(call) do { k = (call); if (k < 0) { w = -1; goto done; } w += k; } while (0)
    /* GROMHOST_TRACE=1: stage times on stderr */
    const int trace = getenv("GROMHOST_TRACE") != NULL;
    struct timespec ts0; if (trace) clock_gettime(CLOCK_MONOTONIC, &ts0);
#define MARK(what) do { if (trace) { struct timespec t1_; clock_gettime(CLOCK_MONOTONIC, &t1_); fprintf(stderr, "[vcf] %-22s %8.2f ms\n", what, (t1_.tv_sec - ts0.tv_sec) * 1e3 + (t1_.tv_nsec - ts0.tv_nsec) * 1e-6); ts0 = t1_; } } while (0)
    gromhost_sv_lists_t L;
    grom_sv_pair *dup2 = NULL, *del2 = NULL, *invf2 = NULL, *invr2 = NULL;
    delrec *small = NULL;
    if (gromhost_sv_lists(p, sv_ev, n_sv_ev, &L) != 0) return -2;
    MARK("sv lists");
    const int64_t n_dup2 = merge_pairs(p, L.dup, L.n_dup, &dup2), n_del2 = merge_pairs(p, L.del, L.n_del, &del2);
    const int64_t n_invf2 = merge_pairs(p, L.inv_f, L.n_inv_f, &invf2), n_invr2 = merge_pairs(p, L.inv_r, L.n_inv_r, &invr2);
    MARK("merge pairs");
    ADD(gromhost_vcf_snv(p, chr_name, fasta, snv, n_snv, snv_ave_rd, buf + w, cap - w));
    MARK("snv records");
    for (int64_t a = 0; a < n_dup2; a++) if (pair_passes(p, &dup2[a], 1)) ADD(pair_record(p, chr_name, "DUP", &dup2[a], buf + w, cap - w));
    ADD(inv_records(p, chr_name, invf2, n_invf2, invr2, n_invr2, 0, buf + w, cap - w));
    ADD(inv_records(p, chr_name, invr2, n_invr2, invf2, n_invf2, 1, buf + w, cap - w));
    ADD(ins_records(p, chr_name, L.ins, L.n_ins, buf + w, cap - w));
    MARK("dup inv ins records");
    ADD(gromhost_vcf_ins(p, chr_name, fasta, chr_len, ins, n_ins, buf + w, cap - w));
    MARK("small insertions");
    {
        int idx;
        small = smalldel_list(p, del_ev, n_del_ev, &idx);
        ADD(smalldel_emit(p, chr_name, fasta, chr_len, small, idx, del2, n_del2, buf + w, cap - w));
        /* paired-end deletions, unless a small deletion with at least as strong evidence covers them (src/GROM.c:16497-16560) */
        const int reach = p->insert_max - 2 * p->lseq;
        for (int64_t a = 0; a < n_del2; a++) {
            const grom_sv_pair *q = &del2[a];
            if (!pair_passes(p, q, 1)) continue;
            int covered = 0;
            for (int b = 0; b < idx && !covered; b++) {
                const delrec *d = &small[b];
                if (!(smalldel_passes(p, d) && abs(q->start.pos - d->start) < reach && abs(q->end.pos - d->end) < reach)) continue;
                double r1, r2;
                overlap_ratios(q->start.pos, q->end.pos, d->start, d->end, 0, 0, &r1, &r2);
                if (r1 >= 0.5 && r2 >= 0.5 && d->s_pr * d->e_pr <= q->start.binom * q->end.binom) covered = 1;
            }
            if (!covered) ADD(pair_record(p, chr_name, "DEL", q, buf + w, cap - w));
        }
    }
    MARK("deletions");
    ADD(gromhost_vcf_cnv(p, chr_name, fasta, chr_len, cnv, n_cnv, buf + w, cap - w));
    MARK("read-depth calls");
done:
    free(dup2); free(del2); free(invf2); free(invr2); free(small);
    gromhost_sv_lists_free(&L);
    return w;
#undef ADD
#undef MARK
}

/* read-depth CNV calls: emission filter and text of src/GROM.c:17197-17240, 17280, 17414 */
int64_t gromhost_vcf_cnv(const grom_params *p, const char *chr_name, const char *fasta, int64_t chr_len,
                         const grom_cnv_call *calls, int64_t n, char *buf, int64_t cap)
{
    int64_t w = 0;
    (void)fasta; (void)chr_len;
    for (int kind = 0; kind < 2; kind++)
        for (int64_t i = 0; i < n; i++) {
            const grom_cnv_call *c = &calls[i];
            if (c->kind != kind || !(c->pvalue < p->rd_pval_threshold)) continue;
            char line[512];
            int m = snprintf(line, sizeof(line), "%s\t%ld\t.\t.\t<%s>\t.\t.\tEND=%ld\tSD:Z:CN:CS\t%e:%e:%.2f:%e\n", chr_name, (long)c->start + 1,
                             kind ? "DUP" : "DEL", (long)c->end + 1, c->z, c->pvalue, c->cn, c->cn_sd);
            if (w + m > cap) return -1;
            memcpy(buf + w, line, m); w += m;
        }
    return w;
}

/* ---- header blocks of <out> and <out>.ctx.vcf as the reference prints them (src/GROM.c:20517-20565, 22639-22677) ---- */
static const char *HDR_COMMON1[] = {
    "##ALT=<ID=DEL,Description=\"Deletion\">", "##ALT=<ID=DUP,Description=\"Duplication\">", "##ALT=<ID=INS,Description=\"Insertion\">",
    "##ALT=<ID=INV,Description=\"Inversion\">", "##INFO=<ID=END,Number=1,Type=Integer,Description=\"End position of the structural variant\">", NULL };
/* FORMAT keys shared by both files: id, type, description */
static const char *HDR_FMT[][3] = {
    {"SPR", "Float", "Probability of start breakpoint evidence occurring by chance"}, {"EPR", "Float", "Probability of end breakpoint evidence occurring by chance"},
    {"SEV", "Integer", "Evidence supporting variant at start breakpoint"}, {"EEV", "Integer", "Evidence supporting variant at end breakpoint"},
    {"SRD", "Integer", "Physical read depth at start breakpoint"}, {"ERD", "Integer", "Physical read depth at end breakpoint"},
    {"SCO", "Integer", "Concordant pairs at start breakpoint"}, {"ECO", "Integer", "Concordant pairs at end breakpoint"},
    {"SOT", "Integer", "Count of distinct SVs with evidence at start breakpoint"}, {"EOT", "Integer", "Count of distinct SVs with evidence at end breakpoint"},
    {"SSC", "Integer", "Soft-clipped reads at start breakpoint"}, {"ESC", "Integer", "Soft-clipped at end breakpoint"},
    {"SFR", "Integer", "Position of first read supporting start breakpoint"}, {"SLR", "Integer", "Position of last read supporting start breakpoint"},
    {"EFR", "Integer", "Position of first read supporting end breakpoint"}, {"ELR", "Integer", "Position of last read supporting end breakpoint"},
    {"AF", "Float", "Allele frequency (high mapping quality reads)"}, {"PR", "Float", "Probability of SNV evidence occurring by chance"},
    {"A", "Integer", "A nucleotides (high mapping quality reads)"}, {"C", "Integer", "C nucleotides (high mapping quality reads)"},
    {"G", "Integer", "G nucleotides (high mapping quality reads)"}, {"T", "Integer", "T nucleotides (high mapping quality reads)"},
    {"AL", "Integer", "A nucleotides (low mapping quality reads)"}, {"CL", "Integer", "C nucleotides (low mapping quality reads)"},
    {"GL", "Integer", "G nucleotides (low mapping quality reads)"}, {"TL", "Integer", "T nucleotides (low mapping quality reads)"},
    {"BQ", "Float", "Average base quality (all reads)"}, {"MQ", "Float", "Average mapping quality (all reads)"},
    {"PIR", "Float", "Average distance of SNV from DNA fragment end)"}, {"FS", "Integer", "SNV reads mapped to forward strand)"}, {NULL, NULL, NULL} };
/* the four read-depth keys of the main file are printed without the closing '>' by the reference; kept byte for byte */
static const char *HDR_CNV[][2] = { {"SD", "CNV standard deviation"}, {"Z", "CNV probability score"}, {"CN", "CNV copy number"}, {"CS", "CNV copy number standard deviation"}, {NULL, NULL} };


int64_t gromhost_vcf_header(const char *fasta_name, int is_ctx, char *buf, int64_t cap)
{
    int64_t w = 0;
#define HP(...) do { const int m_ = snprintf(buf + w, (size_t)(cap - w), __VA_ARGS__); if (m_ < 0 || w + m_ >= cap) return -1; w += m_; } while (0)
    time_t t = time(NULL);
    struct tm tm = *localtime(&t);
    HP("##fileformat=VCFv4.2\n");
    HP("##fileDate=%d%d%d\n", tm.tm_year + 1900, tm.tm_mon + 1, tm.tm_mday);        /* unpadded, like the reference */
    HP("##reference=%s\n", fasta_name);
    for (int i = 0; HDR_COMMON1[i]; i++) HP("%s\n", HDR_COMMON1[i]);
    if (!is_ctx) HP("##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n");
    for (int i = 0; HDR_FMT[i][0]; i++) HP("##FORMAT=<ID=%s,Number=1,Type=%s,Description=\"%s\">\n", HDR_FMT[i][0], HDR_FMT[i][1], HDR_FMT[i][2]);
    if (!is_ctx) for (int i = 0; HDR_CNV[i][0]; i++) HP("##FORMAT=<ID=%s,Number=1,Type=Float,Description=\"%s\"\n", HDR_CNV[i][0], HDR_CNV[i][1]);
    HP("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\n");
#undef HP
    return w;
}
