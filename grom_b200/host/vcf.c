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
#include "gromhost.h"

int64_t gromhost_vcf_snv(const grom_params *p, const char *chr_name, const char *fasta,
                         const grom_snv_cand *snv, int64_t n, double ave_rd, char *buf, int64_t cap)
{
    int64_t w = 0;
    char gt[512];
    for (int64_t i = 0; i < n; i++) {
        const grom_snv_cand *c = &snv[i];
        if (!(c->v[GA_RC_ALL] <= round(p->snv_rd_min_factor * ave_rd) || c->ratio >= p->high_cov_min_snv_ratio)) continue;
        int cn = (int)round(c->ratio * p->ploidy);
        if (cn == 0) cn = 1;
        for (int k = 0; k < p->ploidy && k < 250; k++) { gt[2 * k] = k < cn ? '1' : '0'; gt[2 * k + 1] = k < p->ploidy - 1 ? '/' : '\0'; }
        const int nb = c->v[GA_SNV_A + c->base];
        if (cap - w < 1024) return -1;
        w += snprintf(buf + w, (size_t)(cap - w),
                      "%s\t%d\t\t%c\t%c\t.\t.\t.\tGT:PR:AF:A:C:G:T:AL:CL:GL:TL:BQ:MQ:PIR:FS\t%s:%e:%e:%d:%d:%d:%d:%d:%d:%d:%d:%.2f:%.2f:%.2f:%.2f\n",
                      chr_name, c->pos + 1, fasta[c->pos], "ACGT"[c->base], gt, c->pr, c->ratio,
                      c->v[GA_SNV_A], c->v[GA_SNV_C], c->v[GA_SNV_G], c->v[GA_SNV_T],
                      c->v[GA_SNVLOW_A], c->v[GA_SNVLOW_C], c->v[GA_SNVLOW_G], c->v[GA_SNVLOW_T],
                      (double)c->v[GA_BQ_ALL] / (double)c->v[GA_RC_ALL], (double)c->v[GA_MQ_ALL] / (double)c->v[GA_RC_ALL],
                      (double)c->v[GA_PIR_A + c->base] / (double)nb, (double)c->v[GA_FS_A + c->base] / (double)nb);
    }
    return w;
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

int64_t gromhost_vcf_smalldel(const grom_params *p, const char *chr_name, const char *fasta, int64_t chr_len,
                              const grom_del_event *ev, int64_t n, char *buf, int64_t cap)
{
    /* ---- pairing state machine over the events in (position, start-before-end) order, src/GROM.c:11475-11745 */
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
    /* ---- emission, src/GROM.c:16351-16490 (the loop stops before the entry the index points at) */
    int64_t w = 0;
    for (int a = 0; a < idx; a++) {
        const delrec *d = &L[a];
        if (!(d->s_pr <= p->pval_threshold && d->e_pr <= p->pval_threshold &&
              (double)d->s_w / (double)d->s_rd > p->min_indel_ratio * (double)p->add_factor &&
              (double)d->e_w / (double)d->e_rd > p->min_indel_ratio * (double)p->add_factor)) continue;
        if (d->start < 0 || d->end < 0 || d->end >= chr_len) continue;
        const int hp = homopolymer(fasta, chr_len, d->start, d->end, 1);
        if (hp > 10) continue;
        const int cn = d->end - d->start + 1;
        if (cap - w < 2048) { free(L); return -1; }
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
    free(L);
    return w;
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
