/* ctx.c -- translocation (inter-contig breakpoint) records: the per-contig candidate merge and the genome-level mate pairing that
 * produce the reference's <out>.ctx.vcf.
 *
 *   per contig   cdp_ctx_f_list / cdp_ctx_r_list -> list2: candidates within ins_max - 2 lseq of the kept one are folded into it
 *                (better p-value with no less evidence wins), then the emission filter                     src/GROM.c:16098-16246
 *   per genome   every record looks for a record on its mate contig that points back at it with a matching orientation; of several
 *                paired records describing the same junction the most significant survives                 src/GROM.c:22470-22745
 *
 * The reference passes the per-contig records through a text file ("%e" / "%.1f") and parses them back with atof(): the pairing
 * compares the 7-digit values.  gromhost_ctx_contig applies the same rounding to the fields it returns.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "gromhost.h"

static double through_text(const char *fmt, double v) { char t[64]; snprintf(t, sizeof(t), fmt, v); return atof(t); }

/* candidates of one orientation -> surviving records, appended to out (caller-allocated, capacity >= n) */
static int64_t merge_and_filter(const grom_params *p, int tid, int type, const grom_sv_event *l, int64_t n, grom_ctx_record *out)
{
    grom_sv_event *m = (grom_sv_event *)malloc((size_t)(n + 1) * sizeof(grom_sv_event));
    int64_t n2 = 0, w = 0;
    int begun = 0;
    const int reach = p->insert_max - 2 * p->lseq;
    for (int64_t a = 0; a < n; a++) {
        const grom_sv_event *c = &l[a];
        if (begun) {
            grom_sv_event *k = &m[n2 - 1];
            if (c->pos > k->pos + reach) begun = 0;
            else if (((c->binom < k->binom && k->weight <= c->weight) || (c->binom == k->binom && k->weight < c->weight)) && c->pos >= 0) *k = *c;
        }
        if (!begun && c->pos >= 0 && n2 < 100000 - 1) { begun = 1; m[n2++] = *c; }       /* g_sv_list2_len */
    }
    for (int64_t a = 0; a < n2; a++) {
        const grom_sv_event *k = &m[a];
        if (!((k->binom <= p->pval_threshold || k->hez <= p->pval_threshold) &&
              (double)k->weight / (double)k->rd >= p->min_sv_ratio * (double)p->add_factor)) continue;
        grom_ctx_record *r = &out[w++];
        r->type = type; r->chr = tid; r->pos = k->pos; r->binom = through_text("%e", k->binom);
        r->evidence = through_text("%.1f", (double)k->weight / (double)p->add_factor);
        r->rd = k->rd; r->conc = k->conc; r->other_len = k->other_len; r->mchr = k->mchr; r->mpos = (int32_t)k->dist;
        r->read_start = k->read_start; r->read_end = k->read_end; r->hez = through_text("%e", k->hez);
        r->mate_id = -1; r->keep = 0;
    }
    free(m);
    return w;
}

int64_t gromhost_ctx_contig(const grom_params *p, int tid, const grom_sv_event *ctx_f, int64_t n_f, const grom_sv_event *ctx_r, int64_t n_r,
                            grom_ctx_record *out, int64_t cap)
{
    if (cap < n_f + n_r) return -1;
    int64_t w = merge_and_filter(p, tid, 6, ctx_f, n_f, out);                             /* g_sv_types[6] = CTX_F, [7] = CTX_R */
    w += merge_and_filter(p, tid, 7, ctx_r, n_r, out + w);
    return w;
}

/* rec: the records of all contigs in the order the contigs were processed (modified: mate id, mate position, keep flag) */
int64_t gromhost_ctx_vcf(const grom_params *p, const char *const *target_names, int n_targets, grom_ctx_record *rec, int64_t n, char *buf, int64_t cap)
{
    const int reach = p->insert_max - 2 * p->lseq;
    for (int64_t b = 0; b < n; b++) { rec[b].keep = 0; rec[b].mate_id = -1; }
    for (int64_t b = 0; b < n; b++)
        for (int64_t c = 0; c < n; c++) {
            if (!(rec[b].chr == rec[c].mchr && rec[c].chr == rec[b].mchr)) continue;
            if (!(abs(rec[b].pos - abs(rec[c].mpos)) < reach && abs(rec[c].pos - abs(rec[b].mpos)) < reach)) continue;
            if (!(((rec[b].type == 6 && rec[c].mpos >= 0) || (rec[b].type == 7 && rec[c].mpos < 0)) &&
                  ((rec[c].type == 6 && rec[b].mpos >= 0) || (rec[c].type == 7 && rec[b].mpos < 0)))) continue;
            rec[b].keep = 1; rec[b].mate_id = (int32_t)c;
            rec[b].mpos = rec[b].mpos < 0 ? -rec[c].pos : rec[c].pos;                   /* the mate's own position replaces the running mean */
        }
    for (int64_t b = 0; b < n; b++)
        for (int64_t c = 0; c < n; c++) {
            if (b == c || rec[b].chr != rec[c].chr || rec[b].mchr != rec[c].mchr) continue;
            if (!(abs(rec[b].pos - rec[c].pos) < reach && abs(abs(rec[b].mpos) - abs(rec[c].mpos)) < reach)) continue;
            if (rec[b].keep == 1 && rec[c].keep == 1 && (rec[b].binom > rec[c].binom || (rec[b].binom == rec[c].binom && b > c))) {
                rec[b].keep = 0;
                if (rec[b].mate_id >= 0) rec[rec[b].mate_id].keep = 0;
            }
        }
    int64_t w = 0;
    for (int64_t b = 0; b < n; b++) {
        const grom_ctx_record *r = &rec[b];
        if (r->keep != 1) continue;
        if (r->chr < 0 || r->chr >= n_targets || r->mchr < 0 || r->mchr >= n_targets) continue;
        char alt[512];
        const char *mn = target_names[r->mchr];
        if (r->type == 6 && r->mpos < 0) snprintf(alt, sizeof(alt), "N[%s:%d[", mn, abs(r->mpos));
        else if (r->type == 6) snprintf(alt, sizeof(alt), "N]%s:%d]", mn, abs(r->mpos));
        else if (r->mpos < 0) snprintf(alt, sizeof(alt), "[%s:%d[N", mn, abs(r->mpos));
        else snprintf(alt, sizeof(alt), "]%s:%d]N", mn, abs(r->mpos));
        if (cap - w < 1024) return -1;
        w += snprintf(buf + w, (size_t)(cap - w), "%s\t%d\t%d\tN\t%s\t.\t.\tSVTYPE=BND;MATEID=%d\tSPR:SEV:SRD:SCO:SOT:SFR:SLR:SHPR\t%e:%.1f:%d:%d:%d:%d:%d:%e\n",
                      target_names[r->chr], r->pos + 1, (int)b, alt, r->mate_id, r->binom, r->evidence, r->rd, r->conc, r->other_len,
                      r->read_start + 1, r->read_end + 1, r->hez);
    }
    return w;
}
