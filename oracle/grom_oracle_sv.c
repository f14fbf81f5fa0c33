/* oracle/grom_oracle_sv.c -- TEST INFRASTRUCTURE ONLY (see grom_oracle.h).
 *
 * CPU restatement of the order-dependent evidence the reference accumulates per read:
 *   - CIGAR insertions / deletions into first-seen primary slots          (src/GROM.c:7187-7423)
 *   - split-read (SA/XP) small deletions, large deletions, tandem dups    (src/GROM.c:7425-7950, 7978-8340, 9361-9722)
 *   - discordant / concordant pair ranges with online clusters            (src/GROM.c:7955-10953)
 * The reference repeats one cluster-update template ~20 times with small variations
 * (value clustered, tolerance, which end is the anchor, how first/last read positions
 * are tracked).  Here the template is one function, cl_update(), parameterised by
 * exactly those variations; every call site cites the block it stands for.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "grom_oracle_int.h"

/* C abs() on a double argument: the reference passes doubles to int abs(), i.e. the value is
 * truncated toward zero first (src/GROM.c:8420 and every sibling). */
static inline int iabs_d(double v) { int t = (int)v; return t < 0 ? -t : t; }

enum { RS_SET_RE = 0,     /* join: read_end = v                                   (pair ranges, src/GROM.c:8437) */
       RS_MINMAX = 1,     /* join: read_start = min(.,v), read_end = max(.,v)     (src/GROM.c:8717-8724, 8056-8077) */
       RS_MAX_ONLY = 2 }; /* join: read_end = max(.,v), read_start untouched       (split-read del_f, src/GROM.c:7560-7575) */

typedef struct {
    int cls;          /* CL_* */
    double x;         /* clustered value */
    int w;            /* weight added: add, or add/2 */
    double wd;        /* (double)add, or (double)add/2.0: numerator factor of the running mean */
    double tol;       /* ins_max - ins_min (+ insert_temp for inversions) */
    int v;            /* read position recorded in read_start / read_end */
    int mode;         /* RS_* */
    int is_ctx;       /* translocation: also match mate contig and the sign of the stored value */
    int mchr;
    int sign;         /* ctx: +1 mate forward (stores +mpos), -1 mate reverse (stores -mpos) */
    int w_repl;       /* threshold `cdp_add` of the replacement scan (src/GROM.c:8508) */
} clu;

static oslot *others_at(svctx *c, int64_t x)
{
    if (!c->oth[x]) c->oth[x] = (oslot *)calloc((size_t)c->p->other_len, sizeof(oslot));
    return c->oth[x];
}

static int join_ok(const clu *u, double dist, int w, int mchr)
{
    if (u->is_ctx) {
        if (mchr != u->mchr) return 0;
        double mp = u->sign > 0 ? u->x : -u->x;                     /* +mpos */
        if (u->sign > 0) return iabs_d(dist - mp) <= u->tol * (1.0 + (1.0 / (double)w)) && dist > 0;
        return iabs_d((double)iabs_d(dist) - mp) <= u->tol * (1.0 + (1.0 / (double)w)) && dist < 0;
    }
    return iabs_d(dist - u->x) <= u->tol * (1.0 + (1.0 / (double)w));
}

static void track(int mode, int v, int *rs, int *re)
{
    if (mode == RS_SET_RE) *re = v;
    else if (mode == RS_MINMAX) { if (v < *rs) *rs = v; if (v > *re) *re = v; }
    else { if (v > *re) *re = v; }
}

/* one evidence item for cluster class u->cls at position x */
void cl_update(svctx *c, int64_t x, const clu *u)
{
    if (x < 0 || x >= c->P) return;
    int k = u->cls;
    int32_t *W = c->cw[k] + x, *RS = c->crs[k] + x, *RE = c->cre[k] + x;
    double *D = c->cdist[k] + x;
    int32_t *MC = u->is_ctx ? c->cmchr[k - CL_CTX_F] + x : NULL;
    if (*W == 0) {
        *W = u->w; *D = u->x; *RS = u->v; *RE = u->v;
        if (MC) *MC = u->mchr;
        return;
    }
    if (join_ok(u, *D, *W, MC ? *MC : 0)) {
        *W += u->w;
        *D += u->wd * (u->x - *D) / (double)(*W);
        track(u->mode, u->v, RS, RE);
        return;
    }
    oslot *o = others_at(c, x);
    const int otype = k + 1;
    for (int s = 0; s < c->p->other_len; s++) {
        if (o[s].type == otype) {
            if (join_ok(u, o[s].dist, o[s].w, o[s].mchr)) {
                o[s].w += u->w;
                o[s].dist += u->wd * (u->x - o[s].dist) / (double)o[s].w;
                track(u->mode, u->v, &o[s].rs, &o[s].re);
                if (o[s].w > *W) {               /* overtaken: swap with the primary */
                    oslot t = o[s];
                    o[s].w = *W; o[s].dist = *D; o[s].rs = *RS; o[s].re = *RE;
                    *W = t.w; *D = t.dist; *RS = t.rs; *RE = t.re;
                    if (MC) { o[s].mchr = *MC; *MC = t.mchr; }
                }
                return;
            }
        } else if (o[s].type == OTHER_EMPTY) {
            o[s].w = u->w; o[s].type = otype; o[s].dist = u->x; o[s].rs = u->v; o[s].re = u->v;
            if (u->is_ctx) o[s].mchr = u->mchr;
            return;
        }
    }
    for (int s = 0; s < c->p->other_len; s++) {
        if (o[s].w <= u->w_repl) {
            o[s].w = u->w; o[s].type = otype; o[s].dist = u->x; o[s].rs = u->v; o[s].re = u->v;
            if (u->is_ctx) o[s].mchr = u->mchr;
            return;
        }
    }
}

/* small-indel primary slot (indel_i / indel_d_f / indel_d_r): exact-length match, src/GROM.c:7213-7285, 7292-7350, 7353-7415 */
static void indel_update(svctx *c, int64_t x, int ga_w, int ga_dist, int otype, int len, int add)
{
    if (x < 0 || x >= c->P) return;
    int32_t *W = c->A->a[ga_w] + x, *D = c->A->a[ga_dist] + x;
    if (*W == 0) { *W = add; *D = len; return; }
    if (len == *D) { *W += add; return; }
    oslot *o = others_at(c, x);
    for (int s = 0; s < c->p->other_len; s++) {
        if (o[s].type == otype) {
            if ((uint32_t)len == (uint32_t)(o[s].dist + 0.5)) {
                o[s].w += add;
                if (o[s].w > *W) {
                    int tw = o[s].w; double td = o[s].dist;
                    o[s].w = *W; o[s].dist = (double)*D;
                    *W = tw; *D = (int32_t)(uint32_t)(td + 0.5);
                }
                return;
            }
        } else if (o[s].type == OTHER_EMPTY) {
            o[s].w = add; o[s].type = otype; o[s].dist = (double)len;
            return;
        }
    }
    for (int s = 0; s < c->p->other_len; s++) {
        if (o[s].w <= add) { o[s].w = add; o[s].type = otype; o[s].dist = (double)len; o[s].rs = 0; o[s].re = 0; return; }
    }
}

static inline void rd_inc(svctx *c, int64_t x) { if (x >= 0 && x < c->P) c->A->a[GA_RD][x] += 1; }

/* forward pair range [lo, hi) with a cluster update at every position, anchor = lo (src/GROM.c:8400-8525 and siblings) */
static void range_fwd(svctx *c, int64_t lo, int64_t hi, clu u, int add, int end_adj)
{
    for (int64_t x = lo; x < hi; x++) {
        rd_inc(c, x);
        int full = (end_adj < c->p->sc_min) || x == lo;
        u.w = full ? add : add / 2; u.wd = full ? (double)add : (double)add / 2.0; u.w_repl = add;
        cl_update(c, x, &u);
    }
}
/* backward pair range [lo, hi), anchor = hi - 1 (src/GROM.c:9076-9203 and siblings) */
static void range_bwd(svctx *c, int64_t lo, int64_t hi, clu u, int add, int start_adj)
{
    for (int64_t x = lo; x < hi; x++) {
        rd_inc(c, x);
        int full = (start_adj < c->p->sc_min) || x == hi - 1;
        u.w = full ? add : add / 2; u.wd = full ? (double)add : (double)add / 2.0; u.w_repl = add;
        cl_update(c, x, &u);
    }
}

void sv_evidence_read(svctx *c, const svread *r)
{
    const grom_params *p = c->p;
    const int pos = r->pos, mpos = r->mpos, tlen = r->tlen, flag = r->flag, add = r->add;
    const int lseq = r->lseq, start_adj = r->start_adj, end_adj = r->end_adj, indel = r->end_adj_indel;
    const int ins_min = p->insert_min, ins_max = p->insert_max, ins_mean = p->insert_mean;
    const int paired = (flag & 1) != 0, munmap = (flag & 8) != 0, rev = (flag & 16) != 0, mrev = (flag & 32) != 0;
    const int same = (r->tid == r->mtid);
    const int64_t E = (int64_t)pos - start_adj + lseq - end_adj - indel;          /* reference end of the alignment */
    const int64_t F = (int64_t)pos - start_adj - indel + ins_max - lseq;           /* forward horizon */
    const int64_t Bk0 = (int64_t)pos - start_adj - ins_max + 2 * lseq;             /* backward horizon */
    const int64_t win_lo = r->win_lo, win_hi = r->win_lo + c->W;                   /* the reference's window when this read is applied */
    const int64_t Bk = Bk0 < win_lo ? win_lo : Bk0;
    const double tol = (double)(ins_max - ins_min);
    int insert_temp = ins_mean - 2 * lseq; if (insert_temp < 0) insert_temp = 0;   /* src/GROM.c:7953-7958 */
    const double tolI = (double)(ins_max - ins_min + insert_temp);
#define MIN2(a, b) ((a) < (b) ? (a) : (b))

    /* ---- CIGAR indels, src/GROM.c:7187-7423 */
    {
        int64_t tp = pos; int qoff = 0;
        for (int k = 0; k < r->n_cigar; k++) {
            int op = r->cigar[k] & 15, len = (int)(r->cigar[k] >> 4);
            if (op == 0 || op == 3 || op == 7 || op == 8) { tp += len; if (op != 3) qoff += len; }
            else if (op == 4) qoff += len;
            else if (op == 1) {
                if (tp >= 0 && tp < c->P && c->A->a[GA_INDEL_I][tp] == 0 && len <= c->p->indel_i_seq_len) {
                    /* the primary slot is (re)created: its first `len` sequence characters are overwritten, longer
                     * leftovers of an earlier zero-weight creation stay (src/GROM.c:7219-7228) */
                    if (!c->ins_seq[tp]) c->ins_seq[tp] = (char *)calloc(64, 1);
                    const grom_read_batch *b = r->batch;
                    for (int j = 0; j < len; j++) {
                        uint64_t slot = b->base_off[r->read_index] + (uint64_t)(qoff + j);
                        c->ins_seq[tp][j] = "=ACMGRSVTWYHKDBN"[(b->seq4[slot >> 1] >> ((~slot & 1) << 2)) & 15];
                    }
                }
                indel_update(c, tp, GA_INDEL_I, GA_INDEL_IDIST, OTHER_INDEL_I, len, add);
                qoff += len;
            }
            else if (op == 2) {
                if (tp >= 0 && tp < c->P) c->A->a[GA_INDEL_D_F_RD][tp] += 1;
                indel_update(c, tp, GA_INDEL_D_F, GA_INDEL_D_FDIST, OTHER_INDEL_D_F, len, add);
                int64_t te = tp + len - 1;
                if (te >= 0 && te < c->P) c->A->a[GA_INDEL_D_R_RD][te] += 1;
                indel_update(c, te, GA_INDEL_D_R, GA_INDEL_D_RDIST, OTHER_INDEL_D_R, len, add);
                tp += len;
            }
        }
    }

    /* ---- split-read deletions, src/GROM.c:7425-7950 */
    sv_split_del(c, r);

    /* ---- pairs, src/GROM.c:7963-10953 */
    clu u; memset(&u, 0, sizeof(u)); u.v = pos; u.mode = RS_SET_RE; u.tol = tol;
    if (paired && !munmap) {
        if (same) {
            if (mpos > pos) {
                if (!rev && mrev) {
                    if (tlen >= ins_min && tlen <= ins_max) {                                  /* concordant, src/GROM.c:7976-8366 */
                        if (!sv_split_dup_fwd(c, r)) {
                            int64_t hi = MIN2((int64_t)mpos, win_hi);
                            for (int64_t x = E; x < hi; x++) { rd_inc(c, x); if (x >= 0 && x < c->P) c->A->a[GA_CONC][x] += 1; }
                        }
                    } else if (tlen > 2 * ins_max) {                                           /* src/GROM.c:8370-8526 */
                        int64_t hi = MIN2(MIN2(F, win_hi), (int64_t)mpos);
                        u.cls = CL_DEL_F; u.x = (double)tlen;
                        range_fwd(c, E, hi, u, add, end_adj);
                    } else if (tlen > ins_max) {                                               /* src/GROM.c:8531-8822 */
                        int64_t hi = MIN2((int64_t)mpos, win_hi);
                        for (int64_t x = E; x < hi; x++) {
                            rd_inc(c, x);
                            if (x < F) {
                                clu a = u; a.cls = CL_DEL_F; a.x = (double)tlen;
                                int full = (end_adj < p->sc_min) || x == E;
                                a.w = full ? add : add / 2; a.wd = full ? (double)add : (double)add / 2.0; a.w_repl = add;
                                cl_update(c, x, &a);
                            }
                            if (abs(tlen) <= 2 * ins_max && x > (int64_t)pos - start_adj + tlen - ins_max + lseq) {
                                clu a = u; a.cls = CL_DEL_R; a.x = (double)tlen; a.v = mpos; a.mode = RS_MINMAX;
                                int full = (start_adj < p->sc_min) || x == hi - 1;
                                a.w = full ? add : add / 2; a.wd = full ? (double)add : (double)add / 2.0; a.w_repl = add;
                                cl_update(c, x, &a);
                            }
                        }
                    } else if (tlen < ins_min) {                                               /* src/GROM.c:8823-8873 */
                        int no_ins = 0;
                        if (r->sa_pos >= 0 && r->sa_same && !rev && r->sa_strand == 0 && r->sa_pos < pos && pos < mpos) no_ins = 1;
                        if (!no_ins) {
                            int64_t hi = MIN2((int64_t)mpos, win_hi);
                            for (int64_t x = E; x < hi; x++) { rd_inc(c, x); if (x >= 0 && x < c->P) c->A->a[GA_INS][x] += add; }
                        }
                    }
                } else if (!rev && !mrev) {                                                    /* src/GROM.c:8879-9041 */
                    if (mpos - pos >= 10) {
                        int64_t hi = MIN2(MIN2(F, win_hi), (int64_t)mpos);
                        u.cls = CL_INV_F1; u.x = (double)tlen; u.tol = tolI;
                        range_fwd(c, E, hi, u, add, end_adj);
                    }
                } else if (rev) {                                                              /* src/GROM.c:9042-9345 */
                    if (mpos - pos >= 10) {
                        u.cls = mrev ? CL_INV_R1 : CL_DUP_R; u.x = (double)tlen; u.tol = mrev ? tolI : tol;
                        range_bwd(c, Bk, pos, u, add, start_adj);
                    }
                }
            } else {
                if (rev && !mrev) {                                                            /* src/GROM.c:9355-9868 */
                    if (abs(tlen) >= ins_min && abs(tlen) <= ins_max) {
                        sv_split_dup_rev(c, r);
                    } else if (abs(tlen) > 2 * ins_max) {
                        u.cls = CL_DEL_R; u.x = (double)abs(tlen);
                        range_bwd(c, Bk, pos, u, add, start_adj);
                    }
                } else if (!rev && !mrev) {                                                    /* src/GROM.c:9869-10019 */
                    if (pos - mpos >= 10) {
                        u.cls = CL_INV_F2; u.x = (double)abs(tlen); u.tol = tolI;
                        range_fwd(c, E, MIN2(F, win_hi), u, add, end_adj);
                    }
                } else if (mrev) {                                                             /* src/GROM.c:10020-10312 */
                    if (pos - mpos >= 10) {
                        if (!rev) {
                            u.cls = CL_DUP_F; u.x = (double)abs(tlen);
                            range_fwd(c, E, MIN2(F, win_hi), u, add, end_adj);
                        } else {
                            int64_t lo = Bk0; if (lo < (int64_t)mpos + lseq) lo = (int64_t)mpos + lseq;
                            u.cls = CL_INV_R2; u.x = (double)abs(tlen); u.tol = tolI;
                            range_bwd(c, lo, pos, u, add, start_adj);
                        }
                    }
                }
            }
        } else {                                                                               /* other contig, src/GROM.c:10313-10901 */
            u.is_ctx = 1; u.mchr = r->mtid; u.sign = mrev ? -1 : 1; u.x = mrev ? -(double)mpos : (double)mpos;
            if (!rev) {
                u.cls = CL_CTX_F;
                range_fwd(c, E, MIN2(F, win_hi), u, add, end_adj);
            } else {
                int64_t lo = (int64_t)pos - start_adj + lseq - ins_max + lseq; if (lo < win_lo) lo = win_lo;
                u.cls = CL_CTX_R;
                range_bwd(c, lo, pos, u, add, start_adj);
            }
        }
    } else if (paired && munmap) {                                                             /* src/GROM.c:10902-10952 */
        if (!rev) {
            int64_t hi = MIN2(F, win_hi);
            for (int64_t x = E; x < hi; x++) { rd_inc(c, x); if (x >= 0 && x < c->P) c->A->a[GA_MUNMAPPED_F][x] += add; }
        } else {
            int64_t lo = (int64_t)pos - start_adj + lseq + indel - ins_max + lseq; if (lo < win_lo) lo = win_lo;
            for (int64_t x = lo; x < pos; x++) { rd_inc(c, x); if (x >= 0 && x < c->P) c->A->a[GA_MUNMAPPED_R][x] += add; }
        }
    }
#undef MIN2
}

/* ---- split reads (first SA/XP entry on this contig), src/GROM.c:7425-7950, 7978-8340, 9361-9722 ---- */

static int sa_usable(const svctx *c, const svread *r)
{
    return r->sa_pos >= 0 && r->sa_same && r->sa_mapq >= c->p->min_mapq && r->mapq >= c->p->min_mapq;
}

/* split-read deletion: small gaps also feed the small-indel slots, every gap feeds del_f / del_r with
 * x = gap + insert_mean (src/GROM.c:7425-7950) */
void sv_split_del(svctx *c, const svread *r)
{
    const grom_params *p = c->p;
    if (!(r->sa_pos >= 0 && r->sa_same)) return;
    if (!(r->sa_mapq >= p->min_mapq && r->mapq >= p->min_mapq)) return;
    const int flag = r->flag, rev = (flag & 16) != 0;
    const int paired_same = (flag & 1) && !(flag & 8) && r->tid == r->mtid;
    const int pos = r->pos, apos = r->sa_pos, lseq = r->lseq;
    const int64_t E = (int64_t)pos - r->start_adj + lseq - r->end_adj - r->end_adj_indel;
    const int64_t AE = (int64_t)apos - r->sa_start_adj + lseq - r->sa_end_adj - r->sa_end_adj_indel;
    int sr_del = 0; int64_t S = 0, Eo = 0;
    if (!((!rev && r->sa_strand == 0) || (rev && r->sa_strand == 1))) return;
    if (paired_same) {
        if (!rev && r->sa_strand == 0) {
            if (pos < apos && r->tlen <= p->insert_max && apos < r->mpos) {
                if (apos - E < p->insert_max && apos - E > 0) {
                    if (abs(lseq - r->end_adj - r->sa_start_adj) <= p->max_split_loss && lseq - r->start_adj - r->end_adj - r->end_adj_indel >= p->min_sr_len &&
                        lseq - r->sa_start_adj - r->sa_end_adj - r->sa_end_adj_indel >= p->min_sr_len) { sr_del = 1; S = E; Eo = apos; }
                }
            }
        } else if (rev && r->sa_strand == 1) {
            if (apos < pos && abs(r->tlen) < p->insert_max && r->mpos < apos) {
                if (abs(lseq - r->start_adj - r->sa_end_adj) <= p->max_split_loss && lseq - r->start_adj - r->end_adj - r->end_adj_indel >= p->min_sr_len &&
                    lseq - r->sa_start_adj - r->sa_end_adj - r->sa_end_adj_indel >= p->min_sr_len) { S = AE; Eo = pos; if (S < Eo) sr_del = 1; }
            }
        }
    } else {
        if (!rev && r->sa_strand == 0) {
            if (pos < apos && apos - E < p->insert_max && apos - E > 0) { sr_del = 1; S = E; Eo = apos; }
        } else if (rev && r->sa_strand == 1) {
            if (apos < pos && pos - AE < p->insert_max) { S = AE; Eo = pos; if (S < Eo) sr_del = 1; }
        }
    }
    if (!sr_del) return;
    const int64_t gap = Eo - S;
    if (gap < p->lseq && gap < p->insert_max - p->insert_mean) {
        if (S >= 0 && S < c->P) c->A->a[GA_INDEL_D_F_RD][S] += 1;
        indel_update(c, S, GA_INDEL_D_F, GA_INDEL_D_FDIST, OTHER_INDEL_D_F, (int)gap, r->add);
        if (Eo - 1 >= 0 && Eo - 1 < c->P) c->A->a[GA_INDEL_D_R_RD][Eo - 1] += 1;
        indel_update(c, Eo - 1, GA_INDEL_D_R, GA_INDEL_D_RDIST, OTHER_INDEL_D_R, (int)gap, r->add);
    }
    clu u; memset(&u, 0, sizeof(u));
    u.x = (double)(gap + p->insert_mean); u.w = r->add; u.wd = (double)r->add; u.w_repl = r->add; u.tol = (double)(p->insert_max - p->insert_min);
    rd_inc(c, S);
    u.cls = CL_DEL_F; u.v = pos < apos ? pos : apos; u.mode = RS_MAX_ONLY;
    cl_update(c, S, &u);
    rd_inc(c, Eo - 1);
    u.cls = CL_DEL_R; u.v = pos < apos ? apos : pos; u.mode = RS_MINMAX;
    cl_update(c, Eo - 1, &u);
}

/* tandem-duplication breakpoints from a split read: dup_f at the downstream junction, dup_r at the upstream one,
 * x = span - insert_mean (src/GROM.c:8016-8340, 9402-9722).  Bug-compatible: when the dup_f primary is created its
 * read_end is not written; del_f's read_end at that position is written instead (src/GROM.c:8037-8045, 9423-9429). */
static void split_dup_evidence(svctx *c, const svread *r, int64_t lp_start, int64_t lp_end)
{
    const grom_params *p = c->p;
    const int pos = r->pos, apos = r->sa_pos;
    clu u; memset(&u, 0, sizeof(u));
    u.x = (double)(lp_end - lp_start - p->insert_mean); u.w = r->add; u.wd = (double)r->add; u.w_repl = r->add;
    u.tol = (double)(p->insert_max - p->insert_min); u.mode = RS_MINMAX;
    rd_inc(c, lp_end);
    if (lp_end >= 0 && lp_end < c->P) {
        u.cls = CL_DUP_F; u.v = pos < apos ? apos : pos;
        const int was_empty = c->cw[CL_DUP_F][lp_end] == 0;
        const int32_t old_re = c->cre[CL_DUP_F][lp_end];
        cl_update(c, lp_end, &u);
        if (was_empty) { c->cre[CL_DUP_F][lp_end] = old_re; c->cre[CL_DEL_F][lp_end] = u.v; }
    }
    rd_inc(c, lp_start - 1);
    u.cls = CL_DUP_R; u.v = pos < apos ? pos : apos;
    cl_update(c, lp_start - 1, &u);
}

/* forward read of a concordant-looking pair whose SA lies between read and mate (src/GROM.c:7978-8011) */
int sv_split_dup_fwd(svctx *c, const svread *r)
{
    const grom_params *p = c->p;
    if (!sa_usable(c, r)) return 0;
    const int rev = (r->flag & 16) != 0;
    if (!(!rev && r->sa_strand == 0)) return 0;
    if (!((r->flag & 1) && !(r->flag & 8) && r->tid == r->mtid)) return 0;
    if (!(r->pos < r->sa_pos && r->sa_pos < r->mpos)) return 0;
    const int it = r->end_adj_indel > 0 ? r->end_adj_indel : 0;
    const int ait = r->sa_end_adj_indel > 0 ? r->end_adj_indel : 0;        /* sic: the read's own value, src/GROM.c:7998 */
    if (!(abs(r->lseq - r->start_adj - r->sa_end_adj) <= p->max_split_loss && r->lseq - r->start_adj - r->end_adj - it >= p->min_sr_len &&
          r->lseq - r->sa_start_adj - r->sa_end_adj - ait >= p->min_sr_len)) return 0;
    const int64_t lp_start = r->pos;
    const int64_t lp_end = (int64_t)r->sa_pos - r->sa_start_adj + r->lseq - r->sa_end_adj - r->sa_end_adj_indel;
    split_dup_evidence(c, r, lp_start, lp_end);
    return 1;
}

/* reverse read, mate upstream, SA between mate and read (src/GROM.c:9361-9400) */
void sv_split_dup_rev(svctx *c, const svread *r)
{
    const grom_params *p = c->p;
    if (!sa_usable(c, r)) return;
    const int rev = (r->flag & 16) != 0;
    if (!(rev && r->sa_strand == 1)) return;
    if (!((r->flag & 1) && !(r->flag & 8) && r->tid == r->mtid)) return;
    if (!(r->sa_pos < r->pos && r->mpos < r->sa_pos)) return;
    const int it = r->end_adj_indel > 0 ? r->end_adj_indel : 0;
    const int ait = r->sa_end_adj_indel > 0 ? r->end_adj_indel : 0;
    if (!(abs(r->lseq - r->sa_start_adj - r->end_adj) <= p->max_split_loss && r->lseq - r->start_adj - r->end_adj - it >= p->min_sr_len &&
          r->lseq - r->sa_start_adj - r->sa_end_adj - ait >= p->min_sr_len)) return;
    const int64_t lp_start = r->sa_pos;
    const int64_t lp_end = (int64_t)r->pos - r->start_adj + r->lseq - r->end_adj - r->end_adj_indel;
    if (!(lp_start < lp_end)) return;
    split_dup_evidence(c, r, lp_start, lp_end);
}
