/* oracle/grom_oracle.c -- TEST INFRASTRUCTURE ONLY (see grom_oracle.h).
 *
 * Sequential CPU restatement of the reference's per-chromosome evidence
 * accumulation and SNV scan over chromosome-length arrays.  Each block cites the
 * reference lines it follows.  Written from the behaviour of the reference, not
 * from its text: one pass over the reads in BAM order, no sliding window -- the
 * window geometry of the reference (src/GROM.c:5846-6402) is reproduced only
 * where it is observable (which reads are consumed, which positions are scanned,
 * the look-ahead read length, range truncation at the window ends).
 */
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "grom_oracle_int.h"

#define F_PAIRED 1
#define F_UNMAP 4
#define F_MUNMAP 8
#define F_REVERSE 16
#define F_MREVERSE 32
#define F_DUP 1024

enum { OP_M = 0, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X };
enum { SV_DEL = 0, SV_DUP = 1, SV_INV_F = 8, SV_INV_R = 9, SV_CTX_FF = 11, SV_CTX_FR, SV_CTX_RF, SV_CTX_RR };

static const char NT16[] = "=ACMGRSVTWYHKDBN";


static inline int base_code(const grom_read_batch *b, int64_t i, int k)
{
    uint64_t slot = b->base_off[i] + (uint64_t)k;
    return (b->seq4[slot >> 1] >> ((~slot & 1) << 2)) & 15;
}

/* -M svtype, src/GROM.c:6435-6542 */
static int dup_svtype(int tid, int mtid, int pos, int mpos, int flag)
{
    int rev = (flag & F_REVERSE) != 0, mrev = (flag & F_MREVERSE) != 0;
    if (tid == mtid) {
        if (mpos > pos) {
            if (!rev && mrev) return SV_DEL;
            if (!rev && !mrev) return SV_INV_F;
            return mrev ? SV_INV_R : SV_DUP;
        }
        if (rev && !mrev) return SV_DEL;
        if (!rev && !mrev) return SV_INV_F;
        if (mrev) return rev ? SV_INV_R : SV_DUP;
        return -1;
    }
    if (!rev) return mrev ? SV_CTX_FR : SV_CTX_FF;
    return mrev ? SV_CTX_RR : SV_CTX_RF;
}

typedef struct { int mpos, mtid, lseq, tlen, svtype; } dupkey;

int oracle_run_chr(const grom_params *p, const grom_read_batch *b, const char *fasta, int64_t P,
                   const double *hez_tbl, const double *mq_tbl, oracle_chr_out *out)
{
    const int64_t n = b->n_reads;
    const int q = p->min_mapq, bqmin = p->min_base_qual;
    const int W = grom_window_len(p), first_pos = grom_first_pos(p);
    const int TD = p->max_trials + 1;
    arrs A; A.len = P;
    memset(out->arrays, 0, sizeof(int32_t) * (size_t)GA_COUNT * (size_t)P);
    for (int k = 0; k < GA_COUNT; k++) A.a[k] = out->arrays + (size_t)k * (size_t)P;
    memset(out->read_state, 0, (size_t)n);
    if (out->lookahead_lseq) memset(out->lookahead_lseq, 0, sizeof(int32_t) * (size_t)P);
    out->n_snv = 0; out->scan_first = -1; out->scan_last = -1; out->snv_ave_rd = 0;

    /* per-position storage of the first min_snv distinct read names seen on high-quality
     * mismatches (src/GROM.c:6805-6824); hashes stand in for the strings */
    const int NS = p->min_snv;
    uint64_t *nm_hash = (uint64_t *)calloc((size_t)P * NS, sizeof(uint64_t));
    uint8_t *nm_cnt = (uint8_t *)calloc((size_t)P, 1);

    svctx S; memset(&S, 0, sizeof(S));
    S.p = p; S.P = P; S.A = &A; S.W = W;
    for (int k = 0; k < CL_COUNT; k++) {
        S.cw[k] = (int32_t *)calloc((size_t)P, 4); S.crs[k] = (int32_t *)calloc((size_t)P, 4); S.cre[k] = (int32_t *)calloc((size_t)P, 4);
        S.cdist[k] = (double *)calloc((size_t)P, 8);
    }
    S.cmchr[0] = (int32_t *)calloc((size_t)P, 4); S.cmchr[1] = (int32_t *)calloc((size_t)P, 4);
    S.oth = (oslot **)calloc((size_t)P, sizeof(oslot *));
    S.ins_seq = (char **)calloc((size_t)P, sizeof(char *));

    dupkey *dl = (dupkey *)malloc(sizeof(dupkey) * (size_t)p->rmdup_list_len);
    int dl_n = 0, old_pos = -1;
    int *c_type = (int *)malloc(sizeof(int) * 65536), *c_len = (int *)malloc(sizeof(int) * 65536);

    int64_t i0 = 0;                      /* reads before W/4+1 are skipped outright, src/GROM.c:6406, 14859 */
    while (i0 < n && b->pos[i0] < first_pos) i0++;
    int last_lseq_after_clip = 0;        /* cdp_lseq of the last processed read (with H added), for the final scan */

    for (int64_t i = i0; i < n; i++) {
        const int pos = b->pos[i], flag = b->flag[i], mq = b->mapq[i], mtid = b->mtid[i], mpos = b->mpos[i];
        const int tlen = b->tlen[i];
        int lseq = b->l_qseq[i];
        const int add = (mq >= q) ? p->add_factor : 0;                       /* src/GROM.c:5829-5837 */
        last_lseq_after_clip = lseq;
        if (flag & (F_UNMAP | F_DUP)) continue;                              /* src/GROM.c:6419 */
        int keep = 1;
        /* ---- -M duplicate filter, src/GROM.c:6432-6588 */
        if (p->rmdup > 0 && (flag & F_PAIRED) && !(flag & F_MUNMAP)) {
            int sv = dup_svtype(b->tid, mtid, pos, mpos, flag);
            if (sv >= 0) {
                if (pos != old_pos) { dl_n = 0; old_pos = pos; }
                else {
                    for (int k = 0; k < dl_n; k++)
                        if (mpos == dl[k].mpos && mtid == dl[k].mtid && dl[k].lseq == lseq && dl[k].tlen == tlen &&
                            mq >= q && dl[k].svtype == sv) { keep = 0; break; }
                }
                if (keep && dl_n < p->rmdup_list_len) {
                    dl[dl_n].mpos = mpos; dl[dl_n].mtid = mtid; dl[dl_n].lseq = lseq; dl[dl_n].tlen = tlen; dl[dl_n].svtype = sv;
                    dl_n++;
                }
            }
        }
        out->read_state[i] = keep ? 1 : 2;
        if (!keep) continue;

        const uint32_t *cg = b->cigar + b->cigar_off[i];
        const int ncig_all = b->n_cigar[i];
        /* ---- CNV depth, src/GROM.c:6605-6671: whole CIGAR, N/S/I/H/P do not advance */
        {
            int64_t cp = pos;
            for (int k = 0; k < ncig_all; k++) {
                int op = cg[k] & 15; int64_t len = cg[k] >> 4;
                if (op == OP_M || op == OP_EQ || op == OP_X) {
                    if (cp >= 0 && cp + len < P) {
                        int32_t *dst = (mq >= p->rd_min_mapq) ? A.a[GA_RD_RD] : A.a[GA_RD_LOW];
                        for (int64_t x = cp; x < cp + len; x++) { A.a[GA_RD_MQ][x] += mq; dst[x] += 1; }
                    }
                    cp += len;
                } else if (op == OP_D) cp += len;
            }
        }
        /* ---- pileup, src/GROM.c:6740-7059: first max_cigar_ops ops only */
        int ncig = ncig_all > p->max_cigar_ops ? p->max_cigar_ops : ncig_all;
        for (int k = 0; k < ncig; k++) { c_type[k] = cg[k] & 15; c_len[k] = (int)(cg[k] >> 4); }
        {
            int qi = 0, ri = 0;
            const int fwd = !(flag & F_REVERSE);
            for (int k = 0; k < ncig; k++) {
                int op = c_type[k], len = c_len[k];
                if (op == OP_M || op == OP_EQ || op == OP_X) {
                    if (pos >= 0 && pos < P) {
                        int nrun;
                        if ((int64_t)pos + ri + len >= P) nrun = (int)(P - pos);
                        else if (pos + len < 0) nrun = 1;
                        else nrun = len;
                        for (int t = 0; t < nrun; t++) {
                            int64_t rp = (int64_t)pos + ri;
                            if (rp >= P || qi >= b->l_qseq[i]) { qi++; ri++; continue; }   /* reference would read out of bounds here */
                            int code = base_code(b, i, qi);
                            char sc = NT16[code];
                            int qv = b->qual[b->base_off[i] + (uint64_t)qi];
                            char rc = (char)toupper((unsigned char)fasta[rp]);
                            int bi = (sc == 'A') ? 0 : (sc == 'C') ? 1 : (sc == 'G') ? 2 : (sc == 'T') ? 3 : -1;
                            if (mq >= q && qv >= bqmin) {
                                int skip = 0;
                                if (rc != sc) {                                     /* mate-overlap de-dup on mismatches */
                                    uint64_t h = b->qname_hash[i];
                                    uint64_t *slot = nm_hash + (size_t)rp * NS;
                                    int s;
                                    for (s = 0; s < NS; s++) {
                                        if (s >= nm_cnt[rp]) {                      /* empty slot */
                                            if (b->qname_len[i] < p->read_name_len) { slot[s] = h; nm_cnt[rp] = (uint8_t)(s + 1); }
                                            break;
                                        } else if (slot[s] == h) { skip = 1; break; }
                                    }
                                }
                                if (!skip && bi >= 0) {
                                    A.a[GA_SNV_A + bi][rp] += 1;
                                    A.a[GA_BQ][rp] += qv; A.a[GA_BQ_ALL][rp] += qv;
                                    A.a[GA_MQ][rp] += mq; A.a[GA_MQ_ALL][rp] += mq;
                                    A.a[GA_BQ_RC][rp] += 1; A.a[GA_MQ_RC][rp] += 1; A.a[GA_RC_ALL][rp] += 1;
                                    if (fwd) A.a[GA_FS_A + bi][rp] += 1;
                                    if (rc == sc) A.a[GA_PIR_A + bi][rp] += fwd ? qi : lseq - qi;    /* src/GROM.c:6853-6864 */
                                    else          A.a[GA_PIR_A + bi][rp] += qi;                      /* src/GROM.c:6909 */
                                }
                            } else if (bi >= 0) {                                   /* src/GROM.c:6930-6979 */
                                A.a[GA_SNVLOW_A + bi][rp] += 1;
                                A.a[GA_BQ_ALL][rp] += qv; A.a[GA_MQ_ALL][rp] += mq; A.a[GA_RC_ALL][rp] += 1;
                            }
                            qi++; ri++;
                        }
                    }
                } else if (op == OP_S) qi += len;
                else if (op == OP_H) lseq += len;                                   /* src/GROM.c:6997-7000 */
                else if (op == OP_I) qi += len;
                else if (op == OP_D || op == OP_N) ri += len;
            }
        }
        last_lseq_after_clip = lseq;
        /* ---- clips and physical depth, src/GROM.c:7067-7181 */
        int start_adj = 0, end_adj = 0, end_adj_indel = 0;
        if (ncig > 0) {
            if (c_type[0] == OP_S || c_type[0] == OP_H) start_adj = c_len[0];
            if (c_type[ncig - 1] == OP_S || c_type[ncig - 1] == OP_H) end_adj = c_len[ncig - 1];
        }
        for (int k = 0; k < ncig; k++) {
            if (c_type[k] == OP_I) end_adj_indel += c_len[k];
            else if (c_type[k] == OP_D) end_adj_indel -= c_len[k];
        }
        const int paired = (flag & F_PAIRED) != 0, munmap = (flag & F_MUNMAP) != 0, rev = (flag & F_REVERSE) != 0;
        const int same = (b->tid == mtid);
        const int64_t rend = (int64_t)pos - start_adj + lseq - end_adj - end_adj_indel;
#define BUMP(ARR, RD, CRD, X) do { int64_t x_ = (X); if (x_ >= 0 && x_ < P) { A.a[ARR][x_] += add; A.a[RD][x_] += 1; A.a[CRD][x_] += 1; } } while (0)
        if (start_adj >= p->sc_min) {
            int64_t x = (int64_t)pos - 1;
            if (!paired || (!rev && (munmap || (same && mpos > pos)))) BUMP(GA_SC_LEFT, GA_SC_LEFT_RD, GA_SC_RD, x);
            if (paired && !munmap && !same && rev) BUMP(GA_CTX_SC_LEFT, GA_CTX_SC_LEFT_RD, GA_CTX_SC_RD, x);
            if (paired && !munmap && same && rev && abs(tlen) <= p->insert_max && mpos < pos)
                BUMP(GA_INDEL_SC_LEFT, GA_INDEL_SC_LEFT_RD, GA_INDEL_SC_RD, x);
        }
        if (end_adj >= p->sc_min) {
            if (!paired || (rev && (munmap || (same && mpos < pos)))) BUMP(GA_SC_RIGHT, GA_SC_RIGHT_RD, GA_SC_RD, rend);
            if (paired && !munmap && !same && !rev) BUMP(GA_CTX_SC_RIGHT, GA_CTX_SC_RIGHT_RD, GA_CTX_SC_RD, rend);
            if (paired && !munmap && same && !rev && abs(tlen) <= p->insert_max && mpos > pos)
                BUMP(GA_INDEL_SC_RIGHT, GA_INDEL_SC_RIGHT_RD, GA_INDEL_SC_RD, rend);
        }
#undef BUMP
        for (int64_t x = pos; x < rend; x++) if (x >= 0 && x < P) A.a[GA_RD][x] += 1;
        /* ---- small indels, split reads, pair ranges (grom_oracle_sv.c) */
        {
            svread r;
            r.tid = b->tid; r.pos = pos; r.mpos = mpos; r.mtid = mtid; r.tlen = tlen; r.flag = flag; r.mapq = mq; r.add = add;
            r.lseq = lseq; r.start_adj = start_adj; r.end_adj = end_adj; r.end_adj_indel = end_adj_indel;
            r.cigar = cg; r.n_cigar = ncig;
            r.sa_pos = b->sa_pos[i]; r.sa_strand = b->sa_strand[i]; r.sa_mapq = b->sa_mapq[i]; r.sa_same = b->sa_same_chr[i];
            r.sa_start_adj = b->sa_start_adj[i]; r.sa_end_adj = b->sa_end_adj[i]; r.sa_end_adj_indel = b->sa_end_adj_indel[i];
            if (!p->splitread) r.sa_pos = -1;
            /* window geometry when this read is applied: scan position = max(pos - ins_max, W/4+1), window index advances once
             * per loop iteration incl. the skipped leading reads and wraps 3W/4 -> W/4 (src/GROM.c:5845-5847, 6317, 6408-6411) */
            int64_t pproc = (int64_t)pos - (int64_t)p->overlap_mult * p->insert_max; if (pproc < first_pos) pproc = first_pos;
            int64_t idx = W / 4 + ((i0 + 2 + (pproc - first_pos)) % (W / 2));
            r.win_lo = pproc - idx; r.read_index = i; r.batch = b;
            sv_evidence_read(&S, &r);
        }
    }

    /* ---- scanned range and look-ahead read length (src/GROM.c:6406-6411, 11075-11086) */
    if (i0 < n) {
        int scan_first = first_pos;
        int64_t scan_last = (int64_t)b->pos[n - 1] - (int64_t)p->overlap_mult * p->insert_max;
        if (scan_last < scan_first) scan_last = scan_first;
        if (scan_last >= P) scan_last = P - 1;
        out->scan_first = scan_first; out->scan_last = (int32_t)scan_last;
        int32_t *la = out->lookahead_lseq ? out->lookahead_lseq : (int32_t *)calloc((size_t)P, sizeof(int32_t));
        {
            int64_t j = i0;
            for (int64_t x = scan_first; x <= scan_last; x++) {
                while (j < n && (int64_t)b->pos[j] - (int64_t)p->overlap_mult * p->insert_max <= x) j++;
                la[x] = (j < n) ? b->l_qseq[j] : last_lseq_after_clip;
            }
        }
        /* ---- SNV gate per scanned position, src/GROM.c:11096-11199 */
        for (int64_t x = scan_first; x <= scan_last; x++) {
            if (A.a[GA_RD][x] + A.a[GA_INDEL_SC_RD][x] <= 0) continue;
            char fc = fasta[x];
            if (fc == 'N' || fc == 'n') continue;
            int total = 0, cnt[4];
            for (int k = 0; k < 4; k++) { cnt[k] = A.a[GA_SNV_A + k][x]; total += cnt[k]; }
            int have = 0; grom_snv_cand c; memset(&c, 0, sizeof(c));
            for (int k = 0; k < 4; k++) {
                double ratio = (float)cnt[k] / (float)total;
                double pr, hz;
                if (total > p->max_trials) {
                    int col = cnt[k] * p->max_trials / total;
                    pr = mq_tbl[(size_t)p->max_trials * TD + col]; hz = hez_tbl[(size_t)p->max_trials * TD + col];
                } else { pr = mq_tbl[(size_t)total * TD + cnt[k]]; hz = hez_tbl[(size_t)total * TD + cnt[k]]; }
                if (toupper((unsigned char)fc) != "ACGT"[k] && ratio >= p->min_snv_ratio && cnt[k] >= p->min_snv &&
                    (double)A.a[GA_BQ_ALL][x] / (double)A.a[GA_RC_ALL][x] >= p->min_ave_bq) {
                    if (have) { if (ratio > c.ratio) { c.ratio = ratio; c.base = k; c.pr = pr; c.hez = hz; } }
                    else {
                        have = 1; c.pos = (int32_t)x; c.base = k; c.ratio = ratio; c.pr = pr; c.hez = hz;
                        for (int t = 0; t < GA_PILEUP_COUNT; t++) c.v[t] = A.a[t][x];
                    }
                }
            }
            if (have) { if (out->n_snv < out->snv_cap) out->snv[out->n_snv] = c; out->n_snv++; }
        }
        /* ---- small-insertion gate per scanned position, src/GROM.c:11329-11453 */
        out->n_ins = 0;
        for (int64_t x = scan_first; x <= scan_last; x++) {
            if (A.a[GA_RD][x] + A.a[GA_INDEL_SC_RD][x] <= 0) continue;
            int rdt = 0;
            for (int k = 0; k < 8; k++) rdt += A.a[GA_SNV_A + k][x];
            int it = A.a[GA_INDEL_I][x];
            const int af = p->add_factor;
            if (it / af > rdt) it = rdt * af;
            if (!(it / af >= p->min_disc) || rdt > p->max_trials) continue;
            double pr = mq_tbl[(size_t)rdt * TD + it / af], hz;
            if ((it + A.a[GA_INDEL_SC_LEFT][x]) / af < rdt) {
                hz = hez_tbl[(size_t)rdt * TD + (it + A.a[GA_INDEL_SC_LEFT][x]) / af];
                if ((it + A.a[GA_INDEL_SC_RIGHT][x]) / af < rdt) {
                    double h2 = hez_tbl[(size_t)rdt * TD + (it + A.a[GA_INDEL_SC_RIGHT][x]) / af];
                    if (h2 > hz) hz = h2;
                } else hz = hez_tbl[(size_t)rdt * TD + rdt];
            } else hz = hez_tbl[(size_t)rdt * TD + rdt];
            if (!(pr <= p->pval_threshold1)) continue;
            if (out->ins && out->n_ins < out->ins_cap) {
                grom_ins_cand *c = &out->ins[out->n_ins];
                memset(c, 0, sizeof(*c));
                c->pos = (int32_t)x; c->dist = A.a[GA_INDEL_IDIST][x]; c->pr = pr; c->hez = hz; c->conc = A.a[GA_CONC][x]; c->weight = it; c->rd = rdt;
                c->sc = (x + 1 < P ? A.a[GA_SC_LEFT][x + 1] : 0) + A.a[GA_SC_RIGHT][x];
                int ol = 0; if (S.oth[x]) while (ol < p->other_len && S.oth[x][ol].type != OTHER_EMPTY) ol++;
                c->other_len = ol;
                if (c->dist <= p->indel_i_seq_len && S.ins_seq[x])
                    for (int k = 0; k < c->dist && k < 50; k++) c->seq[k] = S.ins_seq[x][k];
            }
            out->n_ins++;
        }
        /* ---- small-deletion start / end gates per scanned position, src/GROM.c:11454-11745 (events; the pairing state
         * machine runs on the host over this compact stream) */
        out->n_del = 0;
        for (int64_t x = scan_first; x <= scan_last; x++) {
            if (A.a[GA_RD][x] + A.a[GA_INDEL_SC_RD][x] <= 0) continue;
            int base = 0;
            for (int k = 0; k < 8; k++) base += A.a[GA_SNV_A + k][x];
            const int af = p->add_factor;
            int ol = 0; if (S.oth[x]) while (ol < p->other_len && S.oth[x][ol].type != OTHER_EMPTY) ol++;
            for (int kind = 0; kind < 2; kind++) {
                const int wt = A.a[kind ? GA_INDEL_D_R : GA_INDEL_D_F][x];
                const int rdt = wt / af + base;
                if (!(wt / af >= p->min_disc) || rdt > p->max_trials) continue;
                const int scv = A.a[kind ? GA_INDEL_SC_LEFT : GA_INDEL_SC_RIGHT][x];
                const double pr = mq_tbl[(size_t)rdt * TD + wt / af];
                const double hz = ((wt + scv) / af < rdt) ? hez_tbl[(size_t)rdt * TD + (wt + scv) / af] : hez_tbl[(size_t)rdt * TD + rdt];
                if (!(pr <= p->pval_threshold1)) continue;
                if (out->del_ev && out->n_del < out->del_cap) {
                    grom_del_event *e = &out->del_ev[out->n_del];
                    e->pos = (int32_t)x; e->kind = kind; e->pr = pr; e->hez = hz; e->conc = A.a[GA_CONC][x]; e->weight = wt; e->rd = rdt;
                    e->sc = A.a[kind ? GA_SC_LEFT : GA_SC_RIGHT][x]; e->other_len = ol; e->rdist = A.a[GA_INDEL_D_RDIST][x];
                }
                out->n_del++;
            }
        }
        /* ---- structural-variant gates per scanned position, src/GROM.c:11750-13541 (events in the reference's evaluation order:
         * insertion left / right, ctx_f, ctx_r, dup_r, dup_f, del_f, del_r, inv_f1, inv_f2, inv_r1, inv_r2) */
        out->n_sv = 0;
        for (int64_t x = scan_first; x <= scan_last; x++) {
            const int af = p->add_factor, rd = A.a[GA_RD][x], mt = p->max_trials;
            int ol = -1;
#define SV_EMIT(CLS, BIN, HEZ, DIST, WGT, RS, RE, MCHR) do {                                                                      \
                if (ol < 0) { ol = 0; if (S.oth[x]) while (ol < p->other_len && S.oth[x][ol].type != OTHER_EMPTY) ol++; }            \
                if (out->sv_ev && out->n_sv < out->sv_cap) {                                                                        \
                    grom_sv_event *e = &out->sv_ev[out->n_sv]; memset(e, 0, sizeof(*e));                                            \
                    e->pos = (int32_t)x; e->cls = (CLS); e->binom = (BIN); e->hez = (HEZ); e->dist = (DIST); e->weight = (WGT);     \
                    e->rd = rd; e->conc = A.a[GA_CONC][x]; e->read_start = (RS); e->read_end = (RE); e->other_len = ol; e->mchr = (MCHR); \
                }                                                                                                                   \
                out->n_sv++; } while (0)
            if (rd + A.a[GA_SC_RD][x] > 0) {
                /* insertions: soft clips + short-insert pairs (+ mate-unmapped reads in the table column), src/GROM.c:11750-11961 */
                for (int side = 0; side < 2; side++) {
                    const int sc = A.a[side ? GA_SC_RIGHT : GA_SC_LEFT][x], scrd = rd + A.a[side ? GA_SC_RIGHT_RD : GA_SC_LEFT_RD][x];
                    const int mu = A.a[side ? GA_MUNMAPPED_F : GA_MUNMAPPED_R][x], ins = A.a[GA_INS][x];
                    if (!((sc + ins) / af >= p->min_disc) || scrd > mt) continue;
                    const double bin = ((mu + sc + ins) / af < scrd) ? mq_tbl[(size_t)scrd * TD + (mu + sc + ins) / af] : mq_tbl[(size_t)scrd * TD + scrd];
                    if (bin <= p->pval_insertion1) SV_EMIT(GROM_SV_INS_L + side, bin, 2.0, 0.0, ins, 0, 0, 0);
                }
            }
            if (rd > 0) {
                static const int order[10] = { CL_CTX_F, CL_CTX_R, CL_DUP_R, CL_DUP_F, CL_DEL_F, CL_DEL_R, CL_INV_F1, CL_INV_F2, CL_INV_R1, CL_INV_R2 };
                for (int oi = 0; oi < 10; oi++) {
                    const int c = order[oi], fwd = (c == CL_CTX_F || c == CL_DUP_F || c == CL_DEL_F || c == CL_INV_F1 || c == CL_INV_F2);
                    const int w = S.cw[c][x];
                    if (!(w / af >= p->min_disc)) continue;
                    if (fwd ? !((int)x - S.cre[c][x] < p->insert_mean) : !(S.crs[c][x] + la[x] - (int)x < p->insert_mean)) continue;
                    const int side = fwd ? A.a[GA_SC_RIGHT][x] + A.a[GA_MUNMAPPED_F][x] : A.a[GA_SC_LEFT][x] + A.a[GA_MUNMAPPED_R][x];
                    double bin, hz = 2.0;
                    if (rd > mt) {
                        bin = mq_tbl[(size_t)mt * TD + w * mt / (af * rd)];
                        if ((float)side / (float)w <= p->max_evidence_ratio)
                            hz = ((w + side) / af < rd) ? hez_tbl[(size_t)mt * TD + (w + side) * mt / (af * rd)] : hez_tbl[(size_t)mt * TD + mt];
                    } else {
                        bin = mq_tbl[(size_t)rd * TD + w / af];
                        /* ctx_r tests the ctx_f ratio here (src/GROM.c:12074) */
                        const float ratio = c == CL_CTX_R ? (float)(A.a[GA_SC_RIGHT][x] + A.a[GA_MUNMAPPED_F][x]) / (float)S.cw[CL_CTX_F][x] : (float)side / (float)w;
                        if (ratio <= p->max_evidence_ratio)
                            hz = ((w + side) / af < rd) ? hez_tbl[(size_t)rd * TD + (w + side) / af] : hez_tbl[(size_t)rd * TD + rd];
                    }
                    if (bin <= p->pval_threshold1) {
                        SV_EMIT(c, bin, hz, S.cdist[c][x], w, S.crs[c][x], S.cre[c][x], c >= CL_CTX_F ? S.cmchr[c - CL_CTX_F][x] : 0);
                        if (c >= CL_INV_F1 && c <= CL_INV_R2 && out->sv_ev && out->n_sv <= out->sv_cap) {
                            /* depth around the breakpoint, compared between both ends at emission (src/GROM.c:15921-15934) */
                            long sum = 0;
                            for (int64_t y = S.crs[c][x]; y < (int64_t)S.cre[c][x] + p->lseq; y++) if (y >= 0 && y < P) sum += A.a[GA_RD_RD][y] + A.a[GA_RD_LOW][y];
                            out->sv_ev[out->n_sv - 1].reserved = (int32_t)sum;
                        }
                    }
                }
            }
#undef SV_EMIT
        }
        if (!out->lookahead_lseq) free(la);
        /* ---- mean depth for the emission filter, src/GROM.c:15035-15043.  Upper bound = position of
         * window index 0 when the loop ends: (scan_last+1) - index(scan_last), where the window index
         * advances once per loop iteration (including one per skipped leading read) and wraps from
         * 3W/4 back to W/4 (src/GROM.c:5845-5847, 6317). */
        {
            int64_t s = i0;                                   /* leading reads skipped before W/4+1 */
            int64_t idx = W / 4 + ((s + 2 + (scan_last - first_pos)) % (W / 2));
            int64_t bound = scan_last + 1 - idx;
            long tot = 0, cntb = 0;
            for (int64_t x = 0; x < bound && x < P; x++)
                if (fasta[x] != 'N' && fasta[x] != 'n') { tot += (long)A.a[GA_RD_RD][x] + (long)A.a[GA_RD_LOW][x]; cntb++; }
            out->snv_ave_rd = (double)tot / (double)cntb;
        }
    }
    oracle_gc_prepass(p, fasta, P, A.a[GA_GC], A.a[GA_ACGT]);
    for (int k = 0; k < CL_COUNT; k++) {
        if (out->cl_w) memcpy(out->cl_w + (size_t)k * P, S.cw[k], (size_t)P * 4);
        if (out->cl_rs) memcpy(out->cl_rs + (size_t)k * P, S.crs[k], (size_t)P * 4);
        if (out->cl_re) memcpy(out->cl_re + (size_t)k * P, S.cre[k], (size_t)P * 4);
        if (out->cl_dist) memcpy(out->cl_dist + (size_t)k * P, S.cdist[k], (size_t)P * 8);
        free(S.cw[k]); free(S.crs[k]); free(S.cre[k]); free(S.cdist[k]);
    }
    for (int k = 0; k < 2; k++) { if (out->cl_mchr) memcpy(out->cl_mchr + (size_t)k * P, S.cmchr[k], (size_t)P * 4); free(S.cmchr[k]); }
    for (int64_t x = 0; x < P; x++) {
        int ol = 0;
        if (S.oth[x]) { while (ol < p->other_len && S.oth[x][ol].type != OTHER_EMPTY) ol++; free(S.oth[x]); }
        if (out->other_len) out->other_len[x] = ol;
    }
    free(S.oth);
    for (int64_t x = 0; x < P; x++) free(S.ins_seq[x]);
    free(S.ins_seq);
    free(nm_hash); free(nm_cnt); free(dl); free(c_type); free(c_len);
    return 0;
}

/* src/GROM.c:15046-15095 */
int64_t oracle_format_snv_vcf(const grom_params *p, const char *chr_name, const char *fasta,
                              const grom_snv_cand *snv, int64_t n_snv, double ave_rd, char *buf, int64_t cap)
{
    int64_t w = 0;
    char gt[256];
    for (int64_t i = 0; i < n_snv; i++) {
        const grom_snv_cand *c = &snv[i];
        if (!(c->v[GA_RC_ALL] <= round(p->snv_rd_min_factor * ave_rd) || c->ratio >= p->high_cov_min_snv_ratio)) continue;
        int cn = (int)round(c->ratio * p->ploidy);
        if (cn == 0) cn = 1;
        for (int k = 0; k < p->ploidy; k++) {
            gt[2 * k] = k < cn ? '1' : '0';
            gt[2 * k + 1] = k < p->ploidy - 1 ? '/' : '\0';
        }
        int nb = c->v[GA_SNV_A + c->base];
        if (cap - w < 512) return -1;
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

/* src/GROM.c:16253-16340.  Bug-compatible homopolymer rule: the second run is counted against the character
 * code fasta[pos] + 1 (src/GROM.c:16284), not against the next base.  END prints list_end + 1 = 0 (the field is
 * initialised to -1 and never set, src/GROM.c:5545); ECO / EOT are uninitialised in the reference and printed as 0 here. */
int64_t oracle_format_ins_vcf(const grom_params *p, const char *chr_name, const char *fasta, int64_t chr_len,
                              const grom_ins_cand *ins, int64_t n_ins, char *buf, int64_t cap)
{
    int64_t w = 0;
    for (int64_t i = 0; i < n_ins; i++) {
        const grom_ins_cand *c = &ins[i];
        if (!(c->pr <= p->pval_threshold && (double)c->weight / (double)c->rd > p->min_indel_ratio * (double)p->add_factor)) continue;
        int hp = 1;
        char hc = fasta[c->pos];
        for (int k = 1; k < 20; k++) { if (c->pos - k >= 0 && hc == fasta[c->pos - k]) hp++; else break; }
        int hp2 = 1;
        if (fasta[c->pos] + 1 < chr_len) {
            hc = (char)(fasta[c->pos] + 1);
            for (int k = 1; k < 20; k++) { if (c->pos + k + 1 < chr_len && hc == fasta[c->pos + k + 1]) hp2++; else break; }
        }
        if (hp2 > hp) hp = hp2;
        if (hp > 10) continue;                                   /* g_max_homopolymer */
        char alt[64];
        if (c->dist <= p->indel_i_seq_len) { memcpy(alt, c->seq, (size_t)c->dist); alt[c->dist] = 0; } else strcpy(alt, "<INS>");
        if (cap - w < 512) return -1;
        w += snprintf(buf + w, (size_t)(cap - w), "%s\t%d\t.\t.\t%s\t.\t.\tEND=%d\tSPR:SEV:SRD:SCO:ECO:SOT:EOT:SSC:HP\t%e:%.1f:%d:%d:%d:%d:%d:%d:%d\n",
                      chr_name, c->pos + 1, alt, 0, c->pr, (double)c->weight / (double)p->add_factor, c->rd, c->conc, 0, c->other_len, 0, c->sc, hp);
    }
    return w;
}

/* GC / ACGT percentage of the triangular window of half-width insert_mean centred on each position
 * (src/GROM.c:1766-1859).  The reference updates the weighted counts incrementally; the closed form is
 * count(r) = sum_{|d| < M} (M - |d|) * is(r + d), M = insert_mean, defined for r in [M-1, P-(2M-1));
 * value = 100 * count / M^2 (integer division).  Positions outside that range are never written by the
 * reference (malloc'ed memory); they are reported as 0 here. */
void oracle_gc_prepass(const grom_params *p, const char *fasta, int64_t chr_len, int32_t *gc, int32_t *acgt)
{
    const int64_t M = p->insert_mean, W1 = 2 * M - 1, total = M * M;
    memset(gc, 0, sizeof(int32_t) * (size_t)chr_len);
    memset(acgt, 0, sizeof(int32_t) * (size_t)chr_len);
    if (chr_len < W1 + M) return;
    /* S1g[i] = number of G/C in fasta[0..i), likewise for A/C/G/T */
    int64_t *s1g = (int64_t *)calloc((size_t)chr_len + 1, sizeof(int64_t)), *s1a = (int64_t *)calloc((size_t)chr_len + 1, sizeof(int64_t));
    for (int64_t i = 0; i < chr_len; i++) {
        char c = fasta[i];
        int g = (c == 'C' || c == 'G' || c == 'c' || c == 'g');
        int a = g || (c == 'A' || c == 'T' || c == 'a' || c == 't');
        s1g[i + 1] = s1g[i] + g; s1a[i + 1] = s1a[i] + a;
    }
    for (int64_t r = M - 1; r < chr_len - W1; r++) {
        /* sum_{j=0}^{M-1} (S1[r+j+1] - S1[r+j+1-M]) */
        int64_t cg = 0, ca = 0;
        for (int64_t j = 0; j < M; j++) { cg += s1g[r + j + 1] - s1g[r + j + 1 - M]; ca += s1a[r + j + 1] - s1a[r + j + 1 - M]; }
        gc[r] = (int32_t)(100 * cg / total); acgt[r] = (int32_t)(100 * ca / total);
    }
    free(s1g); free(s1a);
}
