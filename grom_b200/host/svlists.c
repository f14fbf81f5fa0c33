/* svlists.c -- host stage of the structural-variant scan: turns the per-position gate events the CUDA library returns
 * (gromgpu_chr_result: grom_sv_event, scan order) into the reference's candidate lists as they stand at the end of the
 * per-position scan, before the list -> list2 merge:
 *
 *   cdp_ctx_f_list / cdp_ctx_r_list       plain appends                                  src/GROM.c:12007-12042, 12087-12122
 *   cdp_dup_list / del / inv_f / inv_r    start events append (dup_r, del_f, inv_f1, inv_r1); end events (dup_f, del_r, inv_f2,
 *                                         inv_r2) look the matching starts up by expected distance and position window and
 *                                         take over the end side when they are better     src/GROM.c:12168-12469, 12516-12846,
 *                                                                                         12888-13197, 13237-13541
 *   cdp_ins_list                          start / end sides grouped within g_sc_range     src/GROM.c:11775-11957
 *
 * Sequential by nature (every end event reads the list built so far), tiny (one event per candidate position).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "gromhost.h"

#define SC_RANGE 35             /* g_sc_range */
#define RANGE_MULT 0.75         /* g_range_mult */
#define SV_LIST_LEN 1000000     /* g_sv_list_len */

/* position in the reference's evaluation order within one scanned position */
static int order_key(int cls)
{
    static const int key[GROM_SV_CLASSES] = { 6, 7, 5, 4, 8, 10, 9, 11, 2, 3, 0, 1 };
    return cls >= 0 && cls < GROM_SV_CLASSES ? key[cls] : 99;
}
static int cmp_event(const void *a, const void *b)
{
    const grom_sv_event *x = a, *y = b;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    return order_key(x->cls) - order_key(y->cls);
}

static void side_from_event(grom_sv_side *s, const grom_sv_event *e)
{
    s->pos = e->pos; s->weight = e->weight; s->rd = e->rd; s->conc = e->conc; s->read_start = e->read_start; s->read_end = e->read_end;
    s->other_len = e->other_len; s->reserved = e->reserved; s->binom = e->binom; s->hez = e->hez;
}

/* The reference's in-line list search (src/GROM.c:12271-12346, also bisect_list 20360-20434) over the start positions collected so
 * far.  `pos_of(k)` for k == n is the not-yet-filled slot, which the reference initialises to -1 (src/GROM.c:5528-5546); the
 * interpolated first guess divides by it, so in practice the guess falls outside the list and the whole list is bisected -- kept
 * literally because a target of exactly 0 does produce an in-range guess.  type 0: first index >= target side; type 1: last <=. */
static int list_search(const grom_sv_pair *l, int n, int target, int type)
{
#define POS_OF(k) ((k) < n ? l[k].start.pos : -1)
    int range = n / 64;
    if (range < 4) range = 4; else if (range > 64) range = 64;
    const double guess = round((double)target * (double)n / (double)POS_OF(n));
    int lo = (int)(guess - range), hi = (int)(guess + range);
    if (lo < 0 || lo >= n) lo = 0;
    else if (POS_OF(lo) > target) { hi = lo; lo = 0; }
    if (hi > n || hi < 0) hi = n;
    else if (POS_OF(hi) < target) { lo = hi; hi = n; }
    int i = lo + (hi - lo) / 2;
    for (;;) {
        if (target < POS_OF(i)) { hi = i; i = lo + (i - lo) / 2; if (hi == i) break; }
        else if (target > POS_OF(i)) { lo = i; i = i + (hi - i) / 2; if (lo == i) break; }
        else break;
    }
    if (type == 0 && target > POS_OF(i) && i < n) i++;
    else if (type == 1 && target < POS_OF(i) && i > 0) i--;
    return i;
#undef POS_OF
}

typedef struct { grom_sv_pair *v; int64_t n, cap; } pair_list;
typedef struct { grom_sv_event *v; int64_t n, cap; } event_list;

static grom_sv_pair *pair_push(pair_list *l)
{
    if (l->n == l->cap) { l->cap = l->cap ? 2 * l->cap : 256; l->v = realloc(l->v, (size_t)l->cap * sizeof(grom_sv_pair)); }
    grom_sv_pair *p = &l->v[l->n++];
    memset(p, 0, sizeof(*p));
    p->start.pos = p->end.pos = -1;
    return p;
}

int gromhost_sv_lists(const grom_params *p, const grom_sv_event *events, int64_t n_events, gromhost_sv_lists_t *out)
{
    memset(out, 0, sizeof(*out));
    grom_sv_event *ev = malloc((size_t)(n_events > 0 ? n_events : 1) * sizeof(grom_sv_event));
    if (!ev) return -1;
    memcpy(ev, events, (size_t)n_events * sizeof(grom_sv_event));
    qsort(ev, (size_t)n_events, sizeof(grom_sv_event), cmp_event);

    pair_list pl[4] = {{0}};                    /* 0 dup, 1 del, 2 inv_f, 3 inv_r */
    pair_list ins = {0};
    event_list ctx[2] = {{0}};
    const double span = RANGE_MULT * (double)(p->insert_max - p->insert_min);
    for (int64_t k = 0; k < n_events; k++) {
        const grom_sv_event *e = &ev[k];
        int li = -1, is_end = 0, tie_ge = 0, base = 0;
        double centre = 0;
        switch (e->cls) {
        case GROM_SV_CTX_F: case GROM_SV_CTX_R: {
            event_list *c = &ctx[e->cls - GROM_SV_CTX_F];
            if (c->n < SV_LIST_LEN - 1) {
                if (c->n == c->cap) { c->cap = c->cap ? 2 * c->cap : 256; c->v = realloc(c->v, (size_t)c->cap * sizeof(grom_sv_event)); }
                c->v[c->n++] = *e;
            }
            continue;
        }
        case GROM_SV_INS_L: case GROM_SV_INS_R: {
            /* one entry collects the best left and the best right gate within g_sc_range of each other */
            const int right = e->cls == GROM_SV_INS_R;
            grom_sv_pair *cur = ins.n ? &ins.v[ins.n - 1] : NULL;
            grom_sv_side *side;
            if (!cur) side = right ? &pair_push(&ins)->end : &pair_push(&ins)->start;
            else if ((cur->start.pos != -1 && e->pos - cur->start.pos > SC_RANGE) || (cur->end.pos != -1 && e->pos - cur->end.pos > SC_RANGE)) {
                if (ins.n - 1 >= SV_LIST_LEN - 1) continue;
                cur = pair_push(&ins);
                side = right ? &cur->end : &cur->start;
            } else {
                side = right ? &cur->end : &cur->start;
                if (!(side->pos == -1 || e->binom < side->binom)) continue;
            }
            side_from_event(side, e);
            continue;
        }
        case GROM_SV_DUP_R: li = 0; break;
        case GROM_SV_DEL_F: li = 1; break;
        case GROM_SV_INV_F1: li = 2; break;
        case GROM_SV_INV_R1: li = 3; break;
        case GROM_SV_DUP_F: li = 0; is_end = 1; centre = e->dist + 2 * p->lseq; base = e->pos - p->insert_mean + 2 * p->lseq; break;
        case GROM_SV_DEL_R: li = 1; is_end = 1; tie_ge = 1; centre = e->dist; base = e->pos + p->insert_mean; break;
        case GROM_SV_INV_F2: li = 2; is_end = 1; centre = e->dist + p->lseq; base = e->pos + p->lseq; break;
        case GROM_SV_INV_R2: li = 3; is_end = 1; centre = e->dist + p->lseq; base = e->pos + p->lseq; break;
        default: continue;
        }
        pair_list *L = &pl[li];
        if (!is_end) {
            if (L->n < SV_LIST_LEN - 1) { grom_sv_pair *q = pair_push(L); side_from_event(&q->start, e); q->dist = e->dist; }
            continue;
        }
        /* expected distance window of the partner and where its start gate must have been */
        const int dmin = (int)((centre - span) + 0.5), dmax = (int)((centre + span) + 0.5);
        int a0 = list_search(L->v, (int)L->n, base - dmin, 0), a1 = list_search(L->v, (int)L->n, base - dmax, 1);
        if (a1 < a0) { const int t = a1; a1 = a0; a0 = t; }
        const int p_lo = base - dmax, p_hi = base - dmin;
        for (int a = a0; a < a1; a++) {
            grom_sv_pair *q = &L->v[a];
            if (!(q->dist >= dmin && q->dist <= dmax && q->start.pos >= p_lo && q->start.pos <= p_hi)) continue;
            const int better = q->end.pos == -1 || (q->end.binom > e->binom && e->weight >= q->end.weight) ||
                               (q->end.binom == e->binom && (tie_ge ? e->weight >= q->end.weight : e->weight > q->end.weight));
            if (better) side_from_event(&q->end, e);
        }
    }
    free(ev);
    out->dup = pl[0].v; out->n_dup = pl[0].n; out->del = pl[1].v; out->n_del = pl[1].n;
    out->inv_f = pl[2].v; out->n_inv_f = pl[2].n; out->inv_r = pl[3].v; out->n_inv_r = pl[3].n;
    out->ins = ins.v; out->n_ins = ins.n;
    out->ctx_f = ctx[0].v; out->n_ctx_f = ctx[0].n; out->ctx_r = ctx[1].v; out->n_ctx_r = ctx[1].n;
    return 0;
}

void gromhost_sv_lists_free(gromhost_sv_lists_t *l)
{
    free(l->dup); free(l->del); free(l->inv_f); free(l->inv_r); free(l->ins); free(l->ctx_f); free(l->ctx_r);
    memset(l, 0, sizeof(*l));
}
