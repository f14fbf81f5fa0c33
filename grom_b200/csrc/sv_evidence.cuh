// sv_evidence.cuh -- order-dependent per-position evidence: CIGAR indel slots, split-read breakpoints and the ten
// discordant-pair breakpoint clusters with their shared "other" slots (reference src/GROM.c:7187-10953).
//
// B200 formulation.  The reference applies each read to a window of positions as it streams the BAM; the result at a
// position depends on the ORDER of the reads that touch it (first-seen primary slot, running-mean distance with an
// int-truncated join test, swap-on-overtake with 50 typed side slots).  Here:
//   1. k_sv_items   one thread per applied read evaluates the reference's class logic once and emits (a) order-free range
//                   adds (rd / conc / ins / munmapped) as +/- difference pairs that a prefix scan finishes, and (b) a
//                   compact list of order-dependent ITEMS (range, class, value, weight rule), in BAM order (count -> scan -> fill);
//   2. k_sv_tiles   per tile of 128 positions, the slice of the item list that can reach it (two binary searches);
//   3. k_sv_apply   one thread per position folds the items that cover it, in list order = BAM order, with the whole
//                   cluster state of the tile in shared memory; only tiles that are touched are written.
// Included by gromgpu.cu (one translation unit, shares c_prm and DevReads).
#pragma once

#define SV_T 128                      // positions per tile / threads per CTA in k_sv_apply
#define SV_NCL 10                     // breakpoint cluster classes
#define SV_NCLS 13                    // + indel_i, indel_d_f, indel_d_r
#define SV_OTHER 50                   // side slots per position (g_other_len)

enum { CL_DEL_F = 0, CL_DEL_R, CL_DUP_F, CL_DUP_R, CL_INV_F1, CL_INV_R1, CL_INV_F2, CL_INV_R2, CL_CTX_F, CL_CTX_R, CL_INDEL_I, CL_INDEL_D_F, CL_INDEL_D_R };
enum { RS_SET_RE = 0, RS_MINMAX = 1, RS_MAX_ONLY = 2 };
enum { AN_FULL = 0, AN_FWD = 1, AN_BWD = 2 };     // which end of a pair range carries the full weight when the read is clipped there

struct __align__(16) SvItem {
    int lo, hi;          // positions [lo, hi)
    int x;               // clustered value (every value the reference clusters is an integer: tlen, gap + mean, +/- mate position, indel length)
    int v;               // read position recorded in read_start / read_end
    int mchr;            // mate contig (translocations)
    int tol;             // join tolerance before the (1 + 1/w) factor
    uint32_t meta;       // [3:0] class  [5:4] RS_*  [7:6] AN_*  [8] clipped at the far end  [9] ctx  [10] ctx mate reverse  [11] split dup_f quirk
    int add;             // cdp_add of the read (6, or 0 below -q)
};
#define SVM_CLIPPED 0x100u
#define SVM_CTX     0x200u
#define SVM_NEG     0x400u
#define SVM_DUPQ    0x800u

struct __align__(8) SvOther { int w, type, mchr, rs, re, pad; double dist; };   // 32 bytes

struct SvDev {
    // outputs (dense, [class][Ppad]); zero-filled before the run, written only for touched tiles
    int32_t *cl_w, *cl_rs, *cl_re, *cl_mchr, *other_len, *ins_src;    // ins_src: [3][Ppad] read index, query offset, length of the stored inserted bases
    double *cl_dist;
    // other-slot pool
    SvOther *pool; int pool_cap; int *pool_used; int *err;
    int2 *ins_pos; int ins_pos_cap; int *n_ins_pos;        // compacted (position, slot mask) whose insertion / deletion-start / deletion-end slot weight reaches min_disc
};

struct SvReadArrays {                 // per-read SA fields (include/grom_reads.h)
    const int32_t *sa_pos, *sa_start_adj, *sa_end_adj, *sa_end_adj_indel;
    const uint8_t *sa_strand, *sa_same; const int16_t *sa_mapq;
};

// ---- per-read evaluation ------------------------------------------------------------------------------------------
template <bool EMIT>
struct SvEmitter {
    int32_t *arrays; int64_t P, Ppad; SvItem *out; int n;      // n = items emitted / counted so far
    int64_t room;                                              // items the buffer still holds from `out` on (a pass that runs out is repeated)
    int max_fwd, max_bwd, pos;
    __device__ __forceinline__ void diff(int ga, int64_t lo, int64_t hi, int val)
    {
        if (!EMIT) return;
        if (lo < 0) lo = 0;
        if (hi > P) hi = P;
        if (lo >= hi || val == 0) return;
        atomicAdd(arrays + (int64_t)ga * Ppad + lo, val);
        if (hi < P) atomicAdd(arrays + (int64_t)ga * Ppad + hi, -val);
    }
    __device__ __forceinline__ void point(int ga, int64_t x, int val) { if (EMIT && x >= 0 && x < P) atomicAdd(arrays + (int64_t)ga * Ppad + x, val); }
    __device__ __forceinline__ void item(int64_t lo, int64_t hi, int cls, int x, int v, int mode, int anchor, bool clipped, int tol, int add,
                                         uint32_t extra = 0, int mchr = 0)
    {
        if (lo < 0) lo = 0;               // positions below 0 / beyond P do not exist; anchors are unaffected because the
        if (hi > P) hi = P;               // reference's own ranges never cross the contig ends with the anchor outside
        if (lo >= hi) return;
        if (EMIT) {
            SvItem it; it.lo = (int)lo; it.hi = (int)hi; it.x = x; it.v = v; it.mchr = mchr; it.tol = tol; it.add = add;
            it.meta = (uint32_t)cls | ((uint32_t)mode << 4) | ((uint32_t)anchor << 6) | (clipped ? SVM_CLIPPED : 0u) | extra;
            if (n < room) out[n] = it;
            max_fwd = max(max_fwd, (int)hi - pos); max_bwd = max(max_bwd, pos - (int)lo);
        }
        n++;
    }
};

template <bool EMIT>
__device__ void sv_read_items(const DevReads &R, const SvReadArrays &SA, int64_t i, int tid, int64_t n_leading, SvEmitter<EMIT> &em)
{
    const int pos = R.pos[i], mpos = R.mpos[i], mtid = R.mtid[i], tlen = R.tlen[i], flag = R.flag[i], mq = R.mapq[i], lq = R.l_qseq[i];
    const int add = (mq >= c_prm.min_mapq) ? c_prm.add_factor : 0;
    const int ins_min = c_prm.insert_min, ins_max = c_prm.insert_max, ins_mean = c_prm.insert_mean, sc_min = c_prm.sc_min;
    em.pos = pos;
    // CIGAR summary (src/GROM.c:6740-6750, 6997-7000, 7067-7099) and CIGAR indel items (src/GROM.c:7187-7423)
    const int ncig_all = R.n_cigar[i], ncig = min(ncig_all, c_prm.max_cigar_ops);
    const uint64_t coff = R.cigar_off[i];
    int start_adj = 0, end_adj = 0, indel = 0, lseq = lq;
    {
        int64_t tp = pos; int qoff = 0;
        for (int k = 0; k < ncig; k++) {
            const uint32_t c = R.cigar[coff + k];
            const int op = c & 15, len = (int)(c >> 4);
            if (op == OP_M || op == OP_N || op == OP_EQ || op == OP_X) { tp += len; if (op != OP_N) qoff += len; }
            else if (op == OP_S) qoff += len;
            else if (op == OP_I) { indel += len; em.item(tp, tp + 1, CL_INDEL_I, len, (int)i, 0, AN_FULL, false, 0, add, 0, qoff); qoff += len; }   // v = read index, mchr = query offset of the inserted bases
            else if (op == OP_D) {
                indel -= len;
                em.point(GA_INDEL_D_F_RD, tp, 1);
                em.item(tp, tp + 1, CL_INDEL_D_F, len, 0, 0, AN_FULL, false, 0, add);
                em.point(GA_INDEL_D_R_RD, tp + len - 1, 1);
                em.item(tp + len - 1, tp + len, CL_INDEL_D_R, len, 0, 0, AN_FULL, false, 0, add);
                tp += len;
            } else if (op == OP_H) lseq += len;
            if (op == OP_S || op == OP_H) { if (k == 0) start_adj = len; if (k == ncig - 1) end_adj = len; }
        }
    }
    const bool paired = flag & F_PAIRED, munmap = flag & F_MUNMAP, rev = flag & F_REVERSE, mrev = flag & F_MREVERSE;
    const bool same = (tid == mtid);
    const int64_t E = (int64_t)pos - start_adj + lseq - end_adj - indel;
    const int64_t F = (int64_t)pos - start_adj - indel + ins_max - lseq;
    const int64_t Bk0 = (int64_t)pos - start_adj - ins_max + 2 * lseq;
    // the reference's window when this read is applied (src/GROM.c:5845-5847, 6317, 6408-6411)
    const int W = 2 * max(c_prm.overlap_mult * 8 * (2 * ins_mean - 1), c_prm.overlap_mult * 8 * (ins_max + 1)), first_pos = W / 4 + 1;   // src/GROM.c:22282-22290
    int64_t pproc = (int64_t)pos - (int64_t)c_prm.overlap_mult * ins_max; if (pproc < first_pos) pproc = first_pos;
    const int64_t win_lo = pproc - (W / 4 + ((n_leading + 2 + (pproc - first_pos)) % (W / 2))), win_hi = win_lo + W;
    const int64_t Bk = max(Bk0, win_lo);
    const int tol = ins_max - ins_min;
    const int tolI = tol + max(0, ins_mean - 2 * lseq);
    const bool clipE = end_adj >= sc_min, clipS = start_adj >= sc_min;

    // split-read fields
    int apos = c_prm.splitread ? SA.sa_pos[i] : -1;
    const bool sa_same = apos >= 0 && SA.sa_same[i];
    const int astrand = SA.sa_strand[i], amq = SA.sa_mapq[i], a_sadj = SA.sa_start_adj[i], a_eadj = SA.sa_end_adj[i], a_indel = SA.sa_end_adj_indel[i];
    const bool sa_q = sa_same && amq >= c_prm.min_mapq && mq >= c_prm.min_mapq;
    const int64_t AE = (int64_t)apos - a_sadj + lseq - a_eadj - a_indel;
    const bool paired_same = paired && !munmap && same;

    // ---- split-read deletion (src/GROM.c:7425-7950)
    if (sa_q && ((!rev && astrand == 0) || (rev && astrand == 1))) {
        bool sr_del = false; int64_t S = 0, Eo = 0;
        if (paired_same) {
            if (!rev) {
                if (pos < apos && tlen <= ins_max && apos < mpos && apos - E < ins_max && apos - E > 0 &&
                    abs(lseq - end_adj - a_sadj) <= c_prm.max_split_loss && lseq - start_adj - end_adj - indel >= c_prm.min_sr_len &&
                    lseq - a_sadj - a_eadj - a_indel >= c_prm.min_sr_len) { sr_del = true; S = E; Eo = apos; }
            } else {
                if (apos < pos && abs(tlen) < ins_max && mpos < apos &&
                    abs(lseq - start_adj - a_eadj) <= c_prm.max_split_loss && lseq - start_adj - end_adj - indel >= c_prm.min_sr_len &&
                    lseq - a_sadj - a_eadj - a_indel >= c_prm.min_sr_len) { S = AE; Eo = pos; sr_del = S < Eo; }
            }
        } else {
            if (!rev) { if (pos < apos && apos - E < ins_max && apos - E > 0) { sr_del = true; S = E; Eo = apos; } }
            else { if (apos < pos && pos - AE < ins_max) { S = AE; Eo = pos; sr_del = S < Eo; } }
        }
        if (sr_del) {
            const int64_t gap = Eo - S;
            if (gap < c_prm.lseq && gap < ins_max - ins_mean) {
                em.point(GA_INDEL_D_F_RD, S, 1);
                em.item(S, S + 1, CL_INDEL_D_F, (int)gap, 0, 0, AN_FULL, false, 0, add);
                em.point(GA_INDEL_D_R_RD, Eo - 1, 1);
                em.item(Eo - 1, Eo, CL_INDEL_D_R, (int)gap, 0, 0, AN_FULL, false, 0, add);
            }
            em.diff(GA_RD, S, (S) + 1, 1);
            em.item(S, S + 1, CL_DEL_F, (int)(gap + ins_mean), pos < apos ? pos : apos, RS_MAX_ONLY, AN_FULL, false, tol, add);
            em.diff(GA_RD, Eo - 1, (Eo - 1) + 1, 1);
            em.item(Eo - 1, Eo, CL_DEL_R, (int)(gap + ins_mean), pos < apos ? apos : pos, RS_MINMAX, AN_FULL, false, tol, add);
        }
    }

    // ---- pairs (src/GROM.c:7963-10953)
    if (paired && !munmap) {
        if (same) {
            if (mpos > pos) {
                if (!rev && mrev) {
                    if (tlen >= ins_min && tlen <= ins_max) {
                        // split tandem-dup evidence instead of the concordant range (src/GROM.c:7978-8340)
                        bool sr_dup = false;
                        if (sa_q && astrand == 0 && pos < apos && apos < mpos) {
                            const int it = indel > 0 ? indel : 0, ait = a_indel > 0 ? indel : 0;      // sic (src/GROM.c:7998)
                            if (abs(lseq - start_adj - a_eadj) <= c_prm.max_split_loss && lseq - start_adj - end_adj - it >= c_prm.min_sr_len &&
                                lseq - a_sadj - a_eadj - ait >= c_prm.min_sr_len) {
                                sr_dup = true;
                                const int64_t ls = pos, le = AE;
                                em.diff(GA_RD, le, (le) + 1, 1);
                                em.item(le, le + 1, CL_DUP_F, (int)(le - ls - ins_mean), pos < apos ? apos : pos, RS_MINMAX, AN_FULL, false, tol, add, SVM_DUPQ);
                                em.diff(GA_RD, ls - 1, (ls - 1) + 1, 1);
                                em.item(ls - 1, ls, CL_DUP_R, (int)(le - ls - ins_mean), pos < apos ? pos : apos, RS_MINMAX, AN_FULL, false, tol, add);
                            }
                        }
                        if (!sr_dup) { const int64_t hi = min((int64_t)mpos, win_hi); em.diff(GA_RD, E, hi, 1); em.diff(GA_CONC, E, hi, 1); }
                    } else if (tlen > 2 * ins_max) {
                        const int64_t hi = min(min(F, win_hi), (int64_t)mpos);
                        em.diff(GA_RD, E, hi, 1);
                        em.item(E, hi, CL_DEL_F, tlen, pos, RS_SET_RE, AN_FWD, clipE, tol, add);
                    } else if (tlen > ins_max) {
                        const int64_t hi = min((int64_t)mpos, win_hi);
                        em.diff(GA_RD, E, hi, 1);
                        em.item(E, min(hi, F), CL_DEL_F, tlen, pos, RS_SET_RE, AN_FWD, clipE, tol, add);
                        if (abs(tlen) <= 2 * ins_max) {
                            const int64_t thr = (int64_t)pos - start_adj + tlen - ins_max + lseq;
                            em.item(max(E, thr + 1), hi, CL_DEL_R, tlen, mpos, RS_MINMAX, AN_BWD, clipS, tol, add);
                        }
                    } else if (tlen < ins_min) {
                        const bool no_ins = sa_same && !rev && astrand == 0 && apos < pos && pos < mpos;
                        if (!no_ins) { const int64_t hi = min((int64_t)mpos, win_hi); em.diff(GA_RD, E, hi, 1); em.diff(GA_INS, E, hi, add); }
                    }
                } else if (!rev && !mrev) {
                    if (mpos - pos >= 10) {
                        const int64_t hi = min(min(F, win_hi), (int64_t)mpos);
                        em.diff(GA_RD, E, hi, 1);
                        em.item(E, hi, CL_INV_F1, tlen, pos, RS_SET_RE, AN_FWD, clipE, tolI, add);
                    }
                } else if (rev) {
                    if (mpos - pos >= 10) {
                        em.diff(GA_RD, Bk, pos, 1);
                        em.item(Bk, pos, mrev ? CL_INV_R1 : CL_DUP_R, tlen, pos, RS_SET_RE, AN_BWD, clipS, mrev ? tolI : tol, add);
                    }
                }
            } else {
                if (rev && !mrev) {
                    if (abs(tlen) >= ins_min && abs(tlen) <= ins_max) {
                        // split tandem-dup evidence, reverse read (src/GROM.c:9361-9722)
                        if (sa_q && astrand == 1 && apos < pos && mpos < apos) {
                            const int it = indel > 0 ? indel : 0, ait = a_indel > 0 ? indel : 0;
                            if (abs(lseq - a_sadj - end_adj) <= c_prm.max_split_loss && lseq - start_adj - end_adj - it >= c_prm.min_sr_len &&
                                lseq - a_sadj - a_eadj - ait >= c_prm.min_sr_len) {
                                const int64_t ls = apos, le = E;
                                if (ls < le) {
                                    em.diff(GA_RD, le, (le) + 1, 1);
                                    em.item(le, le + 1, CL_DUP_F, (int)(le - ls - ins_mean), pos < apos ? apos : pos, RS_MINMAX, AN_FULL, false, tol, add, SVM_DUPQ);
                                    em.diff(GA_RD, ls - 1, (ls - 1) + 1, 1);
                                    em.item(ls - 1, ls, CL_DUP_R, (int)(le - ls - ins_mean), pos < apos ? pos : apos, RS_MINMAX, AN_FULL, false, tol, add);
                                }
                            }
                        }
                    } else if (abs(tlen) > 2 * ins_max) {
                        em.diff(GA_RD, Bk, pos, 1);
                        em.item(Bk, pos, CL_DEL_R, abs(tlen), pos, RS_SET_RE, AN_BWD, clipS, tol, add);
                    }
                } else if (!rev && !mrev) {
                    if (pos - mpos >= 10) {
                        const int64_t hi = min(F, win_hi);
                        em.diff(GA_RD, E, hi, 1);
                        em.item(E, hi, CL_INV_F2, abs(tlen), pos, RS_SET_RE, AN_FWD, clipE, tolI, add);
                    }
                } else if (mrev) {
                    if (pos - mpos >= 10) {
                        if (!rev) {
                            const int64_t hi = min(F, win_hi);
                            em.diff(GA_RD, E, hi, 1);
                            em.item(E, hi, CL_DUP_F, abs(tlen), pos, RS_SET_RE, AN_FWD, clipE, tol, add);
                        } else {
                            const int64_t lo = max(Bk0, (int64_t)mpos + lseq);
                            em.diff(GA_RD, lo, pos, 1);
                            em.item(lo, pos, CL_INV_R2, abs(tlen), pos, RS_SET_RE, AN_BWD, clipS, tolI, add);
                        }
                    }
                }
            }
        } else {
            const uint32_t ex = SVM_CTX | (mrev ? SVM_NEG : 0u);
            if (!rev) {
                const int64_t hi = min(F, win_hi);
                em.diff(GA_RD, E, hi, 1);
                em.item(E, hi, CL_CTX_F, mrev ? -mpos : mpos, pos, RS_SET_RE, AN_FWD, clipE, tol, add, ex, mtid);
            } else {
                const int64_t lo = max((int64_t)pos - start_adj + lseq - ins_max + lseq, win_lo);
                em.diff(GA_RD, lo, pos, 1);
                em.item(lo, pos, CL_CTX_R, mrev ? -mpos : mpos, pos, RS_SET_RE, AN_BWD, clipS, tol, add, ex, mtid);
            }
        }
    } else if (paired && munmap) {
        if (!rev) { const int64_t hi = min(F, win_hi); em.diff(GA_RD, E, hi, 1); em.diff(GA_MUNMAPPED_F, E, hi, add); }
        else {
            const int64_t lo = max((int64_t)pos - start_adj + lseq + indel - ins_max + lseq, win_lo);
            em.diff(GA_RD, lo, pos, 1); em.diff(GA_MUNMAPPED_R, lo, pos, add);
        }
    }
}

__global__ void __launch_bounds__(256) k_sv_count(DevReads R, SvReadArrays SA, int tid, int64_t n_leading, const uint8_t *state, int64_t P, int32_t *item_cnt)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R.n) return;
    int c = 0;
    if (state[i] == 1) {
        SvEmitter<false> em; em.arrays = nullptr; em.P = P; em.Ppad = 0; em.out = nullptr; em.n = 0; em.room = 0; em.max_fwd = em.max_bwd = 0; em.pos = 0;
        sv_read_items<false>(R, SA, i, tid, n_leading, em);
        c = em.n;
    }
    item_cnt[i] = c;
}

// item_cnt holds the INCLUSIVE prefix sum on entry
__global__ void __launch_bounds__(256) k_sv_emit(DevReads R, SvReadArrays SA, int tid, int64_t n_leading, const uint8_t *state, const int32_t *item_incl,
                                                  SvItem *items, int64_t items_cap, int32_t *arrays, int64_t P, int64_t Ppad, int *reach /* [0] fwd [1] bwd */)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int mf = 0, mb = 0;
    if (i < R.n && state[i] == 1) {
        const int64_t first = i ? item_incl[i - 1] : 0;
        SvEmitter<true> em; em.arrays = arrays; em.P = P; em.Ppad = Ppad; em.out = items + first; em.room = items_cap - first; em.n = 0; em.max_fwd = em.max_bwd = 0; em.pos = 0;
        sv_read_items<true>(R, SA, i, tid, n_leading, em);
        mf = em.max_fwd; mb = em.max_bwd;
    }
    mf = __reduce_max_sync(0xffffffffu, mf); mb = __reduce_max_sync(0xffffffffu, mb);
    if ((threadIdx.x & 31) == 0) { if (mf) atomicMax(reach, mf); if (mb) atomicMax(reach + 1, mb); }
}

// per tile: slice [lo, hi) of the item list whose reads lie within reach of the tile
__global__ void __launch_bounds__(256) k_sv_tiles(const int32_t *pos, int64_t n, const int32_t *item_incl, const int *reach, int64_t n_tiles, int64_t items_cap, int2 *tile_rng)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const int64_t a = t * SV_T - reach[0], b = t * SV_T + SV_T + reach[1];     // reads with pos in [a, b)
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t m = (lo + hi) >> 1; if ((int64_t)pos[m] < a) lo = m + 1; else hi = m; }
    const int64_t r0 = lo;
    hi = n;
    while (lo < hi) { const int64_t m = (lo + hi) >> 1; if ((int64_t)pos[m] < b) lo = m + 1; else hi = m; }
    const int64_t r1 = lo;
    tile_rng[t] = make_int2((int)min((int64_t)(r0 ? item_incl[r0 - 1] : 0), items_cap), (int)min((int64_t)(r1 ? item_incl[r1 - 1] : 0), items_cap));
}

// ---- per-position fold ----------------------------------------------------------------------------------------------
struct SvTileState {
    int w[SV_NCLS][SV_T];
    int idist[3][SV_T];
    int rs[SV_NCL][SV_T], re[SV_NCL][SV_T];
    int mchr[2][SV_T];
    int oth[SV_T];                     // index into the other-slot pool, -1 = none
    int ins_src[3][SV_T];              // read index / query offset / length of the inserted sequence last stored (src/GROM.c:7219-7228)
    int oth_n[SV_T];                   // side slots in use (slots are handed out in order and never freed, so slot oth_n is the first empty one)
    double dist[SV_NCL][SV_T];
};

__device__ __forceinline__ int iabs_trunc(double v) { const int t = (int)v; return t < 0 ? -t : t; }     // C abs() applied to a double

__device__ __forceinline__ bool sv_join(const SvItem &it, double x, double dist, int w, int mchr)
{
    const double lim = (double)it.tol * (1.0 + (1.0 / (double)w));
    if (it.meta & SVM_CTX) {
        if (mchr != it.mchr) return false;
        const double mp = (it.meta & SVM_NEG) ? -x : x;                        // +mate position
        if (!(it.meta & SVM_NEG)) return (double)iabs_trunc(dist - mp) <= lim && dist > 0;
        return (double)iabs_trunc((double)iabs_trunc(dist) - mp) <= lim && dist < 0;
    }
    return (double)iabs_trunc(dist - x) <= lim;
}

__device__ __forceinline__ void sv_track(int mode, int v, int &rs, int &re)
{
    if (mode == RS_SET_RE) re = v;
    else if (mode == RS_MINMAX) { if (v < rs) rs = v; if (v > re) re = v; }
    else { if (v > re) re = v; }
}

__device__ SvOther *sv_others(SvTileState &S, int t, const SvDev &D)
{
    if (S.oth[t] < 0) {
        const int k = atomicAdd(D.pool_used, 1);
        if (k >= D.pool_cap) { atomicExch(D.err, 1); return nullptr; }
        S.oth[t] = k; S.oth_n[t] = 0;          // no zero fill: only slots below oth_n are ever read
    }
    return D.pool + (size_t)S.oth[t] * SV_OTHER;
}

// one item at one position (thread t of the tile); mirrors the reference's cluster template, e.g. src/GROM.c:8400-8525
__device__ void sv_apply_cluster(SvTileState &S, int t, const SvItem &it, int p, const SvDev &D)
{
    const int k = it.meta & 15, mode = (it.meta >> 4) & 3, an = (it.meta >> 6) & 3;
    bool full = true;
    if (it.meta & SVM_CLIPPED) full = (an == AN_FWD) ? (p == it.lo) : (an == AN_BWD) ? (p == it.hi - 1) : true;
    const int w = full ? it.add : it.add / 2;
    const double wd = full ? (double)it.add : (double)it.add / 2.0;
    const double x = (double)it.x;
    const bool ctx = it.meta & SVM_CTX;
    int &W = S.w[k][t]; int &RS = S.rs[k][t]; int &RE = S.re[k][t]; double &DI = S.dist[k][t];
    if (W == 0) {
        const int old_re = RE;
        W = w; DI = x; RS = it.v; RE = it.v;
        if (ctx) S.mchr[k - CL_CTX_F][t] = it.mchr;
        if (it.meta & SVM_DUPQ) { RE = old_re; S.re[CL_DEL_F][t] = it.v; }      // src/GROM.c:8037-8045, 9423-9429
        return;
    }
    if (sv_join(it, x, DI, W, ctx ? S.mchr[k - CL_CTX_F][t] : 0)) {
        W += w; DI += wd * (x - DI) / (double)W;
        sv_track(mode, it.v, RS, RE);
        return;
    }
    SvOther *o = sv_others(S, t, D);
    if (!o) return;
    const int otype = k + 1;
    const int nu = S.oth_n[t];
    for (int s = 0; s < nu; s++) {
        if (o[s].type == otype && sv_join(it, x, o[s].dist, o[s].w, o[s].mchr)) {
            o[s].w += w; o[s].dist += wd * (x - o[s].dist) / (double)o[s].w;
            sv_track(mode, it.v, o[s].rs, o[s].re);
            if (o[s].w > W) {
                const int tw = o[s].w, trs = o[s].rs, tre = o[s].re, tm = o[s].mchr; const double td = o[s].dist;
                o[s].w = W; o[s].dist = DI; o[s].rs = RS; o[s].re = RE;
                W = tw; DI = td; RS = trs; RE = tre;
                if (ctx) { o[s].mchr = S.mchr[k - CL_CTX_F][t]; S.mchr[k - CL_CTX_F][t] = tm; }
            }
            return;
        }
    }
    if (nu < SV_OTHER) {                       // first empty slot
        SvOther n_; n_.w = w; n_.type = otype; n_.mchr = ctx ? it.mchr : 0; n_.rs = it.v; n_.re = it.v; n_.pad = 0; n_.dist = x;
        o[nu] = n_; S.oth_n[t] = nu + 1;
        return;
    }
    for (int s = 0; s < SV_OTHER; s++) {       // all 50 in use: overwrite the first whose weight is <= cdp_add (src/GROM.c:8506-8523)
        if (o[s].w <= it.add) {
            o[s].w = w; o[s].type = otype; o[s].dist = x; o[s].rs = it.v; o[s].re = it.v;
            if (ctx) o[s].mchr = it.mchr;
            return;
        }
    }
}

// small-indel slots: exact-length match (src/GROM.c:7213-7285, 7292-7350, 7353-7415)
__device__ void sv_apply_indel(SvTileState &S, int t, const SvItem &it, const SvDev &D)
{
    const int k = it.meta & 15, j = k - CL_INDEL_I, len = it.x, add = it.add;
    int &W = S.w[k][t]; int &DI = S.idist[j][t];
    if (W == 0) {
        if (k == CL_INDEL_I && len <= c_prm.indel_i_seq_len) { S.ins_src[0][t] = it.v; S.ins_src[1][t] = it.mchr; S.ins_src[2][t] = len; }
        W = add; DI = len; return;
    }
    if (len == DI) { W += add; return; }
    SvOther *o = sv_others(S, t, D);
    if (!o) return;
    const int otype = 11 + j;
    const int nu = S.oth_n[t];
    for (int s = 0; s < nu; s++) {
        if (o[s].type == otype && (uint32_t)len == (uint32_t)(o[s].dist + 0.5)) {
            o[s].w += add;
            if (o[s].w > W) { const int tw = o[s].w; const double td = o[s].dist; o[s].w = W; o[s].dist = (double)DI; W = tw; DI = (int)(uint32_t)(td + 0.5); }
            return;
        }
    }
    if (nu < SV_OTHER) {
        SvOther n_; n_.w = add; n_.type = otype; n_.mchr = 0; n_.rs = 0; n_.re = 0; n_.pad = 0; n_.dist = (double)len;
        o[nu] = n_; S.oth_n[t] = nu + 1;
        return;
    }
    for (int s = 0; s < SV_OTHER; s++)
        if (o[s].w <= add) { o[s].w = add; o[s].type = otype; o[s].dist = (double)len; o[s].rs = 0; o[s].re = 0; return; }
}

__global__ void __launch_bounds__(SV_T) k_sv_apply(const SvItem *__restrict__ items, const int2 *__restrict__ tile_rng, int64_t P, int64_t Ppad,
                                                    int32_t *__restrict__ arrays, SvDev D, uint16_t *tile_dirty)
{
    extern __shared__ __align__(16) uint8_t sv_smem_raw[];
    SvTileState &S = *reinterpret_cast<SvTileState *>(sv_smem_raw);
    __shared__ SvItem s_it[SV_T];
    const int t = threadIdx.x;
    const int64_t tile_lo = (int64_t)blockIdx.x * SV_T;
    const int p = (int)tile_lo + t;
    const int2 rng = tile_rng[blockIdx.x];
    // does any candidate item overlap this tile?
    int any = 0;
    for (int k = rng.x + t; k < rng.y; k += SV_T) { const int lo = items[k].lo, hi = items[k].hi; if (hi > tile_lo && lo < tile_lo + SV_T) any = 1; }
    if (!__syncthreads_or(any)) return;
    for (int k = 0; k < SV_NCLS; k++) S.w[k][t] = 0;
    for (int k = 0; k < 3; k++) S.idist[k][t] = 0;
    for (int k = 0; k < SV_NCL; k++) { S.rs[k][t] = 0; S.re[k][t] = 0; S.dist[k][t] = 0; }
    S.mchr[0][t] = S.mchr[1][t] = 0; S.oth[t] = -1; S.oth_n[t] = 0; S.ins_src[0][t] = S.ins_src[1][t] = S.ins_src[2][t] = 0;
    for (int base = rng.x; base < rng.y; base += SV_T) {
        __syncthreads();
        const int cnt = min(SV_T, rng.y - base);
        if (t < cnt) s_it[t] = items[base + t];
        __syncthreads();
        for (int k = 0; k < cnt; k++) {
            const SvItem it = s_it[k];
            if (p >= it.lo && p < it.hi) {
                if ((it.meta & 15) >= CL_INDEL_I) sv_apply_indel(S, t, it, D); else sv_apply_cluster(S, t, it, p, D);
            }
        }
    }
    // write only the classes that occur in this tile (bit k of the tile mask; bit 13 = side slots); k_sv_clear undoes exactly these
    unsigned mask = 0;
    for (int k = 0; k < SV_NCLS; k++) if (S.w[k][t] != 0 || (k < SV_NCL && (S.dist[k][t] != 0 || S.rs[k][t] != 0 || S.re[k][t] != 0)) || (k >= SV_NCL && S.idist[k - SV_NCL][t] != 0)) mask |= 1u << k;
    if (S.oth_n[t]) mask |= 1u << 13;
    __shared__ unsigned s_mask;
    if (t == 0) s_mask = 0;
    __syncthreads();
    mask = __reduce_or_sync(0xffffffffu, mask);
    if ((t & 31) == 0 && mask) atomicOr(&s_mask, mask);
    __syncthreads();
    mask = s_mask;
    if (t == 0) tile_dirty[blockIdx.x] = (uint16_t)mask;
    if (p < P) {
        for (int k = 0; k < SV_NCL; k++) if (mask & (1u << k)) {
            D.cl_w[(int64_t)k * Ppad + p] = S.w[k][t]; D.cl_rs[(int64_t)k * Ppad + p] = S.rs[k][t]; D.cl_re[(int64_t)k * Ppad + p] = S.re[k][t];
            D.cl_dist[(int64_t)k * Ppad + p] = S.dist[k][t];
        }
        if (mask & (1u << CL_CTX_F)) D.cl_mchr[p] = S.mchr[0][t];
        if (mask & (1u << CL_CTX_R)) D.cl_mchr[Ppad + p] = S.mchr[1][t];
        if (mask & (1u << CL_INDEL_I)) {
            arrays[(int64_t)GA_INDEL_I * Ppad + p] = S.w[CL_INDEL_I][t]; arrays[(int64_t)GA_INDEL_IDIST * Ppad + p] = S.idist[0][t];
            D.ins_src[p] = S.ins_src[0][t]; D.ins_src[Ppad + p] = S.ins_src[1][t]; D.ins_src[2 * Ppad + p] = S.ins_src[2][t];
        }
        if (mask & (1u << CL_INDEL_D_F)) { arrays[(int64_t)GA_INDEL_D_F * Ppad + p] = S.w[CL_INDEL_D_F][t]; arrays[(int64_t)GA_INDEL_D_FDIST * Ppad + p] = S.idist[1][t]; }
        if (mask & (1u << CL_INDEL_D_R)) { arrays[(int64_t)GA_INDEL_D_R * Ppad + p] = S.w[CL_INDEL_D_R][t]; arrays[(int64_t)GA_INDEL_D_RDIST * Ppad + p] = S.idist[2][t]; }
        if (mask & (1u << 13)) D.other_len[p] = S.oth_n[t];
        const int af = c_prm.add_factor, md = c_prm.min_disc;
        int km = (S.w[CL_INDEL_I][t] / af >= md ? 1 : 0) | (S.w[CL_INDEL_D_F][t] / af >= md ? 2 : 0) | (S.w[CL_INDEL_D_R][t] / af >= md ? 4 : 0);
#pragma unroll
        for (int c = 0; c < 10; c++) if (S.w[c][t] / af >= md) km |= 8 << c;            // breakpoint classes whose weight reaches the gate
        if (km) {
            const int k = atomicAdd(D.n_ins_pos, 1);
            if (k < D.ins_pos_cap) D.ins_pos[k] = make_int2(p, km); else atomicExch(D.err, 2);
        }
    }
}

// zero the outputs of the tiles the previous run touched (everything else is still zero from allocation time)
__global__ void __launch_bounds__(SV_T) k_sv_clear(uint16_t *tile_dirty, int64_t P, int64_t Ppad, int32_t *__restrict__ arrays, SvDev D)
{
    const unsigned mask = tile_dirty[blockIdx.x];
    if (!mask) return;
    const int64_t p = (int64_t)blockIdx.x * SV_T + threadIdx.x;
    if (p < P) {
        for (int k = 0; k < SV_NCL; k++) if (mask & (1u << k)) {
            D.cl_w[(int64_t)k * Ppad + p] = 0; D.cl_rs[(int64_t)k * Ppad + p] = 0; D.cl_re[(int64_t)k * Ppad + p] = 0; D.cl_dist[(int64_t)k * Ppad + p] = 0;
        }
        if (mask & (1u << CL_CTX_F)) D.cl_mchr[p] = 0;
        if (mask & (1u << CL_CTX_R)) D.cl_mchr[Ppad + p] = 0;
        if (mask & (1u << 13)) D.other_len[p] = 0;
        if (mask & (1u << CL_INDEL_I)) { arrays[(int64_t)GA_INDEL_I * Ppad + p] = 0; arrays[(int64_t)GA_INDEL_IDIST * Ppad + p] = 0; D.ins_src[2 * Ppad + p] = 0; }
        if (mask & (1u << CL_INDEL_D_F)) { arrays[(int64_t)GA_INDEL_D_F * Ppad + p] = 0; arrays[(int64_t)GA_INDEL_D_FDIST * Ppad + p] = 0; }
        if (mask & (1u << CL_INDEL_D_R)) { arrays[(int64_t)GA_INDEL_D_R * Ppad + p] = 0; arrays[(int64_t)GA_INDEL_D_RDIST * Ppad + p] = 0; }
    }
    __syncthreads();
    if (threadIdx.x == 0) tile_dirty[blockIdx.x] = 0;
}
