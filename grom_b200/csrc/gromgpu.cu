// gromgpu.cu -- B200 (sm_100a) implementation of GROM's per-chromosome evidence accumulation and SNV scan
// behind the C ABI of include/gromgpu.h.
//
// Design: the reference walks reads one at a time and scatters into a sliding window
// (reference src/GROM.c:5842-14980).  Here every reference position is owned by exactly one thread
// ("pull" formulation): a CTA owns a tile of consecutive positions, stages the records of the reads that
// overlap the tile in shared memory (BAM order) and every position-thread folds the reads that cover it
// into registers, then stores its counts once, coalesced.  No atomics on the hot arrays, every count array
// is written exactly once, and order-dependent per-position rules (the mate-overlap read-name slots of
// src/GROM.c:6805-6824) are evaluated in BAM order for free.  Sparse point/range updates (soft-clip
// classes, physical depth ranges) are scattered by the per-read prep kernel and finished by a single-pass
// decoupled-look-back prefix scan.
//
// All work is HBM-bound integer/byte work; there is deliberately no tensor-core path.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <algorithm>
#include <atomic>
#include <limits>
#include <vector>
#include <string>
#include <map>
#include <tuple>
#include <chrono>
#include <thread>
#include <sched.h>
#include "gromgpu.h"

#define F_PAIRED 1
#define F_UNMAP 4
#define F_MUNMAP 8
#define F_REVERSE 16
#define F_MREVERSE 32
#define F_DUP 1024

enum { OP_M = 0, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X };

static thread_local char g_err[1024];
static int fail(const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return -1;
}
// CUDA's current device is per host thread: every entry point selects the library's device first (g_params, c_prm and the two
// tables live on ONE device per process -- one process per GPU, like one rank of torch.distributed)
#define ON_DEV() do { if (g_device >= 0) { cudaError_t e_ = cudaSetDevice(g_device); if (e_ != cudaSuccess) return fail("cudaSetDevice(%d) failed: %s", g_device, cudaGetErrorString(e_)); } } while (0)
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// ------------------------------------------------------------------------------------------------ globals
static bool g_inited = false;
static int g_device = -1;
static grom_params g_params;
static double *d_hez = nullptr, *d_mq = nullptr;
static cudaStream_t g_own_stream = nullptr, g_stream = nullptr;
__constant__ grom_params c_prm;

// ------------------------------------------------------------------------------------------------ device data
// per-read record produced by the prep kernel and consumed by the pileup kernel (32 bytes, 2 x 16-byte loads)
struct __align__(16) PrepRec {
    int32_t  pos;
    int32_t  ext_end;     // one past the last reference position any M/=/X base of the read can touch
    uint32_t base16;      // base_off / 16
    uint32_t misc;        // [7:0] mapq  [8] applied  [9] reverse  [10] simple (one M op == whole read)  [11] name storable  [31:16] l_qseq
    uint64_t hash;
    uint32_t cig_off;
    uint32_t n_cigar;
};
#define PR_APPLIED 0x100u
#define PR_REV     0x200u
#define PR_SIMPLE  0x400u
#define PR_NAMEOK  0x800u

struct DevReads {
    int64_t n;
    const int32_t *pos, *mpos, *tlen, *mtid, *l_qseq;
    const uint16_t *flag, *n_cigar;
    const uint8_t *mapq, *qname_len;
    const uint64_t *qname_hash, *cigar_off, *base_off;
    const uint32_t *cigar;
    const uint8_t *seq4, *qual;
};

#ifndef TILE
#define TILE 256           // positions per CTA in the pileup kernel (one per thread)
#endif
#ifndef CHUNK
#define CHUNK 32           // read records staged per shared-memory refill (CHUNK / 32 per lane of the producer warp; a multiple of 32, <= 224: packed counters)
#endif

// reference character -> BAM 4-bit code of toupper(char), 16 if the character is not a code letter
__device__ __forceinline__ int ref_code(unsigned char c)
{
    if (c >= 'a' && c <= 'z') c -= 32;
    switch (c) {
    case '=': return 0; case 'A': return 1; case 'C': return 2; case 'M': return 3; case 'G': return 4; case 'R': return 5;
    case 'S': return 6; case 'V': return 7; case 'T': return 8; case 'W': return 9; case 'Y': return 10; case 'H': return 11;
    case 'K': return 12; case 'D': return 13; case 'B': return 14; case 'N': return 15; default: return 16;
    }
}

// -M svtype (src/GROM.c:6435-6542); -1 = no class
__device__ __forceinline__ int dup_svtype(int tid, int mtid, int pos, int mpos, int flag)
{
    const bool rev = flag & F_REVERSE, mrev = flag & F_MREVERSE;
    if (tid == mtid) {
        if (mpos > pos) { if (!rev && mrev) return 0; if (!rev && !mrev) return 8; return mrev ? 9 : 1; }
        if (rev && !mrev) return 0;
        if (!rev && !mrev) return 8;
        if (mrev) return rev ? 9 : 1;
        return -1;
    }
    if (!rev) return mrev ? 12 : 11;
    return mrev ? 14 : 13;
}

// ---- K3: -M duplicate flags.  A read is a duplicate iff mapq >= q and an earlier read of the same start
// position carries the same (mpos, mtid, l_qseq, tlen, svtype) (src/GROM.c:6548-6588; the first read of a
// key run is always kept, so "an earlier kept read with this key exists" == "an earlier read with this key
// exists").  Reads of one position are adjacent in BAM order, so each thread scans its own short run
// backwards.  state: 0 not applied, 1 applied, 2 duplicate.
__global__ void __launch_bounds__(256) k_read_state(DevReads R, int tid, int first_pos, uint8_t *state, unsigned long long *bad /* [0] out of order [1] l_qseq > 65535 */)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R.n) return;
    const int pos = R.pos[i], flag = R.flag[i];
    // input validation rides along: the pileup relies on coordinate order and keeps l_qseq in 16 bits
    if (i > 0 && R.pos[i - 1] > pos) atomicAdd(bad, 1ull);
    if ((unsigned)R.l_qseq[i] > 65535u) atomicAdd(bad + 1, 1ull);
    uint8_t st = 1;
    if (pos < first_pos || (flag & (F_UNMAP | F_DUP))) st = 0;
    else if (c_prm.rmdup > 0 && (flag & F_PAIRED) && !(flag & F_MUNMAP) && R.mapq[i] >= c_prm.min_mapq) {
        const int mtid = R.mtid[i], mpos = R.mpos[i], tlen = R.tlen[i], lq = R.l_qseq[i];
        const int sv = dup_svtype(tid, mtid, pos, mpos, flag);
        if (sv >= 0) {
            for (int64_t j = i - 1; j >= 0 && R.pos[j] == pos; j--) {
                const int fj = R.flag[j];
                if ((fj & (F_UNMAP | F_DUP)) || !(fj & F_PAIRED) || (fj & F_MUNMAP)) continue;
                if (R.mpos[j] == mpos && R.mtid[j] == mtid && R.l_qseq[j] == lq && R.tlen[j] == tlen &&
                    dup_svtype(tid, mtid, pos, mpos, fj) == sv) { st = 2; break; }
            }
        }
    }
    state[i] = st;
}

// ---- K0: per-read CIGAR summary (src/GROM.c:6740-6750, 7067-7099), soft-clip class point updates
// (src/GROM.c:7103-7170) and the read-span range add for physical depth as a +1/-1 difference pair
// (src/GROM.c:7176-7181; finished by k_scan_inplace).
__global__ void __launch_bounds__(256) k_read_prep(DevReads R, int tid, int64_t P, int64_t Ppad, const uint8_t *state,
                                                    PrepRec *prep, int32_t *arrays, int *max_span,
                                                    unsigned long long *counters /* [0] applied [1] dups [2] aligned bases [3] rec bytes */)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int span = 0;
    unsigned long long n_app = 0, n_dup = 0, n_al = 0, n_bytes = 0;
    if (i < R.n) {
        const int pos = R.pos[i], flag = R.flag[i], mq = R.mapq[i], lq = R.l_qseq[i];
        const int ncig_all = R.n_cigar[i];
        const uint64_t coff = R.cigar_off[i];
        const uint8_t st = state[i];
        PrepRec pr;
        pr.pos = pos; pr.base16 = (uint32_t)(R.base_off[i] >> 4); pr.hash = R.qname_hash[i];
        pr.cig_off = (uint32_t)coff; pr.n_cigar = (uint32_t)ncig_all;
        uint32_t misc = (uint32_t)mq | ((uint32_t)(lq & 0xffff) << 16);
        if (flag & F_REVERSE) misc |= PR_REV;
        if (R.qname_len[i] < c_prm.read_name_len) misc |= PR_NAMEOK;
        int ext_end = pos;
        if (st == 1) {
            misc |= PR_APPLIED;
            const int ncig = min(ncig_all, c_prm.max_cigar_ops);
            int rp = 0, rd_ = 0;            // reference offsets of the pileup walk and of the CNV-depth walk
            int start_adj = 0, end_adj = 0, indel = 0, lseq = lq, al = 0;
            int ext_p = 0, ext_d = 0;
            for (int k = 0; k < ncig_all; k++) {
                const uint32_t c = R.cigar[coff + k];
                const int op = c & 15, len = (int)(c >> 4);
                const bool in_pile = k < ncig;
                if (op == OP_M || op == OP_EQ || op == OP_X) {
                    if (in_pile) { rp += len; ext_p = rp; }
                    rd_ += len; ext_d = rd_; al += len;
                } else if (op == OP_D) { if (in_pile) { rp += len; indel -= len; } rd_ += len; }
                else if (op == OP_N) { if (in_pile) rp += len; }
                else if (op == OP_I) { if (in_pile) indel += len; }
                else if (op == OP_H) { if (in_pile) lseq += len; }
                if (in_pile && (op == OP_S || op == OP_H)) { if (k == 0) start_adj = len; if (k == ncig - 1) end_adj = len; }
            }
            if (ncig_all == 1 && (R.cigar[coff] & 15) == OP_M && (int)(R.cigar[coff] >> 4) == lq) misc |= PR_SIMPLE;
            ext_end = pos + max(ext_p, ext_d);
            span = ext_end - pos;
            n_app = 1; n_al = (unsigned long long)al;
            n_bytes = 40ull + 4ull * ncig_all + (unsigned long long)((lq + 3) / 4) + (unsigned long long)lq;
            // soft-clip classes
            const int add = (mq >= c_prm.min_mapq) ? c_prm.add_factor : 0;
            const int mtid = R.mtid[i], mpos = R.mpos[i], tlen = R.tlen[i];
            const bool paired = flag & F_PAIRED, munmap = flag & F_MUNMAP, rev = flag & F_REVERSE, same = (tid == mtid);
            const int64_t rend = (int64_t)pos - start_adj + lseq - end_adj - indel;
#define BUMP(ARR, RD, CRD, X) do { const int64_t x_ = (X); if (x_ >= 0 && x_ < P) { \
                if (add) atomicAdd(arrays + (int64_t)(ARR) * Ppad + x_, add); \
                atomicAdd(arrays + (int64_t)(RD) * Ppad + x_, 1); atomicAdd(arrays + (int64_t)(CRD) * Ppad + x_, 1); } } while (0)
            if (start_adj >= c_prm.sc_min) {
                const int64_t x = (int64_t)pos - 1;
                if (!paired || (!rev && (munmap || (same && mpos > pos)))) BUMP(GA_SC_LEFT, GA_SC_LEFT_RD, GA_SC_RD, x);
                if (paired && !munmap && !same && rev) BUMP(GA_CTX_SC_LEFT, GA_CTX_SC_LEFT_RD, GA_CTX_SC_RD, x);
                if (paired && !munmap && same && rev && abs(tlen) <= c_prm.insert_max && mpos < pos)
                    BUMP(GA_INDEL_SC_LEFT, GA_INDEL_SC_LEFT_RD, GA_INDEL_SC_RD, x);
            }
            if (end_adj >= c_prm.sc_min) {
                if (!paired || (rev && (munmap || (same && mpos < pos)))) BUMP(GA_SC_RIGHT, GA_SC_RIGHT_RD, GA_SC_RD, rend);
                if (paired && !munmap && !same && !rev) BUMP(GA_CTX_SC_RIGHT, GA_CTX_SC_RIGHT_RD, GA_CTX_SC_RD, rend);
                if (paired && !munmap && same && !rev && abs(tlen) <= c_prm.insert_max && mpos > pos)
                    BUMP(GA_INDEL_SC_RIGHT, GA_INDEL_SC_RIGHT_RD, GA_INDEL_SC_RD, rend);
            }
#undef BUMP
            // physical depth over [pos, rend): difference pair
            if (rend > pos) {
                const int64_t a = max((int64_t)pos, (int64_t)0), e = min(rend, P);
                if (a < e) {
                    atomicAdd(arrays + (int64_t)GA_RD * Ppad + a, 1);
                    if (e < P) atomicAdd(arrays + (int64_t)GA_RD * Ppad + e, -1);
                }
            }
        } else if (st == 2) n_dup = 1;
        pr.ext_end = ext_end; pr.misc = misc;
        prep[i] = pr;
    }
    // block-level reduction of the small statistics, then one atomic per block and counter (same-address
    // atomics from every warp would serialise at L2)
    __shared__ int s_span[8];
    __shared__ unsigned long long s_cnt4[8][4];
    span = __reduce_max_sync(0xffffffffu, span);
    const unsigned app32 = __reduce_add_sync(0xffffffffu, (unsigned)n_app);
    const unsigned dup32 = __reduce_add_sync(0xffffffffu, (unsigned)n_dup);
    const unsigned al32 = __reduce_add_sync(0xffffffffu, (unsigned)n_al);
    const unsigned by32 = __reduce_add_sync(0xffffffffu, (unsigned)n_bytes);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_span[w] = span; s_cnt4[w][0] = app32; s_cnt4[w][1] = dup32; s_cnt4[w][2] = al32; s_cnt4[w][3] = by32; }
    __syncthreads();
    if (threadIdx.x < 4) {
        unsigned long long t = 0;
        for (int k = 0; k < 8; k++) t += s_cnt4[k][threadIdx.x];
        if (t) atomicAdd(counters + threadIdx.x, t);
    } else if (threadIdx.x == 4) {
        int m = 0;
        for (int k = 0; k < 8; k++) m = max(m, s_span[k]);
        if (m) atomicMax(max_span, m);
    }
}

// ---- tile -> index of the first read that reaches it: lower_bound(pos, tile_start - max_span), then forward past the reads that end
// before the tile (one long-span read -- a spliced alignment, a long deletion -- raises max_span for the whole contig, but only the tiles
// under it pay for it)
__global__ void __launch_bounds__(256) k_tile_index(const int32_t *pos, const PrepRec *__restrict__ prep, int64_t n, const int *max_span, int64_t n_tiles, int64_t *tile_first)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const int64_t tile_lo = t * TILE, key = tile_lo - (int64_t)(*max_span);
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if ((int64_t)pos[mid] < key) lo = mid + 1; else hi = mid; }
    while (lo < n && (int64_t)pos[lo] < tile_lo && (!(prep[lo].misc & PR_APPLIED) || (int64_t)prep[lo].ext_end <= tile_lo)) lo++;
    tile_first[t] = lo;
}

// ---- K1: pileup + CNV depth, one thread per reference position (src/GROM.c:6605-6671, 6740-7059)
//
// Per covering read a position-thread classifies one base.  Two paths:
//  * fast path (branch-free, ~97 % of bases): the read base equals the reference base and the reference base is
//    A/C/G/T.  The base index is then a per-thread constant, so the counts live in four dedicated registers
//    (m_hi, m_low, m_fs, m_pir) and every update is one predicated add.
//  * generic path (mismatches, non-ACGT codes, N/IUPAC reference letters): the reference's full rule including
//    the read-name slots, evaluated in BAM order because the staged read list is in BAM order.
// Stage totals: the packed per-stage counters are unpacked into these once per stage.  With PILE_LOCAL_TOTALS they are indexed
// through a run-time zero (c_prm.reserved0), which places them in local memory and frees eight registers of the fast loop.
#ifndef PILE_LOCAL_TOTALS
#define PILE_LOCAL_TOTALS 0
#endif
enum { F_MHI, F_MFS, F_MALL, F_MQ, F_MQALL, F_RDMQ, F_RDRD, F_RDCNT, F_COUNT };
struct PileAcc {
    int m_hi, m_low, m_fs, m_pir, m_all;                // reference-matching bases (m_low is derived: m_all - m_hi)
    // Everything else is touched by ~0.2 % of the bases.  These 16 counters and the name slots are indexed with run-time values on
    // purpose: that places them in (L1-resident) local memory and keeps the fast loop's accumulators in registers without spills.
    int bq, bq_all, mq, mq_all;
    int rd_mq, rd_rd, rd_low;
};
struct PileRare {
    int v[16];                                          // [0..3] snv, [4..7] low, [8..11] pir, [12..15] fs per base
    uint64_t nm[3]; int nm_cnt;
};
struct PileTotals { int f[F_COUNT]; };
enum { RA_SNV = 0, RA_LOW = 4, RA_PIR = 8, RA_FS = 12 };

__device__ __forceinline__ void pile_generic(PileAcc &a, PileRare &x, PileTotals &tt, uint64_t hash, uint32_t misc, int code, int qv, int qi, int lseq, int rc4,
                                          bool hi, int min_snv, int zz)
{
    const int mq = misc & 0xff;
    const int bi = (code == 1) ? 0 : (code == 2) ? 1 : (code == 4) ? 2 : (code == 8) ? 3 : -1;
    if (hi) {
        bool skip = false;
        const bool mism = (code != rc4);
        if (mism) {
            // first min_snv (<= 3) distinct names seen on mismatching high-quality bases (src/GROM.c:6805-6824)
            const int n = x.nm_cnt;
            for (int k = 0; k < n; k++) skip = skip || x.nm[k] == hash;
            if (!skip && n < min_snv && (misc & PR_NAMEOK)) { x.nm[n] = hash; x.nm_cnt = n + 1; }
        }
        if (!skip && bi >= 0) {
            const bool fwd = !(misc & PR_REV);
            const int pir = (mism || fwd) ? qi : lseq - qi;
            a.bq += qv; a.bq_all += qv; tt.f[F_MQ + zz] += mq; tt.f[F_MQALL + zz] += mq;
            x.v[RA_SNV + bi] += 1; x.v[RA_PIR + bi] += pir; x.v[RA_FS + bi] += fwd ? 1 : 0;
        }
    } else if (bi >= 0) {
        a.bq_all += qv; tt.f[F_MQALL + zz] += mq;
        x.v[RA_LOW + bi] += 1;
    }
}

// classify one base (code, quality) of one read at this thread's position and fold it into the accumulators;
// qi = query offset, lseq = read length seen by the position-in-read rule
__device__ __forceinline__ void pile_apply(PileAcc &a, PileRare &x, PileTotals &tt, int code, int qv, uint64_t hash, uint32_t misc, int qi, int lseq, int rc4,
                                           bool ref_acgt, bool mq_ok, int bqmin, int min_snv, int p_rel, int zz)
{
    const bool hi = mq_ok && qv >= bqmin;
    if (ref_acgt && code == rc4) {
        const int mq = misc & 0xff;
        const bool fwd = !(misc & PR_REV);
        a.bq_all += qv; tt.f[F_MQALL + zz] += mq; tt.f[F_MALL + zz] += 1;
        // m_pir holds the sum of (position in read) -+ p_rel, p_rel = position - tile start; the epilogue adds p_rel * (forward - reverse)
        if (hi) { a.bq += qv; tt.f[F_MQ + zz] += mq; tt.f[F_MHI + zz] += 1; a.m_pir += fwd ? qi - p_rel : lseq - qi + p_rel; tt.f[F_MFS + zz] += fwd ? 1 : 0; }
    } else {
        pile_generic(a, x, tt, hash, misc, code, qv, qi, lseq, rc4, hi, min_snv, zz);
    }
}

// ---- mbarrier / bulk-copy (TMA) wrappers: the read bases of a tile are staged in shared memory by
// cp.async.bulk, double-buffered, so the per-hit loads are LDS instead of dependent global loads.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx_only(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

#ifndef MBAR_BACKOFF
#define MBAR_BACKOFF
#endif
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) { MBAR_BACKOFF } }
__device__ __forceinline__ int lds_u8(uint32_t addr) { int v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }

#define QCAP (CHUNK * 160)        // quality bytes per stage (64 reads of 2x150 data); the 4-bit area is half of it
#ifndef NSTAGE
#define NSTAGE 4           // 4 stages of 32 reads measured fastest (profiles/README.md: finer stages, and less shared memory per CTA leaves more L1)
#endif
#define NWARP (TILE / 32)         // consumer warps, one reference position per thread
#define PILE_THREADS (TILE + 32)  // + 1 producer warp
#ifndef PILE_MIN_CTAS
#define PILE_MIN_CTAS 3
#endif
#ifndef SUBTILES
#define SUBTILES 16
#endif
//                              // consecutive tiles handled by one CTA (keeps the producer pipeline full across tiles)

// Per staged read the producer warp leaves three 16-byte records of plain ints in shared memory so that the
// position threads need no bit unpacking (all values are warp-uniform broadcasts):
struct __align__(16) StageA { int pos; uint32_t lq_fast; uint32_t qa; int bq_eff; };     // lq_fast = 0 unless single-M read fully inside the contig; bq_eff = -b if mapq >= -q else 256 (never reached)
// Per-read constants of the fast path, packed so that one add updates several counters (fields are unpacked into the 32-bit
// accumulators once per stage: <= CHUNK = 64 reads, so a count stays below 2^8 and a MAPQ sum below 2^16):
//   u_cov  = 1 | (mapq >= rd_min_mapq) << 8 | mapq << 16     added where the read covers the position           (CNV depth)
//   u_all  = 1 | mapq << 16                                  added where the base equals the A/C/G/T reference  (all qualities)
//   u_hi   = 1 | forward << 8 | mapq << 16                   ... and passes -q / -b
//   v_pir  = forward ? -(pos - tile_lo) : l_qseq + (pos - tile_lo): position in read = +-(p - tile_lo) + v_pir (src/GROM.c:6853-6864)
struct __align__(16) StageB { uint32_t u_cov, u_all, u_hi; int v_pir; };
struct __align__(16) StageD { int pir_c; int pir_s; int lq; uint32_t flags; };          // position-in-read = off * pir_s + pir_c (src/GROM.c:6853-6864)
struct __align__(16) StageC { uint64_t hash; uint32_t cig_off, n_cigar; };
struct __align__(8)  StageE { uint32_t base16; int ext_end; };
struct __align__(16) StageF { uint32_t op[4]; };                  // the CIGAR of a read with <= 4 operations (clipped reads, one indel): no global loads in the per-lane walk
#define SF_COMPLEX 1u             // needs the general CIGAR walk (or CNV-depth bound fails, or bases are not staged)
#define SF_GLOBAL  2u             // bases longer than a stage: read them from global memory
#define SF_REV     4u
#define SF_NAMEOK  8u
#define SF_MQOK    16u

// arguments of the per-position SNV gate that runs in the pileup epilogue (src/GROM.c:11096-11199, 15035-15043)
struct SnvScanArgs {
    int scan_first, scan_last;
    int64_t depth_bound;
    const double *hez, *mqt;
    grom_snv_cand *cand; unsigned int cand_cap; unsigned int *n_cand;
    unsigned long long *depth_sum;      // [0] sum of rd_rd + rd_low over non-N positions below depth_bound, [1] their count
    // small-insertion gate (src/GROM.c:11329-11453)
    grom_ins_cand *ins; unsigned int ins_cap; unsigned int *n_ins;
    grom_del_event *del_ev; unsigned int del_cap; unsigned int *n_del;
    const int32_t *other_len, *ins_src;
};

struct __align__(128) PileSmem {
    uint8_t qual[NSTAGE][QCAP];
    uint8_t seq[NSTAGE][QCAP / 2];
    StageA a[NSTAGE][CHUNK];
    StageB b[NSTAGE][CHUNK];
    StageD d[NSTAGE][CHUNK];
    StageC c[NSTAGE][CHUNK];
    StageE e[NSTAGE][CHUNK];
    StageF f[NSTAGE][CHUNK];
    int2 rng[NSTAGE][NWARP];      // per consumer warp: slice [t0, t1) of the staged reads that can reach its 32 positions
    uint64_t full[NSTAGE], empty[NSTAGE];
    int last[NSTAGE];
};

__global__ void __launch_bounds__(PILE_THREADS, PILE_MIN_CTAS) k_pileup(DevReads R, const PrepRec *__restrict__ prep, const int64_t *__restrict__ tile_first,
                                                          const int *__restrict__ max_span_p, int64_t n_tiles,
                                                          const char *__restrict__ fasta, int64_t P, int64_t Ppad, int32_t *__restrict__ arrays,
                                                          SnvScanArgs sc)
{
    extern __shared__ __align__(128) uint8_t pile_smem_raw[];
    PileSmem &S = *reinterpret_cast<PileSmem *>(pile_smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int q = c_prm.min_mapq, bqmin = c_prm.min_base_qual, rdq = c_prm.rd_min_mapq, min_snv = min(c_prm.min_snv, 3);
    const int64_t tile_begin = (int64_t)blockIdx.x * SUBTILES, tile_end = min(tile_begin + SUBTILES, n_tiles);

    if (threadIdx.x == 0) {
        for (int k = 0; k < NSTAGE; k++) { mbar_init(smem_u32(&S.full[k]), 32); mbar_init(smem_u32(&S.empty[k]), NWARP); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (wid == NWARP) {
        // ================= producer warp: for every tile, stream the reads that can reach it through the stage ring:
        // per chunk up to CHUNK reads (two per lane), compacted to the applied reads that overlap the tile, metadata as
        // plain ints + bulk copies (TMA) of their bases
        const int max_span = *max_span_p;
        int c = 0;
        constexpr int RPL = CHUNK / 32;                           // records per lane and chunk
        auto load_group = [&](int64_t first, PrepRec (&r)[RPL]) {
#pragma unroll
            for (int h = 0; h < RPL; h++) {
                const int64_t i = first + RPL * lane + h;
                r[h].pos = INT32_MAX; r[h].misc = 0; r[h].ext_end = INT32_MIN; r[h].base16 = 0; r[h].hash = 0; r[h].cig_off = r[h].n_cigar = 0;
                if (i < R.n) r[h] = prep[i];
            }
        };
        int64_t next = tile_begin < tile_end ? tile_first[tile_begin] : 0;
        PrepRec r[RPL];
        load_group(next, r);
        for (int64_t tile = tile_begin; tile < tile_end; tile++) {
            const int64_t tile_lo = tile * TILE, tile_hi = min(tile_lo + TILE, P);
            const int64_t next_tile_first = tile + 1 < tile_end ? tile_first[tile + 1] : 0;     // in flight while this tile's chunks are staged
            for (;; c++) {
                const int buf = c % NSTAGE;
                bool v[RPL], use[RPL], big[RPL], want[RPL], st[RPL];
                uint32_t sz[RPL]; int64_t rel[RPL];
                // The bases of consecutive reads lie back to back in the batch, so the stage takes ONE contiguous span per array (two bulk
                // copies per stage): from the first read that is needed to the last one that still fits, whatever lies between included.
                // (Offsets that do not ascend end the chunk early; the read then opens the next one.)
                int first_idx = CHUNK;
#pragma unroll
                for (int h = 0; h < RPL; h++) {
                    v[h] = next + RPL * lane + h < R.n;
                    use[h] = v[h] && (r[h].misc & PR_APPLIED) && (int64_t)r[h].ext_end > tile_lo && (int64_t)r[h].pos < tile_hi;
                    sz[h] = use[h] ? (((r[h].misc >> 16) + 31u) & ~31u) : 0u;
                    big[h] = sz[h] > QCAP;                             // longer than a stage: the position threads read its bases from global memory
                    want[h] = use[h] && !big[h] && sz[h];
                    const unsigned bw = __ballot_sync(0xffffffffu, want[h]);
                    if (bw) first_idx = min(first_idx, RPL * (__ffs(bw) - 1) + h);
                }
                // short CIGARs of the reads that need the general walk travel with the stage (loads issued here, consumed after the waits below)
                StageF cg4[RPL];
#pragma unroll
                for (int h = 0; h < RPL; h++) {
                    cg4[h].op[0] = cg4[h].op[1] = cg4[h].op[2] = cg4[h].op[3] = 0u;
                    const bool fast_h = (r[h].misc & PR_SIMPLE) && !big[h] && (int64_t)r[h].pos + (int64_t)(r[h].misc >> 16) < P;     // same test as below
                    if (use[h] && !fast_h && r[h].n_cigar <= 4u) {
#pragma unroll
                        for (int k = 0; k < 4; k++) if ((uint32_t)k < r[h].n_cigar) cg4[h].op[k] = __ldg(R.cigar + r[h].cig_off + k);
                    }
                }
                uint32_t base_first = 0;
                if (first_idx < CHUNK) {
                    uint32_t cand = 0;
#pragma unroll
                    for (int h = 0; h < RPL; h++) if (first_idx % RPL == h) cand = r[h].base16;
                    base_first = __shfl_sync(0xffffffffu, cand, first_idx / RPL);
                }
                // reads consumed from the stream: the leading run of reads that fit (read order = RPL * lane + h)
                int count = CHUNK;
#pragma unroll
                for (int h = 0; h < RPL; h++) {
                    rel[h] = ((int64_t)r[h].base16 - (int64_t)base_first) * 16;
                    const bool fit = v[h] && (!want[h] || (rel[h] >= 0 && rel[h] + sz[h] <= QCAP));
                    const unsigned nf = ~__ballot_sync(0xffffffffu, fit);
                    if (nf) count = min(count, RPL * (__ffs(nf) - 1) + h);
                }
                bool past = false;
                uint32_t span = 0;
                int below = 0;                                          // staged reads in lower lanes
                const unsigned lt = (1u << lane) - 1u;
#pragma unroll
                for (int h = 0; h < RPL; h++) {
                    const bool in = RPL * lane + h < count;
                    past = past || (in && v[h] && (int64_t)r[h].pos >= tile_hi);
                    st[h] = in && use[h];                               // staged (compacted) reads
                    below += __popc(__ballot_sync(0xffffffffu, st[h]) & lt);
                    if (st[h] && want[h]) span = max(span, (uint32_t)rel[h] + sz[h]);
                }
                const bool last = (__ballot_sync(0xffffffffu, past) != 0u) || (next + count >= R.n);
                span = __reduce_max_sync(0xffffffffu, span);
                // the records of the chunk after this one travel while this one is staged
                const int64_t next_after = last ? next_tile_first : next + count;
                PrepRec nx[RPL];
                load_group(next_after, nx);
                // per consumer warp: staged reads from the first one that reaches the warp's 32 positions (ext_end > wlo) to the last one that
                // starts at or before them (lane w keeps the range of warp w).  No dependence on the contig-wide max_span.
                static_assert(RPL == 1, "the per-warp slices index the compacted stage by lane");
                int2 my_rng = make_int2(0, 0);
                const unsigned stmask = __ballot_sync(0xffffffffu, st[0]);
#pragma unroll
                for (int w = 0; w < NWARP; w++) {
                    const int wlo = (int)tile_lo + 32 * w, whi = wlo + 31;
                    const int t1 = __popc(__ballot_sync(0xffffffffu, st[0] && r[0].pos <= whi));
                    const unsigned reach = __ballot_sync(0xffffffffu, st[0] && r[0].ext_end > wlo);
                    const int t0 = reach ? __popc(stmask & ((1u << (__ffs(reach) - 1)) - 1u)) : t1;
                    if (lane == w) my_rng = make_int2(min(t0, t1), t1);
                }
                // everything above ran while the position threads were still reading this stage's previous contents
                if (c >= NSTAGE) mbar_wait(smem_u32(&S.empty[buf]), (uint32_t)((c / NSTAGE - 1) & 1));
                const uint32_t bar = smem_u32(&S.full[buf]);
                const uint32_t qbase = smem_u32(&S.qual[buf][0]), sbase = smem_u32(&S.seq[buf][0]);
                if (lane == 0 && span) {
                    mbar_expect_tx_only(bar, span + (span >> 1));
                    bulk_g2s(qbase, R.qual + ((uint64_t)base_first << 4), span, bar);
                    bulk_g2s(sbase, R.seq4 + ((uint64_t)base_first << 3), span >> 1, bar);
                }
                if (lane < NWARP) S.rng[buf][lane] = my_rng;
                int t = below;
#pragma unroll
                for (int h = 0; h < RPL; h++) {
                    if (st[h]) {
                        const PrepRec &rr = r[h];
                        const uint32_t off = (uint32_t)rel[h];
                        const int mq = rr.misc & 0xff, lq = (int)(rr.misc >> 16);
                        const bool rev = rr.misc & PR_REV;
                        const bool fast = (rr.misc & PR_SIMPLE) && !big[h] && (int64_t)rr.pos + lq < P;
                        const int prel = rr.pos - (int)tile_lo;
                        StageA A; A.pos = rr.pos; A.lq_fast = fast ? (uint32_t)lq : 0u; A.qa = qbase + off; A.bq_eff = (mq >= q) ? bqmin : 256;
                        StageB B; B.u_cov = 1u | ((mq >= rdq) ? 0x100u : 0u) | ((uint32_t)mq << 16); B.u_all = 1u | ((uint32_t)mq << 16);
                        B.u_hi = 1u | (rev ? 0u : 0x100u) | ((uint32_t)mq << 16); B.v_pir = rev ? lq + prel : -prel;
                        StageD D; D.pir_c = rev ? lq : 0; D.pir_s = rev ? -1 : 1; D.lq = lq;
                        D.flags = (fast ? 0u : SF_COMPLEX) | (big[h] ? SF_GLOBAL : 0u) | (rev ? SF_REV : 0u) | ((rr.misc & PR_NAMEOK) ? SF_NAMEOK : 0u) | ((mq >= q) ? SF_MQOK : 0u);
                        StageC C; C.hash = rr.hash; C.cig_off = rr.cig_off; C.n_cigar = rr.n_cigar;
                        StageE E; E.base16 = rr.base16; E.ext_end = rr.ext_end;
                        S.a[buf][t] = A; S.b[buf][t] = B; S.d[buf][t] = D; S.c[buf][t] = C; S.e[buf][t] = E; S.f[buf][t] = cg4[h];
                        t++;
                    }
                }
                if (lane == 0) S.last[buf] = last ? 1 : 0;
                mbar_arrive(bar);                                  // 32 arrivals publish the records; the bulk copies complete the transaction bytes
                next = next_after;
#pragma unroll
                for (int h = 0; h < RPL; h++) r[h] = nx[h];
                if (last) { c++; break; }
            }
        }
        return;
    }

    // ================= position threads
    const int max_cig = c_prm.max_cigar_ops;
    const int zz = PILE_LOCAL_TOTALS ? c_prm.reserved0 : 0;          // always 0
    unsigned long long dsum = 0; unsigned int dcnt = 0;
    int c = 0;
    for (int64_t tile = tile_begin; tile < tile_end; tile++) {
        const int64_t tile_lo = tile * TILE;
        const int64_t p = tile_lo + threadIdx.x;
        const bool live = p < P;
        const int rc4 = live ? ref_code((unsigned char)fasta[p]) : 16;
        const bool ref_acgt = (rc4 == 1 || rc4 == 2 || rc4 == 4 || rc4 == 8);
        const int rc4m = ref_acgt ? rc4 : 0x10;                                // 0x10: never equals a nibble
        const int wlo = (int)(tile_lo + (threadIdx.x & ~31));
        const int ip = (int)p;
        PileAcc a; PileRare x; PileTotals tt;
        a.m_hi = a.m_low = a.m_fs = a.m_pir = a.m_all = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) x.v[k] = 0;
        a.bq = a.bq_all = a.mq = a.mq_all = a.rd_mq = a.rd_rd = a.rd_low = 0; x.nm[0] = x.nm[1] = x.nm[2] = 0; x.nm_cnt = 0;
#pragma unroll
        for (int k = 0; k < F_COUNT; k++) tt.f[k] = 0;

        for (;; c++) {
            const int buf = c % NSTAGE;
            mbar_wait(smem_u32(&S.full[buf]), (uint32_t)((c / NSTAGE) & 1));
            const bool last = S.last[buf] != 0;
            const int2 rng = S.rng[buf][wid];
            const uint32_t kb = smem_u32(&S.seq[buf][0]) - (smem_u32(&S.qual[buf][0]) >> 1);     // nibble byte of quality address x: (x >> 1) + kb
            uint32_t acc_cov = 0, acc_all = 0, acc_hi = 0;
#pragma unroll 1                                          // measured: 1 beats 2 by 9 %, 3 and 4 are far slower (the slow paths are duplicated into the loop body)
            for (int t = rng.x; t < rng.y; t++) {
                const StageA A = S.a[buf][t];
                const StageB B = S.b[buf][t];
                // fast path for one staged read, hand-written so that every accumulate is a single predicated add
                int off, qv, nib, slow;
                asm volatile("{\n"
                    " .reg .pred ph, pm, pmh, ps;\n"
                    " .reg .b32 aq, as, sh, by;\n"
                    " sub.s32 %0, %10, %11;\n"                       // off = ip - pos
                    " setp.lt.u32 ph, %0, %12;\n"                    // hit = off <u lq_fast
                    " add.u32 aq, %13, %0;\n"
                    " @ph ld.shared.u8 %1, [aq];\n"                  // quality
                    " shr.u32 as, aq, 1;\n"
                    " add.u32 as, as, %15;\n"
                    " @ph ld.shared.u8 by, [as];\n"                  // two 4-bit base codes
                    " not.b32 sh, aq;\n"
                    " and.b32 sh, sh, 1;\n"
                    " shl.b32 sh, sh, 2;\n"
                    " shr.u32 %2, by, sh;\n"
                    " and.b32 %2, %2, 15;\n"                         // this base's code
                    " setp.eq.and.s32 pm, %2, %20, ph;\n"            // equals the (A/C/G/T) reference base
                    " setp.ge.and.s32 pmh, %1, %14, pm;\n"           // ... with mapq >= -q and base quality >= -b
                    " @ph add.u32 %4, %4, %16;\n"                    // covered: depth count | high-MAPQ count | MAPQ sum
                    " @pm add.s32 %5, %5, %1;\n"                     // bq_all
                    " @pm add.u32 %6, %6, %17;\n"                    // matches | MAPQ sum
                    " @pmh add.s32 %7, %7, %1;\n"                    // bq
                    " @pmh add.u32 %8, %8, %18;\n"                   // passing matches | forward | MAPQ sum
                    " @pmh add.s32 %9, %9, %19;\n"                   // position-in-read constant
                    " not.pred ps, pm;\n"
                    " and.pred ps, ps, ph;\n"                        // covered but not a plain match: generic rule
                    " setp.eq.or.u32 ps, %12, 0, ps;\n"              // ... or a read that needs the CIGAR walk
                    " selp.s32 %3, 1, 0, ps;\n"
                    "}"
                    : "=r"(off), "=r"(qv), "=r"(nib), "=r"(slow),
                      "+r"(acc_cov), "+r"(a.bq_all), "+r"(acc_all), "+r"(a.bq), "+r"(acc_hi), "+r"(a.m_pir)
                    : "r"(ip), "r"(A.pos), "r"(A.lq_fast), "r"(A.qa), "r"(A.bq_eff), "r"(kb), "r"(B.u_cov), "r"(B.u_all), "r"(B.u_hi), "r"(B.v_pir), "r"(rc4m));
                if (slow) {
                const StageD D = S.d[buf][t];
                const int r_mq = (int)(B.u_cov >> 16);
                if ((unsigned)off < A.lq_fast) {
                    const StageC C = S.c[buf][t];
                    pile_generic(a, x, tt, C.hash, (uint32_t)r_mq | ((D.flags & SF_REV) ? PR_REV : 0u) | ((D.flags & SF_NAMEOK) ? PR_NAMEOK : 0u), nib, qv, off, D.lq, rc4,
                                 qv >= A.bq_eff, min_snv, zz);
                }
                if (D.flags & SF_COMPLEX) {
                    // general CIGAR (or unstaged bases / depth bound not met): every lane walks the same op list; pileup offsets
                    // advance on M/=/X/D/N, CNV-depth offsets on M/=/X/D only (src/GROM.c:6621-6663), H extends lseq (6997-7000)
                    const StageC C = S.c[buf][t];
                    const StageE E = S.e[buf][t];
                    if (E.ext_end > wlo && live) {
                        const bool glob = D.flags & SF_GLOBAL;
                        const bool mq_ok = D.flags & SF_MQOK;
                        const uint32_t gmisc = (uint32_t)r_mq | ((D.flags & SF_REV) ? PR_REV : 0u) | ((D.flags & SF_NAMEOK) ? PR_NAMEOK : 0u);
                        const int ncig_all = (int)C.n_cigar, ncig = min(ncig_all, max_cig);
                        int qi = 0, rp = A.pos, rdp = A.pos, lseq = D.lq;
                        const bool staged_cigar = ncig_all <= 4;
                        const StageF F = S.f[buf][t];
                        for (int k = 0; k < ncig_all; k++) {
                            const uint32_t cg = staged_cigar ? (k == 0 ? F.op[0] : k == 1 ? F.op[1] : k == 2 ? F.op[2] : F.op[3]) : __ldg(R.cigar + C.cig_off + k);
                            const int op = cg & 15, len = (int)(cg >> 4);
                            const bool in_pile = k < ncig;
                            if (op == OP_M || op == OP_EQ || op == OP_X) {
                                const int od = ip - rdp;
                                if ((unsigned)od < (unsigned)len && rdp >= 0 && (int64_t)rdp + len < P) { tt.f[F_RDMQ + zz] += r_mq; tt.f[F_RDCNT + zz] += 1; tt.f[F_RDRD + zz] += (int)((B.u_cov >> 8) & 1u); }
                                rdp += len;
                                if (in_pile) {
                                    const int o = ip - rp;
                                    if ((unsigned)o < (unsigned)len && qi + o < D.lq) {
                                        const int xq = qi + o;
                                        int qv2, byte2;
                                        if (!glob) { qv2 = lds_u8(A.qa + xq); byte2 = lds_u8(((A.qa + xq) >> 1) + kb); }
                                        else { const uint64_t slot = ((uint64_t)E.base16 << 4) + (uint64_t)xq; qv2 = __ldg(R.qual + slot); byte2 = __ldg(R.seq4 + (slot >> 1)); }
                                        const int code2 = (byte2 >> ((~xq & 1) << 2)) & 15;
                                        pile_apply(a, x, tt, code2, qv2, C.hash, gmisc, xq, lseq, rc4, ref_acgt, mq_ok, bqmin, min_snv, (int)threadIdx.x, zz);
                                    }
                                    qi += len; rp += len;
                                }
                            } else if (op == OP_D) { rdp += len; if (in_pile) rp += len; }
                            else if (op == OP_N) { if (in_pile) rp += len; }
                            else if (op == OP_I || op == OP_S) { if (in_pile) qi += len; }
                            else if (op == OP_H) { if (in_pile) lseq += len; }
                        }
                    }
                }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&S.empty[buf]));
            // unpack the stage's packed counters into the 32-bit accumulators
            tt.f[F_RDCNT + zz] += (int)(acc_cov & 0xffu); tt.f[F_RDRD + zz] += (int)((acc_cov >> 8) & 0xffu); tt.f[F_RDMQ + zz] += (int)(acc_cov >> 16);
            tt.f[F_MALL + zz] += (int)(acc_all & 0xffffu); tt.f[F_MQALL + zz] += (int)(acc_all >> 16);
            tt.f[F_MHI + zz] += (int)(acc_hi & 0xffu); tt.f[F_MFS + zz] += (int)((acc_hi >> 8) & 0xffu); tt.f[F_MQ + zz] += (int)(acc_hi >> 16);
            if (last) { c++; break; }
        }
        a.m_hi += tt.f[F_MHI]; a.m_fs += tt.f[F_MFS]; a.m_all += tt.f[F_MALL]; a.mq += tt.f[F_MQ]; a.mq_all += tt.f[F_MQALL];
        a.rd_mq += tt.f[F_RDMQ]; a.rd_rd += tt.f[F_RDRD];
        const int rd_cnt = tt.f[F_RDCNT];
        int snv[4], low[4], pir[4], fs[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { snv[k] = x.v[RA_SNV + k]; low[k] = x.v[RA_LOW + k]; pir[k] = x.v[RA_PIR + k]; fs[k] = x.v[RA_FS + k]; }
        if (live) {
            // fold the matching-base registers into the per-base counters
            a.m_low = a.m_all - a.m_hi;
            a.m_pir += (int)threadIdx.x * (2 * a.m_fs - a.m_hi);               // position in read of the matching bases (see StageB::v_pir)
            a.rd_low += rd_cnt - a.rd_rd;
            const int rb = (rc4 == 1) ? 0 : (rc4 == 2) ? 1 : (rc4 == 4) ? 2 : 3;
#pragma unroll
            for (int k = 0; k < 4; k++) if (ref_acgt && rb == k) { snv[k] += a.m_hi; low[k] += a.m_low; pir[k] += a.m_pir; fs[k] += a.m_fs; }
            int32_t *o = arrays + p;
            const int tot = snv[0] + snv[1] + snv[2] + snv[3];
            const int lowt = low[0] + low[1] + low[2] + low[3];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                o[(int64_t)(GA_SNV_A + k) * Ppad] = snv[k]; o[(int64_t)(GA_SNVLOW_A + k) * Ppad] = low[k];
                o[(int64_t)(GA_PIR_A + k) * Ppad] = pir[k]; o[(int64_t)(GA_FS_A + k) * Ppad] = fs[k];
            }
            o[(int64_t)GA_BQ * Ppad] = a.bq; o[(int64_t)GA_BQ_ALL * Ppad] = a.bq_all;
            o[(int64_t)GA_MQ * Ppad] = a.mq; o[(int64_t)GA_MQ_ALL * Ppad] = a.mq_all;
            o[(int64_t)GA_BQ_RC * Ppad] = tot; o[(int64_t)GA_MQ_RC * Ppad] = tot; o[(int64_t)GA_RC_ALL * Ppad] = tot + lowt;
            o[(int64_t)GA_RD_MQ * Ppad] = a.rd_mq; o[(int64_t)GA_RD_RD * Ppad] = a.rd_rd; o[(int64_t)GA_RD_LOW * Ppad] = a.rd_low;
        }
        // ---- SNV gate on the counts still in registers (src/GROM.c:11096-11199); candidates are compacted by warp ballot
        bool is_cand = false;
        int c_base = 0; double c_ratio = 0, c_pr = 0, c_hez = 0;
        const int total = snv[0] + snv[1] + snv[2] + snv[3];
        const int rc_all = total + low[0] + low[1] + low[2] + low[3];
        if (live) {
            const char fc = fasta[p];
            const bool is_n = (fc == 'N' || fc == 'n');
            if (p < sc.depth_bound && !is_n) { dsum += (unsigned long long)((long long)a.rd_rd + (long long)a.rd_low); dcnt += 1; }
            if (ip >= sc.scan_first && ip <= sc.scan_last && !is_n) {
                bool any = false;
#pragma unroll
                for (int k = 0; k < 4; k++) any = any || (snv[k] >= c_prm.min_snv && rc4 != (1 << k));
                if (any && arrays[(int64_t)GA_RD * Ppad + p] + arrays[(int64_t)GA_INDEL_SC_RD * Ppad + p] > 0) {
                    const bool bq_ok = (double)a.bq_all / (double)rc_all >= c_prm.min_ave_bq;
                    const int T = c_prm.max_trials, TD = T + 1;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const double ratio = (double)((float)snv[k] / (float)total);
                        if (rc4 != (1 << k) && ratio >= c_prm.min_snv_ratio && snv[k] >= c_prm.min_snv && bq_ok) {
                            if (!is_cand || ratio > c_ratio) {
                                const size_t idx = (total > T) ? (size_t)T * TD + (size_t)(snv[k] * T / total) : (size_t)total * TD + (size_t)snv[k];
                                c_base = k; c_ratio = ratio; c_pr = sc.mqt[idx]; c_hez = sc.hez[idx];
                                is_cand = true;
                            }
                        }
                    }
                }
            }
        }
        const unsigned ball = __ballot_sync(0xffffffffu, is_cand);
        if (ball) {
            unsigned basei = 0;
            if (lane == 0) basei = atomicAdd(sc.n_cand, (unsigned)__popc(ball));
            basei = __shfl_sync(0xffffffffu, basei, 0);
            const unsigned slot = basei + (unsigned)__popc(ball & ((1u << lane) - 1u));
            if (is_cand && slot < sc.cand_cap) {
                grom_snv_cand *cd = sc.cand + slot;
                cd->pos = ip; cd->base = c_base; cd->ratio = c_ratio; cd->pr = c_pr; cd->hez = c_hez; cd->reserved = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) { cd->v[GA_SNV_A + k] = snv[k]; cd->v[GA_SNVLOW_A + k] = low[k]; cd->v[GA_PIR_A + k] = pir[k]; cd->v[GA_FS_A + k] = fs[k]; }
                cd->v[GA_BQ] = a.bq; cd->v[GA_BQ_ALL] = a.bq_all; cd->v[GA_MQ] = a.mq; cd->v[GA_MQ_ALL] = a.mq_all;
                cd->v[GA_BQ_RC] = total; cd->v[GA_MQ_RC] = total; cd->v[GA_RC_ALL] = rc_all;
            }
        }
    }
    // depth sum for the SNV emission filter: one atomic pair per consumer warp per CTA (SUBTILES tiles)
    for (int d = 16; d; d >>= 1) { dsum += __shfl_xor_sync(0xffffffffu, dsum, d); dcnt += __shfl_xor_sync(0xffffffffu, dcnt, d); }
    if (lane == 0 && dcnt) { atomicAdd(sc.depth_sum, dsum); atomicAdd(sc.depth_sum + 1, (unsigned long long)dcnt); }
}

// ---- small-indel gates (src/GROM.c:11329-11745) over the (few) positions whose insertion / deletion-start / deletion-end
// slot carries enough weight; the position list is compacted by k_sv_apply, the depth comes from the pileup arrays
__global__ void __launch_bounds__(128) k_indel_gate(const int2 *__restrict__ pos_list, const unsigned int *__restrict__ n_list, unsigned int list_cap,
                                                     DevReads R, int64_t P, int64_t Ppad, const int32_t *__restrict__ arrays, SnvScanArgs sc)
{
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= min(*n_list, list_cap)) return;
    const int2 pk = pos_list[i];
    const int ip = pk.x;
    const int64_t p = ip;
    if (ip < sc.scan_first || ip > sc.scan_last) return;
    if (!(arrays[(int64_t)GA_RD * Ppad + p] + arrays[(int64_t)GA_INDEL_SC_RD * Ppad + p] > 0)) return;
    const int af = c_prm.add_factor;
    const int base = arrays[(int64_t)GA_RC_ALL * Ppad + p];             // sum of snv + snv_lowmq
    const int TD = c_prm.max_trials + 1;
    const int scl = arrays[(int64_t)GA_INDEL_SC_LEFT * Ppad + p], scr = arrays[(int64_t)GA_INDEL_SC_RIGHT * Ppad + p];
    if (pk.y & 1) {                                                     // insertion, src/GROM.c:11329-11453
        int it = arrays[(int64_t)GA_INDEL_I * Ppad + p];
        const int rdt = base;
        if (it / af > rdt) it = rdt * af;
        if (it / af >= c_prm.min_disc && rdt <= c_prm.max_trials) {
            const double pr = sc.mqt[(size_t)rdt * TD + it / af];
            double hz;
            if ((it + scl) / af < rdt) {
                hz = sc.hez[(size_t)rdt * TD + (it + scl) / af];
                if ((it + scr) / af < rdt) { const double h2 = sc.hez[(size_t)rdt * TD + (it + scr) / af]; if (h2 > hz) hz = h2; }
                else hz = sc.hez[(size_t)rdt * TD + rdt];
            } else hz = sc.hez[(size_t)rdt * TD + rdt];
            if (pr <= c_prm.pval_threshold1) {
                const unsigned slot = atomicAdd(sc.n_ins, 1u);
                if (slot < sc.ins_cap) {
                    grom_ins_cand *c = sc.ins + slot;
                    c->pos = ip; c->dist = arrays[(int64_t)GA_INDEL_IDIST * Ppad + p]; c->pr = pr; c->hez = hz;
                    c->conc = arrays[(int64_t)GA_CONC * Ppad + p]; c->weight = it; c->rd = rdt;
                    c->sc = (p + 1 < P ? arrays[(int64_t)GA_SC_LEFT * Ppad + p + 1] : 0) + arrays[(int64_t)GA_SC_RIGHT * Ppad + p];
                    c->other_len = sc.other_len[p]; c->reserved = 0;
                    for (int k = 0; k < 56; k++) c->seq[k] = 0;
                    if (c->dist <= c_prm.indel_i_seq_len) {
                        // characters of the creation that is still visible; longer leftovers of an earlier zero-weight creation
                        // (the reference keeps them, src/GROM.c:7219-7228) are not tracked on the device
                        const int64_t ri = sc.ins_src[p]; const int qo = sc.ins_src[Ppad + p], sl = min(sc.ins_src[2 * Ppad + p], c->dist);
                        const uint64_t bo = R.base_off[ri];
                        for (int k = 0; k < sl; k++) {
                            const uint64_t slot2 = bo + (uint64_t)(qo + k);
                            c->seq[k] = "=ACMGRSVTWYHKDBN"[(R.seq4[slot2 >> 1] >> ((~(int)slot2 & 1) << 2)) & 15];
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int kind = 0; kind < 2; kind++) {                             // deletion start / end, src/GROM.c:11454-11745
        if (!(pk.y & (2 << kind))) continue;
        const int wt = arrays[(int64_t)(kind ? GA_INDEL_D_R : GA_INDEL_D_F) * Ppad + p];
        const int rdt = wt / af + base;
        if (!(wt / af >= c_prm.min_disc) || rdt > c_prm.max_trials) continue;
        const int scv = kind ? scl : scr;
        const double pr = sc.mqt[(size_t)rdt * TD + wt / af];
        const double hz = ((wt + scv) / af < rdt) ? sc.hez[(size_t)rdt * TD + (wt + scv) / af] : sc.hez[(size_t)rdt * TD + rdt];
        if (!(pr <= c_prm.pval_threshold1)) continue;
        const unsigned slot = atomicAdd(sc.n_del, 1u);
        if (slot < sc.del_cap) {
            grom_del_event *e = sc.del_ev + slot;
            e->pos = ip; e->kind = kind; e->pr = pr; e->hez = hz; e->conc = arrays[(int64_t)GA_CONC * Ppad + p]; e->weight = wt; e->rd = rdt;
            e->sc = arrays[(int64_t)(kind ? GA_SC_LEFT : GA_SC_RIGHT) * Ppad + p]; e->other_len = sc.other_len[p];
            e->rdist = arrays[(int64_t)GA_INDEL_D_RDIST * Ppad + p];
        }
    }
}

// ---- structural-variant gates (src/GROM.c:11963-13541): the ten breakpoint classes at the positions k_sv_apply listed.  One gate per
// (position, class): weight, freshness of the supporting reads relative to the scanned position (reverse classes measure it with the
// length of the look-ahead read, i.e. the first read that starts more than ins_max beyond the position), binomial tail from the mq
// table, side-evidence-corrected tail from the hez table.  Events are compacted with an atomic counter and sorted on the host.
struct SvGateArgs {
    const int32_t *cl_w, *cl_rs, *cl_re, *cl_mchr, *other_len; const double *cl_dist;
    grom_sv_event *ev; unsigned int cap; unsigned int *n_ev;
    int64_t i0, n_reads; int last_lseq, last_lseq_applied; const uint8_t *state;
};
__device__ __forceinline__ void sv_emit(const SvGateArgs &G, int pos, int cls, double bin, double hez, double dist, int w, int rd, int conc, int rs, int re, int ol, int mchr, int dsum = 0)
{
    const unsigned int k = atomicAdd(G.n_ev, 1u);
    if (k >= G.cap) return;
    grom_sv_event e;
    e.pos = pos; e.cls = cls; e.binom = bin; e.hez = hez; e.dist = dist; e.weight = w; e.rd = rd; e.conc = conc; e.read_start = rs; e.read_end = re;
    e.other_len = ol; e.mchr = mchr; e.reserved = dsum;
    G.ev[k] = e;
}
__global__ void __launch_bounds__(128) k_sv_gate(const int2 *__restrict__ pos_list, const unsigned int *__restrict__ n_list, unsigned int list_cap,
                                                  DevReads R, int64_t P, int64_t Ppad, const int32_t *__restrict__ arrays, SnvScanArgs sc, SvGateArgs G)
{
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= min(*n_list, list_cap)) return;
    const int2 pk = pos_list[i];
    const int ip = pk.x, cm = pk.y >> 3;
    const int64_t p = ip;
    if (!cm || ip < sc.scan_first || ip > sc.scan_last) return;
    const int rd = arrays[(int64_t)GA_RD * Ppad + p];
    if (rd <= 0) return;
    const int af = c_prm.add_factor, mt = c_prm.max_trials, TD = mt + 1;
    const int side_f = arrays[(int64_t)GA_SC_RIGHT * Ppad + p] + arrays[(int64_t)GA_MUNMAPPED_F * Ppad + p];
    const int side_r = arrays[(int64_t)GA_SC_LEFT * Ppad + p] + arrays[(int64_t)GA_MUNMAPPED_R * Ppad + p];
    int la = -1;                                                      // look-ahead read length, found on first use
#pragma unroll 1
    for (int c = 0; c < 10; c++) {
        if (!((cm >> c) & 1)) continue;
        const bool fwd = !(c & 1);                                    // del_f dup_f inv_f1 inv_f2 ctx_f sit on even class numbers
        const int w = G.cl_w[(int64_t)c * Ppad + p], rs = G.cl_rs[(int64_t)c * Ppad + p], re = G.cl_re[(int64_t)c * Ppad + p];
        if (fwd) { if (!(ip - re < c_prm.insert_mean)) continue; }
        else {
            if (la < 0) {
                // first read index >= i0 whose start exceeds position + ins_max
                int64_t lo = G.i0, hi = G.n_reads;
                const int key = ip + c_prm.overlap_mult * c_prm.insert_max;
                while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (R.pos[m] <= key) lo = m + 1; else hi = m; }
                la = lo < G.n_reads ? R.l_qseq[lo] : (G.n_reads > 0 && G.state[G.n_reads - 1] == 1 ? G.last_lseq_applied : G.last_lseq);
            }
            if (!(rs + la - ip < c_prm.insert_mean)) continue;
        }
        const int side = fwd ? side_f : side_r;
        double bin, hz = 2.0;
        if (rd > mt) {
            bin = sc.mqt[(size_t)mt * TD + w * mt / (af * rd)];
            if ((double)((float)side / (float)w) <= c_prm.max_evidence_ratio)
                hz = ((w + side) / af < rd) ? sc.hez[(size_t)mt * TD + (w + side) * mt / (af * rd)] : sc.hez[(size_t)mt * TD + mt];
        } else {
            bin = sc.mqt[(size_t)rd * TD + w / af];
            // ctx_r tests the ctx_f ratio in this branch (src/GROM.c:12074)
            const float ratio = c == GROM_SV_CTX_R ? (float)side_f / (float)G.cl_w[(int64_t)GROM_SV_CTX_F * Ppad + p] : (float)side / (float)w;
            if ((double)ratio <= c_prm.max_evidence_ratio)
                hz = ((w + side) / af < rd) ? sc.hez[(size_t)rd * TD + (w + side) / af] : sc.hez[(size_t)rd * TD + rd];
        }
        if (bin <= c_prm.pval_threshold1) {
            int dsum = 0;
            if (c >= GROM_SV_INV_F1 && c <= GROM_SV_INV_R2)          // depth around the breakpoint (src/GROM.c:15921-15934)
                for (int64_t y = max(rs, 0); y < min((int64_t)re + c_prm.lseq, P); y++) dsum += arrays[(int64_t)GA_RD_RD * Ppad + y] + arrays[(int64_t)GA_RD_LOW * Ppad + y];
            sv_emit(G, ip, c, bin, hz, G.cl_dist[(int64_t)c * Ppad + p], w, rd, arrays[(int64_t)GA_CONC * Ppad + p], rs, re, G.other_len[p],
                    c >= GROM_SV_CTX_F ? G.cl_mchr[(int64_t)(c - GROM_SV_CTX_F) * Ppad + p] : 0, dsum);
        }
    }
}
// insertion gates (src/GROM.c:11750-11961): soft-clip weight + short-pair range adds on either side, dense over the scanned range
__global__ void __launch_bounds__(256) k_ins_sv_gate(int64_t Ppad, const int32_t *__restrict__ arrays, SnvScanArgs sc, SvGateArgs G)
{
    const int64_t p = sc.scan_first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p > sc.scan_last) return;
    const int af = c_prm.add_factor, md = c_prm.min_disc;
    const int ins = arrays[(int64_t)GA_INS * Ppad + p], scl = arrays[(int64_t)GA_SC_LEFT * Ppad + p], scr = arrays[(int64_t)GA_SC_RIGHT * Ppad + p];
    const bool l_ok = (scl + ins) / af >= md, r_ok = (scr + ins) / af >= md;
    if (!l_ok && !r_ok) return;
    const int rd = arrays[(int64_t)GA_RD * Ppad + p];
    if (!(rd + arrays[(int64_t)GA_SC_RD * Ppad + p] > 0)) return;
    const int mt = c_prm.max_trials, TD = mt + 1;
    for (int side = 0; side < 2; side++) {
        if (!(side ? r_ok : l_ok)) continue;
        const int sv = side ? scr : scl, scrd = rd + arrays[(int64_t)(side ? GA_SC_RIGHT_RD : GA_SC_LEFT_RD) * Ppad + p];
        if (scrd > mt) continue;
        const int mu = arrays[(int64_t)(side ? GA_MUNMAPPED_F : GA_MUNMAPPED_R) * Ppad + p];
        const double bin = ((mu + sv + ins) / af < scrd) ? sc.mqt[(size_t)scrd * TD + (mu + sv + ins) / af] : sc.mqt[(size_t)scrd * TD + scrd];
        if (bin <= c_prm.pval_insertion1)
            sv_emit(G, (int)p, GROM_SV_INS_L + side, bin, 2.0, 0.0, ins, rd, arrays[(int64_t)GA_CONC * Ppad + p], 0, 0, G.other_len[p], 0);
    }
}

// ---- single-pass inclusive prefix sum, in place (decoupled look-back).  4096 elements per CTA.
#define SCAN_THREADS 256
#define SCAN_ITEMS 16
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)
struct ScanList { int64_t off[8]; };     // element offsets of up to 8 arrays scanned by one launch (blockIdx.y selects the array)
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_inplace(int32_t *data_base, ScanList list, int64_t n, unsigned long long *status_base, unsigned int *ticket_base)
{
    __shared__ int s_tile, s_warp[SCAN_THREADS / 32], s_excl;
    int32_t *data = data_base + list.off[blockIdx.y];
    unsigned long long *status = status_base + (size_t)blockIdx.y * gridDim.x;
    unsigned int *ticket = ticket_base + blockIdx.y;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = s_tile;
    const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k += 4) {
        if (base + k + 3 < n) { const int4 x = *reinterpret_cast<const int4 *>(data + base + k); v[k] = x.x; v[k + 1] = x.y; v[k + 2] = x.z; v[k + 3] = x.w; }
        else { for (int j = 0; j < 4; j++) v[k + j] = (base + k + j < n) ? data[base + k + j] : 0; }
    }
#pragma unroll
    for (int k = 1; k < SCAN_ITEMS; k++) v[k] += v[k - 1];
    const int mine = v[SCAN_ITEMS - 1];
    int incl = mine;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = (lane < SCAN_THREADS / 32) ? s_warp[lane] : 0;
#pragma unroll
        for (int d = 1; d < SCAN_THREADS / 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += y; }
        if (lane < SCAN_THREADS / 32) s_warp[lane] = w;       // inclusive warp totals
    }
    __syncthreads();
    const int warp_excl = wid ? s_warp[wid - 1] : 0;
    const int tile_sum = s_warp[SCAN_THREADS / 32 - 1];
    if (wid == 0) {
        // decoupled look-back by one warp: 32 predecessors per round trip (lane l looks at tile - 1 - l), up to the nearest tile
        // whose inclusive prefix is known
        if (lane == 0) atomicExch(status + tile, ((tile == 0 ? 2ull : 1ull) << 32) | (unsigned int)tile_sum);
        int excl = 0;
        if (tile > 0) {
            for (int j = tile - 1;; j -= 32) {
                const int idx = j - lane;
                unsigned long long st = 2ull << 32;                       // before the first tile: inclusive prefix 0
                if (idx >= 0) { do { st = *((volatile unsigned long long *)(status + idx)); } while ((st >> 32) == 0); }
                const unsigned done = __ballot_sync(0xffffffffu, (st >> 32) == 2);
                const int stop = done ? __ffs(done) - 1 : 31;             // nearest predecessor with an inclusive prefix
                int part = lane <= stop ? (int)(unsigned int)st : 0;
#pragma unroll
                for (int d = 16; d; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                excl += part;
                if (done) break;
            }
            if (lane == 0) { __threadfence(); atomicExch(status + tile, (2ull << 32) | (unsigned int)(excl + tile_sum)); }
        }
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    const int off = s_excl + warp_excl + (incl - mine);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k += 4) {
        if (base + k + 3 < n) *reinterpret_cast<int4 *>(data + base + k) = make_int4(v[k] + off, v[k + 1] + off, v[k + 2] + off, v[k + 3] + off);
        else { for (int j = 0; j < 4; j++) if (base + k + j < n) data[base + k + j] = v[k + j] + off; }
    }
}

// ---- K5a: GC / ACGT percentage of the triangular window (src/GROM.c:1766-1859).  count(r) = sum_{|d|<M} (M-|d|) is(r+d)
// is a second difference of the double prefix sum of the indicator, so each CTA scans a tile + halo of the FASTA in
// shared memory twice (S1, then S2; tile-local constants cancel in the symmetric second difference).
#define GC_TILE 2048
#define GC_THREADS 256
__device__ __forceinline__ int block_excl_scan_256(int v, int *s_w)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
    __syncthreads();
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    int base = 0;
    for (int k = 0; k < w; k++) base += s_w[k];
    return base + incl - v;
}

__global__ void __launch_bounds__(GC_THREADS) k_gc_prepass(const char *__restrict__ fasta, int64_t P, int M, int32_t *__restrict__ gc_out, int32_t *__restrict__ acgt_out)
{
    extern __shared__ int gc_smem[];
    __shared__ int s_w[GC_THREADS / 32];
    const int L = GC_TILE + 2 * M;                 // local index i <-> position tile_lo - M + i
    int *sg = gc_smem, *sa = gc_smem + (L + 1);    // inclusive prefix arrays shifted by one (element 0 = 0)
    const int64_t tile_lo = (int64_t)blockIdx.x * GC_TILE;
    int seg = (L + GC_THREADS - 1) / GC_THREADS; seg |= 1;          // odd segment length: conflict-free strided access
    const int i0 = threadIdx.x * seg, i1 = min(i0 + seg, L);
    // pass 0: indicators -> S1
    int tg = 0, ta = 0;
    for (int i = i0; i < i1; i++) {
        const int64_t x = tile_lo - M + i;
        int g = 0, a = 0;
        if (x >= 0 && x < P) { const char c = fasta[x]; g = (c == 'C' || c == 'G' || c == 'c' || c == 'g'); a = g || (c == 'A' || c == 'T' || c == 'a' || c == 't'); }
        tg += g; ta += a; sg[i + 1] = tg; sa[i + 1] = ta;
    }
    int og = block_excl_scan_256(tg, s_w);
    int oa = block_excl_scan_256(ta, s_w);
    // pass 1: S1 -> S2 (in place)
    tg = 0; ta = 0;
    for (int i = i0; i < i1; i++) { tg += sg[i + 1] + og; ta += sa[i + 1] + oa; sg[i + 1] = tg; sa[i + 1] = ta; }
    og = block_excl_scan_256(tg, s_w);
    oa = block_excl_scan_256(ta, s_w);
    for (int i = i0; i < i1; i++) { sg[i + 1] += og; sa[i + 1] += oa; }
    if (threadIdx.x == 0) { sg[0] = 0; sa[0] = 0; }
    __syncthreads();
    const long long total = (long long)M * M;
    const int64_t valid_lo = M - 1, valid_hi = P - (2 * (int64_t)M - 1);
    for (int k = threadIdx.x; k < GC_TILE; k += GC_THREADS) {
        const int64_t r = tile_lo + k;
        if (r >= P) break;
        int vg = 0, va = 0;
        if (r >= valid_lo && r < valid_hi) {
            const int ir = k + M;                   // local index of r; S2 inclusive at local j is s[j + 1]
            const long long cg = (long long)(sg[ir + M] - sg[ir]) - (long long)(sg[ir] - sg[ir - M]);
            const long long ca = (long long)(sa[ir + M] - sa[ir]) - (long long)(sa[ir] - sa[ir - M]);
            vg = (int)(100 * cg / total); va = (int)(100 * ca / total);
        }
        gc_out[r] = vg; acgt_out[r] = va;
    }
}

#include "sv_evidence.cuh"
#include "cnv.cuh"

__global__ void k_fix_offsets(uint64_t *cigar_off, uint64_t *base_off, int64_t n, uint64_t cig_base, uint64_t slot_base)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { cigar_off[i] += cig_base; base_off[i] += slot_base; }
}

// ---- transport-compact forms of the read batch (include/grom_reads.h: GROM_LAYOUT_*): the canonical device arrays are rebuilt here
// canonical offsets: cigar_off = exclusive sum of n_cigar, base_off = exclusive sum of l_qseq rounded up to GROM_BASE_ALIGN
#define OFF_BLOCK 1024
__device__ __forceinline__ uint64_t pad_slots(int l) { return ((uint64_t)(l > 0 ? l : 0) + GROM_BASE_ALIGN - 1) / GROM_BASE_ALIGN * GROM_BASE_ALIGN; }
__device__ __forceinline__ void block_excl_scan2(uint64_t &a, uint64_t &b, uint64_t *sh /* [2 * 32 + 2] */, uint64_t &tot_a, uint64_t &tot_b)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint64_t ia = a, ib = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint64_t ya = __shfl_up_sync(0xffffffffu, ia, d), yb = __shfl_up_sync(0xffffffffu, ib, d); if (lane >= d) { ia += ya; ib += yb; } }
    if (lane == 31) { sh[wid] = ia; sh[32 + wid] = ib; }
    __syncthreads();
    if (wid == 0) {
        uint64_t wa = lane < nw ? sh[lane] : 0, wb = lane < nw ? sh[32 + lane] : 0;
        const uint64_t oa = wa, ob = wb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint64_t ya = __shfl_up_sync(0xffffffffu, wa, d), yb = __shfl_up_sync(0xffffffffu, wb, d); if (lane >= d) { wa += ya; wb += yb; } }
        if (lane < nw) { sh[lane] = wa - oa; sh[32 + lane] = wb - ob; }
        if (lane == 31) { sh[64] = wa; sh[65] = wb; }
    }
    __syncthreads();
    a = ia - a + sh[wid]; b = ib - b + sh[32 + wid];
    tot_a = sh[64]; tot_b = sh[65];
    __syncthreads();
}
__global__ void __launch_bounds__(256) k_off_sums(const uint16_t *__restrict__ n_cigar, const int32_t *__restrict__ l_qseq, int64_t n, uint64_t *__restrict__ bsum)
{
    __shared__ uint64_t sh[66];
    const int64_t i0 = (int64_t)blockIdx.x * OFF_BLOCK + threadIdx.x * 4;
    uint64_t a = 0, b = 0, ta, tb;
    for (int k = 0; k < 4; k++) if (i0 + k < n) { a += n_cigar[i0 + k]; b += pad_slots(l_qseq[i0 + k]); }
    block_excl_scan2(a, b, sh, ta, tb);
    if (threadIdx.x == 0) { bsum[2 * (int64_t)blockIdx.x] = ta; bsum[2 * (int64_t)blockIdx.x + 1] = tb; }
}
__global__ void __launch_bounds__(1024) k_off_scan(uint64_t *bsum, int64_t nblk)
{
    __shared__ uint64_t sh[66];
    const int64_t per = (nblk + blockDim.x - 1) / blockDim.x, j0 = (int64_t)threadIdx.x * per, j1 = min(j0 + per, nblk);
    uint64_t a = 0, b = 0, ta, tb;
    for (int64_t j = j0; j < j1; j++) { a += bsum[2 * j]; b += bsum[2 * j + 1]; }
    block_excl_scan2(a, b, sh, ta, tb);
    for (int64_t j = j0; j < j1; j++) { const uint64_t va = bsum[2 * j], vb = bsum[2 * j + 1]; bsum[2 * j] = a; bsum[2 * j + 1] = b; a += va; b += vb; }
}
__global__ void __launch_bounds__(256) k_off_write(const uint16_t *__restrict__ n_cigar, const int32_t *__restrict__ l_qseq, int64_t n, const uint64_t *__restrict__ bsum,
                                                   uint64_t cig_base, uint64_t slot_base, uint64_t *__restrict__ cigar_off, uint64_t *__restrict__ base_off)
{
    __shared__ uint64_t sh[66];
    const int64_t i0 = (int64_t)blockIdx.x * OFF_BLOCK + threadIdx.x * 4;
    uint64_t va[4], vb[4], a = 0, b = 0, ta, tb;
    for (int k = 0; k < 4; k++) { va[k] = vb[k] = 0; if (i0 + k < n) { va[k] = n_cigar[i0 + k]; vb[k] = pad_slots(l_qseq[i0 + k]); } a += va[k]; b += vb[k]; }
    block_excl_scan2(a, b, sh, ta, tb);
    a += cig_base + bsum[2 * (int64_t)blockIdx.x]; b += slot_base + bsum[2 * (int64_t)blockIdx.x + 1];
    for (int k = 0; k < 4; k++) if (i0 + k < n) { cigar_off[i0 + k] = a; base_off[i0 + k] = b; a += va[k]; b += vb[k]; }
}
// 4-bit dictionary-coded qualities -> one byte per base slot (16 slots per thread)
struct QualLut { uint8_t v[16]; };
__global__ void __launch_bounds__(256) k_expand_qual(const uint8_t *__restrict__ q4, int64_t n_slots, QualLut lut, uint8_t *__restrict__ qual)
{
    __shared__ uint8_t sl[16];
    if (threadIdx.x < 16) sl[threadIdx.x] = lut.v[threadIdx.x];
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, s0 = g * 16;
    if (s0 >= n_slots) return;
    if (s0 + 16 <= n_slots) {
        const uint2 w = *reinterpret_cast<const uint2 *>(q4 + (s0 >> 1));
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t half = (k < 2 ? w.x : w.y) >> ((k & 1) * 16);          // two source bytes = four slots
            const uint32_t b0 = half & 0xff, b1 = (half >> 8) & 0xff;
            o[k] = (uint32_t)sl[b0 >> 4] | ((uint32_t)sl[b0 & 15] << 8) | ((uint32_t)sl[b1 >> 4] << 16) | ((uint32_t)sl[b1 & 15] << 24);
        }
        *reinterpret_cast<uint4 *>(qual + s0) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
        for (int64_t t = s0; t < n_slots; t++) qual[t] = sl[(q4[t >> 1] >> ((~t & 1) << 2)) & 15];
    }
}
// 2-bit dictionary-coded qualities -> one byte per base slot (16 slots per thread); padding slots are zeroed by k_seq_zero_pad
__global__ void __launch_bounds__(256) k_expand_qual2(const uint8_t *__restrict__ q2, int64_t n_slots, QualLut lut, uint8_t *__restrict__ qual)
{
    __shared__ uint32_t four[256];                                                    // one source byte (four slots, first in the top bits) -> four quality bytes
    {
        const uint32_t b = threadIdx.x;
        four[b] = (uint32_t)lut.v[b >> 6] | ((uint32_t)lut.v[(b >> 4) & 3] << 8) | ((uint32_t)lut.v[(b >> 2) & 3] << 16) | ((uint32_t)lut.v[b & 3] << 24);
    }
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, s0 = g * 16;
    if (s0 >= n_slots) return;
    if (s0 + 16 <= n_slots) {
        const uint32_t w = *reinterpret_cast<const uint32_t *>(q2 + (s0 >> 2));
        *reinterpret_cast<uint4 *>(qual + s0) = make_uint4(four[w & 0xff], four[(w >> 8) & 0xff], four[(w >> 16) & 0xff], four[w >> 24]);
    } else {
        for (int64_t t = s0; t < n_slots; t++) qual[t] = lut.v[(q2[t >> 2] >> ((~t & 3) << 1)) & 3];
    }
}
// 2-bit bases -> BAM nibbles (16 slots per thread), exceptions (non-ACGT codes) patched in, padding slots of every read zeroed
__global__ void __launch_bounds__(256) k_expand_seq(const uint8_t *__restrict__ s2, int64_t n_slots, uint8_t *__restrict__ seq4)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, s0 = g * 16;
    if (s0 >= n_slots) return;
    if (s0 + 16 <= n_slots) {
        const uint32_t w = *reinterpret_cast<const uint32_t *>(s2 + (s0 >> 2));
        uint32_t o[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            uint32_t acc = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t b = (w >> (8 * (2 * k + (j >> 1)))) & 0xff;              // source byte: four slots, first in the top bits
                const uint32_t two = (j & 1) ? (b & 15) : (b >> 4);                     // two slots
                acc |= (((1u << (two >> 2)) << 4) | (1u << (two & 3))) << (8 * j);
            }
            o[k] = acc;
        }
        *reinterpret_cast<uint2 *>(seq4 + (s0 >> 1)) = make_uint2(o[0], o[1]);
    } else {
        for (int64_t t = s0; t < n_slots; t += 2) {
            const uint32_t c0 = (s2[t >> 2] >> ((~t & 3) << 1)) & 3, c1 = t + 1 < n_slots ? (s2[(t + 1) >> 2] >> ((~(t + 1) & 3) << 1)) & 3 : 0;
            seq4[t >> 1] = (uint8_t)(((1u << c0) << 4) | (t + 1 < n_slots ? (1u << c1) : 0u));
        }
    }
}
__global__ void __launch_bounds__(256) k_seq_exceptions(const uint64_t *__restrict__ slot, const uint8_t *__restrict__ code, int64_t n_exc, int64_t n_slots, uint8_t *__restrict__ seq4)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_exc) return;
    const uint64_t s = slot[k];
    if (s >= (uint64_t)n_slots) return;
    // neighbouring exceptions share bytes: atomics on the aligned word (seq4 regions start 16-byte aligned)
    unsigned int *w = reinterpret_cast<unsigned int *>(seq4 + ((s >> 1) & ~(uint64_t)3));
    const unsigned sh = (unsigned)(((s >> 1) & 3) * 8 + ((~s & 1) << 2));
    atomicAnd(w, ~(15u << sh));
    atomicOr(w, ((unsigned)code[k] & 15u) << sh);
}
__global__ void __launch_bounds__(256) k_seq_zero_pad(const int32_t *__restrict__ l_qseq, const uint64_t *__restrict__ base_off, int64_t n, uint8_t *__restrict__ seq4,
                                                      uint8_t *__restrict__ qual)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t b = base_off[i], e = b + pad_slots(l_qseq[i]);
    uint64_t s = b + (uint64_t)max(l_qseq[i], 0);
    if (qual) {                                                                        // this read owns every slot of [b, e)
        uint64_t t = s;
        for (; t < e && (t & 3); t++) qual[t] = 0;
        for (; t + 4 <= e; t += 4) *reinterpret_cast<uint32_t *>(qual + t) = 0u;
        for (; t < e; t++) qual[t] = 0;
    }
    if (!seq4) return;
    if (s < e && (s & 1)) { seq4[s >> 1] &= 0xf0; s++; }
    for (; s < e; s += 2) seq4[s >> 1] = 0;
}
// sparse first-SA-entry fields -> dense per-read arrays (already preset to "none")
struct SaSparse { const int32_t *idx, *pos, *sadj, *eadj, *indel; const int16_t *mapq; const uint8_t *strand, *same; };
__global__ void __launch_bounds__(256) k_scatter_sa(SaSparse S, int64_t n_sa, int64_t n, int32_t *pos, int32_t *sadj, int32_t *eadj, int32_t *indel, uint8_t *strand, int16_t *mapq, uint8_t *same)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_sa) return;
    const int64_t i = S.idx[k];
    if (i < 0 || i >= n) return;
    pos[i] = S.pos[k]; sadj[i] = S.sadj[k]; eadj[i] = S.eadj[k]; indel[i] = S.indel[k]; strand[i] = S.strand[k]; mapq[i] = S.mapq[k]; same[i] = S.same[k];
}

// ------------------------------------------------------------------------------------------------ host side
struct DevBuf {
    void *p = nullptr; size_t cap = 0, size = 0;
    int ensure(size_t need, cudaStream_t s)
    {
        if (need <= cap) return 0;
        size_t ncap = std::max(need, cap + cap / 2);
        ncap = (ncap + 255) & ~(size_t)255;
        void *np = nullptr;
        CK(cudaMalloc(&np, ncap));
        if (p && size) CK(cudaMemcpyAsync(np, p, size, cudaMemcpyDeviceToDevice, s));
        if (p) { CK(cudaStreamSynchronize(s)); CK(cudaFree(p)); }
        p = np; cap = ncap;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = size = 0; }
};

enum { B_POS, B_MPOS, B_TLEN, B_MTID, B_LQSEQ, B_FLAG, B_NCIGAR, B_MAPQ, B_QLEN, B_HASH, B_CIGOFF, B_BASEOFF, B_CIGAR, B_SEQ4, B_QUAL,
       B_SAPOS, B_SASADJ, B_SAEADJ, B_SAINDEL, B_SASTRAND, B_SAMAPQ, B_SASAME, B_COUNT,
       B_QUAL4 = B_COUNT, B_SATMP, B_OFFTMP, B_SEQ2, B_ALL };     // B_QUAL4 also stages qual2      // the last four: staging of the transport-compact forms

struct CnvState;
static void cnv_state_free(CnvState *c);
struct gromgpu_chr {
    int tid = 0;
    int64_t P = 0, Ppad = 0, P_cap = 0;          // P_cap: the length the per-position buffers were allocated for (gromgpu_chr_rebind)
    cudaStream_t stream = nullptr;
    char *d_fasta = nullptr;
    int32_t *d_arrays = nullptr;
    DevBuf rb[B_ALL];
    int64_t n_reads = 0, n_cigar = 0, n_slots = 0;
    int32_t last_pos = -1, last_lseq = 0;
    int64_t n_leading = 0;             // reads before W/4+1 (they advance the reference's window index, src/GROM.c:5845)
    uint8_t *d_state = nullptr; PrepRec *d_prep = nullptr; int64_t *d_tile_first = nullptr; size_t cap_state = 0, cap_tiles = 0;
    int *d_max_span = nullptr; unsigned long long *d_counters = nullptr;   // 4 counters + 2 depth sums
    unsigned long long *d_scan_status = nullptr; unsigned int *d_ticket = nullptr; size_t cap_scan = 0;
    grom_snv_cand *d_cand = nullptr; unsigned int cand_cap = 0; unsigned int *d_ncand = nullptr;
    grom_ins_cand *d_ins = nullptr; unsigned int ins_cap = 0; std::vector<grom_ins_cand> h_ins;
    int2 *d_ins_pos = nullptr; unsigned int ins_pos_cap = 0;        // (position, slot mask) whose indel slots reach min_disc (compacted by k_sv_apply)
    grom_del_event *d_del = nullptr; unsigned int del_cap = 0; std::vector<grom_del_event> h_del;
    grom_sv_event *d_svev = nullptr; unsigned int svev_cap = 0; std::vector<grom_sv_event> h_svev;
    int last_lseq_applied = 0;       // l_qseq of the last read plus its hard clips (the reference's cdp_lseq once that read is applied)
    // SV / indel evidence (sv_evidence.cuh)
    int32_t *d_item_cnt = nullptr; size_t cap_item_cnt = 0;
    SvItem *d_items = nullptr; size_t cap_items = 0;
    int2 *d_sv_tiles = nullptr; uint16_t *d_sv_dirty = nullptr; size_t cap_sv_tiles = 0;
    int *d_sv_small = nullptr;                       // [0] reach fwd [1] reach bwd [2] pool used [3] error flag [4] insertion-position list length
    SvOther *d_pool = nullptr; int pool_cap = 0;
    int32_t *d_cl_int = nullptr;                     // cl_w[10] cl_rs[10] cl_re[10] cl_mchr[2] other_len[1] ins_src[3]  (36 x Ppad int32)
    double *d_cl_dist = nullptr;                     // [10][Ppad]
    int64_t n_items = 0;
    unsigned int need_cand = 0, need_ins = 0, need_ins_pos = 0, need_del = 0, need_svev = 0; int64_t need_items = 0; int need_pool = 0;   // what the last pass asked for
    unsigned long long *h_counts = nullptr;          // pinned landing area of the per-pass counters
    unsigned int n_cnt[4] = {0, 0, 0, 0};            // SNV / insertion / deletion / SV-gate counts of the last pass
    std::vector<grom_snv_cand> h_cand;
    struct CnvState *cnv = nullptr;
    cudaEvent_t ev[12];
    cudaEvent_t ev_push[2] = {nullptr, nullptr}; unsigned push_seq = 0;
    bool ran = false;
    gromgpu_stats stats;
    gromgpu_result res;
};

extern "C" const char *gromgpu_last_error(void) { return g_err; }

extern "C" int gromgpu_init(int device, const double *hez_tbl, const double *mq_tbl, const grom_params *p)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail("gromgpu_init: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail("gromgpu_init: device %d out of range (%d devices)", device, ndev);
    if (g_inited && device != g_device) return fail("gromgpu_init: already initialised on device %d; one device per process (tables and parameters live there)", g_device);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail("gromgpu_init: device %d is sm_%d%d; this build targets sm_100a only", device, prop.major, prop.minor);
    if (!g_own_stream) CK(cudaStreamCreateWithFlags(&g_own_stream, cudaStreamNonBlocking));
    if (!g_stream) g_stream = g_own_stream;
    const size_t tb = sizeof(double) * 1001 * 1001;
    if (!d_hez) CK(cudaMalloc(&d_hez, tb));
    if (!d_mq) CK(cudaMalloc(&d_mq, tb));
    CK(cudaMemcpy(d_hez, hez_tbl, tb, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_mq, mq_tbl, tb, cudaMemcpyHostToDevice));
    g_params = *p;
    CK(cudaMemcpyToSymbol(c_prm, p, sizeof(grom_params)));
    CK(cudaFuncSetAttribute(k_pileup, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PileSmem)));
    CK(cudaFuncSetAttribute(cnv::k_sweep_sum, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SSUM_ST * SSUM_FR * 32 * sizeof(double))));
    g_device = device; g_inited = true;
    return 0;
}

extern "C" void gromgpu_shutdown(void)
{
    if (d_hez) cudaFree(d_hez); if (d_mq) cudaFree(d_mq);
    d_hez = d_mq = nullptr;
    if (g_own_stream) cudaStreamDestroy(g_own_stream);
    g_own_stream = g_stream = nullptr; g_inited = false;
}

extern "C" int gromgpu_set_stream(void *s)
{
    if (!g_inited) return fail("gromgpu_set_stream: call gromgpu_init first");
    g_stream = s ? (cudaStream_t)s : g_own_stream;
    return 0;
}

extern "C" int gromgpu_stream_create(void **out)
{
    if (!g_inited) return fail("gromgpu_stream_create: call gromgpu_init first");
    ON_DEV();
    cudaStream_t s = nullptr;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *out = (void *)s;
    return 0;
}
extern "C" void gromgpu_stream_destroy(void *s) { if (s) { if (g_device >= 0) cudaSetDevice(g_device); cudaStreamDestroy((cudaStream_t)s); } }
extern "C" int64_t gromgpu_device_free_bytes(void)
{
    size_t fr = 0, tot = 0;
    if (g_device >= 0) cudaSetDevice(g_device);
    return cudaMemGetInfo(&fr, &tot) == cudaSuccess ? (int64_t)fr : -1;
}
extern "C" int64_t gromgpu_chr_bytes_estimate(int64_t len, int64_t n_reads, int64_t n_base_slots)
{
    const int64_t Ppad = (len + 1023) & ~(int64_t)1023;
    // per position: GA_COUNT count arrays, 36 int + 10 double of cluster state, FASTA, CNV records / depth / scratch (~60 B);
    // per read: the canonical arrays (~80 B + bases), prepared records, state; pools and candidate buffers ~4 GB
    return Ppad * ((int64_t)sizeof(int32_t) * (GA_COUNT + 36) + 80 + 1 + 60) + n_reads * 160 + n_base_slots * 2 + ((int64_t)4 << 30);
}
extern "C" int gromgpu_chr_begin_on(gromgpu_chr **out, int tid, const char *fasta, int64_t len, void *stream);
extern "C" int gromgpu_chr_begin(gromgpu_chr **out, int tid, const char *fasta, int64_t len) { return gromgpu_chr_begin_on(out, tid, fasta, len, (void *)g_stream); }

extern "C" int gromgpu_chr_begin_on(gromgpu_chr **out, int tid, const char *fasta, int64_t len, void *stream)
{
    if (!g_inited) return fail("gromgpu_chr_begin: call gromgpu_init first");
    ON_DEV();
    if (len <= 0 || len > 0x7fffffff) return fail("gromgpu_chr_begin: chromosome length %lld unsupported", (long long)len);
    gromgpu_chr *h = new gromgpu_chr();
    h->tid = tid; h->P = len; h->P_cap = len; h->Ppad = (len + 1023) & ~(int64_t)1023; h->stream = stream ? (cudaStream_t)stream : g_stream;
    memset(&h->stats, 0, sizeof(h->stats)); memset(&h->res, 0, sizeof(h->res));
    for (int i = 0; i < 12; i++) h->ev[i] = nullptr;
    *out = h;
    CK(cudaMalloc(&h->d_fasta, (size_t)h->Ppad));
    CK(cudaMemcpyAsync(h->d_fasta, fasta, (size_t)len, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMalloc(&h->d_arrays, sizeof(int32_t) * (size_t)GA_COUNT * (size_t)h->Ppad));
    CK(cudaMalloc(&h->d_max_span, sizeof(int)));
    CK(cudaMalloc(&h->d_counters, sizeof(unsigned long long) * 8));
    CK(cudaMalloc(&h->d_ticket, sizeof(unsigned int) * 8));
    CK(cudaMalloc(&h->d_ncand, sizeof(unsigned int) * 4));
    for (int i = 0; i < 12; i++) CK(cudaEventCreate(&h->ev[i]));
    for (int i = 0; i < 2; i++) CK(cudaEventCreateWithFlags(&h->ev_push[i], cudaEventDisableTiming));
    CK(cudaMalloc(&h->d_cl_int, sizeof(int32_t) * 36 * (size_t)h->Ppad));
    CK(cudaMalloc(&h->d_cl_dist, sizeof(double) * 10 * (size_t)h->Ppad));
    CK(cudaMemsetAsync(h->d_cl_int, 0, sizeof(int32_t) * 36 * (size_t)h->Ppad, h->stream));
    CK(cudaMemsetAsync(h->d_cl_dist, 0, sizeof(double) * 10 * (size_t)h->Ppad, h->stream));
    CK(cudaMemsetAsync(h->d_arrays, 0, sizeof(int32_t) * (size_t)GA_COUNT * (size_t)h->Ppad, h->stream));
    CK(cudaMalloc(&h->d_sv_small, sizeof(int) * 8));
    CK(cudaMallocHost(&h->h_counts, 256));
    // one slab of 50 side slots per position that ever holds a second cluster of some class: every position for small
    // contigs, at most 2 M slabs (3.2 GB) for chromosome-sized ones; exhaustion is reported as an error, never ignored
    h->pool_cap = (int)std::min<int64_t>(h->Ppad, std::max<int64_t>(h->Ppad / 32, 2 << 20));
    CK(cudaMalloc(&h->d_pool, sizeof(SvOther) * SV_OTHER * (size_t)h->pool_cap));
    {
        const int64_t nt = (h->P + SV_T - 1) / SV_T;
        h->cap_sv_tiles = (size_t)nt;
        CK(cudaMalloc(&h->d_sv_tiles, sizeof(int2) * (size_t)nt));
        CK(cudaMalloc(&h->d_sv_dirty, sizeof(uint16_t) * (size_t)nt));
        CK(cudaMemsetAsync(h->d_sv_dirty, 0, sizeof(uint16_t) * (size_t)nt, h->stream));
    }
    return 0;
}

extern "C" int gromgpu_chr_reset(gromgpu_chr *h, const char *fasta)
{
    if (!h) return fail("gromgpu_chr_reset: null handle");
    ON_DEV();
    if (fasta) CK(cudaMemcpyAsync(h->d_fasta, fasta, (size_t)h->P, cudaMemcpyHostToDevice, h->stream));
    for (int i = 0; i < B_COUNT; i++) h->rb[i].size = 0;
    h->n_reads = h->n_cigar = h->n_slots = 0; h->last_pos = -1; h->last_lseq = 0; h->n_leading = 0; h->ran = false;
    return 0;
}

extern "C" int gromgpu_chr_rebind(gromgpu_chr *h, int tid, const char *fasta, int64_t len)
{
    if (!h || !fasta) return fail("gromgpu_chr_rebind: null argument");
    ON_DEV();
    if (len <= 0) return fail("gromgpu_chr_rebind: chromosome length %lld unsupported", (long long)len);
    if (len > h->P_cap) return 1;                                     // does not fit: the caller frees the handle and begins a new one
    cudaStream_t s = h->stream;
    h->tid = tid; h->P = len; h->Ppad = (len + 1023) & ~(int64_t)1023;
    // the same state gromgpu_chr_begin leaves: characters uploaded, every per-position array of the (new, shorter) layout zero
    CK(cudaMemcpyAsync(h->d_fasta, fasta, (size_t)len, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(h->d_cl_int, 0, sizeof(int32_t) * 36 * (size_t)h->Ppad, s));
    CK(cudaMemsetAsync(h->d_cl_dist, 0, sizeof(double) * 10 * (size_t)h->Ppad, s));
    CK(cudaMemsetAsync(h->d_arrays, 0, sizeof(int32_t) * (size_t)GA_COUNT * (size_t)h->Ppad, s));
    CK(cudaMemsetAsync(h->d_sv_dirty, 0, sizeof(uint16_t) * h->cap_sv_tiles, s));
    for (int i = 0; i < B_COUNT; i++) h->rb[i].size = 0;
    h->n_reads = h->n_cigar = h->n_slots = 0; h->last_pos = -1; h->last_lseq = 0; h->n_leading = 0; h->ran = false; h->last_lseq_applied = 0;
    h->n_items = 0;
    memset(&h->stats, 0, sizeof(h->stats)); memset(&h->res, 0, sizeof(h->res));
    return 0;
}

extern "C" int gromgpu_chr_sync(gromgpu_chr *h)
{
    if (!h) return fail("gromgpu_chr_sync: null handle");
    ON_DEV();
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" void gromgpu_chr_free(gromgpu_chr *h)
{
    if (!h) return;
    if (g_device >= 0) cudaSetDevice(g_device);
    cudaStreamSynchronize(h->stream);
    for (int i = 0; i < B_ALL; i++) h->rb[i].release();
    cudaFree(h->d_fasta); cudaFree(h->d_arrays); cudaFree(h->d_state); cudaFree(h->d_prep); cudaFree(h->d_tile_first);
    cudaFree(h->d_max_span); cudaFree(h->d_counters); cudaFree(h->d_scan_status); cudaFree(h->d_ticket);
    cudaFree(h->d_cand); cudaFree(h->d_ncand); cudaFree(h->d_ins); cudaFree(h->d_ins_pos); cudaFree(h->d_del); cudaFree(h->d_svev);
    for (int i = 0; i < 12; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i < 2; i++) if (h->ev_push[i]) cudaEventDestroy(h->ev_push[i]);
    cudaFree(h->d_item_cnt); cudaFree(h->d_items); cudaFree(h->d_sv_tiles); cudaFree(h->d_sv_dirty); cudaFree(h->d_sv_small); cudaFree(h->d_pool);
    cudaFree(h->d_cl_int); cudaFree(h->d_cl_dist);
    if (h->h_counts) cudaFreeHost(h->h_counts);
    cnv_state_free(h->cnv);
    delete h;
}

extern "C" int gromgpu_push_reads(gromgpu_chr *h, const grom_read_batch *b)
{
    if (!h || !b) return fail("gromgpu_push_reads: null argument");
    ON_DEV();
    const int64_t n = b->n_reads;
    if (n == 0) return 0;
    if (b->pos[0] < h->last_pos) return fail("gromgpu_push_reads: reads are not in coordinate order");
    if ((h->n_cigar + b->n_cigar_total) > 0xffffffffLL) return fail("gromgpu_push_reads: more than 2^32 CIGAR operations on one chromosome");
    const int lay = b->layout_flags;
    const bool lay_off = lay & GROM_LAYOUT_CANONICAL_OFFSETS, lay_q2 = (lay & GROM_LAYOUT_QUAL2) && b->qual2, lay_q4 = !lay_q2 && (lay & GROM_LAYOUT_QUAL4) && b->qual4, lay_sa = (lay & GROM_LAYOUT_SPARSE_SA) != 0,
               lay_s2 = (lay & GROM_LAYOUT_SEQ2) && b->seq2;
    if (!lay_s2 && !b->seq4) return fail("gromgpu_push_reads: bases missing");
    if (lay_s2 && (b->n_seq_exc < 0 || (b->n_seq_exc && (!b->seq_exc_slot || !b->seq_exc_code)))) return fail("gromgpu_push_reads: bad base exception list");
    if (!lay_sa && !b->sa_pos) return fail("gromgpu_push_reads: SA arrays missing");
    if (!lay_off && (!b->cigar_off || !b->base_off)) return fail("gromgpu_push_reads: offset arrays missing (and GROM_LAYOUT_CANONICAL_OFFSETS not set)");
    if (!lay_q4 && !lay_q2 && !b->qual) return fail("gromgpu_push_reads: qualities missing");
    if (lay_sa && (b->n_sa < 0 || b->n_sa > n || (b->n_sa && !b->sa_index))) return fail("gromgpu_push_reads: bad sparse SA list");
    // src == nullptr: the array is rebuilt on the device from a transport-compact form (below)
    struct { int id; const void *src; size_t elt; int64_t cnt, have; } f[B_COUNT] = {
        { B_POS, b->pos, 4, n, h->n_reads }, { B_MPOS, b->mpos, 4, n, h->n_reads }, { B_TLEN, b->tlen, 4, n, h->n_reads },
        { B_MTID, b->mtid, 4, n, h->n_reads }, { B_LQSEQ, b->l_qseq, 4, n, h->n_reads }, { B_FLAG, b->flag, 2, n, h->n_reads },
        { B_NCIGAR, b->n_cigar, 2, n, h->n_reads }, { B_MAPQ, b->mapq, 1, n, h->n_reads }, { B_QLEN, b->qname_len, 1, n, h->n_reads },
        { B_HASH, b->qname_hash, 8, n, h->n_reads }, { B_CIGOFF, lay_off ? nullptr : b->cigar_off, 8, n, h->n_reads }, { B_BASEOFF, lay_off ? nullptr : b->base_off, 8, n, h->n_reads },
        { B_CIGAR, b->cigar, 4, b->n_cigar_total, h->n_cigar }, { B_SEQ4, lay_s2 ? nullptr : b->seq4, 1, (b->n_base_slots + 1) / 2, h->n_slots / 2 },
        { B_QUAL, (lay_q4 || lay_q2) ? nullptr : b->qual, 1, b->n_base_slots, h->n_slots },
        { B_SAPOS, lay_sa ? nullptr : b->sa_pos, 4, n, h->n_reads }, { B_SASADJ, lay_sa ? nullptr : b->sa_start_adj, 4, n, h->n_reads }, { B_SAEADJ, lay_sa ? nullptr : b->sa_end_adj, 4, n, h->n_reads },
        { B_SAINDEL, lay_sa ? nullptr : b->sa_end_adj_indel, 4, n, h->n_reads }, { B_SASTRAND, lay_sa ? nullptr : b->sa_strand, 1, n, h->n_reads }, { B_SAMAPQ, lay_sa ? nullptr : b->sa_mapq, 2, n, h->n_reads },
        { B_SASAME, lay_sa ? nullptr : b->sa_same_chr, 1, n, h->n_reads } };
    // in pieces, at most a few in flight: the copy engine serves requests in submission order, so a multi-GB transfer queued
    // at once would stall every small copy of a second contig that is computing on another stream
    auto upload = [&](void *dst, const void *src, size_t total) -> int {
        static const size_t piece = []() { const char *e = getenv("GROMGPU_PUSH_PIECE_MB"); const long mb = e ? atol(e) : 4; return (size_t)(mb > 0 ? mb : 4) << 20; }();
        for (size_t o = 0; o < total; o += piece) {
            CK(cudaMemcpyAsync((char *)dst + o, (const char *)src + o, std::min(piece, total - o), cudaMemcpyHostToDevice, h->stream));
            if (total > piece) {
                CK(cudaEventRecord(h->ev_push[h->push_seq & 1], h->stream));
                h->push_seq++;
                if (h->push_seq >= 2) CK(cudaEventSynchronize(h->ev_push[h->push_seq & 1]));       // the piece before the last one has landed
            }
        }
        return 0;
    };
    for (int k = 0; k < B_COUNT; k++) {
        DevBuf &d = h->rb[f[k].id];
        const size_t need = (size_t)(f[k].have + f[k].cnt) * f[k].elt + 64;
        if (d.ensure(need, h->stream)) return -1;
        if (f[k].cnt && f[k].src && upload((char *)d.p + (size_t)f[k].have * f[k].elt, f[k].src, (size_t)f[k].cnt * f[k].elt)) return -1;
        d.size = (size_t)(f[k].have + f[k].cnt) * f[k].elt;
    }
    if (lay_off) {
        const int64_t nblk = (n + OFF_BLOCK - 1) / OFF_BLOCK;
        DevBuf &t = h->rb[B_OFFTMP];
        if (t.ensure(sizeof(uint64_t) * 2 * (size_t)nblk + 64, h->stream)) return -1;
        const uint16_t *nc = (const uint16_t *)h->rb[B_NCIGAR].p + h->n_reads; const int32_t *lq = (const int32_t *)h->rb[B_LQSEQ].p + h->n_reads;
        k_off_sums<<<(unsigned)nblk, 256, 0, h->stream>>>(nc, lq, n, (uint64_t *)t.p);
        k_off_scan<<<1, 1024, 0, h->stream>>>((uint64_t *)t.p, nblk);
        k_off_write<<<(unsigned)nblk, 256, 0, h->stream>>>(nc, lq, n, (const uint64_t *)t.p, (uint64_t)h->n_cigar, (uint64_t)h->n_slots,
                                                          (uint64_t *)h->rb[B_CIGOFF].p + h->n_reads, (uint64_t *)h->rb[B_BASEOFF].p + h->n_reads);
        CK(cudaGetLastError());
    } else if (h->n_cigar || h->n_slots) {
        k_fix_offsets<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>((uint64_t *)h->rb[B_CIGOFF].p + h->n_reads, (uint64_t *)h->rb[B_BASEOFF].p + h->n_reads,
                                                                           n, (uint64_t)h->n_cigar, (uint64_t)h->n_slots);
        CK(cudaGetLastError());
    }
    if (lay_q4 && b->n_base_slots) {
        DevBuf &t = h->rb[B_QUAL4];
        const size_t nb = (size_t)(b->n_base_slots + 1) / 2;
        if (t.ensure(nb + 64, h->stream)) return -1;
        if (upload(t.p, b->qual4, nb)) return -1;
        QualLut lut; memcpy(lut.v, b->qual_lut, 16);
        k_expand_qual<<<(unsigned)((b->n_base_slots + 16 * 256 - 1) / (16 * 256)), 256, 0, h->stream>>>((const uint8_t *)t.p, b->n_base_slots, lut, (uint8_t *)h->rb[B_QUAL].p + h->n_slots);
        CK(cudaGetLastError());
    }
    if (lay_s2 && b->n_base_slots) {
        DevBuf &t = h->rb[B_SEQ2];
        const size_t nb = ((size_t)(b->n_base_slots + 3) / 4 + 15) & ~(size_t)15, ne = (size_t)b->n_seq_exc;
        if (t.ensure(nb + ne * 9 + 64, h->stream)) return -1;
        if (upload(t.p, b->seq2, (size_t)(b->n_base_slots + 3) / 4)) return -1;
        uint8_t *d_seq = (uint8_t *)h->rb[B_SEQ4].p + h->n_slots / 2;
        k_expand_seq<<<(unsigned)((b->n_base_slots + 16 * 256 - 1) / (16 * 256)), 256, 0, h->stream>>>((const uint8_t *)t.p, b->n_base_slots, d_seq);
        if (ne) {
            uint64_t *d_slot = (uint64_t *)((char *)t.p + nb); uint8_t *d_code = (uint8_t *)(d_slot + ne);
            if (upload(d_slot, b->seq_exc_slot, ne * 8) || upload(d_code, b->seq_exc_code, ne)) return -1;
            k_seq_exceptions<<<(unsigned)((ne + 255) / 256), 256, 0, h->stream>>>(d_slot, d_code, (int64_t)ne, b->n_base_slots, d_seq);
        }
    }
    if (lay_q2 && b->n_base_slots) {
        DevBuf &t = h->rb[B_QUAL4];
        const size_t nb = (size_t)(b->n_base_slots + 3) / 4;
        if (t.ensure(nb + 64, h->stream)) return -1;
        if (upload(t.p, b->qual2, nb)) return -1;
        QualLut lut; memcpy(lut.v, b->qual_lut, 16);
        k_expand_qual2<<<(unsigned)((b->n_base_slots + 16 * 256 - 1) / (16 * 256)), 256, 0, h->stream>>>((const uint8_t *)t.p, b->n_base_slots, lut, (uint8_t *)h->rb[B_QUAL].p + h->n_slots);
        CK(cudaGetLastError());
    }
    if ((lay_s2 || lay_q2) && b->n_base_slots) {
        // padding slots of every read back to 0 like in the canonical arrays (offsets are absolute by now)
        k_seq_zero_pad<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>((const int32_t *)h->rb[B_LQSEQ].p + h->n_reads, (const uint64_t *)h->rb[B_BASEOFF].p + h->n_reads, n,
                                                                            lay_s2 ? (uint8_t *)h->rb[B_SEQ4].p : nullptr, lay_q2 ? (uint8_t *)h->rb[B_QUAL].p : nullptr);
        CK(cudaGetLastError());
    }
    if (lay_sa) {
        int32_t *d_pos = (int32_t *)h->rb[B_SAPOS].p + h->n_reads;
        CK(cudaMemsetAsync(d_pos, 0xff, sizeof(int32_t) * (size_t)n, h->stream));              // sa_pos = -1: no entry
        for (int id : {B_SASADJ, B_SAEADJ, B_SAINDEL}) CK(cudaMemsetAsync((int32_t *)h->rb[id].p + h->n_reads, 0, sizeof(int32_t) * (size_t)n, h->stream));
        CK(cudaMemsetAsync((uint8_t *)h->rb[B_SASTRAND].p + h->n_reads, 0, (size_t)n, h->stream));
        CK(cudaMemsetAsync((int16_t *)h->rb[B_SAMAPQ].p + h->n_reads, 0xff, sizeof(int16_t) * (size_t)n, h->stream));     // -1 like sa_pos
        CK(cudaMemsetAsync((uint8_t *)h->rb[B_SASAME].p + h->n_reads, 0, (size_t)n, h->stream));
        const int64_t m = b->n_sa;
        if (m) {
            DevBuf &t = h->rb[B_SATMP];
            const size_t m4 = ((size_t)m * 4 + 15) & ~(size_t)15, m2 = ((size_t)m * 2 + 15) & ~(size_t)15, m1 = ((size_t)m + 15) & ~(size_t)15;
            if (t.ensure(5 * m4 + m2 + 2 * m1 + 64, h->stream)) return -1;
            char *q = (char *)t.p;
            SaSparse S;
            S.idx = (const int32_t *)q; if (upload(q, b->sa_index, (size_t)m * 4)) return -1; q += m4;
            S.pos = (const int32_t *)q; if (upload(q, b->sas_pos, (size_t)m * 4)) return -1; q += m4;
            S.sadj = (const int32_t *)q; if (upload(q, b->sas_start_adj, (size_t)m * 4)) return -1; q += m4;
            S.eadj = (const int32_t *)q; if (upload(q, b->sas_end_adj, (size_t)m * 4)) return -1; q += m4;
            S.indel = (const int32_t *)q; if (upload(q, b->sas_end_adj_indel, (size_t)m * 4)) return -1; q += m4;
            S.mapq = (const int16_t *)q; if (upload(q, b->sas_mapq, (size_t)m * 2)) return -1; q += m2;
            S.strand = (const uint8_t *)q; if (upload(q, b->sas_strand, (size_t)m)) return -1; q += m1;
            S.same = (const uint8_t *)q; if (upload(q, b->sas_same_chr, (size_t)m)) return -1;
            k_scatter_sa<<<(unsigned)((m + 255) / 256), 256, 0, h->stream>>>(S, m, n, d_pos, (int32_t *)h->rb[B_SASADJ].p + h->n_reads, (int32_t *)h->rb[B_SAEADJ].p + h->n_reads,
                                                                             (int32_t *)h->rb[B_SAINDEL].p + h->n_reads, (uint8_t *)h->rb[B_SASTRAND].p + h->n_reads,
                                                                             (int16_t *)h->rb[B_SAMAPQ].p + h->n_reads, (uint8_t *)h->rb[B_SASAME].p + h->n_reads);
            CK(cudaGetLastError());
        }
    }
    // host-side bookkeeping the scan range needs (src/GROM.c:6406, 11075-11083)
    const int first_pos = grom_first_pos(&g_params);
    for (int64_t i = 0; i < n && b->pos[i] < first_pos; i++) h->n_leading++;
    h->last_pos = b->pos[n - 1]; h->last_lseq = b->l_qseq[n - 1];
    {
        // hard clips of the last read extend its length once it is applied (src/GROM.c:6997-7000; first max_cigar_ops operations)
        int hsum = 0;
        const uint64_t c0 = lay_off ? (uint64_t)(b->n_cigar_total - b->n_cigar[n - 1]) : b->cigar_off[n - 1];
        const int nc = std::min<int>(b->n_cigar[n - 1], g_params.max_cigar_ops);
        for (int k = 0; k < nc; k++) if ((b->cigar[c0 + k] & 15) == 5) hsum += (int)(b->cigar[c0 + k] >> 4);
        h->last_lseq_applied = h->last_lseq + hsum;
    }
    h->n_reads += n; h->n_cigar += b->n_cigar_total;
    h->n_slots += (b->n_base_slots + GROM_BASE_ALIGN - 1) / GROM_BASE_ALIGN * GROM_BASE_ALIGN;
    h->ran = false;
    return 0;
}

static DevReads dev_reads(const gromgpu_chr *h)
{
    DevReads R;
    R.n = h->n_reads;
    R.pos = (const int32_t *)h->rb[B_POS].p; R.mpos = (const int32_t *)h->rb[B_MPOS].p; R.tlen = (const int32_t *)h->rb[B_TLEN].p;
    R.mtid = (const int32_t *)h->rb[B_MTID].p; R.l_qseq = (const int32_t *)h->rb[B_LQSEQ].p; R.flag = (const uint16_t *)h->rb[B_FLAG].p;
    R.n_cigar = (const uint16_t *)h->rb[B_NCIGAR].p; R.mapq = (const uint8_t *)h->rb[B_MAPQ].p; R.qname_len = (const uint8_t *)h->rb[B_QLEN].p;
    R.qname_hash = (const uint64_t *)h->rb[B_HASH].p; R.cigar_off = (const uint64_t *)h->rb[B_CIGOFF].p; R.base_off = (const uint64_t *)h->rb[B_BASEOFF].p;
    R.cigar = (const uint32_t *)h->rb[B_CIGAR].p; R.seq4 = (const uint8_t *)h->rb[B_SEQ4].p; R.qual = (const uint8_t *)h->rb[B_QUAL].p;
    return R;
}

// One pass over everything pushed so far.  Result buffers (candidates, gate events, evidence items, listed positions) keep their
// size from the previous run; the device counts what it would have written, and a pass that outgrew a buffer is repeated by
// gromgpu_chr_run with larger ones (the pass resets all of its state first, so it is idempotent).
static int chr_run_once(gromgpu_chr *h, bool *again)
{
    cudaStream_t s = h->stream;
    const int64_t n = h->n_reads, P = h->P, Ppad = h->Ppad;
    const int64_t n_tiles = (P + TILE - 1) / TILE;
    const int64_t n_scan_tiles = (Ppad + SCAN_TILE - 1) / SCAN_TILE;
    // scratch sized to the current input
    if ((size_t)n > h->cap_state) {
        cudaFree(h->d_state); cudaFree(h->d_prep);
        h->cap_state = (size_t)n + (size_t)n / 8 + 1024;
        CK(cudaMalloc(&h->d_state, h->cap_state));
        CK(cudaMalloc(&h->d_prep, h->cap_state * sizeof(PrepRec)));
    }
    if ((size_t)n_tiles > h->cap_tiles) { cudaFree(h->d_tile_first); h->cap_tiles = (size_t)n_tiles; CK(cudaMalloc(&h->d_tile_first, sizeof(int64_t) * h->cap_tiles)); }
    // first sizes scale with the contig; a run that needs more is repeated with what it asked for (gromgpu_chr_run)
    const bool tiny = getenv("GROMGPU_TEST_SMALL_BUFFERS") != nullptr;          // tests: first sizes of a few entries, so the repeat path runs
    auto want = [&](unsigned int &cap, unsigned int &need, int64_t first) { if (!cap) cap = tiny ? 16u : (unsigned int)std::min<int64_t>(first, 1 << 30); if (need > cap) cap = (unsigned int)std::min<uint64_t>((uint64_t)need + need / 4 + 1024, 0x7fffffffu); need = 0; };
    { const unsigned int c0 = h->cand_cap; want(h->cand_cap, h->need_cand, std::max<int64_t>(1 << 20, Ppad / 64));
      if (h->cand_cap != c0) { cudaFree(h->d_cand); h->d_cand = nullptr; CK(cudaMalloc(&h->d_cand, sizeof(grom_snv_cand) * (size_t)h->cand_cap)); } }
    { const unsigned int c0 = h->ins_cap; want(h->ins_cap, h->need_ins, std::max<int64_t>(1 << 18, Ppad / 256));
      if (h->ins_cap != c0) { cudaFree(h->d_ins); h->d_ins = nullptr; CK(cudaMalloc(&h->d_ins, sizeof(grom_ins_cand) * (size_t)h->ins_cap)); } }
    { const unsigned int c0 = h->ins_pos_cap; want(h->ins_pos_cap, h->need_ins_pos, std::max<int64_t>(1 << 20, Ppad / 32));
      if (h->ins_pos_cap != c0) { cudaFree(h->d_ins_pos); h->d_ins_pos = nullptr; CK(cudaMalloc(&h->d_ins_pos, sizeof(int2) * (size_t)h->ins_pos_cap)); } }
    { const unsigned int c0 = h->del_cap; want(h->del_cap, h->need_del, std::max<int64_t>(1 << 19, Ppad / 128));
      if (h->del_cap != c0) { cudaFree(h->d_del); h->d_del = nullptr; CK(cudaMalloc(&h->d_del, sizeof(grom_del_event) * (size_t)h->del_cap)); } }
    { const unsigned int c0 = h->svev_cap; want(h->svev_cap, h->need_svev, std::max<int64_t>(1 << 16, Ppad / 16));
      if (h->svev_cap != c0) { cudaFree(h->d_svev); h->d_svev = nullptr; CK(cudaMalloc(&h->d_svev, sizeof(grom_sv_event) * (size_t)h->svev_cap)); } }
    {   // evidence items: a few per discordant / clipped / indel-carrying read
        const size_t first = tiny ? 64 : (size_t)n / 4 + 65536, need = (size_t)h->need_items + (size_t)h->need_items / 4 + 1024;
        const size_t capw = h->need_items > (int64_t)h->cap_items ? need : (h->cap_items ? h->cap_items : first);
        if (capw != h->cap_items) { cudaFree(h->d_items); h->d_items = nullptr; h->cap_items = capw; CK(cudaMalloc(&h->d_items, sizeof(SvItem) * h->cap_items)); }
        h->need_items = 0;
    }
    if (h->need_pool > h->pool_cap) {
        cudaFree(h->d_pool); h->d_pool = nullptr;
        h->pool_cap = (int)std::min<int64_t>(Ppad, (int64_t)h->need_pool + h->need_pool / 4 + 1024);
        CK(cudaMalloc(&h->d_pool, sizeof(SvOther) * SV_OTHER * (size_t)h->pool_cap));
    }
    h->need_pool = 0;
    const int64_t n_cnt_pad = (n + 1023) & ~(int64_t)1023;
    const int64_t n_cnt_tiles = (n_cnt_pad + SCAN_TILE - 1) / SCAN_TILE;
    if ((size_t)n_cnt_pad > h->cap_item_cnt) { cudaFree(h->d_item_cnt); h->cap_item_cnt = (size_t)n_cnt_pad + (size_t)n_cnt_pad / 8; CK(cudaMalloc(&h->d_item_cnt, sizeof(int32_t) * h->cap_item_cnt)); }
    const int64_t n_sv_tiles = (P + SV_T - 1) / SV_T;
    {   // scan status: 5 position arrays + 1 item-count array
        const size_t need = (size_t)n_scan_tiles * 5 + (size_t)n_cnt_tiles + 8;
        if (need > h->cap_scan) { cudaFree(h->d_scan_status); h->cap_scan = need; CK(cudaMalloc(&h->d_scan_status, sizeof(unsigned long long) * h->cap_scan)); }
    }

    const int first_pos = grom_first_pos(&g_params);
    int scan_first = -1, scan_last = -1;
    if (h->n_leading < n) {
        scan_first = first_pos;
        int64_t sl = (int64_t)h->last_pos - (int64_t)g_params.overlap_mult * g_params.insert_max;
        if (sl < scan_first) sl = scan_first;
        if (sl >= P) sl = P - 1;
        scan_last = (int)sl;
    }
    int64_t depth_bound = 0;
    if (scan_first >= 0) {
        const int W = grom_window_len(&g_params);
        const int64_t idx = W / 4 + ((h->n_leading + 2 + ((int64_t)scan_last - first_pos)) % (W / 2));
        depth_bound = (int64_t)scan_last + 1 - idx;
    }
    int launches = 0;
    DevReads R = dev_reads(h);
    CK(cudaEventRecord(h->ev[0], s));
    // zero only the arrays that are scatter targets (rd .. indel_d_r_rd); the pileup and depth arrays are fully overwritten
    CK(cudaMemsetAsync(h->d_arrays + (int64_t)GA_RD * Ppad, 0, sizeof(int32_t) * (size_t)(GA_INDEL_I - GA_RD) * (size_t)Ppad, s));   // rd, clips, conc, ins, munmapped
    CK(cudaMemsetAsync(h->d_arrays + (int64_t)GA_INDEL_D_F_RD * Ppad, 0, sizeof(int32_t) * (size_t)Ppad, s));
    CK(cudaMemsetAsync(h->d_arrays + (int64_t)GA_INDEL_D_R_RD * Ppad, 0, sizeof(int32_t) * (size_t)Ppad, s));
    CK(cudaMemsetAsync(h->d_sv_small, 0, sizeof(int) * 8, s));
    CK(cudaMemsetAsync(h->d_item_cnt, 0, sizeof(int32_t) * (size_t)n_cnt_pad, s));
    CK(cudaMemsetAsync(h->d_max_span, 0, sizeof(int), s));
    CK(cudaMemsetAsync(h->d_counters, 0, sizeof(unsigned long long) * 8, s));
    CK(cudaMemsetAsync(h->d_ticket, 0, sizeof(unsigned int) * 8, s));
    CK(cudaMemsetAsync(h->d_ncand, 0, sizeof(unsigned int) * 4, s));
    CK(cudaMemsetAsync(h->d_scan_status, 0, sizeof(unsigned long long) * ((size_t)n_scan_tiles * 5 + (size_t)n_cnt_tiles), s));
    CK(cudaEventRecord(h->ev[1], s));
    {
        const int M = g_params.insert_mean;
        const size_t smem = sizeof(int) * 2 * (size_t)(GC_TILE + 2 * M + 1);
        if (smem > 200 * 1024) return fail("gromgpu_chr_run: insert_mean %d too large for the GC pre-pass tile", M);
        static size_t gc_smem_set = 0;                 // the attribute sticks to the function: set it only when the window grows
        if (smem > gc_smem_set) { CK(cudaFuncSetAttribute(k_gc_prepass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); gc_smem_set = smem; }
        k_gc_prepass<<<(unsigned)((P + GC_TILE - 1) / GC_TILE), GC_THREADS, smem, s>>>(h->d_fasta, P, M, h->d_arrays + (int64_t)GA_GC * Ppad,
                                                                                       h->d_arrays + (int64_t)GA_ACGT * Ppad); launches++;
    }
    CK(cudaEventRecord(h->ev[8], s));
    const unsigned rb = (unsigned)((n + 255) / 256);
    if (n) { k_read_state<<<rb, 256, 0, s>>>(R, h->tid, first_pos, h->d_state, h->d_counters + 6); launches++; }
    CK(cudaEventRecord(h->ev[2], s));
    if (n) { k_read_prep<<<rb, 256, 0, s>>>(R, h->tid, P, Ppad, h->d_state, h->d_prep, h->d_arrays, h->d_max_span, h->d_counters); launches++; }
    CK(cudaEventRecord(h->ev[3], s));
    k_tile_index<<<(unsigned)((n_tiles + 255) / 256), 256, 0, s>>>(R.pos, h->d_prep, n, h->d_max_span, n_tiles, h->d_tile_first); launches++;
    CK(cudaEventRecord(h->ev[4], s));
    // ---- SV / indel evidence: items in BAM order, per-tile fold, then the five range-add prefix scans
    SvDev SD;
    SD.cl_w = h->d_cl_int; SD.cl_rs = h->d_cl_int + 10 * Ppad; SD.cl_re = h->d_cl_int + 20 * Ppad; SD.cl_mchr = h->d_cl_int + 30 * Ppad;
    SD.other_len = h->d_cl_int + 32 * Ppad; SD.ins_src = h->d_cl_int + 33 * Ppad; SD.cl_dist = h->d_cl_dist; SD.pool = h->d_pool; SD.pool_cap = h->pool_cap;
    SD.pool_used = h->d_sv_small + 2; SD.err = h->d_sv_small + 3;
    SD.ins_pos = h->d_ins_pos; SD.ins_pos_cap = (int)h->ins_pos_cap; SD.n_ins_pos = h->d_sv_small + 4;
    h->n_items = 0;
    k_sv_clear<<<(unsigned)n_sv_tiles, SV_T, 0, s>>>(h->d_sv_dirty, P, Ppad, h->d_arrays, SD); launches++;
    if (n) {
        SvReadArrays SA;
        SA.sa_pos = (const int32_t *)h->rb[B_SAPOS].p; SA.sa_start_adj = (const int32_t *)h->rb[B_SASADJ].p; SA.sa_end_adj = (const int32_t *)h->rb[B_SAEADJ].p;
        SA.sa_end_adj_indel = (const int32_t *)h->rb[B_SAINDEL].p; SA.sa_strand = (const uint8_t *)h->rb[B_SASTRAND].p;
        SA.sa_mapq = (const int16_t *)h->rb[B_SAMAPQ].p; SA.sa_same = (const uint8_t *)h->rb[B_SASAME].p;
        k_sv_count<<<rb, 256, 0, s>>>(R, SA, h->tid, h->n_leading, h->d_state, P, h->d_item_cnt); launches++;
        ScanList sl0; memset(&sl0, 0, sizeof(sl0));
        k_scan_inplace<<<dim3((unsigned)n_cnt_tiles, 1), SCAN_THREADS, 0, s>>>(h->d_item_cnt, sl0, n_cnt_pad, h->d_scan_status + (size_t)n_scan_tiles * 5, h->d_ticket + 5); launches++;
        // no round trip for the item total: the emitter stops at the buffer's end and the total is read with the other counters
        k_sv_emit<<<rb, 256, 0, s>>>(R, SA, h->tid, h->n_leading, h->d_state, h->d_item_cnt, h->d_items, (int64_t)h->cap_items, h->d_arrays, P, Ppad, h->d_sv_small); launches++;
        k_sv_tiles<<<(unsigned)((n_sv_tiles + 255) / 256), 256, 0, s>>>(R.pos, n, h->d_item_cnt, h->d_sv_small, n_sv_tiles, (int64_t)h->cap_items, h->d_sv_tiles); launches++;
        k_sv_apply<<<(unsigned)n_sv_tiles, SV_T, sizeof(SvTileState), s>>>(h->d_items, h->d_sv_tiles, P, Ppad, h->d_arrays, SD, h->d_sv_dirty); launches++;
    }
    CK(cudaEventRecord(h->ev[9], s));
    {
        ScanList sl; memset(&sl, 0, sizeof(sl));
        sl.off[0] = (int64_t)GA_RD * Ppad; sl.off[1] = (int64_t)GA_CONC * Ppad; sl.off[2] = (int64_t)GA_INS * Ppad;
        sl.off[3] = (int64_t)GA_MUNMAPPED_F * Ppad; sl.off[4] = (int64_t)GA_MUNMAPPED_R * Ppad;
        k_scan_inplace<<<dim3((unsigned)n_scan_tiles, 5), SCAN_THREADS, 0, s>>>(h->d_arrays, sl, Ppad, h->d_scan_status, h->d_ticket); launches++;
    }
    CK(cudaEventRecord(h->ev[5], s));
    SnvScanArgs sca;
    sca.scan_first = scan_first; sca.scan_last = scan_last; sca.depth_bound = depth_bound; sca.hez = d_hez; sca.mqt = d_mq;
    sca.cand = h->d_cand; sca.cand_cap = h->cand_cap; sca.n_cand = h->d_ncand; sca.depth_sum = h->d_counters + 4;
    sca.ins = h->d_ins; sca.ins_cap = h->ins_cap; sca.n_ins = h->d_ncand + 1; sca.del_ev = h->d_del; sca.del_cap = h->del_cap; sca.n_del = h->d_ncand + 2; sca.other_len = SD.other_len; sca.ins_src = SD.ins_src;
    k_pileup<<<(unsigned)((n_tiles + SUBTILES - 1) / SUBTILES), PILE_THREADS, sizeof(PileSmem), s>>>(R, h->d_prep, h->d_tile_first, h->d_max_span, n_tiles,
                                                                                                   h->d_fasta, P, Ppad, h->d_arrays, sca); launches++;
    CK(cudaEventRecord(h->ev[6], s));
    k_indel_gate<<<(h->ins_pos_cap + 127) / 128, 128, 0, s>>>(h->d_ins_pos, (const unsigned int *)(h->d_sv_small + 4), h->ins_pos_cap, R, P, Ppad, h->d_arrays, sca); launches++;
    {
        SvGateArgs G;
        G.cl_w = SD.cl_w; G.cl_rs = SD.cl_rs; G.cl_re = SD.cl_re; G.cl_mchr = SD.cl_mchr; G.other_len = SD.other_len; G.cl_dist = SD.cl_dist;
        G.ev = h->d_svev; G.cap = h->svev_cap; G.n_ev = h->d_ncand + 3; G.i0 = h->n_leading; G.n_reads = h->n_reads;
        G.last_lseq = h->last_lseq; G.last_lseq_applied = h->last_lseq_applied; G.state = h->d_state;
        k_sv_gate<<<(h->ins_pos_cap + 127) / 128, 128, 0, s>>>(h->d_ins_pos, (const unsigned int *)(h->d_sv_small + 4), h->ins_pos_cap, R, P, Ppad, h->d_arrays, sca, G); launches++;
        if (scan_last >= scan_first) { k_ins_sv_gate<<<(unsigned)(((int64_t)scan_last - scan_first + 256) / 256), 256, 0, s>>>(Ppad, h->d_arrays, sca, G); launches++; }
    }
    CK(cudaEventRecord(h->ev[7], s));
    CK(cudaGetLastError());
    // every counter of the pass in one pinned landing area
    unsigned long long *cnt = h->h_counts;                         // [0..7] counters
    unsigned int *ncnt = (unsigned int *)(cnt + 8);                // [0..3] candidates / events
    int *small = (int *)(ncnt + 4);                                // [0..7] reach, pool, error, listed positions
    int32_t *items_total = small + 8;
    CK(cudaMemcpyAsync(cnt, h->d_counters, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ncnt, h->d_ncand, sizeof(unsigned int) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(small, h->d_sv_small, sizeof(int) * 8, cudaMemcpyDeviceToHost, s));
    *items_total = 0;
    if (n) CK(cudaMemcpyAsync(items_total, h->d_item_cnt + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (cnt[6]) return fail("gromgpu_chr_run: %llu reads are out of coordinate order", cnt[6]);
    if (cnt[7]) return fail("gromgpu_chr_run: %llu reads are longer than 65535 bases (unsupported)", cnt[7]);
    if (*items_total < 0) return fail("gromgpu_chr_run: more than 2^31 evidence items on one chromosome");
    h->n_items = *items_total;
    if ((size_t)*items_total > h->cap_items) { h->need_items = *items_total; *again = true; }
    if (ncnt[0] > h->cand_cap) { h->need_cand = ncnt[0]; *again = true; }
    if (ncnt[1] > h->ins_cap) { h->need_ins = ncnt[1]; *again = true; }
    if (ncnt[2] > h->del_cap) { h->need_del = ncnt[2]; *again = true; }
    if (ncnt[3] > h->svev_cap) { h->need_svev = ncnt[3]; *again = true; }
    if ((unsigned int)small[4] > h->ins_pos_cap) { h->need_ins_pos = (unsigned int)small[4]; *again = true; }
    if (small[2] > h->pool_cap) { if (h->pool_cap >= Ppad) return fail("gromgpu_chr_run: other-slot pool of %d position slabs exhausted", h->pool_cap); h->need_pool = small[2]; *again = true; }
    if (*again) return 0;
    float ms;
    gromgpu_stats &st = h->stats;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[7]); st.ms_total = ms;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); st.ms_clear = ms;
    cudaEventElapsedTime(&ms, h->ev[1], h->ev[8]); st.ms_gc = ms;
    cudaEventElapsedTime(&ms, h->ev[8], h->ev[2]); st.ms_dup = ms;
    cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]); st.ms_prep = ms;
    cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]); st.ms_index = ms;
    cudaEventElapsedTime(&ms, h->ev[4], h->ev[9]); st.ms_sv = ms;
    cudaEventElapsedTime(&ms, h->ev[9], h->ev[5]); st.ms_rdscan = ms;
    cudaEventElapsedTime(&ms, h->ev[5], h->ev[6]); st.ms_pileup = ms;
    st.ms_snvscan = 0.f;      // the SNV gate runs in the pileup kernel's epilogue
    st.launches = launches;
    if (small[3]) return fail("gromgpu_chr_run: device error flag %d after a pass whose buffers were large enough", small[3]);
    st.n_sv_items = h->n_items; st.n_other_slabs = small[2];
    h->n_cnt[0] = ncnt[0]; h->n_cnt[1] = ncnt[1]; h->n_cnt[2] = ncnt[2]; h->n_cnt[3] = ncnt[3];
    st.n_reads = n; st.n_applied = (int64_t)cnt[0]; st.n_dups = (int64_t)cnt[1]; st.aligned_bases = (int64_t)cnt[2]; st.bytes_reads = (int64_t)cnt[3];
    h->res.scan_first = scan_first; h->res.scan_last = scan_last;
    h->res.snv_ave_rd = (double)(long)cnt[4] / (double)(long)cnt[5];
    h->ran = true;
    return 0;
}

extern "C" int gromgpu_chr_run(gromgpu_chr *h)
{
    if (!h) return fail("gromgpu_chr_run: null handle");
    ON_DEV();
    for (int attempt = 0; attempt < 6; attempt++) {
        bool again = false;
        if (chr_run_once(h, &again)) return -1;
        if (!again) return 0;
    }
    return fail("gromgpu_chr_run: result buffers still too small after five repeats");
}

extern "C" int gromgpu_chr_result(gromgpu_chr *h, gromgpu_result *out)
{
    if (!h || !h->ran) return fail("gromgpu_chr_result: call gromgpu_chr_run first");
    ON_DEV();
    const unsigned int nc = h->n_cnt[0];
    if (nc > h->cand_cap) return fail("gromgpu_chr_result: %u SNV candidates exceed the buffer of %u", nc, h->cand_cap);
    h->h_cand.resize(nc);
    if (nc) CK(cudaMemcpy(h->h_cand.data(), h->d_cand, sizeof(grom_snv_cand) * (size_t)nc, cudaMemcpyDeviceToHost));
    std::sort(h->h_cand.begin(), h->h_cand.end(), [](const grom_snv_cand &a, const grom_snv_cand &b) { return a.pos < b.pos; });
    h->res.n_snv = nc; h->res.snv = h->h_cand.data();
    const unsigned int ni = h->n_cnt[1];
    if (ni > h->ins_cap) return fail("gromgpu_chr_result: %u insertion candidates exceed the buffer of %u", ni, h->ins_cap);
    h->h_ins.resize(ni);
    if (ni) CK(cudaMemcpy(h->h_ins.data(), h->d_ins, sizeof(grom_ins_cand) * (size_t)ni, cudaMemcpyDeviceToHost));
    std::sort(h->h_ins.begin(), h->h_ins.end(), [](const grom_ins_cand &a, const grom_ins_cand &b) { return a.pos < b.pos; });
    h->res.n_ins = ni; h->res.ins = h->h_ins.data();
    const unsigned int nd = h->n_cnt[2];
    if (nd > h->del_cap) return fail("gromgpu_chr_result: %u deletion events exceed the buffer of %u", nd, h->del_cap);
    h->h_del.resize(nd);
    if (nd) CK(cudaMemcpy(h->h_del.data(), h->d_del, sizeof(grom_del_event) * (size_t)nd, cudaMemcpyDeviceToHost));
    std::sort(h->h_del.begin(), h->h_del.end(), [](const grom_del_event &a, const grom_del_event &b) { return a.pos != b.pos ? a.pos < b.pos : a.kind < b.kind; });
    h->res.n_del = nd; h->res.del_ev = h->h_del.data();
    const unsigned int ns = h->n_cnt[3];
    if (ns > h->svev_cap) return fail("gromgpu_chr_result: %u structural-variant gate events exceed the buffer of %u", ns, h->svev_cap);
    h->h_svev.resize(ns);
    if (ns) CK(cudaMemcpy(h->h_svev.data(), h->d_svev, sizeof(grom_sv_event) * (size_t)ns, cudaMemcpyDeviceToHost));
    {
        // scan order: by position, then in the order the reference evaluates the gates at one position
        static const int key[GROM_SV_CLASSES] = { 6, 7, 5, 4, 8, 10, 9, 11, 2, 3, 0, 1 };
        std::sort(h->h_svev.begin(), h->h_svev.end(), [](const grom_sv_event &a, const grom_sv_event &b) { return a.pos != b.pos ? a.pos < b.pos : key[a.cls] < key[b.cls]; });
    }
    h->res.n_sv = ns; h->res.sv_ev = h->h_svev.data();
    *out = h->res;
    return 0;
}

extern "C" int gromgpu_chr_finish(gromgpu_chr *h, gromgpu_result *out)
{
    if (gromgpu_chr_run(h)) return -1;
    return gromgpu_chr_result(h, out);
}

extern "C" int gromgpu_chr_stats(const gromgpu_chr *h, gromgpu_stats *out)
{
    if (!h || !h->ran) return fail("gromgpu_chr_stats: call gromgpu_chr_run first");
    *out = h->stats;
    return 0;
}

extern "C" int gromgpu_debug_fetch(gromgpu_chr *h, int ga, int32_t *dst, int64_t p0, int64_t p1)
{
    if (!h || !h->ran) return fail("gromgpu_debug_fetch: call gromgpu_chr_run first");
    ON_DEV();
    if (ga < 0 || ga >= GA_COUNT || p0 < 0 || p1 > h->P || p0 > p1) return fail("gromgpu_debug_fetch: bad array %d or range [%lld,%lld)", ga, (long long)p0, (long long)p1);
    CK(cudaMemcpy(dst, h->d_arrays + (int64_t)ga * h->Ppad + p0, sizeof(int32_t) * (size_t)(p1 - p0), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int gromgpu_debug_fetch_cluster(gromgpu_chr *h, int what, int cls, void *dst, int64_t p0, int64_t p1)
{
    if (!h || !h->ran) return fail("gromgpu_debug_fetch_cluster: call gromgpu_chr_run first");
    ON_DEV();
    if (p0 < 0 || p1 > h->P || p0 > p1 || what < 0 || what > 5 || cls < 0 || cls >= 10) return fail("gromgpu_debug_fetch_cluster: bad arguments");
    const int64_t Ppad = h->Ppad; const size_t cntp = (size_t)(p1 - p0);
    if (what == 3) { CK(cudaMemcpy(dst, h->d_cl_dist + (int64_t)cls * Ppad + p0, sizeof(double) * cntp, cudaMemcpyDeviceToHost)); return 0; }
    int64_t row;
    if (what <= 2) row = what * 10 + cls; else if (what == 4) { if (cls > 1) return fail("gromgpu_debug_fetch_cluster: ctx class index must be 0 or 1"); row = 30 + cls; } else row = 32;
    CK(cudaMemcpy(dst, h->d_cl_int + row * Ppad + p0, sizeof(int32_t) * cntp, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int gromgpu_fetch_read_state(gromgpu_chr *h, uint8_t *dst, int64_t i0, int64_t i1)
{
    if (!h || !h->ran) return fail("gromgpu_fetch_read_state: call gromgpu_chr_run first");
    ON_DEV();
    if (i0 < 0 || i1 > h->n_reads || i0 > i1) return fail("gromgpu_fetch_read_state: bad range");
    CK(cudaMemcpy(dst, h->d_state + i0, (size_t)(i1 - i0), cudaMemcpyDeviceToHost));
    return 0;
}


// ================================================================================================ read-depth CNV path (cnv.cuh)
// Host threads one call of gromgpu_chr_cnv may use for its host stages: the cores this process may run on (its affinity mask), shared
// by the processes of the node and the contigs in flight in each of them when the caller says how many there are
// (GROMGPU_HOST_THREADS = threads per call; set by the genome drivers and bench.py), never more than `most`.
static unsigned cnv_host_threads(unsigned most)
{
    static const unsigned avail = []() -> unsigned {
        if (const char *e = getenv("GROMGPU_HOST_THREADS")) { const long v = atol(e); if (v > 0) return (unsigned)v; }
        cpu_set_t set; CPU_ZERO(&set);
        if (sched_getaffinity(0, sizeof(set), &set) == 0) { const int n = CPU_COUNT(&set); if (n > 0) return (unsigned)n; }
        const unsigned h = std::thread::hardware_concurrency();
        return h ? h : 1u;
    }();
    return std::max(1u, std::min(most, avail));
}
struct CnvState {
    int32_t *d_depth = nullptr; uint8_t *d_mq8 = nullptr; uint32_t *d_rec = nullptr, *d_seed = nullptr;   // d_seed: [2][words]
    double *d_z = nullptr;                                             // z of the deletion scan per position (k_zfill)
    cnv::PreOut *d_pre = nullptr; unsigned long long *d_hist = nullptr;
    cnv::RepRec *d_rep = nullptr; unsigned int *d_nrep = nullptr; unsigned int rep_cap = 0;
    uint8_t *d_tile = nullptr;                                         // [4][n_tiles]: last/in of the mask stage, last/in of the z stage
    uint32_t *h_rec = nullptr, *h_seed = nullptr, *h_wp = nullptr, *h_land = nullptr; cnv::SeedCall *h_spec = nullptr;   // pinned
    uint32_t *d_blk = nullptr, *d_wp = nullptr, *d_land = nullptr; uint32_t land_cap = 0, spec_cap = 0; int nb = 0;
    cnv::SeedCall *d_spec = nullptr; unsigned int *d_nspec = nullptr; double *d_winsd = nullptr;
    uint32_t *d_u1 = nullptr; int32_t *d_ends = nullptr; static constexpr uint32_t ENDS_CAP = 1u << 20;
    int64_t P_cap = 0, words_cap = 0;                                  // what the buffers were sized for (a handle rebound to a shorter contig keeps them)
    std::vector<grom_cnv_call> calls; std::vector<double> tail_gm;
    std::vector<double> win_sd, win_thr, bin_d; std::vector<int64_t> win_cnt, bin_n;
    int64_t P = 0, words = 0;
    int q = 0;
    std::vector<double> sd_tbl, wtab;
    cnv::Grow tmp[16], jump, flags, hop_out, hop_sink, gather_rec, heads, head_out, mid, walk, spec_g[5];
    void *h_specg = nullptr; size_t h_specg_cap = 0; cudaEvent_t ev_spec = nullptr;       // early gather of the calls made at run heads (their copy numbers are computed beside the device work)
    uint32_t *d_open = nullptr;                              // [2 kinds][2 classes][words]: seeds still open after the first round
    cnv::SeedHead *h_heads = nullptr; cnv::HeadOutcome *h_head_out = nullptr;      // pinned
    uint32_t *h_win = nullptr; size_t h_win_cap = 0;         // pinned landing area of the record windows of the run heads
    void *h_gather = nullptr; size_t h_gather_cap = 0;      // pinned landing area of the gathered call ranges (depth, GC byte, record per position)
    cudaStream_t copy_stream = nullptr; cudaEvent_t ev_z = nullptr, ev_copied = nullptr, e0 = nullptr, e1 = nullptr;   // packed records travel to the host while the sweep runs
};
static void cnv_state_free(CnvState *c)
{
    if (!c) return;
    if (c->h_gather) cudaFreeHost(c->h_gather);
    cudaFree(c->d_depth); cudaFree(c->d_mq8); cudaFree(c->d_rec); cudaFree(c->d_z); cudaFree(c->d_seed); cudaFree(c->d_pre); cudaFree(c->d_hist);
    cudaFree(c->d_rep); cudaFree(c->d_nrep); cudaFree(c->d_tile);
    for (auto &g : c->tmp) if (g.p) cudaFree(g.p);
    for (cnv::Grow *g : {&c->jump, &c->flags, &c->hop_out, &c->hop_sink, &c->gather_rec, &c->heads, &c->head_out, &c->mid, &c->walk, &c->spec_g[0], &c->spec_g[1], &c->spec_g[2], &c->spec_g[3], &c->spec_g[4]}) if (g->p) cudaFree(g->p);
    if (c->h_specg) cudaFreeHost(c->h_specg);
    if (c->ev_spec) cudaEventDestroy(c->ev_spec);
    cudaFree(c->d_open);
    if (c->h_heads) cudaFreeHost(c->h_heads);
    if (c->h_head_out) cudaFreeHost(c->h_head_out);
    if (c->h_win) cudaFreeHost(c->h_win);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->ev_z) cudaEventDestroy(c->ev_z);
    if (c->ev_copied) cudaEventDestroy(c->ev_copied);
    if (c->e0) cudaEventDestroy(c->e0);
    if (c->e1) cudaEventDestroy(c->e1);
    if (c->h_rec) cudaFreeHost(c->h_rec);
    if (c->h_seed) cudaFreeHost(c->h_seed);
    if (c->h_wp) cudaFreeHost(c->h_wp);
    if (c->h_land) cudaFreeHost(c->h_land);
    if (c->h_spec) cudaFreeHost(c->h_spec);
    cudaFree(c->d_blk); cudaFree(c->d_wp); cudaFree(c->d_land); cudaFree(c->d_spec); cudaFree(c->d_nspec); cudaFree(c->d_winsd); cudaFree(c->d_u1); cudaFree(c->d_ends);
    delete c;
}

// the reference's bisections run literally (src/GROM.c:21630-21744); STRICT = bisect_right
template <bool STRICT> static long ref_bisect(const int *a, int v, long s, long e)
{
    auto less = [](int x, int y) { return STRICT ? x < y : x <= y; };
    long lo = s, hi = e, i = s + (e - s) / 2;
    for (;;) {
        if (i <= s) return less(v, a[s]) ? s : s + 1;
        if (i >= e - 1) return less(v, a[e - 1]) ? e - 1 : e;
        if (less(v, a[i])) { hi = i; i = lo + (i - lo) / 2; if (hi == i) return i + 1; }
        else { lo = i; i = i + (hi - i) / 2; if (lo == i) return i + 1; }
    }
}

namespace {
struct DevTmp {             // scoped device allocation
    void *p = nullptr;
    ~DevTmp() { if (p) cudaFree(p); }
    template <class T> T *as() { return (T *)p; }
};
}

extern "C" int gromgpu_chr_cnv(gromgpu_chr *h, const double *p2s_p, const double *p2s_sd, int n_p2s, int ploidy, gromgpu_cnv_result *out)
{
    using namespace cnv;
    if (!h || !h->ran) return fail("gromgpu_chr_cnv: call gromgpu_chr_run first");
    ON_DEV();
    if (!out || !p2s_p || !p2s_sd) return fail("gromgpu_chr_cnv: null argument");
    if (n_p2s != P2S) return fail("gromgpu_chr_cnv: the p-value table must have %d entries (got %d)", P2S, n_p2s);
    if (ploidy <= 0) return fail("gromgpu_chr_cnv: ploidy must be positive");
    const grom_params &prm = g_params;
    const int64_t P = h->P, Ppad = h->Ppad, M = prm.insert_mean, W1 = 2 * M - 1, lo = M - 1, hi = P - W1, half = M / 2;
    const int q = prm.rd_min_mapq, A = prm.windows_sampling_factor, Lmin = prm.min_rd_window_len, Lmax = prm.max_rd_window_len;
    const long cap = prm.sample_lists_len;
    if (Lmin < 1 || Lmax < Lmin || A < 1 || half < 1) return fail("gromgpu_chr_cnv: bad window parameters");
    memset(out, 0, sizeof(*out));
    out->biased_repeat = -1;
    cudaStream_t s = h->stream;
    if (!h->cnv) h->cnv = new CnvState();
    CnvState &c = *h->cnv;
    c.calls.clear(); c.q = q; c.P = P;
    c.sd_tbl.assign(p2s_sd, p2s_sd + P2S);
    c.win_sd.assign(Lmax + 1, 0.0); c.win_cnt.assign(Lmax + 1, 0); c.bin_d.assign(4 * NLIST, 0.0); c.bin_n.assign(NLIST, 0);
    out->win_sd = c.win_sd.data(); out->win_cnt = c.win_cnt.data();
    out->bin_ave = c.bin_d.data(); out->bin_sd = c.bin_d.data() + NLIST; out->bin_del_thr = c.bin_d.data() + 2 * NLIST; out->bin_dup_thr = c.bin_d.data() + 3 * NLIST;
    out->bin_n = c.bin_n.data();
    if (hi <= lo) return 0;                                            // contig shorter than the GC window: nothing is analysed
    if (!c.e0) { CK(cudaEventCreate(&c.e0)); CK(cudaEventCreate(&c.e1)); }
    cudaEvent_t e0 = c.e0, e1 = c.e1;
    const auto t_begin = std::chrono::steady_clock::now();
    double ms_dev = 0;
    int n_launch = 0;
    int64_t d2h = 0;
    const bool trace = getenv("GROMGPU_CNV_TRACE") != nullptr;
    auto t_last = t_begin;
    auto mark = [&](const char *what) {
        if (!trace) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[cnv] %-28s %8.3f ms (device so far %.3f)\n", what, std::chrono::duration<double, std::milli>(now - t_last).count(), ms_dev);
        t_last = now;
    };
    auto dev_begin = [&]() { cudaEventRecord(e0, s); };
    auto dev_end = [&]() { cudaEventRecord(e1, s); cudaEventSynchronize(e1); float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms_dev += ms; };

    const int32_t *A_mq = h->d_arrays + (int64_t)GA_RD_MQ * Ppad, *A_rd = h->d_arrays + (int64_t)GA_RD_RD * Ppad, *A_low = h->d_arrays + (int64_t)GA_RD_LOW * Ppad;
    const int32_t *A_gc = h->d_arrays + (int64_t)GA_GC * Ppad, *A_acgt = h->d_arrays + (int64_t)GA_ACGT * Ppad;
    const int64_t n_blk = (P + BLK_UNIT - 1) / BLK_UNIT, n_tiles = (P + CTILE - 1) / CTILE, words = (P + 31) / 32;
    c.words = words;
    if (c.d_depth && P > c.P_cap) return fail("gromgpu_chr_cnv: the handle's read-depth buffers hold %lld positions, the contig has %lld", (long long)c.P_cap, (long long)P);
    if (!c.d_depth) {
        c.P_cap = P; c.words_cap = words;
        CK(cudaMalloc(&c.d_depth, sizeof(int32_t) * P)); CK(cudaMalloc(&c.d_mq8, P)); CK(cudaMalloc(&c.d_rec, sizeof(uint32_t) * P)); CK(cudaMalloc(&c.d_z, sizeof(double) * P));
        CK(cudaMalloc(&c.d_seed, sizeof(uint32_t) * 3 * words));      // deletion seeds, duplication seeds, positions with a z value
        CK(cudaMalloc(&c.d_pre, sizeof(PreOut) * n_blk));
        CK(cudaMalloc(&c.d_open, sizeof(uint32_t) * 6 * words));       // [2][2][words] open seeds + [2][words] positions under calls made at run heads
        CK(cudaMalloc(&c.d_hist, sizeof(unsigned long long) * HIST_ALL));
        c.rep_cap = (unsigned int)(P / 20 + 2);
        CK(cudaMalloc(&c.d_rep, sizeof(RepRec) * c.rep_cap)); CK(cudaMalloc(&c.d_nrep, sizeof(unsigned int)));
        CK(cudaMalloc(&c.d_tile, 4 * n_tiles));
        CK(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c.ev_z, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c.ev_copied, cudaEventDisableTiming));
        c.nb = (int)((words + SEED_WORDS - 1) / SEED_WORDS); c.land_cap = (uint32_t)(P / 4 + 1024);
        CK(cudaMalloc(&c.d_blk, sizeof(uint32_t) * (3 * c.nb + 4))); CK(cudaMalloc(&c.d_wp, sizeof(uint32_t) * 3 * words));
        CK(cudaMalloc(&c.d_land, sizeof(uint32_t) * 4 * (size_t)c.land_cap));
        c.spec_cap = (uint32_t)std::min<int64_t>(P / 8 + 1024, (int64_t)1 << 28);
        CK(cudaMalloc(&c.d_spec, sizeof(SeedCall) * (size_t)c.spec_cap));
        CK(cudaMalloc(&c.d_nspec, 16 * sizeof(unsigned int))); CK(cudaMalloc(&c.d_winsd, sizeof(double) * (3 * (Lmax + 1) + 1)));      // win_sd, win_thr, then the tail bound gm [Lmax + 2]
        CK(cudaMalloc(&c.d_u1, sizeof(uint32_t) * 8 * words)); CK(cudaMalloc(&c.d_ends, sizeof(int32_t) * CnvState::ENDS_CAP));      // d_u1: [2][words] listed seeds, [2][words] safe stretch ends, [4][words] runs whose last seed stayed open
    }

    // ---- stage 1: pre-statistics + repeat runs
    dev_begin();
    CK(cudaMemsetAsync(c.d_hist, 0, sizeof(unsigned long long) * HIST_ALL, s));
    CK(cudaMemsetAsync(c.d_nrep, 0, sizeof(unsigned int), s));
    k_pre<<<(unsigned)n_blk, 256, 0, s>>>(A_mq, A_rd, A_low, A_acgt, h->d_fasta, P, lo, hi, c.d_depth, c.d_mq8, c.d_pre, c.d_hist); n_launch++;
    k_repeats<<<(unsigned)((hi - lo + 255) / 256), 256, 0, s>>>(h->d_fasta, c.d_depth, lo, hi, c.d_rep, c.rep_cap, c.d_nrep); n_launch++;
    std::vector<PreOut> pre(n_blk);
    std::vector<unsigned long long> hist(HIST_ALL);
    unsigned int n_rep = 0;
    CK(cudaMemcpyAsync(pre.data(), c.d_pre, sizeof(PreOut) * n_blk, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(hist.data(), c.d_hist, sizeof(unsigned long long) * HIST_ALL, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&n_rep, c.d_nrep, sizeof(n_rep), cudaMemcpyDeviceToHost, s));
    dev_end();
    CK(cudaGetLastError());
    mark("stage1 kernels+D2H");
    if (n_rep > c.rep_cap) return fail("gromgpu_chr_cnv: %u repeat runs exceed the buffer of %u", n_rep, c.rep_cap);
    std::vector<RepRec> reps(n_rep);
    if (n_rep) CK(cudaMemcpy(reps.data(), c.d_rep, sizeof(RepRec) * n_rep, cudaMemcpyDeviceToHost));
    std::sort(reps.begin(), reps.end(), [](const RepRec &a, const RepRec &b) { return a.s < b.s; });
    out->n_repeats = n_rep;

    // contig mean / sd (src/GROM.c:16648-16686); the sd sums squares per depth value instead of per position
    unsigned long long ave_sum = 0, ave_cnt = 0, acgt_sum = 0, acgt_cnt = 0;
    for (const PreOut &b : pre) { ave_sum += b.ave_sum; ave_cnt += b.ave_cnt; acgt_sum += b.acgt_sum; acgt_cnt += b.acgt_cnt; }
    double chr_ave = (double)ave_sum, chr_sd = 0;
    if (ave_cnt > 0) chr_ave = chr_ave / (double)(long)ave_cnt;
    for (int d = 0; d < HIST_ALL; d++) if (hist[d]) {
        const double term = (double)d < 2 * chr_ave ? ((double)d - chr_ave) * ((double)d - chr_ave) : chr_ave * chr_ave;
        chr_sd += term * (double)hist[d];
    }
    chr_sd = ave_cnt > 1 ? sqrt(chr_sd / ((double)ave_cnt - 1.0)) : 0.0;
    out->chr_ave = chr_ave; out->chr_sd = chr_sd;
    // per-type repeat depth and the most biased type (src/GROM.c:16693-16774)
    int biased = -1;
    {
        double r_ave[10] = {0}, r_sd[10] = {0}; long r_cnt[10] = {0};
        std::vector<double> rl(n_rep);
        for (unsigned i = 0; i < n_rep; i++) {
            rl[i] = (double)reps[i].depth_sum / (double)(reps[i].e - reps[i].s);
            r_ave[reps[i].type] += rl[i] < 2 * chr_ave ? rl[i] : 2 * chr_ave;
            r_cnt[reps[i].type]++;
        }
        for (int k = 0; k < 10; k++) r_ave[k] = r_ave[k] / (double)r_cnt[k];
        for (unsigned i = 0; i < n_rep; i++) {
            const int t = reps[i].type;
            const double x = rl[i] < 2 * chr_ave ? rl[i] : 2 * chr_ave;
            r_sd[t] += (x - r_ave[t]) * (x - r_ave[t]);
        }
        long best = 0;
        for (int k = 0; k < 10; k++) {
            r_sd[k] = r_cnt[k] > 1 ? sqrt(r_sd[k] / ((double)r_cnt[k] - 1.0)) : 0.0;
            if (r_cnt[k] > NO_COMBINE && r_ave[k] + 1.5 * r_sd[k] < chr_ave && chr_ave - 1.5 * chr_sd > r_ave[k] && r_cnt[k] > best) { biased = k; best = r_cnt[k]; }
        }
    }
    out->biased_repeat = biased;
    // 10 kb blocks above twice the contig mean -> runs -> their complement is what gets sampled (src/GROM.c:16784-16990)
    std::vector<int64_t> sb_s, sb_e;
    {
        const int64_t nfull = P / BLK_UNIT;
        const double blk_ave = (double)(long)acgt_sum / (double)(long)acgt_cnt, thr = 2 * blk_ave;
        out->blk_ave = blk_ave;
        std::vector<long> over;
        for (int64_t k = 0; k < nfull; k++) if ((double)(long)pre[k].blk_sum / (double)BLK_UNIT > thr) over.push_back((long)k);
        std::vector<long> bs(10001, 0), be(10001, 0);
        long run = 0, r_s = 0, r_e = 0, bi = 0;
        for (size_t a = 1; a < over.size(); a++) {
            if (run == 0) {
                if (run + 1 > (over[a] - over[a - 1]) / 4) { r_e = over[a] + 1; run++; } else r_e = over[a - 1] + 1;
                r_s = over[a - 1]; run++;
            } else {
                if (run + 1 > (over[a - 1] - r_s) / 4) { r_e = over[a - 1] + 1; run++; }
                else { if (run >= 4) bi++; r_s = over[a - 1]; r_e = over[a - 1] + 1; run = 1; }
                if (run >= 4 && bi < 10000) { bs[bi] = r_s * BLK_UNIT; be[bi] = r_e * BLK_UNIT; }
            }
        }
        if (run >= 4) bi++;
        std::vector<long> ls(bi + 3, 0), le(bi + 3, 0);
        long li = 0;
        for (long a = 0; a < bi && a < 10000; a++) if (be[a] - bs[a] >= 10000) { le[li] = bs[a]; ls[li + 1] = be[a]; li++; }
        li++;
        le[li - 1] = P;
        for (long k = 0; k < li; k++) {
            if (ls[k] < lo) ls[k] = lo; else if (ls[k] >= hi) ls[k] = hi;
            if (le[k] < lo) le[k] = lo; else if (le[k] >= hi) le[k] = hi;
            if (le[k] - ls[k] >= Lmin) { sb_s.push_back(ls[k]); sb_e.push_back(le[k]); }
        }
    }
    mark("pre-statistics host");
    const int n_sb = (int)sb_s.size();
    out->n_sample_blocks = n_sb;

    // ---- stage 2: depth samples -> per-bin sorted lists (host: reservoir with libc-rand semantics), statistics, rank tables
    std::vector<int64_t> sb_first(n_sb + 1, 0);
    for (int k = 0; k < n_sb; k++) sb_first[k + 1] = sb_first[k] + (sb_e[k] - sb_s[k] + half - 1) / half;
    const int64_t n_samples = sb_first[n_sb];
    out->n_samples = n_samples;
    std::vector<Sample> samples(n_samples);
    Grow &t_sb = c.tmp[0], &t_first = c.tmp[1], &t_samples = c.tmp[2];
    if (n_samples) {
        if (!t_sb.ensure(sizeof(int64_t) * n_sb) || !t_first.ensure(sizeof(int64_t) * (n_sb + 1)) || !t_samples.ensure(sizeof(Sample) * n_samples)) return fail("gromgpu_chr_cnv: out of device memory");
        dev_begin();
        CK(cudaMemcpyAsync(t_sb.p, sb_s.data(), sizeof(int64_t) * n_sb, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(t_first.p, sb_first.data(), sizeof(int64_t) * (n_sb + 1), cudaMemcpyHostToDevice, s));
        k_samples<<<(unsigned)((n_samples + 255) / 256), 256, 0, s>>>(c.d_depth, A_rd, A_low, c.d_mq8, A_gc, A_acgt, t_sb.as<int64_t>(), t_first.as<int64_t>(), n_sb,
                                                                     n_samples, half, q, t_samples.as<Sample>()); n_launch++;
        CK(cudaMemcpyAsync(samples.data(), t_samples.p, sizeof(Sample) * n_samples, cudaMemcpyDeviceToHost, s));
        dev_end();
        CK(cudaGetLastError());
    }
    mark("samples kernel+D2H");
    GlibcRand rng((unsigned)prm.rand_seed);
    // most-biased repeat: depth samples by distance segment around every run of that type (src/GROM.c:18262-18367)
    constexpr int SEG = 10;
    std::vector<int64_t> rp_start, rp_first;      // gathered ranges around the biased repeats
    std::vector<int32_t> rp_depth; std::vector<uint8_t> rp_gc;
    SampleList rsl[SEG];
    double rs_ave[SEG] = {0}, rs_sd[SEG] = {0};
    std::vector<unsigned> biased_idx;
    auto gather = [&](const std::vector<int64_t> &starts, const std::vector<int64_t> &firsts, std::vector<int32_t> &o_depth, std::vector<uint8_t> &o_gc, std::vector<uint32_t> *o_rec = nullptr, void **pinned_out = nullptr) -> int {
        const int n_seg = (int)starts.size();
        const int64_t total = firsts.back();
        if (!pinned_out) { o_depth.resize(total); o_gc.resize(total); if (o_rec) o_rec->resize(total); }
        if (!total) return 0;
        Grow &a = c.tmp[3], &b = c.tmp[4], &od = c.tmp[5], &og = c.tmp[6];
        if (!a.ensure(sizeof(int64_t) * n_seg) || !b.ensure(sizeof(int64_t) * (n_seg + 1)) || !od.ensure(sizeof(int32_t) * total) || !og.ensure(total) ||
            (o_rec && !c.gather_rec.ensure(sizeof(uint32_t) * total))) return fail("gromgpu_chr_cnv: out of device memory");
        dev_begin();
        CK(cudaMemcpyAsync(a.p, starts.data(), sizeof(int64_t) * n_seg, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(b.p, firsts.data(), sizeof(int64_t) * (n_seg + 1), cudaMemcpyHostToDevice, s));
        k_gather<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(c.d_depth, A_gc, A_acgt, a.as<int64_t>(), b.as<int64_t>(), n_seg, total, od.as<int32_t>(), og.as<uint8_t>(), c.d_rec, o_rec ? c.gather_rec.as<uint32_t>() : nullptr); n_launch++;
        if (pinned_out) {
            // copy-number ranges: straight into a pinned landing area (no zero-filled vectors, no pageable staging)
            const size_t need = (size_t)total * 9 + 64;
            if (need > c.h_gather_cap) {
                if (c.h_gather) cudaFreeHost(c.h_gather);
                c.h_gather = nullptr; c.h_gather_cap = 0;
                CK(cudaMallocHost(&c.h_gather, need + need / 4));
                c.h_gather_cap = need + need / 4;
            }
            int32_t *pd = (int32_t *)c.h_gather; uint32_t *pr = (uint32_t *)(pd + total); uint8_t *pg = (uint8_t *)(pr + total);
            CK(cudaMemcpyAsync(pd, od.p, sizeof(int32_t) * total, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(pr, c.gather_rec.p, sizeof(uint32_t) * total, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(pg, og.p, total, cudaMemcpyDeviceToHost, s));
            pinned_out[0] = pd; pinned_out[1] = pr; pinned_out[2] = pg;
            dev_end();
            CK(cudaGetLastError());
            return 0;
        }
        CK(cudaMemcpyAsync(o_depth.data(), od.p, sizeof(int32_t) * total, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(o_gc.data(), og.p, total, cudaMemcpyDeviceToHost, s));
        if (o_rec) CK(cudaMemcpyAsync(o_rec->data(), c.gather_rec.p, sizeof(uint32_t) * total, cudaMemcpyDeviceToHost, s));
        dev_end();
        CK(cudaGetLastError());
        return 0;
    };
    auto rep_segment = [&](int64_t p, const RepRec &r) -> int {
        if (p < r.s) return (int)((SEG - 1) * (p - (r.s - half)) / half);
        if (p >= r.e) return (int)((SEG - 1) * ((r.e + half) - p) / half);
        return SEG - 1;
    };
    if (biased != -1) {
        rp_first.push_back(0);
        for (unsigned i = 0; i < n_rep; i++) if (reps[i].type == biased) {
            const int64_t a = std::max<int64_t>(0, reps[i].s - half), b = std::min<int64_t>(P, reps[i].e + half);
            biased_idx.push_back(i); rp_start.push_back(a); rp_first.push_back(rp_first.back() + (b - a));
        }
        if (gather(rp_start, rp_first, rp_depth, rp_gc)) return -1;
        for (size_t k = 0; k < biased_idx.size(); k++) {
            const RepRec &r = reps[biased_idx[k]];
            for (int64_t j = rp_first[k]; j < rp_first[k + 1]; j++) {
                const int64_t p = rp_start[k] + (j - rp_first[k]);
                if (rp_gc[j] & 0x80) rsl[rep_segment(p, r)].add(rp_depth[j], cap, rng);
            }
        }
        for (int k = 0; k < SEG; k++) {
            std::sort(rsl[k].v.begin(), rsl[k].v.end());
            const long n = (long)rsl[k].v.size();
            if (n > 0) {
                const long a = n / 20, b = n - a, m = b - a;
                double sm = 0, v = 0;
                for (long j = a; j < b; j++) sm += rsl[k].v[j];
                rs_ave[k] = sm / m;
                for (long j = a; j < b; j++) v += (rsl[k].v[j] - rs_ave[k]) * (rsl[k].v[j] - rs_ave[k]);
                rs_sd[k] = m > 1 ? sqrt(v / (m - 1)) : v;
            }
        }
    }
    // GC-stratified lists: an uncovered sample joins the list of the last covered one (src/GROM.c:18373-18456)
    std::vector<SampleList> lists(NLIST);
    {
        int last_low = 0;
        for (int64_t j = 0; j < n_samples; j++) {
            const int code = samples[j].code;
            if (!(code & 1)) continue;
            const int cls = (code >> 1) & 3, g = code >> 8;
            int to_low;
            if (cls == 2) to_low = last_low; else to_low = last_low = cls;
            lists[to_low * NB + g].add(samples[j].depth, cap, rng);
        }
    }
    // depths are small non-negative integers: counting sort
    auto sort_depths = [](std::vector<int> &v) {
        if (v.size() < 64) { std::sort(v.begin(), v.end()); return; }
        int mx = 0, mn = 0;
        for (int x : v) { mx = std::max(mx, x); mn = std::min(mn, x); }
        if (mn < 0 || mx > (1 << 20)) { std::sort(v.begin(), v.end()); return; }
        std::vector<int> h(mx + 1, 0);
        for (int x : v) h[x]++;
        size_t k = 0;
        for (int d = 0; d <= mx; d++) for (int j = 0; j < h[d]; j++) v[k++] = d;
    };
    // the per-list work below is independent per list: a few host threads share it
    auto par_lists = [&](auto &&fn) {
        const int T = (int)cnv_host_threads(8);
        std::vector<std::thread> pool;
        for (int t = 1; t < T; t++) pool.emplace_back([&, t]() { for (int l = t; l < NLIST; l += T) fn(l); });
        for (int l = 0; l < NLIST; l += T) fn(l);
        for (auto &x : pool) x.join();
    };
    par_lists([&](int l) { sort_depths(lists[l].v); });
    {
        // thin bins (20 <= n < 100) borrow the original samples of the two bins on either side (src/GROM.c:18481-18548)
        std::vector<std::vector<int>> grown(NLIST);
        par_lists([&](int l) {
            const int m = l / NB, b = l % NB;
            if (m >= 2 || b < 2 || b >= NB - 2) return;
            const auto &me = lists[l].v;
            if ((long)me.size() < MIN_WINDOWS || (long)me.size() >= NO_COMBINE) return;
            std::vector<int> g(me);
            for (int a = b - 2; a <= b + 2; a++) if (a != b) for (int x : lists[m * NB + a].v) if ((long)g.size() < cap) g.push_back(x);
            sort_depths(g);
            grown[l] = std::move(g);
        });
        for (int l = 0; l < NLIST; l++) if (!grown[l].empty()) lists[l].v = std::move(grown[l]);
    }
    const double del_f = 1.0 - 0.6 / ploidy, dup_f = 1.0 + 0.6 / ploidy;
    std::vector<double> ave(NLIST, 0.0), sdv(NLIST, 0.0), del_thr(NLIST, 0.0), dup_thr(NLIST, 0.0);
    std::vector<int32_t> nlist(NLIST, 0), small(2 * NLIST, 0), top(NLIST, 0);
    par_lists([&](int l) {
        const auto &v = lists[l].v;
        const long n = (long)v.size();
        nlist[l] = (int32_t)n;
        if (n > 0) {
            double sm = 0, var = 0;
            for (long j = 0; j < n; j++) sm += v[j];
            ave[l] = sm / n; del_thr[l] = del_f * ave[l]; dup_thr[l] = dup_f * ave[l];
            for (long j = 0; j < n; j++) var += (v[j] - ave[l]) * (v[j] - ave[l]);
            sdv[l] = n > 1 ? sqrt(var / (n - 1)) : var;
            top[l] = v.back();
            small[2 * l] = v[0]; small[2 * l + 1] = n > 1 ? v[1] : v[0];
        }
        c.bin_d[l] = ave[l]; c.bin_d[NLIST + l] = sdv[l]; c.bin_d[2 * NLIST + l] = del_thr[l]; c.bin_d[3 * NLIST + l] = dup_thr[l]; c.bin_n[l] = n;
    });
    int D = 0;
    for (int l = 0; l < NLIST; l++) D = std::max(D, top[l]);
    if ((int64_t)NLIST * (D + 1) > (int64_t)1 << 30) return fail("gromgpu_chr_cnv: sampled depth %d is too large for the rank tables", D);
    std::vector<int32_t> cum((size_t)NLIST * (D + 1), 0);
    par_lists([&](int l) {
        int32_t *row = cum.data() + (size_t)l * (D + 1);
        for (int x : lists[l].v) row[x]++;
        for (int d = 1; d <= D; d++) row[d] += row[d - 1];
    });
    Grow &t_cum = c.tmp[7], &t_n = c.tmp[8], &t_small = c.tmp[9], &t_dbl = c.tmp[10];
    if (!t_cum.ensure(sizeof(int32_t) * cum.size()) || !t_n.ensure(sizeof(int32_t) * NLIST) || !t_small.ensure(sizeof(int32_t) * 2 * NLIST) ||
        !t_dbl.ensure(sizeof(double) * (3 * NLIST + 2 * P2S + 256))) return fail("gromgpu_chr_cnv: out of device memory");
    std::vector<double> dbl(3 * NLIST + 2 * P2S + 256);
    c.wtab.resize(256);
    for (int m = 0; m < 256; m++) c.wtab[m] = dbl[3 * NLIST + 2 * P2S + m] = 0.5 + (1.0 - 0.5) * (m - q) / (double)(RD_MAX_MAPQ - q);      // rec_z's weight, per mean MAPQ
    std::copy(ave.begin(), ave.end(), dbl.begin()); std::copy(del_thr.begin(), del_thr.end(), dbl.begin() + NLIST); std::copy(dup_thr.begin(), dup_thr.end(), dbl.begin() + 2 * NLIST);
    std::copy(p2s_p, p2s_p + P2S, dbl.begin() + 3 * NLIST); std::copy(p2s_sd, p2s_sd + P2S, dbl.begin() + 3 * NLIST + P2S);
    Tables T;
    T.cum = t_cum.as<int32_t>(); T.D = D; T.n = t_n.as<int32_t>(); T.small = t_small.as<int32_t>();
    T.ave = t_dbl.as<double>(); T.del_thr = T.ave + NLIST; T.dup_thr = T.ave + 2 * NLIST; T.p2s_p = T.ave + 3 * NLIST; T.p2s_sd = T.p2s_p + P2S;

    mark("lists+tables host");
    // ---- stage 3: mask, z, seeds, window sweep
    // walk of every sample block at each -A offset, cut into frames of Lmax elements
    std::vector<SweepBlock> sw(n_sb);
    int64_t n_frames = 0;
    for (int k = 0; k < n_sb; k++) {
        int64_t total = 0;
        for (int a = 0; a < A; a++) total += std::max<int64_t>(0, sb_e[k] - (sb_s[k] + (int64_t)a * Lmax / A));
        sw[k].start = sb_s[k]; sw[k].end = sb_e[k]; sw[k].first_frame = n_frames; sw[k].n_frames = (total + Lmax - 1) / Lmax;
        n_frames += sw[k].n_frames;
    }
    out->n_frames = n_frames;
    const int n_len = Lmax - Lmin + 1;
    Grow &t_sw = c.tmp[11], &t_X = c.tmp[12], &t_wsq = c.tmp[13], &t_wcnt = c.tmp[14];
    if (n_frames) {
        if (!t_sw.ensure(sizeof(SweepBlock) * n_sb) || !t_X.ensure(sizeof(double) * (size_t)n_len * n_frames) || !t_wsq.ensure(sizeof(double) * n_len) ||
            !t_wcnt.ensure(sizeof(long long) * n_len)) return fail("gromgpu_chr_cnv: out of device memory for the window sweep (%lld frames)", (long long)n_frames);
    }
    uint8_t *tl_mask = c.d_tile, *ti_mask = c.d_tile + n_tiles, *tl_z = c.d_tile + 2 * n_tiles, *ti_z = c.d_tile + 3 * n_tiles;
    std::vector<double> wsq(n_len, 0.0); std::vector<long long> wcnt(n_len, 0);
    dev_begin();
    CK(cudaMemcpyAsync(t_cum.p, cum.data(), sizeof(int32_t) * cum.size(), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(t_n.p, nlist.data(), sizeof(int32_t) * NLIST, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(t_small.p, small.data(), sizeof(int32_t) * 2 * NLIST, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(t_dbl.p, dbl.data(), sizeof(double) * dbl.size(), cudaMemcpyHostToDevice, s));
    k_tile_last_mask<<<(unsigned)n_tiles, 256, 0, s>>>(c.d_depth, c.d_mq8, A_acgt, lo, hi, q, tl_mask); n_launch++;
    k_carry_scan<<<1, 1024, 0, s>>>(tl_mask, ti_mask, (int)n_tiles); n_launch++;
    k_mask<<<(unsigned)n_tiles, 256, 0, s>>>(c.d_depth, c.d_mq8, A_gc, A_acgt, P, lo, hi, q, T.n, ti_mask, c.d_rec, tl_z); n_launch++;
    k_carry_scan<<<1, 1024, 0, s>>>(tl_z, ti_z, (int)n_tiles); n_launch++;
    k_z<<<(unsigned)n_tiles, 256, 0, s>>>(c.d_depth, A_gc, P, lo, hi, q, T, ti_z, c.d_rec, c.d_seed, c.d_seed + words); n_launch++;
    uint32_t seed_tot[2] = {0, 0};
    k_seed_blocksum<<<dim3((unsigned)c.nb, 2), 256, 0, s>>>(c.d_seed, words, c.d_blk, c.nb); n_launch++;
    k_seed_blockscan<<<2, 1024, 0, s>>>(c.d_blk, c.nb, c.d_blk + 3 * c.nb); n_launch++;
    k_seed_rank<<<dim3((unsigned)c.nb, 2), 256, 0, s>>>(c.d_seed, words, c.d_blk, c.nb, c.d_wp); n_launch++;
    CK(cudaMemcpyAsync(seed_tot, c.d_blk + 3 * c.nb, sizeof(seed_tot), cudaMemcpyDeviceToHost, s));
    if (n_frames) {
        CK(cudaMemcpyAsync(t_sw.p, sw.data(), sizeof(SweepBlock) * n_sb, cudaMemcpyHostToDevice, s));
        k_sweep<<<(unsigned)((n_frames + 31) / 32), 32 * SW_WARPS, 0, s>>>(c.d_rec, t_sw.as<SweepBlock>(), n_sb, n_frames, A, Lmin, Lmax, q, T.p2s_sd, T.p2s_sd + P2S, t_X.as<double>()); n_launch++;
        k_sweep_sum<<<(unsigned)((n_len + 31) / 32), 256, SSUM_ST * SSUM_FR * 32 * sizeof(double), s>>>(t_X.as<double>(), n_frames, n_len, t_wsq.as<double>(), t_wcnt.as<long long>()); n_launch++;
        CK(cudaMemcpyAsync(wsq.data(), t_wsq.p, sizeof(double) * n_len, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(wcnt.data(), t_wcnt.p, sizeof(long long) * n_len, cudaMemcpyDeviceToHost, s));
    }
    dev_end();
    CK(cudaGetLastError());
    mark("stage3 kernels+D2H");
    // the packed records stay on the device; only the rare paths below (biased-repeat override, seed tables that outgrew their
    // buffers) pull all of them to the host
    bool have_host_rec = false;
    auto pull_records = [&]() -> int {
        if (have_host_rec) return 0;
        // pinned landing areas of the rare host paths, sized for the handle's capacity, allocated the first time one of them runs
        if (!c.h_rec) { CK(cudaMallocHost(&c.h_rec, sizeof(uint32_t) * c.P_cap)); CK(cudaMallocHost(&c.h_seed, sizeof(uint32_t) * 2 * c.words_cap)); CK(cudaMallocHost(&c.h_wp, sizeof(uint32_t) * 2 * c.words_cap)); }
        CK(cudaMemcpyAsync(c.h_rec, c.d_rec, sizeof(uint32_t) * P, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(c.h_seed, c.d_seed, sizeof(uint32_t) * 2 * words, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(c.h_wp, c.d_wp, sizeof(uint32_t) * 2 * words, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        d2h += 4 * P + 16 * words;
        have_host_rec = true;
        return 0;
    };
    for (int L = Lmin; L <= Lmax; L++) {
        c.win_cnt[L] = wcnt[L - Lmin];
        c.win_sd[L] = wcnt[L - Lmin] > 1 ? sqrt(wsq[L - Lmin] / (double)(wcnt[L - Lmin] - 1)) : 0.0;
    }
    // pre-filter of SegCtx::scores: 2.97 sd per window length, +inf where the reference's `win_sd > 0` test fails
    c.win_thr.assign(Lmax + 1, std::numeric_limits<double>::infinity());
    for (int L = Lmin; L <= Lmax; L++) if (c.win_sd[L] > 0) c.win_thr[L] = 2.97 * c.win_sd[L];

    // most-biased repeat override of the z list, after the sweep like the reference (src/GROM.c:19023-19150)
    if (biased != -1) {
        if (pull_records()) return -1;
        for (size_t k = 0; k < biased_idx.size(); k++) {
            const RepRec &r = reps[biased_idx[k]];
            for (int64_t j = rp_first[k]; j < rp_first[k + 1]; j++) {
                const int64_t p = rp_start[k] + (j - rp_first[k]);
                uint32_t &rec = c.h_rec[p];
                if (rec & R_MASK) continue;
                const int sg = rep_segment(p, r), d = rp_depth[j];
                const std::vector<int> &v = rsl[sg].v;
                const long n = (long)v.size();
                long i1, i2; bool neg;
                if ((double)d < rs_ave[sg]) { i1 = ref_bisect<true>(v.data(), d, 0, n); i2 = ref_bisect<false>(v.data(), d, 0, n); neg = false; }
                else {
                    if ((double)d > 2 * rs_ave[sg]) i1 = ref_bisect<false>(v.data(), (int)(2 * rs_ave[sg]), 0, n); else i1 = ref_bisect<false>(v.data(), d, 0, n);
                    i2 = ref_bisect<true>(v.data(), d, 0, n);
                    i1 = n - i1; i2 = n - i2; neg = true;
                }
                const double prob = ((i1 <= 0 ? 0.5 : (double)i1) + (i2 <= 0 ? 0.5 : (double)i2)) / (double)(2 * n);
                long kk = std::upper_bound(p2s_p, p2s_p + P2S, prob) - p2s_p;
                if (kk >= P2S) kk = P2S - 1;
                rec = (rec & ~(R_NEG | (1023u << R_K))) | R_NZ | R_OVR | (neg ? R_NEG : 0u) | ((uint32_t)kk << R_K);
            }
        }
        CK(cudaMemcpyAsync(c.d_rec, c.h_rec, sizeof(uint32_t) * P, cudaMemcpyHostToDevice, s));      // keeps gromgpu_cnv_fetch consistent
        CK(cudaStreamSynchronize(s));
    }

    // ---- stage 4: greedy segmentation (two host threads: deletions, duplications) and copy number
    std::vector<Call> found[2];
    int64_t seed_tot_all = 0, n_spec_all = 0;
    // copy number of one call from its gathered positions (src/GROM.c:20071-20153): ratios depth / bin mean over the unmasked positions,
    // the reference's qsort (merge tree, low-word comparator), 10 %-trimmed mean times the ploidy, sd around it over all ratios
    auto copy_number = [&](const int32_t *g_depth_, const uint32_t *g_rec_, const uint8_t *g_gc_, int64_t n_pos, std::vector<double> &buf, std::vector<double> &tmp, double *cn, double *cn_sd) {
        *cn = -1; *cn_sd = 0;
        buf.clear();
        for (int64_t j = 0; j < n_pos; j++) {
            const uint32_t r = g_rec_[j];
            if (r & R_MASK) continue;
            const int l = ((((r >> R_CLASS) & 3) == 0) ? 0 : NB) + (g_gc_[j] & 0x7f);
            if (ave[l] > 0) buf.push_back((double)g_depth_[j] / ave[l]);
        }
        const long n = (long)buf.size();
        if (n <= 0) return;
        tmp.resize(n);
        lowword_msort_par(buf.data(), n, tmp.data(), n >= 65536 ? 3 : (n >= 16384 ? 2 : 0));
        const long a = (long)(0.1 * n), b = n - a;
        double tot = 0;
        for (long j = a; j < b; j++) tot += buf[j];
        if (b - a > 0) {
            *cn = (tot / (b - a)) * ploidy;
            double v = 0;
            for (long j = 0; j < n; j++) { const double dd = ploidy * buf[j] - *cn; v += dd * dd; }
            *cn_sd = sqrt(v / n);
        }
    };
    // The calls made at run heads are the long ones and are known long before the path is: their positions are gathered right away and
    // their copy numbers computed by a background thread while the device runs the second round and the hop; the final pass below
    // takes them from here (a call that turns out not to be on the path is simply not used).
    struct SpecCN {
        std::thread th;
        std::vector<int64_t> start, end, first; std::vector<int> kind; std::vector<double> cn, cn_sd;
        ~SpecCN() { if (th.joinable()) th.join(); }
    } spec_cn;
    {
        SegCtx ctx[2];
        for (int k = 0; k < 2; k++) { ctx[k].rec = c.d_rec; ctx[k].len = P; ctx[k].end = hi - Lmin; ctx[k].q = q; ctx[k].Lmin = Lmin; ctx[k].Lmax = Lmax; ctx[k].bound = SEED_BOUND; ctx[k].sd = T.p2s_sd; ctx[k].win_sd = c.d_winsd; ctx[k].win_thr = c.d_winsd + (Lmax + 1); ctx[k].zarr = c.d_z; ctx[k].dup = k == 1; ctx[k].wtab = T.p2s_sd + P2S; }
        // every seed evaluated on the device (bounded); a seed list that outgrew its buffer, or the biased-repeat override (it rewrites
        // z on the host copy after the sweep), leaves the evaluation to the host
        unsigned int n_spec = 0;
        bool device_hop = false;
        const bool have_land = !getenv("GROMGPU_CNV_NO_SEED_TABLES") && seed_tot[0] <= c.land_cap && seed_tot[1] <= c.land_cap && (int64_t)Lmax + SEED_BOUND < ((int64_t)1 << LAND_SHIFT) && hi - Lmin > lo;
        if (have_land) {
            dev_begin();
            CK(cudaMemcpyAsync(c.d_winsd, c.win_sd.data(), sizeof(double) * (Lmax + 1), cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(c.d_winsd + (Lmax + 1), c.win_thr.data(), sizeof(double) * (Lmax + 1), cudaMemcpyHostToDevice, s));
            {
                // suffix minimum of L * win_thr[L]: the smallest total a window of at least L positions needs to score (k_tail_check)
                c.tail_gm.assign(Lmax + 2, std::numeric_limits<double>::infinity());
                for (int L = Lmax; L >= 0; L--) c.tail_gm[L] = std::min(c.tail_gm[L + 1], L >= Lmin ? (double)L * c.win_thr[L] : std::numeric_limits<double>::infinity());
                CK(cudaMemcpyAsync(c.d_winsd + 2 * (Lmax + 1), c.tail_gm.data(), sizeof(double) * (Lmax + 2), cudaMemcpyHostToDevice, s));
                CK(cudaMemsetAsync(c.d_u1, 0, sizeof(uint32_t) * 8 * (size_t)words, s));
            }
            CK(cudaMemsetAsync(c.d_nspec, 0, 16 * sizeof(unsigned int), s));
            k_zfill<<<148 * 8, 256, 0, s>>>(c.d_rec, P, T.p2s_sd, T.p2s_sd + P2S, c.d_z); n_launch++;
            // every (seed, carried class) can stay open after the first round (a contig full of long events): room for all of them
            const uint32_t todo_cap = (uint32_t)std::min<uint64_t>(2ull * ((uint64_t)seed_tot[0] + seed_tot[1]) + 64, 1ull << 28);
            Grow &t_todo = c.tmp[15];
            if (!t_todo.ensure(sizeof(SeedTodo) * (size_t)todo_cap)) return fail("gromgpu_chr_cnv: out of device memory");
            // jump table: level k holds the node reached after 2^k hops; node ids of the duplication scan sit behind the deletion scan's
            const uint32_t n_nodes = 2 * seed_tot[0] + 1 + 2 * seed_tot[1] + 1, base[2] = {0u, 2 * seed_tot[0] + 1};
            int levels = 1;
            while ((1ull << levels) < (unsigned long long)2 * std::max(seed_tot[0], seed_tot[1]) + 2) levels++;
            // no room for the jump table (levels x nodes words): the seed tables still serve the host scan below
            const bool jump_ok = !getenv("GROMGPU_CNV_HOST_SCAN") && c.jump.ensure(sizeof(uint32_t) * (size_t)levels * n_nodes) && c.flags.ensure((size_t)n_nodes) &&
                                 c.hop_out.ensure(sizeof(HopCall) * (size_t)c.spec_cap) && c.hop_sink.ensure(2 * sizeof(HopSink));
            if (!jump_ok) cudaGetLastError();
            uint32_t *J = jump_ok ? c.jump.as<uint32_t>() : nullptr;
            uint8_t *flag = jump_ok ? c.flags.as<uint8_t>() : nullptr;
            const uint32_t most = std::max(seed_tot[0], seed_tot[1]);
            CK(cudaMemsetAsync(c.d_open, 0, sizeof(uint32_t) * 4 * (size_t)words, s));
            // positions with a z value (after the biased-repeat override, if any), ranked like the seed bitmaps
            uint32_t *d_nz = c.d_seed + 2 * words, *d_nzwp = c.d_wp + 2 * words;
            k_nz_bits<<<(unsigned)((words * 32 + 255) / 256), 256, 0, s>>>(c.d_rec, P, d_nz); n_launch++;
            k_seed_blocksum<<<dim3((unsigned)c.nb, 1), 256, 0, s>>>(d_nz, words, c.d_blk + 2 * c.nb, c.nb); n_launch++;
            k_seed_blockscan<<<1, 1024, 0, s>>>(c.d_blk + 2 * c.nb, c.nb, c.d_blk + 3 * c.nb + 2); n_launch++;
            k_seed_rank<<<dim3((unsigned)c.nb, 1), 256, 0, s>>>(d_nz, words, c.d_blk + 2 * c.nb, c.nb, d_nzwp); n_launch++;
            // the stretch-end check is a handful of warps walking 10,000 positions each (latency, not work): on a side stream, beside pass one
            CK(cudaEventRecord(c.ev_z, s));
            CK(cudaStreamWaitEvent(c.copy_stream, c.ev_z, 0));
            k_tail_ends<<<(unsigned)((words + 255) / 256), 256, 0, c.copy_stream>>>(d_nz, words, c.d_ends, CnvState::ENDS_CAP, c.d_nspec + 12); n_launch++;
            k_tail_check<<<148 * 8, 128, 0, c.copy_stream>>>(c.d_rec, c.d_z, P, c.d_ends, c.d_nspec + 12, CnvState::ENDS_CAP, Lmax, c.d_winsd + 2 * (Lmax + 1), c.d_u1 + 2 * words, words); n_launch++;
            CK(cudaEventRecord(c.ev_copied, c.copy_stream));
            if (most) {
                Grow &t_mid = c.mid;
                if (!t_mid.ensure(sizeof(SeedTodo) * (size_t)todo_cap)) return fail("gromgpu_chr_cnv: out of device memory");
                ctx[0].bound = ctx[1].bound = std::min(SEED_BOUND0, Lmin + 4);          // just past the first window: what survives it is classified, not walked further
                k_seed_eval<<<dim3((unsigned)((words + SEED_CTA_WORDS - 1) / SEED_CTA_WORDS), 2), 256, 0, s>>>(ctx[0], ctx[1], c.d_seed, words, c.d_wp, c.d_land, c.land_cap, seed_tot[0], seed_tot[1], c.d_spec, c.spec_cap,
                                                                                                                   c.d_nspec, t_mid.as<SeedTodo>(), todo_cap, J, c.d_u1); n_launch++;
                ctx[0].bound = ctx[1].bound = SEED_BOUND;
                Grow &t_wl = c.walk;
                if (!t_wl.ensure(sizeof(SeedTodo) * (size_t)todo_cap)) return fail("gromgpu_chr_cnv: out of device memory");
                CK(cudaStreamWaitEvent(s, c.ev_copied, 0));                       // safe stretch ends are in place
                k_seed_eval_mid<0><<<148 * 8, 128, 0, s>>>(ctx[0], ctx[1], c.d_seed, words, c.d_wp, c.d_land, c.land_cap, seed_tot[0], seed_tot[1], c.d_spec, c.spec_cap, c.d_nspec,
                                                           t_mid.as<SeedTodo>(), todo_cap, t_todo.as<SeedTodo>(), todo_cap, J, c.d_open, d_nz, d_nzwp, c.d_u1, c.d_u1 + 2 * words, c.d_u1 + 4 * words, t_wl.as<SeedTodo>()); n_launch++;
                k_seed_eval_mid<2><<<148 * 8, 128, 0, s>>>(ctx[0], ctx[1], c.d_seed, words, c.d_wp, c.d_land, c.land_cap, seed_tot[0], seed_tot[1], c.d_spec, c.spec_cap, c.d_nspec,
                                                           t_mid.as<SeedTodo>(), todo_cap, t_todo.as<SeedTodo>(), todo_cap, J, c.d_open, d_nz, d_nzwp, c.d_u1, c.d_u1 + 2 * words, c.d_u1 + 4 * words, t_wl.as<SeedTodo>()); n_launch++;
                k_seed_eval_mid<1><<<148 * 8, 128, 0, s>>>(ctx[0], ctx[1], c.d_seed, words, c.d_wp, c.d_land, c.land_cap, seed_tot[0], seed_tot[1], c.d_spec, c.spec_cap, c.d_nspec,
                                                           t_mid.as<SeedTodo>(), todo_cap, t_todo.as<SeedTodo>(), todo_cap, J, c.d_open, d_nz, d_nzwp, c.d_u1, c.d_u1 + 2 * words, c.d_u1 + 4 * words, t_wl.as<SeedTodo>()); n_launch++;
            }
            // Open seeds (ran past the first round's bound: genuine events and long stretches without coverage).  The heads of their runs are
            // evaluated exactly by the host, all in parallel, over windows of records fetched in one go (a long walk is a dependent chain:
            // ~1 us per position for a lone GPU thread, ~10 ns on a host core); then the open seeds that no call made at a head covers get
            // the full growth phase on the device, one thread each.  Whatever is still open after that is a sink of the jump table.
            constexpr uint32_t HEAD_CAP = 1u << 16;
            if (!c.heads.ensure(sizeof(SeedHead) * (size_t)HEAD_CAP) || !c.head_out.ensure(sizeof(HeadOutcome) * (size_t)HEAD_CAP)) return fail("gromgpu_chr_cnv: out of device memory");
            if (!c.h_heads) { CK(cudaMallocHost(&c.h_heads, sizeof(SeedHead) * (size_t)HEAD_CAP + 64)); CK(cudaMallocHost(&c.h_head_out, sizeof(HeadOutcome) * (size_t)HEAD_CAP)); }
            unsigned int *n_heads_d = c.d_nspec + 6;
            uint32_t *d_cover = c.d_open + 4 * words;                    // [2 kinds][words]
            CK(cudaMemsetAsync(d_cover, 0, sizeof(uint32_t) * 2 * (size_t)words, s));
            CK(cudaStreamWaitEvent(s, c.ev_copied, 0));                           // (also when there were no seeds at all)
            k_open_heads<<<592, 256, 0, s>>>(t_todo.as<SeedTodo>(), c.d_nspec + 1, todo_cap, c.d_open, words, c.heads.as<SeedHead>(), HEAD_CAP, n_heads_d); n_launch++;
            ctx[0].bound = ctx[1].bound = Lmax;                     // second round: the whole growth phase
            unsigned int *h_cnt = (unsigned int *)((char *)c.h_heads + sizeof(SeedHead) * (size_t)HEAD_CAP);       // [0] open seeds [1] heads
            CK(cudaMemcpyAsync(h_cnt, c.d_nspec + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(h_cnt + 1, n_heads_d, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(c.h_heads, c.heads.p, sizeof(SeedHead) * 4096, cudaMemcpyDeviceToHost, s));                 // the usual case in the same round trip
            CK(cudaStreamSynchronize(s));
            const unsigned int n_open = h_cnt[0];
            const unsigned int n_heads_all = std::min(h_cnt[1], HEAD_CAP);
            if (n_heads_all > 4096) { CK(cudaMemcpyAsync(c.h_heads, c.heads.p, sizeof(SeedHead) * (size_t)n_heads_all, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s)); }
            unsigned int n_heads = 0;                                   // heads the device round left open, compacted to the front
            for (unsigned int i = 0; i < n_heads_all; i++) if (!c.h_heads[i].pad) c.h_heads[n_heads++] = c.h_heads[i];
            mark("  seeds, first round + heads (device)");
            if (trace) for (unsigned int i = 0; i < std::min(n_heads, 6u); i++) fprintf(stderr, "[cnv]     head %u: kind %d class %d pos %d run end %d (%d positions)\n", i, c.h_heads[i].kind, c.h_heads[i].variant, c.h_heads[i].pos, c.h_heads[i].run_end, c.h_heads[i].run_end - c.h_heads[i].pos);
            device_hop = jump_ok;                              // GROMGPU_CNV_HOST_SCAN forces the host scan (tests)
            unsigned int n_head_done = 0;
            if (n_heads) {
                // windows: [head, end of its run + the sliding phase's look-ahead); overlapping windows share one copy
                const SeedHead *hd = c.h_heads;
                std::vector<uint32_t> order(n_heads);
                for (uint32_t i = 0; i < n_heads; i++) order[i] = i;
                std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return hd[x].pos < hd[y].pos; });
                const int64_t look = 2 * (int64_t)Lmax + 2048;
                struct Seg { int64_t a, b, off; };
                std::vector<Seg> segs; std::vector<uint32_t> seg_of(n_heads);
                int64_t total = 0;
                const int64_t budget = (int64_t)64 << 20;                    // records (256 MB) fetched for heads at most; the rest stays open
                uint32_t n_take = 0;
                for (uint32_t oi = 0; oi < n_heads; oi++) {
                    const SeedHead &h0 = hd[order[oi]];
                    const int64_t a0 = h0.pos, b0 = std::min<int64_t>(P, (int64_t)h0.run_end + look);
                    if (!segs.empty() && a0 <= segs.back().b) { const int64_t grow = std::max<int64_t>(0, b0 - segs.back().b); if (total + grow > budget) break; segs.back().b += grow; total += grow; }
                    else { if (total + (b0 - a0) > budget) break; segs.push_back({a0, b0, total}); total += b0 - a0; }
                    seg_of[oi] = (uint32_t)segs.size() - 1; n_take = oi + 1;
                }
                if ((size_t)total * 4 > c.h_win_cap) {
                    if (c.h_win) cudaFreeHost(c.h_win);
                    c.h_win = nullptr; c.h_win_cap = 0;
                    CK(cudaMallocHost(&c.h_win, (size_t)total * 5));
                    c.h_win_cap = (size_t)total * 5;
                }
                for (const Seg &g : segs) CK(cudaMemcpyAsync(c.h_win + g.off, c.d_rec + g.a, sizeof(uint32_t) * (size_t)(g.b - g.a), cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
                d2h += 4 * total + (int64_t)sizeof(SeedHead) * n_heads_all;
                mark("  head windows D2H");
                HeadOutcome *ho = c.h_head_out;
                std::vector<uint8_t> ok(n_take, 0);
                auto eval_one = [&](uint32_t oi, int64_t *call_end) {
                    const SeedHead &h0 = hd[order[oi]];
                    const Seg &g = segs[seg_of[oi]];
                    SegCtx hc = ctx[h0.kind];
                    hc.rec = c.h_win + g.off - g.a; hc.len = g.b; hc.sd = c.sd_tbl.data(); hc.win_sd = c.win_sd.data(); hc.win_thr = c.win_thr.data(); hc.wtab = c.wtab.data(); hc.zarr = nullptr;
                    const int c0 = hc.cls(h0.pos);
                    const Outcome o = eval_seed<false>(hc, h0.pos, c0 != 2 ? c0 : h0.variant);
                    if (g.b < P && o.far >= g.b) return;                  // the sliding phase ran into the window's end: left open (the path evaluates it if it gets there)
                    HeadOutcome &r = ho[oi];
                    r.rank = h0.rank; r.kind = h0.kind; r.variant = h0.variant; r.seg = (uint16_t)o.kind; r.pos = h0.pos;
                    r.rel_next = (int32_t)(o.next - h0.pos); r.c_end = o.c_end; r.c_z = o.c_z;
                    ok[oi] = 1;
                    if (o.kind == SEG_CALL) *call_end = o.c_end;
                };
                // Heads come in clusters (the fragments of one event); the path enters a cluster at its first head and the call made there
                // usually jumps over most of the others.  So a cluster is walked in position order by ONE thread that skips the heads under
                // the calls it has made so far, and the clusters are spread over the threads.  Skipped heads stay open (and end up under the
                // cover bitmap); should the path reach one after all, it is evaluated then.
                std::vector<uint32_t> cl_first;                               // index (in position order) of the first head of every cluster
                for (uint32_t oi = 0; oi < n_take; oi++)
                    if (oi == 0 || hd[order[oi]].pos - hd[order[oi - 1]].pos > 2 * (int64_t)Lmax) cl_first.push_back(oi);
                cl_first.push_back(n_take);
                const uint32_t n_cl = (uint32_t)cl_first.size() - 1;
                std::atomic<uint32_t> next_cl(0);
                auto cluster_worker = [&]() {
                    for (;;) {
                        const uint32_t k = next_cl.fetch_add(1);
                        if (k >= n_cl) return;
                        int64_t under[2] = {-1, -1};                          // end of the calls made so far, per kind
                        for (uint32_t oi = cl_first[k]; oi < cl_first[k + 1]; oi++) {
                            const SeedHead &h0 = hd[order[oi]];
                            if (h0.pos <= under[h0.kind]) continue;
                            int64_t ce = -1;
                            eval_one(oi, &ce);
                            if (ce + 1 > under[h0.kind]) under[h0.kind] = ce + 1;
                        }
                    }
                };
                {
                    const unsigned T = std::max(1u, std::min<unsigned>(cnv_host_threads(16), n_cl));
                    std::vector<std::thread> pool;
                    for (unsigned t = 1; t < T; t++) pool.emplace_back(cluster_worker);
                    cluster_worker();
                    for (auto &x : pool) x.join();
                }
                for (uint32_t oi = 0; oi < n_take; oi++) if (ok[oi]) { if (n_head_done != oi) ho[n_head_done] = ho[oi]; n_head_done++; }
                mark("  heads evaluated (host)");
                if (n_head_done) {
                    CK(cudaMemcpyAsync(c.head_out.p, ho, sizeof(HeadOutcome) * (size_t)n_head_done, cudaMemcpyHostToDevice, s));
                    k_apply_heads<<<(n_head_done * 32 + 127) / 128, 128, 0, s>>>(ctx[0], ctx[1], c.head_out.as<HeadOutcome>(), n_head_done, c.d_seed, words, c.d_wp, c.d_land, c.land_cap, c.d_spec, c.spec_cap,
                                                                                 c.d_nspec, seed_tot[0], seed_tot[1], J, d_cover); n_launch++;
                    // early gather of the calls made here (at most 32 M positions), copy numbers beside the device work
                    spec_cn.first.push_back(0);
                    for (uint32_t i = 0; i < n_head_done; i++) if (ho[i].seg == SEG_CALL && ho[i].c_end > ho[i].pos) {
                        spec_cn.kind.push_back(ho[i].kind); spec_cn.start.push_back(ho[i].pos); spec_cn.end.push_back(ho[i].c_end);
                        spec_cn.first.push_back(spec_cn.first.back() + (ho[i].c_end - ho[i].pos));
                    }
                    const int n_sp = (int)spec_cn.start.size();
                    const int64_t tot_sp = spec_cn.first.back();
                    bool sp_ok = n_sp > 0 && tot_sp <= ((int64_t)32 << 20) && !getenv("GROMGPU_CNV_NO_EARLY_CN");
                    if (sp_ok) {
                        const size_t need = (size_t)tot_sp * 9 + 64;
                        if (need > c.h_specg_cap) {
                            if (c.h_specg) cudaFreeHost(c.h_specg);
                            c.h_specg = nullptr; c.h_specg_cap = 0;
                            if (cudaMallocHost(&c.h_specg, need + need / 4) == cudaSuccess) c.h_specg_cap = need + need / 4; else { cudaGetLastError(); sp_ok = false; }
                        }
                        if (!c.ev_spec && cudaEventCreateWithFlags(&c.ev_spec, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); sp_ok = false; }
                        sp_ok = sp_ok && c.spec_g[0].ensure(sizeof(int64_t) * n_sp) && c.spec_g[1].ensure(sizeof(int64_t) * (n_sp + 1)) && c.spec_g[2].ensure(sizeof(int32_t) * (size_t)tot_sp) &&
                                c.spec_g[3].ensure((size_t)tot_sp) && c.spec_g[4].ensure(sizeof(uint32_t) * (size_t)tot_sp);
                    }
                    if (sp_ok) {
                        CK(cudaMemcpyAsync(c.spec_g[0].p, spec_cn.start.data(), sizeof(int64_t) * n_sp, cudaMemcpyHostToDevice, s));
                        CK(cudaMemcpyAsync(c.spec_g[1].p, spec_cn.first.data(), sizeof(int64_t) * (n_sp + 1), cudaMemcpyHostToDevice, s));
                        k_gather<<<(unsigned)((tot_sp + 255) / 256), 256, 0, s>>>(c.d_depth, A_gc, A_acgt, c.spec_g[0].as<int64_t>(), c.spec_g[1].as<int64_t>(), n_sp, tot_sp, c.spec_g[2].as<int32_t>(),
                                                                                  c.spec_g[3].as<uint8_t>(), c.d_rec, c.spec_g[4].as<uint32_t>()); n_launch++;
                        int32_t *pd = (int32_t *)c.h_specg; uint32_t *pr = (uint32_t *)(pd + tot_sp); uint8_t *pg = (uint8_t *)(pr + tot_sp);
                        CK(cudaMemcpyAsync(pd, c.spec_g[2].p, sizeof(int32_t) * (size_t)tot_sp, cudaMemcpyDeviceToHost, s));
                        CK(cudaMemcpyAsync(pr, c.spec_g[4].p, sizeof(uint32_t) * (size_t)tot_sp, cudaMemcpyDeviceToHost, s));
                        CK(cudaMemcpyAsync(pg, c.spec_g[3].p, (size_t)tot_sp, cudaMemcpyDeviceToHost, s));
                        CK(cudaEventRecord(c.ev_spec, s));
                        d2h += 9 * tot_sp;
                        spec_cn.cn.assign(n_sp, -1.0); spec_cn.cn_sd.assign(n_sp, 0.0);
                        const int dev_id = g_device;
                        cudaEvent_t ev = c.ev_spec;
                        spec_cn.th = std::thread([&, pd, pr, pg, n_sp, dev_id, ev]() {
                            cudaSetDevice(dev_id);
                            if (cudaEventSynchronize(ev) != cudaSuccess) { for (int i = 0; i < n_sp; i++) spec_cn.cn_sd[i] = -1.0; return; }      // cn_sd < 0: not computed
                            std::vector<uint32_t> order(n_sp);
                            for (int i = 0; i < n_sp; i++) order[i] = (uint32_t)i;
                            std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return spec_cn.first[a + 1] - spec_cn.first[a] > spec_cn.first[b + 1] - spec_cn.first[b]; });
                            std::atomic<int> next(0);
                            auto w = [&]() {
                                std::vector<double> buf, tmp;
                                for (;;) {
                                    const int oi = next.fetch_add(1);
                                    if (oi >= n_sp) break;
                                    const int i = (int)order[oi];
                                    const int64_t f0 = spec_cn.first[i], np = spec_cn.first[i + 1] - f0;
                                    copy_number(pd + f0, pr + f0, pg + f0, np, buf, tmp, &spec_cn.cn[i], &spec_cn.cn_sd[i]);
                                }
                            };
                            const unsigned T = std::max(1u, std::min<unsigned>(cnv_host_threads(16) / 2, (unsigned)n_sp));
                            std::vector<std::thread> pool;
                            for (unsigned t = 1; t < T; t++) pool.emplace_back(w);
                            w();
                            for (auto &x : pool) x.join();
                        });
                    } else { spec_cn.start.clear(); spec_cn.end.clear(); spec_cn.kind.clear(); }
                }
            }
            if (n_open) {
                const uint32_t n_todo = std::min(n_open, todo_cap);
                k_seed_filter2<<<(n_todo + 255) / 256, 256, 0, s>>>(c.d_land, c.land_cap, c.d_nspec, t_todo.as<SeedTodo>(), todo_cap, d_cover, words, c.mid.as<SeedTodo>()); n_launch++;
                k_seed_eval2<<<(n_todo + 63) / 64, 64, 0, s>>>(ctx[0], ctx[1], c.d_seed, words, c.d_wp, c.d_land, c.land_cap, c.d_spec, c.spec_cap, c.d_nspec, c.mid.as<SeedTodo>(),
                                                               seed_tot[0], seed_tot[1], J); n_launch++;
                if (trace) {
                    unsigned int dbg[16];
                    CK(cudaMemcpyAsync(dbg, c.d_nspec, sizeof(dbg), cudaMemcpyDeviceToHost, s));
                    CK(cudaStreamSynchronize(s));
                    mark("  seeds, second round");
                    fprintf(stderr, "[cnv]   %u open seeds after the first round (%u zero-stretch seeds closed in O(1) before that); %u run heads, %u left to the host; second round: %u closed, %u skipped under calls made at heads\n",
                            n_open, dbg[7], n_heads_all, n_heads, dbg[4], dbg[5]);
                    fprintf(stderr, "[cnv]   %u seeds past the first bound; %u closed as tails of stretches without z (%u stretch ends); %u left open behind an open run end\n", dbg[8], dbg[10], dbg[12], dbg[11]);
                    if (const char *dump = getenv("GROMGPU_CNV_DUMP")) {
                        // development aid: the open-seed list with each seed's final table entry and cover bit
                        std::vector<SeedTodo> td(n_todo); std::vector<uint32_t> ld(4 * (size_t)c.land_cap), cov(2 * (size_t)words);
                        CK(cudaMemcpy(td.data(), t_todo.p, sizeof(SeedTodo) * n_todo, cudaMemcpyDeviceToHost));
                        CK(cudaMemcpy(ld.data(), c.d_land, sizeof(uint32_t) * ld.size(), cudaMemcpyDeviceToHost));
                        CK(cudaMemcpy(cov.data(), d_cover, sizeof(uint32_t) * cov.size(), cudaMemcpyDeviceToHost));
                        std::vector<SeedCall> sc(std::min(dbg[0], c.spec_cap));
                        if (!sc.empty()) CK(cudaMemcpy(sc.data(), c.d_spec, sizeof(SeedCall) * sc.size(), cudaMemcpyDeviceToHost));
                        {
                            // ... and every seed that ran past the first bound, with its final table entry
                            const unsigned int n_mid = std::min(dbg[8], todo_cap);
                            std::vector<SeedTodo> md(n_mid);
                            if (n_mid) CK(cudaMemcpy(md.data(), c.mid.p, sizeof(SeedTodo) * n_mid, cudaMemcpyDeviceToHost));
                            std::string name = std::string(dump) + ".mid";
                            if (FILE *f = fopen(name.c_str(), "wb")) {
                                for (size_t mi = 0; mi < md.size(); mi += 8) {                       // a sample: every 8th
                                    const SeedTodo &t = md[mi];
                                    const uint32_t e = ld[((size_t)t.kind * c.land_cap + t.rank) * 2 + t.variant];
                                    const int32_t row[4] = {t.pos, t.kind * 2 + t.variant, (int32_t)e, t.pad};
                                    fwrite(row, sizeof(row), 1, f);
                                }
                                fclose(f);
                            }
                        }
                        if (FILE *f = fopen(dump, "wb")) {
                            for (const SeedTodo &t : td) {
                                const uint32_t e = ld[((size_t)t.kind * c.land_cap + t.rank) * 2 + t.variant];
                                const int64_t cend = (e != LAND_NOT && (e >> LAND_SHIFT) == SEG_CALL) ? sc[e & ((1u << LAND_SHIFT) - 1u)].c_end : -1;
                                const int32_t row[6] = {t.pos, t.kind, t.variant, (int32_t)e, (int32_t)((cov[(size_t)t.kind * words + (t.pos >> 5)] >> (t.pos & 31)) & 1u), (int32_t)cend};
                                fwrite(row, sizeof(row), 1, f);
                            }
                            fclose(f);
                        }
                    }
                }
            }
            if (device_hop) {
            if (most == 0) CK(cudaMemsetAsync(J, 0, sizeof(uint32_t) * n_nodes, s));
            if (seed_tot[0] == 0 || seed_tot[1] == 0) {                // a scan without seeds: its END node loops on itself
                const uint32_t e0 = base[0] + 2 * seed_tot[0], e1 = base[1] + 2 * seed_tot[1];
                if (seed_tot[0] == 0) CK(cudaMemcpyAsync(J + e0, &e0, 4, cudaMemcpyHostToDevice, s));
                if (seed_tot[1] == 0) CK(cudaMemcpyAsync(J + e1, &e1, 4, cudaMemcpyHostToDevice, s));
            }
            for (int k = 1; k < levels; k++) { k_hop_double<<<(n_nodes + 255) / 256, 256, 0, s>>>(J + (size_t)(k - 1) * n_nodes, J + (size_t)k * n_nodes, n_nodes); n_launch++; }
            CK(cudaMemsetAsync(flag, 0, (size_t)n_nodes, s));
            unsigned int *n_hop = c.d_nspec + 2;                        // [2]: calls collected per scan
            // The path is found leg by leg: open seeds are sinks of the jump table, so its top level leads from a leg's start straight to
            // the first open seed on the path; that seed is evaluated here on a window of records and the next leg starts where the
            // evaluation lands.  One tiny launch and one round trip per leg.
            std::vector<Call> by_host[2];
            std::vector<uint32_t> window;
            HopLeg leg[2] = {{lo, 0, 1}, {lo, 0, 1}};
            int64_t n_legs = 0;
            int no_call = 0; bool too_many_legs = false;
            while (leg[0].active || leg[1].active) {
                if (++n_legs > 512) { too_many_legs = true; break; }      // open seeds keep turning up on the path: the host scan below takes over
                k_hop_advance<<<1, 2, 0, s>>>(ctx[0], ctx[1], c.d_seed, c.d_wp, words, J + (size_t)(levels - 1) * n_nodes, leg[0], leg[1], seed_tot[0], seed_tot[1], flag,
                                              c.hop_sink.as<HopSink>()); n_launch++;
                HopSink sink[2];
                CK(cudaMemcpyAsync(sink, c.hop_sink.p, sizeof(sink), cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
                for (int k = 0; k < 2; k++) if (leg[k].active) {
                    if (!sink[k].found) { leg[k].active = 0; continue; }
                    const int64_t pos = sink[k].pos;
                    Outcome o;
                    int c0 = 0;
                    for (int64_t span = std::max<int64_t>(3 * (int64_t)Lmax + 2048, 1 << 17);; span *= 2) {
                        const int64_t w1 = std::min<int64_t>(P, pos + span);
                        window.resize(w1 - pos);
                        CK(cudaMemcpyAsync(window.data(), c.d_rec + pos, sizeof(uint32_t) * (w1 - pos), cudaMemcpyDeviceToHost, s));
                        CK(cudaStreamSynchronize(s));
                        d2h += 4 * (w1 - pos);
                        SegCtx hc = ctx[k];
                        hc.rec = window.data() - pos; hc.len = w1; hc.zarr = nullptr; hc.sd = c.sd_tbl.data(); hc.win_sd = c.win_sd.data(); hc.win_thr = c.win_thr.data(); hc.wtab = c.wtab.data();
                        c0 = hc.cls(pos);
                        o = eval_seed<false>(hc, pos, c0 != 2 ? c0 : sink[k].variant);
                        if (w1 == P || o.far < w1) break;
                    }
                    if (o.kind == SEG_CALL) by_host[k].push_back({pos, o.c_end, o.c_z}); else no_call++;
                    leg[k].x = o.next; leg[k].s = c0 != 2 ? c0 : sink[k].variant;
                }
            }
            mark("  jump table + legs");
            if (too_many_legs) {
                dev_end(); device_hop = false;
                if (trace) fprintf(stderr, "[cnv] more than 512 open seeds on the path: leaving the hop to the host scan\n");
            } else {
            for (int k = levels - 1; k >= 0; k--) { k_hop_mark<<<(n_nodes + 255) / 256, 256, 0, s>>>(J + (size_t)k * n_nodes, flag, n_nodes); n_launch++; }
            for (int k = 0; k < 2; k++) if (seed_tot[k]) {
                k_hop_collect<<<(2 * seed_tot[k] + 255) / 256, 256, 0, s>>>(flag + base[k], c.d_land + (size_t)k * 2 * c.land_cap, c.d_seed + k * words, c.d_wp + k * words, words,
                                                                             seed_tot[k], c.d_spec, c.hop_out.as<HopCall>() + (size_t)k * (c.spec_cap / 2), c.spec_cap / 2, n_hop + k); n_launch++;
            }
            unsigned int n_got[2] = {0, 0};
            CK(cudaMemcpyAsync(n_got, n_hop, sizeof(n_got), cudaMemcpyDeviceToHost, s));
            dev_end();
            CK(cudaGetLastError());
            for (int k = 0; k < 2; k++) {
                if (n_got[k] > c.spec_cap / 2) return fail("gromgpu_chr_cnv: %u calls exceed the buffer", n_got[k]);
                std::vector<HopCall> got(n_got[k]);
                if (n_got[k]) CK(cudaMemcpy(got.data(), c.hop_out.as<HopCall>() + (size_t)k * (c.spec_cap / 2), sizeof(HopCall) * n_got[k], cudaMemcpyDeviceToHost));
                d2h += (int64_t)sizeof(HopCall) * n_got[k];
                for (const HopCall &g : got) found[k].push_back({g.pos, g.c_end, g.c_z});
                for (const Call &g : by_host[k]) found[k].push_back(g);
                std::sort(found[k].begin(), found[k].end(), [](const Call &a, const Call &b) { return a.start < b.start; });
            }
            seed_tot_all = 0; n_spec_all = 0;
            if (trace) fprintf(stderr, "[cnv] seeds del %u dup %u; %u open after the first round, %u run heads, %u evaluated by the host up front; jump table %d levels x %u nodes, %lld legs (%d sinks met on the path without a call); calls from the device %u + %u, from seeds met on the path %zu + %zu\n",
                               seed_tot[0], seed_tot[1], n_open, n_heads_all, n_head_done, levels, n_nodes, (long long)n_legs, no_call, n_got[0], n_got[1], by_host[0].size(), by_host[1].size());
            }
            } else { dev_end(); if (trace) fprintf(stderr, "[cnv] no jump table: scanning on the host\n"); }
        }
        if (!have_land || !device_hop) {
            // host scan over the packed records, in parallel pieces; with the device-evaluated seed tables when they exist (the hop is
            // then a table walk and only unresolved seeds are evaluated here), without them every seed is evaluated on the host
            if (pull_records()) return -1;
            unsigned int n_spec_host = 0;
            if (have_land) {
                if (!c.h_land) { CK(cudaMallocHost(&c.h_land, sizeof(uint32_t) * 4 * (size_t)c.land_cap)); CK(cudaMallocHost(&c.h_spec, sizeof(SeedCall) * (size_t)c.spec_cap)); }
                for (int k = 0; k < 2; k++) if (seed_tot[k]) CK(cudaMemcpyAsync(c.h_land + (size_t)k * 2 * c.land_cap, c.d_land + (size_t)k * 2 * c.land_cap, sizeof(uint32_t) * 2 * seed_tot[k], cudaMemcpyDeviceToHost, s));
                CK(cudaMemcpyAsync(&n_spec_host, c.d_nspec, sizeof(n_spec_host), cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
                n_spec_host = std::min(n_spec_host, c.spec_cap);
                if (n_spec_host) CK(cudaMemcpy(c.h_spec, c.d_spec, sizeof(SeedCall) * n_spec_host, cudaMemcpyDeviceToHost));
                d2h += 8 * ((int64_t)seed_tot[0] + seed_tot[1]) + 16 * (int64_t)n_spec_host;
            }
            Segmenter sg[2];
            for (int k = 0; k < 2; k++) {
                sg[k].C = ctx[k]; sg[k].C.zarr = nullptr; sg[k].C.rec = c.h_rec; sg[k].C.sd = c.sd_tbl.data(); sg[k].C.win_sd = c.win_sd.data(); sg[k].C.win_thr = c.win_thr.data(); sg[k].C.wtab = c.wtab.data(); sg[k].seeds = c.h_seed + k * words; sg[k].lo = lo;
                if (have_land) { sg[k].wp = c.h_wp + k * words; sg[k].land = c.h_land + (size_t)k * 2 * c.land_cap; sg[k].spec = c.h_spec; }
            }
            const int hw = (int)cnv_host_threads(16), per_scan = std::max(1, std::min(8, hw / 2));
            std::thread th([&]() { sg[1].run(found[1], per_scan); });
            sg[0].run(found[0], per_scan);
            th.join();
        }
    }
    mark("segmentation host");
    std::vector<int64_t> seg_start, seg_first(1, 0);
    for (int k = 0; k < 2; k++) for (const Call &cl : found[k]) { seg_start.push_back(cl.start); seg_first.push_back(seg_first.back() + std::max<int64_t>(0, cl.end - cl.start)); }
    std::vector<int32_t> g_depth; std::vector<uint8_t> g_gc; std::vector<uint32_t> g_rec;
    void *g_pin[3] = {nullptr, nullptr, nullptr};
    if (!seg_start.empty() && gather(seg_start, seg_first, g_depth, g_gc, &g_rec, g_pin)) return -1;
    const int32_t *gp_depth = (const int32_t *)g_pin[0]; const uint32_t *gp_rec = (const uint32_t *)g_pin[1]; const uint8_t *gp_gc = (const uint8_t *)g_pin[2];
    const int64_t g_total = seg_first.back();
    mark("gather");
    {
        // copy number per call (src/GROM.c:20071-20224): independent per call, spread over a few host threads
        std::vector<const Call *> flat; std::vector<int> flat_kind;
        for (int k = 0; k < 2; k++) for (const Call &cl : found[k]) { flat.push_back(&cl); flat_kind.push_back(k); }
        c.calls.assign(flat.size(), grom_cnv_call());
        // calls are handed out longest first (their cost is n log n in the call's length and a few calls hold most positions)
        std::vector<uint32_t> by_len(flat.size());
        for (size_t i = 0; i < by_len.size(); i++) by_len[i] = (uint32_t)i;
        std::sort(by_len.begin(), by_len.end(), [&](uint32_t a, uint32_t b) { const int64_t la = seg_first[a + 1] - seg_first[a], lb = seg_first[b + 1] - seg_first[b]; return la != lb ? la > lb : a < b; });
        std::atomic<size_t> next_call(0);
        // what the background thread has ready: (kind, start, end) -> index
        if (spec_cn.th.joinable()) spec_cn.th.join();
        std::map<std::tuple<int, int64_t, int64_t>, int> early;
        for (size_t i = 0; i < spec_cn.cn.size(); i++) if (spec_cn.cn_sd[i] >= 0) early[std::make_tuple(spec_cn.kind[i], spec_cn.start[i], spec_cn.end[i])] = (int)i;
        auto work = [&]() {
            std::vector<double> buf, tmp;
            for (;;) {
                const size_t oi = next_call.fetch_add(1);
                if (oi >= by_len.size()) break;
                const size_t si = by_len[oi];
                const Call &cl = *flat[si];
                grom_cnv_call o; o.start = cl.start; o.end = cl.end; o.kind = flat_kind[si]; o.reserved = 0; o.z = cl.z; o.cn = -1; o.cn_sd = 0;
                const auto hit = early.find(std::make_tuple(flat_kind[si], (int64_t)cl.start, (int64_t)cl.end));
                if (hit != early.end()) { o.cn = spec_cn.cn[hit->second]; o.cn_sd = spec_cn.cn_sd[hit->second]; }     // computed beside the device work
                else copy_number(gp_depth + seg_first[si], gp_rec + seg_first[si], gp_gc + seg_first[si], seg_first[si + 1] - seg_first[si], buf, tmp, &o.cn, &o.cn_sd);
                // one-sided normal tail through the reference's own erf variant, t = 1 / (1 + p + x) (src/GROM.c:17163-17172)
                const double x = fabs(o.z) / sqrt(2.0), t = 1.0 / (1.0 + 0.3275911 + x);
                const double erf_ = 1.0 - ((0.254829592 * t + -0.284496736 * (t * t) + 1.421413741 * pow(t, 3) + -1.453152027 * pow(t, 4) + 1.061405429 * pow(t, 5)) * exp(-(x * x)));
                o.pvalue = (1.0 - erf_) / 2.0;
                c.calls[si] = o;
            }
        };
        const size_t nc = flat.size(), T = std::max<size_t>(1, std::min<size_t>(cnv_host_threads(16), nc / 64 + 1));
        std::vector<std::thread> pool;
        for (size_t t = 1; t < T; t++) pool.emplace_back(work);
        work();
        for (auto &x : pool) x.join();
    }
    mark("copy number");
    out->n_calls = (int64_t)c.calls.size(); out->calls = c.calls.data();
    const double ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    d2h += (int64_t)sizeof(cnv::PreOut) * n_blk + 8 * HIST_ALL + (int64_t)sizeof(RepRec) * n_rep + (int64_t)sizeof(Sample) * n_samples +
           8 * ((int64_t)seed_tot_all) + 16 * (int64_t)n_spec_all + 9 * g_total + 5 * (int64_t)rp_depth.size() + 16 * (int64_t)n_len;
    out->launches = n_launch; out->d2h_bytes = d2h;
    out->ms_device = (float)ms_dev; out->ms_total = (float)ms_total; out->ms_host = (float)(ms_total - ms_dev);
    return 0;
}

__global__ void k_cnv_decode(const uint32_t *__restrict__ rec, const double *__restrict__ sd, int q, int what, int64_t p0, int64_t n, double *__restrict__ oz, uint8_t *__restrict__ ob)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = rec[p0 + i];
    if (what == 0) oz[i] = cnv::rec_z(r, q, sd); else ob[i] = (uint8_t)(r & cnv::R_MASK);
}

extern "C" int gromgpu_cnv_fetch(gromgpu_chr *h, int what, void *dst, int64_t p0, int64_t p1)
{
    if (!h || !h->cnv || !h->cnv->d_rec) return fail("gromgpu_cnv_fetch: call gromgpu_chr_cnv first");
    ON_DEV();
    if (p0 < 0 || p1 > h->P || p0 > p1) return fail("gromgpu_cnv_fetch: bad range");
    CnvState &c = *h->cnv;
    const int64_t n = p1 - p0;
    if (!n) return 0;
    if (what == 2) { CK(cudaMemcpy(dst, c.d_mq8 + p0, n, cudaMemcpyDeviceToHost)); return 0; }
    if (what == 3) { CK(cudaMemcpy(dst, c.d_depth + p0, sizeof(int32_t) * n, cudaMemcpyDeviceToHost)); return 0; }
    if (what != 0 && what != 1) return fail("gromgpu_cnv_fetch: unknown selector %d", what);
    cnv::DevTmpRaw o, sd;
    CK(cudaMalloc(&o.p, what == 0 ? sizeof(double) * n : (size_t)n)); CK(cudaMalloc(&sd.p, sizeof(double) * cnv::P2S));
    CK(cudaMemcpy(sd.p, c.sd_tbl.data(), sizeof(double) * cnv::P2S, cudaMemcpyHostToDevice));
    k_cnv_decode<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(c.d_rec, (const double *)sd.p, c.q, what, p0, n, (double *)o.p, (uint8_t *)o.p);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(dst, o.p, what == 0 ? sizeof(double) * n : (size_t)n, cudaMemcpyDeviceToHost));
    return 0;
}
