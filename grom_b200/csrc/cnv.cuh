// cnv.cuh -- read-depth CNV path of the reference (SURVEY.md §8 rows a13 side lists, a14, a15, a17) on the device.
//
// Reference: src/GROM.c:1684-1764 (dinucleotide-repeat runs), 16633-16990 (pre-statistics), 18228-20355 (detect_del_dup),
// 17146-17240 (p-value + -V filter), 21630-21860 (bisections).
//
// Split of the work (DESIGN.md "CNV path"):
//   device, one thread per reference position or better: mean MAPQ, depth, 10 kb block sums, depth histogram, repeat runs,
//           depth samples, mask, rank -> sd transform (cumulative-count tables instead of per-base binary searches),
//           per-frame window sweep for every window length (sequential inside a frame = the reference's summation order),
//           ordered reduction over frames, seed bitmaps for the segmentation
//   host,   O(samples) / O(calls) sequential logic with libc-rand / qsort semantics: reservoir lists, block clustering,
//           greedy DEL / DUP segmentation over the packed per-position records, copy number
//
// Every double that reaches the output is produced by the same operation sequence as the reference (no FMA contraction:
// __dmul_rn / __dadd_rn where a product feeds a sum).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace cnv {

constexpr int NB = 101;                 // g_num_gc_bins
constexpr int NLIST = 2 * NB;           // [mq class][gc bin]
constexpr int BLK_UNIT = 10000;         // g_block_unit_size
constexpr int CTILE = 2048;              // positions per carry tile
constexpr int HIST = 4096;              // depth histogram bins kept on chip
constexpr int HIST_ALL = 65536;         // bins of the global histogram (last bin collects everything deeper)
constexpr int MIN_ACGT = 99;            // g_insert_min_acgt
constexpr int NO_COMBINE = 100;         // g_rd_no_combine_min_windows
constexpr int MIN_WINDOWS = 20;         // g_rd_min_windows
constexpr int RD_MAX_MAPQ = 60;         // g_rd_max_mapq
constexpr int P2S = 1001;               // p-value -> sd table length

// packed per-position record (uint32)
constexpr uint32_t R_MASK = 1u;                          // rd_low_acgt_or_windows_list
constexpr int      R_CLASS = 1;                          // 2 bits: 0 mean MAPQ >= q, 1 covered with low MAPQ, 2 uncovered
constexpr uint32_t R_DEL0 = 1u << 3, R_DEL1 = 1u << 4;   // depth <= deletion threshold of the high / low list of the bin
constexpr uint32_t R_DUP0 = 1u << 5, R_DUP1 = 1u << 6;   // depth >= duplication threshold
constexpr uint32_t R_WIN0 = 1u << 7, R_WIN1 = 1u << 8;   // list has more than one sample
constexpr uint32_t R_NEG = 1u << 9, R_NZ = 1u << 10, R_OVR = 1u << 11;
constexpr int      R_K = 12, R_MQ = 22;                  // 10 bits table index, 8 bits mean MAPQ
constexpr uint32_t R_USABLE = 1u << 30;                  // mask == 0 and the list of the position's own MAPQ class has > 1 sample (the z kernel's carry walk reads it)

struct Tables {                 // device pointers
    const int32_t *cum;         // [NLIST][D+1]  number of samples <= d
    int32_t D;
    const int32_t *n;           // [NLIST]
    const int32_t *small;       // [NLIST][2]    the samples themselves when n < 3 (the reference's bisection is not a true bisection there)
    const double *ave, *del_thr, *dup_thr;   // [NLIST]
    const double *p2s_p, *p2s_sd;            // [P2S]
};

__host__ __device__ inline double rec_z(uint32_t r, int q, const double *sd)
{
    if (!(r & R_NZ)) return 0.0;
    const int mq = (r >> R_MQ) & 255, k = (r >> R_K) & 1023;
    double w;
    if (r & R_OVR) w = 1.0;
    else if (((r >> R_CLASS) & 3) == 0) w = 0.5 + (1.0 - 0.5) * (mq - q) / (double)(RD_MAX_MAPQ - q);
    else w = 0.5;
    const double z = w * sd[k];
    return (r & R_NEG) ? -z : z;
}
__host__ __device__ inline bool rec_usable(uint32_t r)
{
    return !(r & R_MASK) && ((((r >> R_CLASS) & 3) == 0) ? (r & R_WIN0) != 0 : (r & R_WIN1) != 0);
}

// ---- K1: mean MAPQ, depth, per-10kb block sums, contig sums, depth histogram (src/GROM.c:16637-16686, 16812-16831)
struct PreOut { unsigned long long blk_sum, acgt_sum, acgt_cnt, ave_sum, ave_cnt; };
__global__ void __launch_bounds__(256) k_pre(const int32_t *__restrict__ mqsum, const int32_t *__restrict__ rd, const int32_t *__restrict__ low,
                                             const int32_t *__restrict__ acgt, const char *__restrict__ fasta, int64_t P, int64_t lo, int64_t hi,
                                             int32_t *__restrict__ depth, uint8_t *__restrict__ mq8, PreOut *__restrict__ out,
                                             unsigned long long *__restrict__ hist)
{
    __shared__ unsigned int sh[HIST];
    __shared__ unsigned long long red[5];
    for (int i = threadIdx.x; i < HIST; i += blockDim.x) sh[i] = 0;
    if (threadIdx.x < 5) red[threadIdx.x] = 0;
    __syncthreads();
    const int64_t p0 = (int64_t)blockIdx.x * BLK_UNIT, p1 = min(p0 + BLK_UNIT, P);
    unsigned long long s_blk = 0, s_acgt = 0, n_acgt = 0, s_ave = 0, n_ave = 0;
    // the trip count is warp-uniform (BLK_UNIT rounded up to the block size) so that the warp-aggregated histogram update below can
    // use full-mask collectives
    for (int64_t b = p0; b < p1; b += blockDim.x) {
        const int64_t p = b + threadIdx.x;
        int hbin = -1;
        if (p < p1) {
            const int d = rd[p] + low[p];
            const int m = d > 0 ? mqsum[p] / d : mqsum[p];
            depth[p] = d; mq8[p] = (uint8_t)min(max(m, 0), 255);
            s_blk += (unsigned)d;
            const char c = fasta[p] & 0xDF;
            if (c == 'A' || c == 'C' || c == 'G' || c == 'T') { s_acgt += (unsigned)d; n_acgt++; }
            if (p >= lo && p < hi && acgt[p] >= MIN_ACGT) { s_ave += (unsigned)d; n_ave++; hbin = min(d, HIST_ALL - 1); }
        }
        // neighbouring positions mostly share their depth: one atomic per distinct value in the warp
        const unsigned peers = __match_any_sync(0xffffffffu, hbin);
        if (hbin >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) {
            if (hbin < HIST) atomicAdd(&sh[hbin], (unsigned)__popc(peers)); else atomicAdd(&hist[hbin], (unsigned long long)__popc(peers));
        }
    }
    for (int o = 16; o; o >>= 1) {
        s_blk += __shfl_xor_sync(0xffffffffu, s_blk, o); s_acgt += __shfl_xor_sync(0xffffffffu, s_acgt, o); n_acgt += __shfl_xor_sync(0xffffffffu, n_acgt, o);
        s_ave += __shfl_xor_sync(0xffffffffu, s_ave, o); n_ave += __shfl_xor_sync(0xffffffffu, n_ave, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&red[0], s_blk); atomicAdd(&red[1], s_acgt); atomicAdd(&red[2], n_acgt); atomicAdd(&red[3], s_ave); atomicAdd(&red[4], n_ave); }
    __syncthreads();
    if (threadIdx.x == 0) { PreOut o; o.blk_sum = red[0]; o.acgt_sum = red[1]; o.acgt_cnt = red[2]; o.ave_sum = red[3]; o.ave_cnt = red[4]; out[blockIdx.x] = o; }
    for (int i = threadIdx.x; i < HIST; i += blockDim.x) if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

// ---- K2: dinucleotide-repeat runs >= 20 (src/GROM.c:1727-1764): the thread at a run start walks the run
struct RepRec { int32_t s, e, type, pad; long long depth_sum; };
__device__ __forceinline__ int dinuc_type(char c0, char c1)
{
    if (((c0 ^ c1) & 0x20) != 0) return 10;                 // mixed case never matches
    auto code = [](char c) { c &= 0xDF; return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; };
    int a = code(c0), b = code(c1);
    if (a < 0 || b < 0) return 10;
    if (a > b) { const int t = a; a = b; b = t; }
    return a * 4 - a * (a - 1) / 2 + (b - a);               // 0..9 over the unordered pairs AA AC AG AT CC CG CT GG GT TT
}
__global__ void __launch_bounds__(256) k_repeats(const char *__restrict__ fasta, const int32_t *__restrict__ depth, int64_t lo, int64_t hi,
                                                 RepRec *__restrict__ out, unsigned int cap, unsigned int *__restrict__ n_out)
{
    const int64_t p = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= hi) return;
    const int t = dinuc_type(fasta[p], fasta[p + 1]);
    if (t == 10) return;
    if (p > lo && dinuc_type(fasta[p - 1], fasta[p]) == t) return;          // not the first pair of its run
    int64_t e = p;
    long long sum = depth[p];
    while (e + 1 < hi && dinuc_type(fasta[e + 1], fasta[e + 2]) == t) { e++; sum += depth[e]; }
    if (e + 1 >= hi) return;                                                 // a run still open at the end of the span is never flushed
    if (e - p < 19) return;
    const unsigned int k = atomicAdd(n_out, 1u);             // [s, e + 1) is the reference's half-open run
    if (k < cap) { RepRec r; r.s = (int32_t)p; r.e = (int32_t)(e + 1); r.type = t; r.pad = 0; r.depth_sum = sum; out[k] = r; }
}

// ---- K3: depth samples every insert_mean/2 bases of the sample blocks (src/GROM.c:18373-18456)
struct Sample { int32_t depth; int32_t code; };     // code: bit0 valid, bits1-2 class (0 high, 1 low, 2 uncovered), bits 8.. gc bin
__global__ void __launch_bounds__(256) k_samples(const int32_t *__restrict__ depth, const int32_t *__restrict__ rd, const int32_t *__restrict__ low,
                                                 const uint8_t *__restrict__ mq8, const int32_t *__restrict__ gc, const int32_t *__restrict__ acgt,
                                                 const int64_t *__restrict__ blk_start, const int64_t *__restrict__ blk_first, int n_blk,
                                                 int64_t n_samples, int64_t step, int q, Sample *__restrict__ out)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_samples) return;
    int b = 0;
    while (b + 1 < n_blk && blk_first[b + 1] <= j) b++;
    const int64_t p = blk_start[b] + (j - blk_first[b]) * step;
    Sample s; s.depth = depth[p]; s.code = 0;
    if (acgt[p] >= MIN_ACGT) {
        const int cls = (rd[p] == 0 && low[p] == 0) ? 2 : (mq8[p] >= q ? 0 : 1);
        s.code = 1 | (cls << 1) | (gc[p] << 8);
    }
    out[j] = s;
}

// ---- carry tiles: "class of the last setter position before p" (the reference's ddd_last_low_mq) ------------------------
// summary: class (0/1) of the last setter in the tile, 2 if none
__global__ void __launch_bounds__(256) k_tile_last_mask(const int32_t *__restrict__ depth, const uint8_t *__restrict__ mq8, const int32_t *__restrict__ acgt,
                                                        int64_t lo, int64_t hi, int q, uint8_t *__restrict__ tile_last)
{
    __shared__ int best;
    if (threadIdx.x == 0) best = -1;
    __syncthreads();
    const int64_t t0 = (int64_t)blockIdx.x * CTILE;
    int mine = -1;
    for (int i = threadIdx.x; i < CTILE; i += blockDim.x) {
        const int64_t p = t0 + i;
        if (p >= lo && p < hi && acgt[p] >= MIN_ACGT && depth[p] > 0) mine = max(mine, i);
    }
    if (mine >= 0) atomicMax(&best, mine);
    __syncthreads();
    if (threadIdx.x == 0) tile_last[blockIdx.x] = best < 0 ? 2 : (mq8[t0 + best] >= q ? 0 : 1);
}
// exclusive "last non-2" scan over the tile summaries, one block
__global__ void __launch_bounds__(1024) k_carry_scan(const uint8_t *__restrict__ tile_last, uint8_t *__restrict__ tile_in, int n_tiles)
{
    __shared__ uint8_t chunk_last[1024], chunk_in[1024];
    const int per = (n_tiles + 1023) / 1024, a = threadIdx.x * per, b = min(a + per, n_tiles);
    uint8_t last = 2;
    for (int t = a; t < b; t++) if (tile_last[t] != 2) last = tile_last[t];
    chunk_last[threadIdx.x] = last;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint8_t c = 0;                                   // ddd_last_low_mq starts at 0
        for (int i = 0; i < 1024; i++) { chunk_in[i] = c; if (chunk_last[i] != 2) c = chunk_last[i]; }
    }
    __syncthreads();
    uint8_t c = chunk_in[threadIdx.x];
    for (int t = a; t < b; t++) { tile_in[t] = c; if (tile_last[t] != 2) c = tile_last[t]; }
}

// ---- K5: mask (src/GROM.c:18681-18720), class, list-size bits; summary of the z-stage setters per tile
__global__ void __launch_bounds__(256) k_mask(const int32_t *__restrict__ depth, const uint8_t *__restrict__ mq8, const int32_t *__restrict__ gc,
                                              const int32_t *__restrict__ acgt, int64_t P, int64_t lo, int64_t hi, int q, const int32_t *__restrict__ nlist,
                                              const uint8_t *__restrict__ tile_in, uint32_t *__restrict__ rec, uint8_t *__restrict__ tile_last_z)
{
    __shared__ int best;
    if (threadIdx.x == 0) best = -1;
    __syncthreads();
    const int64_t t0 = (int64_t)blockIdx.x * CTILE;
    int mine = -1;
    for (int i = threadIdx.x; i < CTILE; i += blockDim.x) {
        const int64_t p = t0 + i;
        if (p >= P) break;
        const int d = depth[p], m = mq8[p];
        const int cls = m >= q ? 0 : (d > 0 ? 1 : 2);
        uint32_t r = (uint32_t)cls << R_CLASS | (uint32_t)m << R_MQ | R_MASK;
        if (p >= lo) {
            const int g = gc[p];
            if (nlist[g] > 1) r |= R_WIN0;
            if (nlist[NB + g] > 1) r |= R_WIN1;
        }
        if (p >= lo && p < hi) {
            const int g = gc[p];
            if (acgt[p] >= MIN_ACGT) {
                int mi;
                if (d > 0) mi = m >= q ? 0 : 1;
                else {
                    // uncovered: the list of the last covered position with enough ACGT context (walk back inside the tile, else the tile's carry-in)
                    mi = -1;
                    for (int64_t b = p - 1; b >= t0 && b >= lo; b--) if (acgt[b] >= MIN_ACGT && depth[b] > 0) { mi = mq8[b] >= q ? 0 : 1; break; }
                    if (mi < 0) mi = tile_in[blockIdx.x];
                }
                if (nlist[mi * NB + g] >= NO_COMBINE) r &= ~R_MASK;
            }
            if (rec_usable(r)) { r |= R_USABLE; if (cls != 2) mine = max(mine, i); }
        }
        rec[p] = r;
    }
    if (mine >= 0) atomicMax(&best, mine);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint8_t v = 2;
        if (best >= 0) v = mq8[t0 + best] >= q ? 0 : 1;
        tile_last_z[blockIdx.x] = v;
    }
}

// the reference's bisections on a sorted sample list, answered from the cumulative counts (exact for n >= 3; n < 3 spelled out)
__device__ __forceinline__ int rank_le(const Tables &T, int list, int v)   // bisect_right
{
    const int n = T.n[list];
    if (n >= 3) return v < 0 ? 0 : T.cum[(int64_t)list * (T.D + 1) + min(v, T.D)];
    if (n == 2) return v < T.small[list * 2 + 1] ? 1 : 2;
    return v < T.small[list * 2] ? 0 : 1;
}
__device__ __forceinline__ int rank_lt(const Tables &T, int list, int v)   // bisect_left
{
    const int n = T.n[list];
    if (n >= 3) return v <= 0 ? 0 : T.cum[(int64_t)list * (T.D + 1) + min(v - 1, T.D)];
    if (n == 2) return v <= T.small[list * 2 + 1] ? 1 : 2;
    return v <= T.small[list * 2] ? 0 : 1;
}

// ---- K6: rank -> probability -> sd units (src/GROM.c:18754-18963), threshold bits, seed bitmaps
__global__ void __launch_bounds__(256) k_z(const int32_t *__restrict__ depth, const int32_t *__restrict__ gc, int64_t P, int64_t lo, int64_t hi, int q,
                                           Tables T, const uint8_t *__restrict__ tile_in, uint32_t *__restrict__ rec,
                                           uint32_t *__restrict__ seed_del, uint32_t *__restrict__ seed_dup)
{
    __shared__ double sp[P2S];
    for (int i = threadIdx.x; i < P2S; i += blockDim.x) sp[i] = T.p2s_p[i];
    __syncthreads();
    const int64_t t0 = (int64_t)blockIdx.x * CTILE;
    for (int i = threadIdx.x; i < CTILE; i += blockDim.x) {
        const int64_t p = t0 + i;
        uint32_t r = 0;
        bool s_del = false, s_dup = false;
        if (p < P) {
            r = rec[p];
            const int d = depth[p];
            if (p >= lo) {
                // thresholds of the position's GC bin; past the analysed span the reference reads bin 0 of a zero-filled array
                const int g = gc[p];
                if ((double)d <= T.del_thr[g]) r |= R_DEL0;
                if ((double)d <= T.del_thr[NB + g]) r |= R_DEL1;
                if ((double)d >= T.dup_thr[g]) r |= R_DUP0;
                if ((double)d >= T.dup_thr[NB + g]) r |= R_DUP1;
            }
            if (p >= lo && p < hi) {
                const int g = gc[p];
                const int cls = (r >> R_CLASS) & 3;
                s_del = cls == 0 ? (r & R_DEL0) : cls == 1 ? (r & R_DEL1) : (r & (R_DEL0 | R_DEL1));
                s_dup = cls == 0 ? (r & R_DUP0) : cls == 1 ? (r & R_DUP1) : (r & (R_DUP0 | R_DUP1));
                if (r & R_USABLE) {
                    int mi;
                    if (cls == 0) mi = 0;
                    else if (cls == 1) mi = 1;
                    else {
                        mi = -1;
                        for (int64_t b = p - 1; b >= t0 && b >= lo; b--) { const uint32_t rb = rec[b]; if ((rb & R_USABLE) && ((rb >> R_CLASS) & 3) != 2) { mi = (rb >> R_CLASS) & 3; break; } }
                        if (mi < 0) mi = tile_in[blockIdx.x];
                    }
                    const int list = mi * NB + g, n = T.n[list];
                    if (n > 0) {
                        const double ave = T.ave[list];
                        int i1, i2;
                        bool neg;
                        if ((double)d < ave) { i1 = rank_le(T, list, d); i2 = rank_lt(T, list, d); neg = false; }
                        else {
                            if ((double)d > 2 * ave) i1 = rank_lt(T, list, (int)(2 * ave)); else i1 = rank_lt(T, list, d);
                            i2 = rank_le(T, list, d);
                            i1 = n - i1; i2 = n - i2; neg = true;
                        }
                        const double prob = ((i1 <= 0 ? 0.5 : (double)i1) + (i2 <= 0 ? 0.5 : (double)i2)) / (double)(2 * (long long)n);
                        int a = 0, b = P2S;                       // first table entry > prob (bisect_right over the ascending p-value table)
                        while (a < b) { const int m = (a + b) >> 1; if (prob < sp[m]) b = m; else a = m + 1; }
                        if (a >= P2S) a = P2S - 1;
                        r |= R_NZ | (neg ? R_NEG : 0u) | ((uint32_t)a << R_K);
                    }
                }
            }
            rec[p] = r;
        }
        const unsigned bd = __ballot_sync(0xffffffffu, s_del), bu = __ballot_sync(0xffffffffu, s_dup);
        if ((threadIdx.x & 31) == 0 && p < ((P + 31) / 32) * 32) { seed_del[p >> 5] = bd; seed_dup[p >> 5] = bu; }
    }
}

// ---- K7: window-length sweep (src/GROM.c:18967-19018).  The walk over a sample block, repeated at each -A offset without resetting
// the running frame, is cut into frames of Lmax elements; every prefix mean of a frame is summed in element order (the reference's
// summation order), so each of the 9,901 values per frame is the reference's double.
struct SweepBlock { int64_t start, end; int64_t first_frame, n_frames; };
__device__ __forceinline__ void sweep_advance(int &a, int64_t &p, int64_t n, int64_t s, int64_t e, int A, int Lmax)
{
    // move n elements along the concatenated walk; a == A means the walk is over
    while (a < A) {
        const int64_t rem = max((int64_t)0, e - p);
        if (n < rem) { p += n; return; }
        n -= rem; a++; p = s + (int64_t)a * Lmax / A;
    }
}
// One warp owns 32 consecutive frames, one lane per frame.  Per chunk of 32 elements: (A) the 32 x 32 records are fetched row by row
// (coalesced: a row is 32 consecutive elements of one frame's walk), decoded and parked in shared memory; (B) every lane folds ITS frame's
// 32 values into its running sum in element order -- one dependent DADD per element, the reference's summation order -- leaving the
// prefix sums in place; (C) the tile is read back transposed: lane = window length, so the division, the square and the store of
// X[frame][L] run in parallel over 32 lengths and the stores are coalesced rows.
#define SWEEP_PITCH 33
__global__ void __launch_bounds__(32) k_sweep(const uint32_t *__restrict__ rec, const SweepBlock *__restrict__ blocks, int n_blocks, int64_t n_frames,
                                              int A, int Lmin, int Lmax, int q, const double *__restrict__ p2s_sd, const double *__restrict__ wtab_g, double *__restrict__ X)
{
    __shared__ double sd[P2S];
    __shared__ double wtab[256];
    __shared__ double zt[32 * SWEEP_PITCH];
    __shared__ int st_a[32], st_p[32], st_s[32], st_e[32], st_n0[32];
    __shared__ unsigned st_um[32], st_vm[32];
    const int lane = threadIdx.x;
    for (int i = lane; i < P2S; i += 32) sd[i] = p2s_sd[i];
    for (int i = lane; i < 256; i += 32) wtab[i] = wtab_g[i];
    const int64_t f0 = (int64_t)blockIdx.x * 32, f = f0 + lane;
    const int n_rows = (int)min((int64_t)32, n_frames - f0);
    const int n_len = Lmax - Lmin + 1;
    // this lane's frame: block, walk state at the frame start
    int a = A; int64_t p = 0, s = 0, e = 0;
    if (f < n_frames) {
        int lo_b = 0, hi_b = n_blocks - 1;                       // last block whose first frame is <= f
        while (lo_b < hi_b) { const int m = (lo_b + hi_b + 1) >> 1; if (blocks[m].first_frame <= f) lo_b = m; else hi_b = m - 1; }
        s = blocks[lo_b].start; e = blocks[lo_b].end; a = 0; p = s;
        sweep_advance(a, p, (f - blocks[lo_b].first_frame) * (int64_t)Lmax, s, e, A, Lmax);
    }
    st_s[lane] = (int)s; st_e[lane] = (int)e;
    double tot = 0.0;
    int n_us = 0;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int w0 = 0; w0 < Lmax; w0 += 32) {
        st_a[lane] = a; st_p[lane] = (int)p;
        __syncwarp();
        unsigned my_um = 0, my_vm = 0;
        const bool in_len = w0 + lane < Lmax;
#pragma unroll 8
        for (int r = 0; r < 32; r++) {                           // (A) row r = frame f0 + r, this lane = element w0 + lane of its walk
            int la = st_a[r]; int64_t lp = st_p[r];
            sweep_advance(la, lp, lane, (int64_t)st_s[r], (int64_t)st_e[r], A, Lmax);
            const bool valid = la < A && in_len;
            const uint32_t rr = valid ? __ldg(rec + lp) : R_MASK;
            const bool us = valid && rec_usable(rr);
            double z = 0.0;
            if (us && (rr & R_NZ)) {
                const double wt = (rr & R_OVR) ? 1.0 : (((rr >> R_CLASS) & 3) == 0 ? wtab[(rr >> R_MQ) & 255] : 0.5);
                z = __dmul_rn(wt, sd[(rr >> R_K) & 1023]);
                if (rr & R_NEG) z = -z;
            }
            const unsigned um = __ballot_sync(0xffffffffu, us), vm = __ballot_sync(0xffffffffu, valid);
            zt[r * SWEEP_PITCH + lane] = z;
            if (lane == r) { my_um = um; my_vm = vm; }
        }
        __syncwarp();
        const int n0 = n_us;                                      // (B) the ordered fold of this lane's frame
        double *mine = zt + lane * SWEEP_PITCH;
#pragma unroll
        for (int j = 0; j < 32; j++) { if ((my_um >> j) & 1u) tot = __dadd_rn(tot, mine[j]); mine[j] = tot; }
        n_us += __popc(my_um);
        st_um[lane] = my_um; st_vm[lane] = my_vm; st_n0[lane] = n0;
        __syncwarp();
        const int w = w0 + lane + 1;                              // (C) lane = window length
        if (w >= Lmin && w <= Lmax) {
            const unsigned below = 0xffffffffu >> (31 - lane);
#pragma unroll 4
            for (int r = 0; r < n_rows; r++) {
                const int n = st_n0[r] + __popc(st_um[r] & below);
                double x2 = nan;
                if (((st_vm[r] >> lane) & 1u) && n > 0) { const double x = zt[r * SWEEP_PITCH + lane] / (double)n; x2 = __dmul_rn(x, x); }
                __stcs(X + (f0 + r) * (int64_t)n_len + (w - Lmin), x2);
            }
        }
        __syncwarp();
        sweep_advance(a, p, 32, s, e, A, Lmax);
    }
}
// ordered sum over the frames per window length: a CTA owns 32 lengths; all eight warps stream tiles of 64 frames x 32 lengths into a
// four-stage shared-memory ring (8-byte async copies), warp 0 adds them in frame order (one lane per length)
#define SSUM_FR 64
#define SSUM_ST 4
__global__ void __launch_bounds__(256) k_sweep_sum(const double *__restrict__ X, int64_t n_frames, int n_len, double *__restrict__ wsq, long long *__restrict__ wcnt)
{
    extern __shared__ __align__(16) double ssum_buf[];           // [SSUM_ST][SSUM_FR][32]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int L = blockIdx.x * 32 + lane;
    const bool l_ok = L < n_len;
    const int64_t n_tiles = (n_frames + SSUM_FR - 1) / SSUM_FR;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    auto issue = [&](int64_t t) {
        if (t < n_tiles) {
            double *dst = ssum_buf + (size_t)(t % SSUM_ST) * SSUM_FR * 32;
#pragma unroll
            for (int k = 0; k < SSUM_FR / 8; k++) {
                const int row = wid + 8 * k;
                const int64_t fr = t * SSUM_FR + row;
                double *d = dst + row * 32 + lane;
                if (l_ok && fr < n_frames) {
                    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(d);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(X + fr * (int64_t)n_len + L) : "memory");
                } else *d = nan;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int t = 0; t < SSUM_ST - 1; t++) issue(t);
    double sum = 0.0;
    long long cnt = 0;
    for (int64_t t = 0; t < n_tiles; t++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(SSUM_ST - 2) : "memory");
        __syncthreads();                                          // tile t has landed for everyone; warp 0 is done with tile t - 1
        issue(t + SSUM_ST - 1);                                   // into the buffer tile t - 1 occupied
        if (wid == 0) {
            const double *src = ssum_buf + (size_t)(t % SSUM_ST) * SSUM_FR * 32 + lane;
            double v[SSUM_FR];
#pragma unroll
            for (int i = 0; i < SSUM_FR; i++) v[i] = src[i * 32];
#pragma unroll
            for (int i = 0; i < SSUM_FR; i++) if (v[i] == v[i]) { sum = __dadd_rn(sum, v[i]); cnt++; }
        }
    }
    if (wid == 0 && l_ok) { wsq[L] = sum; wcnt[L] = cnt; }
}

// ---- greedy segmentation (src/GROM.c:19361-19678 deletions, 19702-20010 duplications) --------------------------------------------
// What happens at one seed -- a position whose depth is beyond the threshold -- depends only on the seed and, for an uncovered
// seed, on the carried MAPQ class: walk Lmin positions and give up as soon as fewer than half are beyond the threshold; otherwise
// grow the window to Lmax scoring every length, slide it, trim the end.  eval_seed() is that function, shared by the device (every
// seed evaluated speculatively in parallel, bounded to SEED_BOUND positions) and the host (the sequential hop from seed to seed,
// plus the few evaluations the device left unresolved).
struct SegCtx {
    const uint32_t *rec; int64_t len, end; int q, Lmin, Lmax, bound; const double *sd, *win_sd; bool dup;
    const double *wtab;          // MAPQ weight per mean MAPQ value [256]: the expression of rec_z evaluated once per value (same doubles, no division per base)
    const double *win_thr;       // [Lmax + 1] 2.97 * win_sd[L], +inf where win_sd[L] <= 0: the cheap side of scores()
    const double *zarr;          // device only: z of the deletion scan per position, unpacked once (k_zfill); nullptr = unpack from rec
    __host__ __device__ inline int cls(int64_t p) const { return (rec[p] >> R_CLASS) & 3; }
    __host__ __device__ inline uint32_t beyond_bit(int mi) const { return dup ? (mi ? R_DUP1 : R_DUP0) : (mi ? R_DEL1 : R_DEL0); }
    __host__ __device__ inline bool beyond(int64_t p, int mi) const { return (rec[p] & beyond_bit(mi)) != 0; }
    __host__ __device__ inline bool win_gt1(int64_t p, int mi) const { return (rec[p] & (mi ? R_WIN1 : R_WIN0)) != 0; }
    // z of the deletion scan (the duplication scan sees the opposite sign)
    __host__ __device__ inline double z_del_of(const uint32_t r) const
    {
        if (!(r & R_NZ)) return 0.0;
        const double w = (r & R_OVR) ? 1.0 : (((r >> R_CLASS) & 3) == 0 ? wtab[(r >> R_MQ) & 255] : 0.5);
#ifdef __CUDA_ARCH__
        const double v = __dmul_rn(w, sd[(r >> R_K) & 1023]);
#else
        const double v = w * sd[(r >> R_K) & 1023];
#endif
        return (r & R_NEG) ? -v : v;
    }
    __host__ __device__ inline double z(int64_t p) const
    {
        const double v = zarr ? zarr[p] : z_del_of(rec[p]);
        return dup ? 0.0 - v : v;            // 0.0 - v: exact, and keeps a zero positive like the reference's literal 0.0
    }
    // score >= 3 ?  The division is only carried out when the quotient can be anywhere near the threshold (a 1 % margin dwarfs
    // the rounding of the products and the quotient), so the outcome is the reference's in every case; tot > 0 and win_sd > 0,
    // which the reference tests first, follow from score >= 3 and from the +inf entries of win_thr.
    __host__ __device__ inline bool scores(double tot, int cnt, int L, double *score) const
    {
        if (!(tot >= cnt * win_thr[L])) return false;
        *score = tot / (cnt * win_sd[L]);
        return *score >= 3;
    }
};
enum { SEG_RESUME = 0, SEG_CALL = 1, SEG_UNRESOLVED = 2 };
struct Outcome { int kind; int64_t next, c_end; double c_z; int64_t far; };   // far: one past the last position the sliding phase looked at
constexpr int SEED_BOUND = 1024;       // first round, every seed; the second round gives the compacted unresolved ones the full growth phase (Lmax)

template <bool BOUNDED> __host__ __device__ inline Outcome eval_seed(const SegCtx &C, const int64_t pos, int mi)
{
#define CNV_STEP(var, p) do { const int c_ = C.cls(p); if (c_ != 2) var = c_; } while (0)
    const int64_t Lmin = C.Lmin, Lmax = C.Lmax, end = C.end, max_gap = Lmax + 500;
    Outcome o; o.kind = SEG_RESUME; o.next = pos + 1; o.c_end = 0; o.c_z = 0; o.far = 0;
    bool stop = false, begun = false;
    int wlen = 0, cnt = 0, cnt2 = 0;                        // window length and counters fit 32 bits (Lmax positions at most)
    int64_t pa, c_start = 0, c_end = 0, last_good = 0;
    double tot = 0, c_z = 0, tz;
    const uint32_t *rp = C.rec + pos;                        // the first two phases index relative to the seed (32-bit offsets)
    const double *zp = C.zarr ? C.zarr + pos : nullptr;
    auto zrel = [&](int i, uint32_t r) { const double v = zp ? zp[i] : C.z_del_of(r); return C.dup ? 0.0 - v : v; };
    const int iLmin = (int)Lmin, iLmax = (int)Lmax;
    for (int i = 0; i < iLmin; i++) {
        wlen++;
        bool ok = false;
        const uint32_t r = rp[i];
        if (!(r & R_MASK)) { const int c_ = (r >> R_CLASS) & 3; if (c_ != 2) mi = c_; ok = (r & C.beyond_bit(mi)) != 0; }
        if (ok) cnt2++;
        else if (2 * cnt2 < wlen) { o.next = pos + i + 1; return o; }       // give up inside the first window: resume after the offender
    }
    cnt = iLmin;
    for (int i = 0; i < iLmin; i++) { const uint32_t r = rp[i]; cnt -= (int)(r & R_MASK); tot += zrel(i, r); }
    if (cnt > 0 && C.scores(tot, cnt, iLmin, &tz)) {
        begun = true; c_start = pos; last_good = c_end = pos + Lmin; c_z = tz;
    }
    {
        const int i_end = end - pos < (int64_t)iLmax ? (int)(end - pos) : iLmax;      // first offset at or past `end` (>= Lmin is not guaranteed)
        int good = -1;                                                               // last scoring offset of this phase
        const uint32_t m0 = C.beyond_bit(0), m1 = C.beyond_bit(1);
        for (int i = iLmin; i < iLmax; i++) {
            wlen++;
            if (BOUNDED && wlen > C.bound) { o.kind = SEG_UNRESOLVED; return o; }
            if (i >= i_end) { stop = true; break; }
            bool ok = false;
            const uint32_t r = rp[i];
            if (!(r & R_MASK)) {
                const int c_ = (r >> R_CLASS) & 3; if (c_ != 2) mi = c_;
                tot += zrel(i, r); cnt++;
                ok = (r & (mi ? m1 : m0)) != 0;
                if (ok) {
                    cnt2++;
                    if (C.scores(tot, cnt, wlen, &tz)) {
                        good = i;
                        if (tz > c_z) c_z = tz;                                      // c_z starts at 0 and every scoring tz is >= 3
                    }
                }
            }
            if (!ok && 2 * cnt2 < wlen) { stop = true; break; }
        }
        if (good >= 0) {
            last_good = pos + good;
            if (!begun) { begun = true; c_start = pos; }
            c_end = pos + good;
        }
    }
    if (!stop && begun) {
        if (BOUNDED) { o.kind = SEG_UNRESOLVED; return o; }
        int mi_b = mi;
        pa = pos + Lmax; tot = 0; cnt = 0;
        while (pa < C.len && pa - last_good <= max_gap) {
            if (pa == pos + Lmax) {
                for (int64_t pb = pa - Lmax + 1; pb < pa + 1; pb++) {
                    CNV_STEP(mi_b, pb);
                    if (!(C.rec[pb] & R_MASK) && C.win_gt1(pb, mi_b)) { tot += C.z(pb); cnt++; }
                }
            } else {
                const int64_t pb = pa - Lmax;
                CNV_STEP(mi_b, pb);
                if (!(C.rec[pb] & R_MASK) && C.win_gt1(pb, mi_b)) { tot -= C.z(pb); cnt--; }
                CNV_STEP(mi, pa);
                if (!(C.rec[pa] & R_MASK) && C.win_gt1(pa, mi)) { tot += C.z(pa); cnt++; }
            }
            if (cnt > 0 && C.scores(tot, cnt, iLmax, &tz)) {
                last_good = pa; c_end = pa;
                if (tz > c_z) c_z = tz;
            }
            pa++;
        }
        o.far = pa;
    }
    if (!begun) return o;                                                  // gave up while growing (or never scored): resume at seed + 1
    int64_t t = c_end;                                                     // trim the end back to a stretch that is still mostly beyond
    while (t > c_start + Lmin) {
        CNV_STEP(mi, t);
        if (!C.beyond(t, mi)) { t--; c_end = t; }
        else {
            int64_t c2 = 0, c3 = 0;
            bool halt = false;
            int mi_a = mi;
            pa = c_end;
            while (pa > c_start + Lmin && !halt) {
                if (!(C.rec[pa] & R_MASK)) { CNV_STEP(mi_a, pa); c3++; if (C.beyond(pa, mi_a)) c2++; }
                if (c3 == 0 || c2 / (double)c3 < 0.5) { c_end = pa - 1; halt = true; }
                pa--;
            }
            t = pa;
        }
    }
    o.kind = SEG_CALL; o.c_end = c_end; o.c_z = c_z; o.next = c_end + 2;
    return o;
#undef CNV_STEP
}

// land[2*rank + class] (uint32): 0xFFFFFFFF not a seed under that class; top bits SEG_*; RESUME: low bits = distance to the next
// position; CALL: low bits = index into the speculative call list
constexpr uint32_t LAND_NOT = 0xFFFFFFFFu;
constexpr int LAND_SHIFT = 29;
struct SeedCall { int64_t c_end; double c_z; };
constexpr int SEED_WORDS = 1024;            // seed-bitmap words per CTA
__global__ void __launch_bounds__(256) k_seed_blocksum(const uint32_t *__restrict__ seeds, int64_t words, uint32_t *__restrict__ blk, int nb)
{
    __shared__ uint32_t red[8];
    const uint32_t *sd = seeds + (int64_t)blockIdx.y * words;
    uint32_t n = 0;
    for (int j = 0; j < 4; j++) { const int64_t w = (int64_t)blockIdx.x * SEED_WORDS + threadIdx.x * 4 + j; if (w < words) n += __popc(sd[w]); }
    for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int i = 0; i < 8; i++) t += red[i]; blk[blockIdx.y * nb + blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024) k_seed_blockscan(uint32_t *__restrict__ blk, int nb, uint32_t *__restrict__ totals)
{
    __shared__ uint32_t part[1024];
    uint32_t *b = blk + blockIdx.x * nb;
    const int per = (nb + 1023) / 1024, i0 = threadIdx.x * per, i1 = min(i0 + per, nb);
    uint32_t t = 0;
    for (int i = i0; i < i1; i++) t += b[i];
    part[threadIdx.x] = t;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t run = 0; for (int i = 0; i < 1024; i++) { const uint32_t v = part[i]; part[i] = run; run += v; } totals[blockIdx.x] = run; }
    __syncthreads();
    uint32_t run = part[threadIdx.x];
    for (int i = i0; i < i1; i++) { const uint32_t v = b[i]; b[i] = run; run += v; }
}
// per-word seed ranks (exclusive prefix of the bitmap popcounts), one CTA per SEED_WORDS words
__global__ void __launch_bounds__(256) k_seed_rank(const uint32_t *__restrict__ seeds, int64_t words, const uint32_t *__restrict__ blk, int nb, uint32_t *__restrict__ wp)
{
    __shared__ uint32_t part[256];
    const int kind = blockIdx.y;
    const uint32_t *sd = seeds + (int64_t)kind * words;
    uint32_t bits[4], n = 0;
    const int64_t w0 = (int64_t)blockIdx.x * SEED_WORDS + threadIdx.x * 4;
    for (int j = 0; j < 4; j++) { bits[j] = w0 + j < words ? sd[w0 + j] : 0u; n += __popc(bits[j]); }
    part[threadIdx.x] = n;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t run = blk[kind * nb + blockIdx.x]; for (int i = 0; i < 256; i++) { const uint32_t v = part[i]; part[i] = run; run += v; } }
    __syncthreads();
    uint32_t rank = part[threadIdx.x];
    for (int j = 0; j < 4; j++) { if (w0 + j < words) wp[(int64_t)kind * words + w0 + j] = rank; rank += __popc(bits[j]); }
}
// ---- the hop on the device.  Node = (seed rank, carried class); its successor is a pure function of the node's outcome: where
// the outer loop resumes, which seed it meets next and which class it carries there (src/GROM.c:19370-19389).  With the successor
// array in hand the visited path is found by pointer doubling: J_k = J_{k-1} o J_{k-1}, then marks spread from the start node
// through J_{K-1} .. J_0, which reaches exactly the nodes at every path index.  Unresolved nodes are sinks (self loops).
__device__ __forceinline__ uint32_t next_node(const SegCtx &C, const uint32_t *__restrict__ sd, const uint32_t *__restrict__ wpk, int64_t x, int s, uint32_t n_seeds)
{
    const int64_t end = C.end;
    if (x >= end) return 2 * n_seeds;                                            // END
    int64_t w = x >> 5;
    uint32_t bits = sd[w] & (0xffffffffu << (x & 31));
    const int64_t wend = (end + 31) >> 5;
    while (!bits) { if (++w >= wend) return 2 * n_seeds; bits = sd[w]; }
    const int bit = __ffs(bits) - 1;
    const int64_t q = (w << 5) + bit;
    if (q >= end) return 2 * n_seeds;
    int v = C.cls(q);
    if (v == 2) { v = s; for (int64_t b = q - 1; b >= x; b--) { const int c = C.cls(b); if (c != 2) { v = c; break; } } }
    return 2 * (wpk[w] + __popc(sd[w] & ((1u << bit) - 1u))) + (uint32_t)v;
}
__device__ __forceinline__ uint32_t successor(const SegCtx &C, const uint32_t *__restrict__ sd, const uint32_t *__restrict__ wpk, int64_t p, int c0, int v, uint32_t e,
                                              const SeedCall *__restrict__ calls, uint32_t rank, uint32_t n_seeds)
{
    const uint32_t kind = e >> LAND_SHIFT, low = e & ((1u << LAND_SHIFT) - 1u);
    const int s = c0 != 2 ? c0 : v;                                              // class carried past this seed
    if (e == LAND_NOT) return next_node(C, sd, wpk, p + 1, s, n_seeds);
    if (kind == SEG_RESUME) return next_node(C, sd, wpk, p + low, s, n_seeds);
    if (kind == SEG_CALL) return next_node(C, sd, wpk, calls[low].c_end + 2, s, n_seeds);
    return 2 * rank + (uint32_t)v;                                               // unresolved: sink
}
__global__ void __launch_bounds__(256) k_hop_double(const uint32_t *__restrict__ a, uint32_t *__restrict__ b, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = a[a[i]];
}
__global__ void __launch_bounds__(256) k_hop_mark(const uint32_t *__restrict__ j, uint8_t *__restrict__ flag, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) flag[j[i]] = 1;
}
struct HopCall { int64_t pos, c_end; double c_z; };
struct HopSink { int64_t pos; int32_t variant, found; };        // found: 1 = the path stops at an unresolved seed, 0 = it ran to the end
// calls of the flagged (= visited) nodes; unresolved nodes on the path were evaluated by the host (k_hop_advance)
__global__ void __launch_bounds__(256) k_hop_collect(const uint8_t *__restrict__ flag, const uint32_t *__restrict__ land, const uint32_t *__restrict__ sd,
                                                     const uint32_t *__restrict__ wpk, int64_t words, uint32_t n_seeds, const SeedCall *__restrict__ calls,
                                                     HopCall *__restrict__ out, uint32_t out_cap, unsigned int *__restrict__ n_out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;               // node id inside this kind
    if (i >= 2 * n_seeds || !flag[i]) return;
    const uint32_t e = land[i];
    if (e == LAND_NOT || (e >> LAND_SHIFT) != SEG_CALL) return;
    const uint32_t low = e & ((1u << LAND_SHIFT) - 1u), rank = i >> 1;
    int64_t a = 0, b = words;
    while (a < b) { const int64_t m = (a + b) >> 1; if (wpk[m] <= rank) a = m + 1; else b = m; }
    const int64_t w = a - 1, pos = (w << 5) + __fns(sd[w], 0, (int)(rank - wpk[w]) + 1);
    const unsigned int k = atomicAdd(n_out, 1u);
    if (k < out_cap) { out[k].pos = pos; out[k].c_end = calls[low].c_end; out[k].c_z = calls[low].c_z; }
}
// One leg of the path per scan (thread 0 deletions, thread 1 duplications): flag the node the outer loop meets first from position x
// carrying class s -- the marks later spread from these leg starts -- and follow the top level of the jump table to the leg's fixed
// point: the END node, or an unresolved seed (a sink), which the host evaluates before the next leg starts behind it.
struct HopLeg { int64_t x; int32_t s, active; };
__global__ void k_hop_advance(SegCtx C0, SegCtx C1, const uint32_t *__restrict__ seeds, const uint32_t *__restrict__ wp, int64_t words, const uint32_t *__restrict__ jtop,
                              HopLeg leg0, HopLeg leg1, uint32_t n0, uint32_t n1, uint8_t *__restrict__ flag, HopSink *__restrict__ sink)
{
    const int k = threadIdx.x;
    const HopLeg leg = k ? leg1 : leg0;
    if (!leg.active) return;
    const SegCtx &C = k ? C1 : C0;
    const uint32_t n_seeds = k ? n1 : n0, base = k ? 2 * n0 + 1 : 0;
    const uint32_t *sd = seeds + (int64_t)k * words, *wpk = wp + (int64_t)k * words;
    uint32_t node = base + next_node(C, sd, wpk, leg.x, leg.s, n_seeds);
    flag[node] = 1;
    for (;;) { const uint32_t nx = jtop[node]; if (nx == node) break; node = nx; }
    HopSink r; r.pos = 0; r.variant = 0; r.found = 0;
    if (node != base + 2 * n_seeds) {
        const uint32_t i = node - base, rank = i >> 1;
        int64_t a = 0, b = words;
        while (a < b) { const int64_t m = (a + b) >> 1; if (wpk[m] <= rank) a = m + 1; else b = m; }
        const int64_t w = a - 1;
        r.pos = (w << 5) + __fns(sd[w], 0, (int)(rank - wpk[w]) + 1); r.variant = (int)(i & 1); r.found = 1;
    }
    sink[k] = r;
}

// one thread per seed (blockIdx.y = deletions / duplications): rank -> position through the per-word ranks, then evaluate the seed
// (both carried classes when it is uncovered).  Seeds that run past the bound are appended to `todo` for the second round, which
// runs one thread per (kind, rank, class) entry of that list.
struct SeedTodo { uint32_t rank; uint8_t kind, variant; uint16_t pad; };
__device__ __forceinline__ int64_t seed_position(const uint32_t *__restrict__ sd, const uint32_t *__restrict__ wpk, int64_t words, uint32_t rank)
{
    int64_t a = 0, b = words;                                  // first word whose exclusive rank exceeds `rank`, minus one
    while (a < b) { const int64_t m = (a + b) >> 1; if (wpk[m] <= rank) a = m + 1; else b = m; }
    const int64_t w = a - 1;
    return (w << 5) + __fns(sd[w], 0, (int)(rank - wpk[w]) + 1);
}
__device__ __forceinline__ uint32_t seed_outcome(const SegCtx &C, int64_t p, int mi, SeedCall *__restrict__ calls, uint32_t call_cap, unsigned int *__restrict__ n_calls)
{
    const uint32_t unres = (uint32_t)SEG_UNRESOLVED << LAND_SHIFT;
    const Outcome o = eval_seed<true>(C, p, mi);
    if (o.kind == SEG_RESUME) return (uint32_t)(o.next - p);
    if (o.kind == SEG_CALL) {
        const unsigned int k = atomicAdd(n_calls, 1u);
        if (k < call_cap) { calls[k].c_end = o.c_end; calls[k].c_z = o.c_z; return ((uint32_t)SEG_CALL << LAND_SHIFT) | k; }
    }
    return unres;
}
__global__ void __launch_bounds__(128, 8) k_seed_eval(SegCtx Cdel, SegCtx Cdup, const uint32_t *__restrict__ seeds, int64_t words, const uint32_t *__restrict__ wp,
                                                   uint32_t *__restrict__ land, uint32_t cap, uint32_t n_del, uint32_t n_dup, SeedCall *__restrict__ calls, uint32_t call_cap,
                                                   unsigned int *__restrict__ n_calls, SeedTodo *__restrict__ todo, uint32_t todo_cap, uint32_t *__restrict__ jump0)
{
    const int kind = blockIdx.y;
    const uint32_t rank = blockIdx.x * blockDim.x + threadIdx.x;
    if (rank >= (kind ? n_dup : n_del) || rank >= cap) return;
    const SegCtx &C = kind ? Cdup : Cdel;
    const int64_t p = seed_position(seeds + (int64_t)kind * words, wp + (int64_t)kind * words, words, rank);
    uint32_t res[2] = {LAND_NOT, LAND_NOT};
    const uint32_t unres = (uint32_t)SEG_UNRESOLVED << LAND_SHIFT;
    if (p < C.end) {
        const int c0 = C.cls(p);
        for (int v = 0; v < 2; v++) {
            if (c0 != 2 && v == 1) { res[1] = res[0]; break; }
            const int mi = c0 == 2 ? v : c0;
            if (!C.beyond(p, mi)) continue;
            res[v] = seed_outcome(C, p, mi, calls, call_cap, n_calls);
            if (res[v] == unres) {
                const unsigned int k = atomicAdd(n_calls + 1, 1u);
                if (k < todo_cap) { SeedTodo t; t.rank = rank; t.kind = (uint8_t)kind; t.variant = (uint8_t)v; t.pad = 0; todo[k] = t; }
            }
        }
    }
    land[((int64_t)kind * cap + rank) * 2] = res[0]; land[((int64_t)kind * cap + rank) * 2 + 1] = res[1];
    if (jump0) {
        // successors; node ids are local to the kind, the table stores them behind the kind's base offset (2 * n_del + 1 for duplications)
        const uint32_t n_seeds = kind ? n_dup : n_del, base = kind ? 2 * n_del + 1 : 0;
        const uint32_t *sd = seeds + (int64_t)kind * words, *wpk = wp + (int64_t)kind * words;
        const int c0 = p < C.end ? C.cls(p) : 0;
        for (int v = 0; v < 2; v++)
            jump0[base + 2 * rank + v] = base + (p < C.end ? successor(C, sd, wpk, p, c0, v, res[v], calls, rank, n_seeds) : 2 * n_seeds);
        if (rank == 0) jump0[base + 2 * n_seeds] = base + 2 * n_seeds;      // END loops on itself
    }
}
// Second round over the compacted open seeds (full growth phase, one thread each).  Deep inside a long event every seed walks the
// whole growth phase only to stay open (it enters the sliding phase), and none of them is ever visited: the path meets the first
// open seed of the event, the host evaluates it and the call jumps past the rest.  So the round runs in two passes over blocks of
// OPEN_BLOCK positions: pass 0 evaluates the first open seed of every block; pass 1 evaluates the others unless the block's first
// seed and the next block's both stayed open.  Skipping is only a guess about cost -- a skipped seed stays open, and an open seed
// on the path is evaluated exactly by the host like any other.
constexpr int OPEN_BLOCK_SHIFT = 8;
enum { OPEN_NONE = 0, OPEN_STAYS = 1, OPEN_CLOSED = 2 };
__global__ void __launch_bounds__(256) k_open_first(const uint32_t *__restrict__ seeds, int64_t words, const uint32_t *__restrict__ wp, const SeedTodo *__restrict__ todo,
                                                    uint32_t n_todo, uint32_t *__restrict__ first, int64_t n_blocks)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_todo) return;
    const SeedTodo t = todo[i];
    const int64_t p = seed_position(seeds + (int64_t)t.kind * words, wp + (int64_t)t.kind * words, words, t.rank);
    atomicMin(first + (int64_t)(t.kind * 2 + t.variant) * n_blocks + (p >> OPEN_BLOCK_SHIFT), (uint32_t)p);
}
__global__ void __launch_bounds__(64) k_seed_eval2(SegCtx Cdel, SegCtx Cdup, const uint32_t *__restrict__ seeds, int64_t words, const uint32_t *__restrict__ wp,
                                                   uint32_t *__restrict__ land, uint32_t cap, SeedCall *__restrict__ calls, uint32_t call_cap,
                                                   unsigned int *__restrict__ n_calls, const SeedTodo *__restrict__ todo, uint32_t n_todo, uint32_t n_del, uint32_t n_dup,
                                                   uint32_t *__restrict__ jump0, int pass, const uint32_t *__restrict__ first, uint8_t *__restrict__ state, int64_t n_blocks)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_todo) return;
    const SeedTodo t = todo[i];
    const SegCtx &C = t.kind ? Cdup : Cdel;
    const int64_t p = seed_position(seeds + (int64_t)t.kind * words, wp + (int64_t)t.kind * words, words, t.rank);
    const uint32_t unres = (uint32_t)SEG_UNRESOLVED << LAND_SHIFT;
    const int64_t key = (int64_t)(t.kind * 2 + t.variant) * n_blocks + (p >> OPEN_BLOCK_SHIFT);
    if (pass == 2) { if (land[((int64_t)t.kind * cap + t.rank) * 2 + t.variant] != unres) return; }    // sweep of whatever is still open
    else if (pass >= 0) {
        const bool is_first = first[key] == (uint32_t)p;
        if (is_first != (pass == 0)) return;
        if (pass == 1 && state[key] == OPEN_STAYS && (p >> OPEN_BLOCK_SHIFT) + 1 < n_blocks && state[key + 1] == OPEN_STAYS) { atomicAdd(n_calls + 5, 1u); return; }
    }
    const int c0 = C.cls(p);
    const uint32_t e = seed_outcome(C, p, c0 == 2 ? t.variant : c0, calls, call_cap, n_calls);
    if (pass == 0) state[key] = e == unres ? OPEN_STAYS : OPEN_CLOSED;
    if (e == unres) return;                                                                // enters the sliding phase: a long call, left to the host
    land[((int64_t)t.kind * cap + t.rank) * 2 + t.variant] = e;
    if (c0 != 2) land[((int64_t)t.kind * cap + t.rank) * 2 + 1] = e;
    atomicAdd(n_calls + 4, 1u);                                                            // seeds closed by the second round
    if (jump0) {
        const uint32_t n_seeds = t.kind ? n_dup : n_del, base = t.kind ? 2 * n_del + 1 : 0;
        const uint32_t *sd = seeds + (int64_t)t.kind * words, *wpk = wp + (int64_t)t.kind * words;
        const uint32_t nx = base + successor(C, sd, wpk, p, c0, t.variant, e, calls, t.rank, n_seeds);
        jump0[base + 2 * t.rank + t.variant] = nx;
        if (c0 != 2) jump0[base + 2 * t.rank + 1] = base + successor(C, sd, wpk, p, c0, 1, e, calls, t.rank, n_seeds);
    }
}

// ---- K8: depth and GC bin of the called segments, packed back to back (copy-number step, src/GROM.c:20071-20224)
// out_gc: bits 0-6 GC bin, bit 7 = ACGT context >= 99 %
__global__ void __launch_bounds__(256) k_gather(const int32_t *__restrict__ depth, const int32_t *__restrict__ gc, const int32_t *__restrict__ acgt, const int64_t *__restrict__ seg_start,
                                                const int64_t *__restrict__ seg_first, int n_seg, int64_t total, int32_t *__restrict__ out_depth, uint8_t *__restrict__ out_gc,
                                                const uint32_t *__restrict__ rec, uint32_t *__restrict__ out_rec)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= total) return;
    int a = 0, b = n_seg - 1;
    while (a < b) { const int m = (a + b + 1) >> 1; if (seg_first[m] <= j) a = m; else b = m - 1; }
    const int64_t p = seg_start[a] + (j - seg_first[a]);
    out_depth[j] = depth[p]; out_gc[j] = (uint8_t)((gc[p] & 0x7f) | (acgt[p] >= MIN_ACGT ? 0x80 : 0));
    if (out_rec) out_rec[j] = rec[p];
}

// =====================================================================================================================
// host side: sequential logic with the reference's libc semantics
// =====================================================================================================================

// glibc rand(): TYPE_3 additive feedback generator, seeded as srandom_r does
struct GlibcRand {
    int32_t r[34]; int f = 3, b = 0;
    explicit GlibcRand(unsigned seed)
    {
        int32_t w = (int32_t)seed; if (w == 0) w = 1;
        r[0] = w;
        for (int i = 1; i < 31; i++) { long h = w / 127773, l = w % 127773, t = 16807 * l - 2836 * h; if (t < 0) t += 2147483647; w = (int32_t)t; r[i] = w; }
        for (int i = 0; i < 310; i++) next();
    }
    int next() { const uint32_t v = (uint32_t)r[f] + (uint32_t)r[b]; r[f] = (int32_t)v; f = (f + 1) % 31; b = (b + 1) % 31; return (int)(v >> 1); }
    long below(long max)            // grom_rand, src/GROM.c:1185-1203
    {
        long val = 0, scale = 1;
        while (scale < max) { long t = (next() % 10) * scale; while (t + val >= max) t = (next() % 10) * scale; val += t; scale *= 10; }
        return val;
    }
};

struct SampleList {
    std::vector<int> v; long n_all = 0;
    void add(int x, long cap, GlibcRand &g)
    {
        if ((long)v.size() < cap) { v.push_back(x); n_all++; }
        else { if (g.below(n_all) == 0) v[g.below((long)v.size())] = x; n_all++; }
    }
};

struct Call { int64_t start, end; double z; };
struct DevTmpRaw { void *p = nullptr; ~DevTmpRaw() { if (p) cudaFree(p); } };
struct Grow {                  // grow-only device buffer kept across calls
    void *p = nullptr; size_t cap = 0;
    bool ensure(size_t n) { if (n <= cap) return true; if (p) cudaFree(p); p = nullptr; cap = n + n / 4 + 256; return cudaMalloc(&p, cap) == cudaSuccess; }
    template <class T> T *as() { return (T *)p; }
};

// the reference's qsort on the copy-number ratios: glibc merge sort with the comparator `*(int*)a - *(int*)b`, i.e. ordered by the
// low word of each double with wrapping subtraction (src/GROM.c:1105, 20113)
inline void lowword_msort(double *b, size_t n, double *tmp)
{
    if (n <= 1) return;
    const size_t n1 = n / 2, n2 = n - n1;
    lowword_msort(b, n1, tmp); lowword_msort(b + n1, n2, tmp);
    size_t i = 0, j = n1, k = 0;
    auto key = [](const double &d) { uint64_t u; memcpy(&u, &d, 8); return (uint32_t)u; };
    while (i < n1 && j < n) {
        if ((int32_t)(key(b[i]) - key(b[j])) <= 0) tmp[k++] = b[i++]; else tmp[k++] = b[j++];
    }
    while (i < n1) tmp[k++] = b[i++];
    memcpy(b, tmp, k * sizeof(double));
}

// the sequential part that is left: hop from seed to seed (src/GROM.c:19370-19389, 19670-19676)
struct Segmenter {
    SegCtx C; const uint32_t *seeds; int64_t lo;
    const uint32_t *wp = nullptr, *land = nullptr; const SeedCall *spec = nullptr;     // optional device-evaluated seeds (k_seed_eval)
    long n_table = 0, n_host = 0, host_span = 0;
    inline int64_t next_seed(int64_t p, int64_t end) const
    {
        if (p >= end) return end;
        int64_t w = p >> 5;
        uint32_t bits = seeds[w] & (0xffffffffu << (p & 31));
        const int64_t wend = (end + 31) >> 5;
        while (!bits) { if (++w >= wend) return end; bits = seeds[w]; }
        const int64_t r = (w << 5) + __builtin_ctz(bits);
        return r < end ? r : end;
    }
    struct Seen { int64_t pos; int mi; };
    struct Piece { std::vector<Seen> seen; std::vector<std::pair<size_t, Call>> calls; int64_t exit_pos = 0; int exit_low = 0; };
    // Hop from (pos, last_low) until pos >= limit.  Every evaluated seed is appended to piece.seen, every call to piece.calls (with the
    // index of its seed).  If `probe` is given, stop as soon as the next seed to evaluate is one the probe path evaluated with the same
    // class: from there on both paths are identical; returns that index, else -1.
    long hop(int64_t pos, int last_low, int64_t limit, Piece &piece, const Piece *probe)
    {
        const int64_t end = C.end;
        if (limit > end) limit = end;
        int mi = 0;
        while (pos < limit) {
            // positions between seeds only move last_low: it is the class of the last covered position the outer loop passed
            const int64_t nx = next_seed(pos, end);
            for (int64_t b = nx - 1; b >= pos; b--) { const int c = C.cls(b); if (c != 2) { last_low = c; break; } }
            pos = nx;
            if (pos >= end) break;
            const int c0 = C.cls(pos);
            if (c0 != 2) { mi = c0; last_low = c0; } else mi = last_low;
            uint32_t v = LAND_NOT - 1;                     // "no table"
            if (land) {
                v = land[2 * ((int64_t)wp[pos >> 5] + __builtin_popcount(seeds[pos >> 5] & ((1u << (pos & 31)) - 1u))) + mi];
                if (v == LAND_NOT) { pos++; continue; }
            } else if (!C.beyond(pos, mi)) { pos++; continue; }
            if (pos >= limit) break;                       // the seed belongs to the next piece
            if (probe) {
                auto it = std::lower_bound(probe->seen.begin(), probe->seen.end(), pos, [](const Seen &s, int64_t x) { return s.pos < x; });
                if (it != probe->seen.end() && it->pos == pos && it->mi == mi) { piece.exit_pos = pos; piece.exit_low = last_low; return (long)(it - probe->seen.begin()); }
            }
            piece.seen.push_back({pos, mi});
            const uint32_t kind = v >> LAND_SHIFT, low = v & ((1u << LAND_SHIFT) - 1u);
            if (land && kind == SEG_RESUME) { pos += low; n_table++; continue; }
            if (land && kind == SEG_CALL) { piece.calls.push_back({piece.seen.size() - 1, Call{pos, spec[low].c_end, spec[low].c_z}}); pos = spec[low].c_end + 2; n_table++; continue; }
            const Outcome o = eval_seed<false>(C, pos, mi);
            n_host++; host_span += o.next - pos;
            if (o.kind == SEG_CALL) piece.calls.push_back({piece.seen.size() - 1, Call{pos, o.c_end, o.c_z}});
            pos = o.next;
        }
        piece.exit_pos = pos; piece.exit_low = last_low;
        return -1;
    }
    // The scan is a chain (each seed decides where the next one is), so pieces of the contig are hopped speculatively in parallel from
    // their own start, then stitched in order: the true path entering a piece is followed until it meets the speculative one.
    void run(std::vector<Call> &out, int n_threads)
    {
        const int64_t end = C.end;
        int K = std::max(1, n_threads);
        if (end - lo < (int64_t)K * 200000) K = (int)std::max<int64_t>(1, (end - lo) / 200000);
        std::vector<Piece> pieces(K);
        std::vector<int64_t> bound(K + 1);
        for (int t = 0; t <= K; t++) bound[t] = lo + (end - lo) * t / K;
        std::vector<Segmenter> workers(K, *this);
        std::vector<std::thread> th;
        for (int t = 0; t < K; t++) {
            auto job = [&, t]() {
                int guess = 0;
                for (int64_t b = bound[t] - 1; b >= lo && b >= bound[t] - 100000; b--) { const int c = C.cls(b); if (c != 2) { guess = c; break; } }
                workers[t].hop(bound[t], t == 0 ? 0 : guess, bound[t + 1], pieces[t], nullptr);
            };
            if (t + 1 < K) th.emplace_back(job); else job();
        }
        for (auto &x : th) x.join();
        for (auto &w : workers) { n_table += w.n_table; n_host += w.n_host; host_span += w.host_span; }
        for (auto &c : pieces[0].calls) out.push_back(c.second);
        int64_t pos = pieces[0].exit_pos; int last_low = pieces[0].exit_low;
        for (int t = 1; t < K; t++) {
            if (pos >= bound[t + 1]) continue;                                   // a call reached past this whole piece
            Piece link;
            const long j = hop(pos, last_low, bound[t + 1], link, &pieces[t]);
            for (auto &c : link.calls) out.push_back(c.second);
            if (j >= 0) {
                for (auto &c : pieces[t].calls) if ((long)c.first >= j) out.push_back(c.second);
                pos = pieces[t].exit_pos; last_low = pieces[t].exit_low;
            } else { pos = link.exit_pos; last_low = link.exit_low; }
        }
    }
};

}  // namespace cnv
