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
    // pair types of the CTA's positions (one evaluation per position) plus one before and a few behind: ty[i] = type of the pair at base - 1 + i
    constexpr int AHEAD = 30;
    __shared__ uint8_t ty[256 + 1 + AHEAD];
    const int64_t base = lo + (int64_t)blockIdx.x * blockDim.x;
    for (int i = threadIdx.x; i < 256 + 1 + AHEAD; i += blockDim.x) {
        const int64_t q = base - 1 + i;
        ty[i] = (q >= lo && q < hi) ? (uint8_t)dinuc_type(fasta[q], fasta[q + 1]) : (uint8_t)(q < lo ? 11 : 12);      // 11 / 12: outside the span, equal to nothing
    }
    __syncthreads();
    const int64_t p = base + threadIdx.x;
    if (p >= hi) return;
    const int t = ty[threadIdx.x + 1];
    if (t == 10) return;
    if (ty[threadIdx.x] == t) return;                                        // not the first pair of its run
    int64_t e = p;
    long long sum = depth[p];
    for (;;) {                                                               // the thread at a run start walks the run
        if (e + 1 >= hi) return;                                             // a run still open at the end of the span is never flushed
        const int64_t i = e + 1 - (base - 1);
        const int tn = i < 256 + 1 + AHEAD ? (int)ty[i] : dinuc_type(fasta[e + 1], fasta[e + 2]);
        if (tn != t) break;
        e++; sum += depth[e];
    }
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

// Both kernels below give each thread CITEMS = 8 consecutive positions of a 2048-position tile (vector loads) and need, per position, the
// class of the LAST "setter" position before it (the reference's ddd_last_low_mq carried through uncovered stretches).  That is an
// exclusive max-scan of setter indices over the tile -- inside the thread, then across the block -- never a backward walk: a tile
// inside a zero-coverage stretch (full-loss segment, N run) costs the same as any other.
constexpr int CITEMS = CTILE / 256;
static_assert(CITEMS == 8, "the carry kernels load 8 positions per thread");
__device__ __forceinline__ int block_excl_max_256(int v, int *s_w /* [8] */)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl = max(incl, y); }
    int excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = -1;
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    int base = -1;
    for (int k = 0; k < w; k++) base = max(base, s_w[k]);
    __syncthreads();
    return max(base, excl);
}
template <class T> __device__ __forceinline__ void load8(const T *__restrict__ a, int64_t p0, int64_t n, T (&v)[8], T fill)
{
    if (p0 + 8 <= n) {
        if (sizeof(T) == 4) { const int4 x = *reinterpret_cast<const int4 *>(a + p0), y = *reinterpret_cast<const int4 *>(a + p0 + 4);
            v[0] = (T)x.x; v[1] = (T)x.y; v[2] = (T)x.z; v[3] = (T)x.w; v[4] = (T)y.x; v[5] = (T)y.y; v[6] = (T)y.z; v[7] = (T)y.w; }
        else { const uint2 x = *reinterpret_cast<const uint2 *>(a + p0);
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = (T)(((k < 4 ? x.x : x.y) >> (8 * (k & 3))) & 0xff); }
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = p0 + k < n ? a[p0 + k] : fill;
    }
}

// ---- K5: mask (src/GROM.c:18681-18720), class, list-size bits; summary of the z-stage setters per tile
__global__ void __launch_bounds__(256) k_mask(const int32_t *__restrict__ depth, const uint8_t *__restrict__ mq8, const int32_t *__restrict__ gc,
                                              const int32_t *__restrict__ acgt, int64_t P, int64_t lo, int64_t hi, int q, const int32_t *__restrict__ nlist,
                                              const uint8_t *__restrict__ tile_in, uint32_t *__restrict__ rec, uint8_t *__restrict__ tile_last_z)
{
    __shared__ int s_w[8];
    __shared__ int best;
    if (threadIdx.x == 0) best = -1;
    const int64_t t0 = (int64_t)blockIdx.x * CTILE, p0 = t0 + (int64_t)threadIdx.x * CITEMS;
    int d[8], g[8], ac[8]; uint8_t m[8];
    load8(depth, p0, P, d, 0); load8(gc, p0, P, g, 0); load8(acgt, p0, P, ac, 0); load8(mq8, p0, P, m, (uint8_t)0);
    // last mask-stage setter (covered, enough ACGT context, inside the analysed span) strictly before each position
    int before[8], last = -1;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int64_t p = p0 + k;
        before[k] = last;
        if (p >= lo && p < hi && ac[k] >= MIN_ACGT && d[k] > 0) last = threadIdx.x * CITEMS + k;
    }
    const int carry = block_excl_max_256(last, s_w);
    const int cin = tile_in[blockIdx.x];
    int mine = -1;
    uint32_t out[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int64_t p = p0 + k;
        const int cls = m[k] >= q ? 0 : (d[k] > 0 ? 1 : 2);
        uint32_t r = (uint32_t)cls << R_CLASS | (uint32_t)m[k] << R_MQ | R_MASK;
        if (p >= lo && p < P) {
            if (nlist[g[k]] > 1) r |= R_WIN0;
            if (nlist[NB + g[k]] > 1) r |= R_WIN1;
        }
        if (p >= lo && p < hi) {
            if (ac[k] >= MIN_ACGT) {
                int mi;
                if (d[k] > 0) mi = m[k] >= q ? 0 : 1;
                else {
                    // uncovered: the list of the last covered position with enough ACGT context (in the tile, else the tile's carry-in)
                    const int idx = max(before[k], carry);
                    mi = idx >= 0 ? (mq8[t0 + idx] >= q ? 0 : 1) : cin;
                }
                if (nlist[mi * NB + g[k]] >= NO_COMBINE) r &= ~R_MASK;
            }
            if (rec_usable(r)) { r |= R_USABLE; if (cls != 2) mine = max(mine, threadIdx.x * CITEMS + k); }
        }
        out[k] = r;
    }
    if (p0 + 8 <= P) { *reinterpret_cast<uint4 *>(rec + p0) = make_uint4(out[0], out[1], out[2], out[3]); *reinterpret_cast<uint4 *>(rec + p0 + 4) = make_uint4(out[4], out[5], out[6], out[7]); }
    else { for (int k = 0; k < 8; k++) if (p0 + k < P) rec[p0 + k] = out[k]; }
    mine = __reduce_max_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31) == 0 && mine >= 0) atomicMax(&best, mine);
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
    __shared__ int s_w[8];
    __shared__ uint8_t s_cls[CTILE];                 // class of every position of the tile (read back at the carried setter index)
    for (int i = threadIdx.x; i < P2S; i += blockDim.x) sp[i] = T.p2s_p[i];
    const int64_t t0 = (int64_t)blockIdx.x * CTILE, p0 = t0 + (int64_t)threadIdx.x * CITEMS;
    int d[8], g[8]; uint32_t r[8];
    load8(depth, p0, P, d, 0); load8(gc, p0, P, g, 0); load8(rec, p0, P, r, (uint32_t)R_MASK);
    int before[8], last = -1;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        before[k] = last;
        const int cls = (r[k] >> R_CLASS) & 3;
        s_cls[threadIdx.x * CITEMS + k] = (uint8_t)cls;
        if ((r[k] & R_USABLE) && cls != 2 && p0 + k < P) last = threadIdx.x * CITEMS + k;
    }
    const int carry = block_excl_max_256(last, s_w);          // (its barriers also publish sp and s_cls)
    const int cin = tile_in[blockIdx.x];
    unsigned bits_del = 0, bits_dup = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int64_t p = p0 + k;
        uint32_t rr = r[k];
        bool s_del = false, s_dup = false;
        if (p < P) {
            const int dd = d[k];
            if (p >= lo) {
                // thresholds of the position's GC bin; past the analysed span the reference reads bin 0 of a zero-filled array
                const int gg = g[k];
                if ((double)dd <= T.del_thr[gg]) rr |= R_DEL0;
                if ((double)dd <= T.del_thr[NB + gg]) rr |= R_DEL1;
                if ((double)dd >= T.dup_thr[gg]) rr |= R_DUP0;
                if ((double)dd >= T.dup_thr[NB + gg]) rr |= R_DUP1;
            }
            if (p >= lo && p < hi) {
                const int gg = g[k];
                const int cls = (rr >> R_CLASS) & 3;
                s_del = cls == 0 ? (rr & R_DEL0) : cls == 1 ? (rr & R_DEL1) : (rr & (R_DEL0 | R_DEL1));
                s_dup = cls == 0 ? (rr & R_DUP0) : cls == 1 ? (rr & R_DUP1) : (rr & (R_DUP0 | R_DUP1));
                if (rr & R_USABLE) {
                    int mi;
                    if (cls == 0) mi = 0;
                    else if (cls == 1) mi = 1;
                    else { const int idx = max(before[k], carry); mi = idx >= 0 ? (int)s_cls[idx] : cin; }
                    const int list = mi * NB + gg, n = T.n[list];
                    if (n > 0) {
                        const double ave = T.ave[list];
                        int i1, i2;
                        bool neg;
                        if ((double)dd < ave) { i1 = rank_le(T, list, dd); i2 = rank_lt(T, list, dd); neg = false; }
                        else {
                            if ((double)dd > 2 * ave) i1 = rank_lt(T, list, (int)(2 * ave)); else i1 = rank_lt(T, list, dd);
                            i2 = rank_le(T, list, dd);
                            i1 = n - i1; i2 = n - i2; neg = true;
                        }
                        const double prob = ((i1 <= 0 ? 0.5 : (double)i1) + (i2 <= 0 ? 0.5 : (double)i2)) / (double)(2 * (long long)n);
                        int a = 0, b = P2S;                       // first table entry > prob (bisect_right over the ascending p-value table)
                        while (a < b) { const int mm = (a + b) >> 1; if (prob < sp[mm]) b = mm; else a = mm + 1; }
                        if (a >= P2S) a = P2S - 1;
                        rr |= R_NZ | (neg ? R_NEG : 0u) | ((uint32_t)a << R_K);
                    }
                }
            }
        }
        r[k] = rr;
        if (s_del) bits_del |= 1u << k;
        if (s_dup) bits_dup |= 1u << k;
    }
    if (p0 + 8 <= P) { *reinterpret_cast<uint4 *>(rec + p0) = make_uint4(r[0], r[1], r[2], r[3]); *reinterpret_cast<uint4 *>(rec + p0 + 4) = make_uint4(r[4], r[5], r[6], r[7]); }
    else { for (int k = 0; k < 8; k++) if (p0 + k < P) rec[p0 + k] = r[k]; }
    // seed words: four threads (32 positions) share one word of each bitmap
    const int lane = threadIdx.x & 31, sub = lane & 3;
    unsigned wd = bits_del << (8 * sub), wu = bits_dup << (8 * sub);
    wd |= __shfl_xor_sync(0xffffffffu, wd, 1); wd |= __shfl_xor_sync(0xffffffffu, wd, 2);
    wu |= __shfl_xor_sync(0xffffffffu, wu, 1); wu |= __shfl_xor_sync(0xffffffffu, wu, 2);
    if (sub == 0 && p0 < ((P + 31) / 32) * 32) { seed_del[p0 >> 5] = wd; seed_dup[p0 >> 5] = wu; }
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
// A CTA owns 32 consecutive frames and walks them in chunks of SW_CH elements through a double-buffered shared-memory tile
// [32 frames][SW_CH]:
//  (A) all eight warps fetch the records row by row (a row = SW_CH consecutive elements of one frame's walk: coalesced), decode z and park it;
//  (B) warp 0, one lane per frame, folds ITS frame's SW_CH values into its running sum in element order -- one dependent DADD per
//      element, the reference's summation order -- leaving the prefix sums in place;
//  (C) all warps read the tile back transposed (lane = window length): division, square and the coalesced store of X[frame][L].
// (B) of chunk c overlaps (A) of chunk c + 1 (the other buffer), so the serial chain hides behind the fetches.
#define SW_CH 64
#define SW_PITCH (SW_CH + 1)
#define SW_WARPS 8
__global__ void __launch_bounds__(32 * SW_WARPS) k_sweep(const uint32_t *__restrict__ rec, const SweepBlock *__restrict__ blocks, int n_blocks, int64_t n_frames,
                                              int A, int Lmin, int Lmax, int q, const double *__restrict__ p2s_sd, const double *__restrict__ wtab_g, double *__restrict__ X)
{
    __shared__ double sd[P2S];
    __shared__ double wtab[256];
    __shared__ double zt[2][32 * SW_PITCH];
    __shared__ int st_a[2][32], st_p[2][32], st_s[32], st_e[32], st_n0[2][32];
    __shared__ unsigned st_um[2][32][SW_CH / 32], st_vm[2][32][SW_CH / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < P2S; i += blockDim.x) sd[i] = p2s_sd[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) wtab[i] = wtab_g[i];
    const int64_t f0 = (int64_t)blockIdx.x * 32;
    const int n_rows = (int)min((int64_t)32, n_frames - f0);
    const int n_len = Lmax - Lmin + 1;
    const int n_chunks = (Lmax + SW_CH - 1) / SW_CH;
    // warp 0: lane = frame; walk state at the frame start
    int a = A; int64_t p = 0, s = 0, e = 0;
    double tot = 0.0;
    int n_us = 0;
    if (wid == 0) {
        const int64_t f = f0 + lane;
        if (f < n_frames) {
            int lo_b = 0, hi_b = n_blocks - 1;                   // last block whose first frame is <= f
            while (lo_b < hi_b) { const int m = (lo_b + hi_b + 1) >> 1; if (blocks[m].first_frame <= f) lo_b = m; else hi_b = m - 1; }
            s = blocks[lo_b].start; e = blocks[lo_b].end; a = 0; p = s;
            sweep_advance(a, p, (f - blocks[lo_b].first_frame) * (int64_t)Lmax, s, e, A, Lmax);
        }
        st_s[lane] = (int)s; st_e[lane] = (int)e; st_a[0][lane] = a; st_p[0][lane] = (int)p;
    }
    __syncthreads();
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    auto fetch = [&](int c, int w_first, int w_step) {            // (A) chunk c -> buffer c & 1, rows split over warps w_first, w_first + w_step ...
        const int b = c & 1, w0 = c * SW_CH;
        for (int r = w_first; r < 32; r += w_step) {
            const int ra = st_a[b][r]; const int64_t rp = st_p[b][r], rs = st_s[r], re = st_e[r];
#pragma unroll
            for (int h = 0; h < SW_CH / 32; h++) {
                const int j = h * 32 + lane;
                int la = ra; int64_t lp = rp;
                sweep_advance(la, lp, j, rs, re, A, Lmax);
                const bool valid = la < A && w0 + j < Lmax;
                const uint32_t rr = valid ? __ldg(rec + lp) : R_MASK;
                const bool us = valid && rec_usable(rr);
                double z = 0.0;
                if (us && (rr & R_NZ)) {
                    const double wt = (rr & R_OVR) ? 1.0 : (((rr >> R_CLASS) & 3) == 0 ? wtab[(rr >> R_MQ) & 255] : 0.5);
                    z = __dmul_rn(wt, sd[(rr >> R_K) & 1023]);
                    if (rr & R_NEG) z = -z;
                }
                const unsigned um = __ballot_sync(0xffffffffu, us), vm = __ballot_sync(0xffffffffu, valid);
                zt[b][r * SW_PITCH + j] = z;
                if (lane == 0) { st_um[b][r][h] = um; st_vm[b][r][h] = vm; }
            }
        }
    };
    auto next_state = [&](int c) {                                // warp 0: walk state at the start of chunk c (the lane's state is chunk c - 1's)
        sweep_advance(a, p, SW_CH, s, e, A, Lmax);
        st_a[c & 1][lane] = a; st_p[c & 1][lane] = (int)p;
    };
    fetch(0, wid, SW_WARPS);
    if (wid == 0) next_state(1);
    __syncthreads();
    for (int c = 0; c < n_chunks; c++) {
        const int b = c & 1, w0 = c * SW_CH;
        if (wid == 0) {                                           // (B) the ordered fold of chunk c ...
            double *mine = zt[b] + lane * SW_PITCH;
            st_n0[b][lane] = n_us;
#pragma unroll
            for (int h = 0; h < SW_CH / 32; h++) {
                const unsigned um = st_um[b][lane][h];
#pragma unroll
                for (int j = 0; j < 32; j++) { if ((um >> j) & 1u) tot = __dadd_rn(tot, mine[h * 32 + j]); mine[h * 32 + j] = tot; }
                n_us += __popc(um);
            }
        } else if (c + 1 < n_chunks) fetch(c + 1, wid - 1, SW_WARPS - 1);      // ... while the other warps fetch chunk c + 1 into the other buffer
        __syncthreads();
        if (wid == 0 && c + 2 < n_chunks) next_state(c + 2);      // slot (c + 2) & 1 == b: its last readers (fetch of chunk c) are long done
        {                                                         // (C) lane = window length
#pragma unroll
            for (int h = 0; h < SW_CH / 32; h++) {
                const int j = h * 32 + lane, w = w0 + j + 1;
                if (w >= Lmin && w <= Lmax) {
                    const unsigned below = 0xffffffffu >> (31 - lane);
                    for (int r = wid; r < n_rows; r += SW_WARPS) {
                        int n = st_n0[b][r];
                        for (int hh = 0; hh < h; hh++) n += __popc(st_um[b][r][hh]);
                        n += __popc(st_um[b][r][h] & below);
                        double x2 = nan;
                        if (((st_vm[b][r][h] >> lane) & 1u) && n > 0) { const double x = zt[b][r * SW_PITCH + j] / (double)n; x2 = __dmul_rn(x, x); }
                        __stcs(X + (f0 + r) * (int64_t)n_len + (w - Lmin), x2);
                    }
                }
            }
        }
        __syncthreads();                                          // tile b is drained: the next iteration's fetch (chunk c + 2) may overwrite it
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
    int64_t slide_max = 0;       // bounded evaluation: positions the sliding phase may advance before the seed is left open (0: never enters it)
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
// z of the deletion scan per position, unpacked once from the packed records (same two factors, same __dmul_rn as z_del_of): the seed
// evaluations then read 8 coalesced bytes per position instead of gathering from the p-value and MAPQ-weight tables lane by lane
__global__ void __launch_bounds__(256) k_zfill(const uint32_t *__restrict__ rec, int64_t P, const double *__restrict__ sd, const double *__restrict__ wtab, double *__restrict__ z)
{
    __shared__ double s_sd[1024], s_w[256];
    for (int i = threadIdx.x; i < 1024; i += 256) s_sd[i] = i < P2S ? sd[i] : 0.0;
    s_w[threadIdx.x] = wtab[threadIdx.x];
    __syncthreads();
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < P; p += (int64_t)gridDim.x * 256) {
        const uint32_t r = rec[p];
        double v = 0.0;
        if (r & R_NZ) {
            const double w = (r & R_OVR) ? 1.0 : (((r >> R_CLASS) & 3) == 0 ? s_w[(r >> R_MQ) & 255] : 0.5);
            v = __dmul_rn(w, s_sd[(r >> R_K) & 1023]);
            if (r & R_NEG) v = -v;
        }
        z[p] = v;
    }
}
enum { SEG_RESUME = 0, SEG_CALL = 1, SEG_UNRESOLVED = 2 };
struct Outcome { int kind; int64_t next, c_end; double c_z; int64_t far; };   // far: one past the last position the sliding phase looked at
constexpr int SEED_BOUND0 = 128;       // first round, pass one (every seed)
constexpr bool MID_EVAL = true;      // evaluate the seeds that ran past the first bound once more at SEED_BOUND before calling them open (off: straight to the run heads + second round)
constexpr int SEED_BOUND = 1024;       // first round, every seed: closes everything but genuine events and long uncovered stretches (chance dips of the coverage end within a few hundred positions); what runs past it is "open"

template <bool BOUNDED, bool DUP, bool HASZ> __host__ __device__ inline Outcome eval_seed_k(const SegCtx &C, const int64_t pos, int mi)
{
#define CNV_STEP(var, p) do { const int c_ = C.cls(p); if (c_ != 2) var = c_; } while (0)
    const int64_t Lmin = C.Lmin, Lmax = C.Lmax, end = C.end, max_gap = Lmax + 500;
    Outcome o; o.kind = SEG_RESUME; o.next = pos + 1; o.c_end = 0; o.c_z = 0; o.far = 0;
    bool stop = false, begun = false;
    int wlen = 0, cnt = 0, cnt2 = 0;                        // window length and counters fit 32 bits (Lmax positions at most)
    int64_t pa, c_start = 0, c_end = 0, last_good = 0;
    double tot = 0, c_z = 0, tz;
    const uint32_t *rp = C.rec + pos;                        // the first two phases index relative to the seed (32-bit offsets)
    const double *zp = HASZ ? C.zarr + pos : nullptr;         // HASZ: z unpacked per position (device); otherwise derived from the record (host)
    const uint32_t m0 = DUP ? R_DUP0 : R_DEL0, m1 = DUP ? R_DUP1 : R_DEL1;
    // z as the scan sees it: 0.0 - v for duplications (exact, and keeps a zero positive like the reference's literal 0.0)
    auto zrel = [&](int i, uint32_t r) { const double v = HASZ ? zp[i] : C.z_del_of(r); return DUP ? 0.0 - v : v; };
    const int iLmin = (int)Lmin, iLmax = (int)Lmax;
    cnt = iLmin;
    if (HASZ) {
        // one pass: the give-up test of the first window, and on the side the window's sum and mask count (the reference walks twice;
        // the sum adds the same values in the same order, and is simply dropped when the walk gives up)
        int w2 = 0;
        for (int i = 0; i < iLmin; i++) {
            const uint32_t r = rp[i];
            const bool um = !(r & R_MASK);
            const int c_ = (int)(r >> R_CLASS) & 3;
            mi = (um && c_ != 2) ? c_ : mi;
            tot += zrel(i, r);
            cnt -= (int)(r & R_MASK);
            const bool ok = um && (r & (mi ? m1 : m0)) != 0;
            w2 += ok ? 1 : -1;
            if (!ok && w2 < 0) { o.next = pos + i + 1; return o; }          // give up inside the first window: resume after the offender
        }
        cnt2 = (w2 + iLmin) / 2; wlen = iLmin;
    } else {
    for (int i = 0; i < iLmin; i++) {
        wlen++;
        bool ok = false;
        const uint32_t r = rp[i];
        if (!(r & R_MASK)) { const int c_ = (r >> R_CLASS) & 3; if (c_ != 2) mi = c_; ok = (r & (mi ? m1 : m0)) != 0; }
        if (ok) cnt2++;
        else if (2 * cnt2 < wlen) { o.next = pos + i + 1; return o; }       // give up inside the first window: resume after the offender
    }
    for (int i = 0; i < iLmin; i++) { const uint32_t r = rp[i]; cnt -= (int)(r & R_MASK); tot += zrel(i, r); }
    }
    if (cnt > 0 && C.scores(tot, cnt, iLmin, &tz)) {
        begun = true; c_start = pos; last_good = c_end = pos + Lmin; c_z = tz;
    }
    {
        const int i_end = end - pos < (int64_t)iLmax ? (int)(end - pos) : iLmax;      // first offset at or past `end` (>= Lmin is not guaranteed)
        int good = -1;                                                               // last scoring offset of this phase
        // the loop runs up to the first of: the largest window, the bound of a bounded evaluation, the end of the analysed block
        int i_lim = i_end < iLmax ? i_end : iLmax;
        if (BOUNDED && C.bound < i_lim) i_lim = C.bound;
        const double *thr = C.win_thr;
        double dcnt = (double)cnt;                                                   // cnt as a double, kept in step (exact: small integers)
        int i = iLmin;
        if (HASZ) {
            // device form (z unpacked per position): the same steps with selects instead of nested branches -- a masked position adds an
            // exact zero (z is 0 where the mask is set, and x + (+-0) == x), leaves class and counters alone and is never "ok"; the two
            // rare events (a window that may score, giving up) stay branches
            int w2 = 2 * cnt2 - i;                                                   // 2 * cnt2 - wlen; every step adds +-1
            // one step; `ii` is the offset of the position, `r` / `zv` its record and z
#define CNV_GROW_STEP(ii, r, zv)                                                                                          \
            {                                                                                                            \
                const bool um = !((r) & R_MASK);                                                                         \
                const int c_ = (int)((r) >> R_CLASS) & 3;                                                                \
                mi = (um && c_ != 2) ? c_ : mi;                                                                          \
                tot += DUP ? 0.0 - (zv) : (zv);                                                                          \
                dcnt += um ? 1.0 : 0.0;                                                                                  \
                const bool ok = um && ((r) & (mi ? m1 : m0)) != 0;                                                       \
                w2 += ok ? 1 : -1;                                                                                       \
                if (ok) {                                                                                                \
                    if (tot >= dcnt * thr[(ii) + 1]) {                       /* SegCtx::scores, its cheap side inline */ \
                        tz = tot / (dcnt * C.win_sd[(ii) + 1]);                                                          \
                        if (tz >= 3) { good = (ii); if (tz > c_z) c_z = tz; }                                            \
                    }                                                                                                    \
                } else if (w2 < 0) { stop = true; i = (ii); break; }                                                     \
            }
            // four positions per trip: their records and z values are fetched together (the walk is a dependent chain per seed; the
            // loads are not), then folded in order
            for (; i + 4 <= i_lim; i += 4) {
                const uint32_t r0 = rp[i], r1 = rp[i + 1], r2 = rp[i + 2], r3 = rp[i + 3];
                const double z0 = zp[i], z1 = zp[i + 1], z2 = zp[i + 2], z3 = zp[i + 3];
                CNV_GROW_STEP(i, r0, z0) CNV_GROW_STEP(i + 1, r1, z1) CNV_GROW_STEP(i + 2, r2, z2) CNV_GROW_STEP(i + 3, r3, z3)
            }
            if (!stop) for (; i < i_lim; i++) { const uint32_t r0 = rp[i]; const double z0 = zp[i]; CNV_GROW_STEP(i, r0, z0) }
#undef CNV_GROW_STEP
            cnt = (int)dcnt;
        } else
        for (; i < i_lim; i++) {
            const uint32_t r = rp[i];
            bool ok = false;
            if (!(r & R_MASK)) {
                const int c_ = (r >> R_CLASS) & 3; if (c_ != 2) mi = c_;
                tot += zrel(i, r); cnt++; dcnt += 1.0;
                ok = (r & (mi ? m1 : m0)) != 0;
                if (ok) {
                    cnt2++;
                    if (tot >= dcnt * thr[i + 1]) {                                  // SegCtx::scores, its cheap side inline
                        tz = tot / (dcnt * C.win_sd[i + 1]);
                        if (tz >= 3) { good = i; if (tz > c_z) c_z = tz; }           // c_z starts at 0 and every scoring tz is >= 3
                    }
                }
            }
            if (!ok && 2 * cnt2 < i + 1) { stop = true; break; }
        }
        if (!stop && i < iLmax) {
            // left the loop early: the bound is tested before the block's end, like the per-step order of the tests it replaces
            if (BOUNDED && i >= C.bound) { o.kind = SEG_UNRESOLVED; return o; }
            stop = true;
        }
        wlen = i + 1;
        if (good >= 0) {
            last_good = pos + good;
            if (!begun) { begun = true; c_start = pos; }
            c_end = pos + good;
        }
    }
    if (!stop && begun) {
        if (BOUNDED && C.slide_max <= 0) { o.kind = SEG_UNRESOLVED; return o; }
        int mi_b = mi;
        pa = pos + Lmax; tot = 0; cnt = 0;
        while (pa < C.len && pa - last_good <= max_gap) {
            if (BOUNDED && pa - (pos + Lmax) > C.slide_max) { o.kind = SEG_UNRESOLVED; return o; }
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
template <bool BOUNDED> __host__ __device__ inline Outcome eval_seed(const SegCtx &C, const int64_t pos, int mi)
{
#ifdef __CUDA_ARCH__
    return C.dup ? eval_seed_k<BOUNDED, true, true>(C, pos, mi) : eval_seed_k<BOUNDED, false, true>(C, pos, mi);       // device contexts always carry zarr (k_zfill)
#else
    return C.dup ? eval_seed_k<BOUNDED, true, false>(C, pos, mi) : eval_seed_k<BOUNDED, false, false>(C, pos, mi);     // host contexts never do
#endif
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

// positions [a, b] of a call made at a head -> the `cover` bitmap of its kind (seeds under it are jumped over by the path)
__device__ __forceinline__ void cover_mark(uint32_t *__restrict__ cov, int64_t a, int64_t b, int lane, int n_lanes)
{
    if (b < a) return;
    const int64_t w0 = a >> 5, w1 = b >> 5;
    for (int64_t w = w0 + lane; w <= w1; w += n_lanes) {
        uint32_t m = 0xffffffffu;
        if (w == w0) m &= 0xffffffffu << (a & 31);
        if (w == w1) m &= 0xffffffffu >> (31 - (b & 31));
        if (m == 0xffffffffu) cov[w] = m; else atomicOr(cov + w, m);
    }
}
__device__ __forceinline__ void head_publish(const SegCtx &C, int kind, uint32_t rank, int variant, int64_t p, uint32_t e, const uint32_t *__restrict__ seeds, int64_t words,
                                             const uint32_t *__restrict__ wp, uint32_t *__restrict__ land, uint32_t cap, const SeedCall *__restrict__ calls,
                                             uint32_t n_del, uint32_t n_dup, uint32_t *__restrict__ jump0)
{
    const int c0 = C.cls(p);
    land[((int64_t)kind * cap + rank) * 2 + variant] = e;
    if (c0 != 2) land[((int64_t)kind * cap + rank) * 2 + 1] = e;
    if (jump0) {
        const uint32_t n_seeds = kind ? n_dup : n_del, base = kind ? 2 * n_del + 1 : 0;
        const uint32_t *sd = seeds + (int64_t)kind * words, *wpk = wp + (int64_t)kind * words;
        jump0[base + 2 * rank + variant] = base + successor(C, sd, wpk, p, c0, variant, e, calls, rank, n_seeds);
        if (c0 != 2) jump0[base + 2 * rank + 1] = base + successor(C, sd, wpk, p, c0, 1, e, calls, rank, n_seeds);
    }
}
// bitmap of the positions that carry a z value (R_NZ), with per-word ranks (k_seed_rank): "does this stretch hold any z at all" in O(1)
__global__ void __launch_bounds__(256) k_nz_bits(const uint32_t *__restrict__ rec, int64_t P, uint32_t *__restrict__ nz)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned b = __ballot_sync(0xffffffffu, p < P && (rec[p] & R_NZ));
    if ((threadIdx.x & 31) == 0 && p < ((P + 31) / 32) * 32) nz[p >> 5] = b;
}
__device__ __forceinline__ uint32_t nz_rank(const uint32_t *__restrict__ nz, const uint32_t *__restrict__ nzwp, int64_t words, int64_t x)
{
    const int64_t w = x >> 5;
    if (w >= words) return nzwp[words - 1] + __popc(nz[words - 1]);
    return nzwp[w] + __popc(nz[w] & ((1u << (x & 31)) - 1u));
}

// ---- tails of stretches without z values.  A seed s inside such a stretch that ends at b (the first position with a z value) sees a
// running sum of exactly 0 up to b and the sums T_b(x) = z(b) + .. + z(x) behind it, whatever s is; only its counters differ.  Having
// passed the first window it resumes at s + 1 unless some window scores, and a score at x needs T_b(x) >= cnt * win_thr[wlen] with
// wlen = (x - b + 1) + (b - s) and cnt >= wlen - (masked positions in reach).  k_tail_check tests that necessary condition once per
// stretch end, for every x and the most favourable window length (suffix minimum `gm` of L * win_thr[L]), on an approximate T with a
// margin far above any rounding difference to the reference's sum; where it can never hold, every seed at least TAIL_KMIN positions
// before b is closed in O(1) (the reference walks up to 2 * (b - s) positions from each of them: src/GROM.c:19402-19470).
constexpr int TAIL_KMIN = 64;
__global__ void __launch_bounds__(256) k_tail_ends(const uint32_t *__restrict__ nz, int64_t words, int32_t *__restrict__ ends, uint32_t cap, unsigned int *__restrict__ n_ends)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < 2 || w >= words) return;
    const uint32_t m = nz[w];
    if (!m || nz[w - 1] || nz[w - 2]) return;                 // the lowest z position of the word, behind at least 64 positions without one
    const unsigned int k = atomicAdd(n_ends, 1u);
    if (k < cap) ends[k] = (int32_t)((w << 5) + __ffs(m) - 1);
}
__global__ void __launch_bounds__(128) k_tail_check(const uint32_t *__restrict__ rec, const double *__restrict__ z, int64_t P, const int32_t *__restrict__ ends,
                                                    const unsigned int *__restrict__ n_ends_p, uint32_t cap, int Lmax, const double *__restrict__ gm,
                                                    uint32_t *__restrict__ safe, int64_t words)
{
    const int lane = threadIdx.x & 31;
    const uint32_t n_ends = min(*n_ends_p, cap);
    for (uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < n_ends; e += (gridDim.x * blockDim.x) >> 5) {     // one warp per stretch end
    const int64_t b = ends[e];
    int Ms = 0;                                               // masked positions a seed of this tail can have before b
    for (int64_t p = max((int64_t)0, b - Lmax) + lane; p < b; p += 32) Ms += (int)(rec[p] & R_MASK);
    for (int o = 16; o; o >>= 1) Ms += __shfl_xor_sync(0xffffffffu, Ms, o);
    double T = 0, A = 0;
    int Mb = 0;
    bool bad0 = false, bad1 = false;
    const int64_t limit = min(P, b + Lmax - TAIL_KMIN);
    for (int64_t x0 = b; x0 < limit; x0 += 32) {
        const int64_t x = x0 + lane;
        const bool valid = x < limit;
        const uint32_t r = valid ? rec[x] : 0u;
        int sm = (int)(r & R_MASK);
        double sv = (valid && !sm) ? z[x] : 0.0, sa = fabs(sv);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double yv = __shfl_up_sync(0xffffffffu, sv, o), ya = __shfl_up_sync(0xffffffffu, sa, o);
            const int ym = __shfl_up_sync(0xffffffffu, sm, o);
            if (lane >= o) { sv += yv; sa += ya; sm += ym; }
        }
        if (valid) {
            const double Tx = T + sv, margin = 1e-10 * (A + sa) + 1e-300;
            const int Mx = Ms + Mb + sm, L = (int)(x - b) + 1 + TAIL_KMIN;
            if (2 * Mx >= L) bad0 = bad1 = true;
            else {
                const double bound = gm[L] * (1.0 - (double)Mx / (double)L);
                if (Tx + margin >= bound) bad0 = true;
                if (margin - Tx >= bound) bad1 = true;
            }
        }
        T += __shfl_sync(0xffffffffu, sv, 31); A += __shfl_sync(0xffffffffu, sa, 31); Mb += __shfl_sync(0xffffffffu, sm, 31);
        if (__any_sync(0xffffffffu, bad0) && __any_sync(0xffffffffu, bad1)) break;
    }
    bad0 = __any_sync(0xffffffffu, bad0); bad1 = __any_sync(0xffffffffu, bad1);
    if (lane == 0) {
        if (!bad0) atomicOr(safe + (b >> 5), 1u << (b & 31));
        if (!bad1) atomicOr(safe + words + (b >> 5), 1u << (b & 31));
    }
    }
}
// first position with a z value at or after p, looking no further than word wmax; -1 if there is none
__device__ __forceinline__ int64_t nz_next(const uint32_t *__restrict__ nz, const uint32_t *__restrict__ nzwp, int64_t words, int64_t p, int64_t wmax)
{
    const int64_t w = p >> 5;
    if (w >= words) return -1;
    const uint32_t r = nzwp[w] + __popc(nz[w] & ((1u << (p & 31)) - 1u));
    int64_t lo = w, hi = min(words - 1, wmax);
    if (nzwp[hi] + __popc(nz[hi]) <= r) return -1;
    while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (nzwp[m] + __popc(nz[m]) > r) hi = m; else lo = m + 1; }
    uint32_t bits = nz[lo];
    if (lo == w) bits &= ~((1u << (p & 31)) - 1u);
    return (lo << 5) + __ffs(bits) - 1;
}

// Seed evaluation, first round, in two passes.  Pass one, every seed, bounded to SEED_BOUND0 positions (the first window and a little
// more: almost every seed gives up after a handful of positions): a CTA owns 256 words of the seed bitmap of one kind (blockIdx.y =
// deletions / duplications), the set bits are compacted into shared memory, so consecutive threads take consecutive seeds and a seed's
// rank is the block's first rank + its index (no search).  Each seed is evaluated under its own class, or under both carried classes
// when it is uncovered; outcomes go to `land`, successors to level 0 of the jump table.  Seeds that ran past the bound -- they sit
// together inside events, whole CTAs of them -- go to a list, and pass two (k_seed_eval_mid, bound SEED_BOUND) takes that list one
// thread per entry, spread evenly over the device; what runs past that bound too is "open": `todo` list + open bitmaps [kind][class].
struct SeedTodo { uint32_t rank; uint8_t kind, variant; uint16_t pad; int32_t pos; };
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
constexpr int SEED_CTA_WORDS = 256;
__global__ void __launch_bounds__(256, 4) k_seed_eval(SegCtx Cdel, SegCtx Cdup, const uint32_t *__restrict__ seeds, int64_t words, const uint32_t *__restrict__ wp,
                                                   uint32_t *__restrict__ land, uint32_t cap, uint32_t n_del, uint32_t n_dup, SeedCall *__restrict__ calls, uint32_t call_cap,
                                                   unsigned int *__restrict__ n_calls, SeedTodo *__restrict__ mid, uint32_t mid_cap, uint32_t *__restrict__ jump0, uint32_t *__restrict__ u1)
{
    __shared__ uint16_t lst[SEED_CTA_WORDS * 32];
    __shared__ uint32_t s_warp[8];
    const int kind = blockIdx.y;
    const SegCtx &C = kind ? Cdup : Cdel;
    const uint32_t *sd = seeds + (int64_t)kind * words, *wpk = wp + (int64_t)kind * words;
    const int64_t w0 = (int64_t)blockIdx.x * SEED_CTA_WORDS, w = w0 + threadIdx.x;
    const uint32_t bits = w < words ? sd[w] : 0u;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = __popc(bits);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    uint32_t base = 0, total = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { if (k < wid) base += s_warp[k]; total += s_warp[k]; }
    {
        uint32_t o = base + incl - __popc(bits), b = bits;
        while (b) { const int bit = __ffs(b) - 1; b &= b - 1; lst[o++] = (uint16_t)((threadIdx.x << 5) | bit); }
    }
    __syncthreads();
    if (!total) return;
    const uint32_t rank0 = wpk[w0], n_seeds = kind ? n_dup : n_del, jbase = kind ? 2 * n_del + 1 : 0;
    const uint32_t unres = (uint32_t)SEG_UNRESOLVED << LAND_SHIFT;
    for (uint32_t idx0 = 0; idx0 < total; idx0 += blockDim.x) {               // uniform trip count: the appends below are warp-wide
        const uint32_t idx = idx0 + threadIdx.x, rank = rank0 + idx;
        const bool valid = idx < total && rank < cap;
        uint32_t res[2] = {LAND_NOT, LAND_NOT};
        int c0 = 0;
        int64_t p = 0;
        if (valid) {
            const uint32_t loc = lst[idx];
            p = ((w0 + (loc >> 5)) << 5) + (loc & 31);
            if (p < C.end) {
                c0 = C.cls(p);
                for (int v = 0; v < 2; v++) {
                    if (c0 != 2 && v == 1) { res[1] = res[0]; break; }
                    const int mi = c0 == 2 ? v : c0;
                    if (!C.beyond(p, mi)) continue;
                    res[v] = seed_outcome(C, p, mi, calls, call_cap, n_calls);
                }
            }
        }
        // past the short bound of this pass: queued for the pass over such seeds (k_seed_eval_mid); a warp appends its seeds in
        // lane order -- consecutive seeds -- so that pass finds neighbouring seeds in neighbouring lanes
        for (int v = 0; v < 2; v++) {
            const bool q = valid && res[v] == unres && (v == 0 || c0 == 2);
            const unsigned m = __ballot_sync(0xffffffffu, q);
            if (!m) continue;
            unsigned int k0 = 0;
            if (lane == __ffs(m) - 1) k0 = atomicAdd(n_calls + 8, (unsigned int)__popc(m));
            k0 = __shfl_sync(0xffffffffu, k0, __ffs(m) - 1);
            if (q) {
                const unsigned int k = k0 + __popc(m & ((1u << lane) - 1u));
                if (k < mid_cap) { SeedTodo t; t.rank = rank; t.kind = (uint8_t)kind; t.variant = (uint8_t)v; t.pad = 0; t.pos = (int32_t)p; mid[k] = t; }
                atomicOr(u1 + (int64_t)kind * words + (p >> 5), 1u << (p & 31));
            }
        }
        if (!valid) continue;
        land[((int64_t)kind * cap + rank) * 2] = res[0]; land[((int64_t)kind * cap + rank) * 2 + 1] = res[1];
        if (jump0) {
            // successors; node ids are local to the kind, the table stores them behind the kind's base offset (2 * n_del + 1 for duplications)
            for (int v = 0; v < 2; v++)
                jump0[jbase + 2 * rank + v] = jbase + (p < C.end ? successor(C, sd, wpk, p, c0, v, res[v], calls, rank, n_seeds) : 2 * n_seeds);
        }
    }
    if (jump0 && blockIdx.x == 0 && threadIdx.x == 0) jump0[jbase + 2 * n_seeds] = jbase + 2 * n_seeds;      // END loops on itself
}

// Pass two over the seeds that ran past the first bound, in two launches (PHASE 0, then PHASE 1):
//  * closed in O(1) where nothing can score: no z value within reach, or the tail of a stretch without z values whose end k_tail_check
//    found safe;
//  * a seed whose right-hand neighbour is also on the list is deferred to phase 1: consecutive seeds give up in order of position (the
//    give-up rule compares the count of positions beyond the threshold with the window length, and a seed further left has at least
//    the lead of the one to its right), so if the last seed of such a run stays open the others are left open without a walk -- which
//    is always safe, an open seed is evaluated exactly if the path reaches it -- and only runs whose last seed closed are walked in full.
//    Inside an event that is the walk of one seed per gap instead of every position;
//  * everything else is walked to SEED_BOUND.
// One thread per seed, each thread on its own grid-stride loop and no warp-wide step inside it: the seeds of a warp stop after very
// different numbers of positions (the give-up rule is a first-passage time), and a lane that is done moves on to its next seed while
// its neighbours still walk -- lanes that meet again in the walk's loop are issued together.
template <int PHASE>
__global__ void __launch_bounds__(128, 8) k_seed_eval_mid(SegCtx Cdel, SegCtx Cdup, const uint32_t *__restrict__ seeds, int64_t words, const uint32_t *__restrict__ wp,
                                                       uint32_t *__restrict__ land, uint32_t cap, uint32_t n_del, uint32_t n_dup, SeedCall *__restrict__ calls, uint32_t call_cap,
                                                       unsigned int *__restrict__ n_calls, SeedTodo *__restrict__ mid, uint32_t mid_cap, SeedTodo *__restrict__ todo, uint32_t todo_cap,
                                                       uint32_t *__restrict__ jump0, uint32_t *__restrict__ open_bits, const uint32_t *__restrict__ nz, const uint32_t *__restrict__ nzwp,
                                                       const uint32_t *__restrict__ u1, const uint32_t *__restrict__ safe, uint32_t *__restrict__ run_open, SeedTodo *__restrict__ wl)
{
    // PHASE 0 classifies the list, PHASE 2 walks the run ends it set aside (list `wl`, compact: full warps), PHASE 1 the deferred seeds
    const uint32_t n_mid = PHASE == 2 ? min(n_calls[13], mid_cap) : min(n_calls[8], mid_cap);
    const uint32_t unres = (uint32_t)SEG_UNRESOLVED << LAND_SHIFT;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_mid; i += gridDim.x * blockDim.x) {
        const SeedTodo t = PHASE == 2 ? wl[i] : mid[i];
        const SegCtx &C = t.kind ? Cdup : Cdel;
        const int64_t p = t.pos;
        const int c0 = C.cls(p);
        uint32_t e = unres;
        bool walk = true;
        if (PHASE == 0) {
            // It passed the first window.  If no position of everything the growth phase can reach carries a z value (an uncovered
            // stretch: full-copy loss, the start of the contig), the running sum stays exactly 0, no window length can score and
            // the reference resumes at the next position after walking all of it (src/GROM.c:19402-19470): O(1) here, before any walk.
            const int64_t reach = min(C.len, p + max((int64_t)C.Lmin, min((int64_t)C.Lmax, C.end - p)));
            if (nz_rank(nz, nzwp, words, reach) == nz_rank(nz, nzwp, words, p)) { e = 1u; atomicAdd(n_calls + 7, 1u); }
            else {
                const int64_t w = p >> 5;
                if (!(nz[w] >> (p & 31)) && (w + 1 >= words || !nz[w + 1])) {                 // nothing within the next 32 positions: look for the end of the stretch
                    const int64_t f = nz_next(nz, nzwp, words, p, (p + C.Lmax) >> 5);
                    if (f >= 0 && f - p >= TAIL_KMIN && ((safe[(int64_t)t.kind * words + (f >> 5)] >> (f & 31)) & 1u)) { e = 1u; atomicAdd(n_calls + 10, 1u); }
                }
            }
            if (e == unres && ((u1[(int64_t)t.kind * words + ((p + 1) >> 5)] >> ((p + 1) & 31)) & 1u)) { mid[i].pad = 1; continue; }      // deferred
            if (e == unres && MID_EVAL) { const unsigned int k = atomicAdd(n_calls + 13, 1u); if (k < mid_cap) { wl[k] = t; continue; } }         // a run end: walked by the next launch
        } else if (PHASE == 1) {
            if (!t.pad) continue;                                                              // dealt with in phase 0
            // the last seed of the run of listed seeds this one sits in stayed open in phase 0: open as well, no walk
            if ((run_open[((int64_t)(t.kind * 2 + t.variant)) * words + (p >> 5)] >> (p & 31)) & 1u) { walk = false; atomicAdd(n_calls + 11, 1u); }
        }
        if (PHASE != 0 && e == unres && walk && MID_EVAL) e = seed_outcome(C, p, c0 == 2 ? t.variant : c0, calls, call_cap, n_calls);
        if (PHASE != 1 && e == unres) {
            // a run end that stays open: flag the whole run of listed seeds to its left (phase 1 leaves them open without a walk)
            const uint32_t *ub = u1 + (int64_t)t.kind * words;
            uint32_t *ro0 = run_open + ((int64_t)(t.kind * 2 + t.variant)) * words, *ro1 = c0 != 2 ? run_open + ((int64_t)(t.kind * 2 + 1)) * words : nullptr;
            int64_t w = p >> 5;
            uint32_t m = 0xffffffffu >> (31 - (p & 31));                                       // bits at or below p
            for (;;) {
                const uint32_t inv = ~ub[w] & m;
                if (inv) m &= ~((2u << (31 - __clz(inv))) - 1u);                               // keep the ones above the highest gap
                if (m) { atomicOr(ro0 + w, m); if (ro1) atomicOr(ro1 + w, m); }
                if (inv || w == 0) break;
                w--; m = 0xffffffffu;
            }
        }
        if (e == unres) {
            const unsigned int k = atomicAdd(n_calls + 1, 1u);
            if (k < todo_cap) { SeedTodo o = t; o.pad = 0; todo[k] = o; }
            atomicOr(open_bits + ((int64_t)(t.kind * 2 + t.variant)) * words + (p >> 5), 1u << (p & 31));
            if (c0 != 2) atomicOr(open_bits + ((int64_t)(t.kind * 2 + 1)) * words + (p >> 5), 1u << (p & 31));
            continue;
        }
        head_publish(C, t.kind, t.rank, t.variant, p, e, seeds, words, wp, land, cap, calls, n_del, n_dup, jump0);
    }
}

// Open seeds (ran past the bound: inside genuine events, or the uncovered stretch before the first applied read) come in runs, and the
// path enters a run at its head -- the call made there jumps over the rest.  A head is an open seed with no open seed of its
// (kind, class) in the HEAD_GAP positions before it; for every head the extent of its run (last open bit before a gap of HEAD_GAP) is
// measured by the warp.  Heads are evaluated exactly (unbounded) by the host, all of them in parallel, over windows of records
// fetched in one go; open seeds that turn out to lie on the path all the same are evaluated one by one like before (k_hop_advance).
constexpr int HEAD_GAP = 64;
struct SeedHead { uint32_t rank; uint8_t kind, variant; uint16_t pad; int32_t pos, run_end; };
__global__ void __launch_bounds__(256) k_open_heads(const SeedTodo *__restrict__ todo, const unsigned int *__restrict__ n_todo_p, uint32_t todo_cap, const uint32_t *__restrict__ open_bits,
                                                    int64_t words, SeedHead *__restrict__ heads, uint32_t head_cap, unsigned int *__restrict__ n_heads)
{
    const uint32_t n_todo = min(*n_todo_p, todo_cap);
    const int lane = threadIdx.x & 31;
    for (uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; i0 < n_todo; i0 += gridDim.x * blockDim.x) {
        const uint32_t i = i0 + lane;
        bool head = false;
        SeedTodo t; t.rank = 0; t.kind = 0; t.variant = 0; t.pad = 0; t.pos = 0;
        if (i < n_todo) {
            t = todo[i];
            const uint32_t *ob = open_bits + (int64_t)(t.kind * 2 + t.variant) * words;
            const int64_t p = t.pos, a = max((int64_t)0, p - HEAD_GAP);
            head = true;
            for (int64_t w = a >> 5; w <= (p >> 5) && head; w++) {
                uint32_t m = ob[w];
                if (w == (a >> 5)) m &= 0xffffffffu << (a & 31);
                if (w == (p >> 5)) m &= (p & 31) ? (0xffffffffu >> (32 - (p & 31))) : 0u;
                if (m) head = false;
            }
        }
        unsigned hb = __ballot_sync(0xffffffffu, head);
        while (hb) {
            const int src = __ffs(hb) - 1; hb &= hb - 1;
            const int kind = __shfl_sync(0xffffffffu, (int)t.kind, src), variant = __shfl_sync(0xffffffffu, (int)t.variant, src);
            const int pos = __shfl_sync(0xffffffffu, t.pos, src);
            const uint32_t *ob = open_bits + (int64_t)(kind * 2 + variant) * words;
            // last open position of the run: scan 32 words per step until two consecutive empty words follow the last set bit
            int64_t last = pos;
            for (int64_t wbase = pos >> 5;; wbase += 32) {
                const int64_t w = wbase + lane;
                uint32_t m = w < words ? ob[w] : 0u;
                if (w == (pos >> 5)) m &= 0xffffffffu << (pos & 31);
                const unsigned nz = __ballot_sync(0xffffffffu, m != 0);
                // first pair of consecutive empty words at or after the word holding `last`
                bool stop = false;
                for (int k = 0; k < 32; k++) {
                    const uint32_t mk = __shfl_sync(0xffffffffu, m, k);
                    const int64_t wk = wbase + k;
                    if (mk) last = (wk << 5) + (31 - __clz(mk));
                    else if ((wk << 5) - last > HEAD_GAP) { stop = true; break; }
                }
                (void)nz;
                if (stop || wbase + 32 >= words) break;
            }
            if (lane == src) {
                const unsigned int k = atomicAdd(n_heads, 1u);
                if (k < head_cap) { SeedHead h; h.rank = t.rank; h.kind = t.kind; h.variant = t.variant; h.pad = 0; h.pos = t.pos; h.run_end = (int32_t)last; heads[k] = h; }
            }
        }
    }
}
// outcomes of the host-evaluated heads -> land, speculative call list, level 0 of the jump table, cover bitmap (one warp per head)
struct HeadOutcome { uint32_t rank; uint8_t kind, variant; uint16_t seg; int32_t pos, rel_next; int64_t c_end; double c_z; };
__global__ void __launch_bounds__(128) k_apply_heads(SegCtx Cdel, SegCtx Cdup, const HeadOutcome *__restrict__ ho, uint32_t n, const uint32_t *__restrict__ seeds, int64_t words,
                                                     const uint32_t *__restrict__ wp, uint32_t *__restrict__ land, uint32_t cap, SeedCall *__restrict__ calls, uint32_t call_cap,
                                                     unsigned int *__restrict__ n_calls, uint32_t n_del, uint32_t n_dup, uint32_t *__restrict__ jump0, uint32_t *__restrict__ cover)
{
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const HeadOutcome h = ho[i];
    const SegCtx &C = h.kind ? Cdup : Cdel;
    uint32_t e = 0;
    int ok = 1;
    if (lane == 0) {
        if (h.seg == SEG_RESUME) e = (uint32_t)h.rel_next;
        else {
            const unsigned int k = atomicAdd(n_calls, 1u);
            if (k >= call_cap) ok = 0;                              // stays open: the path evaluates it if it gets there
            else { calls[k].c_end = h.c_end; calls[k].c_z = h.c_z; e = ((uint32_t)SEG_CALL << LAND_SHIFT) | k; }
        }
        if (ok) head_publish(C, h.kind, h.rank, h.variant, h.pos, e, seeds, words, wp, land, cap, calls, n_del, n_dup, jump0);
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    if (ok && h.seg == SEG_CALL) cover_mark(cover + (int64_t)h.kind * words, h.pos, h.c_end, lane, 32);
}

// Second round: the open seeds that are neither run heads nor under a call made at a head (the path jumps over those) get the full
// growth phase, one thread each.  What enters the sliding phase stays open (a call longer than the largest window) and is evaluated by
// the host if the path ever reaches it.
// pass one: the open seeds that are still unresolved (not a head the host closed) and not under a call made at a head, compacted in
// list order (neighbouring seeds stay in neighbouring lanes); pass two: one thread per compacted seed
__global__ void __launch_bounds__(256) k_seed_filter2(const uint32_t *__restrict__ land, uint32_t cap, unsigned int *__restrict__ n_calls, const SeedTodo *__restrict__ todo,
                                                      uint32_t todo_cap, const uint32_t *__restrict__ cover, int64_t words, SeedTodo *__restrict__ out)
{
    const uint32_t n_todo = min(n_calls[1], todo_cap);
    const uint32_t unres = (uint32_t)SEG_UNRESOLVED << LAND_SHIFT;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool keep = false, covered = false;
    SeedTodo t; t.rank = 0; t.kind = 0; t.variant = 0; t.pad = 0; t.pos = 0;
    if (i < n_todo) {
        t = todo[i];
        if (land[((int64_t)t.kind * cap + t.rank) * 2 + t.variant] == unres) {
            covered = (cover[(int64_t)t.kind * words + (t.pos >> 5)] >> (t.pos & 31)) & 1u;
            keep = !covered;
        }
    }
    const unsigned mc = __ballot_sync(0xffffffffu, covered), m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0 && mc) atomicAdd(n_calls + 5, (unsigned int)__popc(mc));                       // under a call made at a head
    if (!m) return;
    unsigned int k0 = 0;
    if (lane == __ffs(m) - 1) k0 = atomicAdd(n_calls + 9, (unsigned int)__popc(m));
    k0 = __shfl_sync(0xffffffffu, k0, __ffs(m) - 1);
    if (keep) out[k0 + __popc(m & ((1u << lane) - 1u))] = t;
}
__global__ void __launch_bounds__(64, 16) k_seed_eval2(SegCtx Cdel, SegCtx Cdup, const uint32_t *__restrict__ seeds, int64_t words, const uint32_t *__restrict__ wp,
                                                   uint32_t *__restrict__ land, uint32_t cap, SeedCall *__restrict__ calls, uint32_t call_cap,
                                                   unsigned int *__restrict__ n_calls, const SeedTodo *__restrict__ todo, uint32_t n_del, uint32_t n_dup,
                                                   uint32_t *__restrict__ jump0)
{
    const uint32_t n_todo = n_calls[9];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_todo) return;
    const SeedTodo t = todo[i];
    const SegCtx &C = t.kind ? Cdup : Cdel;
    const uint32_t unres = (uint32_t)SEG_UNRESOLVED << LAND_SHIFT;
    const int64_t p = t.pos;
    const int c0 = C.cls(p);
    const uint32_t e = seed_outcome(C, p, c0 == 2 ? t.variant : c0, calls, call_cap, n_calls);
    if (e == unres) return;
    atomicAdd(n_calls + 4, 1u);                                                                  // seeds closed by the second round
    head_publish(C, t.kind, t.rank, t.variant, p, e, seeds, words, wp, land, cap, calls, n_del, n_dup, jump0);
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
    bool ensure(size_t n)
    {
        if (n <= cap && p) return true;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = n + n / 4 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) {              // with less headroom before giving up; a failed buffer never looks allocated
            cudaGetLastError();
            if (cudaMalloc(&p, n + 256) != cudaSuccess) { cudaGetLastError(); p = nullptr; return false; }
            cap = n + 256;
            return true;
        }
        cap = want;
        return true;
    }
    template <class T> T *as() { return (T *)p; }
};

// the reference's qsort on the copy-number ratios: glibc merge sort with the comparator `*(int*)a - *(int*)b`, i.e. ordered by the
// low word of each double with wrapping subtraction (src/GROM.c:1105, 20113)
inline void lowword_merge(double *b, size_t n1, size_t n, double *tmp)
{
    size_t i = 0, j = n1, k = 0;
    auto key = [](const double &d) { uint64_t u; memcpy(&u, &d, 8); return (uint32_t)u; };
    while (i < n1 && j < n) {
        if ((int32_t)(key(b[i]) - key(b[j])) <= 0) tmp[k++] = b[i++]; else tmp[k++] = b[j++];
    }
    while (i < n1) tmp[k++] = b[i++];
    memcpy(b, tmp, k * sizeof(double));
}
inline void lowword_msort(double *b, size_t n, double *tmp)
{
    if (n <= 1) return;
    const size_t n1 = n / 2, n2 = n - n1;
    lowword_msort(b, n1, tmp); lowword_msort(b + n1, n2, tmp);
    lowword_merge(b, n1, n, tmp);
}
// the same merge tree with the two halves of the top `fork` levels sorted by different threads: every merge sees the same two
// inputs as in the sequential order of evaluation, so the result is identical (the comparator is not transitive, the tree is what counts)
inline void lowword_msort_par(double *b, size_t n, double *tmp, int fork)
{
    if (fork <= 0 || n < 8192) { lowword_msort(b, n, tmp); return; }
    const size_t n1 = n / 2, n2 = n - n1;
    std::thread left([=]() { lowword_msort_par(b, n1, tmp, fork - 1); });
    lowword_msort_par(b + n1, n2, tmp + n1, fork - 1);
    left.join();
    lowword_merge(b, n1, n, tmp);
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
