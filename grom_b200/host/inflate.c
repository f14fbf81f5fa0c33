/* inflate.c -- raw DEFLATE (RFC 1951) decoder for BGZF blocks.
 *
 * The host batcher (bamio.c) spends most of its time inflating 64 KB BGZF blocks (the reference reads through samtools'
 * bgzf layer over zlib, reference src/GROM.c:82-324 and 981-992: one block at a time, one thread).  A BGZF block is a complete
 * deflate stream with a known output size, which allows a decoder that zlib's streaming interface cannot be: whole-buffer
 * input and output, a 64-bit bit reservoir refilled with one unaligned load, multi-bit table lookups with the extra bits
 * folded into the entry, up to three literals per refill, and word-wide match copies while at least FAST_OUT bytes of
 * output room remain.  The last bytes of every block and anything irregular go through a careful byte-wise loop.
 *
 * Contract: grom_inflate_raw() returns 0 only if the stream is well formed, ends with a final block and produces exactly
 * out_len bytes; it never reads outside [in, in+in_len) or writes outside [out, out+out_len).  Any other outcome is -1,
 * and the caller (bamio.c) then gives the block to zlib, so a stream this decoder refuses is still judged by zlib.
 */
#include <stdint.h>
#include <string.h>
#include "inflate.h"

#define LL_BITS 11                 /* main table of the literal/length code */
#define D_BITS 8                   /* main table of the distance code */
#define PRE_BITS 7
#define LL_SYMS 288
#define D_SYMS 32
#define LL_TABLE (2048 + 288 * 16)
#define D_TABLE (256 + 32 * 128)
#define FAST_OUT 330               /* a fast-loop iteration writes at most 3 literals + a 258-byte match rounded up to whole 8/16-byte words */
#define FAST_IN 32                 /* ... and consumes at most 8 bytes per refill, three refills */

/* Table entry (uint32), one load per symbol:
 *   bits 0-5   bits to drop from the reservoir: the code word and, for length / distance bases, its extra bits as well (bits 6-7 stay 0,
 *              so the low byte is a shift count as it stands); for a subtable pointer the main-table bits
 *   bits 8-11  length of the code word alone (what to shift out to reach the extra bits); for a subtable pointer the subtable's index bits
 *   bit  12    end of block     bit 13  subtable pointer     bit 14  exceptional (end of block, pointer, or -- alone -- invalid)
 *   bits 16-30 literal byte | length or distance base | subtable start
 *   bit  31    literal
 * An all-zero entry never occurs in a built table: unused patterns carry E_EXC alone. */
typedef uint32_t ent_t;
#define E_LIT 0x80000000u
#define E_EOB 0x1000u
#define E_SUB 0x2000u
#define E_EXC 0x4000u
#define E_TOTAL(e) ((e) & 0x3f)
#define E_CODE(e) (((e) >> 8) & 15)
#define E_VAL(e) (((e) >> 16) & 0x7fff)

struct grom_inflate_ctx {
    ent_t ll[LL_TABLE];
    ent_t ds[D_TABLE];
    ent_t pre[1 << PRE_BITS];
    ent_t fixed_ll[LL_TABLE];
    ent_t fixed_ds[D_TABLE];
    int fixed_ready;
};

static const uint16_t len_base[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
static const uint8_t len_extra[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
static const uint16_t dist_base[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
static const uint8_t dist_extra[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };

/* the low n (1..15) bits of v in reverse order */
static inline uint32_t rev_bits(uint32_t v, int n)
{
    static const uint8_t r4[16] = { 0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15 };
    const uint32_t r16 = ((uint32_t)r4[v & 15] << 12) | ((uint32_t)r4[(v >> 4) & 15] << 8) | ((uint32_t)r4[(v >> 8) & 15] << 4) | r4[(v >> 12) & 15];
    return r16 >> (16 - n);
}

/* what symbol `sym` of the code `which` (0 literal/length, 1 distance, 2 code-length code) decodes to, for a code word of `cl` bits still
 * to be dropped at this table level */
static inline ent_t make_entry(int which, int sym, int cl)
{
    if (which == 0) {
        if (sym < 256) return E_LIT | ((uint32_t)sym << 16) | ((uint32_t)cl << 8) | (uint32_t)cl;
        if (sym == 256) return E_EXC | E_EOB | ((uint32_t)cl << 8) | (uint32_t)cl;
        if (sym <= 285) return ((uint32_t)len_base[sym - 257] << 16) | ((uint32_t)cl << 8) | (uint32_t)(cl + len_extra[sym - 257]);
        return E_EXC;                                      /* 286, 287 take part in the code but must not occur */
    }
    if (which == 1) {
        if (sym < 30) return ((uint32_t)dist_base[sym] << 16) | ((uint32_t)cl << 8) | (uint32_t)(cl + dist_extra[sym]);
        return E_EXC;
    }
    return E_LIT | ((uint32_t)sym << 16) | ((uint32_t)cl << 8) | (uint32_t)cl;
}

/* Canonical Huffman code of `n` symbols with the given lengths -> lookup table indexed by the next `tb` stream bits (LSB first), longer
 * codes through subtables.  Returns 0, or -1 for an over-subscribed code.  Incomplete codes are accepted (their unused patterns stay
 * invalid and fail when met), as zlib accepts them for a single distance code. */
static int build_table(int which, const uint8_t *lens, int n, int tb, ent_t *tab, int tab_cap)
{
    int count[16]; memset(count, 0, sizeof(count));
    for (int i = 0; i < n; i++) count[lens[i]]++;
    count[0] = 0;
    int left = 1;
    for (int l = 1; l <= 15; l++) { left = (left << 1) - count[l]; if (left < 0) return -1; }
    uint32_t next[16]; uint32_t code = 0;
    for (int l = 1; l <= 15; l++) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
    const int main_n = 1 << tb;
    if (left > 0) for (int k = 0; k < main_n; k++) tab[k] = E_EXC;         /* incomplete code: some patterns stay unused */
    /* longest code behind every main-table prefix that needs a subtable */
    uint8_t sub_len[1 << LL_BITS];
    int any_long = 0;
    for (int l = tb + 1; l <= 15; l++) if (count[l]) any_long = 1;
    if (any_long) memset(sub_len, 0, (size_t)main_n);
    uint32_t codes[LL_SYMS];
    for (int i = 0; i < n; i++) {
        const int l = lens[i];
        if (!l) continue;
        const uint32_t r = rev_bits(next[l]++, l);
        codes[i] = r;
        if (l > tb) { const uint32_t p = r & (uint32_t)(main_n - 1); if (sub_len[p] < l) sub_len[p] = (uint8_t)l; }
    }
    int used = main_n;
    if (any_long) {
        for (int p = 0; p < main_n; p++) {
            if (!sub_len[p]) continue;
            const int sb = sub_len[p] - tb, sz = 1 << sb;
            if (used + sz > tab_cap) return -1;
            for (int k = 0; k < sz; k++) tab[used + k] = E_EXC;
            tab[p] = E_EXC | E_SUB | ((uint32_t)used << 16) | ((uint32_t)sb << 8) | (uint32_t)tb;
            used += sz;
        }
    }
    for (int i = 0; i < n; i++) {
        const int l = lens[i];
        if (!l) continue;
        const uint32_t r = codes[i];
        if (l <= tb) {
            const ent_t e = make_entry(which, i, l);
            for (uint32_t k = r; k < (uint32_t)main_n; k += 1u << l) tab[k] = e;
        } else {
            const ent_t m = tab[r & (uint32_t)(main_n - 1)];
            const uint32_t sb = E_CODE(m); ent_t *s = tab + E_VAL(m);
            const ent_t e = make_entry(which, i, l - tb);
            for (uint32_t k = r >> tb; k < (1u << sb); k += 1u << (l - tb)) s[k] = e;
        }
    }
    return 0;
}

size_t grom_inflate_ctx_size(void) { return sizeof(struct grom_inflate_ctx); }
void grom_inflate_ctx_init(struct grom_inflate_ctx *c) { c->fixed_ready = 0; }

static inline uint64_t load64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }   /* little-endian host (x86-64 / aarch64 LE) */
static inline void store64(uint8_t *p, uint64_t v) { memcpy(p, &v, 8); }

/* bit reservoir: `bb` holds `bn` counted bits, LSB = next bit of the stream.  After a fast refill the bits above `bn` are stream bits too
 * (the bytes at ip), so refilling again ORs the same values onto them. */
#define REFILL_FAST() do { bb |= load64(ip) << bn; ip += (63 - bn) >> 3; bn |= 56; } while (0)
#define REFILL_SAFE() do { while (bn <= 56 && ip < in_end) { bb |= (uint64_t)*ip++ << bn; bn += 8; } } while (0)
#define DROP(n) do { bb >>= (n); bn -= (n); } while (0)
#define DROP_E(e) do { bb >>= (uint8_t)(e); bn -= (int)E_TOTAL(e); } while (0)
/* base + extra bits of a length / distance entry; `sv` is the reservoir before the entry's bits were dropped */
#define BASE_PLUS_EXTRA(e, sv) (E_VAL(e) + (uint32_t)(((sv) >> E_CODE(e)) & ((1u << (E_TOTAL(e) - E_CODE(e))) - 1)))

#define INFLATE_FN inflate_generic
#define INFLATE_ATTR
#include "inflate_body.inc"
#undef INFLATE_FN
#undef INFLATE_ATTR

#if defined(__x86_64__) && defined(__GNUC__)
#define INFLATE_FN inflate_bmi2
#define INFLATE_ATTR __attribute__((target("bmi2")))
#include "inflate_body.inc"
#undef INFLATE_FN
#undef INFLATE_ATTR
#define GROM_INFLATE_BMI2 1
#endif

int grom_inflate_raw(struct grom_inflate_ctx *c, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len)
{
#ifdef GROM_INFLATE_BMI2
    static int have = -1;
    if (have < 0) have = __builtin_cpu_supports("bmi2") ? 1 : 0;
    if (have) return inflate_bmi2(c, in, in_len, out, out_len);
#endif
    return inflate_generic(c, in, in_len, out, out_len);
}

/* the plain x86-64 / portable variant whatever the CPU offers (tests) */
int grom_inflate_raw_generic(struct grom_inflate_ctx *c, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len)
{
    return inflate_generic(c, in, in_len, out, out_len);
}
