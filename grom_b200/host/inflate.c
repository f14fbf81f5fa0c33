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

int grom_inflate_raw(struct grom_inflate_ctx *c, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len)
{
    const uint8_t *ip = in, *const in_end = in + in_len;
    uint8_t *op = out, *const out_end = out + out_len;
    uint64_t bb = 0; int bn = 0;
    int last;
    do {
        REFILL_SAFE();
        if (bn < 3) return -1;
        last = (int)(bb & 1); const int type = (int)((bb >> 1) & 3); DROP(3);
        const ent_t *ll, *ds;
        if (type == 0) {                                   /* stored: skip to the byte boundary, LEN, ~LEN, bytes */
            DROP(bn & 7);
            REFILL_SAFE();
            if (bn < 32) return -1;
            const uint32_t len = (uint32_t)(bb & 0xffff), nlen = (uint32_t)((bb >> 16) & 0xffff); DROP(32);
            if ((len ^ 0xffff) != nlen) return -1;
            /* bytes still in the reservoir belong to the stored data */
            uint32_t left = len;
            while (left && bn >= 8) { if (op >= out_end) return -1; *op++ = (uint8_t)bb; DROP(8); left--; }
            if (left) {
                if (bn != 0) return -1;
                if ((size_t)(in_end - ip) < left || (size_t)(out_end - op) < left) return -1;
                memcpy(op, ip, left); ip += left; op += left;
                bb = 0; bn = 0;
            }
            continue;
        } else if (type == 1) {
            if (!c->fixed_ready) {
                uint8_t l[LL_SYMS];
                for (int i = 0; i < 144; i++) l[i] = 8;
                for (int i = 144; i < 256; i++) l[i] = 9;
                for (int i = 256; i < 280; i++) l[i] = 7;
                for (int i = 280; i < 288; i++) l[i] = 8;
                if (build_table(0, l, 288, LL_BITS, c->fixed_ll, LL_TABLE) < 0) return -1;
                for (int i = 0; i < 32; i++) l[i] = 5;
                if (build_table(1, l, 32, D_BITS, c->fixed_ds, D_TABLE) < 0) return -1;
                c->fixed_ready = 1;
            }
            ll = c->fixed_ll; ds = c->fixed_ds;
        } else if (type == 2) {
            REFILL_SAFE();
            if (bn < 14) return -1;
            const int hlit = (int)(bb & 31) + 257, hdist = (int)((bb >> 5) & 31) + 1, hclen = (int)((bb >> 10) & 15) + 4; DROP(14);
            if (hlit > 286 || hdist > 30) return -1;
            static const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
            uint8_t pl[19]; memset(pl, 0, sizeof(pl));
            for (int i = 0; i < hclen; i++) {
                REFILL_SAFE();
                if (bn < 3) return -1;
                pl[order[i]] = (uint8_t)(bb & 7); DROP(3);
            }
            if (build_table(2, pl, 19, PRE_BITS, c->pre, 1 << PRE_BITS) < 0) return -1;
            uint8_t lens[LL_SYMS + D_SYMS + 140];
            int n = 0; const int total = hlit + hdist;
            while (n < total) {
                REFILL_SAFE();
                const ent_t e = c->pre[bb & ((1 << PRE_BITS) - 1)];
                if (!(e & E_LIT) || (int)E_TOTAL(e) > bn) return -1;
                DROP_E(e);
                const int sym = (int)E_VAL(e);
                if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
                int rep, val = 0;
                if (sym == 16) { if (n == 0 || bn < 2) return -1; val = lens[n - 1]; rep = 3 + (int)(bb & 3); DROP(2); }
                else if (sym == 17) { if (bn < 3) return -1; rep = 3 + (int)(bb & 7); DROP(3); }
                else { if (bn < 7) return -1; rep = 11 + (int)(bb & 127); DROP(7); }
                if (n + rep > total) return -1;
                memset(lens + n, val, (size_t)rep); n += rep;
            }
            if (lens[256] == 0) return -1;                  /* no end-of-block code */
            uint8_t l2[LL_SYMS]; memset(l2, 0, sizeof(l2)); memcpy(l2, lens, (size_t)hlit);
            if (build_table(0, l2, 288, LL_BITS, c->ll, LL_TABLE) < 0) return -1;
            uint8_t d2[D_SYMS]; memset(d2, 0, sizeof(d2)); memcpy(d2, lens + hlit, (size_t)hdist);
            if (build_table(1, d2, 32, D_BITS, c->ds, D_TABLE) < 0) return -1;
            ll = c->ll; ds = c->ds;
        } else return -1;

        /* ---- fast loop: room for a whole iteration on both sides, no bounds tests inside.  `e` is looked up ahead of its use (before the
         * match copy of the previous symbol) so that the table latency overlaps the copy. */
        if ((size_t)(in_end - ip) >= FAST_IN && (size_t)(out_end - op) >= FAST_OUT) {
            const uint8_t *const in_fast = in_end - FAST_IN; uint8_t *const out_fast = out_end - FAST_OUT;
            REFILL_FAST();
            ent_t e = ll[bb & ((1 << LL_BITS) - 1)];
            for (;;) {
                /* here: bn >= 56 counted bits, e = entry of the next symbol */
                if (e & E_LIT) {
                    DROP_E(e); *op++ = (uint8_t)(e >> 16);
                    e = ll[bb & ((1 << LL_BITS) - 1)];
                    if (e & E_LIT) {
                        DROP_E(e); *op++ = (uint8_t)(e >> 16);
                        e = ll[bb & ((1 << LL_BITS) - 1)];
                        if (e & E_LIT) {                                       /* at most 33 bits so far */
                            DROP_E(e); *op++ = (uint8_t)(e >> 16);
                            if (ip > in_fast || op > out_fast) break;
                            e = ll[bb & ((1 << LL_BITS) - 1)]; REFILL_FAST();          /* at least 19 stream bits are still in the reservoir */
                            continue;
                        }
                    }
                    REFILL_FAST();                                              /* up to 22 bits used; a match may need 48 */
                }
                if (e & E_EXC) {
                    if (!(e & E_SUB)) {
                        if (e & E_EOB) { DROP_E(e); goto block_done; }
                        return -1;
                    }
                    DROP_E(e);
                    e = ll[E_VAL(e) + (bb & ((1u << E_CODE(e)) - 1))];
                    if (e & E_LIT) {
                        DROP_E(e); *op++ = (uint8_t)(e >> 16);
                        if (ip > in_fast || op > out_fast) break;
                        REFILL_FAST(); e = ll[bb & ((1 << LL_BITS) - 1)];
                        continue;
                    }
                    if (e & E_EXC) {
                        if (e & E_EOB) { DROP_E(e); goto block_done; }
                        return -1;
                    }
                }
                uint64_t sv = bb;
                DROP_E(e);
                const uint32_t len = BASE_PLUS_EXTRA(e, sv);
                ent_t d = ds[bb & ((1 << D_BITS) - 1)];
                if (d & E_EXC) {
                    if (!(d & E_SUB)) return -1;
                    DROP_E(d);
                    d = ds[E_VAL(d) + (bb & ((1u << E_CODE(d)) - 1))];
                    if (d & E_EXC) return -1;
                }
                sv = bb;
                DROP_E(d);
                const uint32_t dist = BASE_PLUS_EXTRA(d, sv);
                if (dist > (size_t)(op - out)) return -1;
                const uint8_t *src = op - dist; uint8_t *dst = op; op += len;
                const int more = !(ip > in_fast || op > out_fast);
                /* a refill leaves all 64 bits of the reservoir valid (56+ of them counted), a match drops at most 48: the next entry can be
                 * looked up before the refill, which takes the refill (it waits for the bit count) off the path from entry to entry */
                if (more) { e = ll[bb & ((1 << LL_BITS) - 1)]; REFILL_FAST(); }
                if (dist >= 8) {
                    store64(dst, load64(src)); store64(dst + 8, load64(src + 8));            /* most matches are short */
                    if (len > 16) { uint8_t *const e2 = op; dst += 16; src += 16; do { store64(dst, load64(src)); dst += 8; src += 8; } while (dst < e2); }
                } else if (dist == 1) {
                    const uint64_t v = 0x0101010101010101ULL * src[0];
                    uint8_t *const e2 = op; do { store64(dst, v); dst += 8; } while (dst < e2);
                } else {
                    uint8_t *const e2 = op; do { *dst++ = *src++; } while (dst < e2);
                }
                if (!more) break;
            }
        }
        /* ---- careful loop: the tail of the block / of the output */
        for (;;) {
            REFILL_SAFE();
            ent_t e = ll[bb & ((1 << LL_BITS) - 1)];
            if ((e & (E_EXC | E_SUB)) == (E_EXC | E_SUB)) {
                if ((int)E_TOTAL(e) > bn) return -1;
                DROP_E(e);
                e = ll[E_VAL(e) + (bb & ((1u << E_CODE(e)) - 1))];
            }
            if ((int)E_TOTAL(e) > bn) return -1;
            if (e & E_LIT) { if (op >= out_end) return -1; DROP_E(e); *op++ = (uint8_t)(e >> 16); continue; }
            if (e & E_EXC) {
                if (e & E_EOB) { DROP_E(e); goto block_done; }
                return -1;
            }
            uint64_t sv = bb;
            DROP_E(e);
            const uint32_t len = BASE_PLUS_EXTRA(e, sv);
            REFILL_SAFE();
            ent_t d = ds[bb & ((1 << D_BITS) - 1)];
            if ((d & (E_EXC | E_SUB)) == (E_EXC | E_SUB)) {
                if ((int)E_TOTAL(d) > bn) return -1;
                DROP_E(d);
                d = ds[E_VAL(d) + (bb & ((1u << E_CODE(d)) - 1))];
            }
            if ((d & (E_EXC | E_LIT)) || (int)E_TOTAL(d) > bn) return -1;
            sv = bb;
            DROP_E(d);
            const uint32_t dist = BASE_PLUS_EXTRA(d, sv);
            if (dist > (size_t)(op - out) || len > (size_t)(out_end - op)) return -1;
            const uint8_t *src = op - dist;
            for (uint32_t k = 0; k < len; k++) op[k] = src[k];
            op += len;
        }
block_done: ;
    } while (!last);
    return (op == out_end) ? 0 : -1;
}
