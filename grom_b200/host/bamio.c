/* bamio.c -- host read supply: BGZF/BAM/BAI reader -> packed SoA batch, and the
 * batch -> BAM+BAI serialiser used by the test/bench tooling.
 *
 * Replaces the reference's read supply (reference src/GROM.c:82-324, 981-992)
 * and its per-record field/aux extraction (src/GROM.c:5743-5824): all records of
 * one contig are inflated block-parallel (OpenMP) and unpacked into the
 * structure-of-arrays batch of include/grom_reads.h.  Independent of
 * samtools/htslib; needs only zlib.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE            /* mremap */
#endif
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <ctype.h>
#include <zlib.h>
#include <time.h>
#include <sys/mman.h>
#include <sys/stat.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "gromhost.h"
#include "inflate.h"

static __thread char g_err[512];
const char *gromhost_last_error(void) { return g_err; }
static int fail(const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return -1;
}
/* the same for the other files of the library (not part of the public header) */
int gromhost_fail(const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return -1;
}

/* ------------------------------------------------------------------ BAM open */

struct grom_bam {
    char *path;
    FILE *f;
    int n_targets;
    char **names;
    int64_t *lens;
    uint64_t first_voff;      /* virtual offset of the first alignment record */
    int has_index;
    uint64_t *tgt_beg;        /* per target: smallest chunk_beg in the .bai, UINT64_MAX if none */
    uint64_t *tgt_end;        /* per target: largest chunk_end */
    int64_t *tgt_mapped, *tgt_unmapped;   /* per target: record counts of the index's metadata pseudo-bin, -1 = the index has none */
    const uint8_t *map;       /* the whole file, mapped read-only (or read into memory where mapping is refused) */
    int64_t map_len;
    int map_is_mmap;
};

static inline uint32_t rd_u32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline int32_t  rd_i32(const uint8_t *p) { return (int32_t)rd_u32(p); }

/* header of the BGZF block at h (`avail` bytes readable): gzip member with the FEXTRA flag whose extra field holds the 'B','C' subfield
 * (block size - 1).  htslib writes that subfield alone (XLEN = 6); other subfields before or after it are stepped over.  Returns 0 and
 * the block size and the offset of the deflate stream, or -1. */
static int bgzf_block_header(const uint8_t *h, int64_t avail, int *bsize, int *coff)
{
    if (avail < 18 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return -1;
    const int xlen = h[10] | (h[11] << 8);
    if (12 + (int64_t)xlen > avail) return -1;
    int bs = -1;
    for (int p = 12; p + 4 <= 12 + xlen; ) {
        const int slen = h[p + 2] | (h[p + 3] << 8);
        if (p + 4 + slen > 12 + xlen) return -1;
        if (h[p] == 'B' && h[p + 1] == 'C' && slen == 2) bs = (h[p + 4] | (h[p + 5] << 8)) + 1;
        p += 4 + slen;
    }
    if (bs < 12 + xlen + 8 + 2) return -1;             /* header + (at least) the empty deflate stream + CRC32 + ISIZE */
    *bsize = bs; *coff = 12 + xlen;
    return 0;
}

/* inflate one BGZF block located at file offset `off`; returns isize or -1; *bsize_out = block size */
static int bgzf_inflate_at(FILE *f, int64_t off, uint8_t *raw, uint8_t *dst, int *bsize_out)
{
    uint8_t hdr[12 + 1024];
    if (fseeko(f, off, SEEK_SET) != 0) return -1;
    if (fread(hdr, 1, 18, f) != 18) return -1;
    const int xlen = hdr[10] | (hdr[11] << 8);
    if (xlen > 1024) return -1;
    if (xlen > 6 && fread(hdr + 18, 1, (size_t)xlen - 6, f) != (size_t)xlen - 6) return -1;
    int bsize, coff;
    if (bgzf_block_header(hdr, 12 + (xlen > 6 ? xlen : 6), &bsize, &coff) < 0) return -1;
    if (fseeko(f, off + coff, SEEK_SET) != 0) return -1;
    int clen = bsize - coff;
    if (clen > 65536 + 64) return -1;
    if ((int)fread(raw, 1, clen, f) != clen) return -1;
    uint32_t isize = rd_u32(raw + clen - 4);
    z_stream s; memset(&s, 0, sizeof(s));
    s.next_in = raw; s.avail_in = clen - 8; s.next_out = dst; s.avail_out = 65536;
    if (inflateInit2(&s, -15) != Z_OK) return -1;
    int rc = inflate(&s, Z_FINISH);
    inflateEnd(&s);
    if (rc != Z_STREAM_END || s.total_out != isize) return -1;
    *bsize_out = bsize;
    return (int)isize;
}

/* CRC-32 of the BGZF trailer (the gzip polynomial).  On x86-64 with carry-less multiply the bulk is folded 64 bytes at a time (the
 * folding constants of the gzip polynomial: x^(512+64), x^512, x^(128+64), x^128, x^64 mod P, P and its Barrett inverse); the last
 * n mod 16 bytes, short buffers and every other machine go through zlib's crc32, which also serves as the check in tests/test_inflate.py. */
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
__attribute__((target("pclmul,sse4.1")))
static uint32_t crc32_clmul(const uint8_t *buf, size_t len, uint32_t crc)   /* len >= 64, multiple of 16; crc = running state (inverted form) */
{
    static const uint64_t __attribute__((aligned(16))) k1k2[] = { 0x0154442bd4ULL, 0x01c6e41596ULL };
    static const uint64_t __attribute__((aligned(16))) k3k4[] = { 0x01751997d0ULL, 0x00ccaa009eULL };
    static const uint64_t __attribute__((aligned(16))) k5k0[] = { 0x0163cd6124ULL, 0x0000000000ULL };
    static const uint64_t __attribute__((aligned(16))) poly[] = { 0x01db710641ULL, 0x01f7011641ULL };
    __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
    x1 = _mm_loadu_si128((const __m128i *)(buf + 0x00));
    x2 = _mm_loadu_si128((const __m128i *)(buf + 0x10));
    x3 = _mm_loadu_si128((const __m128i *)(buf + 0x20));
    x4 = _mm_loadu_si128((const __m128i *)(buf + 0x30));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
    x0 = _mm_load_si128((const __m128i *)k1k2);
    buf += 64; len -= 64;
    while (len >= 64) {
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
        x7 = _mm_clmulepi64_si128(x3, x0, 0x00); x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
        x3 = _mm_clmulepi64_si128(x3, x0, 0x11); x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
        y5 = _mm_loadu_si128((const __m128i *)(buf + 0x00)); y6 = _mm_loadu_si128((const __m128i *)(buf + 0x10));
        y7 = _mm_loadu_si128((const __m128i *)(buf + 0x20)); y8 = _mm_loadu_si128((const __m128i *)(buf + 0x30));
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5); x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
        x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7); x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
        buf += 64; len -= 64;
    }
    x0 = _mm_load_si128((const __m128i *)k3k4);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
    while (len >= 16) {
        x2 = _mm_loadu_si128((const __m128i *)buf);
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
        buf += 16; len -= 16;
    }
    x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
    x3 = _mm_setr_epi32(~0, 0, ~0, 0);
    x1 = _mm_srli_si128(x1, 8); x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_loadl_epi64((const __m128i *)k5k0);
    x2 = _mm_srli_si128(x1, 4); x1 = _mm_and_si128(x1, x3); x1 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_load_si128((const __m128i *)poly);
    x2 = _mm_and_si128(x1, x3); x2 = _mm_clmulepi64_si128(x2, x0, 0x10); x2 = _mm_and_si128(x2, x3); x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}
#define GROM_HAVE_CLMUL 1
#endif

uint32_t gromhost_crc32(const uint8_t *p, int64_t n)
{
    uint32_t c = 0;
#ifdef GROM_HAVE_CLMUL
    static int have = -1;
    if (have < 0) have = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    if (have && n >= 64) { const int64_t m = n & ~(int64_t)15; c = ~crc32_clmul(p, (size_t)m, ~c); p += m; n -= m; }
#endif
    return (uint32_t)crc32(c, p, (uInt)n);
}

/* inflate the deflate stream of one BGZF block (`clen` bytes, header and trailer stripped) into exactly `isize` bytes at dst and check
 * the CRC-32 of the trailer.  The batcher's own decoder (inflate.c) runs first; what it refuses is judged by zlib. */
static int bgzf_inflate_block(struct grom_inflate_ctx *ctx, const uint8_t *src, int clen, uint8_t *dst, int isize, uint32_t crc)
{
    static int generic = -1;                     /* GROMHOST_INFLATE=generic: the variant compiled without BMI2 (measurements, tests) */
    if (generic < 0) { const char *v = getenv("GROMHOST_INFLATE"); generic = (v && !strcmp(v, "generic")) ? 1 : 0; }
    if (!ctx || (generic ? grom_inflate_raw_generic(ctx, src, (size_t)clen, dst, (size_t)isize) : grom_inflate_raw(ctx, src, (size_t)clen, dst, (size_t)isize)) != 0) {
        z_stream s; memset(&s, 0, sizeof(s));
        s.next_in = (Bytef *)src; s.avail_in = (uInt)clen; s.next_out = dst; s.avail_out = (uInt)isize;
        if (inflateInit2(&s, -15) != Z_OK) return -1;
        const int rc = inflate(&s, Z_FINISH);
        inflateEnd(&s);
        if (!(rc == Z_STREAM_END && s.total_out == (uLong)isize)) return -1;
    }
    return (gromhost_crc32(dst, isize) == crc) ? 0 : -1;
}

/* test / tooling entry: one raw deflate stream with a known output size through the batcher's own decoder only (no zlib fallback) */
int gromhost_inflate_raw(const uint8_t *src, int64_t src_len, uint8_t *dst, int64_t dst_len)
{
    struct grom_inflate_ctx *ctx = (struct grom_inflate_ctx *)malloc(grom_inflate_ctx_size());
    if (!ctx) return fail("out of memory");
    grom_inflate_ctx_init(ctx);
    const char *v = getenv("GROMHOST_INFLATE");                     /* "generic": the variant compiled without BMI2 */
    const int rc = (v && !strcmp(v, "generic")) ? grom_inflate_raw_generic(ctx, src, (size_t)src_len, dst, (size_t)dst_len)
                                                : grom_inflate_raw(ctx, src, (size_t)src_len, dst, (size_t)dst_len);
    free(ctx);
    return rc == 0 ? 0 : fail("not a well-formed deflate stream of %lld bytes", (long long)dst_len);
}

/* sequential reader used only for the header */
typedef struct { FILE *f; int64_t addr, next; uint8_t raw[65536 + 64], blk[65536]; int len, off; } seqrd;
static int seq_fill(seqrd *r)
{
    int bs;
    r->addr = r->next;
    int n = bgzf_inflate_at(r->f, r->addr, r->raw, r->blk, &bs);
    if (n < 0) return -1;
    r->next = r->addr + bs; r->len = n; r->off = 0;
    return 0;
}
static int seq_read(seqrd *r, void *dst, int n)
{
    uint8_t *d = (uint8_t *)dst; int got = 0;
    while (got < n) {
        if (r->off >= r->len) { if (seq_fill(r) < 0) return got; if (r->len == 0) continue; }
        int k = r->len - r->off; if (k > n - got) k = n - got;
        memcpy(d + got, r->blk + r->off, k); r->off += k; got += k;
    }
    return got;
}

int gromhost_bam_open(const char *path, grom_bam **out)
{
    FILE *f = fopen(path, "rb");
    if (!f) return fail("Could not open %s", path);
    seqrd *r = (seqrd *)calloc(1, sizeof(seqrd));
    r->f = f;
    char magic[4]; int32_t l_text, n_ref;
    if (seq_read(r, magic, 4) != 4 || memcmp(magic, "BAM\1", 4)) { free(r); fclose(f); return fail("%s: not a BAM file", path); }
    seq_read(r, &l_text, 4);
    char *text = (char *)malloc((size_t)l_text + 1);
    seq_read(r, text, l_text); free(text);
    seq_read(r, &n_ref, 4);
    grom_bam *b = (grom_bam *)calloc(1, sizeof(*b));
    b->path = strdup(path); b->f = f; b->n_targets = n_ref;
    b->names = (char **)calloc(n_ref > 0 ? n_ref : 1, sizeof(char *));
    b->lens = (int64_t *)calloc(n_ref > 0 ? n_ref : 1, sizeof(int64_t));
    for (int i = 0; i < n_ref; i++) {
        int32_t l_name, l_ref;
        seq_read(r, &l_name, 4);
        b->names[i] = (char *)calloc((size_t)l_name + 1, 1);
        seq_read(r, b->names[i], l_name);
        seq_read(r, &l_ref, 4);
        b->lens[i] = l_ref;
    }
    if (r->off >= r->len) { b->first_voff = (uint64_t)r->next << 16; }
    else b->first_voff = ((uint64_t)r->addr << 16) | (uint64_t)r->off;
    free(r);

    /* optional index: keep only [min chunk_beg, max chunk_end] per target */
    /* "<bam>.bai", else "<stem>.bai" for a name that ends in "bam" (the two places samtools' bam_index_load looks, src/GROM.c:220, 22129) */
    char iname[4096]; snprintf(iname, sizeof(iname), "%s.bai", path);
    FILE *fi = fopen(iname, "rb");
    if (!fi) {
        const size_t pl = strlen(path);
        if (pl >= 3 && pl < sizeof(iname) && !strcmp(path + pl - 3, "bam")) { snprintf(iname, sizeof(iname), "%s", path); iname[pl - 1] = 'i'; fi = fopen(iname, "rb"); }
    }
    if (fi) {
        char m[4]; int32_t nr = 0;
        if (fread(m, 1, 4, fi) == 4 && !memcmp(m, "BAI\1", 4) && fread(&nr, 4, 1, fi) == 1 && nr == n_ref) {
            b->tgt_beg = (uint64_t *)malloc(sizeof(uint64_t) * (n_ref > 0 ? n_ref : 1));
            b->tgt_end = (uint64_t *)malloc(sizeof(uint64_t) * (n_ref > 0 ? n_ref : 1));
            b->tgt_mapped = (int64_t *)malloc(sizeof(int64_t) * (n_ref > 0 ? n_ref : 1));
            b->tgt_unmapped = (int64_t *)malloc(sizeof(int64_t) * (n_ref > 0 ? n_ref : 1));
            for (int i = 0; i < n_ref; i++) b->tgt_mapped[i] = b->tgt_unmapped[i] = -1;
            int ok = 1;
            for (int i = 0; i < n_ref && ok; i++) {
                int32_t n_bin, n_intv;
                b->tgt_beg[i] = UINT64_MAX; b->tgt_end[i] = 0;
                if (fread(&n_bin, 4, 1, fi) != 1) { ok = 0; break; }
                for (int j = 0; j < n_bin && ok; j++) {
                    uint32_t bin; int32_t n_chunk;
                    if (fread(&bin, 4, 1, fi) != 1 || fread(&n_chunk, 4, 1, fi) != 1) { ok = 0; break; }
                    for (int k = 0; k < n_chunk; k++) {
                        uint64_t be[2];
                        if (fread(be, 8, 2, fi) != 2) { ok = 0; break; }
                        if (bin == 37450) {              /* metadata pseudo-bin: (first, last virtual offset), (mapped, unmapped records) */
                            if (k == 1) { b->tgt_mapped[i] = (int64_t)be[0]; b->tgt_unmapped[i] = (int64_t)be[1]; }
                            continue;
                        }
                        if (be[0] < b->tgt_beg[i]) b->tgt_beg[i] = be[0];
                        if (be[1] > b->tgt_end[i]) b->tgt_end[i] = be[1];
                    }
                }
                if (!ok || fread(&n_intv, 4, 1, fi) != 1) { ok = 0; break; }
                fseeko(fi, (off_t)n_intv * 8, SEEK_CUR);
            }
            b->has_index = ok;
        }
        fclose(fi);
    }
    /* the alignment blocks are read in place */
    struct stat st;
    if (fstat(fileno(f), &st) != 0 || st.st_size <= 0) { gromhost_bam_close(b); return fail("%s: cannot stat", path); }
    b->map_len = (int64_t)st.st_size;
    void *m = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fileno(f), 0);
    if (m != MAP_FAILED) { b->map = (const uint8_t *)m; b->map_is_mmap = 1; }
    else {
        uint8_t *buf = (uint8_t *)malloc((size_t)st.st_size);
        if (!buf || fseeko(f, 0, SEEK_SET) != 0 || fread(buf, 1, (size_t)st.st_size, f) != (size_t)st.st_size) { free(buf); gromhost_bam_close(b); return fail("%s: cannot map or read", path); }
        b->map = buf;
    }
    *out = b;
    return 0;
}

void gromhost_bam_close(grom_bam *b)
{
    if (!b) return;
    if (b->map) { if (b->map_is_mmap) munmap((void *)b->map, (size_t)b->map_len); else free((void *)b->map); }
    for (int i = 0; i < b->n_targets; i++) free(b->names[i]);
    free(b->names); free(b->lens); free(b->tgt_beg); free(b->tgt_end); free(b->tgt_mapped); free(b->tgt_unmapped); free(b->path);
    if (b->f) fclose(b->f);
    free(b);
}
int gromhost_bam_n_targets(const grom_bam *b) { return b->n_targets; }
const char *gromhost_bam_target_name(const grom_bam *b, int tid) { return (tid >= 0 && tid < b->n_targets) ? b->names[tid] : NULL; }
int64_t gromhost_bam_target_len(const grom_bam *b, int tid) { return (tid >= 0 && tid < b->n_targets) ? b->lens[tid] : -1; }
int gromhost_bam_has_index(const grom_bam *b) { return b->has_index; }
int gromhost_bam_target_reads(const grom_bam *b, int tid, int64_t *mapped, int64_t *unmapped)
{
    if (!b->has_index || tid < 0 || tid >= b->n_targets || !b->tgt_mapped) return -1;
    if (b->tgt_mapped[tid] < 0) {
        if (b->tgt_beg[tid] != UINT64_MAX) return -1;             /* records, but no count of them (an index from before samtools 0.1.8) */
        if (mapped) *mapped = 0;                                   /* no bins at all: a target without records */
        if (unmapped) *unmapped = 0;
        return 0;
    }
    if (mapped) *mapped = b->tgt_mapped[tid];
    if (unmapped) *unmapped = b->tgt_unmapped[tid];
    return 0;
}

/* ------------------------------------------------------------------ batch */

struct grom_batch {
    grom_read_batch v;
    int64_t cap_reads, cap_cigar, cap_slots, cap_names;
    int32_t *pos, *mpos, *tlen, *mtid, *l_qseq, *sa_pos, *sa_start_adj, *sa_end_adj, *sa_end_adj_indel;
    uint16_t *flag, *n_cigar; int16_t *sa_mapq;
    uint8_t *mapq, *qname_len, *sa_strand, *sa_same_chr;
    uint64_t *qname_hash, *cigar_off, *base_off, *qname_off;
    uint32_t *cigar; uint8_t *seq4, *qual; char *qname_pool;
    /* transport-compact forms (grom_reads.h GROM_LAYOUT_*), built once the canonical arrays are filled */
    uint8_t *seq2, *seq_exc_code, *qual2; uint64_t *seq_exc_slot;
    size_t big_seq4, big_qual, big_seq2, big_qual2;      /* lengths of the arrays that are zero-filled anonymous mappings (big_zalloc), 0 = malloc'd */
    uint8_t *qual4; int32_t *sa_index, *sas_pos, *sas_start_adj, *sas_end_adj, *sas_end_adj_indel; int16_t *sas_mapq; uint8_t *sas_strand, *sas_same_chr;
};

/* the large per-base arrays: zero-filled anonymous mappings, huge pages where the kernel grants them (far fewer page faults while the
 * threads fill them); small ones come from calloc.  *len = 0 means "free() it". */
static void *big_zalloc(size_t bytes, size_t *len)
{
    *len = 0;
    if (bytes < ((size_t)8 << 20)) return calloc(bytes, 1);
    const size_t n = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    void *p = mmap(NULL, n, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return calloc(bytes, 1);
#ifdef MADV_HUGEPAGE
    if (!getenv("GROMHOST_NO_HUGEPAGE")) madvise(p, n, MADV_HUGEPAGE);
#endif
    *len = n;
    return p;
}
static void big_free(void *p, size_t len) { if (!p) return; if (len) munmap(p, len); else free(p); }

void gromhost_batch_view(const grom_batch *bt, grom_read_batch *view) { *view = bt->v; }

void gromhost_batch_free(grom_batch *t)
{
    if (!t) return;
    free(t->pos); free(t->mpos); free(t->tlen); free(t->mtid); free(t->l_qseq); free(t->sa_pos);
    free(t->sa_start_adj); free(t->sa_end_adj); free(t->sa_end_adj_indel); free(t->flag); free(t->n_cigar);
    free(t->sa_mapq); free(t->mapq); free(t->qname_len); free(t->sa_strand); free(t->sa_same_chr);
    free(t->qname_hash); free(t->cigar_off); free(t->base_off); free(t->qname_off); free(t->cigar);
    big_free(t->seq4, t->big_seq4); big_free(t->qual, t->big_qual); free(t->qname_pool);
    big_free(t->seq2, t->big_seq2); free(t->seq_exc_code); free(t->seq_exc_slot); big_free(t->qual2, t->big_qual2);
    free(t->qual4); free(t->sa_index); free(t->sas_pos); free(t->sas_start_adj); free(t->sas_end_adj); free(t->sas_end_adj_indel);
    free(t->sas_mapq); free(t->sas_strand); free(t->sas_same_chr); free(t);
}

static void batch_publish(grom_batch *t)
{
    grom_read_batch *v = &t->v;
    v->pos = t->pos; v->mpos = t->mpos; v->tlen = t->tlen; v->mtid = t->mtid; v->l_qseq = t->l_qseq;
    v->flag = t->flag; v->n_cigar = t->n_cigar; v->mapq = t->mapq; v->qname_len = t->qname_len;
    v->qname_hash = t->qname_hash; v->cigar_off = t->cigar_off; v->base_off = t->base_off;
    v->cigar = t->cigar; v->seq4 = t->seq4; v->qual = t->qual; v->sa_pos = t->sa_pos;
    v->sa_start_adj = t->sa_start_adj; v->sa_end_adj = t->sa_end_adj; v->sa_end_adj_indel = t->sa_end_adj_indel;
    v->sa_strand = t->sa_strand; v->sa_mapq = t->sa_mapq; v->sa_same_chr = t->sa_same_chr;
    v->qname_off = t->qname_off; v->qname_pool = t->qname_pool;
}

/* ---- aux: first XP else SA entry (reference src/GROM.c:5763-5824, 6686-6733) ---- */

static const uint8_t *aux_find(const uint8_t *s, const uint8_t *e, char a, char b)
{
    while (s + 3 <= e) {
        int hit = (s[0] == (uint8_t)a && s[1] == (uint8_t)b);
        uint8_t t = s[2];
        const uint8_t *v = s + 2;
        s += 3;
        if (hit) return v;
        switch (t) {
        case 'A': case 'c': case 'C': s += 1; break;
        case 's': case 'S': s += 2; break;
        case 'i': case 'I': case 'f': s += 4; break;
        case 'd': s += 8; break;
        case 'Z': case 'H': while (s < e && *s) s++; s++; break;
        case 'B': { if (s + 5 > e) return NULL; uint8_t st = s[0]; uint32_t n = rd_u32(s + 1);
                    int w = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
                    s += 5 + (size_t)n * w; break; }
        default: return NULL;
        }
    }
    return NULL;
}

/* split on ',' like strtok (skips empty fields); returns number of tokens, up to 5 */
static int split_commas(char *s, char *tok[5])
{
    int n = 0;
    while (n < 5) {
        while (*s == ',') s++;
        if (!*s) break;
        tok[n++] = s;
        while (*s && *s != ',') s++;
        if (*s) *s++ = 0;
    }
    return n;
}

static void parse_sa(const uint8_t *aux, int l_aux, const char *target_name,
                     int32_t *sa_pos, uint8_t *strand, int16_t *mq, uint8_t *same,
                     int32_t *start_adj, int32_t *end_adj, int32_t *end_adj_indel)
{
    *sa_pos = -1; *strand = 0; *mq = -1; *same = 0; *start_adj = *end_adj = *end_adj_indel = 0;
    if (!(l_aux > 0 && l_aux < 100)) return;
    const uint8_t *e = aux + l_aux;
    int is_xp = 1;
    const uint8_t *v = aux_find(aux, e, 'X', 'P');
    if (!v) { is_xp = 0; v = aux_find(aux, e, 'S', 'A'); }
    if (!v) return;
    char buf[128];
    const uint8_t *src = (v[0] == 'Z') ? v + 1 : v;
    int n = 0;
    while (src + n < e && src[n] && n < 127) { buf[n] = (char)src[n]; n++; }
    buf[n] = 0;
    char *tok[5];
    int nt = split_commas(buf, tok);
    const char *cig = NULL;
    if (is_xp) {           /* chr,[+-]pos,cigar,mq */
        if (nt < 4) return;
        *strand = (tok[1][0] == '+') ? 0 : 1;
        *sa_pos = atoi(tok[1] + 1);
        cig = tok[2]; *mq = (int16_t)atoi(tok[3]);
    } else {               /* chr,pos,strand,cigar,mq,... */
        if (nt < 5) return;
        *sa_pos = atoi(tok[1]);
        *strand = (tok[2][0] == '+') ? 0 : 1;
        cig = tok[3]; *mq = (int16_t)atoi(tok[4]);
    }
    *same = (strncmp(target_name, tok[0], strlen(target_name)) == 0) ? 1 : 0;
    /* aux CIGAR: digits accumulate until an alphabetic op letter; non-alnum characters
     * are skipped without resetting the digit buffer (src/GROM.c:6693-6708) */
    char digits[32]; int nd = 0; int first = 1; char last_t = 0; long last_l = 0;
    for (const char *c = cig; *c; c++) {
        if (isdigit((unsigned char)*c)) { if (nd < 31) digits[nd++] = *c; }
        else if (isalpha((unsigned char)*c)) {
            digits[nd] = 0; long len = strtol(digits, NULL, 10); nd = 0;
            if (first) { if (*c == 'S') *start_adj = (int32_t)len; first = 0; }
            if (*c == 'I') *end_adj_indel += (int32_t)len;
            else if (*c == 'D') *end_adj_indel -= (int32_t)len;
            last_t = *c; last_l = len;
        }
    }
    if (last_t == 'S') *end_adj = (int32_t)last_l;
}

/* ------------------------------------------------------------------ read one target */

/* GROMHOST_TRACE=1: phase times of gromhost_bam_read_target on stderr */
static double now_ms(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
#define TRACE_MARK(what) do { if (trace) { double t_ = now_ms(); fprintf(stderr, "[bamio] %-22s %8.2f ms\n", what, t_ - t_last); t_last = t_; } } while (0)

typedef struct { int64_t off; int bsize, coff, isize; int64_t uoff; } blkinfo;      /* coff: where the deflate stream starts inside the block */

/* reads with a first XP / SA entry, collected per thread during the fill pass (a small minority of the reads) */
typedef struct { int32_t idx, pos, start_adj, end_adj, end_adj_indel; int16_t mapq; uint8_t strand, same_chr; } sa_ent;
typedef struct { sa_ent *e; int64_t n, cap; uint8_t present[256]; char pad[64]; } fill_local;

/* the records of target `tid` in the inflated stream, in file order: offsets of the records and the running sums that become
 * cigar_off / base_off / qname_off (the canonical offsets of include/grom_reads.h) */
typedef struct { int64_t n, cap, n_cig, n_slots, n_name; int64_t *recoff; uint64_t *cig_off, *base_off, *name_off; } reclist;

static int reclist_grow(reclist *r, int keep_names)
{
    int64_t cap = r->cap ? r->cap * 2 : 4096;
    int64_t *a = (int64_t *)realloc(r->recoff, sizeof(int64_t) * (size_t)cap); if (!a) return -1; r->recoff = a;
    uint64_t *b = (uint64_t *)realloc(r->cig_off, sizeof(uint64_t) * (size_t)cap); if (!b) return -1; r->cig_off = b;
    uint64_t *c = (uint64_t *)realloc(r->base_off, sizeof(uint64_t) * (size_t)cap); if (!c) return -1; r->base_off = c;
    if (keep_names) { uint64_t *d = (uint64_t *)realloc(r->name_off, sizeof(uint64_t) * (size_t)(cap + 1)); if (!d) return -1; r->name_off = d; }
    r->cap = cap;
    return 0;
}

/* One pass over the record chain from offset p up to (not including) offset `stop`: validates the layout of every record of the target
 * (a truncated or corrupt file must not make the fill pass read past the inflated data) and lists them.  `started` says whether records
 * of the target came before p.  Outcome in *term: 0 = arrived exactly at `stop`; 1 = the chain ended (a record past the target or in the
 * unplaced tail, or a block_size that cannot be one); 5 = the data ended (inside a record, or exactly behind one): the chain may go on
 * in data not yet inflated; 2 = stepped over `stop` (it was no record boundary); 3 = corrupt record at *err_at; 4 = memory.
 * *stop_at = where the walk stopped; *lead = the first record seen was not one of the target. */
typedef struct { int term, lead; int64_t err_at, stop_at; } walk_end;
static void walk_records(const uint8_t *u, int64_t utotal, int64_t p, int64_t stop, int started, int tid, int keep_names, reclist *r, walk_end *w)
{
    int first = 1;
    w->term = 5; w->lead = 0; w->err_at = -1;
    while (p + 36 <= utotal) {
        if (p >= stop) { w->term = (p == stop) ? 0 : 2; w->stop_at = p; return; }
        const int32_t bl = rd_i32(u + p);
        if (bl < 32) { w->term = 1; w->stop_at = p; return; }
        if (p + 4 + bl > utotal) { w->stop_at = p; return; }
        const int32_t rtid = rd_i32(u + p + 4);
        if (rtid == tid) {
            started = 1;
            const uint32_t bmq = rd_u32(u + p + 12), fnc = rd_u32(u + p + 16); const int32_t lq = rd_i32(u + p + 20);
            if (lq < 0 || 32 + (int64_t)(bmq & 0xff) + 4 * (int64_t)(fnc & 0xffff) + ((int64_t)lq + 1) / 2 + (int64_t)lq > (int64_t)bl) { w->term = 3; w->err_at = p; w->stop_at = p; return; }
            if (r->n == r->cap && reclist_grow(r, keep_names) < 0) { w->term = 4; w->stop_at = p; return; }
            r->recoff[r->n] = p; r->cig_off[r->n] = (uint64_t)r->n_cig; r->base_off[r->n] = (uint64_t)r->n_slots;
            if (keep_names) r->name_off[r->n] = (uint64_t)r->n_name;
            r->n++; r->n_cig += fnc & 0xffff; r->n_slots += (lq + GROM_BASE_ALIGN - 1) / GROM_BASE_ALIGN * GROM_BASE_ALIGN; r->n_name += bmq & 0xff;
        } else {
            if (first) w->lead = 1;
            if (started || rtid > tid || rtid < 0) { w->term = 1; w->stop_at = p; return; }      /* coordinate-sorted: past the target (or into the unplaced tail) */
        }
        first = 0;
        p += 4 + bl;
    }
    w->stop_at = p;
    if (p >= stop) w->term = (p == stop) ? 0 : 2;                /* the chain may end exactly where the next range begins */
}

static int corrupt_record(const char *path, const uint8_t *u, int64_t p)
{
    const uint32_t bmq = rd_u32(u + p + 12), fnc = rd_u32(u + p + 16);
    return fail("%s: corrupt BAM record at uncompressed offset %lld (block_size %d cannot hold name %u + %u CIGAR ops + %d bases)",
                path, (long long)p, rd_i32(u + p), bmq & 0xff, fnc & 0xffff, rd_i32(u + p + 20));
}

/* does a well-formed record start at p?  (used to guess entry points into the chain; a guess is only kept when the walk of the range
 * before it arrives exactly there) */
static inline int record_plausible(const uint8_t *u, int64_t utotal, int64_t p, int n_targets)
{
    if (p + 36 > utotal) return 0;
    const int32_t bl = rd_i32(u + p);
    if (bl < 32 || bl > (1 << 24) || p + 4 + bl > utotal) return 0;
    const int32_t rtid = rd_i32(u + p + 4), pos = rd_i32(u + p + 8), lq = rd_i32(u + p + 20), mtid = rd_i32(u + p + 24), mpos = rd_i32(u + p + 28);
    if (rtid < -1 || rtid >= n_targets || mtid < -1 || mtid >= n_targets || pos < -1 || mpos < -1 || lq < 0) return 0;
    const uint32_t bmq = rd_u32(u + p + 12), fnc = rd_u32(u + p + 16);
    const int l_qname = bmq & 0xff;
    if (l_qname < 1 || 32 + (int64_t)l_qname + 4 * (int64_t)(fnc & 0xffff) + ((int64_t)lq + 1) / 2 + (int64_t)lq > (int64_t)bl) return 0;
    return u[p + 36 + l_qname - 1] == 0;
}

#define WALK_SYNC_CHAIN 8          /* consecutive plausible records that make an entry-point guess */
#define WALK_PAR_MIN (8 << 20)     /* inflated bytes below which the chain is walked by one thread */

/* The record chain of a target, listed by all threads: BAM records carry no synchronisation marks, so every thread but the first guesses
 * an entry point (the first offset of its share at which WALK_SYNC_CHAIN plausible records follow one another) and walks from there; a
 * guess counts only if the walk of the share before it arrives exactly at it, which makes it a boundary of the true chain by induction
 * from the known first record.  Any miss falls back to the one-thread walk, so the result never depends on the guesses.
 * Returns 0, -1 (corrupt record, message set) or -2 (memory); the lists of all shares are concatenated into *out. */
static int walk_records_parallel(const uint8_t *u, int64_t utotal, int64_t p0, int started0, int tid, int n_targets, int keep_names, int n_threads,
                                 const char *path, reclist *out, walk_end *end)
{
    memset(out, 0, sizeof(*out));
    end->term = 5; end->lead = 0; end->err_at = -1; end->stop_at = p0;
    int T = n_threads;
    int64_t par_min = WALK_PAR_MIN;
    { const char *e = getenv("GROMHOST_WALK_PAR_MIN"); if (e && *e) par_min = atoll(e); }      /* tests: force / forbid the all-thread walk */
    if (utotal - p0 < par_min || T < 2) T = 1;
    if (T > 64) T = 64;
    reclist rl[64]; walk_end we[64]; int64_t start[65];
    memset(rl, 0, sizeof(reclist) * (size_t)T);
    int rc = 0, n_used = 0;
    if (T > 1) {
        start[0] = p0; start[T] = INT64_MAX;
        #pragma omp parallel for schedule(static, 1) num_threads(T)
        for (int k = 1; k < T; k++) {
            const int64_t g = p0 + (utotal - p0) / T * k, lim = p0 + (utotal - p0) / T * (k + 1);
            int64_t found = -1;
            for (int64_t q = g; q < lim && found < 0; q++) {
                if (!record_plausible(u, utotal, q, n_targets)) continue;
                int64_t c = q; int ok = 1;
                for (int j = 0; j < WALK_SYNC_CHAIN && ok; j++) {
                    c += 4 + (int64_t)rd_i32(u + c);
                    if (c + 36 > utotal) break;                   /* the data end inside the chain: nothing left to contradict the guess */
                    ok = record_plausible(u, utotal, c, n_targets);
                }
                if (ok) found = q;
            }
            start[k] = found;
        }
        for (int k = T - 1; k >= 1; k--) if (start[k] < 0) start[k] = start[k + 1];      /* a share without an entry point joins the one before it */
        int grow_bad = 0;
        #pragma omp parallel for schedule(static, 1) num_threads(T)
        for (int k = 0; k < T; k++) {
            if (reclist_grow(&rl[k], keep_names) < 0) { we[k].term = 4; grow_bad = 1; continue; }
            if (k > 0 && start[k] == start[k + 1]) { we[k].term = 0; we[k].lead = 0; we[k].err_at = -1; continue; }    /* empty share */
            walk_records(u, utotal, start[k], start[k + 1], k == 0 ? started0 : 0, tid, keep_names, &rl[k], &we[k]);
        }
        /* which shares are on the true chain: the first is; a share is reached when the one before arrived exactly at its start */
        int started = started0, fallback = grow_bad;
        for (int k = 0; k < T && !fallback; k++) {
            const int empty = k > 0 && start[k] == start[k + 1];
            if (started && we[k].lead && !empty) { end->term = 1; end->stop_at = start[k]; break; }          /* a foreign record after the target began ends the chain */
            if (we[k].term == 2 || we[k].term == 4) { fallback = 1; break; }
            n_used = k + 1;
            if (rl[k].n) started = 1;
            if (we[k].term == 3) { rc = corrupt_record(path, u, we[k].err_at); break; }
            if (!empty) { end->term = we[k].term; end->stop_at = we[k].stop_at; }
            if (we[k].term == 1 || we[k].term == 5) break;
            /* a share that began after records of the target (started) but met a foreign record first was handled above; one that began
             * before the target and skipped foreign records is what the one-thread walk does as well */
        }
        if (getenv("GROMHOST_TRACE")) fprintf(stderr, "[bamio] record chain: %d shares, %d on the chain, fallback=%d\n", T, n_used, fallback);
        if (fallback) { for (int k = 0; k < T; k++) { free(rl[k].recoff); free(rl[k].cig_off); free(rl[k].base_off); free(rl[k].name_off); } memset(rl, 0, sizeof(reclist) * (size_t)T); T = 1; n_used = 0; rc = 0; }
    }
    if (T == 1) {
        if (reclist_grow(&rl[0], keep_names) < 0) { free(rl[0].recoff); free(rl[0].cig_off); free(rl[0].base_off); free(rl[0].name_off); return -2; }
        walk_records(u, utotal, p0, INT64_MAX, started0, tid, keep_names, &rl[0], &we[0]);
        if (we[0].term == 3) rc = corrupt_record(path, u, we[0].err_at);
        else if (we[0].term == 4) rc = -2;
        if (rc == 0) { *out = rl[0]; *end = we[0]; return 0; }
        free(rl[0].recoff); free(rl[0].cig_off); free(rl[0].base_off); free(rl[0].name_off);
        return rc;
    }
    if (rc == 0) {
        int64_t base_n[65], base_c[65], base_s[65], base_m[65];
        base_n[0] = base_c[0] = base_s[0] = base_m[0] = 0;
        for (int k = 0; k < n_used; k++) { base_n[k + 1] = base_n[k] + rl[k].n; base_c[k + 1] = base_c[k] + rl[k].n_cig; base_s[k + 1] = base_s[k] + rl[k].n_slots; base_m[k + 1] = base_m[k] + rl[k].n_name; }
        const int64_t n = base_n[n_used];
        out->n = n; out->cap = n > 0 ? n : 1; out->n_cig = base_c[n_used]; out->n_slots = base_s[n_used]; out->n_name = base_m[n_used];
        out->recoff = (int64_t *)malloc(sizeof(int64_t) * (size_t)out->cap); out->cig_off = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)out->cap);
        out->base_off = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)out->cap);
        out->name_off = keep_names ? (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(out->cap + 1)) : NULL;
        if (!out->recoff || !out->cig_off || !out->base_off || (keep_names && !out->name_off)) {
            free(out->recoff); free(out->cig_off); free(out->base_off); free(out->name_off); memset(out, 0, sizeof(*out)); rc = -2;
        } else {
            #pragma omp parallel for schedule(static, 1) num_threads(T)
            for (int k = 0; k < n_used; k++) {
                const int64_t o = base_n[k]; const uint64_t c = (uint64_t)base_c[k], sl = (uint64_t)base_s[k], m = (uint64_t)base_m[k];
                memcpy(out->recoff + o, rl[k].recoff, sizeof(int64_t) * (size_t)rl[k].n);
                for (int64_t i = 0; i < rl[k].n; i++) { out->cig_off[o + i] = rl[k].cig_off[i] + c; out->base_off[o + i] = rl[k].base_off[i] + sl; }
                if (keep_names) for (int64_t i = 0; i < rl[k].n; i++) out->name_off[o + i] = rl[k].name_off[i] + m;
            }
        }
    }
    for (int k = 0; k < T; k++) { free(rl[k].recoff); free(rl[k].cig_off); free(rl[k].base_off); free(rl[k].name_off); }
    return rc;
}

/* ---- the batch under construction: every array grows with the windows of the target that have been decoded so far */
typedef struct { int64_t reads, cig, slots, names; } batch_caps;

static void *grow_zeroed(void *p, size_t old_bytes, size_t new_bytes)
{
    void *q = realloc(p, new_bytes);
    if (q && new_bytes > old_bytes) memset((char *)q + old_bytes, 0, new_bytes - old_bytes);
    return q;
}
/* a large per-base array: stays a zero-filled anonymous mapping (moved by the kernel when it has to grow: no copy) */
static int big_grow(uint8_t **p, size_t *maplen, size_t old_bytes, size_t new_bytes)
{
    if (*p == NULL) { *p = (uint8_t *)big_zalloc(new_bytes, maplen); return *p ? 0 : -1; }
    if (*maplen) {
        if (new_bytes <= *maplen) return 0;
        const size_t n = (new_bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
        void *q = mremap(*p, *maplen, n, MREMAP_MAYMOVE);
        if (q == MAP_FAILED) return -1;
#ifdef MADV_HUGEPAGE
        if (!getenv("GROMHOST_NO_HUGEPAGE")) madvise(q, n, MADV_HUGEPAGE);
#endif
        *p = (uint8_t *)q; *maplen = n;
        return 0;
    }
    if (new_bytes >= ((size_t)8 << 20)) {                          /* outgrew calloc: move into a mapping */
        size_t ml; uint8_t *q = (uint8_t *)big_zalloc(new_bytes, &ml);
        if (!q) return -1;
        memcpy(q, *p, old_bytes); free(*p); *p = q; *maplen = ml;
        return 0;
    }
    uint8_t *q = (uint8_t *)grow_zeroed(*p, old_bytes, new_bytes);
    if (!q) return -1;
    *p = q;
    return 0;
}

/* room for at least the given totals (reads, CIGAR operations, base slots, name bytes); grows by doubling */
static int batch_reserve(grom_batch *t, batch_caps *c, int64_t reads, int64_t cig, int64_t slots, int64_t names, int keep_names, int with_seq2, int64_t **exc_at)
{
    /* per-read arrays, CIGAR operations and names are written in full by the fill pass: no zero fill (it would touch every page from one
     * thread); only the per-base arrays rely on zero padding, and those are fresh pages of a mapping */
#define GR(ptr, type, oldn, newn) do { (void)(oldn); void *q_ = realloc(t->ptr, (size_t)(newn) * sizeof(type)); if (!q_) return -1; t->ptr = (type *)q_; } while (0)
    if (reads > c->reads) {
        int64_t n = c->reads ? c->reads * 2 : 1; if (n < reads) n = reads;
        const int64_t o = c->reads;
        GR(pos, int32_t, o, n); GR(mpos, int32_t, o, n); GR(tlen, int32_t, o, n); GR(mtid, int32_t, o, n); GR(l_qseq, int32_t, o, n);
        GR(sa_pos, int32_t, o, n); GR(sa_start_adj, int32_t, o, n); GR(sa_end_adj, int32_t, o, n); GR(sa_end_adj_indel, int32_t, o, n);
        GR(flag, uint16_t, o, n); GR(n_cigar, uint16_t, o, n); GR(sa_mapq, int16_t, o, n);
        GR(mapq, uint8_t, o, n); GR(qname_len, uint8_t, o, n); GR(sa_strand, uint8_t, o, n); GR(sa_same_chr, uint8_t, o, n);
        GR(qname_hash, uint64_t, o, n); GR(cigar_off, uint64_t, o, n); GR(base_off, uint64_t, o, n);
        if (keep_names) GR(qname_off, uint64_t, o ? o + 1 : 0, n + 1);
        if (with_seq2) { void *q_ = realloc(*exc_at, (size_t)(n + 1) * sizeof(int64_t)); if (!q_) return -1; *exc_at = (int64_t *)q_; }
        c->reads = n;
    }
    if (cig > c->cig) {
        int64_t n = c->cig ? c->cig * 2 : 1; if (n < cig) n = cig;
        GR(cigar, uint32_t, c->cig, n);
        c->cig = n;
    }
#undef GR
    if (slots > c->slots) {
        int64_t n = c->slots ? c->slots * 2 : 0; if (n < slots) n = slots;
        const size_t o = (size_t)c->slots;
        if (big_grow(&t->qual, &t->big_qual, c->slots ? o + 16 : 0, (size_t)n + 16) < 0) return -1;
        if (big_grow(&t->seq4, &t->big_seq4, c->slots ? o / 2 + 16 : 0, (size_t)n / 2 + 16) < 0) return -1;
        if (with_seq2 && big_grow(&t->seq2, &t->big_seq2, c->slots ? o / 4 + 16 : 0, (size_t)n / 4 + 16) < 0) return -1;
        if (with_seq2 && big_grow(&t->qual2, &t->big_qual2, c->slots ? o / 4 + 16 : 0, (size_t)n / 4 + 16) < 0) return -1;      /* (2-bit qualities: packed window by window) */
        c->slots = n;
    }
    if (keep_names && names > c->names) {
        int64_t n = c->names ? c->names * 2 : 1; if (n < names) n = names;
        char *q = (char *)realloc(t->qname_pool, (size_t)n + 1);
        if (!q) return -1;
        t->qname_pool = q; c->names = n;
    }
    return 0;
}

#define WINDOW_BLOCKS_PER_THREAD 64     /* BGZF blocks a thread inflates per window (4 MB of records) */

/* ---- a target decoded in pieces: the iterator keeps what belongs to the target (its blocks, the window of inflated data with whatever the
 * last window left over, how far the record chain has come); every gromhost_bam_iter_next() decodes windows into a new batch until the batch
 * holds at least max_reads records or the target ends.  gromhost_bam_read_target() is one call with no limit. */
struct grom_target_iter {
    grom_bam *b; int tid, keep_names, n_threads, own;
    uint64_t vbeg;
    blkinfo *blk; int64_t nblk, utotal, WB, w0, inflated;
    uint8_t *win; size_t win_len;
    int64_t carry;
    int started, done, n_batches;
    fill_local *loc;
    uint8_t S2[256], SX[256];
};

int gromhost_bam_iter_open(grom_bam *b, int tid, int keep_names, int n_threads, grom_target_iter **out)
{
    if (tid < 0 || tid >= b->n_targets) return fail("target id %d out of range", tid);
    uint64_t vbeg = b->first_voff, vend = UINT64_MAX;
    if (b->has_index) {
        if (b->tgt_beg[tid] == UINT64_MAX) { vbeg = vend = 0; }
        else { vbeg = b->tgt_beg[tid]; vend = b->tgt_end[tid]; }
    }
    const int trace = getenv("GROMHOST_TRACE") != NULL;
    double t_last = trace ? now_ms() : 0;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
    /* the compressed blocks in [vbeg, vend] (the file is mapped: headers and trailers are read in place) */
    blkinfo *blk = NULL; int64_t nblk = 0, capblk = 0;
    const uint8_t *cf = b->map; const int64_t cf_len = b->map_len;
    if (vend != 0) {
        int64_t off = (int64_t)(vbeg >> 16), endoff = (vend == UINT64_MAX) ? INT64_MAX : (int64_t)(vend >> 16);
        while (off <= endoff && off + 18 <= cf_len) {
            int bsize, coff;
            if (bgzf_block_header(cf + off, cf_len - off, &bsize, &coff) < 0) { free(blk); return fail("%s: bad BGZF block at %lld", b->path, (long long)off); }
            if (off + bsize > cf_len) break;                      /* truncated last block */
            const uint32_t isize = rd_u32(cf + off + bsize - 4);
            if (isize > 65536) { free(blk); return fail("%s: bad BGZF block at %lld", b->path, (long long)off); }
            if (nblk == capblk) { capblk = capblk ? capblk * 2 : 1024; blk = (blkinfo *)realloc(blk, capblk * sizeof(blkinfo)); }
            blk[nblk].off = off; blk[nblk].bsize = bsize; blk[nblk].coff = coff; blk[nblk].isize = (int)isize; nblk++;
            off += bsize;
        }
    }
    int64_t utotal = 0;
    for (int64_t i = 0; i < nblk; i++) utotal += blk[i].isize;
    TRACE_MARK("enumerate blocks");
    /* The target is decoded a window of blocks at a time: inflate (every block straight to its place behind what the window before left
     * over), list the records, make room in the batch, fill.  Only one window of inflated data exists at any time (host memory of a call =
     * the batch + one window, not the batch + the whole inflated target), and it is the same, already touched buffer every time. */
    int64_t WB = (int64_t)WINDOW_BLOCKS_PER_THREAD * n_threads;
    if (WB < 128) WB = 128;
    { const char *e = getenv("GROMHOST_WINDOW_BLOCKS"); if (e && atoll(e) > 0) WB = atoll(e); }      /* tests: tiny windows */
    if (WB > nblk) WB = nblk > 0 ? nblk : 1;
    size_t win_len = ((size_t)WB * 65536 + ((size_t)1 << 20) + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    uint8_t *win = (uint8_t *)mmap(NULL, win_len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (win == (uint8_t *)MAP_FAILED) { free(blk); return fail("out of memory (%lld bytes of inflated BAM)", (long long)win_len); }
#ifdef MADV_HUGEPAGE
    if (!getenv("GROMHOST_NO_HUGEPAGE")) madvise(win, win_len, MADV_HUGEPAGE);
#endif
    const char *force = getenv("GROMHOST_INFLATE");
    grom_target_iter *it = (grom_target_iter *)calloc(1, sizeof(*it));
    fill_local *loc = (fill_local *)calloc((size_t)n_threads, sizeof(fill_local));
    if (!it || !loc) { free(it); free(loc); free(blk); munmap(win, win_len); return fail("out of memory (target iterator)"); }
    it->b = b; it->tid = tid; it->keep_names = keep_names; it->n_threads = n_threads;
    it->own = !(force && !strcmp(force, "zlib"));                   /* GROMHOST_INFLATE=zlib: every block through zlib */
    it->vbeg = vbeg; it->blk = blk; it->nblk = nblk; it->utotal = utotal; it->WB = WB; it->win = win; it->win_len = win_len; it->loc = loc;
    for (int v = 0; v < 256; v++) {
        const int hi = v >> 4, lo = v & 15;
        const int th = hi == 1 ? 0 : hi == 2 ? 1 : hi == 4 ? 2 : hi == 8 ? 3 : -1, tl = lo == 1 ? 0 : lo == 2 ? 1 : lo == 4 ? 2 : lo == 8 ? 3 : -1;
        it->S2[v] = (uint8_t)(((th < 0 ? 0 : th) << 2) | (tl < 0 ? 0 : tl)); it->SX[v] = (uint8_t)((th < 0) + (tl < 0));
    }
    *out = it;
    return 0;
}

void gromhost_bam_iter_close(grom_target_iter *it)
{
    if (!it) return;
    for (int k = 0; k < it->n_threads; k++) free(it->loc[k].e);
    free(it->loc); free(it->blk);
    if (it->win) munmap(it->win, it->win_len);
    free(it);
}

int gromhost_bam_iter_next(grom_target_iter *it, int64_t max_reads, grom_batch **out)
{
    if (it->n_batches > 0 && (it->done || it->w0 >= it->nblk)) return 1;          /* nothing left (an empty target still gives one empty batch) */
    if (max_reads <= 0) max_reads = INT64_MAX;
    grom_bam *b = it->b;
    const int tid = it->tid, keep_names = it->keep_names, n_threads = it->n_threads, own = it->own;
    const uint64_t vbeg = it->vbeg;
    blkinfo *blk = it->blk; const int64_t nblk = it->nblk, WB = it->WB, utotal = it->utotal;
    const uint8_t *cf = b->map;
    const uint8_t *S2 = it->S2, *SX = it->SX;
    fill_local *loc = it->loc;
    uint8_t *win = it->win; size_t win_len = it->win_len;
    int64_t carry = it->carry, w0 = it->w0, inflated_before = it->inflated;
    int started = it->started, done = it->done;
    const int trace = getenv("GROMHOST_TRACE") != NULL;
    double t_last = trace ? now_ms() : 0, t_infl = 0, t_walk = 0, t_fill = 0, t_mark = 0;
    for (int k = 0; k < n_threads; k++) { memset(loc[k].present, 0, sizeof(loc[k].present)); loc[k].n = 0; }
    grom_batch *t = (grom_batch *)calloc(1, sizeof(*t));
    if (!t) return fail("out of memory (read batch)");
    t->v.tid = tid;
    batch_caps caps; memset(&caps, 0, sizeof(caps));
    int64_t *exc_at = NULL;
    const int do_seq2 = 1;
    int64_t n_reads = 0, n_cig = 0, n_slots = 0, n_name = 0;
    /* what is collected across the windows besides the arrays: quality values seen, per thread (it->loc); reads with an XP / SA entry, in read order */
    sa_ent *sa_all = NULL; int64_t n_sa = 0, cap_sa = 0;
    /* 2-bit qualities are packed at the end of every window, while the window's qualities are still in cache, with the dictionary of the
     * values seen so far; that stands as long as the set of values does not change afterwards (instruments use a fixed handful, all of
     * which show up in the first window) -- otherwise the whole batch is packed again at the end.  0 nothing packed, 1 packed up to
     * q2_upto under q2_vals, 2 the set changed or is too large */
    int q2_state = 0, q2_nv = 0; uint8_t q2_vals[4] = { 0, 0, 0, 0 }; int64_t q2_upto = 0;
    const char *tname = b->names[tid];
    int rc = 0;
    int64_t inflated_so_far = 0;
#define FAIL_OUT(code) do { rc = (code); goto out; } while (0)
    for (; w0 < nblk && !done && n_reads < max_reads; w0 += WB) {
        const int64_t w1 = w0 + WB < nblk ? w0 + WB : nblk;
        int64_t wbytes = 0;
        for (int64_t i = w0; i < w1; i++) { blk[i].uoff = carry + wbytes; wbytes += blk[i].isize; }
        if ((size_t)(carry + wbytes) + 64 > win_len) {              /* a record longer than the spare room was carried over */
            const size_t n = ((size_t)(carry + wbytes) + 64 + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
            void *q = mremap(win, win_len, n, MREMAP_MAYMOVE);
            if (q == MAP_FAILED) FAIL_OUT(fail("out of memory (%lld bytes of inflated BAM)", (long long)n));
            win = (uint8_t *)q; win_len = n;
        }
        if (trace) t_mark = now_ms();
        int bad = 0;
        #pragma omp parallel num_threads(n_threads)
        {
            struct grom_inflate_ctx *ctx = own ? (struct grom_inflate_ctx *)malloc(grom_inflate_ctx_size()) : NULL;
            if (ctx) grom_inflate_ctx_init(ctx);
            #pragma omp for schedule(dynamic, 8)
            for (int64_t i = w0; i < w1; i++) {
                const uint8_t *blkp = cf + blk[i].off;
                if (bgzf_inflate_block(ctx, blkp + blk[i].coff, blk[i].bsize - blk[i].coff - 8, win + blk[i].uoff, blk[i].isize, rd_u32(blkp + blk[i].bsize - 8)) < 0) {
                    #pragma omp atomic write
                    bad = 1;
                }
            }
            free(ctx);
        }
        if (bad) FAIL_OUT(fail("%s: BGZF inflate failed (corrupt deflate stream or CRC mismatch)", b->path));
        if (trace) { const double x = now_ms(); t_infl += x - t_mark; t_mark = x; }
        const int64_t have = carry + wbytes;
        inflated_so_far += wbytes;
        /* the record chain of the window: count, validate, offsets */
        reclist rl; walk_end we;
        const int wrc = walk_records_parallel(win, have, w0 == 0 ? (int64_t)(vbeg & 0xffff) : 0, started, tid, b->n_targets, keep_names, n_threads, b->path, &rl, &we);
        if (wrc < 0) FAIL_OUT(wrc == -2 ? fail("out of memory (record list)") : -1);
        if (rl.n) started = 1;
        if (trace) { const double x = now_ms(); t_walk += x - t_mark; t_mark = x; }
        /* room for the window's records; the first window sizes the whole batch from its share of the inflated bytes */
        {
            int64_t wr = n_reads + rl.n, wc = n_cig + rl.n_cig, wsl = n_slots + rl.n_slots, wn = n_name + rl.n_name;
            if (caps.reads == 0 && w1 < nblk && inflated_so_far > 0 && wr > 0) {
                /* what is left of the target, were it all like this window -- or the caller's limit plus one window, if that is less */
                double f = 1.03 * (double)(utotal - inflated_before) / (double)inflated_so_far;
                if (max_reads != INT64_MAX && (double)wr * f > (double)max_reads + 2.0 * (double)wr) f = ((double)max_reads + 2.0 * (double)wr) / (double)wr;
                wr = (int64_t)((double)wr * f) + 1024; wc = (int64_t)((double)wc * f) + 1024; wsl = (int64_t)((double)wsl * f) + 32768; wn = (int64_t)((double)wn * f) + 32768;
            }
            if (wr < 1) wr = 1;
            if (wc < 1) wc = 1;
            if (batch_reserve(t, &caps, wr, wc, wsl, wn, keep_names, do_seq2, &exc_at) < 0) { free(rl.recoff); free(rl.cig_off); free(rl.base_off); free(rl.name_off); FAIL_OUT(fail("out of memory (read batch)")); }
        }
        /* fill, parallel over contiguous ranges of the window's reads.  In the same pass over a record: the 2-bit form of its bases
         * (table-driven, two nibbles per lookup) with the number of its non-A/C/G/T bases, the set of quality values seen, and the
         * reads with an XP / SA entry. */
        const int64_t wn_reads = rl.n;
        int fill_oom = 0;
        #pragma omp parallel num_threads(n_threads)
        {
#ifdef _OPENMP
            const int T = omp_get_num_threads(), me = omp_get_thread_num();
#else
            const int T = 1, me = 0;
#endif
            fill_local *L = &loc[me];
            L->n = 0;
            const int64_t i0 = wn_reads * me / T, i1 = wn_reads * (me + 1) / T;
            for (int64_t j = i0; j < i1; j++) {
                const int64_t i = n_reads + j;
                const uint8_t *r = win + rl.recoff[j];
                const int32_t bl = rd_i32(r);
                const uint32_t bmq = rd_u32(r + 12), fnc = rd_u32(r + 16);
                const int l_qname = bmq & 0xff, ncig = fnc & 0xffff; const int32_t lq = rd_i32(r + 20);
                t->pos[i] = rd_i32(r + 8); t->mapq[i] = (bmq >> 8) & 0xff; t->flag[i] = (uint16_t)(fnc >> 16); t->n_cigar[i] = (uint16_t)ncig;
                t->l_qseq[i] = lq; t->mtid[i] = rd_i32(r + 24); t->mpos[i] = rd_i32(r + 28); t->tlen[i] = rd_i32(r + 32);
                const uint64_t c0 = (uint64_t)n_cig + rl.cig_off[j], b0 = (uint64_t)n_slots + rl.base_off[j];
                t->cigar_off[i] = c0; t->base_off[i] = b0;
                const uint8_t *d = r + 36;
                const int nl = (int)strnlen((const char *)d, l_qname);
                t->qname_len[i] = (uint8_t)(nl > 255 ? 255 : nl);
                t->qname_hash[i] = grom_qname_hash((const char *)d, nl);
                if (keep_names) { const uint64_t m0 = (uint64_t)n_name + rl.name_off[j]; t->qname_off[i] = m0; memcpy(t->qname_pool + m0, d, l_qname); }
                memcpy(t->cigar + c0, d + l_qname, (size_t)ncig * 4);
                const uint8_t *sq = d + l_qname + ncig * 4;
                const int nb = (lq + 1) / 2;
                memcpy(t->seq4 + b0 / 2, sq, (size_t)nb);
                const uint8_t *ql = sq + nb;
                memcpy(t->qual + b0, ql, (size_t)lq);
                for (int k = 0; k < lq; k++) L->present[ql[k]] = 1;
                if (do_seq2) {
                    uint8_t *o2 = t->seq2 + b0 / 4; int64_t ne = 0;
                    const int nfull = (lq & 1) ? nb - 1 : nb;                     /* bytes whose two nibbles are both bases */
                    int k = 0;
                    for (; k + 2 <= nfull; k += 2) { const uint8_t x = sq[k], y = sq[k + 1]; o2[k >> 1] = (uint8_t)((S2[x] << 4) | S2[y]); ne += SX[x] + SX[y]; }
                    if (k < nb) {                                                  /* one or two bytes left; the padding nibble of an odd length is no base */
                        uint8_t x = sq[k], y = 0x11;
                        if (k + 1 < nb) y = sq[k + 1];
                        if (lq & 1) { if (k + 1 < nb) y = (uint8_t)((y & 0xf0) | 1); else x = (uint8_t)((x & 0xf0) | 1); }
                        o2[k >> 1] = (uint8_t)((S2[x] << 4) | S2[y]); ne += SX[x] + SX[y];
                    }
                    exc_at[i + 1] = ne;
                }
                const uint8_t *aux = ql + lq;
                const int l_aux = (int)((r + 4 + bl) - aux);
                parse_sa(aux, l_aux, tname, &t->sa_pos[i], &t->sa_strand[i], &t->sa_mapq[i], &t->sa_same_chr[i],
                         &t->sa_start_adj[i], &t->sa_end_adj[i], &t->sa_end_adj_indel[i]);
                if (t->sa_pos[i] != -1 || t->sa_mapq[i] != -1 || t->sa_strand[i] || t->sa_same_chr[i] || t->sa_start_adj[i] || t->sa_end_adj[i] || t->sa_end_adj_indel[i]) {
                    if (L->n == L->cap) {
                        const int64_t nc = L->cap ? L->cap * 2 : 1024;
                        sa_ent *q = (sa_ent *)realloc(L->e, sizeof(sa_ent) * (size_t)nc);
                        if (!q) {
                            #pragma omp atomic write
                            fill_oom = 1;
                            continue;
                        }
                        L->e = q; L->cap = nc;
                    }
                    sa_ent *e = &L->e[L->n++];                /* parse_sa leaves every other read at (-1, 0, -1, 0, 0, 0, 0) */
                    e->idx = (int32_t)i; e->pos = t->sa_pos[i]; e->start_adj = t->sa_start_adj[i]; e->end_adj = t->sa_end_adj[i];
                    e->end_adj_indel = t->sa_end_adj_indel[i]; e->mapq = t->sa_mapq[i]; e->strand = t->sa_strand[i]; e->same_chr = t->sa_same_chr[i];
                }
            }
        }
        if (fill_oom) { free(rl.recoff); free(rl.cig_off); free(rl.base_off); free(rl.name_off); FAIL_OUT(fail("out of memory (SA list)")); }
        /* the threads' SA entries of this window, in thread order = read order */
        for (int k = 0; k < n_threads; k++) {
            if (!loc[k].n) continue;
            if (n_sa + loc[k].n > cap_sa) {
                cap_sa = cap_sa ? cap_sa * 2 : 4096; if (cap_sa < n_sa + loc[k].n) cap_sa = n_sa + loc[k].n;
                sa_ent *q = (sa_ent *)realloc(sa_all, sizeof(sa_ent) * (size_t)cap_sa);
                if (!q) { free(rl.recoff); free(rl.cig_off); free(rl.base_off); free(rl.name_off); FAIL_OUT(fail("out of memory (SA list)")); }
                sa_all = q;
            }
            memcpy(sa_all + n_sa, loc[k].e, sizeof(sa_ent) * (size_t)loc[k].n); n_sa += loc[k].n; loc[k].n = 0;
        }
        n_reads += rl.n; n_cig += rl.n_cig; n_slots += rl.n_slots; n_name += rl.n_name;
        free(rl.recoff); free(rl.cig_off); free(rl.base_off); free(rl.name_off);
        if (q2_state != 2 && n_slots > q2_upto && t->qual2) {
            uint8_t seen[256]; memset(seen, 0, sizeof(seen));
            for (int k = 0; k < n_threads; k++) for (int q = 0; q < 256; q++) seen[q] |= loc[k].present[q];
            int nv = 0; uint8_t vals[4] = { 0, 0, 0, 0 }, inv2[256]; memset(inv2, 0, sizeof(inv2));
            for (int q = 0; q < 256; q++) if (seen[q]) { if (nv < 4) { vals[nv] = (uint8_t)q; inv2[q] = (uint8_t)nv; } nv++; }
            if (nv > 4 || (q2_state == 1 && (nv != q2_nv || memcmp(vals, q2_vals, 4)))) q2_state = 2;
            else if (nv >= 1) {
                const int64_t s0 = q2_upto, s1 = n_slots;
                uint8_t *const q2 = t->qual2; const uint8_t *const qq = t->qual;
                #pragma omp parallel for schedule(static) num_threads(n_threads)
                for (int64_t sidx = s0; sidx < s1; sidx += 4) q2[sidx >> 2] = (uint8_t)((inv2[qq[sidx]] << 6) | (inv2[qq[sidx + 1]] << 4) | (inv2[qq[sidx + 2]] << 2) | inv2[qq[sidx + 3]]);
                q2_state = 1; q2_nv = nv; memcpy(q2_vals, vals, 4); q2_upto = s1;
            }
        }
        if (trace) { const double x = now_ms(); t_fill += x - t_mark; t_mark = x; }
        /* what the walk left: the chain ended (a record past the target, a block_size that cannot be one), or the data of the window
         * did; then the bytes from there on are the head of a record that continues in the next window */
        if (we.term != 5) done = 1;
        else {
            carry = have - we.stop_at;
            if (carry > 0 && w1 < nblk) memmove(win, win + we.stop_at, (size_t)carry);
        }
    }
    if (trace) {
        fprintf(stderr, "[bamio] %-22s %8.2f ms\n[bamio] %-22s %8.2f ms\n[bamio] %-22s %8.2f ms\n", "inflate", t_infl, "record chain", t_walk, "allocate + fill", t_fill);
        t_last = now_ms();
    }
    if (batch_reserve(t, &caps, n_reads > 0 ? n_reads : 1, n_cig > 0 ? n_cig : 1, n_slots, n_name, keep_names, do_seq2, &exc_at) < 0) FAIL_OUT(fail("out of memory (read batch)"));
    if (keep_names) t->qname_off[n_reads] = (uint64_t)n_name;
    t->v.n_reads = n_reads; t->v.n_cigar_total = n_cig; t->v.n_base_slots = n_slots;
    batch_publish(t);
    {
    /* transport-compact forms (include/grom_reads.h GROM_LAYOUT_*), all lossless; the CUDA library rebuilds the canonical device arrays.
     * The offsets are running sums by construction here; base qualities of current instruments take a handful of distinct values, so a
     * 4- or 16-entry dictionary shrinks them to 2 or 4 bits; bases travel as 2 bits + a list of everything that is not A/C/G/T; the
     * first-SA-entry fields exist for a small minority of reads. */
    const int64_t ns = n_slots;
    grom_read_batch *v = &t->v;
    int flags = GROM_LAYOUT_CANONICAL_OFFSETS;
    /* 3. the forms that need a whole-batch decision.  Qualities: <= 4 distinct values on the bases (padding slots aside) -> 2 bits per
     * slot, <= 16 -> 4 bits, else the bytes travel.  inv[0] is 0 either way, so the zero padding slots pack to 0 without a test. */
    int hist[256]; memset(hist, 0, sizeof(hist));
    for (int k = 0; k < n_threads; k++) for (int q = 0; q < 256; q++) hist[q] |= loc[k].present[q];
    int nv = 0, qmode = 0; uint8_t inv[256]; memset(inv, 0, sizeof(inv));
    if (ns > 0) {
        for (int k = 0; k < 256; k++) if (hist[k]) { if (nv < 16) { v->qual_lut[nv] = (uint8_t)k; inv[k] = (uint8_t)nv; } nv++; }
        if (nv >= 1 && nv <= 4 && (t->qual2 || (t->qual2 = (uint8_t *)big_zalloc((size_t)(ns / 4 + 16), &t->big_qual2)))) qmode = 2;
        else if (nv >= 1 && nv <= 16 && !hist[0] && (t->qual4 = (uint8_t *)malloc((size_t)(ns / 2 + 16)))) {
            /* 4-bit form: padding slots (0 in the canonical array) must decode to 0 as well, so 0 takes a dictionary entry */
            if (nv == 16) { free(t->qual4); t->qual4 = NULL; memset(v->qual_lut, 0, 16); }
            else {
                for (int k = nv; k > 0; k--) v->qual_lut[k] = v->qual_lut[k - 1];
                v->qual_lut[0] = 0;
                for (int k = 0; k < 256; k++) if (hist[k]) inv[k]++;
                inv[0] = 0;
                qmode = 4;
            }
        } else if (nv >= 1 && nv <= 16 && hist[0] && (t->qual4 = (uint8_t *)malloc((size_t)(ns / 2 + 16)))) qmode = 4;
        else memset(v->qual_lut, 0, 16);
    }
    t->sa_index = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_sa + 1)); t->sas_pos = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_sa + 1));
    t->sas_start_adj = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_sa + 1)); t->sas_end_adj = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_sa + 1));
    t->sas_end_adj_indel = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_sa + 1)); t->sas_mapq = (int16_t *)malloc(sizeof(int16_t) * (size_t)(n_sa + 1));
    t->sas_strand = (uint8_t *)malloc((size_t)(n_sa + 1)); t->sas_same_chr = (uint8_t *)malloc((size_t)(n_sa + 1));
    const int sa_ok = t->sa_index && t->sas_pos && t->sas_start_adj && t->sas_end_adj && t->sas_end_adj_indel && t->sas_mapq && t->sas_strand && t->sas_same_chr;
    /* the windows' packing stands if it covers the batch under the dictionary the whole batch ends up with */
    const int q2_done = qmode == 2 && q2_state == 1 && q2_upto == ns && q2_nv == nv && !memcmp(q2_vals, v->qual_lut, (size_t)nv);
    if (qmode != 2 && t->qual2) { big_free(t->qual2, t->big_qual2); t->qual2 = NULL; t->big_qual2 = 0; }
    int64_t ne = 0; int seq2_ok = 0;
    if (do_seq2 && ns > 0 && t->seq2 && exc_at) {
        exc_at[0] = 0;
        for (int64_t i = 0; i < n_reads; i++) exc_at[i + 1] += exc_at[i];
        ne = exc_at[n_reads];
        if (ne <= ns / 16) {
            t->seq_exc_slot = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(ne + 1)); t->seq_exc_code = (uint8_t *)malloc((size_t)(ne + 1));
            seq2_ok = t->seq_exc_slot && t->seq_exc_code;
        }
    }
    #pragma omp parallel num_threads(n_threads)
    {
        if (qmode == 2 && !q2_done) {
            #pragma omp for schedule(dynamic, 1 << 16) nowait
            for (int64_t s = 0; s < ns; s += 4) t->qual2[s >> 2] = (uint8_t)((inv[t->qual[s]] << 6) | (inv[t->qual[s + 1]] << 4) | (inv[t->qual[s + 2]] << 2) | inv[t->qual[s + 3]]);
        } else if (qmode == 4) {
            #pragma omp for schedule(dynamic, 1 << 16) nowait
            for (int64_t s = 0; s < ns; s += 2) t->qual4[s >> 1] = (uint8_t)((inv[t->qual[s]] << 4) | inv[t->qual[s + 1]]);
        }
        if (seq2_ok) {
            #pragma omp for schedule(dynamic, 4096) nowait
            for (int64_t i = 0; i < n_reads; i++) {
                if (exc_at[i + 1] == exc_at[i]) continue;
                const uint64_t b0 = t->base_off[i]; const int lq = t->l_qseq[i]; int64_t w = exc_at[i];
                for (int k = 0; k < lq; k++) {
                    const uint64_t sl = b0 + (uint64_t)k;
                    const int code = (t->seq4[sl >> 1] >> ((~sl & 1) << 2)) & 15;
                    if (code != 1 && code != 2 && code != 4 && code != 8) { t->seq_exc_slot[w] = sl; t->seq_exc_code[w] = (uint8_t)code; w++; }
                }
            }
        }
        if (sa_ok) {
            #pragma omp for schedule(static) nowait
            for (int64_t w = 0; w < n_sa; w++) {
                const sa_ent *e = &sa_all[w];
                t->sa_index[w] = e->idx; t->sas_pos[w] = e->pos; t->sas_start_adj[w] = e->start_adj; t->sas_end_adj[w] = e->end_adj;
                t->sas_end_adj_indel[w] = e->end_adj_indel; t->sas_mapq[w] = e->mapq; t->sas_strand[w] = e->strand; t->sas_same_chr[w] = e->same_chr;
            }
        }
    }
    if (qmode == 2) { v->qual2 = t->qual2; flags |= GROM_LAYOUT_QUAL2; }
    else if (qmode == 4) { v->qual4 = t->qual4; flags |= GROM_LAYOUT_QUAL4; }
    if (seq2_ok) {
        v->seq2 = t->seq2; v->n_seq_exc = ne; v->seq_exc_slot = t->seq_exc_slot; v->seq_exc_code = t->seq_exc_code;
        flags |= GROM_LAYOUT_SEQ2;
    }
    if (sa_ok) {
        v->n_sa = n_sa; v->sa_index = t->sa_index; v->sas_pos = t->sas_pos; v->sas_start_adj = t->sas_start_adj; v->sas_end_adj = t->sas_end_adj;
        v->sas_end_adj_indel = t->sas_end_adj_indel; v->sas_mapq = t->sas_mapq; v->sas_strand = t->sas_strand; v->sas_same_chr = t->sas_same_chr;
        flags |= GROM_LAYOUT_SPARSE_SA;
    }
    v->layout_flags = flags;
    }
    TRACE_MARK("compact forms");
out:
    free(exc_at); free(sa_all);
    it->win = win; it->win_len = win_len; it->carry = carry; it->w0 = w0; it->started = started; it->done = done; it->inflated = inflated_before + inflated_so_far;
    if (rc != 0) { gromhost_batch_free(t); return rc; }
    it->n_batches++;
    *out = t;
    return 0;
#undef FAIL_OUT
}

int gromhost_bam_read_target(grom_bam *b, int tid, int keep_names, int n_threads, grom_batch **out)
{
    grom_target_iter *it = NULL;
    int rc = gromhost_bam_iter_open(b, tid, keep_names, n_threads, &it);
    if (rc) return rc;
    rc = gromhost_bam_iter_next(it, 0, out);
    gromhost_bam_iter_close(it);
    return rc;
}

/* ------------------------------------------------------------------ library statistics straight from the file */

/* find_insert_mean (reference src/GROM.c:1205-1318) over the records of the file in file order, without building batches: the BGZF
 * blocks are inflated a window at a time by all threads, only the core fields of the records are looked at, and the scan ends with the
 * window in which the sample fills up (the reference stops reading there as well).  Same accumulator as gromhost_libstats_add, so the
 * result equals feeding it the per-contig batches in contig order. */
int gromhost_bam_library_stats(grom_bam *b, int rd_min_mapq, int n_threads, int *insert_mean, int *lseq, int *insert_min, int *insert_max, int64_t *mapped_reads)
{
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
    const uint8_t *cf = b->map; const int64_t cf_len = b->map_len;
    const int W = 32 * n_threads;                                   /* blocks per window */
    blkinfo *blk = (blkinfo *)malloc(sizeof(blkinfo) * (size_t)W);
    size_t win_cap = (size_t)W * 65536 + (1 << 20);
    uint8_t *win = (uint8_t *)malloc(win_cap);
    int64_t cap = (int64_t)W * 65536 / 36 + 16;
    int32_t *pos = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap), *mpos = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap), *tlen = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap),
            *mtid = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap), *lq = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap), *rt = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
    uint16_t *flag = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)cap); uint8_t *mapq = (uint8_t *)malloc((size_t)cap);
    gromhost_libstats *s = gromhost_libstats_new(rd_min_mapq);
    int rc = 0;
    if (!blk || !win || !pos || !mpos || !tlen || !mtid || !lq || !rt || !flag || !mapq || !s) rc = fail("out of memory (library statistics)");
    int64_t off = (int64_t)(b->first_voff >> 16), skip = (int64_t)(b->first_voff & 0xffff), carry = 0;
    int full = 0, ended = 0;
    while (rc == 0 && !full && !ended) {
        int nb = 0; int64_t utotal = 0;
        while (nb < W && off + 18 <= cf_len) {
            int bsize, coff;
            if (bgzf_block_header(cf + off, cf_len - off, &bsize, &coff) < 0) { rc = fail("%s: bad BGZF block at %lld", b->path, (long long)off); break; }
            if (off + bsize > cf_len) { ended = 1; break; }
            const uint32_t isize = rd_u32(cf + off + bsize - 4);
            if (isize > 65536) { rc = fail("%s: bad BGZF block at %lld", b->path, (long long)off); break; }
            blk[nb].off = off; blk[nb].bsize = bsize; blk[nb].coff = coff; blk[nb].isize = (int)isize; blk[nb].uoff = utotal; utotal += isize; nb++;
            off += bsize;
        }
        if (rc) break;
        if (nb < W) ended = 1;
        if ((size_t)(carry + utotal) + 64 > win_cap) {
            win_cap = (size_t)(carry + utotal) + (1 << 20);
            uint8_t *w2 = (uint8_t *)realloc(win, win_cap);
            if (!w2) { rc = fail("out of memory (library statistics)"); break; }
            win = w2;
        }
        int bad = 0;
        #pragma omp parallel num_threads(n_threads)
        {
            struct grom_inflate_ctx *ctx = (struct grom_inflate_ctx *)malloc(grom_inflate_ctx_size());
            if (ctx) grom_inflate_ctx_init(ctx);
            #pragma omp for schedule(dynamic, 4)
            for (int i = 0; i < nb; i++) {
                const uint8_t *bp = cf + blk[i].off;
                if (bgzf_inflate_block(ctx, bp + blk[i].coff, blk[i].bsize - blk[i].coff - 8, win + carry + blk[i].uoff, blk[i].isize, rd_u32(bp + blk[i].bsize - 8)) < 0) {
                    #pragma omp atomic write
                    bad = 1;
                }
            }
            free(ctx);
        }
        if (bad) { rc = fail("%s: BGZF inflate failed (corrupt deflate stream or CRC mismatch)", b->path); break; }
        const int64_t have = carry + utotal;
        int64_t p = skip, n = 0; skip = 0;
        if (p > have) { skip = p - have; p = have; }                       /* (a header longer than the window) */
        while (p + 36 <= have) {
            const int32_t bl = rd_i32(win + p);
            if (bl < 32) { ended = 1; break; }
            if (p + 4 + (int64_t)bl > have) break;                          /* the record continues in the next window */
            if (n == cap) break;                                            /* cannot happen: cap covers the smallest possible records */
            const uint32_t bmq = rd_u32(win + p + 12), fnc = rd_u32(win + p + 16);
            rt[n] = rd_i32(win + p + 4); pos[n] = rd_i32(win + p + 8); mapq[n] = (uint8_t)((bmq >> 8) & 0xff); flag[n] = (uint16_t)(fnc >> 16);
            lq[n] = rd_i32(win + p + 20); mtid[n] = rd_i32(win + p + 24); mpos[n] = rd_i32(win + p + 28); tlen[n] = rd_i32(win + p + 32);
            n++; p += 4 + bl;
        }
        /* feed the accumulator, one run of equal target id at a time (its pair test compares mtid with the batch's tid) */
        for (int64_t i0 = 0; i0 < n && !full; ) {
            int64_t i1 = i0 + 1;
            while (i1 < n && rt[i1] == rt[i0]) i1++;
            grom_read_batch v; memset(&v, 0, sizeof(v));
            v.n_reads = i1 - i0; v.tid = rt[i0]; v.pos = pos + i0; v.mpos = mpos + i0; v.tlen = tlen + i0; v.mtid = mtid + i0; v.l_qseq = lq + i0; v.flag = flag + i0; v.mapq = mapq + i0;
            full = gromhost_libstats_add(s, &v);
            i0 = i1;
        }
        carry = have - p;
        if (carry > 0 && !ended) {
            if ((size_t)carry > win_cap / 2) {                               /* one record larger than half the window: make room for it */
                win_cap = (size_t)carry * 2 + (size_t)W * 65536 + (1 << 20);
                uint8_t *w2 = (uint8_t *)malloc(win_cap);
                if (!w2) { rc = fail("out of memory (library statistics)"); break; }
                memcpy(w2, win + p, (size_t)carry); free(win); win = w2;
            } else memmove(win, win + p, (size_t)carry);
        } else if (ended) carry = 0;
    }
    if (rc == 0 && gromhost_libstats_finish(s, insert_mean, lseq, insert_min, insert_max, mapped_reads)) rc = fail("%s: no reads to estimate the insert size from", b->path);
    gromhost_libstats_free(s);
    free(blk); free(win); free(pos); free(mpos); free(tlen); free(mtid); free(lq); free(rt); free(flag); free(mapq);
    return rc;
}

/* ------------------------------------------------------------------ writer (tooling) */

typedef struct {
    FILE *f; uint8_t buf[0xff00]; int fill; int64_t faddr; int level; uint8_t *cbuf;
} bgzf_w;

static int bgzf_flush(bgzf_w *w)
{
    z_stream s; memset(&s, 0, sizeof(s));
    if (deflateInit2(&s, w->level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return -1;
    s.next_in = w->buf; s.avail_in = w->fill; s.next_out = w->cbuf + 18; s.avail_out = 65536 - 18 - 8;
    int rc = deflate(&s, Z_FINISH);
    if (rc != Z_STREAM_END) { deflateEnd(&s); return -1; }
    int clen = (int)s.total_out; deflateEnd(&s);
    int bsize = clen + 26;
    static const uint8_t h[12] = { 0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0 };
    memcpy(w->cbuf, h, 12); w->cbuf[12] = 'B'; w->cbuf[13] = 'C'; w->cbuf[14] = 2; w->cbuf[15] = 0;
    w->cbuf[16] = (uint8_t)((bsize - 1) & 0xff); w->cbuf[17] = (uint8_t)((bsize - 1) >> 8);
    uint32_t crc = (uint32_t)crc32(crc32(0L, NULL, 0), w->buf, w->fill);
    uint8_t *t = w->cbuf + 18 + clen;
    t[0] = crc & 0xff; t[1] = (crc >> 8) & 0xff; t[2] = (crc >> 16) & 0xff; t[3] = (crc >> 24) & 0xff;
    uint32_t is = (uint32_t)w->fill;
    t[4] = is & 0xff; t[5] = (is >> 8) & 0xff; t[6] = (is >> 16) & 0xff; t[7] = (is >> 24) & 0xff;
    if ((int)fwrite(w->cbuf, 1, bsize, w->f) != bsize) return -1;
    w->faddr += bsize; w->fill = 0;
    return 0;
}
static int bgzf_write(bgzf_w *w, const void *src, int n)
{
    const uint8_t *s = (const uint8_t *)src;
    while (n > 0) {
        int k = (int)sizeof(w->buf) - w->fill; if (k > n) k = n;
        memcpy(w->buf + w->fill, s, k); w->fill += k; s += k; n -= k;
        if (w->fill == (int)sizeof(w->buf) && bgzf_flush(w) < 0) return -1;
    }
    return 0;
}
static inline uint64_t bgzf_tell(const bgzf_w *w) { return ((uint64_t)w->faddr << 16) | (uint64_t)w->fill; }

static int reg2bin(int64_t beg, int64_t end)
{
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

typedef struct { uint64_t *be; int n, cap; } chunklist;
#define BAI_NBIN 37450

int gromhost_bam_write(const char *path, int n_targets, const char *const *names, const int64_t *lens,
                       int n_batches, const grom_read_batch *batches,
                       const uint64_t *const *aux_off, const uint8_t *const *aux_pool, int level)
{
    FILE *f = fopen(path, "wb");
    if (!f) return fail("cannot create %s", path);
    bgzf_w *w = (bgzf_w *)calloc(1, sizeof(*w));
    w->f = f; w->level = level <= 0 ? 1 : level; w->cbuf = (uint8_t *)malloc(65536 + 64);
    /* header */
    char *text = (char *)malloc(64 + (size_t)n_targets * 128); int lt = 0;
    lt += sprintf(text + lt, "@HD\tVN:1.5\tSO:coordinate\n");
    for (int i = 0; i < n_targets; i++) lt += sprintf(text + lt, "@SQ\tSN:%s\tLN:%lld\n", names[i], (long long)lens[i]);
    int32_t v = lt;
    bgzf_write(w, "BAM\1", 4); bgzf_write(w, &v, 4); bgzf_write(w, text, lt); free(text);
    v = n_targets; bgzf_write(w, &v, 4);
    for (int i = 0; i < n_targets; i++) {
        v = (int32_t)strlen(names[i]) + 1; bgzf_write(w, &v, 4); bgzf_write(w, names[i], v);
        v = (int32_t)lens[i]; bgzf_write(w, &v, 4);
    }
    bgzf_flush(w);

    char iname[4096]; snprintf(iname, sizeof(iname), "%s.bai", path);
    FILE *fi = fopen(iname, "wb");
    if (!fi) { fclose(f); return fail("cannot create %s", iname); }
    int32_t nt = n_targets; fwrite("BAI\1", 1, 4, fi); fwrite(&nt, 4, 1, fi);

    int bi = 0; int rc = 0;
    uint8_t *rec = (uint8_t *)malloc(1 << 20);
    for (int tid = 0; tid < n_targets && rc == 0; tid++) {
        const grom_read_batch *bt = NULL;
        if (bi < n_batches && batches[bi].tid == tid) bt = &batches[bi];
        if (!bt || bt->n_reads == 0) {
            int32_t z = 0; fwrite(&z, 4, 1, fi); fwrite(&z, 4, 1, fi);
            if (bt) bi++;
            continue;
        }
        if (!bt->qname_off || !bt->qname_pool) { rc = fail("batch for target %d carries no read names", tid); break; }
        chunklist *bins = (chunklist *)calloc(BAI_NBIN + 1, sizeof(chunklist));
        int64_t nlin = (lens[tid] >> 14) + 2;
        uint64_t *lin = (uint64_t *)calloc((size_t)nlin, 8);
        int cur_bin = -1; uint64_t chunk_beg = 0;
        uint64_t ref_beg = bgzf_tell(w), n_mapped = 0, n_unmapped = 0;
        for (int64_t i = 0; i < bt->n_reads && rc == 0; i++) {
            int nc = bt->n_cigar[i]; int32_t lq = bt->l_qseq[i];
            int64_t qo = (int64_t)bt->qname_off[i]; int l_qname = (int)(bt->qname_off[i + 1] - bt->qname_off[i]);
            const uint32_t *cg = bt->cigar + bt->cigar_off[i];
            int64_t end = bt->pos[i];
            if (!(bt->flag[i] & 4)) for (int k = 0; k < nc; k++) { int op = cg[k] & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) end += cg[k] >> 4; }
            if (end <= bt->pos[i]) end = (int64_t)bt->pos[i] + 1;
            int bin = reg2bin(bt->pos[i], end);
            int l_aux = (aux_off && aux_off[bi]) ? (int)(aux_off[bi][i + 1] - aux_off[bi][i]) : 0;
            int32_t bl = 32 + l_qname + nc * 4 + (lq + 1) / 2 + lq + l_aux;
            if (bl + 4 > (1 << 20)) { rc = fail("record too large"); break; }
            uint8_t *q = rec;
            int32_t x[9];
            x[0] = bl; x[1] = tid; x[2] = bt->pos[i];
            x[3] = (int32_t)(((uint32_t)bin << 16) | ((uint32_t)bt->mapq[i] << 8) | (uint32_t)l_qname);
            x[4] = (int32_t)(((uint32_t)bt->flag[i] << 16) | (uint32_t)nc);
            x[5] = lq; x[6] = bt->mtid[i]; x[7] = bt->mpos[i]; x[8] = bt->tlen[i];
            memcpy(q, x, 36); q += 36;
            memcpy(q, bt->qname_pool + qo, l_qname); q += l_qname;
            memcpy(q, cg, (size_t)nc * 4); q += nc * 4;
            memcpy(q, bt->seq4 + bt->base_off[i] / 2, (size_t)(lq + 1) / 2);
            if (lq & 1) q[(lq + 1) / 2 - 1] &= 0xf0;
            q += (lq + 1) / 2;
            memcpy(q, bt->qual + bt->base_off[i], (size_t)lq); q += lq;
            if (l_aux) { memcpy(q, aux_pool[bi] + aux_off[bi][i], l_aux); q += l_aux; }
            uint64_t vo = bgzf_tell(w);
            if (w->fill == (int)sizeof(w->buf)) vo = ((uint64_t)w->faddr << 16);
            if (bin != cur_bin) {
                if (cur_bin >= 0) {
                    chunklist *c = &bins[cur_bin];
                    if (c->n == c->cap) { c->cap = c->cap ? c->cap * 2 : 4; c->be = (uint64_t *)realloc(c->be, (size_t)c->cap * 16); }
                    c->be[2 * c->n] = chunk_beg; c->be[2 * c->n + 1] = vo; c->n++;
                }
                cur_bin = bin; chunk_beg = vo;
            }
            for (int64_t wdw = bt->pos[i] >> 14; wdw <= (end - 1) >> 14 && wdw < nlin; wdw++)
                if (lin[wdw] == 0 || vo < lin[wdw]) lin[wdw] = vo;
            if (bt->flag[i] & 4) n_unmapped++; else n_mapped++;
            if (bgzf_write(w, rec, 4 + bl) < 0) rc = fail("write failed");
        }
        uint64_t ref_end = bgzf_tell(w);
        if (cur_bin >= 0) {
            chunklist *c = &bins[cur_bin];
            if (c->n == c->cap) { c->cap = c->cap ? c->cap * 2 : 4; c->be = (uint64_t *)realloc(c->be, (size_t)c->cap * 16); }
            c->be[2 * c->n] = chunk_beg; c->be[2 * c->n + 1] = ref_end; c->n++;
        }
        int32_t n_bin = 1;
        for (int k = 0; k < BAI_NBIN; k++) if (bins[k].n) n_bin++;
        fwrite(&n_bin, 4, 1, fi);
        for (int k = 0; k < BAI_NBIN; k++) if (bins[k].n) {
            uint32_t bk = (uint32_t)k; int32_t ncnk = bins[k].n;
            fwrite(&bk, 4, 1, fi); fwrite(&ncnk, 4, 1, fi); fwrite(bins[k].be, 16, (size_t)ncnk, fi);
            free(bins[k].be);
        }
        { uint32_t bk = BAI_NBIN; int32_t two = 2; uint64_t m[4] = { ref_beg, ref_end, n_mapped, n_unmapped };
          fwrite(&bk, 4, 1, fi); fwrite(&two, 4, 1, fi); fwrite(m, 8, 4, fi); }
        int64_t last = 0;
        for (int64_t k = 0; k < nlin; k++) if (lin[k]) last = k + 1;
        for (int64_t k = 1; k < last; k++) if (lin[k] == 0) lin[k] = lin[k - 1];
        int32_t n_intv = (int32_t)last;
        fwrite(&n_intv, 4, 1, fi); fwrite(lin, 8, (size_t)last, fi);
        free(lin); free(bins);
        bi++;
    }
    free(rec);
    if (w->fill) bgzf_flush(w);
    static const uint8_t eof[28] = { 0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 0x42, 0x43, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    fwrite(eof, 1, 28, f);
    fclose(f); fclose(fi); free(w->cbuf); free(w);
    return rc;
}
