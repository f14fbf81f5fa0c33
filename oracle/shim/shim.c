/* oracle/shim/shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * zlib-only implementation of the legacy samtools entry points that the
 * reference names (reference src/GROM.c:216-258, 989, 1221-1269, 5759-5795,
 * 20474, 22118-22138): BGZF block reader with virtual-offset seek, BAM header
 * and record reader, BAI loader (only the smallest chunk offset per target is
 * kept) and a bam_fetch that seeks there and streams the target's records.
 * Written for this repository; it shares no code with samtools/htslib.
 */
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include "sam.h"

const char bam_nt16_rev_table[] = "=ACMGRSVTWYHKDBN";

struct shim_bgzf {
    FILE *f;
    uint8_t *raw;      /* compressed block */
    uint8_t *blk;      /* inflated block */
    int blk_len;       /* bytes valid in blk */
    int blk_off;       /* read cursor in blk */
    int64_t blk_addr;  /* file offset of the block held in blk */
    int eof;
};

struct shim_bai {
    int n_ref;
    uint64_t *first;   /* smallest chunk_beg per target, UINT64_MAX if none */
};

static int bgzf_load(shim_bgzf_t *z)
{
    uint8_t hdr[18];
    z->blk_addr = ftell(z->f);
    z->blk_len = z->blk_off = 0;
    if (fread(hdr, 1, 18, z->f) != 18) { z->eof = 1; return -1; }
    if (hdr[0] != 0x1f || hdr[1] != 0x8b) { z->eof = 1; return -1; }
    /* assumes the canonical single "BC" extra subfield (xlen == 6) */
    int xlen = hdr[10] | (hdr[11] << 8);
    int bsize = (hdr[16] | (hdr[17] << 8)) + 1;
    if (xlen != 6) { z->eof = 1; return -1; }
    int clen = bsize - 18;
    if ((int)fread(z->raw, 1, clen, z->f) != clen) { z->eof = 1; return -1; }
    uint32_t isize = z->raw[clen - 4] | (z->raw[clen - 3] << 8) | (z->raw[clen - 2] << 16) | ((uint32_t)z->raw[clen - 1] << 24);
    z_stream s;
    memset(&s, 0, sizeof(s));
    s.next_in = z->raw; s.avail_in = clen - 8;
    s.next_out = z->blk; s.avail_out = 65536;
    if (inflateInit2(&s, -15) != Z_OK) { z->eof = 1; return -1; }
    int rc = inflate(&s, Z_FINISH);
    inflateEnd(&s);
    if (rc != Z_STREAM_END || s.total_out != isize) { z->eof = 1; return -1; }
    z->blk_len = (int)isize;
    return 0;
}

static int bgzf_read(shim_bgzf_t *z, void *dst, int n)
{
    uint8_t *d = (uint8_t *)dst;
    int got = 0;
    while (got < n) {
        if (z->blk_off >= z->blk_len) {
            if (z->eof) break;
            if (bgzf_load(z) < 0) break;
            if (z->blk_len == 0) continue;   /* empty (EOF marker) block */
        }
        int k = z->blk_len - z->blk_off;
        if (k > n - got) k = n - got;
        memcpy(d + got, z->blk + z->blk_off, k);
        z->blk_off += k; got += k;
    }
    return got;
}

static void bgzf_seek(shim_bgzf_t *z, uint64_t voff)
{
    z->eof = 0;
    fseek(z->f, (long)(voff >> 16), SEEK_SET);
    if (bgzf_load(z) == 0) z->blk_off = (int)(voff & 0xffff);
}

bamFile bam_open(const char *fn, const char *mode)
{
    (void)mode;
    FILE *f = fopen(fn, "rb");
    if (!f) return NULL;
    shim_bgzf_t *z = (shim_bgzf_t *)calloc(1, sizeof(*z));
    z->f = f;
    z->raw = (uint8_t *)malloc(65536 + 64);
    z->blk = (uint8_t *)malloc(65536);
    return z;
}

int bam_close(bamFile z)
{
    if (!z) return 0;
    fclose(z->f); free(z->raw); free(z->blk); free(z);
    return 0;
}

bam_header_t *bam_header_read(bamFile z)
{
    char magic[4];
    int32_t l_text, n_ref, i;
    if (bgzf_read(z, magic, 4) != 4 || memcmp(magic, "BAM\1", 4)) return NULL;
    bam_header_t *h = (bam_header_t *)calloc(1, sizeof(*h));
    bgzf_read(z, &l_text, 4);
    h->l_text = l_text;
    h->text = (char *)calloc(l_text + 1, 1);
    bgzf_read(z, h->text, l_text);
    bgzf_read(z, &n_ref, 4);
    h->n_targets = n_ref;
    h->target_name = (char **)calloc(n_ref, sizeof(char *));
    h->target_len = (uint32_t *)calloc(n_ref, sizeof(uint32_t));
    for (i = 0; i < n_ref; i++) {
        int32_t l_name;
        bgzf_read(z, &l_name, 4);
        h->target_name[i] = (char *)calloc(l_name + 1, 1);
        bgzf_read(z, h->target_name[i], l_name);
        bgzf_read(z, &h->target_len[i], 4);
    }
    return h;
}

void bam_header_destroy(bam_header_t *h)
{
    int i;
    if (!h) return;
    for (i = 0; i < h->n_targets; i++) free(h->target_name[i]);
    free(h->target_name); free(h->target_len); free(h->text); free(h);
}

bam1_t *bam_init1(void) { return (bam1_t *)calloc(1, sizeof(bam1_t)); }
void bam_destroy1(bam1_t *b) { if (b) { free(b->data); free(b); } }

int bam_read1(bamFile z, bam1_t *b)
{
    int32_t block_len;
    uint32_t x[8];
    if (bgzf_read(z, &block_len, 4) != 4) return -1;
    if (bgzf_read(z, x, 32) != 32) return -3;
    b->core.tid = (int32_t)x[0];
    b->core.pos = (int32_t)x[1];
    b->core.bin = x[2] >> 16; b->core.qual = (x[2] >> 8) & 0xff; b->core.l_qname = x[2] & 0xff;
    b->core.flag = x[3] >> 16; b->core.n_cigar = x[3] & 0xffff;
    b->core.l_qseq = (int32_t)x[4];
    b->core.mtid = (int32_t)x[5];
    b->core.mpos = (int32_t)x[6];
    b->core.isize = (int32_t)x[7];
    b->l_data = block_len - 32;
    if (b->m_data < b->l_data) {
        b->m_data = (b->l_data + 31) / 32 * 32;
        b->data = (uint8_t *)realloc(b->data, b->m_data);
    }
    if (bgzf_read(z, b->data, b->l_data) != b->l_data) return -4;
    return 4 + block_len;
}

samfile_t *samopen(const char *fn, const char *mode, const void *aux)
{
    (void)aux;
    bamFile z = bam_open(fn, mode);
    if (!z) return NULL;
    samfile_t *s = (samfile_t *)calloc(1, sizeof(*s));
    s->fp = z;
    s->header = bam_header_read(z);
    if (!s->header) { bam_close(z); free(s); return NULL; }
    return s;
}

int samread(samfile_t *s, bam1_t *b) { return bam_read1(s->fp, b); }

void samclose(samfile_t *s)
{
    if (!s) return;
    bam_header_destroy(s->header); bam_close(s->fp); free(s);
}

bam_index_t *bam_index_load(const char *fn)
{
    char name[4096];
    snprintf(name, sizeof(name), "%s.bai", fn);
    FILE *f = fopen(name, "rb");
    if (!f) return NULL;
    char magic[4];
    int32_t n_ref, i, j;
    if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "BAI\1", 4)) { fclose(f); return NULL; }
    if (fread(&n_ref, 4, 1, f) != 1) { fclose(f); return NULL; }
    bam_index_t *idx = (bam_index_t *)calloc(1, sizeof(*idx));
    idx->n_ref = n_ref;
    idx->first = (uint64_t *)malloc(sizeof(uint64_t) * (n_ref > 0 ? n_ref : 1));
    for (i = 0; i < n_ref; i++) {
        int32_t n_bin, n_intv;
        idx->first[i] = UINT64_MAX;
        if (fread(&n_bin, 4, 1, f) != 1) break;
        for (j = 0; j < n_bin; j++) {
            uint32_t bin; int32_t n_chunk, k;
            if (fread(&bin, 4, 1, f) != 1 || fread(&n_chunk, 4, 1, f) != 1) break;
            for (k = 0; k < n_chunk; k++) {
                uint64_t be[2];
                if (fread(be, 8, 2, f) != 2) break;
                if (bin != 37450 && be[0] < idx->first[i]) idx->first[i] = be[0];
            }
        }
        if (fread(&n_intv, 4, 1, f) != 1) break;
        fseek(f, (long)n_intv * 8, SEEK_CUR);
    }
    fclose(f);
    return idx;
}

void bam_index_destroy(bam_index_t *idx)
{
    if (idx) { free(idx->first); free(idx); }
}

static int32_t ref_end(const bam1_t *b)
{
    const uint32_t *c = bam1_cigar(b);
    int32_t e = b->core.pos;
    uint32_t k;
    if (b->core.flag & BAM_FUNMAP) return e + 1;
    for (k = 0; k < b->core.n_cigar; k++) {
        int op = bam_cigar_op(c[k]);
        if (op == BAM_CMATCH || op == BAM_CDEL || op == BAM_CREF_SKIP || op == BAM_CEQUAL || op == BAM_CDIFF)
            e += bam_cigar_oplen(c[k]);
    }
    return e > b->core.pos ? e : b->core.pos + 1;
}

int bam_fetch(bamFile z, const bam_index_t *idx, int tid, int beg, int end, void *data, bam_fetch_f func)
{
    if (!idx || tid < 0 || tid >= idx->n_ref || idx->first[tid] == UINT64_MAX) return 0;
    bam1_t *b = bam_init1();
    bgzf_seek(z, idx->first[tid]);
    while (bam_read1(z, b) > 0) {
        if (b->core.tid != tid || b->core.pos >= end) break;
        if (ref_end(b) > beg) func(b, data);
    }
    bam_destroy1(b);
    return 0;
}

uint8_t *bam_aux_get(const bam1_t *b, const char tag[2])
{
    uint8_t *s = bam1_aux(b), *e = b->data + b->l_data;
    while (s + 3 <= e) {
        int hit = (s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1]);
        uint8_t t = s[2];
        uint8_t *v = s + 2;          /* points at the type byte, like the original API */
        s += 3;
        if (hit) return v;
        switch (t) {
        case 'A': case 'c': case 'C': s += 1; break;
        case 's': case 'S': s += 2; break;
        case 'i': case 'I': case 'f': s += 4; break;
        case 'd': s += 8; break;
        case 'Z': case 'H': while (s < e && *s) s++; s++; break;
        case 'B': {
            if (s + 5 > e) return NULL;
            uint8_t st = s[0]; uint32_t n; memcpy(&n, s + 1, 4);
            int w = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
            s += 5 + (size_t)n * w; break; }
        default: return NULL;
        }
    }
    return NULL;
}
