/* fuzz_bamio.c -- robustness of the host batcher against damaged files (tooling, not part of the library build).
 *
 *   gcc -O1 -g -fsanitize=address,undefined -fopenmp -Iinclude tools/fuzz_bamio.c grom_b200/host/{bamio,inflate,fasta,libstats}.c -lz -lm -o /tmp/fuzz_bamio
 *   GROMHOST_WALK_PAR_MIN=1 /tmp/fuzz_bamio tests/golden/g1.bam 2000
 *
 * Takes a good BAM, inflates it, and per iteration damages either the record stream (then re-deflates it into valid BGZF blocks with
 * correct CRCs, so the damage reaches the record walk / fill / SA parser) or the compressed file itself (reaches the block enumeration,
 * the DEFLATE decoder and the CRC check), writes it beside its index and runs open + library statistics + every target through the
 * batcher.  Every outcome but a crash / sanitizer report is fine: an error return, or a batch whose arrays are then read in full. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <unistd.h>
#include <zlib.h>
#include "gromhost.h"

static uint64_t rs = 88172645463325252ULL;
static uint32_t rnd(void) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 16); }

static uint8_t *slurp(const char *p, size_t *n)
{
    FILE *f = fopen(p, "rb"); if (!f) { perror(p); exit(2); }
    fseek(f, 0, SEEK_END); *n = (size_t)ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t *b = malloc(*n + 1); if (fread(b, 1, *n, f) != *n) exit(2); fclose(f); return b;
}
static size_t bgzf_pack(const uint8_t *raw, size_t n, uint8_t *out, size_t chunk)
{
    size_t w = 0;
    for (size_t o = 0; o <= n; o += chunk) {
        const size_t len = (o + chunk <= n) ? chunk : n - o;
        z_stream s; memset(&s, 0, sizeof(s));
        deflateInit2(&s, 1, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        s.next_in = (Bytef *)(raw + o); s.avail_in = (uInt)len; s.next_out = out + w + 18; s.avail_out = 70000;
        deflate(&s, Z_FINISH);
        const size_t cl = s.total_out; deflateEnd(&s);
        static const uint8_t h[16] = { 0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0 };
        memcpy(out + w, h, 16);
        const size_t bs = cl + 26;
        out[w + 16] = (uint8_t)((bs - 1) & 0xff); out[w + 17] = (uint8_t)((bs - 1) >> 8);
        const uint32_t crc = (uint32_t)crc32(0, raw + o, (uInt)len), is = (uint32_t)len;
        memcpy(out + w + 18 + cl, &crc, 4); memcpy(out + w + 22 + cl, &is, 4);
        w += bs;
        if (len == 0) break;
        if (o + chunk >= n) { /* EOF marker next */ }
    }
    return w;
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: fuzz_bamio good.bam iterations [seed]\n"); return 2; }
    const int iters = atoi(argv[2]);
    if (argc > 3) rs ^= (uint64_t)atoll(argv[3]) * 0x9E3779B97F4A7C15ULL;
    size_t fn; uint8_t *file = slurp(argv[1], &fn);
    char ip[4096]; snprintf(ip, sizeof(ip), "%s.bai", argv[1]);
    size_t in_ = 0; uint8_t *idx = access(ip, R_OK) == 0 ? slurp(ip, &in_) : NULL;
    /* inflate the whole file */
    uint8_t *raw = malloc(fn * 12 + 65536); size_t rn = 0;
    for (size_t o = 0; o + 28 <= fn;) {
        const size_t bs = (size_t)(file[o + 16] | (file[o + 17] << 8)) + 1; uint32_t is; memcpy(&is, file + o + bs - 4, 4);
        z_stream s; memset(&s, 0, sizeof(s)); inflateInit2(&s, -15);
        s.next_in = file + o + 18; s.avail_in = (uInt)(bs - 26); s.next_out = raw + rn; s.avail_out = 65536; inflate(&s, Z_FINISH); inflateEnd(&s);
        rn += is; o += bs;
    }
    uint8_t *mut = malloc(rn + 16), *packed = malloc(rn * 2 + (rn / 1000 + 8) * 64 + 65536);
    char tmp[256], tmpi[256];
    snprintf(tmp, sizeof(tmp), "/tmp/fuzz_bamio_case_%d.bam", (int)getpid()); snprintf(tmpi, sizeof(tmpi), "%s.bai", tmp);
    long ok = 0, failed = 0, reads = 0;
    for (int it = 0; it < iters; it++) {
        size_t pn;
        const int mode = (int)(rnd() % 4);
        if (mode < 3) {                                     /* damage the record stream */
            memcpy(mut, raw, rn);
            size_t n = rn;
            const int k = 1 + (int)(rnd() % 6);
            for (int j = 0; j < k; j++) {
                const size_t at = rnd() % n;
                switch (rnd() % 5) {
                case 0: mut[at] ^= (uint8_t)(1u << (rnd() % 8)); break;
                case 1: mut[at] = (uint8_t)rnd(); break;
                case 2: { uint32_t v = rnd() % 3 == 0 ? 0x7fffffffu : rnd() % 3 == 1 ? 0xffffffffu : rnd(); memcpy(mut + (at & ~(size_t)3), &v, at + 4 <= n ? 4 : 1); break; }
                case 3: mut[at] = 0; break;
                default: if (rnd() % 4 == 0) n = at + 1; break;      /* truncate */
                }
            }
            pn = bgzf_pack(mut, n, packed, 1000 + rnd() % 64000);
        } else {                                            /* damage the compressed file */
            memcpy(packed, file, fn); pn = fn;
            const int k = 1 + (int)(rnd() % 3);
            for (int j = 0; j < k; j++) { const size_t at = rnd() % pn; if (rnd() % 5 == 0) pn = at + 1; else packed[at] ^= (uint8_t)(1u << (rnd() % 8)); }
        }
        FILE *f = fopen(tmp, "wb"); fwrite(packed, 1, pn, f); fclose(f);
        if (idx && rnd() % 2) { f = fopen(tmpi, "wb"); fwrite(idx, 1, in_, f); fclose(f); } else unlink(tmpi);
        grom_bam *b = NULL;
        if (gromhost_bam_open(tmp, &b)) { failed++; continue; }
        int m, l, lo, hi; int64_t mp;
        if (gromhost_bam_library_stats(b, 20, 1 + (int)(rnd() % 3), &m, &l, &lo, &hi, &mp)) failed++;
        const int nt = gromhost_bam_n_targets(b);
        for (int t = 0; t < nt && t < 8; t++) {
            grom_batch *bt = NULL;
            if (gromhost_bam_read_target(b, t, (int)(rnd() % 2), 1 + (int)(rnd() % 4), &bt)) { failed++; continue; }
            grom_read_batch v; gromhost_batch_view(bt, &v);
            /* touch everything the view promises */
            uint64_t acc = 0;
            for (int64_t i = 0; i < v.n_reads; i++) {
                acc += (uint64_t)v.pos[i] + v.flag[i] + v.mapq[i] + v.qname_hash[i] + (uint64_t)v.sa_pos[i];
                for (int c = 0; c < v.n_cigar[i]; c++) acc += v.cigar[v.cigar_off[i] + (uint64_t)c];
                for (int q = 0; q < v.l_qseq[i]; q++) acc += v.qual[v.base_off[i] + (uint64_t)q] + v.seq4[(v.base_off[i] + (uint64_t)q) >> 1];
            }
            for (int64_t i = 0; i < v.n_seq_exc; i++) acc += v.seq_exc_slot[i] + v.seq_exc_code[i];
            for (int64_t i = 0; i < v.n_sa; i++) acc += (uint64_t)v.sa_index[i] + (uint64_t)v.sas_pos[i];
            if (v.qual2) for (int64_t sidx = 0; sidx < v.n_base_slots / 4; sidx++) acc += v.qual2[sidx];
            if (v.seq2) for (int64_t sidx = 0; sidx < v.n_base_slots / 4; sidx++) acc += v.seq2[sidx];
            reads += v.n_reads + (long)(acc & 1);
            ok++;
            gromhost_batch_free(bt);
        }
        /* the same target in pieces */
        if (nt > 0 && rnd() % 2) {
            grom_target_iter *itr = NULL;
            if (gromhost_bam_iter_open(b, (int)(rnd() % (uint32_t)nt), (int)(rnd() % 2), 1 + (int)(rnd() % 3), &itr) == 0) {
                const int64_t lim = 1 + (int64_t)(rnd() % 3000);
                for (int guard = 0; guard < 100000; guard++) {
                    grom_batch *bt = NULL;
                    const int rc = gromhost_bam_iter_next(itr, lim, &bt);
                    if (rc == 1) break;
                    if (rc) { failed++; break; }
                    grom_read_batch v; gromhost_batch_view(bt, &v);
                    uint64_t acc = 0;
                    for (int64_t i = 0; i < v.n_reads; i++) { acc += (uint64_t)v.pos[i]; for (int q = 0; q < v.l_qseq[i]; q++) acc += v.qual[v.base_off[i] + (uint64_t)q]; }
                    reads += v.n_reads + (long)(acc & 1); ok++;
                    gromhost_batch_free(bt);
                }
                gromhost_bam_iter_close(itr);
            } else failed++;
        }
        gromhost_bam_close(b);
    }
    printf("%d cases: %ld targets decoded (%ld reads), %ld calls refused\n", iters, ok, reads, failed);
    unlink(tmp); unlink(tmpi);
    free(file); free(idx); free(raw); free(mut); free(packed);
    return 0;
}
