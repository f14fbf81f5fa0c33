/* oracle/hooks.c -- TEST INFRASTRUCTURE ONLY: sinks for the dump hooks of hooks.h.
 * GROM_DUMP_DIR=<dir> switches dumping on; GROM_SEED=<n> pins the reference's
 * srand(time) (reference src/GROM.c:1584) so CNV reservoir sampling is repeatable. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "hooks.h"

int g_hook_on = -1;
static char g_dir[2048];
static FILE *g_scan, *g_reads;
static char g_scan_chr[256], g_reads_chr[256];

static void hook_init(void) __attribute__((constructor));
static void hook_init(void)
{
    const char *d = getenv("GROM_DUMP_DIR");
    g_hook_on = (d && *d) ? 1 : 0;
    if (g_hook_on) snprintf(g_dir, sizeof(g_dir), "%s", d);
}

static FILE *open_for(const char *kind, const char *chr)
{
    char p[4096];
    snprintf(p, sizeof(p), "%s/%s_%s.bin", g_dir, kind, chr);
    FILE *f = fopen(p, "wb");
    if (!f) { fprintf(stderr, "hooks: cannot create %s\n", p); exit(3); }
    setvbuf(f, NULL, _IOFBF, 1 << 22);
    return f;
}

void grom_hook_scan(const char *chr, int pos, const int *v, const double *d)
{
    if (!g_scan || strcmp(chr, g_scan_chr)) {
        if (g_scan) fclose(g_scan);
        snprintf(g_scan_chr, sizeof(g_scan_chr), "%s", chr);
        g_scan = open_for("scan", chr);
    }
    fwrite(&pos, 4, 1, g_scan); fwrite(v, 4, GH_NI, g_scan); fwrite(d, 8, GH_ND, g_scan);
}

void grom_hook_read(const char *chr, int pos, int mpos, int tlen, int flag, int mapq, int keep)
{
    if (!g_reads || strcmp(chr, g_reads_chr)) {
        if (g_reads) fclose(g_reads);
        snprintf(g_reads_chr, sizeof(g_reads_chr), "%s", chr);
        g_reads = open_for("reads", chr);
    }
    int r[6] = { pos, mpos, tlen, flag, mapq, keep };
    fwrite(r, 4, 6, g_reads);
}

void grom_hook_depth(const char *chr, long len, const int *mq, const int *rd, const int *low)
{
    /* the per-position scan of this contig is complete once the depth arrays are final */
    if (g_scan) { fclose(g_scan); g_scan = NULL; }
    if (g_reads) { fclose(g_reads); g_reads = NULL; }
    FILE *f = open_for("depth", chr);
    fwrite(mq, 4, len, f); fwrite(rd, 4, len, f); fwrite(low, 4, len, f);
    fclose(f);
}

void grom_hook_gc(const char *chr, long len, const int *gc, const int *acgt)
{
    FILE *f = open_for("gc", chr);
    fwrite(gc, 4, len, f); fwrite(acgt, 4, len, f);
    fclose(f);
}

void grom_hook_srand(unsigned seed)
{
    const char *s = getenv("GROM_SEED");
    srand(s ? (unsigned)strtoul(s, NULL, 10) : seed);
}
