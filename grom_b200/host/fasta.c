/* fasta.c -- reference FASTA: header index and per-contig character load.
 *
 * Restates the reference's two FASTA passes with their exact line rules, because the characters (case included) and the contig length
 * feed the hot path directly:
 *   - the index pass (reference src/GROM.c:1332-1417): every line that begins with '>' opens a contig; its name is the header's first
 *     word (up to the first character that is not isgraph), lower-cased, at most 49 characters;
 *   - the load of one contig (src/GROM.c:21011-21045): lines up to the next '>' line; from every line the characters up to and including
 *     its last alphabetic one are kept -- but that cut is only re-evaluated when the line's length differs from the previous line's
 *     (first line: always), and a line without any alphabetic character keeps one character.
 * Both passes of the reference read with fgets into a 1000-byte buffer, so a "line" is at most 999 characters; longer physical lines are
 * seen as several lines.  That rule is kept too (fa_line below).  The file is mapped; lines are found with memchr and copied with
 * memcpy, so loading a contig runs at memory speed instead of a character at a time.
 */
#include <ctype.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>
#include "gromhost.h"

#define FA_LINE_MAX 999            /* fgets(line, 1000, f), src/GROM.c:20451 and 1332 */
#define FA_NAME_MAX 49             /* max_chr_name_len - 1, src/GROM.c:636, 1372-1380 */

int gromhost_fail(const char *fmt, ...);      /* bamio.c: sets gromhost_last_error(), returns -1 */

typedef struct { char name[FA_NAME_MAX + 1]; int64_t seq_off, seq_end; } fa_contig;
struct grom_fasta {
    char *path;
    const char *map; int64_t len; int is_mmap;
    fa_contig *c; int n, cap;
};

/* the line fgets would return at offset p: up to and including the next '\n', at most FA_LINE_MAX characters; returns its length */
static inline int64_t fa_line(const char *m, int64_t p, int64_t end)
{
    int64_t room = end - p;
    if (room > FA_LINE_MAX) room = FA_LINE_MAX;
    const char *nl = (const char *)memchr(m + p, '\n', (size_t)room);
    return nl ? (int64_t)(nl - (m + p)) + 1 : room;
}

int gromhost_fasta_open(const char *path, grom_fasta **out)
{
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return gromhost_fail("Could not open %s", path);
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return gromhost_fail("%s: cannot stat", path); }
    grom_fasta *fa = (grom_fasta *)calloc(1, sizeof(*fa));
    fa->path = strdup(path); fa->len = (int64_t)st.st_size;
    if (fa->len > 0) {
        void *m = mmap(NULL, (size_t)fa->len, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m != MAP_FAILED) { fa->map = (const char *)m; fa->is_mmap = 1; }
        else {
            char *buf = (char *)malloc((size_t)fa->len);
            int64_t got = 0;
            while (buf && got < fa->len) { const ssize_t r = read(fd, buf + got, (size_t)(fa->len - got)); if (r <= 0) break; got += r; }
            if (!buf || got != fa->len) { free(buf); close(fd); free(fa->path); free(fa); return gromhost_fail("%s: cannot map or read", path); }
            fa->map = buf;
        }
    }
    close(fd);
    /* index pass: only the first character of every line matters, so the body is crossed with memchr */
    const char *m = fa->map; const int64_t end = fa->len;
    int64_t p = 0;
    while (p < end) {
        const int64_t ll = fa_line(m, p, end);
        if (m[p] == '>') {
            if (fa->n) fa->c[fa->n - 1].seq_end = p;
            if (fa->n == fa->cap) { fa->cap = fa->cap ? 2 * fa->cap : 64; fa->c = (fa_contig *)realloc(fa->c, sizeof(fa_contig) * (size_t)fa->cap); }
            fa_contig *c = &fa->c[fa->n++];
            /* name: characters 1 .. (first index > 0 that is not isgraph) - 1, lower-cased, capped (src/GROM.c:1353-1381) */
            int64_t cut = ll;
            for (int64_t k = 1; k < ll; k++) if (!isgraph((unsigned char)m[p + k])) { cut = k; break; }
            if (cut >= FA_NAME_MAX + 1) cut = FA_NAME_MAX + 1;
            int w = 0;
            for (int64_t k = 1; k < cut; k++) c->name[w++] = (char)tolower((unsigned char)m[p + k]);
            c->name[w] = 0;
            c->seq_off = p + ll; c->seq_end = end;
        }
        p += ll;
    }
    *out = fa;
    return 0;
}

void gromhost_fasta_close(grom_fasta *fa)
{
    if (!fa) return;
    if (fa->map) { if (fa->is_mmap) munmap((void *)fa->map, (size_t)fa->len); else free((void *)fa->map); }
    free(fa->c); free(fa->path); free(fa);
}

int gromhost_fasta_n(const grom_fasta *fa) { return fa->n; }
const char *gromhost_fasta_name(const grom_fasta *fa, int k) { return (k >= 0 && k < fa->n) ? fa->c[k].name : NULL; }
int64_t gromhost_fasta_raw_bytes(const grom_fasta *fa, int k) { return (k >= 0 && k < fa->n) ? fa->c[k].seq_end - fa->c[k].seq_off : -1; }

int gromhost_fasta_find(const grom_fasta *fa, const char *name)
{
    char low[FA_NAME_MAX + 1]; size_t n = strlen(name);
    if (n > FA_NAME_MAX) return -1;
    for (size_t i = 0; i <= n; i++) low[i] = (char)tolower((unsigned char)name[i]);
    for (int k = 0; k < fa->n; k++) if (fa->c[k].name[0] && !strcmp(fa->c[k].name, low)) return k;
    return -1;
}

int64_t gromhost_fasta_load(const grom_fasta *fa, int k, char *dst, int64_t cap)
{
    if (k < 0 || k >= fa->n) { gromhost_fail("FASTA contig %d out of range", k); return -1; }
    const char *m = fa->map; const int64_t end = fa->c[k].seq_end;
    int64_t p = fa->c[k].seq_off, w = 0, line_len = -1, keep = 0;
    while (p < end) {
        const int64_t ll = fa_line(m, p, end);
        if (m[p] == '>') break;
        if (w == 0 || ll != line_len) {
            /* the cut is re-evaluated only when the line length changes (src/GROM.c:21016-21024) */
            line_len = ll;
            int64_t q = ll - 1;
            while (q > 0 && !isalpha((unsigned char)m[p + q])) q--;
            keep = q + 1;
        }
        if (w + keep > cap) { gromhost_fail("%s: contig %s is longer than the %lld characters provided for", fa->path, fa->c[k].name, (long long)cap); return -1; }
        memcpy(dst + w, m + p, (size_t)keep);
        w += keep;
        p += ll;
    }
    return w;
}
