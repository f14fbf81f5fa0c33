/* grom_b200.c -- the host program of the B200 build: GROM's command line and output files over libgromhost (BAM batcher, tables,
 * candidate post-processing, record text) and libgromgpu (the hot path on the GPU, include/gromgpu.h).
 *
 * It replaces, with the same observable contract, the reference's main / find_disc_svs / -P scheduler:
 *   option letters and their defaults                         reference src/GROM.c:21907-22106
 *   <out> and <out>.ctx(.vcf) header blocks                   src/GROM.c:20517-20565, 22639-22677   (output format contract)
 *   ctx file naming                                           src/GROM.c:22431-22445
 *   library statistics before anything else                   src/GROM.c:22230-22262 (find_insert_mean)
 *   per-contig loop, contigs missing from either file skipped src/GROM.c:20900-21130; chrY skipped unless -g 1 (20979-20988)
 *   -P N: N contigs in parallel, largest first                src/GROM.c:22318-22336, 549-599; outputs merged in BAM header order (21121-21126)
 *
 * Here -P N means N GPUs (one worker PROCESS per GPU, started with exec like the reference's children; no CUDA state is shared) and
 * each worker keeps GROM_LANES contigs in flight on its GPU (default 3: decode / upload of one contig overlaps the kernels and the host
 * stages of the others).  Workers write per-contig part files; the parent concatenates them in BAM header order and pairs the
 * translocation records of all contigs, exactly what the reference's parent does with its children's files.
 * Under torch.distributed (bench.py --workload wgs) every rank runs one worker (--rank R --world W --device D --parts-only) and rank 0
 * merges (--merge-only).
 *
 * No CPU fallback: without a CUDA device gromgpu_init fails and the program exits 1 with the library's message.
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <errno.h>
#include <getopt.h>
#include <libgen.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>
#include "gromgpu.h"
#include "gromhost.h"

static double now_s(void) { struct timeval tv; gettimeofday(&tv, NULL); return tv.tv_sec + 1e-6 * tv.tv_usec; }
static void die(const char *fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
#include <stdarg.h>
static void die(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); exit(1); }

/* ------------------------------------------------------------------------------------------------ output format contract */
/* the reference's header blocks (src/GROM.c:20517-20565, 22639-22677) come from libgromhost (gromhost_vcf_header) */
static void write_header(FILE *f, const char *fasta_name, int is_ctx)
{
    char buf[16384];
    const int64_t n = gromhost_vcf_header(fasta_name, is_ctx, buf, (int64_t)sizeof(buf));
    if (n < 0) die("GROM_b200: header does not fit");
    fwrite(buf, 1, (size_t)n, f);
}

/* <out>.vcf -> <out>.ctx.vcf, anything else -> <out>.ctx */
static void ctx_name(const char *out, char *dst, size_t cap)
{
    const size_t n = strlen(out);
    if (n > 4 && !strcmp(out + n - 4, ".vcf")) snprintf(dst, cap, "%.*s.ctx.vcf", (int)(n - 4), out);
    else snprintf(dst, cap, "%s.ctx", out);
}

/* ------------------------------------------------------------------------------------------------ FASTA */
/* index and per-contig load are libgromhost's (gromhost_fasta_*: the reference's line rules, src/GROM.c:1332-1417 and 21011-21045) */
typedef grom_fasta fa_index;

static int fasta_find(const fa_index *ix, const char *lname) { return gromhost_fasta_find(ix, lname); }
/* characters of one contig, line ends removed, case preserved (the hot path compares through toupper and tests 'N'/'n' literally) */
static char *fasta_load(const fa_index *ix, int k, int64_t *len)
{
    const int64_t raw = gromhost_fasta_raw_bytes(ix, k);
    if (raw < 0) die("FASTA contig %d out of range", k);
    char *s = (char *)malloc((size_t)raw + 1);
    if (!s) die("out of memory (%lld characters of reference)", (long long)raw);
    const int64_t n = gromhost_fasta_load(ix, k, s, raw);
    if (n < 0) die("%s", gromhost_last_error());
    *len = n;
    return s;
}

/* ------------------------------------------------------------------------------------------------ options */
typedef struct {
    const char *bam, *fasta, *out;
    grom_params prm;
    int P;                       /* -P: GPUs (worker processes) */
    int rank, world, device, lanes, threads, parts_only, merge_only, have_stats;
    long slice_reads;            /* --slice-reads: decode and push a contig in pieces of about this many records (0 = whole contig) */
    int st_mean, st_lseq, st_min, st_max;
    const char *stats_json;
} options;

static void usage(void)
{
    printf("\nGROM_b200 (B200 build of GROM's hot path)\n"
           "Usage: GROM_b200 -i <BAM input file> -r <REFERENCE input file> -o <output file> [optional parameters]\n\n"
           "\t-i BAM (coordinate sorted; <bam>.bai is used when present)   -r FASTA   -o output (.vcf -> also <out>.ctx.vcf)\n"
           "\t-M remove duplicates   -P <GPUs>   -q <min MAPQ 20>   -b <min base quality 20>   -v <SV/SNV p-value 0.001>\n"
           "\t-e <insertion p-value 1e-10>   -V <read-depth p-value 1e-9>   -p <ploidy 2>   -g <1 = process chrY>   -A <window sampling 2>\n"
           "\t-S no split reads  -n -d -a -y -z -j -m -u -x -l -W -X: as in GROM\n"
           "\t--lanes <contigs in flight per GPU, 3>   --threads <decode threads per lane>   --stats <json file>\n"
           "\t--slice-reads <n>: decode and upload a contig in pieces of about n records (bounds host memory; needs an index with record counts)\n");
}

static int parse(int argc, char **argv, options *o)
{
    memset(o, 0, sizeof(*o));
    grom_params_default(&o->prm);
    o->rank = 0; o->world = 0; o->device = -1; o->lanes = getenv("GROM_LANES") ? atoi(getenv("GROM_LANES")) : 3;
    static const struct option lo[] = { {"rank", 1, 0, 1000}, {"world", 1, 0, 1001}, {"device", 1, 0, 1002}, {"lanes", 1, 0, 1003}, {"parts-only", 0, 0, 1004},
                                        {"merge-only", 0, 0, 1005}, {"libstats", 1, 0, 1006}, {"threads", 1, 0, 1007}, {"stats", 1, 0, 1008}, {"slice-reads", 1, 0, 1009}, {0, 0, 0, 0} };
    int c;
    /* the reference's option string; letters this build has no counterpart for are accepted only with the reference's default value */
    while ((c = getopt_long(argc, argv, "Z:W:X:Q:A:Y:B:D:E:K:N:V:U:L:F:SP:c:R:MG:i:r:o:p:q:s:v:g:l:d:b:n:a:y:z:e:fj:k:m:u:w:x:h", lo, NULL)) != -1) {
        switch (c) {
        case 'i': o->bam = optarg; break;
        case 'r': o->fasta = optarg; break;
        case 'o': o->out = optarg; break;
        case 'S': o->prm.splitread = 0; break;
        case 'M': o->prm.rmdup = 1; break;
        case 'P': o->P = atoi(optarg); if (o->P < 0 || o->P > 256) o->P = 0; break;
        case 'W': o->prm.min_rd_window_len = atoi(optarg); break;
        case 'X': o->prm.max_rd_window_len = atoi(optarg); break;
        case 'A': o->prm.windows_sampling_factor = atoi(optarg); break;
        case 'V': o->prm.rd_pval_threshold = atof(optarg); break;
        case 'p': o->prm.ploidy = atoi(optarg); break;
        case 'q': o->prm.min_mapq = atoi(optarg); break;
        case 'v': o->prm.pval_threshold = atof(optarg); break;
        case 'g': o->prm.gender = atoi(optarg); break;
        case 'l': o->prm.overlap_mult = atoi(optarg); break;
        case 'd': o->prm.min_disc = atoi(optarg); break;
        case 'b': o->prm.min_base_qual = atoi(optarg); break;
        case 'n': o->prm.min_snv = atoi(optarg); break;
        case 'a': o->prm.min_snv_ratio = atof(optarg); break;
        case 'y': o->prm.max_split_loss = atoi(optarg); break;
        case 'z': o->prm.min_sr_len = atoi(optarg); break;
        case 'e': o->prm.pval_insertion = atof(optarg); break;
        case 'j': o->prm.min_sv_ratio = atof(optarg); break;
        case 'm': o->prm.min_indel_ratio = atof(optarg); break;
        case 'u': o->prm.max_evidence_ratio = atof(optarg); break;
        case 'x': o->prm.min_ave_bq = atof(optarg); break;
        case 'Q': break;                                   /* overwritten by -q in the reference too (src/GROM.c:22102) */
        case 'h': usage(); exit(0);
        case 'f': die("GROM_b200: -f (tab-separated debug output) is not part of this build");
        case 'c': case 'R': die("GROM_b200: -%c (the reference's child / sub-region mode) is replaced by --rank / --world", c);
        case 'Z': case 'Y': case 'B': case 'D': case 'E': case 'K': case 'N': case 'U': case 'L': case 'F': case 'G': case 's': case 'k': case 'w':
            fprintf(stderr, "GROM_b200: option -%c %s is accepted for command-line compatibility; this build keeps the reference's default for it\n", c, optarg);
            break;
        case 1000: o->rank = atoi(optarg); break;
        case 1001: o->world = atoi(optarg); break;
        case 1002: o->device = atoi(optarg); break;
        case 1003: o->lanes = atoi(optarg); break;
        case 1004: o->parts_only = 1; break;
        case 1005: o->merge_only = 1; break;
        case 1006: if (sscanf(optarg, "%d,%d,%d,%d", &o->st_mean, &o->st_lseq, &o->st_min, &o->st_max) != 4) die("--libstats wants mean,lseq,min,max"); o->have_stats = 1; break;
        case 1007: o->threads = atoi(optarg); break;
        case 1008: o->stats_json = optarg; break;
        case 1009: o->slice_reads = atol(optarg); break;
        default: return 1;
        }
    }
    o->prm.pval_threshold1 = o->prm.pval_threshold;          /* src/GROM.c:22101 */
    o->prm.rd_min_mapq = o->prm.min_mapq;                    /* src/GROM.c:22102 */
    if (o->lanes < 1) o->lanes = 1;
    if (!o->bam) die("ERROR: No bam file specified.");
    if (!o->out) die("ERROR: No output file specified.");
    if (!o->fasta) die("ERROR: No reference file specified.");
    return 0;
}

/* ------------------------------------------------------------------------------------------------ work list */
typedef struct { int tid, fa; int64_t len; } contig;

static char *lower_dup(const char *s) { char *d = strdup(s); for (char *p = d; *p; p++) *p = (char)tolower((unsigned char)*p); return d; }

/* contigs present in both files, chrY / y only with -g 1; largest-first greedy assignment to `world` workers (the -P policy with slot = GPU) */
static int plan(const grom_bam *bam, const fa_index *fa, const grom_params *prm, int world, int rank, contig **mine, int *n_all, contig **all_out)
{
    const int nt = gromhost_bam_n_targets(bam);
    contig *all = (contig *)malloc(sizeof(contig) * (size_t)(nt > 0 ? nt : 1));
    int n = 0;
    for (int t = 0; t < nt; t++) {
        char *ln = lower_dup(gromhost_bam_target_name(bam, t));
        const int k = fasta_find(fa, ln);
        const int is_y = !strcmp(ln, "chry") || !strcmp(ln, "y");
        free(ln);
        if (k < 0 || (is_y && prm->gender == 0)) continue;
        all[n].tid = t; all[n].fa = k; all[n].len = gromhost_bam_target_len(bam, t); n++;
    }
    /* load of a contig: its records where the index counts them (coverage differs between contigs), else its length */
    double *wgt = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    int counted = 1;
    for (int i = 0; i < n; i++) { int64_t mp = 0, um = 0; if (gromhost_bam_target_reads(bam, all[i].tid, &mp, &um)) { counted = 0; break; } wgt[i] = (double)(mp + um); }
    /* order by length descending, ties by tid (the weights move along) */
    for (int i = 1; i < n; i++) {
        contig x = all[i]; const double wx = wgt[i]; int j = i - 1;
        while (j >= 0 && (all[j].len < x.len || (all[j].len == x.len && all[j].tid > x.tid))) { all[j + 1] = all[j]; wgt[j + 1] = wgt[j]; j--; }
        all[j + 1] = x; wgt[j + 1] = wx;
    }
    double *load = (double *)calloc((size_t)(world > 0 ? world : 1), sizeof(double));
    contig *m = (contig *)malloc(sizeof(contig) * (size_t)(n > 0 ? n : 1));
    int nm = 0;
    for (int i = 0; i < n; i++) {
        int r = 0;
        for (int k = 1; k < world; k++) if (load[k] < load[r]) r = k;
        load[r] += counted ? wgt[i] : (double)all[i].len;
        if (r == rank) m[nm++] = all[i];
    }
    free(load); free(wgt);
    *mine = m; *n_all = n; *all_out = all;
    return nm;
}

static void part_name(const char *out, int tid, const char *kind, char *dst, size_t cap) { snprintf(dst, cap, "%s.%s.%d", out, kind, tid); }

/* ------------------------------------------------------------------------------------------------ worker */
typedef struct {
    const options *o; grom_bam *bam_main; const fa_index *fa; contig *work; int n_work, next; pthread_mutex_t pick, bus, mem;
    pthread_cond_t mem_cv; int64_t mem_budget, mem_used; int running;
    const double *p2s_p, *p2s_sd; int n_p2s;
    double t_decode, t_upload, t_gpu_run, t_cnv, t_text; int64_t reads, bases, records; float ms_dev_run, ms_dev_cnv;
    int failed; char err[1024];
} worker;

static void worker_fail(worker *w, const char *what, const char *msg)
{
    pthread_mutex_lock(&w->pick);
    if (!w->failed) { w->failed = 1; snprintf(w->err, sizeof(w->err), "%s: %s", what, msg); }
    pthread_mutex_unlock(&w->pick);
}

static void *lane_main(void *arg)
{
    worker *w = (worker *)arg;
    const options *o = w->o;
    grom_bam *bam = NULL;
    if (gromhost_bam_open(o->bam, &bam)) { worker_fail(w, "gromhost_bam_open", gromhost_last_error()); return NULL; }
    void *stream = NULL;
    if (gromgpu_stream_create(&stream)) { worker_fail(w, "gromgpu_stream_create", gromgpu_last_error()); gromhost_bam_close(bam); return NULL; }
    size_t cap = 1 << 20; char *text = (char *)malloc(cap);
    /* one handle per lane: begun for the lane's first contig (the largest it will see: the queue is sorted largest first) and
     * rebound -- same device buffers, new contig -- for the rest; its memory reservation is kept for as long as the handle lives */
    gromgpu_chr *h = NULL;
    int64_t held = 0;
    for (;;) {
        pthread_mutex_lock(&w->pick);
        const int k = (w->failed || w->next >= w->n_work) ? -1 : w->next++;
        pthread_mutex_unlock(&w->pick);
        if (k < 0) break;
        const contig c = w->work[k];
        char *lname = lower_dup(gromhost_bam_target_name(bam, c.tid));
        int64_t flen = 0;
        char *chars = fasta_load(w->fa, c.fa, &flen);
        if (flen != c.len) fprintf(stderr, "GROM_b200: warning: %s is %lld bases in the FASTA and %lld in the BAM header\n", lname, (long long)flen, (long long)c.len);
        double t0 = now_s();
        grom_batch *bt = NULL;
        grom_read_batch v; memset(&v, 0, sizeof(v));
        /* --slice-reads: the contig's totals come from the index and its reads are decoded piece by piece while they are pushed (below);
         * otherwise the whole contig is decoded here */
        int64_t cnt_m = 0, cnt_u = 0;
        const int sliced = o->slice_reads > 0 && gromhost_bam_target_reads(bam, c.tid, &cnt_m, &cnt_u) == 0;
        if (sliced) { v.n_reads = cnt_m + cnt_u; v.n_base_slots = v.n_reads * (int64_t)((o->prm.lseq + 31) / 32 * 32); }
        else {
            if (gromhost_bam_read_target(bam, c.tid, 0, o->threads, &bt)) { worker_fail(w, "gromhost_bam_read_target", gromhost_last_error()); free(chars); free(lname); break; }
            gromhost_batch_view(bt, &v);
        }
        double t1 = now_s();
        /* admission: the handle's device memory must fit beside the contigs already in flight; a contig that fits nowhere runs alone */
        const int64_t need = gromgpu_chr_bytes_estimate(flen, v.n_reads, v.n_base_slots);
        gromgpu_result res; gromgpu_cnv_result cnv; gromgpu_stats st;
        int bad = 0;
        double t2 = t1, t3 = t1, t4 = t1, t5 = t1;
        memset(&st, 0, sizeof(st)); memset(&cnv, 0, sizeof(cnv)); memset(&res, 0, sizeof(res));
        int rb = h ? gromgpu_chr_rebind(h, c.tid, chars, flen) : 1;
        if (rb < 0) bad = 1;
        else if (rb == 1) {
            if (h) {                                          /* (a larger contig after a smaller one: only if the queue was not sorted) */
                gromgpu_chr_free(h); h = NULL;
                pthread_mutex_lock(&w->mem); w->mem_used -= held; w->running--; held = 0; pthread_cond_broadcast(&w->mem_cv); pthread_mutex_unlock(&w->mem);
            }
            pthread_mutex_lock(&w->mem);
            while (w->running && w->mem_used + need > w->mem_budget) pthread_cond_wait(&w->mem_cv, &w->mem);
            w->mem_used += need; w->running++; held = need;
            pthread_mutex_unlock(&w->mem);
            if (gromgpu_chr_begin_on(&h, c.tid, chars, flen, stream)) bad = 1;
        }
        int64_t n_reads = v.n_reads;
        if (!bad && !sliced) {
            pthread_mutex_lock(&w->bus);                      /* one upload at a time: the PCIe link is the shared resource */
            bad = gromgpu_push_reads(h, &v) || gromgpu_chr_sync(h);
            pthread_mutex_unlock(&w->bus);
            t2 = now_s();
        }
        if (!bad && sliced) {
            grom_target_iter *it = NULL;
            if (gromhost_bam_iter_open(bam, c.tid, 0, o->threads, &it)) { worker_fail(w, "gromhost_bam_iter_open", gromhost_last_error()); bad = 2; }
            n_reads = 0;
            while (!bad) {
                grom_batch *piece = NULL;
                const int rc = gromhost_bam_iter_next(it, o->slice_reads, &piece);
                if (rc == 1) break;
                if (rc) { worker_fail(w, "gromhost_bam_iter_next", gromhost_last_error()); bad = 2; break; }
                grom_read_batch pv; gromhost_batch_view(piece, &pv);
                if (pv.n_reads) {
                    pthread_mutex_lock(&w->bus);
                    bad = gromgpu_push_reads(h, &pv) || gromgpu_chr_sync(h);
                    pthread_mutex_unlock(&w->bus);
                    n_reads += pv.n_reads;
                }
                gromhost_batch_free(piece);
            }
            gromhost_bam_iter_close(it);
            t2 = now_s();
        }
        gromhost_batch_free(bt); bt = NULL;                   /* the reads live on the device now */
        if (!bad) { bad = gromgpu_chr_finish(h, &res); t3 = now_s(); }
        if (!bad) { bad = gromgpu_chr_cnv(h, w->p2s_p, w->p2s_sd, w->n_p2s, o->prm.ploidy, &cnv); t4 = now_s(); }
        if (!bad) gromgpu_chr_stats(h, &st);
        if (bad == 1) worker_fail(w, "gromgpu", gromgpu_last_error());
        int64_t nrec = 0;
        if (!bad) {
            int64_t n;
            {   /* room for every candidate as a record up front: a refused call would format the contig again */
                const size_t want = (size_t)res.n_snv * 224 + (size_t)(res.n_ins + res.n_del + res.n_sv + cnv.n_calls) * 320 + (1 << 16);
                if (want > cap) { cap = want; text = (char *)realloc(text, cap); }
            }
            for (;;) {
                n = gromhost_vcf_contig(&o->prm, lname, chars, flen, res.snv, res.n_snv, res.snv_ave_rd, res.ins, res.n_ins, res.del_ev, res.n_del,
                                        res.sv_ev, res.n_sv, cnv.calls, cnv.n_calls, text, (int64_t)cap);
                if (n != -1) break;
                cap *= 4; text = (char *)realloc(text, cap);
            }
            if (n < 0) { worker_fail(w, "gromhost_vcf_contig", gromhost_last_error()); bad = 1; }
            else {
                char pn[4200]; part_name(o->out, c.tid, "part", pn, sizeof(pn));
                FILE *f = fopen(pn, "wb");
                if (!f || (int64_t)fwrite(text, 1, (size_t)n, f) != n) { worker_fail(w, pn, strerror(errno)); bad = 1; }
                if (f) fclose(f);
                for (int64_t i = 0; i < n; i++) nrec += text[i] == '\n';
            }
        }
        if (!bad) {
            /* translocation records of this contig (candidate merge + filter); the mate pairing needs all contigs and runs in the merge step */
            gromhost_sv_lists_t L;
            if (gromhost_sv_lists(&o->prm, res.sv_ev, res.n_sv, &L)) { worker_fail(w, "gromhost_sv_lists", gromhost_last_error()); bad = 1; }
            else {
                const int64_t capr = L.n_ctx_f + L.n_ctx_r + 1;
                grom_ctx_record *rec = (grom_ctx_record *)calloc((size_t)capr, sizeof(grom_ctx_record));
                const int64_t nr = gromhost_ctx_contig(&o->prm, c.tid, L.ctx_f, L.n_ctx_f, L.ctx_r, L.n_ctx_r, rec, capr);
                char pn[4200]; part_name(o->out, c.tid, "ctxpart", pn, sizeof(pn));
                FILE *f = fopen(pn, "wb");
                if (nr < 0 || !f || (nr && (int64_t)fwrite(rec, sizeof(grom_ctx_record), (size_t)nr, f) != nr)) { worker_fail(w, pn, "cannot write translocation records"); bad = 1; }
                if (f) fclose(f);
                free(rec); gromhost_sv_lists_free(&L);
            }
            t5 = now_s();
        }
        pthread_mutex_lock(&w->pick);
        w->t_decode += t1 - t0; w->t_upload += t2 - t1; w->t_gpu_run += t3 - t2; w->t_cnv += t4 - t3; w->t_text += t5 - t4;
        w->reads += n_reads; w->bases += st.aligned_bases; w->records += nrec; w->ms_dev_run += st.ms_total; w->ms_dev_cnv += cnv.ms_device;
        pthread_mutex_unlock(&w->pick);
        free(chars); free(lname);
        if (bad) break;
    }
    if (h) gromgpu_chr_free(h);
    if (held) { pthread_mutex_lock(&w->mem); w->mem_used -= held; w->running--; pthread_cond_broadcast(&w->mem_cv); pthread_mutex_unlock(&w->mem); }
    free(text);
    gromgpu_stream_destroy(stream);
    gromhost_bam_close(bam);
    return NULL;
}

static void exe_dir(const char *argv0, char *dst, size_t cap)
{
    char buf[4096];
    ssize_t n = readlink("/proc/self/exe", buf, sizeof(buf) - 1);
    if (n > 0) buf[n] = 0; else snprintf(buf, sizeof(buf), "%s", argv0);
    snprintf(dst, cap, "%s", dirname(buf));
}

/* find_insert_mean over the records in file order until its sample is full (src/GROM.c:1205-1318) */
static void library_stats(const options *o, grom_bam *bam, int *mean, int *lseq, int *imin, int *imax)
{
    int64_t mapped = 0;
    if (gromhost_bam_library_stats(bam, o->prm.min_mapq, 0, mean, lseq, imin, imax, &mapped)) die("GROM_b200: %s", gromhost_last_error());
}

static int run_worker(options *o, const char *argv0)
{
    grom_bam *bam = NULL;
    if (gromhost_bam_open(o->bam, &bam)) die("\n%s", gromhost_last_error());
    fa_index *fa = NULL;
    if (gromhost_fasta_open(o->fasta, &fa)) die("\n%s", gromhost_last_error());
    if (!o->have_stats) { library_stats(o, bam, &o->st_mean, &o->st_lseq, &o->st_min, &o->st_max); o->have_stats = 1; }
    o->prm.insert_mean = o->st_mean > o->st_lseq ? o->st_mean : o->st_lseq;          /* src/GROM.c:22260 */
    o->prm.lseq = o->st_lseq; o->prm.insert_min = o->st_min; o->prm.insert_max = o->st_max;
    if (o->rank == 0) {
        printf("insert mean, insert minimum, insert maximum: %d %d %d\n", o->prm.insert_mean, o->prm.insert_min, o->prm.insert_max);
        printf("median read length: %d\n", o->prm.lseq);
    }
    double *hez = (double *)malloc(sizeof(double) * GROM_TABLE_DIM * GROM_TABLE_DIM), *mq = (double *)malloc(sizeof(double) * GROM_TABLE_DIM * GROM_TABLE_DIM);
    char dir[4096]; exe_dir(argv0, dir, sizeof(dir));
    if (gromhost_tables_get(dir, o->prm.min_mapq, 0, hez, mq)) die("%s", gromhost_last_error());
    double p2s_p[1001], p2s_sd[1001];
    const int n_p2s = gromhost_pval2sd(p2s_p, p2s_sd, 1001);
    const int dev = o->device >= 0 ? o->device : 0;
    if (gromgpu_init(dev, hez, mq, &o->prm)) die("GROM_b200: %s", gromgpu_last_error());
    free(hez); free(mq);

    worker w; memset(&w, 0, sizeof(w));
    w.o = o; w.fa = fa; w.p2s_p = p2s_p; w.p2s_sd = p2s_sd; w.n_p2s = n_p2s;
    contig *all = NULL; int n_all = 0;
    const int world = o->world > 0 ? o->world : 1;
    w.n_work = plan(bam, fa, &o->prm, world, o->rank, &w.work, &n_all, &all);
    pthread_mutex_init(&w.pick, NULL); pthread_mutex_init(&w.bus, NULL); pthread_mutex_init(&w.mem, NULL); pthread_cond_init(&w.mem_cv, NULL);
    w.mem_budget = (int64_t)(0.9 * (double)gromgpu_device_free_bytes());
    int lanes = o->lanes < w.n_work ? o->lanes : w.n_work;
    if (lanes < 1) lanes = 1;
    if (o->threads <= 0) { long nc = sysconf(_SC_NPROCESSORS_ONLN); int per = (int)(nc / ((long)world * lanes)); ((options *)o)->threads = per > 1 ? per : 1; }
    { char tv[32]; snprintf(tv, sizeof(tv), "%d", o->threads > 2 ? o->threads : 2); setenv("GROMGPU_HOST_THREADS", tv, 0); }   /* host stages of gromgpu_chr_cnv: same share of the cores */
    const double t0 = now_s();
    pthread_t th[64];
    if (lanes > 64) lanes = 64;
    for (int i = 0; i < lanes; i++) pthread_create(&th[i], NULL, lane_main, &w);
    for (int i = 0; i < lanes; i++) pthread_join(th[i], NULL);
    const double wall = now_s() - t0;
    if (w.failed) die("GROM_b200: %s", w.err);
    if (o->stats_json) {
        FILE *f = fopen(o->stats_json, "w");
        if (f) {
            fprintf(f, "{\"rank\": %d, \"world\": %d, \"device\": %d, \"lanes\": %d, \"decode_threads_per_lane\": %d, \"contigs\": %d, \"reads\": %lld, \"aligned_bases\": %lld, \"records\": %lld, "
                       "\"wall_s\": %.4f, \"decode_s\": %.4f, \"upload_s\": %.4f, \"run_s\": %.4f, \"cnv_s\": %.4f, \"text_s\": %.4f, \"device_ms_run\": %.3f, \"device_ms_cnv\": %.3f, "
                       "\"insert_mean\": %d, \"lseq\": %d, \"insert_min\": %d, \"insert_max\": %d}\n",
                    o->rank, world, dev, lanes, o->threads, w.n_work, (long long)w.reads, (long long)w.bases, (long long)w.records, wall, w.t_decode, w.t_upload, w.t_gpu_run, w.t_cnv, w.t_text,
                    w.ms_dev_run, w.ms_dev_cnv, o->prm.insert_mean, o->prm.lseq, o->prm.insert_min, o->prm.insert_max);
            fclose(f);
        }
    }
    gromgpu_shutdown();
    gromhost_bam_close(bam);
    free(all); free(w.work);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ merge (parent / rank 0) */
static int run_merge(const options *o)
{
    grom_bam *bam = NULL;
    if (gromhost_bam_open(o->bam, &bam)) die("\n%s", gromhost_last_error());
    const int nt = gromhost_bam_n_targets(bam);
    FILE *out = fopen(o->out, "w");
    if (!out) die("\nCould not open %s", o->out);
    write_header(out, o->fasta, 0);
    grom_ctx_record *rec = NULL; int64_t n_rec = 0, cap_rec = 0;
    char *buf = (char *)malloc(1 << 20);
    for (int t = 0; t < nt; t++) {                            /* BAM header order */
        char pn[4200]; part_name(o->out, t, "part", pn, sizeof(pn));
        FILE *f = fopen(pn, "rb");
        if (f) { size_t g; while ((g = fread(buf, 1, 1 << 20, f)) > 0) fwrite(buf, 1, g, out); fclose(f); remove(pn); }
        part_name(o->out, t, "ctxpart", pn, sizeof(pn));
        f = fopen(pn, "rb");
        if (f) {
            struct stat sb; fstat(fileno(f), &sb);
            const int64_t k = (int64_t)sb.st_size / (int64_t)sizeof(grom_ctx_record);
            if (n_rec + k > cap_rec) { cap_rec = (n_rec + k) * 2 + 16; rec = (grom_ctx_record *)realloc(rec, sizeof(grom_ctx_record) * (size_t)cap_rec); }
            if (k && (int64_t)fread(rec + n_rec, sizeof(grom_ctx_record), (size_t)k, f) != k) die("%s: short read", pn);
            n_rec += k; fclose(f); remove(pn);
        }
    }
    fclose(out); free(buf);
    /* translocations: mate pairing over the records of all contigs (src/GROM.c:22470-22745) */
    char cn[4200]; ctx_name(o->out, cn, sizeof(cn));
    FILE *fc = fopen(cn, "w");
    if (!fc) die("Error opening file %s", cn);
    write_header(fc, o->fasta, 1);
    char **names = (char **)malloc(sizeof(char *) * (size_t)(nt > 0 ? nt : 1));
    for (int t = 0; t < nt; t++) names[t] = lower_dup(gromhost_bam_target_name(bam, t));
    size_t cap = 1 << 16; char *text = (char *)malloc(cap);
    grom_ctx_record *work = (grom_ctx_record *)malloc(sizeof(grom_ctx_record) * (size_t)(n_rec + 1));
    int64_t n;
    for (;;) {
        memcpy(work, rec, sizeof(grom_ctx_record) * (size_t)n_rec);       /* the pairing writes into the records */
        n = gromhost_ctx_vcf(&o->prm, (const char *const *)names, nt, work, n_rec, text, (int64_t)cap);
        if (n != -1) break;
        cap *= 4; text = (char *)realloc(text, cap);
    }
    if (n < 0) die("GROM_b200: %s", gromhost_last_error());
    fwrite(text, 1, (size_t)n, fc);
    fclose(fc);
    for (int t = 0; t < nt; t++) free(names[t]);
    free(names); free(text); free(work); free(rec);
    gromhost_bam_close(bam);
    return 0;
}

int main(int argc, char **argv)
{
    options o;
    if (parse(argc, argv, &o)) return 1;
    if (o.merge_only) {
        /* the merge needs the library statistics only for the translocation filter's insert-size terms */
        if (!o.have_stats) { grom_bam *b = NULL; if (gromhost_bam_open(o.bam, &b)) die("\n%s", gromhost_last_error()); library_stats(&o, b, &o.st_mean, &o.st_lseq, &o.st_min, &o.st_max); gromhost_bam_close(b); }
        o.prm.insert_mean = o.st_mean > o.st_lseq ? o.st_mean : o.st_lseq; o.prm.lseq = o.st_lseq; o.prm.insert_min = o.st_min; o.prm.insert_max = o.st_max;
        return run_merge(&o);
    }
    if (o.world > 0 || o.parts_only) {                       /* one rank of a multi-process job */
        if (o.world <= 0) o.world = 1;
        const int rc = run_worker(&o, argv[0]);
        if (rc || o.parts_only) return rc;
        return run_merge(&o);
    }
    printf("bam %s\nref %s\nresults %s\n", o.bam, o.fasta, o.out);
    { FILE *f = fopen(o.out, "w"); if (!f) die("\nCould not open %s", o.out); fclose(f); }
    if (o.P <= 1) {                                           /* one GPU, this process */
        o.world = 1; o.rank = 0;
        const int rc = run_worker(&o, argv[0]);
        return rc ? rc : run_merge(&o);
    }
    /* -P N: library statistics once, then N worker processes (exec: no CUDA state crosses), then the merge */
    {
        grom_bam *b = NULL;
        if (gromhost_bam_open(o.bam, &b)) die("\n%s", gromhost_last_error());
        library_stats(&o, b, &o.st_mean, &o.st_lseq, &o.st_min, &o.st_max); o.have_stats = 1;
        gromhost_bam_close(b);
    }
    char self[4096]; ssize_t sl = readlink("/proc/self/exe", self, sizeof(self) - 1);
    if (sl <= 0) die("GROM_b200: cannot locate the executable");
    self[sl] = 0;
    pid_t *pid = (pid_t *)calloc((size_t)o.P, sizeof(pid_t));
    for (int r = 0; r < o.P; r++) {
        pid[r] = fork();
        if (pid[r] < 0) die("fork: %s", strerror(errno));
        if (pid[r] == 0) {
            char **av = (char **)calloc((size_t)argc + 16, sizeof(char *));
            int n = 0;
            av[n++] = self;
            for (int i = 1; i < argc; i++) av[n++] = argv[i];
            char a[5][64];
            snprintf(a[0], 64, "%d", r); snprintf(a[1], 64, "%d", o.P); snprintf(a[2], 64, "%d", r); snprintf(a[3], 64, "%d,%d,%d,%d", o.st_mean, o.st_lseq, o.st_min, o.st_max);
            av[n++] = "--rank"; av[n++] = a[0]; av[n++] = "--world"; av[n++] = a[1]; av[n++] = "--device"; av[n++] = a[2]; av[n++] = "--libstats"; av[n++] = a[3]; av[n++] = "--parts-only";
            av[n] = NULL;
            execv(self, av);
            _exit(127);
        }
    }
    int bad = 0;
    for (int r = 0; r < o.P; r++) { int stw = 0; waitpid(pid[r], &stw, 0); if (!WIFEXITED(stw) || WEXITSTATUS(stw)) bad = 1; }
    free(pid);
    if (bad) die("GROM_b200: a worker failed");
    o.prm.insert_mean = o.st_mean > o.st_lseq ? o.st_mean : o.st_lseq; o.prm.lseq = o.st_lseq; o.prm.insert_min = o.st_min; o.prm.insert_max = o.st_max;
    return run_merge(&o);
}
