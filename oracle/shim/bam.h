/* oracle/shim/bam.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Minimal declaration-level stand-in for the legacy samtools-0.1 style API that
 * the reference translation unit includes as "bam.h"/"sam.h" (reference
 * src/GROM.c:26-27; Makefile:3-7 would take it from samtools-1.3.1, whose tarball
 * is absent from the mount).  Only the 14 externs and the handful of macros that
 * GROM.c actually names are provided; the implementation (shim.c) is a small
 * zlib-only BGZF/BAM/BAI reader written for this repository.
 */
#ifndef GROM_ORACLE_SHIM_BAM_H
#define GROM_ORACLE_SHIM_BAM_H
#include <stdint.h>
#include <stdio.h>

#define BAM_FPAIRED        1
#define BAM_FPROPER_PAIR   2
#define BAM_FUNMAP         4
#define BAM_FMUNMAP        8
#define BAM_FREVERSE      16
#define BAM_FMREVERSE     32
#define BAM_FREAD1        64
#define BAM_FREAD2       128
#define BAM_FSECONDARY   256
#define BAM_FQCFAIL      512
#define BAM_FDUP        1024

#define BAM_CMATCH      0
#define BAM_CINS        1
#define BAM_CDEL        2
#define BAM_CREF_SKIP   3
#define BAM_CSOFT_CLIP  4
#define BAM_CHARD_CLIP  5
#define BAM_CPAD        6
#define BAM_CEQUAL      7
#define BAM_CDIFF       8

typedef struct {
    int32_t tid;
    int32_t pos;
    uint32_t bin:16, qual:8, l_qname:8;
    uint32_t flag:16, n_cigar:16;
    int32_t l_qseq;
    int32_t mtid;
    int32_t mpos;
    int32_t isize;
} bam1_core_t;

typedef struct {
    bam1_core_t core;
    int l_data, m_data;
    uint8_t *data;
    uint64_t id;
} bam1_t;
#define data_len l_data

typedef struct {
    int32_t n_targets;
    char **target_name;
    uint32_t *target_len;
    uint32_t l_text;
    char *text;
} bam_header_t;

typedef struct shim_bgzf shim_bgzf_t;
typedef shim_bgzf_t *bamFile;
typedef struct shim_bai bam_index_t;
typedef int (*bam_fetch_f)(const bam1_t *b, void *data);

#define bam_cigar_op(c)     ((c) & 0xf)
#define bam_cigar_oplen(c)  ((c) >> 4)
#define bam1_qname(b)  ((char *)(b)->data)
#define bam1_cigar(b)  ((uint32_t *)((b)->data + (b)->core.l_qname))
#define bam1_seq(b)    ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname)
#define bam1_qual(b)   ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1))
#define bam1_aux(b)    ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (b)->core.l_qseq + (((b)->core.l_qseq + 1) >> 1))
#define bam1_seqi(s, i) ((s)[(i) >> 1] >> ((~(i) & 1) << 2) & 0xf)
#define bam_get_l_aux(b) ((b)->l_data - ((b)->core.n_cigar << 2) - (b)->core.l_qname - (b)->core.l_qseq - (((b)->core.l_qseq + 1) >> 1))

extern const char bam_nt16_rev_table[];

bamFile bam_open(const char *fn, const char *mode);
int bam_close(bamFile fp);
bam_header_t *bam_header_read(bamFile fp);
void bam_header_destroy(bam_header_t *h);
bam_index_t *bam_index_load(const char *fn);
void bam_index_destroy(bam_index_t *idx);
int bam_fetch(bamFile fp, const bam_index_t *idx, int tid, int beg, int end, void *data, bam_fetch_f func);
bam1_t *bam_init1(void);
void bam_destroy1(bam1_t *b);
uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]);
int bam_read1(bamFile fp, bam1_t *b);
#endif
