/* grom_reads.h -- the packed read-record batch that crosses the host/GPU seam.
 *
 * The reference consumes one `bam1_t` at a time through `my_samread`
 * (reference src/GROM.c:981-992) and copies 8 core fields plus the first
 * SA/XP aux entry out of it (src/GROM.c:5743-5824, 10971-11068, 14864-14955).
 * This header is the structure-of-arrays form of exactly those fields for all
 * reads of ONE contig, in BAM order.  It is what the host batcher
 * (include/gromhost.h) produces and what the CUDA library (include/gromgpu.h)
 * consumes; plain pointers and sizes only.
 *
 * Layout rules
 *  - every per-read array has n_reads entries, index = BAM order on the contig;
 *  - cigar[] holds BAM-encoded ops (len<<4|op); read i owns
 *    cigar[cigar_off[i] .. cigar_off[i]+n_cigar[i]);
 *  - bases: read i owns base slots base_off[i] .. base_off[i]+l_qseq[i);
 *    base_off[i] is a multiple of GROM_BASE_ALIGN = 32 (a read's quals start 32-byte
 *    aligned and its nibbles 16-byte aligned, which is what the bulk-copy (TMA)
 *    staging of the pileup kernel requires); qual[slot] is the phred byte, and the
 *    4-bit BAM base code ("=ACMGRSVTWYHKDBN") of a slot is
 *    (seq4[slot>>1] >> ((~slot&1)<<2)) & 15, i.e. the BAM nibble order.
 *    4-bit codes rather than 2-bit are kept on purpose: the reference compares
 *    the decoded character (including IUPAC codes and 'N') with toupper(ref)
 *    (src/GROM.c:6806) and a 2-bit+mask form could not reproduce that bit-exactly;
 *  - qname_hash is a 64-bit FNV-1a hash of the read name, qname_len its strlen
 *    (capped at 255).  The reference stores/compares names only as strings of
 *    length < 50 (src/GROM.c:6810-6821);
 *  - sa_*: first entry of the XP tag, else of the SA tag, parsed only when
 *    0 < l_aux < 100 (src/GROM.c:5763); sa_pos = -1 when absent.  sa_same_chr is
 *    the reference's prefix test strncmp(target_name, sa_chr, strlen(target_name))==0
 *    (src/GROM.c:7431).  sa_start_adj/sa_end_adj/sa_end_adj_indel follow
 *    src/GROM.c:6686-6733 (leading/trailing 'S' length, sum(I)-sum(D)).
 */
#ifndef GROM_READS_H
#define GROM_READS_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GROM_BASE_ALIGN 32

typedef struct grom_read_batch {
    int64_t n_reads;
    int64_t n_cigar_total;      /* entries in cigar[] */
    int64_t n_base_slots;       /* entries in qual[]; seq4 has (n_base_slots+1)/2 bytes */
    int32_t tid;                /* BAM target id shared by every read of the batch */
    int32_t layout_flags;       /* 0 = canonical arrays only; GROM_LAYOUT_* bits announce the transport-compact forms below */
    const int32_t  *pos;        /* 0-based leftmost position (core.pos) */
    const int32_t  *mpos;       /* core.mpos */
    const int32_t  *tlen;       /* core.isize */
    const int32_t  *mtid;       /* core.mtid */
    const int32_t  *l_qseq;     /* core.l_qseq */
    const uint16_t *flag;       /* core.flag */
    const uint16_t *n_cigar;    /* core.n_cigar */
    const uint8_t  *mapq;       /* core.qual */
    const uint8_t  *qname_len;
    const uint64_t *qname_hash;
    const uint64_t *cigar_off;
    const uint64_t *base_off;
    const uint32_t *cigar;
    const uint8_t  *seq4;
    const uint8_t  *qual;
    const int32_t  *sa_pos;     /* as written in the tag (1-based), -1 = none (src/GROM.c:5808) */
    const int32_t  *sa_start_adj;
    const int32_t  *sa_end_adj;
    const int32_t  *sa_end_adj_indel;
    const uint8_t  *sa_strand;  /* 0 '+', 1 otherwise */
    const int16_t  *sa_mapq;
    const uint8_t  *sa_same_chr;
    /* host-only, optional (NULL allowed): read names for BAM serialisation */
    const uint64_t *qname_off;  /* [n_reads+1] */
    const char     *qname_pool;
    /* ---- transport-compact forms (read only when the matching layout_flags bit is set).  They carry the same information
     * in fewer bytes across PCIe; the CUDA library rebuilds the canonical device arrays from them (bit-identical), so
     * nothing downstream changes.  319 -> 163 bytes per 150 bp read with offsets + qual4 + seq2 + sparse SA, 123 with qual2. */
    const uint8_t  *qual4;      /* GROM_LAYOUT_QUAL4: per base slot a 4-bit index into qual_lut, nibble order of seq4
                                   (usable when the batch holds <= 16 distinct quality values); qual may then be NULL */
    uint8_t         qual_lut[16];
    int64_t         n_sa;       /* GROM_LAYOUT_SPARSE_SA: the first-SA-entry fields of only the n_sa reads that have one, */
    const int32_t  *sa_index;   /*   read indices ascending; every other read has sa_pos = sa_mapq = -1 and zeros elsewhere. */
    const int32_t  *sas_pos, *sas_start_adj, *sas_end_adj, *sas_end_adj_indel;   /* [n_sa]; the dense sa_* arrays may then be NULL */
    const int16_t  *sas_mapq;
    const uint8_t  *sas_strand, *sas_same_chr;
    const uint8_t  *seq2;       /* GROM_LAYOUT_SEQ2: 2 bits per base slot (A C G T = 0 1 2 3 = log2 of the BAM code), slot s in byte s>>2 at
                                   bits ((~s&3)<<1); every base that is not A/C/G/T is listed as an exception with its 4-bit BAM code.
                                   Padding slots (beyond l_qseq) become code 0 on the device like in the canonical array */
    int64_t         n_seq_exc;
    const uint64_t *seq_exc_slot;   /* [n_seq_exc] base slot (same index space as base_off) */
    const uint8_t  *seq_exc_code;   /* [n_seq_exc] BAM 4-bit code */
    const uint8_t  *qual2;      /* GROM_LAYOUT_QUAL2: like qual4 with 2-bit indices into qual_lut[0..3], slot s in byte s>>2 at bits ((~s&3)<<1)
                                   (usable when the bases of the batch -- padding slots aside -- carry <= 4 distinct qualities, e.g. the four
                                   bins of current Illumina instruments); padding slots become 0 on the device.  Preferred over qual4 */
} grom_read_batch;

/* GROM_LAYOUT_CANONICAL_OFFSETS: cigar_off[i] = sum of n_cigar[0..i) and base_off[i] = sum of l_qseq[0..i) each rounded up to
 * GROM_BASE_ALIGN (what the host batcher always produces); the two offset arrays are then derived on the device and may be NULL */
#define GROM_LAYOUT_CANONICAL_OFFSETS 1
#define GROM_LAYOUT_QUAL4             2
#define GROM_LAYOUT_SPARSE_SA         4
#define GROM_LAYOUT_SEQ2              8
#define GROM_LAYOUT_QUAL2            16

static inline uint64_t grom_qname_hash(const char *s, int len)
{
    uint64_t h = 1469598103934665603ULL;
    for (int i = 0; i < len; i++) { h ^= (uint8_t)s[i]; h *= 1099511628211ULL; }
    return h;
}

#ifdef __cplusplus
}
#endif
#endif
