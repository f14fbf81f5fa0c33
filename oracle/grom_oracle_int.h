/* oracle/grom_oracle_int.h -- TEST INFRASTRUCTURE ONLY: internal types shared by the oracle's translation units. */
#ifndef GROM_ORACLE_INT_H
#define GROM_ORACLE_INT_H
#include "grom_oracle.h"

typedef struct { int64_t len; int32_t *a[GA_COUNT]; } arrs;

/* cluster classes in the order of the white-box dump (oracle/hooks.h GH_CL) */
enum { CL_DEL_F = 0, CL_DEL_R, CL_DUP_F, CL_DUP_R, CL_INV_F1, CL_INV_R1, CL_INV_F2, CL_INV_R2, CL_CTX_F, CL_CTX_R, CL_COUNT };
/* `other` slot types, reference src/GROM.c:668-681; for the ten cluster classes type = class + 1 */
enum { OTHER_EMPTY = 0, OTHER_INDEL_I = 11, OTHER_INDEL_D_F = 12, OTHER_INDEL_D_R = 13 };

typedef struct { int w; int type; int mchr; double dist; int rs, re; } oslot;

typedef struct {
    const grom_params *p;
    int64_t P;
    arrs *A;
    int32_t *cw[CL_COUNT], *crs[CL_COUNT], *cre[CL_COUNT];
    double *cdist[CL_COUNT];
    int32_t *cmchr[2];
    oslot **oth;                 /* per position, allocated on first use: other_len slots */
    char **ins_seq;              /* per position, allocated on first use: the 50-char inserted-sequence row set (src/GROM.c:7219-7228) */
    int W;
} svctx;

typedef struct {
    int tid, pos, mpos, mtid, tlen, flag, mapq, add;
    int lseq;                    /* l_qseq + hard-clip lengths (src/GROM.c:6997-7000) */
    int start_adj, end_adj, end_adj_indel;
    const uint32_t *cigar; int n_cigar;       /* capped at max_cigar_ops */
    int sa_pos, sa_strand, sa_mapq, sa_same, sa_start_adj, sa_end_adj, sa_end_adj_indel;
    int64_t win_lo;              /* reference position of window index 0 when this read is applied */
    int64_t read_index;
    const grom_read_batch *batch;
} svread;


void sv_evidence_read(svctx *c, const svread *r);
int  sv_split_dup_fwd(svctx *c, const svread *r);
void sv_split_dup_rev(svctx *c, const svread *r);
void sv_split_del(svctx *c, const svread *r);
#endif
