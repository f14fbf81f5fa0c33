/* oracle/shim/sam.h -- TEST INFRASTRUCTURE ONLY; see bam.h in this directory. */
#ifndef GROM_ORACLE_SHIM_SAM_H
#define GROM_ORACLE_SHIM_SAM_H
#include "bam.h"
typedef struct {
    bamFile fp;
    bam_header_t *header;
} samfile_t;
samfile_t *samopen(const char *fn, const char *mode, const void *aux);
int samread(samfile_t *fp, bam1_t *b);
void samclose(samfile_t *fp);
#endif
