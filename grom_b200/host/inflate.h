/* inflate.h -- the host batcher's raw DEFLATE decoder (inflate.c); internal to libgromhost */
#ifndef GROM_INFLATE_H
#define GROM_INFLATE_H
#include <stddef.h>
#include <stdint.h>

struct grom_inflate_ctx;                                  /* decode tables; one per thread, reusable across calls */
size_t grom_inflate_ctx_size(void);
void grom_inflate_ctx_init(struct grom_inflate_ctx *c);
/* 0 = well-formed stream that ended with a final block after exactly out_len bytes; -1 = anything else (nothing outside the two
 * buffers is touched either way) */
int grom_inflate_raw(struct grom_inflate_ctx *c, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len);
/* the same through the variant compiled without BMI2 (grom_inflate_raw picks the BMI2 one where the CPU has it) */
int grom_inflate_raw_generic(struct grom_inflate_ctx *c, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len);

#endif
