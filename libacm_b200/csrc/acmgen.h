/* acmgen.h -- synthetic ACM stream generator (see acmgen.c). */
#ifndef ACMGEN_H
#define ACMGEN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ACMGEN_DIST_FALLOUT = 0, ACMGEN_DIST_STRESS = 1, ACMGEN_DIST_SINGLE = 2 };
enum { ACMGEN_INJECT_NONE = 0, ACMGEN_INJECT_BAD_IND = 1, ACMGEN_INJECT_BAD_TCODE = 2 };

typedef struct acmgen_params {
	uint32_t level;        /* 0..15: cols = 1<<level (decode.c:747, :802) */
	uint32_t rows;         /* 1..4095 (decode.c:748-750) */
	uint32_t channels;     /* 1 or 2 (decode.c:739-741) */
	uint32_t rate;         /* >= 4096 (decode.c:743-745) */
	uint32_t total_values; /* PCM words, all channels (decode.c:734-738) */
	uint32_t wavc;         /* prepend the 28-byte WAVC header */
	uint32_t dist;         /* ACMGEN_DIST_* */
	uint32_t single_ind;   /* filler selector for ACMGEN_DIST_SINGLE */
	uint32_t pzero;        /* P(zero symbol) * 256 for the k-codes */
	uint32_t inject;       /* ACMGEN_INJECT_*: one deliberate defect */
	uint32_t inject_block, inject_col, inject_value;
	uint32_t reserved;
	uint64_t seed;
} acmgen_params;

size_t acmgen_bound(const acmgen_params *p);
size_t acmgen_write(const acmgen_params *p, uint8_t *out, size_t cap);
size_t acmgen_write_many(const acmgen_params *params, size_t n, uint8_t *blob, size_t cap,
			 uint64_t *offs, uint32_t *lens);

#ifdef __cplusplus
}
#endif
#endif
