/*
 * acm_tables.h -- code tables for the filler decoders, built on the host once and
 * uploaded to the device (kernels stage them into shared memory).
 *
 * The reference decodes the prefix ("k") codes one GET_BITS at a time
 * (decode.c:208-403) and the radix ("t") codes with % and / (decode.c:405-476).
 * Here both become table lookups:
 *
 *   k8[kt][b]   kt = 0..7 for selectors 17,18,20,21,23,24,26,27; b = next 8 stream
 *               bits (LSB first).  One 64-bit entry describes EVERY whole symbol that
 *               fits in those 8 bits, up to 7 output values:
 *                 bits  0..3   nv      values produced (1..7)
 *                 bits  4..31  cum[j]  bits consumed once value j+1 is complete (j<7);
 *                              both values of a "0 = two zeros" symbol carry the same
 *                              cum, which is the reference's tail rule (decode.c:216-218:
 *                              the symbol's bits are consumed even if one row remains)
 *                 bits 32..59  val[j]  the value, 4-bit two's complement (-4..4)
 *   t[tt][b]    tt = 0,1,2 for selectors 19 (t15), 22 (t27), 29 (t37); b = the 5/7-bit
 *               code.  16-bit entry: three 4-bit two's complement digits, bit 15 set
 *               when the code is out of range (ACM_ERR_CORRUPT, decode.c:412/438/464).
 *   sel13_r16[w]  scan-only, for the 16-row block shape: w = the 13 stream bits at a column
 *               boundary (5-bit selector + first 8 payload bits).  16-bit entry:
 *                 bits 0..8    bits to advance: 5 + the whole payload of a fixed-size filler
 *                              (zero, linear, t15/t27/t37), or 5 + the first k8 step
 *                 bits 9..12   rows still to come after that first step; non-zero exactly for
 *                              the prefix-coded fillers (a step yields at most 7 of 16 rows)
 *                 bits 13..15  k8 table number of that filler; a bad selector (f_bad,
 *                              decode.c:190-194) is entry >> 9 == 0x70 (table 7, no rows)
 *   kstep[kt][m][b]  scan-only: one prefix-code step with m = min(rows remaining, 7) and b = the
 *               next 8 stream bits.  8-bit entry: bits 0..3 bits consumed, bits 4..6 values
 *               produced (k8's nv and cum with the row cap already applied).
 */
#ifndef ACM_TABLES_H
#define ACM_TABLES_H

#include <stdint.h>

#define ACM_K8_TYPES 8
#define ACM_K8_SIZE (ACM_K8_TYPES * 256)
#define ACM_T_SIZE (3 * 128)

typedef struct acm_tables {
	uint64_t k8[ACM_K8_SIZE];
	uint16_t t[ACM_T_SIZE];
	uint16_t sel13_r16[8192];
	uint8_t kstep[8 * 8 * 256];
	uint8_t kind[32];  /* per selector: class | (subtype << 3); see ACM_CLS_* */
	uint8_t pad[32];
} acm_tables;

enum { ACM_CLS_ZERO = 0, ACM_CLS_LINEAR = 1, ACM_CLS_K = 2, ACM_CLS_T = 3, ACM_CLS_BAD = 4 };

#ifdef __cplusplus
extern "C" {
#endif
void acm_tables_build(acm_tables *t);
#ifdef __cplusplus
}
#endif

#endif
