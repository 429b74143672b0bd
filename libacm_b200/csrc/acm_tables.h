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
 *   k8w[kt][b]  the unpack-side variant of k8 (acm_fast2.cu): every whole symbol that fits in
 *               the 8 bits b while at most 8 rows are produced.  64-bit entry:
 *                 bits  0..31  the (up to 8) values, 4-bit two's complement, value j in nibble j
 *                 bits 32..35  bits consumed (1..8)
 *                 bits 40..45  4 * rows produced (4..32)
 *               No row cap: the caller accumulates 16 nibbles in a 64-bit register and lets
 *               rows past the 16th shift out; where the column ends is known from the scan.
 *   uni16[page][b]  the whole column-length walk of a 16-row block as ONE state machine: an entry
 *               is (bits to advance) | (next page id << 8).  The state "at a column selector" is
 *               the 8192-entry page 0 (index = 13 stream bits: selector + first payload byte);
 *               "inside a prefix-coded column of type kt with rem rows to come" is page
 *               ACM_UNI_K0 + kt*15 + rem-1 (index = ACM_UNI_KBITS stream bits; whole symbols, row cap applied);
 *               a fixed-size filler advances over its whole payload in one step (the 261 bits of a
 *               16-bit linear column as 255 + the SKIP6 page); a bad selector (f_bad,
 *               decode.c:190-194) leads to the BAD page; HALT and BAD entries advance 0 bits
 *               and stay, so a lane that is done, parked or corrupt runs the same instructions.
 *   nib2w[b]    b = two nibbles (rows r, r+1) -> (int16)lo | (int16)hi << 16.
 */
#ifndef ACM_TABLES_H
#define ACM_TABLES_H

#include <stdint.h>

#define ACM_K8_TYPES 8
#define ACM_K8_SIZE (ACM_K8_TYPES * 256)
#define ACM_T_SIZE (3 * 128)
/* uni16 page ids; a page = 2^ACM_UNI_KBITS entries, so (id << (ACM_UNI_KBITS + 1)) is the page's
 * byte offset */
#ifndef ACM_UNI_KBITS
#define ACM_UNI_KBITS 8   /* index width of the prefix-code pages (7 or 8) */
#endif
#define ACM_UNI_PSIZE (1 << ACM_UNI_KBITS)
#define ACM_UNI_SEL 0                            /* the 8192-entry selector page: ids 0 .. K0-1 */
#define ACM_UNI_K0 (8192 >> ACM_UNI_KBITS)       /* (kt, rem) -> K0 + kt * 15 + (rem - 1), rem = 1..15 */
#define ACM_UNI_HALT (ACM_UNI_K0 + 120)
#define ACM_UNI_BAD (ACM_UNI_K0 + 121)
#define ACM_UNI_SKIP6 (ACM_UNI_K0 + 122)
#define ACM_UNI_PAGES (ACM_UNI_K0 + 123)

typedef struct acm_tables {
	uint64_t k8[ACM_K8_SIZE];
	uint16_t t[ACM_T_SIZE];
	uint8_t kind[32];  /* per selector: class | (subtype << 3); see ACM_CLS_* */
	uint8_t pad[32];
	uint64_t k8w[ACM_K8_SIZE]; /* worker-side prefix-code step, see above */
	uint16_t uni16[ACM_UNI_PAGES * ACM_UNI_PSIZE]; /* scan walk of the fast kernel, see above */
	uint32_t nib2w[256];       /* two 4-bit two's complement values -> two int16 in one word */
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
