/*
 * acmgen.c -- synthetic InterPlay ACM stream generator.
 *
 * markokr/libacm ships no encoder and no sample files, so every input this
 * repository decodes is produced here.  The generator emits *valid* bitstreams
 * (verified against the reference decoder) that exercise
 * every filler of the reference's table (decode.c:480-489), every level/row
 * shape and both container variants (plain 14-byte header, decode.c:712-752;
 * 28-byte WAVC pre-header, decode.c:687-710), and obeys the constraints that
 * keep the reference's behaviour defined (Appendix D): pwr >= need(ind) so the
 * dequantisation table covers every index (decode.c:591-600), t-codes below
 * their radix limit, no bad selector unless a negative test asks for one.
 *
 * It draws code SYMBOLS, not PCM: any valid symbol sequence is a valid stream,
 * and the decoder under test (reference, oracle or GPU) defines the PCM.
 *
 * Built both as a shared library (ctypes: acmgen_bound / acmgen_write /
 * acmgen_write_many) and, with -DACMGEN_MAIN, as a CLI.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "acmgen.h"

/* ------------------------------------------------------------ rng */

typedef struct { uint64_t s; } rng_t;

static inline uint64_t rng_next(rng_t *r)
{ /* splitmix64 */
	uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

static inline uint32_t rng_below(rng_t *r, uint32_t n)
{
	return (uint32_t)(((rng_next(r) >> 32) * (uint64_t)n) >> 32);
}

/* ------------------------------------------------------------ bit writer */

typedef struct {
	uint8_t *p;
	size_t cap, n; /* bytes written */
	uint64_t acc;
	unsigned fill;
	int overflow;
} bitwr;

static inline void bw_put(bitwr *w, uint32_t v, unsigned nbits)
{ /* LSB-first, little-endian fields (decode.c:117-133) */
	w->acc |= (uint64_t)(v & ((nbits >= 32) ? 0xFFFFFFFFu : ((1u << nbits) - 1))) << w->fill;
	w->fill += nbits;
	while (w->fill >= 8) {
		if (w->n < w->cap)
			w->p[w->n] = (uint8_t)w->acc;
		else
			w->overflow = 1;
		w->n++;
		w->acc >>= 8;
		w->fill -= 8;
	}
}

static inline void bw_flush(bitwr *w)
{
	if (w->fill)
		bw_put(w, 0, 8 - w->fill);
}

/* ------------------------------------------------------------ fillers */

static const uint8_t valid_inds[26] = { 0, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16,
					17, 18, 19, 20, 21, 22, 23, 24, 26, 27, 29 };

/* smallest pwr for which every index the filler can emit lies in
 * [-2^pwr, 2^pwr - 1] (decode.c:592-600) */
static unsigned need_pwr(unsigned ind)
{
	if (ind == 0) return 0;
	if (ind >= 3 && ind <= 16) return ind - 1;
	switch (ind) {
	case 17: case 18: case 19: return 1;
	case 20: case 21: case 22: case 23: case 24: return 2;
	default: return 3; /* 26 27 29 (and bad codes: irrelevant) */
	}
}

/* Emit the payload of one column.  pz = probability (0..255)/256 of choosing the
 * zero symbol where the code has one. */
static void emit_column(bitwr *w, rng_t *r, unsigned ind, unsigned rows, unsigned pz)
{
	unsigned i = 0;
	if (ind == 0)
		return;
	if (ind >= 3 && ind <= 16) {
		for (; i < rows; i++)
			bw_put(w, (uint32_t)rng_next(r), ind);
		return;
	}
	switch (ind) {
	case 19: /* t15: 5-bit code < 27, 3 values */
		for (; i < rows; i += 3)
			bw_put(w, rng_below(r, 27), 5);
		return;
	case 22: /* t27: 7-bit code < 125, 3 values */
		for (; i < rows; i += 3)
			bw_put(w, rng_below(r, 125), 7);
		return;
	case 29: /* t37: 7-bit code < 121, 2 values */
		for (; i < rows; i += 2)
			bw_put(w, rng_below(r, 121), 7);
		return;
	}
	/* k-codes (Appendix A.4) */
	{
		int pair = (ind == 17 || ind == 20 || ind == 23 || ind == 26);
		while (i < rows) {
			uint32_t u = (uint32_t)rng_next(r);
			if ((u & 0xFF) < pz) {
				if (pair && ((u >> 8) & 1)) {
					bw_put(w, 0, 1); /* "0": two zeros (one at the tail) */
					i += 2;
				} else if (pair) {
					bw_put(w, 1, 2); /* "1 0": one zero */
					i += 1;
				} else {
					bw_put(w, 0, 1);
					i += 1;
				}
				continue;
			}
			u >>= 9;
			switch (ind) {
			case 17: bw_put(w, 3 | ((u & 1) << 2), 3); break;          /* 1 1 x   */
			case 18: bw_put(w, 1 | ((u & 1) << 1), 2); break;          /* 1 x     */
			case 20: bw_put(w, 3 | ((u & 3) << 2), 4); break;          /* 1 1 xx  */
			case 21: bw_put(w, 1 | ((u & 3) << 1), 3); break;          /* 1 xx    */
			case 23:
				if (u & 4) bw_put(w, 3 | ((u & 1) << 3), 4);       /* 1 1 0 x  */
				else bw_put(w, 7 | ((u & 3) << 3), 5);             /* 1 1 1 xx */
				break;
			case 24:
				if (u & 4) bw_put(w, 1 | ((u & 1) << 2), 3);       /* 1 0 x   */
				else bw_put(w, 3 | ((u & 3) << 2), 4);             /* 1 1 xx  */
				break;
			case 26: bw_put(w, 3 | ((u & 7) << 2), 5); break;          /* 1 1 xxx */
			default: bw_put(w, 1 | ((u & 7) << 1), 4); break;          /* 27: 1 xxx */
			}
			i += 1;
		}
	}
}

/* ------------------------------------------------------------ stream */

static unsigned pick_ind(rng_t *r, const acmgen_params *p, unsigned col, unsigned cols)
{
	switch (p->dist) {
	case ACMGEN_DIST_SINGLE:
		return p->single_ind;
	case ACMGEN_DIST_STRESS:
		return valid_inds[rng_below(r, 26)];
	default: { /* ACMGEN_DIST_FALLOUT: SURVEY.md section 8(d).  No real game
		    * files exist here, so this is an ASSUMED spectrum: energy falls
		    * with the subband (column) index.  Bands by f = col/cols:
		    *   f < 1/16  linear 7..10 bits     f < 1/8  linear 5..7 bits
		    *   f < 1/4   linear 4..5 bits      else     zero / k- / t-codes
		    * which lands at about 3.4 bit/sample with pzero = 0.5. */
		static const uint8_t hi[16] = { 0, 17, 18, 19, 20, 21, 22, 22, 23, 24, 24, 26, 27, 27, 29, 29 };
		unsigned f16 = (col * 16) / cols;
		if (cols < 16)
			f16 = col ? 4 : 0;
		if (f16 < 1)
			return 7 + rng_below(r, 4);
		if (f16 < 2)
			return 5 + rng_below(r, 3);
		if (f16 < 4)
			return 4 + rng_below(r, 2);
		return hi[rng_below(r, 16)];
	}
	}
}

size_t acmgen_bound(const acmgen_params *p)
{
	uint64_t cols = 1ull << p->level, blen = cols * p->rows;
	uint64_t nblocks = (p->total_values + blen - 1) / blen;
	uint64_t bits = 20 + cols * (5 + 16ull * p->rows);
	return (size_t)(42 + (nblocks * bits + 7) / 8 + 8);
}

size_t acmgen_write(const acmgen_params *p, uint8_t *out, size_t cap)
{
	bitwr w;
	rng_t r;
	uint32_t cols = 1u << p->level, blen = cols * p->rows;
	uint64_t nblocks = ((uint64_t)p->total_values + blen - 1) / blen, b;
	uint8_t *inds;
	unsigned c;

	memset(&w, 0, sizeof(w));
	w.p = out;
	w.cap = cap;
	r.s = p->seed * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull;
	inds = malloc(cols);
	if (!inds)
		return 0;

	if (p->wavc) { /* Appendix A.2 */
		uint32_t raw = p->total_values * 2;
		bw_put(&w, 0x564157, 24);
		bw_put(&w, 'C', 8);
		bw_put(&w, 0x3156, 16);
		bw_put(&w, 0x302E, 16);
		bw_put(&w, raw & 0xFFFF, 16);
		bw_put(&w, raw >> 16, 16);
		bw_put(&w, 0, 16); /* compressed size: unchecked (decode.c:701-703) */
		bw_put(&w, 0, 16);
		bw_put(&w, 28, 16);
		bw_put(&w, 0, 16);
		bw_put(&w, p->channels, 16);
		bw_put(&w, 16, 16);
		bw_put(&w, p->rate, 16);
		bw_put(&w, 0, 16);
	}
	bw_put(&w, 0x032897, 24); /* ACM_ID libacm.h:28 */
	bw_put(&w, 1, 8);
	bw_put(&w, p->total_values & 0xFFFF, 16);
	bw_put(&w, p->total_values >> 16, 16);
	bw_put(&w, p->channels, 16);
	bw_put(&w, p->rate, 16);
	bw_put(&w, p->level, 4);
	bw_put(&w, p->rows, 12);

	for (b = 0; b < nblocks; b++) {
		unsigned need = 0, pwr, val;
		for (c = 0; c < cols; c++) {
			unsigned ind = pick_ind(&r, p, c, cols), n;
			if (p->inject == ACMGEN_INJECT_BAD_IND && b == p->inject_block && c == p->inject_col)
				ind = p->inject_value;
			inds[c] = (uint8_t)ind;
			n = need_pwr(ind);
			if (n > need)
				need = n;
		}
		if (p->dist == ACMGEN_DIST_FALLOUT) {
			pwr = need + rng_below(&r, 3);
			val = 1 + rng_below(&r, 4096);
		} else {
			pwr = need + rng_below(&r, 16 - need);
			val = rng_below(&r, 65536);
		}
		if (pwr > 15)
			pwr = 15;
		bw_put(&w, pwr, 4);
		bw_put(&w, val, 16);
		for (c = 0; c < cols; c++) {
			unsigned ind = inds[c];
			bw_put(&w, ind, 5);
			if (p->inject == ACMGEN_INJECT_BAD_TCODE && b == p->inject_block &&
			    c == p->inject_col && (ind == 19 || ind == 22 || ind == 29)) {
				/* first code of the column is out of range (decode.c:412/438/464) */
				unsigned width = ind == 19 ? 5 : 7, lim = ind == 19 ? 27 : (ind == 22 ? 125 : 121);
				unsigned step = ind == 29 ? 2 : 3, i;
				bw_put(&w, lim + rng_below(&r, (1u << width) - lim), width);
				for (i = step; i < p->rows; i += step)
					bw_put(&w, 0, width);
				continue;
			}
			if (ind == 1 || ind == 2 || ind == 25 || ind == 28 || ind >= 30)
				continue; /* bad selector: decoder stops here */
			emit_column(&w, &r, ind, p->rows, p->pzero);
		}
	}
	bw_flush(&w);
	free(inds);
	if (w.overflow)
		return 0;
	return w.n;
}

/*
 * Batch helper: n streams, stream i uses params[i]; images are packed
 * back-to-back, each starting on a 16-byte boundary, into blob[0..cap).
 * offs[i]/lens[i] receive the placement.  Returns bytes used or 0 on overflow.
 */
size_t acmgen_write_many(const acmgen_params *params, size_t n, uint8_t *blob, size_t cap,
			 uint64_t *offs, uint32_t *lens)
{
	size_t i, at = 0;
	for (i = 0; i < n; i++) {
		size_t got;
		at = (at + 15) & ~(size_t)15;
		if (at >= cap)
			return 0;
		got = acmgen_write(&params[i], blob + at, cap - at);
		if (!got)
			return 0;
		offs[i] = at;
		lens[i] = (uint32_t)got;
		at += got;
	}
	return at;
}

#ifdef ACMGEN_MAIN
int main(int argc, char **argv)
{
	acmgen_params p;
	uint8_t *buf;
	size_t cap, n;
	FILE *f;
	if (argc < 8) {
		fprintf(stderr, "usage: acmgen OUT level rows channels rate total_values seed [dist 0|1|2] [wavc] [single_ind]\n");
		return 2;
	}
	memset(&p, 0, sizeof(p));
	p.level = atoi(argv[2]);
	p.rows = atoi(argv[3]);
	p.channels = atoi(argv[4]);
	p.rate = atoi(argv[5]);
	p.total_values = strtoul(argv[6], NULL, 0);
	p.seed = strtoull(argv[7], NULL, 0);
	p.dist = argc > 8 ? atoi(argv[8]) : ACMGEN_DIST_FALLOUT;
	p.wavc = argc > 9 ? atoi(argv[9]) : 0;
	p.single_ind = argc > 10 ? atoi(argv[10]) : 0;
	p.pzero = 128;
	cap = acmgen_bound(&p);
	buf = malloc(cap);
	n = acmgen_write(&p, buf, cap);
	if (!n || !(f = fopen(argv[1], "wb")))
		return 1;
	fwrite(buf, 1, n, f);
	fclose(f);
	return 0;
}
#endif
