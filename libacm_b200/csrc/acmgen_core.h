/*
 * acmgen_core.h -- the synthetic-stream generator proper (see acmgen.c for what it emits and
 * why), written so that the same code builds as C for the host library (acmgen.c) and as CUDA
 * device code for the on-GPU corpus generator (acm_gen.cu): the
 * two produce byte-identical images for the same parameters.
 */
#ifndef ACMGEN_CORE_H
#define ACMGEN_CORE_H

#include <stddef.h>
#include <stdint.h>

#include "acmgen.h"

#if defined(__CUDACC__)
#define ACMGEN_HD static __host__ __device__ __forceinline__
#else
#define ACMGEN_HD static inline
#endif

/* ------------------------------------------------------------ rng */

typedef struct { uint64_t s; } rng_t;

ACMGEN_HD uint64_t rng_next(rng_t *r)
{ /* splitmix64 */
	uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

ACMGEN_HD uint32_t rng_below(rng_t *r, uint32_t n)
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

ACMGEN_HD void bw_put(bitwr *w, uint32_t v, unsigned nbits)
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

ACMGEN_HD void bw_flush(bitwr *w)
{
	if (w->fill)
		bw_put(w, 0, 8 - w->fill);
}

/* ------------------------------------------------------------ fillers */

/* the 26 valid selectors (decode.c:480-489), k-th in ascending order */
ACMGEN_HD unsigned valid_ind(unsigned k)
{
	if (k == 0)
		return 0;
	if (k <= 22)
		return k + 2; /* 3..24 */
	return k == 23 ? 26 : (k == 24 ? 27 : 29);
}

/* smallest pwr for which every index the filler can emit lies in
 * [-2^pwr, 2^pwr - 1] (decode.c:592-600) */
ACMGEN_HD unsigned need_pwr(unsigned ind)
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
ACMGEN_HD void emit_column(bitwr *w, rng_t *r, unsigned ind, unsigned rows, unsigned pz)
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

ACMGEN_HD unsigned pick_ind(rng_t *r, const acmgen_params *p, unsigned col, unsigned cols)
{
	switch (p->dist) {
	case ACMGEN_DIST_SINGLE:
		return p->single_ind;
	case ACMGEN_DIST_STRESS:
		return valid_ind(rng_below(r, 26));
	default: { /* ACMGEN_DIST_FALLOUT: SURVEY.md section 8(d).  No real game
		    * files exist here, so this is an ASSUMED spectrum: energy falls
		    * with the subband (column) index.  Bands by f = col/cols:
		    *   f < 1/16  linear 7..10 bits     f < 1/8  linear 5..7 bits
		    *   f < 1/4   linear 4..5 bits      else     zero / k- / t-codes
		    * which lands at about 3.4 bit/sample with pzero = 0.5. */
		/* { 0, 17, 18, 19, 20, 21, 22, 22, 23, 24, 24, 26, 27, 27, 29, 29 } as 16 bytes */
		const uint64_t hi_lo = 0x1616151413121100ull, hi_hi = 0x1D1D1B1B1A181817ull;
		unsigned f16 = (col * 16) / cols;
		if (cols < 16)
			f16 = col ? 4 : 0;
		if (f16 < 1)
			return 7 + rng_below(r, 4);
		if (f16 < 2)
			return 5 + rng_below(r, 3);
		if (f16 < 4)
			return 4 + rng_below(r, 2);
		{
			const unsigned k = rng_below(r, 16);
			return (unsigned)(((k < 8 ? hi_lo : hi_hi) >> (8 * (k & 7))) & 0xFF);
		}
	}
	}
}

ACMGEN_HD size_t acmgen_bound_core(const acmgen_params *p)
{
	uint64_t cols = 1ull << p->level, blen = cols * p->rows;
	uint64_t nblocks = (p->total_values + blen - 1) / blen;
	uint64_t bits = 20 + cols * (5 + 16ull * p->rows);
	return (size_t)(42 + (nblocks * bits + 7) / 8 + 8);
}

/* Writes the stream for p into out[0..cap) and returns its size in bytes; with count_only the
 * size is returned and nothing is written (cap is ignored).  inds = cols bytes of scratch.
 * Returns 0 if the image does not fit. */
ACMGEN_HD size_t acmgen_write_core(const acmgen_params *p, uint8_t *out, size_t cap, uint8_t *inds, int count_only)
{
	bitwr w;
	rng_t r;
	uint32_t cols = 1u << p->level, blen = cols * p->rows;
	uint64_t nblocks = ((uint64_t)p->total_values + blen - 1) / blen, b;
	unsigned c;

	w.p = out;
	w.cap = count_only ? 0 : cap;
	w.n = 0;
	w.acc = 0;
	w.fill = 0;
	w.overflow = 0;
	r.s = p->seed * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull;

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
	if (w.overflow && !count_only)
		return 0;
	return w.n;
}


#endif
