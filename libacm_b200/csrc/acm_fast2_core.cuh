/*
 * acm_fast2_core.cuh -- the lane-local pieces of the level-7 / 16-row kernel (acm_fast2.cu):
 * the table-driven column walk and the three column unpackers.  They are pure functions of
 * the staged bits and the code tables, written __host__ __device__ so that the CPU-only test
 * tier (tests/emu/emu_fast2.cpp) runs the very code the kernel runs.
 */
#pragma once

#include <stdint.h>

#include "acm_device.cuh"

#ifndef F2_NIB_ALU
#define F2_NIB_ALU 1   /* nibbles -> int16 pairs by arithmetic (0: nib2w table in shared memory) */
#endif

namespace acm {
namespace fast2 {

constexpr int ROWS = 16;

/* funnel shift right: the low 32 bits of (hi:lo) >> (sh & 31) */
ACM_HD uint32_t fsr(uint32_t lo, uint32_t hi, uint32_t sh)
{
#if defined(__CUDA_ARCH__)
	return __funnelshift_r(lo, hi, sh);
#else
	sh &= 31u;
	return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}

/* ------------------------------------------------------------------ walk */

/* uni16 addressing in bytes: page id << UNI_PSHIFT; the selector page starts at 0 */
constexpr uint32_t UNI_PSHIFT = ACM_UNI_KBITS + 1;
constexpr uint32_t UNI_HALT8 = (uint32_t)ACM_UNI_HALT << UNI_PSHIFT;
constexpr uint32_t UNI_BAD8 = (uint32_t)ACM_UNI_BAD << UNI_PSHIFT;
/* index masks, doubled (a table entry is two bytes) */
constexpr uint32_t MSK_SEL = 0x1FFFu << 1, MSK_K = ((1u << ACM_UNI_KBITS) - 1u) << 1;

/*
 * State of one lane's column walk.  Q = bit position - 1, s8 = byte offset of the current page
 * in uni16, msk = doubled index mask of that page (13 bits at a selector, ACM_UNI_KBITS inside a
 * prefix-coded column).  One step: w1 = the 32 stream bits at Q, i.e. the bits at the position
 * shifted up by one, which is the byte offset of the 16-bit entry once masked; pages are aligned
 * to their size, so page offset and entry offset combine with an OR (one LOP3 with the mask);
 * e = the uni16 entry at walk_index(s, w1); walk_next(s, e).
 */
struct Walk {
	uint32_t Q, s8, msk;
	uint32_t Q32; /* Q << 5, kept alongside (the scan's ring address is a mask of it) */
};

ACM_HD uint32_t walk_index(const Walk &s, uint32_t w1) { return s.s8 | (w1 & s.msk); }

/* apply entry e; returns true when the lane is at a column selector afterwards */
ACM_HD bool walk_next(Walk &s, uint32_t e)
{
	s.Q += e & 0xFFu;
	s.Q32 = s.Q << 5;
	s.s8 = (e & 0xFF00u) << (UNI_PSHIFT - 8u);
	const bool at_sel = s.s8 == 0u;
	s.msk = at_sel ? MSK_SEL : MSK_K;
	return at_sel;
}

/* the same for a lane that may have to sit this step out (its bits have not landed): with
 * go == false nothing moves.  On the device the advance is one dot-product instruction,
 * Q + byte0(e) * go, which keeps the chain from the table load to the next ring address short. */
ACM_HD bool walk_next_if(Walk &s, uint32_t e, bool go)
{
#if defined(__CUDA_ARCH__)
	s.Q = __dp4a(e, go ? 1u : 0u, s.Q);
	s.Q32 = __dp4a(e, go ? 32u : 0u, s.Q32);
#else
	s.Q += go ? (e & 0xFFu) : 0u;
	s.Q32 = s.Q << 5;
#endif
	const uint32_t nx = (e & 0xFF00u) << (UNI_PSHIFT - 8u);
	s.s8 = go ? nx : s.s8;
	const bool at_sel = s.s8 == 0u;
	s.msk = at_sel ? MSK_SEL : MSK_K;
	return at_sel;
}

/* ------------------------------------------------------------------ unpack */

/* sixteen int16 of a column in two 128-bit halves (rows 0-7, rows 8-15) */
struct ColOut {
	uint32_t *h0, *h1;
};

ACM_HD void store8(uint32_t *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
#if defined(__CUDA_ARCH__)
	*reinterpret_cast<uint4 *>(p) = make_uint4(a, b, c, d);
#else
	p[0] = a; p[1] = b; p[2] = c; p[3] = d;
#endif
}

/* two 4-bit two's-complement values in a byte -> two int16 in a word, without a table: spread the
 * nibbles to bits 0-3 and 16-19 (x 0x1001, mask), then fill bits 4-15 / 20-31 of each half from
 * its sign bit (0x0008 x 0x1FFE = 0xFFF0; the two products cannot meet) */
ACM_HD uint32_t nib2w_alu(uint32_t byte)
{
	const uint32_t x = (byte * 0x1001u) & 0x000F000Fu;
	return (x & 0x00080008u) * 0x1FFEu + x;
}

/* 16 nibbles (a0: rows 0-7, a1: rows 8-15) -> sixteen int16 */
ACM_HD void store_nibbles(uint32_t a0, uint32_t a1, const ColOut &o, const uint32_t *nib2w)
{
#if F2_NIB_ALU
	(void)nib2w;
	store8(o.h0, nib2w_alu(a0 & 255u), nib2w_alu((a0 >> 8) & 255u), nib2w_alu((a0 >> 16) & 255u), nib2w_alu(a0 >> 24));
	store8(o.h1, nib2w_alu(a1 & 255u), nib2w_alu((a1 >> 8) & 255u), nib2w_alu((a1 >> 16) & 255u), nib2w_alu(a1 >> 24));
#else
	store8(o.h0, nib2w[a0 & 255u], nib2w[(a0 >> 8) & 255u], nib2w[(a0 >> 16) & 255u], nib2w[a0 >> 24]);
	store8(o.h1, nib2w[a1 & 255u], nib2w[(a1 >> 8) & 255u], nib2w[(a1 >> 16) & 255u], nib2w[a1 >> 24]);
#endif
}

/*
 * Prefix codes (decode.c:208-403).  One k8w entry per step: up to 8 values, the bits and the
 * rows they cover.  No row cap: the 16 rows are nibbles of a 64-bit accumulator and whatever is
 * decoded past the 16th row (the bits of the next column) shifts out; the reference's tail rule
 * (a "two zeros" symbol at the last row emits one, decode.c:216-218) is the same thing.
 * SR::word(i) = 32-bit word i of the stream; words up to (P >> 5) + 3 are read.
 */
template <typename SR>
ACM_HD void unpack_k(const SR &sr, uint32_t P, uint32_t sub, const ColOut &o, const uint64_t *k8w,
		     const uint32_t *nib2w)
{
	const uint32_t i = P >> 5, sh = P & 31u;
	const uint32_t w0 = sr.word(i), w1 = sr.word(i + 1), w2 = sr.word(i + 2), w3 = sr.word(i + 3);
	uint32_t lo = fsr(w0, w1, sh), mid = fsr(w1, w2, sh), hi = fsr(w2, w3, sh);
	const uint64_t *tab = k8w + sub * 256u;
	uint32_t a0 = 0u, a1 = 0u, r4 = 0u;
	do {
		const uint64_t e = tab[lo & 255u];
		const uint32_t ev = (uint32_t)e, em = (uint32_t)(e >> 32);
		const uint32_t bits = em & 15u;
		lo = fsr(lo, mid, bits);
		mid = fsr(mid, hi, bits);
		hi >>= bits;
		const unsigned long long vv = (unsigned long long)ev << r4;
		a0 |= (uint32_t)vv;
		a1 |= (uint32_t)(vv >> 32);
		r4 += em >> 8;
	} while (r4 < 64u);
	store_nibbles(a0, a1, o, nib2w);
}

/* f_t15 / f_t27 / f_t37 (decode.c:405-476): all codes sit in one 64-bit window.  Returns
 * non-zero if a code that the reference gets to read is out of range (decode.c:412/:438/:464). */
template <typename SR>
ACM_HD int unpack_t(const SR &sr, uint32_t P, uint32_t limit, uint32_t sub, const ColOut &o, const uint16_t *tt,
		    const uint32_t *nib2w, bool store)
{
	const uint32_t width = sub == 0 ? 5u : 7u, per = sub == 2 ? 2u : 3u;
	const uint32_t ncodes = sub == 2 ? 8u : 6u, cmask = (1u << width) - 1u;
	const uint32_t vmask = sub == 2 ? 0xFFu : 0xFFFu;
	const uint16_t *tab = tt + sub * 128u;
	const uint32_t i = P >> 5, sh = P & 31u;
	const uint32_t w0 = sr.word(i), w1 = sr.word(i + 1), w2 = sr.word(i + 2);
	const unsigned long long win = (unsigned long long)fsr(w0, w1, sh) | ((unsigned long long)fsr(w1, w2, sh) << 32);
	/* codes the reference gets to read before the stream runs dry: all of them, except in the
	 * last block of a truncated stream */
	const uint32_t nread = P + ncodes * width <= limit ? ncodes : (limit > P ? (limit - P) / width : 0u);
	uint32_t a0 = 0u, a1 = 0u, seen = 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int q = 0; q < 8; q++) {
		if ((uint32_t)q < ncodes) {
			const uint32_t e = tab[(uint32_t)(win >> (q * width)) & cmask];
			seen |= (uint32_t)q < nread ? e : 0u;
			const unsigned long long vv = (unsigned long long)(e & vmask) << (4u * q * per);
			a0 |= (uint32_t)vv;
			a1 |= (uint32_t)(vv >> 32);
		}
	}
	if (store)
		store_nibbles(a0, a1, o, nib2w);
	return (seen & 0x8000u) != 0u;
}

/* f_linear (decode.c:196-206): sliding 64-bit window; the word that the next refill will need
 * is fetched one refill ahead, so that no value waits on a load */
template <typename SR>
ACM_HD void unpack_linear(const SR &sr, uint32_t P, uint32_t ind, const ColOut &o)
{
	const uint32_t mask = (1u << ind) - 1u;
	const int mid = 1 << (ind - 1);
	const uint32_t i = P >> 5, sh = P & 31u;
	unsigned long long win = (((unsigned long long)sr.word(i + 1) << 32) | sr.word(i)) >> sh;
	uint32_t avail = 64u - sh, nx = i + 3;
	uint32_t nextw = sr.word(i + 2);
	uint32_t w[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int r = 0; r < ROWS; r++) {
		const uint32_t v = (uint32_t)((int)((uint32_t)win & mask) - mid);
		if (r & 1)
			w[r >> 1] |= v << 16;
		else
			w[r >> 1] = v & 0xFFFFu;
		win >>= ind;
		avail -= ind;
		if (avail <= 32u) {
			win |= (unsigned long long)nextw << avail;
			avail += 32u;
			nextw = sr.word(nx);
			nx++;
		}
	}
	store8(o.h0, w[0], w[1], w[2], w[3]);
	store8(o.h1, w[4], w[5], w[6], w[7]);
}

} // namespace fast2
} // namespace acm
