/*
 * acm_fast2_core.cuh -- the lane-local pieces of the level-7 / 16-row kernel (acm_fast2.cu):
 * the table-driven column walk and the three column unpackers.  They are pure functions of
 * the staged bits and the code tables, written __host__ __device__ so that the CPU-only test
 * tier (tests/emu/emu_fast2.cpp) runs the very code the kernel runs.
 */
#pragma once

#include <stdint.h>

#include "acm_device.cuh"

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

/* "does any lane of the warp ...": lets a warp skip work none of its lanes has (on the host, where
 * the test tier runs one lane at a time, the lane's own answer) */
#if defined(__CUDA_ARCH__)
#define ACM_WARP_ANY(x) (__any_sync(0xFFFFFFFFu, (x)) != 0)
#else
#define ACM_WARP_ANY(x) (x)
#endif

/*
 * The three column unpackers.  A decode lane owns whole columns (column = lane mod 32, the
 * layout lifting stages 1 and 2 run in), so a column's sixteen values never leave the lane's
 * registers: prefix- and radix-coded columns end up as sixteen 4-bit two's complement
 * values in two registers (a0: rows 0-7, a1: rows 8-15), linear columns as sixteen
 * dequantised words.  `lo`, `mid`, `hi` = the 96 stream bits that start at the column's
 * payload.
 */

/* sign-extend nibble j of x and dequantise (set_pos / midbuf, decode.c:174-177, :591-600) */
ACM_HD uint32_t nib_val(uint32_t x, int j, int val) { return (uint32_t)(nib_s(x, j) * val); }

/*
 * Prefix codes (decode.c:208-403).  One k8w entry per step: up to 8 values, the bits and the
 * rows they cover.  No row cap: the 16 rows are nibbles of a 64-bit accumulator and whatever is
 * decoded past the 16th row (the bits of the next column) shifts out; the reference's tail rule
 * (a "two zeros" symbol at the last row emits one, decode.c:216-218) is the same thing.
 * A column is at most 80 bits, so the 96-bit window never runs dry.
 */
ACM_HD void unpack_k(uint32_t lo, uint32_t mid, uint32_t hi, uint32_t sub, const uint64_t *k8w, uint32_t &a0,
		     uint32_t &a1)
{
	const uint64_t *tab = k8w + sub * 256u;
	uint32_t r4 = 0u;
	a0 = a1 = 0u;
	do {
		const uint64_t e = tab[lo & 255u];
		const uint32_t ev = (uint32_t)e, em = (uint32_t)(e >> 32);
		/* em = bits consumed (1..8, bits 4-7 clear) | 4 * rows << 8: a funnel shift takes its
		 * count from the low five bits, so em itself is the shift operand */
		lo = fsr(lo, mid, em);
		mid = fsr(mid, hi, em);
		hi = fsr(hi, 0u, em);
		const unsigned long long vv = (unsigned long long)ev << r4;
		a0 |= (uint32_t)vv;
		a1 |= (uint32_t)(vv >> 32);
#if defined(__CUDA_ARCH__)
		r4 = __dp4a(em, 0x00000100u, r4); /* += byte 1 of em: one instruction, on the other pipe */
#else
		r4 += (em >> 8) & 0xFFu;
#endif
	} while (r4 < 64u);
}

/* f_t15 / f_t27 / f_t37 (decode.c:405-476): all codes sit in the first 64 bits of the window.
 * P = position of the payload, limit = first position that cannot be read.  Returns non-zero
 * if a code that the reference gets to read is out of range (decode.c:412/:438/:464). */
ACM_HD int unpack_t(uint32_t lo, uint32_t mid, uint32_t P, uint32_t limit, uint32_t sub, const uint16_t *tt,
		    uint32_t &a0, uint32_t &a1)
{
	const uint32_t width = sub == 0 ? 5u : 7u, per = sub == 2 ? 2u : 3u;
	const uint32_t ncodes = sub == 2 ? 8u : 6u, cmask = (1u << width) - 1u;
	const uint32_t vmask = sub == 2 ? 0xFFu : 0xFFFu;
	const uint16_t *tab = tt + sub * 128u;
	const unsigned long long win = (unsigned long long)lo | ((unsigned long long)mid << 32);
	/* codes the reference gets to read before the stream runs dry: all of them, except in the
	 * last block of a truncated stream */
	const uint32_t nread = P + ncodes * width <= limit ? ncodes : (limit > P ? (limit - P) / width : 0u);
	uint32_t seen = 0u;
	a0 = a1 = 0u;
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
	return (seen & 0x8000u) != 0u;
}

/*
 * The same two unpackers for FOUR columns at once (a decode lane's columns 32 p + lane, p = 0..3).
 * One column is one chain of dependent operations -- table lookup, shift, next lookup -- and a
 * warp that has a single chain in flight issues an instruction every ten cycles or so; four
 * independent chains per lane fill those gaps.  cls[p] = ACM_CLS_* of column p (anything but
 * ACM_CLS_K / ACM_CLS_T: the column is left alone), sub[p] its sub-type.
 */
ACM_HD void unpack_k4(uint32_t (&lo)[4], uint32_t (&mid)[4], uint32_t (&hi)[4], const uint32_t (&cls)[4],
		      const uint32_t (&sub)[4], const uint64_t *k8w, uint32_t (&a0)[4], uint32_t (&a1)[4])
{
	uint32_t r4[4];
	const uint64_t *tab[4];
	uint32_t busy = 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int p = 0; p < 4; p++) {
		const bool isk = cls[p] == ACM_CLS_K;
		r4[p] = isk ? 0u : 64u;
		tab[p] = k8w + (isk ? sub[p] : 0u) * 256u;
		busy |= isk ? 1u : 0u;
	}
	while (busy) {
		busy = 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (int p = 0; p < 4; p++) {
			/* a finished (or other) column looks up nothing: entry 0 = no bits, no rows, no values */
			const uint64_t e = r4[p] < 64u ? tab[p][lo[p] & 255u] : 0ull;
			const uint32_t ev = (uint32_t)e, em = (uint32_t)(e >> 32);
			const uint32_t bits = em & 15u;
			lo[p] = fsr(lo[p], mid[p], bits);
			mid[p] = fsr(mid[p], hi[p], bits);
			hi[p] >>= bits;
			const unsigned long long vv = (unsigned long long)ev << (r4[p] & 63u);
			a0[p] |= (uint32_t)vv;
			a1[p] |= (uint32_t)(vv >> 32);
			r4[p] += em >> 8;
			busy |= r4[p] < 64u ? 1u : 0u;
		}
	}
}

/* P[p] = position of column p's payload; returns non-zero if a radix code that the reference gets
 * to read is out of range */
ACM_HD int unpack_t4(const uint32_t (&lo)[4], const uint32_t (&mid)[4], const uint32_t (&P)[4], uint32_t limit,
		     const uint32_t (&cls)[4], const uint32_t (&sub)[4], const uint16_t *tt, uint32_t (&a0)[4],
		     uint32_t (&a1)[4])
{
	uint32_t seen = 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int p = 0; p < 4; p++) {
		const bool ist = cls[p] == ACM_CLS_T;
		if (!ACM_WARP_ANY(ist))
			continue;
		const uint32_t sb = ist ? sub[p] : 0u;
		const uint32_t width = sb == 0 ? 5u : 7u, per = sb == 2 ? 2u : 3u;
		const uint32_t ncodes = !ist ? 0u : sb == 2 ? 8u : 6u, cmask = (1u << width) - 1u;
		const uint32_t vmask = sb == 2 ? 0xFFu : 0xFFFu;
		const uint16_t *tab = tt + sb * 128u;
		const unsigned long long win = (unsigned long long)lo[p] | ((unsigned long long)mid[p] << 32);
		const uint32_t nread =
			P[p] + ncodes * width <= limit ? ncodes : (limit > P[p] ? (limit - P[p]) / width : 0u);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (int q = 0; q < 8; q++) {
			if ((uint32_t)q < ncodes) {
				const uint32_t e = tab[(uint32_t)(win >> (q * width)) & cmask];
				seen |= (uint32_t)q < nread ? e : 0u;
				const unsigned long long vv = (unsigned long long)(e & vmask) << (4u * q * per);
				a0[p] |= (uint32_t)vv;
				a1[p] |= (uint32_t)(vv >> 32);
			}
		}
	}
	return (seen & 0x8000u) != 0u;
}

/* f_linear (decode.c:196-206): sliding 64-bit window; the word that the next refill will need
 * is fetched one refill ahead, so that no value waits on a load.  SR::word(i) = 32-bit word i
 * of the stream; v[r] = (code - 2^(ind-1)) * val. */
template <typename SR>
ACM_HD void unpack_linear(const SR &sr, uint32_t P, uint32_t ind, int val, uint32_t (&v)[ROWS])
{
	const uint32_t mask = (1u << ind) - 1u;
	const uint32_t bias = (uint32_t)(-(1 << (ind - 1)) * val);
	const uint32_t i = P >> 5, sh = P & 31u;
	unsigned long long win = (((unsigned long long)sr.word(i + 1) << 32) | sr.word(i)) >> sh;
	uint32_t avail = 64u - sh, nx = i + 3;
	uint32_t nextw = sr.word(i + 2);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int r = 0; r < ROWS; r++) {
		v[r] = ((uint32_t)win & mask) * (uint32_t)val + bias;
		win >>= ind;
		avail -= ind;
		if (avail <= 32u) {
			win |= (unsigned long long)nextw << avail;
			avail += 32u;
			nextw = sr.word(nx);
			nx++;
		}
	}
}

} // namespace fast2
} // namespace acm
