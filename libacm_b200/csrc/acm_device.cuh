/*
 * acm_device.cuh -- data structures and the per-phase decode logic shared by the
 * kernels.  Everything here is __host__ __device__ so that tests can run the very
 * same code on the CPU (tests/emu) where no GPU is available; the product only
 * ever calls it from the kernels in acm_kernels.cu.
 *
 * Coordinates: every stream is addressed through a 16-byte aligned base pointer
 * (so that 128-bit loads and TMA bulk copies can fetch it); bit positions ("P")
 * count from bit 0 of that base, so the first data bit (right after the 14/42-byte
 * header) sits at P = bit0 in {0,8,...,120}.
 * file_end is the P one past the last bit of the file image; the reference's
 * single zero byte at EOF (decode.c:57-61) makes limit = file_end + 8 the first
 * position that cannot be read.  A GET_BITS of n bits at P succeeds iff
 * P + n <= limit (decode.c:108-135).
 */
#pragma once

#include <stdint.h>

#include "acm_tables.h"

#if defined(__CUDACC__)
#define ACM_HD __host__ __device__ __forceinline__
#else
#define ACM_HD inline
#endif

namespace acm {

/* per-stream device descriptor, built on the host by acm_batch.cu */
struct DevStream {
	uint64_t base_off;    /* byte offset (multiple of 16) of the stream base inside the blob */
	uint64_t out_off;     /* byte offset of the PCM inside out */
	uint32_t bit0;        /* P of the first data bit */
	uint32_t file_end;    /* P one past the last file bit */
	uint32_t words_limit; /* words the reference's read loop delivers at most */
	uint32_t pad_words;   /* zero-fill [words, pad_words) afterwards (0 = no padding) */
	uint32_t n_attempt;   /* blocks the reference attempts to decode */
	uint32_t index;       /* slot in the result arrays (caller's order) */
	uint32_t rows;
	uint32_t level;
	uint32_t resume;      /* 1: continue a stream (history comes from KernelArgs::resume_hist) */
	uint32_t reserved;
};

struct Format {
	int wordlen;   /* 2, 3, 4 */
	int be;        /* big endian */
	uint32_t bias; /* 0 or the sign bit (unsigned formats, decode.c:640) */
	int checksums;
};

/* block scan verdicts */
enum { SCAN_OK = 1, SCAN_EOF = 0 /* clean EOF: acm_read returns 0 */ };

/* ------------------------------------------------------------------ bits */

struct BitReader {
	const uint32_t *base;
	uint32_t file_end;
	uint32_t widx, w0, w1;

	ACM_HD uint32_t word(uint32_t i) const
	{
		uint32_t last = file_end >> 5, tail = file_end & 31u;
		if (i < last)
#if defined(__CUDA_ARCH__)
			return __ldg(base + i);
#else
			return base[i];
#endif
		if (i == last && tail) {
#if defined(__CUDA_ARCH__)
			return __ldg(base + i) & ((1u << tail) - 1u);
#else
			return base[i] & ((1u << tail) - 1u);
#endif
		}
		return 0u; /* the zero byte of decode.c:57-61 and everything beyond */
	}
	ACM_HD void init(const uint32_t *b, uint32_t fe)
	{
		base = b;
		file_end = fe;
		widx = 0xFFFFFFF0u;
		w0 = w1 = 0;
	}
	/* 32 bits starting at P (bits past file_end read as zero) */
	ACM_HD uint32_t peek(uint32_t P)
	{
		uint32_t i = P >> 5, s = P & 31u;
		if (i != widx) {
			if (i == widx + 1) {
				w0 = w1;
				w1 = word(i + 1);
			} else {
				w0 = word(i);
				w1 = word(i + 1);
			}
			widx = i;
		}
		return s ? (w0 >> s) | (w1 << (32 - s)) : w0;
	}
};

/* ------------------------------------------------------------------ columns */

ACM_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }

/* sign-extend the 4-bit field j of x */
ACM_HD int nib_s(uint32_t x, int j) { return ((int32_t)(x << (28 - 4 * j))) >> 28; }

/*
 * Length scan of one column payload: returns the P just past it.  Mirrors the bit
 * consumption of the reference fillers (decode.c:181-476) without producing values.
 */
template <typename BR>
ACM_HD uint32_t scan_column(BR &br, uint32_t P, uint32_t ind, uint32_t kind, uint32_t rows,
			    const uint64_t *k8)
{
	uint32_t cls = kind & 7u, sub = kind >> 3;
	if (cls == ACM_CLS_ZERO)
		return P;
	if (cls == ACM_CLS_LINEAR)
		return P + rows * ind; /* f_linear decode.c:196-206 */
	if (cls == ACM_CLS_T) {        /* f_t15/t27/t37: fixed-width codes */
		uint32_t per = sub == 2 ? 2u : 3u, width = sub == 0 ? 5u : 7u;
		return P + ((rows + per - 1) / per) * width;
	}
	/* k-codes: walk whole symbols, up to 7 values per 8-bit window */
	const uint64_t *tab = k8 + sub * 256;
	uint32_t rem = rows;
	while (rem) {
		uint32_t e = (uint32_t)tab[br.peek(P) & 255u];
		uint32_t k = umin32(e & 15u, rem);
		P += (e >> (4 * k)) & 15u;
		rem -= k;
	}
	return P;
}

/*
 * Decode the payload of one column into dst[r * stride], r < rows, as idx * val
 * (set_pos + midbuf, decode.c:174-177, :591-600).  `avail` = limit - P0 bounds which
 * t-codes may be inspected: a code is range-checked only if it lies entirely before
 * the limit, exactly the codes the reference gets to read.  Returns 0, or
 * ACM_ERR_CORRUPT (-6) if such a code is out of range.
 */
template <typename T, typename BR>
ACM_HD int decode_column(BR &br, uint32_t P, uint32_t limit, uint32_t ind, uint32_t kind,
			 uint32_t rows, int val, T *dst, uint32_t stride, const uint64_t *k8,
			 const uint16_t *tt)
{
	uint32_t cls = kind & 7u, sub = kind >> 3, r = 0;
	if (cls == ACM_CLS_ZERO) {
		for (; r < rows; r++)
			dst[(size_t)r * stride] = 0;
		return 0;
	}
	if (cls == ACM_CLS_LINEAR) {
		int mid = 1 << (ind - 1);
		uint32_t mask = (1u << ind) - 1u;
		for (; r < rows; r++, P += ind)
			dst[(size_t)r * stride] = (T)(((int)(br.peek(P) & mask) - mid) * val);
		return 0;
	}
	if (cls == ACM_CLS_T) {
		uint32_t per = sub == 2 ? 2u : 3u, width = sub == 0 ? 5u : 7u;
		const uint16_t *tab = tt + sub * 128;
		int bad = 0;
		while (r < rows) {
			uint32_t e = tab[br.peek(P) & ((1u << width) - 1u)];
			if (P + width <= limit && (e & 0x8000u))
				bad = 1;
			P += width;
			for (uint32_t j = 0; j < per && r < rows; j++, r++)
				dst[(size_t)r * stride] = (T)(nib_s(e, j) * val);
		}
		return bad ? -6 : 0;
	}
	const uint64_t *tab = k8 + sub * 256;
	while (r < rows) {
		uint64_t e64 = tab[br.peek(P) & 255u];
		uint32_t e = (uint32_t)e64, hi = (uint32_t)(e64 >> 32);
		uint32_t k = umin32(e & 15u, rows - r);
		P += (e >> (4 * k)) & 15u;
		for (uint32_t j = 0; j < k; j++, r++)
			dst[(size_t)r * stride] = (T)(nib_s(hi, j) * val);
	}
	return 0;
}

/* ------------------------------------------------------------------ block scan */

struct ScanResult {
	int status;       /* SCAN_OK, SCAN_EOF, or ACM_ERR_* */
	int val;          /* block multiplier */
	uint32_t ncols;   /* columns whose payload was scanned completely */
	uint32_t end;     /* P after the block (valid when status == SCAN_OK) */
};

/*
 * Serial walk over one block: header (decode.c:588-589), then for every column the
 * 5-bit selector (decode.c:496) and the payload length.  coloff[c] receives the P of
 * column c's selector.  Stops at the first read the reference could not perform:
 *   - pwr / val / selector past the limit  -> SCAN_EOF   (GET_BITS_EXPECT_EOF)
 *   - bad selector                          -> ACM_ERR_CORRUPT (f_bad decode.c:190)
 *   - payload past the limit                -> ACM_ERR_UNEXPECTED_EOF, unless a t-code
 *     that still fits is out of range       -> ACM_ERR_CORRUPT (checked by the caller's
 *     decode of column ncols, which this function leaves to decode_column)
 */
template <typename OFF, typename BR>
ACM_HD ScanResult scan_block(BR &br, uint32_t P, uint32_t limit, uint32_t cols,
			     uint32_t rows, OFF *coloff, uint32_t off_base, const uint8_t *kind,
			     const uint64_t *k8, uint32_t pitch = 1)
{
	ScanResult s;
	s.status = SCAN_OK;
	s.ncols = 0;
	s.val = 0;
	s.end = P;
	if (P + 20 > limit) {
		/* pwr (4) or val (16) cannot be read */
		s.status = SCAN_EOF;
		return s;
	}
	s.val = (int)((br.peek(P) >> 4) & 0xFFFFu);
	P += 20;
	for (uint32_t c = 0; c < cols; c++) {
		if (P + 5 > limit) {
			s.status = SCAN_EOF;
			break;
		}
		uint32_t ind = br.peek(P) & 31u, k = kind[ind];
		coloff[c * pitch] = (OFF)(P - off_base);
		if ((k & 7u) == ACM_CLS_BAD) {
			s.status = -6;
			break;
		}
		P = scan_column(br, P + 5, ind, k, rows, k8);
		if (P > limit) {
			s.status = -7;
			break;
		}
		s.ncols = c + 1;
	}
	s.end = P;
	return s;
}

/* ------------------------------------------------------------------ transform */

/*
 * One output of lifting stage with column count C (flat form, SURVEY.md B.3;
 * reference juggle decode.c:508-526): cur = stage input, h = last 2C inputs of the
 * previous block for this stage.
 */
ACM_HD uint32_t juggle_at(const uint32_t *cur, const uint32_t *h, uint32_t m, uint32_t C)
{
	uint32_t a = cur[m];
	uint32_t p1 = m >= C ? cur[m - C] : h[C + m];
	uint32_t p2 = m >= 2 * C ? cur[m - 2 * C] : h[m];
	uint32_t s = a + p2;
	return ((m / C) & 1u) ? 2u * p1 - s : 2u * p1 + s;
}

/* ------------------------------------------------------------------ output */

/* output_values (decode.c:617-677) generalised to wordlen 2..4; returns the word as
 * an unsigned integer (checksum input) */
ACM_HD uint32_t emit_word(uint8_t *dst, int32_t v, const Format &f)
{
	uint32_t u = (uint32_t)v + f.bias;
	if (f.wordlen == 2) {
		u &= 0xFFFFu;
		uint16_t w = f.be ? (uint16_t)((u >> 8) | (u << 8)) : (uint16_t)u;
		*(uint16_t *)dst = w; /* out_off is 16-byte aligned, so words are aligned */
	} else if (f.wordlen == 4) {
		uint32_t w = u;
		if (f.be)
			w = (u >> 24) | ((u >> 8) & 0xFF00u) | ((u << 8) & 0xFF0000u) | (u << 24);
		*(uint32_t *)dst = w;
	} else {
		u &= 0xFFFFFFu;
		if (f.be) {
			dst[0] = (uint8_t)(u >> 16); dst[1] = (uint8_t)(u >> 8); dst[2] = (uint8_t)u;
		} else {
			dst[0] = (uint8_t)u; dst[1] = (uint8_t)(u >> 8); dst[2] = (uint8_t)(u >> 16);
		}
	}
	return u;
}

} // namespace acm
