/*
 * tests/emu/emu_fast2.cpp -- CPU run of the lane-local code of the level-7 / 16-row kernel
 * (TEST ONLY; never part of the product).
 *
 * libacm_b200/csrc/acm_fast2_core.cuh (the uni16 column walk and the three column unpackers)
 * is __host__ __device__: here the very same functions walk and unpack every block of a stream
 * and are compared, column by column, with the generic building blocks (scan_block /
 * decode_column of acm_device.cuh), which tests/test_emu.py pins against the oracle.
 */
#include <cstdlib>
#include <cstring>
#include <vector>

#include "acm_fast2_core.cuh"
#include "acm_host.h"
#include "libacm.h"

using namespace acm;

namespace {
struct WordReader {
	BitReader *br;
	uint32_t word(uint32_t i) const { return br->word(i); }
};
} // namespace

/* returns the number of blocks checked, or -(1000*block + code) at the first difference */
extern "C" long emu_fast2_check(const uint8_t *img, uint32_t len)
{
	acm_header h;
	static acm_tables tab;
	static bool built = false;
	if (!built) {
		acm_tables_build(&tab);
		built = true;
	}
	if (acm_parse_header(img, len, 0, &h) < 0 || h.level != 7 || h.rows != 16)
		return -1;
	std::vector<uint32_t> words((len + 64) / 4 + 4, 0xA5A5A5A5u); /* garbage after the image */
	memcpy(words.data(), img, len);
	BitReader br;
	const uint32_t file_end = len * 8u, limit = file_end + 8u;
	br.init(words.data(), file_end);
	uint32_t P = h.header_len * 8u;
	const uint32_t nblocks = (h.total_values + 2047u) / 2048u;
	long checked = 0;
	for (uint32_t b = 0; b < nblocks; b++, checked++) {
		uint16_t off[128];
		const ScanResult sc = scan_block(br, P, limit, 128u, 16u, off, P, tab.kind, tab.k8);
		/* ---- the walk, exactly as fast_step drives it */
		fast2::Walk s;
		s.Q = P + 20u - 1u;
		s.s8 = 0u;
		s.msk = fast2::MSK_SEL;
		uint32_t col = 0, guard = 0;
		uint16_t woff[129];
		bool ok_walk = P + 20u <= limit;
		while (ok_walk && s.s8 != fast2::UNI_HALT8 && s.s8 != fast2::UNI_BAD8 && guard++ < 100000u) {
			const uint32_t w = br.peek(s.Q);
			if (s.msk == fast2::MSK_SEL)
				woff[col++] = (uint16_t)(s.Q + 1u - P);
			const uint32_t e = *reinterpret_cast<const uint16_t *>(
				reinterpret_cast<const unsigned char *>(tab.uni16) + fast2::walk_index(s, w));
			const bool at_sel = fast2::walk_next(s, e);
			if (at_sel && col == 128u) {
				s.s8 = fast2::UNI_HALT8;
				s.msk = fast2::MSK_K;
			}
		}
		const bool fast_ok = ok_walk && s.s8 == fast2::UNI_HALT8 && s.Q + 1u <= limit;
		if (fast_ok != (sc.status == SCAN_OK))
			return -(1000L * b + 1);
		if (!fast_ok)
			break; /* the kernel re-walks this block with scan_block itself */
		if (s.Q + 1u != sc.end)
			return -(1000L * b + 2);
		for (uint32_t c = 0; c < 128; c++)
			if (woff[c] != off[c])
				return -(1000L * b + 3);
		/* ---- the unpackers, fed like the kernel feeds them: the 96 stream bits at the payload */
		WordReader sr{ &br };
		for (uint32_t c = 0; c < 128; c++) {
			const uint32_t Pc = P + off[c], Pp = Pc + 5u;
			const uint32_t ind = br.peek(Pc) & 31u, kind = tab.kind[ind], cls = kind & 7u, sub = kind >> 3;
			int ref[16];
			uint32_t got[16];
			const int val = 1 + (int)((b * 131u + c * 7u) % 65535u);
			const int rc = decode_column(br, Pp, limit, ind, kind, 16u, val, ref, 1u, tab.k8, tab.t);
			const uint32_t lo = br.peek(Pp), mid = br.peek(Pp + 32u), hi = br.peek(Pp + 64u);
			uint32_t a0 = 0u, a1 = 0u;
			int bad = 0;
			if (cls == ACM_CLS_K)
				fast2::unpack_k(lo, mid, hi, sub, tab.k8w, a0, a1);
			else if (cls == ACM_CLS_T)
				bad = fast2::unpack_t(lo, mid, Pp, limit, sub, tab.t, a0, a1);
			if (cls == ACM_CLS_LINEAR)
				fast2::unpack_linear(sr, Pp, ind, val, got);
			else
				for (int r = 0; r < 16; r++)
					got[r] = fast2::nib_val(r < 8 ? a0 : a1, r & 7, val);
			if ((rc == -6) != (bad != 0))
				return -(1000L * b + 4);
			if (!bad)
				for (int r = 0; r < 16; r++)
					if (got[r] != (uint32_t)ref[r])
						return -(1000L * b + 5);
		}
		/* ---- and four columns at once, as a decode lane holds them (columns j, j+32, j+64, j+96) */
		for (uint32_t j = 0; j < 32; j++) {
			uint32_t lo[4], mid[4], hi[4], cls[4], sub[4], Pp[4], a0[4] = { 0, 0, 0, 0 }, a1[4] = { 0, 0, 0, 0 };
			int bad_ref = 0;
			for (int p = 0; p < 4; p++) {
				const uint32_t Pc = P + off[j + 32 * p];
				const uint32_t ind = br.peek(Pc) & 31u, kind = tab.kind[ind];
				Pp[p] = Pc + 5u;
				cls[p] = kind & 7u;
				sub[p] = kind >> 3;
				lo[p] = br.peek(Pp[p]);
				mid[p] = br.peek(Pp[p] + 32u);
				hi[p] = br.peek(Pp[p] + 64u);
			}
			const int bad4 = fast2::unpack_t4(lo, mid, Pp, limit, cls, sub, tab.t, a0, a1);
			fast2::unpack_k4(lo, mid, hi, cls, sub, tab.k8w, a0, a1);
			for (int p = 0; p < 4; p++) {
				if (cls[p] == ACM_CLS_LINEAR)
					continue;
				int ref[16];
				const uint32_t Pc = P + off[j + 32 * p], ind = br.peek(Pc) & 31u;
				const int rc = decode_column(br, Pc + 5u, limit, ind, tab.kind[ind], 16u, 3, ref, 1u, tab.k8, tab.t);
				bad_ref |= rc == -6;
				if (rc != -6)
					for (int r = 0; r < 16; r++)
						if (fast2::nib_val(r < 8 ? a0[p] : a1[p], r & 7, 3) != (uint32_t)ref[r])
							return -(1000L * b + 6);
			}
			if ((bad_ref != 0) != (bad4 != 0))
				return -(1000L * b + 7);
		}
		P = sc.end;
	}
	return checked;
}

/* walk statistics of a level-7 / 16-row image (tuning aid): out[0] blocks, [1] walk steps,
 * [2] steps that advance more than 31 bits, [3] steps the walk would take if no step could
 * advance more than 31 bits, [4] columns, [5] selector steps by class: zero, [6] linear, [7] k,
 * [8] t, [9] total bits */
extern "C" long emu_fast2_stats(const uint8_t *img, uint32_t len, uint64_t *out)
{
	acm_header h;
	static acm_tables tab;
	static bool built = false;
	if (!built) {
		acm_tables_build(&tab);
		built = true;
	}
	if (acm_parse_header(img, len, 0, &h) < 0 || h.level != 7 || h.rows != 16)
		return -1;
	std::vector<uint32_t> words((len + 64) / 4 + 4, 0u);
	memcpy(words.data(), img, len);
	BitReader br;
	const uint32_t file_end = len * 8u, limit = file_end + 8u;
	br.init(words.data(), file_end);
	uint32_t P = h.header_len * 8u;
	const uint32_t nblocks = (h.total_values + 2047u) / 2048u;
	for (uint32_t b = 0; b < nblocks; b++) {
		if (P + 20u > limit)
			break;
		fast2::Walk s;
		s.Q = P + 20u - 1u;
		s.s8 = 0u;
		s.msk = fast2::MSK_SEL;
		uint32_t col = 0, guard = 0;
		while (s.s8 != fast2::UNI_HALT8 && s.s8 != fast2::UNI_BAD8 && guard++ < 100000u) {
			const uint32_t w = br.peek(s.Q);
			const bool sel = s.msk == fast2::MSK_SEL;
			if (sel) {
				col++;
				const uint32_t ind = (w >> 1) & 31u, cls = tab.kind[ind] & 7u;
				out[4]++;
				out[cls == ACM_CLS_LINEAR ? 6 : cls == ACM_CLS_K ? 7 : cls == ACM_CLS_T ? 8 : 5]++;
			}
			const uint32_t e = *reinterpret_cast<const uint16_t *>(
				reinterpret_cast<const unsigned char *>(tab.uni16) + fast2::walk_index(s, w));
			const uint32_t q0 = s.Q;
			const bool at_sel = fast2::walk_next(s, e);
			const uint32_t adv = s.Q - q0;
			out[1]++;
			out[2] += adv > 31u;
			out[3] += adv > 31u ? (adv + 30u) / 31u : 1u;
			if (at_sel && col == 128u) {
				s.s8 = fast2::UNI_HALT8;
				s.msk = fast2::MSK_K;
			}
		}
		if (s.s8 != fast2::UNI_HALT8 || s.Q + 1u > limit)
			break;
		out[9] += s.Q + 1u - P;
		P = s.Q + 1u;
		out[0]++;
	}
	return 0;
}
