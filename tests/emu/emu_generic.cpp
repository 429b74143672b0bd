/*
 * tests/emu/emu_generic.cpp -- CPU emulation of the generic decode kernel's control
 * flow (TEST ONLY; never part of the product).
 *
 * The __host__ __device__ building blocks of libacm_b200/csrc/acm_device.cuh
 * (BitReader, scan_block, decode_column, juggle_at, emit_word), the code tables and
 * the host descriptor logic (acm_parse_header, acm_make_devstream) are compiled here
 * with g++ and driven with NT sequential "threads" per phase, phase boundaries
 * standing in for __syncthreads().  This lets the CPU-only test run the same logic
 * the GPU runs and compare it with the oracle before any GPU time is spent.  It is
 * deliberately a line-for-line mirror of acm_decode_generic_kernel.
 */
#include <cstdlib>
#include <cstring>
#include <vector>

#include "acm_device.cuh"
#include "acm_gpu.h"
#include "acm_host.h"
#include "libacm.h"

using namespace acm;

extern "C" int emu_decode(const uint8_t *blob, uint64_t blob_len, uint64_t in_off, uint32_t in_len,
			  int force_chans, int be, int wordlen, int sgned, int pad_tail, int nthreads,
			  uint8_t *out, uint32_t *words_out, int *status_out, uint64_t *cks_out,
			  uint32_t *total_values_out)
{
	acm_header h;
	acm_gpu_stream g;
	DevStream d;
	acm_tables tab;
	Format fmt;
	const int NT = nthreads > 0 ? nthreads : 256;

	*words_out = 0;
	*status_out = 0;
	*cks_out = 0;
	*total_values_out = 0;
	int err = acm_parse_header(blob + in_off, in_len, force_chans, &h);
	if (err < 0)
		return err;
	memset(&g, 0, sizeof(g));
	g.in_off = in_off;
	g.in_len = in_len;
	g.total_values = h.total_values;
	g.channels = h.channels;
	g.acm_channels = h.acm_channels;
	g.rate = h.rate;
	g.level = h.level;
	g.rows = h.rows;
	g.wavc = h.wavc;
	*total_values_out = h.total_values;
	err = acm_make_devstream(&g, 0, pad_tail, &d);
	if (err < 0)
		return err;
	acm_tables_build(&tab);
	fmt.wordlen = wordlen;
	fmt.be = be;
	fmt.bias = sgned ? 0u : (1u << (8 * wordlen - 1));
	fmt.checksums = 1;

	const uint32_t level = d.level, cols = 1u << level, rows = d.rows;
	const uint32_t blen = rows * cols, limit = d.file_end + 8u;
	std::vector<uint32_t> buf0(blen), buf1(blen), hist(2 * cols, 0u), coloff(cols);
	std::vector<BitReader> brs(NT);
	for (int t = 0; t < NT; t++)
		brs[t].init((const uint32_t *)(blob + d.base_off), d.file_end);
	uint32_t *cur = buf0.data(), *nxt = buf1.data();
	uint32_t P = d.bit0, pos = 0;
	int st = 0;
	unsigned long long cks = 0;

	for (uint32_t b = 0; b < d.n_attempt; b++) {
		ScanResult sc = scan_block(brs[0], P, limit, cols, rows, coloff.data(), 0u, tab.kind, tab.k8);
		int bad = 0;
		const uint32_t ncheck = sc.ncols + (sc.status == -7 ? 1u : 0u);
		for (int t = 0; t < NT; t++)
			for (uint32_t c = t; c < ncheck; c += NT) {
				uint32_t Pc = coloff[c];
				uint32_t ind = brs[t].peek(Pc) & 31u;
				int r = decode_column(brs[t], Pc + 5u, limit, ind, tab.kind[ind], rows, sc.val,
						      cur + c, cols, tab.k8, tab.t);
				if (r < 0)
					bad = 1;
			}
		if (bad)
			st = -6;
		else if (sc.status != SCAN_OK)
			st = sc.status == SCAN_EOF ? 0 : sc.status;
		if (bad || sc.status != SCAN_OK)
			break;
		P = sc.end;
		uint32_t hoff = 0;
		for (uint32_t l = 1; l <= level; l++) {
			const uint32_t C = cols >> l;
			uint32_t *hh = hist.data() + hoff;
			for (uint32_t m = 0; m < blen; m++) {
				uint32_t v = juggle_at(cur, hh, m, C);
				if (l == 1 && (m & (C - 1u)) == 0u)
					v += 1u;
				nxt[m] = v;
			}
			for (uint32_t i = 0; i < 2 * C; i++)
				hh[i] = cur[blen - 2 * C + i];
			uint32_t *t = cur; cur = nxt; nxt = t;
			hoff += 2 * C;
		}
		uint32_t n = blen;
		if (n > d.words_limit - pos)
			n = d.words_limit - pos;
		for (uint32_t m = 0; m < n; m++) {
			uint32_t u = emit_word(out + (size_t)(pos + m) * wordlen, (int32_t)cur[m] >> level, fmt);
			cks += (unsigned long long)(pos + m + 1u) * (unsigned long long)(u + 1ull);
		}
		pos += n;
	}
	if (d.pad_words > pos)
		memset(out + (size_t)pos * wordlen, 0, (size_t)(d.pad_words - pos) * wordlen);
	*words_out = pos;
	*status_out = st;
	*cks_out = cks;
	return 0;
}
