/*
 * acm_gen2.cu -- the GENERAL decode path: any level (cols = 1 << level), any row count, any
 * output format, as three kernels around a table of block records:
 *
 *   scan      acm_scan_kernel: one warp per stream, its first lane walks the stream block by
 *             block (the walk is serial: where column c+1 starts is only known once column c has
 *             been walked) and leaves, per block, a BlockRec (bit position, val,
 *             verdict) and the position of every column selector.  This is fill_block's control
 *             flow (decode.c:491-502) without its data flow.
 *   decode    acm_blocks_kernel: with the records in place, BLOCKS are independent units of
 *             work except for the transform's history -- and the history a block needs is a
 *             function of the 2*cols-2 words before it, i.e. of the
 *             previous block alone (of the previous two when rows == 1).  A work item is a run
 *             of consecutive blocks of one stream plus that warm-up; a CTA unpacks (one column
 *             per thread: filler dispatch, code tables, idx*val), runs the `level` lifting stages
 *             as flat 3-tap stencils and writes the PCM of its run.  Long streams are decoded by
 *             many CTAs at once; nothing waits for a scan thread any more.
 *   finalise  acm_finish_kernel: per stream, what the reference's read loop would report --
 *             words delivered up to the first block that failed (scan verdict, or an out-of-range
 *             radix code found by a decode item), status, checksum (sum of the per-block ones) --
 *             and the zero padding of the undelivered tail (acmtool.c:293-310).
 *
 * The level-7 / 16-row shape has its own fused kernel (acm_fast2.cu); everything else, and the
 * 24/32-bit formats, come here.
 */
#include "acm_kernels.cuh"

namespace acm {

namespace {

constexpr int G2_THREADS = 128;   /* decode CTA */
constexpr int G2_SCAN_WARPS = 4;  /* scan CTA: one stream per warp */

struct TablesSmem {
	uint64_t k8[ACM_K8_SIZE];
	uint16_t t[ACM_T_SIZE];
	uint8_t kind[32];
};

__device__ __forceinline__ void load_tables(TablesSmem &s, const acm_tables *g, int tid, int nt)
{
	for (int i = tid; i < ACM_K8_SIZE; i += nt)
		s.k8[i] = g->k8[i];
	for (int i = tid; i < ACM_T_SIZE; i += nt)
		s.t[i] = g->t[i];
	if (tid < 32)
		s.kind[tid] = g->kind[tid];
}

/* ------------------------------------------------------------------ scan */

__global__ void __launch_bounds__(32 * G2_SCAN_WARPS) acm_scan_kernel(KernelArgs a, Gen2Args g)
{
	__shared__ TablesSmem tab;
	load_tables(tab, a.tables, threadIdx.x, 32 * G2_SCAN_WARPS);
	__syncthreads();
	const uint32_t si = blockIdx.x * G2_SCAN_WARPS + (threadIdx.x >> 5);
	if (si >= a.count || (threadIdx.x & 31) != 0)
		return;
	const DevStream d = a.streams[si];
	const Gen2Stream gs = g.gs[si];
	const uint32_t cols = 1u << d.level, limit = d.file_end + 8u;
	BitReader br;
	br.init((const uint32_t *)(a.blob + d.base_off), d.file_end);
	uint32_t P = d.bit0, b = 0;
	const uint32_t nmax = d.n_attempt < gs.max_blocks ? d.n_attempt : gs.max_blocks;
	for (; b < nmax; b++) {
		uint32_t *coff = g.coff + gs.coff_base + (size_t)b * cols;
		const ScanResult sc = scan_block(br, P, limit, cols, d.rows, coff, 0u, tab.kind, tab.k8);
		BlockRec r;
		r.P = P;
		r.end = sc.end;
		r.val = sc.val;
		r.status = sc.status;
		r.ncols = sc.ncols;
		r.pad0 = r.pad1 = r.pad2 = 0u;
		g.rec[gs.rec_base + b] = r;
		if (sc.status != SCAN_OK) {
			b++;
			break; /* the stream ends with this block */
		}
		P = sc.end;
	}
	g.nscan[si] = b;
}

/* ------------------------------------------------------------------ decode */

__global__ void __launch_bounds__(G2_THREADS) acm_blocks_kernel(KernelArgs a, Gen2Args g, GenericScratch scr)
{
	__shared__ TablesSmem tab;
	__shared__ uint32_t s_item;
	__shared__ int s_bad;
	__shared__ unsigned long long s_cks;
	const int tid = threadIdx.x;
	load_tables(tab, a.tables, tid, G2_THREADS);
	__syncthreads();

	uint32_t *buf0 = scr.buf + (size_t)blockIdx.x * scr.stride;
	uint32_t *buf1 = buf0 + scr.max_blen;
	uint32_t *hist = buf1 + scr.max_blen; /* 2 * max_cols words */

	for (;;) {
		if (tid == 0)
			s_item = atomicAdd(g.item_counter, 1u);
		__syncthreads();
		const uint32_t it = s_item;
		if (it >= g.n_items)
			break;
		const Gen2Item item = g.items[it];
		const uint32_t si = item.stream;
		const DevStream d = a.streams[si];
		const Gen2Stream gs = g.gs[si];
		const uint32_t level = d.level, cols = 1u << level, rows = d.rows;
		const uint32_t blen = rows * cols, limit = d.file_end + 8u;
		const uint32_t nscan = g.nscan[si];
		BitReader br;
		br.init((const uint32_t *)(a.blob + d.base_off), d.file_end);
		uint8_t *out = a.out + d.out_off;
		const uint32_t bfirst = item.b0 - item.warm, bend = item.b0 + item.nb;

		/* zeroed history (decode.c:812) for a stream's first block; for a run that starts inside the
		 * stream the warm-up blocks rebuild it */
		for (uint32_t i = tid; i < 2 * cols; i += G2_THREADS)
			hist[i] = 0u;
		uint32_t *cur = buf0, *nxt = buf1;
		__syncthreads();

		for (uint32_t b = bfirst; b < bend && b < nscan; b++) {
			const BlockRec rec = g.rec[gs.rec_base + b];
			const uint32_t *coff = g.coff + gs.coff_base + (size_t)b * cols;
			const bool ok = rec.status == SCAN_OK;
			if (tid == 0) {
				s_bad = 0;
				s_cks = 0ull;
			}
			__syncthreads();
			/* ---- unpack (column rec.ncols is included when its payload ran past the limit: a
			 * radix code that still fits may be out of range first, decode.c:412/:438/:464) */
			const uint32_t ncheck = rec.ncols + (rec.status == -7 ? 1u : 0u);
			for (uint32_t c = tid; c < ncheck; c += G2_THREADS) {
				const uint32_t Pc = coff[c];
				const uint32_t ind = br.peek(Pc) & 31u;
				const int r = decode_column(br, Pc + 5u, limit, ind, tab.kind[ind], rows, rec.val, cur + c, cols,
							    tab.k8, tab.t);
				if (r < 0)
					s_bad = 1;
			}
			__syncthreads();
			if (s_bad && b >= item.b0 && tid == 0)
				atomicMin(g.first_bad + si, b); /* warm-up blocks are reported by their own items */
			if ((s_bad && b >= item.b0) || !ok)
				break; /* uniform: decided from shared state */

			/* ---- juggle (decode.c:528-577 in flat form) */
			uint32_t hoff = 0;
			for (uint32_t l = 1; l <= level; l++) {
				const uint32_t C = cols >> l;
				uint32_t *h = hist + hoff;
				for (uint32_t m = tid; m < blen; m += G2_THREADS) {
					uint32_t v = juggle_at(cur, h, m, C);
					if (l == 1 && (m & (C - 1u)) == 0u)
						v += 1u; /* decode.c:561-564 */
					nxt[m] = v;
				}
				__syncthreads();
				for (uint32_t i = tid; i < 2 * C; i += G2_THREADS)
					h[i] = cur[blen - 2 * C + i];
				__syncthreads();
				uint32_t *t = cur; cur = nxt; nxt = t;
				hoff += 2 * C;
			}

			/* ---- output (decode.c:849-866) */
			if (b >= item.b0) {
				const uint64_t pos = (uint64_t)b * blen;
				uint32_t n = pos < d.words_limit ? (uint32_t)(d.words_limit - pos < blen ? d.words_limit - pos : blen) : 0u;
				unsigned long long cks = 0ull;
				for (uint32_t m = tid; m < n; m += G2_THREADS) {
					const uint32_t u = emit_word(out + (size_t)(pos + m) * a.fmt.wordlen, (int32_t)cur[m] >> level, a.fmt);
					if (a.fmt.checksums)
						cks += (unsigned long long)(pos + m + 1u) * (unsigned long long)(u + 1ull);
				}
				if (a.fmt.checksums) {
					for (int o = 16; o; o >>= 1)
						cks += __shfl_xor_sync(0xFFFFFFFFu, cks, o);
					if ((tid & 31) == 0)
						atomicAdd(&s_cks, cks);
					__syncthreads();
					if (tid == 0)
						g.cks_blk[gs.rec_base + b] = s_cks;
				}
			}
			__syncthreads();
		}
		__syncthreads();
	}
}

/* ------------------------------------------------------------------ finalise */

__global__ void __launch_bounds__(G2_THREADS) acm_finish_kernel(KernelArgs a, Gen2Args g)
{
	__shared__ unsigned long long s_cks;
	const int tid = threadIdx.x;
	for (uint32_t si = blockIdx.x; si < a.count; si += gridDim.x) {
		const DevStream d = a.streams[si];
		const Gen2Stream gs = g.gs[si];
		const uint32_t blen = d.rows << d.level;
		const uint32_t ns = g.nscan[si], fb = g.first_bad[si];
		int last = SCAN_OK;
		if (ns)
			last = g.rec[gs.rec_base + ns - 1].status;
		const uint32_t nok_scan = ns ? (last == SCAN_OK ? ns : ns - 1u) : 0u;
		uint32_t nok = nok_scan;
		int st = last == SCAN_OK || last == SCAN_EOF ? 0 : last;
		if (fb <= nok_scan) { /* a corrupt radix code comes first (fb = 0xFFFFFFFF: none) */
			nok = fb;
			st = -6;
		}
		const uint64_t w64 = (uint64_t)nok * blen;
		const uint32_t words = w64 < d.words_limit ? (uint32_t)w64 : d.words_limit;
		if (tid == 0)
			s_cks = 0ull;
		__syncthreads();
		if (a.fmt.checksums) {
			unsigned long long c = 0ull;
			for (uint32_t b = tid; b < nok; b += G2_THREADS)
				if ((uint64_t)b * blen < d.words_limit)
					c += g.cks_blk[gs.rec_base + b];
			for (int o = 16; o; o >>= 1)
				c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
			if ((tid & 31) == 0)
				atomicAdd(&s_cks, c);
		}
		/* zero padding of the undelivered tail, up to the 16-byte boundary that ends this stream's
		 * slot: no stale bytes in the gaps (and nothing of what later runs may have written past
		 * the first failing block survives) */
		{
			uint8_t *p = a.out + d.out_off + (size_t)words * a.fmt.wordlen;
			const size_t nbytes = d.pad_words >= words && d.pad_words
						      ? (((size_t)d.pad_words * a.fmt.wordlen + 15u) & ~(size_t)15u) -
								(size_t)words * a.fmt.wordlen
						      : 0;
			for (size_t i = tid; i < nbytes; i += G2_THREADS)
				p[i] = 0;
		}
		__syncthreads();
		if (tid == 0) {
			a.status[d.index] = st;
			a.words[d.index] = words;
			a.cks[d.index] = a.fmt.checksums ? s_cks : 0ull;
		}
		__syncthreads();
	}
}

} // namespace

int gen2_ctas_per_sm()
{
	static int cached = 0;
	if (!cached) {
		int nb = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, acm_blocks_kernel, G2_THREADS, 0) != cudaSuccess || nb < 1)
			nb = 4;
		cached = nb > 8 ? 8 : nb;
	}
	return cached;
}

cudaError_t launch_gen2(const KernelArgs &a, const Gen2Args &g, const GenericScratch &s, int n_ctas, cudaStream_t st)
{
	if (a.count == 0)
		return cudaSuccess;
	const unsigned scan_grid = (a.count + G2_SCAN_WARPS - 1) / G2_SCAN_WARPS;
	acm_scan_kernel<<<scan_grid, 32 * G2_SCAN_WARPS, 0, st>>>(a, g);
	if (g.n_items)
		acm_blocks_kernel<<<n_ctas, G2_THREADS, 0, st>>>(a, g, s);
	return launch_gen2_finish(a, g, st);
}

cudaError_t launch_gen2_finish(const KernelArgs &a, const Gen2Args &g, cudaStream_t st)
{
	if (a.count == 0)
		return cudaSuccess;
	unsigned fin = a.count < 4096u ? a.count : 4096u;
	acm_finish_kernel<<<fin, G2_THREADS, 0, st>>>(a, g);
	return cudaGetLastError();
}

} // namespace acm
