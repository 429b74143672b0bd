/*
 * acm_kernels.cu -- the GENERIC decode kernel: any level (cols = 1 << level), any
 * row count, any stream.  One CTA works on one stream at a time (persistent grid,
 * atomic work queue), keeps the int32 block and the per-stage history in a per-CTA
 * global scratch area, and runs the three stages of the reference's decode_block
 * (decode.c:580-611) plus output_values (decode.c:657-677) per block:
 *
 *   1. scan    one thread walks the block serially and records where every column
 *              starts (column boundaries are data dependent: SURVEY.md H1).  It is the
 *              first lane of a ninth warp and runs one block AHEAD of the other eight
 *              warps (two column-offset buffers), so a block costs max(scan, decode),
 *              not their sum
 *   2. unpack  all threads: one column each -- filler dispatch, table-driven code
 *              decode, dequantisation idx*val, scattered store block[row*cols+col]
 *   3. juggle  level stages, each a flat 3-tap stencil over the block (Appendix B.3),
 *              ping-ponging between two scratch buffers; history = last 2C inputs
 *   4. output  shift, bias, byte order, store; optional checksum
 *
 * It is the correctness backstop and the path for unusual shapes; the common shape
 * (level 7, 16 rows) is served by the scan-CTA / decode-CTA kernel in acm_fast2.cu.
 */
#include "acm_kernels.cuh"

namespace acm {

/* A block costs max(scan, decode), and the scan -- ONE thread, ~400 dependent table steps per
 * 2048-word block -- is by far the longer of the two: what this kernel needs is many scan threads
 * in flight, i.e. many small CTAs per SM, not many decode threads per CTA. */
#ifndef GEN_THREADS
#define GEN_THREADS 128             /* decode threads (warps 0..3) */
#endif
#define GEN_BLOCK (GEN_THREADS + 32) /* + the scan warp (only its first lane works) */

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

/* barrier of the decode threads only (the scan thread is busy with the next block) */
__device__ __forceinline__ void decode_sync() { asm volatile("bar.sync 1, %0;" ::"n"(GEN_THREADS) : "memory"); }

__global__ void __launch_bounds__(GEN_BLOCK)
acm_decode_generic_kernel(KernelArgs a, GenericScratch scr)
{
	__shared__ TablesSmem tab;
	__shared__ int s_si;
	/* scan -> decode hand-over, all through shared-memory atomics (ordered by fences; racecheck
	 * understands atomics and barriers, not flags): the scan thread publishes a block's verdict in
	 * s_hand[b & 1] and bumps s_scanned; decode thread 0 copies it to s_scan for the other decode
	 * threads (behind their barrier) and bumps s_consumed once the block's offsets are free */
	__shared__ uint32_t s_hand[2][4];
	__shared__ ScanResult s_scan;
	__shared__ int s_bad;
	__shared__ unsigned long long s_cks;
	__shared__ uint32_t s_scanned, s_consumed, s_stop;

	const int tid = threadIdx.x;
	const bool scanner = tid >= GEN_THREADS;
	load_tables(tab, a.tables, tid, GEN_BLOCK);
	__syncthreads();

	uint32_t *buf0 = scr.buf + (size_t)blockIdx.x * scr.stride;
	uint32_t *buf1 = buf0 + scr.max_blen;
	uint32_t *hist = buf1 + scr.max_blen;               /* 2 * max_cols words */
	uint32_t *coloff0 = hist + 2 * (size_t)scr.max_cols; /* two buffers of max_cols words */

	for (;;) {
		if (tid == 0) {
			s_si = (int)atomicAdd(a.counter, 1u);
			s_cks = 0ull;
			s_scanned = 0;
			s_consumed = 0;
			s_stop = 0;
		}
		__syncthreads();
		const uint32_t si = (uint32_t)s_si;
		if (si >= a.count)
			break;
		const DevStream d = a.streams[si];
		const uint32_t level = d.level, cols = 1u << level, rows = d.rows;
		const uint32_t blen = rows * cols, limit = d.file_end + 8u;
		BitReader br;
		br.init((const uint32_t *)(a.blob + d.base_off), d.file_end);
		uint8_t *out = a.out + d.out_off;
		uint32_t *rh = a.resume_hist ? a.resume_hist + (size_t)si * a.resume_stride : nullptr;
		uint32_t P = d.bit0, pos = 0;
		int st = 0;
		uint32_t nok = 0;
		unsigned long long cks = 0ull;

		if (scanner) {
			/* ---- 1. scan, one block ahead of the decode */
			if (tid == GEN_THREADS) {
				for (uint32_t b = 0; b < d.n_attempt; b++) {
					while (b - atomicAdd(&s_consumed, 0u) >= 2u && !atomicAdd(&s_stop, 0u))
						__nanosleep(40);
					if (atomicAdd(&s_stop, 0u))
						break;
					const ScanResult sc = scan_block(br, P, limit, cols, rows, coloff0 + (b & 1u) * scr.max_cols,
									 0u, tab.kind, tab.k8);
					atomicExch(&s_hand[b & 1u][0], (uint32_t)sc.status);
					atomicExch(&s_hand[b & 1u][1], (uint32_t)sc.val);
					atomicExch(&s_hand[b & 1u][2], sc.ncols);
					atomicExch(&s_hand[b & 1u][3], sc.end);
					__threadfence_block(); /* the offsets (global scratch, same CTA) and the verdict, then the count */
					atomicExch(&s_scanned, b + 1u);
					if (sc.status != SCAN_OK)
						break; /* the stream ends with this block */
					P = sc.end;
				}
			}
		} else {
			for (uint32_t i = tid; i < 2 * cols; i += GEN_THREADS)
				hist[i] = (rh && d.resume) ? rh[i] : 0u; /* zeroed history: decode.c:812 */
			uint32_t *cur = buf0, *nxt = buf1;
			decode_sync();

			for (uint32_t b = 0; b < d.n_attempt; b++) {
				if (tid == 0) {
					while (atomicAdd(&s_scanned, 0u) <= b)
						__nanosleep(20);
					__threadfence_block();
					s_scan.status = (int)atomicAdd(&s_hand[b & 1u][0], 0u);
					s_scan.val = (int)atomicAdd(&s_hand[b & 1u][1], 0u);
					s_scan.ncols = atomicAdd(&s_hand[b & 1u][2], 0u);
					s_scan.end = atomicAdd(&s_hand[b & 1u][3], 0u);
					s_bad = 0;
				}
				decode_sync();
				const ScanResult sc = s_scan;
				const uint32_t *coloff = coloff0 + (b & 1u) * scr.max_cols;
				/* ---- 2. unpack (column sc.ncols is included when its payload ran past
				 * the limit: a t-code that still fits may be out of range first) */
				const uint32_t ncheck = sc.ncols + (sc.status == -7 ? 1u : 0u);
				for (uint32_t c = tid; c < ncheck; c += GEN_THREADS) {
					uint32_t Pc = coloff[c];
					uint32_t ind = br.peek(Pc) & 31u;
					int r = decode_column(br, Pc + 5u, limit, ind, tab.kind[ind], rows, sc.val,
							      cur + c, cols, tab.k8, tab.t);
					if (r < 0)
						s_bad = 1;
				}
				decode_sync();
				if (tid == 0)
					atomicExch(&s_consumed, b + 1u); /* this block's offsets may be overwritten */
				if (s_bad)
					st = -6;
				else if (sc.status != SCAN_OK)
					st = sc.status == SCAN_EOF ? 0 : sc.status;
				if (s_bad || sc.status != SCAN_OK)
					break; /* uniform: decided from shared state */
				P = sc.end;

				/* ---- 3. juggle (decode.c:528-577 in flat form) */
				uint32_t hoff = 0;
				for (uint32_t l = 1; l <= level; l++) {
					const uint32_t C = cols >> l;
					uint32_t *h = hist + hoff;
					for (uint32_t m = tid; m < blen; m += GEN_THREADS) {
						uint32_t v = juggle_at(cur, h, m, C);
						if (l == 1 && (m & (C - 1u)) == 0u)
							v += 1u; /* decode.c:561-564 */
						nxt[m] = v;
					}
					decode_sync();
					for (uint32_t i = tid; i < 2 * C; i += GEN_THREADS)
						h[i] = cur[blen - 2 * C + i];
					decode_sync();
					uint32_t *t = cur; cur = nxt; nxt = t;
					hoff += 2 * C;
				}

				/* ---- 4. output (decode.c:849-866) */
				uint32_t n = blen;
				if (n > d.words_limit - pos)
					n = d.words_limit - pos;
				for (uint32_t m = tid; m < n; m += GEN_THREADS) {
					uint32_t u = emit_word(out + (size_t)(pos + m) * a.fmt.wordlen,
							       (int32_t)cur[m] >> level, a.fmt);
					if (a.fmt.checksums)
						cks += (unsigned long long)(pos + m + 1u) * (unsigned long long)(u + 1ull);
				}
				pos += n;
				nok = b + 1;
				decode_sync();
			}
			if (tid == 0)
				atomicExch(&s_stop, 1u); /* the scan thread may be a block ahead of a stream that just ended */

			/* zero padding of the undelivered tail (acmtool.c:293-310) */
			{
				uint8_t *p = out + (size_t)pos * a.fmt.wordlen;
				/* up to the 16-byte boundary that ends this stream's slot: no stale bytes in the gaps */
				size_t nbytes = d.pad_words >= pos && d.pad_words
							? (((size_t)d.pad_words * a.fmt.wordlen + 15u) & ~(size_t)15u) -
								  (size_t)pos * a.fmt.wordlen
							: 0;
				for (size_t i = tid; i < nbytes; i += GEN_THREADS)
					p[i] = 0;
			}
			if (a.fmt.checksums) {
				for (int o = 16; o; o >>= 1)
					cks += __shfl_xor_sync(0xFFFFFFFFu, cks, o);
				if ((tid & 31) == 0)
					atomicAdd(&s_cks, cks);
			}
			decode_sync();
			if (rh)
				for (uint32_t i = tid; i < 2 * cols; i += GEN_THREADS)
					rh[i] = hist[i];
			if (tid == 0) {
				a.status[d.index] = st;
				a.words[d.index] = pos;
				a.cks[d.index] = s_cks;
				if (a.end_pos) {
					a.end_pos[2 * si] = P;
					a.end_pos[2 * si + 1] = nok;
				}
			}
		}
		__syncthreads(); /* both sides are done with this stream */
	}
}

/* CTAs of the generic kernel that fit one SM (threads, registers, the 17 KB of code tables) */
int generic_ctas_per_sm()
{
	static int cached = 0;
	if (!cached) {
		int nb = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, acm_decode_generic_kernel, GEN_BLOCK, 0) != cudaSuccess || nb < 1)
			nb = 4;
		cached = nb > 8 ? 8 : nb; /* beyond 8 the tables leave no L1 for the scan threads' loads */
	}
	return cached;
}

cudaError_t launch_generic(const KernelArgs &a, const GenericScratch &s, int n_ctas, cudaStream_t st)
{
	if (a.count == 0)
		return cudaSuccess;
	acm_decode_generic_kernel<<<n_ctas, GEN_BLOCK, 0, st>>>(a, s);
	return cudaGetLastError();
}

/* ------------------------------------------------------------------ header gather */

__global__ void acm_gather_headers_kernel(const uint8_t *blob, uint64_t blob_len,
					  const uint64_t *in_off, const uint32_t *in_len,
					  uint8_t *dst, uint64_t n)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	uint64_t off = in_off[i];
	uint32_t len = in_len[i];
	for (uint32_t k = 0; k < 48; k++) {
		uint8_t v = 0;
		if (k < len && off + k < blob_len)
			v = blob[off + k];
		dst[i * 48 + k] = v;
	}
}

cudaError_t launch_gather_headers(const uint8_t *blob, uint64_t blob_len, const uint64_t *in_off,
				  const uint32_t *in_len, uint8_t *dst, uint64_t n, cudaStream_t st)
{
	if (n == 0)
		return cudaSuccess;
	unsigned blocks = (unsigned)((n + 127) / 128);
	acm_gather_headers_kernel<<<blocks, 128, 0, st>>>(blob, blob_len, in_off, in_len, dst, n);
	return cudaGetLastError();
}

} // namespace acm
