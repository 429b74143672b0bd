/*
 * acm_split.cu -- the THROUGHPUT path for 16-row blocks of 128 columns (level 7: the shape of
 * BASELINE configs 1, 2, 4): decode_block (decode.c:580-611) split into its three stages, one
 * kernel each, with a table of block records between them:
 *
 *   walk    acm_walk_kernel: fill_block's control flow (decode.c:491-502) without its data flow.
 *           One stream per LANE, warps in lock step, the walk one table-driven state machine
 *           (acm_walk.cuh).  A lane that finishes a block retires it on its own (record + 128
 *           column offsets, written as coalesced rows by the whole warp) and goes on with the
 *           next block; a lane out of blocks takes the next stream from the queue.  The walk is
 *           the only serial part of the path; it is latency bound, so a batch of
 *           few long streams is spread thin over all SMs (one warp per sub-partition).
 *   unpack  acm_unpack_kernel: fill_block's data flow, the 14 fillers (decode.c:181-476).  With
 *           the column positions known every COLUMN is an independent unit of work.  A CTA takes
 *           a tile of 16 blocks = 2048 columns, sorts them by filler class (and the prefix-coded
 *           ones by length) with a counting sort in shared memory, and runs every class as
 *           straight-line code on full warps -- no divergence between a zero, a linear, a radix
 *           and a prefix-coded column, and lanes of one warp finish together.  A column leaves as
 *           sixteen signed bytes (the quantiser index, decode.c:174-177, times 2) in the layout
 *           the transform's lanes read; the rare index that does not fit a byte (linear columns
 *           of 8 bits and more) goes to a side array as int16 and is flagged in a per-block mask.
 *   lift    acm_lift_kernel: midbuf dequantisation (decode.c:591-600), juggle_block
 *           (decode.c:528-577) and output_values (decode.c:617-677) for runs of consecutive
 *           blocks of one stream, one warp per run.  The transform is a 7-stage FIR cascade in
 *           flat form: what a block needs of its predecessor is the
 *           last two rows of quantiser indices, so a run rebuilds the reference's wrapbuf
 *           (decode.c:803) from the previous block's bytes and runs are independent of each
 *           other.  Inside a run the history stays in registers.  Stages 1-2 (C = 64, 32) with
 *           lane j owning the words m = j mod 32, one transpose through shared memory, stages
 *           3-7 over a recomputed halo, 128-bit PCM stores.
 *   finish  acm_finish_kernel (acm_gen2.cu): per stream, what the reference's read loop reports.
 *
 * Nothing here waits for anything inside a kernel: the stages are ordered by the CUDA stream.
 * Intermediates (block records, column offsets, index bytes: 2.3 KB per block + the side array)
 * live in an arena that a plan sizes for one GROUP of streams and reuses from group to group.
 */
#include <cstdio>
#include "acm_walk.cuh"
#include "acm_kernels.cuh"

namespace acm {

namespace split {

constexpr int LEVEL = 7;
constexpr int COLS = 128;
constexpr int ROWS = 16;
constexpr int BLEN = COLS * ROWS;
/* Everything the transform touches is carried times 2^QS: the lifting is linear modulo 2^32
 * (decode.c:512), and with QS = 1 the 16 bits the output wants -- bits LEVEL..LEVEL+15 of the
 * reference's word, decode.c:620 -- are bytes 1 and 2 of ours: two results are packed by one
 * byte permute.  The factor rides in the index bytes (unpack stores 2 * idx), so that the
 * dequantisation is ONE two-way dot product per word: byte x val. */
constexpr int QS = 1;
constexpr int OSH = LEVEL + QS; /* 8 */

/* ================================================================== walk */

using namespace walk;

constexpr int WALK_THREADS_MAX = 32 * SW;

__global__ void __launch_bounds__(WALK_THREADS_MAX, 1) acm_walk_kernel(KernelArgs a, SplitArgs g)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	SmemWalk &sm = *reinterpret_cast<SmemWalk *>(smem_raw);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	for (int i = tid; i < ACM_UNI_PAGES * ACM_UNI_PSIZE / 2; i += blockDim.x)
		reinterpret_cast<uint32_t *>(sm.uni16)[i] = reinterpret_cast<const uint32_t *>(a.tables->uni16)[i];
	for (int i = tid; i < (RW + 5) * RROW; i += blockDim.x)
		(&sm.ring[0][0])[i] = 0u;
	__syncthreads();

	Ring ring;
	const uint32_t lane4 = 4u * (uint32_t)(warp * 32 + lane);
	ring.rw = &sm.ring[0][warp * 32 + lane];
	ring.pol = l2_keep_policy();
	ring.safe = a.blob;
	ring.hold_c0 = 0u;
	ring.idle();
	uint16_t *const off0 = reinterpret_cast<uint16_t *>(&sm.off[warp][lane * OFFP]);
	const uint32_t cp0 = (uint32_t)__cvta_generic_to_shared(off0);
	const unsigned char *uni = reinterpret_cast<const unsigned char *>(sm.uni16);

	/* mode: 0 no stream, 1 block header pending, 2 walking, 3 block walked (to retire) */
	int mode = 0;
	bool exhausted = false, hdr_eof = false;
	uint32_t cur = 0, P = 0, blk = 0, limit = 0, nmax = 0, val = 0;
	uint64_t rec_base = 0;
	uint32_t cp = cp0;
	const uint32_t cpend = cp0 + 2u * COLS;
	Walk s;
	s.Q = 0u;
	s.Q32 = 0u;
	s.s8 = UNI_HALT8;
	s.msk = MSK_K;

#ifdef WALK_PROF
	/* tuning build: cycles of warp 0 of CTA 0 (the batch's longest streams) per phase -> prof[40..] */
	unsigned long long pt0 = clock64(), pacc[5] = { 0, 0, 0, 0, 0 };
	unsigned long long pperiods = 0, pouter = 0;
#define WPROF(k) do { const unsigned long long t_ = clock64(); pacc[k] += t_ - pt0; pt0 = t_; } while (0)
#else
#define WPROF(k) do { } while (0)
#endif
	for (;;) {
		WPROF(4);
#ifdef WALK_PROF
		pouter++;
#endif
		/* ---- retire walked blocks: verdict, record, the 128 column offsets as one 256-byte row */
		int status = SCAN_EOF;
		uint32_t ncols = 0, pend = P;
		if (mode == 3) {
			if (hdr_eof) {
				/* pwr / val cannot be read: GET_BITS_EXPECT_EOF decode.c:588-589 */
			} else if (s.s8 == UNI_HALT8 && s.Q + 1u <= limit) {
				status = SCAN_OK; /* 128 columns, every read inside the stream */
				ncols = COLS;
				pend = s.Q + 1u;
			} else {
				/* bad selector, or the stream ended inside the block: walk it again with the
				 * reference's verdicts (at most once per stream) */
				const DevStream d = a.streams[cur];
				BitReader br;
				br.init(reinterpret_cast<const uint32_t *>(a.blob + d.base_off), d.file_end);
				const ScanResult sc = scan_block(br, P, limit, (uint32_t)COLS, (uint32_t)ROWS, off0, P,
								 a.tables->kind, a.tables->k8);
				status = sc.status;
				ncols = sc.ncols;
				pend = sc.end;
				val = (uint32_t)sc.val;
				if (a.prof)
					atomicAdd(a.prof + 33, 1ull);
			}
		}
		{
			const unsigned wm = __ballot_sync(0xFFFFFFFFu, mode == 3);
			uint16_t *const row = g.coff16 + (rec_base + blk) * (uint64_t)COLS;
			const unsigned long long row64 = (unsigned long long)(uintptr_t)row;
			__syncwarp();
			for (unsigned mm = wm; mm; mm &= mm - 1u) {
				const int i = __ffs((int)mm) - 1;
				const unsigned long long r = __shfl_sync(0xFFFFFFFFu, row64, i);
				const uint2 v = *reinterpret_cast<const uint2 *>(&sm.off[warp][i * OFFP + 8 * lane]);
				reinterpret_cast<uint2 *>((uintptr_t)r)[lane] = v;
			}
			__syncwarp();
		}
		if (mode == 3) {
			uint4 *rp = reinterpret_cast<uint4 *>(g.rec + rec_base + blk);
			rp[0] = make_uint4(P, pend, val, (uint32_t)status);
			rp[1] = make_uint4(ncols, cur, blk, g.epoch);
			blk++;
			if (status != SCAN_OK || blk >= nmax) {
				g.nscan[cur] = blk; /* the stream ends with this block */
				mode = 0;
			} else {
				P = pend;
				mode = 1;
			}
		}
		/* ---- a lane without a stream takes the next one from the queue */
		if (mode == 0 && !exhausted) {
			const uint32_t idx = atomicAdd(a.counter, 1u);
			if (idx < a.count) {
				const DevStream d = a.streams[idx];
				const Gen2Stream gs = g.gs[idx];
				cur = idx;
				P = d.bit0;
				blk = 0;
				limit = d.file_end + 8u;
				nmax = d.n_attempt < gs.max_blocks ? d.n_attempt : gs.max_blocks;
				rec_base = gs.rec_base;
				if (nmax == 0) {
					g.nscan[cur] = 0u;
				} else {
					ring.start(a.blob + d.base_off, a.blob_room > d.base_off ? a.blob_room - d.base_off : 0,
						   d.file_end, P);
					mode = 1;
				}
			} else {
				exhausted = true;
				ring.idle();
				P = 0;
			}
		} else if (mode == 0) {
			/* out of streams: over the two zero ring words, on the HALT page */
		}
		if (mode == 1) {
			s.Q = P - 1u; /* P = 0 is a position like any other (Q wraps) */
			s.Q32 = s.Q << 5;
			s.s8 = UNI_HALT8;
			s.msk = MSK_K;
			cp = cp0;
			hdr_eof = false;
		} else if (mode == 0) {
			s.Q = 0u;
			s.Q32 = 0u;
			s.s8 = UNI_HALT8;
			s.msk = MSK_K;
		}
		if (!__any_sync(0xFFFFFFFFu, mode != 0 || !exhausted))
			break;
		WPROF(0); /* 0: retire + fetch */
		/* ---- walk until some lane has a block to retire */
		do {
#ifdef WALK_PROF
			pperiods++;
#endif
			ring.topup(s.Q + 1u);
			WPROF(1); /* 1: top-up */
			if (mode == 1) {
				/* pwr(4) / val(16): decode.c:588-589 */
				if (s.Q + 21u > limit) {
					hdr_eof = true;
					mode = 3;
				} else if (s.Q + 1u <= ring.ready_p) {
					const uint32_t *rp = ring_word(&sm.ring[0][0], lane4, s.Q32);
					const uint32_t w1 = fsr(rp[0], rp[RROW], s.Q);
					val = (w1 >> 5) & 0xFFFFu;
					s.Q += 20u;
					s.Q32 = s.Q << 5;
					s.s8 = 0u;
					s.msk = MSK_SEL;
					mode = 2;
				}
			}
			WPROF(2); /* 2: header */
#pragma unroll 4
			for (int k = 0; k < PERIOD; k++)
				step(s, cp, cpend, P - 1u, &sm.ring[0][0], lane4, ring.ready_p, uni);
			if (mode == 2 && (s.s8 == UNI_HALT8 || s.s8 == UNI_BAD8))
				mode = 3;
			WPROF(3); /* 3: steps */
		} while (!__any_sync(0xFFFFFFFFu, mode == 3 || (mode == 0 && !exhausted)));
	}
#ifdef WALK_PROF
	if (blockIdx.x == 0 && tid == 0 && a.prof) {
		for (int k = 0; k < 5; k++)
			a.prof[40 + k] = pacc[k];
		a.prof[45] = pperiods;
		a.prof[46] = pouter;
	}
#endif
}

/* ================================================================== walk, one stream */

/*
 * The walk of ONE stream, for the streaming API (acm_stream.cu): what bounds acm_read on a single
 * stream is the latency of this serial walk, so here it is a scalar routine on one lane, with
 * branches, instead of 32 streams in lock step.  Same state machine (uni16).  The CTA's other
 * seven warps keep a shared-memory ring of the stream's bytes filled ahead of the walker (with the
 * reference's end-of-file rule applied: bits past the file read as zero, decode.c:57-61), so every
 * load on the walker's chain is a shared-memory load of known latency:
 *   column start:  ring address (2 ALU) -> the 128 stream bits from the position on (4 LDS, together)
 *                  -> funnel shift -> mask -> table LDS -> position += advance
 *   inside a prefix-coded column the next 96 bits stay in a register window:
 *                  table LDS -> funnel shift -> mask -> table LDS
 * Walks blocks [b0, b0 + nb) from bit position P0 and leaves the same records and column offsets
 * as the batch walk; nscan[0] = the blocks walked so far (b0 + those of this launch).
 */
constexpr int W1_THREADS = 256;
constexpr int W1_STAGERS = W1_THREADS - 32;
constexpr uint32_t W1_PBYTES = 4u << ACM_UNI_KBITS; /* bytes per page of the widened table */
constexpr uint32_t W1_RING = 4096;                  /* ring words (16 KB) */
constexpr uint32_t W1_UNI_BYTES = 4u * ACM_UNI_PAGES * ACM_UNI_PSIZE;
/* words a block can span at most (header, 128 selectors, 128 columns of 16 x 16 bits) + the window's read-ahead */
constexpr uint32_t W1_BLOCK_WORDS = (20u + 128u * (5u + 256u) + 31u) / 32u + 8u;
static_assert(W1_BLOCK_WORDS * 2u < W1_RING, "the ring holds more than a block");

/*
 * Shared memory of the one-stream walker, by ABSOLUTE shared-window address (the dynamic
 * allocation starts somewhere in the first kilobytes; everything sits above 32 K):
 *   [32 K, 187 K)        uni32: uni16 widened to 32-bit entries, bits to advance (byte 0) | ABSOLUTE
 *                        address of the next page (a multiple of 1024) | W1_DONE when the next step is
 *                        at a selector again (the page is then the selector page: a lookup issued
 *                        before that is known reads a valid address) | W1_BAD on the page a bad
 *                        selector leads to
 *   [192 K, 208 K)       ring: word i of the stream at word i % W1_RING
 *   [208 K, 208 K + 16)  the ring's first four words again, so that the four words from any index
 *                        on are consecutive
 *   then the block's column offsets and the control words of the two sides.
 * Aligned like this, "ring base | masked position" and "table base | masked window" are one LOP3
 * each, and a prefix-code step's next address is (entry & ~1023) | (window & 1020): no add on
 * the walker's chain.  The entry is used as it is as a funnel-shift count (the low five bits
 * count; no step advances 32 bits or more where the window is shifted), and its byte 0 is added
 * to the position by a dot-product instruction.
 */
constexpr uint32_t W1_A_UNI = 32768u;
constexpr uint32_t W1_A_RING = 196608u;
constexpr uint32_t W1_DONE = 0x100u, W1_BAD = 0x200u; /* entry flags (pages are 1024-byte aligned: bits 8, 9 are free) */
static_assert(W1_A_UNI + W1_UNI_BYTES <= W1_A_RING, "the table ends below the ring");
constexpr uint32_t W1_A_TAIL = W1_A_RING + 4u * W1_RING;
constexpr uint32_t W1_A_OFF = W1_A_TAIL + 16u;               /* u16 off[COLS + 8] */
constexpr uint32_t W1_A_CTL = W1_A_OFF + 2u * (COLS + 8);    /* wpos, fill, quit, s_target, s_quit */
constexpr uint32_t W1_A_END = W1_A_CTL + 32u;
constexpr size_t W1_SMEM = W1_A_END; /* the allocation starts above address 0: this much always reaches W1_A_END */

__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}
/* the hand-over words between the walker and the stagers (wpos, fill, quit): shared-memory atomics,
 * ordered by __threadfence_block() on both sides (racecheck understands atomics, not flag words) */
__device__ __forceinline__ uint32_t lds32_volatile(uint32_t addr)
{
	uint32_t v;
	asm volatile("atom.shared.or.b32 %0, [%1], 0;" : "=r"(v) : "r"(addr) : "memory");
	return v;
}
__device__ __forceinline__ void sts32_volatile(uint32_t addr, uint32_t v)
{
	uint32_t old;
	asm volatile("atom.shared.exch.b32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
}
/* plain volatile accesses: the stagers' own broadcast words, ordered by their barrier */
__device__ __forceinline__ uint32_t lds32_plain(uint32_t addr)
{
	uint32_t v;
	asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
	return v;
}
__device__ __forceinline__ void sts32_plain(uint32_t addr, uint32_t v)
{
	asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v)
{
	asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v)
{
	asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

/* the four ring words from the one that holds bit R on */
struct Win4 {
	uint32_t w0, w1, w2, w3;
};
__device__ __forceinline__ Win4 ring4(uint32_t R)
{
	/* (R >> 3) & 0x3FFC: SHF + LOP3; the ring's base is the loads' immediate offset */
	const uint32_t a0 = (R >> 3) & ((W1_RING - 1u) << 2);
	Win4 w;
	/* a0 + 4 k stays inside [0, 16 K + 12): the tail copy sits right behind the ring */
	asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(w.w0) : "r"(a0), "n"(W1_A_RING));
	asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(w.w1) : "r"(a0), "n"(W1_A_RING + 4u));
	asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(w.w2) : "r"(a0), "n"(W1_A_RING + 8u));
	asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(w.w3) : "r"(a0), "n"(W1_A_RING + 12u));
	return w;
}

__global__ void __launch_bounds__(W1_THREADS, 1) acm_walk1_kernel(KernelArgs a, SplitArgs g, uint32_t b0, uint32_t nb, uint32_t P0)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int tid = threadIdx.x, lane = tid & 31;
	const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw);
	/* the layout is by absolute address: the allocation must start at or below the table (it does:
	 * the window's first kilobyte is the system's) and reach W1_A_END */
	if (sbase > W1_A_UNI) {
		if (tid == 0)
			atomicExch(a.errflag, 1u);
		return;
	}
	for (int i = tid; i < ACM_UNI_PAGES * ACM_UNI_PSIZE; i += W1_THREADS) {
		const uint32_t e = a.tables->uni16[i];
		const uint32_t page = e >> 8;
		uint32_t w = (e & 0xFFu) | (page ? W1_A_UNI + page * W1_PBYTES : W1_A_UNI | W1_DONE);
		if ((uint32_t)i / ACM_UNI_PSIZE == (uint32_t)ACM_UNI_BAD)
			w = W1_A_UNI | W1_DONE | W1_BAD; /* where a bad selector leads: 0 bits, done, flagged */
		asm volatile("st.shared.u32 [%0], %1;" ::"r"(W1_A_UNI + 4u * (uint32_t)i), "r"(w) : "memory");
	}
	const DevStream d = a.streams[0];
	const Gen2Stream gs = g.gs[0];
	/* a launch that continues where the previous one stopped (P0 = 0, b0 > 0) finds its start in
	 * the previous block's record */
	if (b0 && !P0)
		P0 = __ldcg(&g.rec[gs.rec_base + b0 - 1u].end);
	constexpr uint32_t A_WPOS = W1_A_CTL, A_FILL = W1_A_CTL + 4u, A_QUIT = W1_A_CTL + 8u, A_TARGET = W1_A_CTL + 12u,
			   A_SQUIT = W1_A_CTL + 16u;
	if (tid == 0) {
		sts32_volatile(A_WPOS, P0 >> 5);
		sts32_volatile(A_FILL, (P0 >> 5) & ~3u);
		sts32_volatile(A_QUIT, 0u);
	}
	__syncthreads();

	if (tid >= 32) {
		/* ---- stagers: 16-byte chunks of the stream into the ring, up to a ring's length ahead */
		const int st = tid - 32;
		const uint8_t *base = a.blob + d.base_off;
		const uint32_t room16 = a.blob_room > d.base_off ? (uint32_t)((a.blob_room - d.base_off) >> 4) : 0u;
		const uint32_t fe_byte = d.file_end >> 3;
		const uint32_t full16 = fe_byte >> 4 < room16 ? fe_byte >> 4 : room16;
		uint32_t fill = (P0 >> 5) & ~3u;
		for (;;) {
			if (st == 0) {
				sts32_plain(A_TARGET, (lds32_volatile(A_WPOS) & ~3u) + W1_RING);
				sts32_plain(A_SQUIT, lds32_volatile(A_QUIT));
			}
			asm volatile("bar.sync 1, %0;" ::"n"(W1_STAGERS) : "memory");
			const uint32_t target = lds32_plain(A_TARGET);
			if (lds32_plain(A_SQUIT))
				break;
			uint32_t n = (target - fill) >> 2;
			n = n < (uint32_t)W1_STAGERS ? n : (uint32_t)W1_STAGERS;
			if ((uint32_t)st < n) {
				const uint32_t c = (fill >> 2) + (uint32_t)st;
				uint4 v = make_uint4(0u, 0u, 0u, 0u);
				if (c < full16) {
					v = __ldg(reinterpret_cast<const uint4 *>(base) + c);
				} else if (c < room16 && c * 16u < fe_byte) {
					/* the chunk that holds the end of the file: 1 .. 15 of its bytes exist */
					v = __ldg(reinterpret_cast<const uint4 *>(base) + c);
					const uint32_t nbytes = fe_byte - c * 16u;
					uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
					for (int k = 0; k < 4; k++) {
						const uint32_t m = nbytes > 4u * k ? nbytes - 4u * k : 0u;
						w[k] = m >= 4u ? w[k] : m ? w[k] & ((1u << (8u * m)) - 1u) : 0u;
					}
					v = make_uint4(w[0], w[1], w[2], w[3]);
				}
				const uint32_t at = (c << 4) & (4u * W1_RING - 1u);
				sts128(W1_A_RING + at, v);
				if (at == 0u)
					sts128(W1_A_TAIL, v);
			}
			asm volatile("bar.sync 1, %0;" ::"n"(W1_STAGERS) : "memory");
			fill += 4u * n;
			if (st == 0) {
				__threadfence_block();
				sts32_volatile(A_FILL, fill);
			}
			if (n == 0u)
				__nanosleep(2000);
		}
		return;
	}

	/* ---- the walker: lane 0 walks, the warp writes the block's column offsets */
	const uint32_t limit = d.file_end + 8u;
	constexpr uint32_t M_SEL = 0x1FFFu << 2, M_K = ((1u << ACM_UNI_KBITS) - 1u) << 2;
	uint32_t P = P0, b = b0;
#ifdef W1_DEBUG
	unsigned long long t0c = clock64(), t0n, waitc = 0;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0n));
#endif
	for (; b < b0 + nb; b++) {
		int status = SCAN_EOF;
		uint32_t ncols = 0, pend = P, val = 0;
		if (lane == 0 && P + 20u <= limit) {
			/* the whole block is in the ring before its walk starts */
			sts32_volatile(A_WPOS, P >> 5);
#ifdef W1_DEBUG
			unsigned long long tw = clock64();
#endif
			while ((int32_t)(lds32_volatile(A_FILL) - ((P >> 5) + W1_BLOCK_WORDS)) < 0)
				;
#ifdef W1_DEBUG
			waitc += clock64() - tw;
#endif
			__threadfence_block();
			bool clean = true;
			Win4 w = ring4(P);
			val = (fsr(w.w0, w.w1, P) >> 4) & 0xFFFFu; /* pwr(4) val(16): decode.c:588-589 */
			/* R = position - 2: the 32 bits at R, masked, are the byte offset of a 4-byte table entry */
			uint32_t R = P + 18u;
			w = ring4(R);
			uint32_t flags = 0u;
#pragma unroll 4
			for (uint32_t c = 0; c < (uint32_t)COLS; c++) {
				uint32_t lo = fsr(w.w0, w.w1, R);
				uint32_t e;
				asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(e) : "r"(lo & M_SEL), "n"(W1_A_UNI));
				uint32_t mid = fsr(w.w1, w.w2, R), hi = fsr(w.w2, w.w3, R);
				sts16(W1_A_OFF + 2u * c, R + 2u - P);
				R = __dp4a(e, 1u, R); /* += byte 0 */
				/* the next column's window, as if this one were of fixed size (it is, more often than not):
				 * on its way before the branch below is resolved */
				w = ring4(R);
				if (!(e & W1_DONE)) {
					/* inside a prefix-coded column (<= 80 payload bits, all in the window), or on the way to
					 * the SKIP6 / BAD page.  Every lookup is issued BEFORE the branch on the entry before it
					 * resolves (a finished column's entry points at the selector page: the one lookup too
					 * many reads a valid address and is dropped), so the branch waits in the shadow of the
					 * load instead of on the chain: table LDS -> funnel shift -> mask -> table LDS */
#define W1_SHIFT_LOOKUP(ec, out)                                                                                       \
	do {                                                                                                            \
		lo = fsr(lo, mid, ec);                                                                                  \
		mid = fsr(mid, hi, ec);                                                                                 \
		hi = fsr(hi, 0u, ec);                                                                                   \
		uint32_t page_, at_;                                                                                    \
		asm volatile("and.b32 %0, %1, %2;" : "=r"(page_) : "r"(ec), "n"(~(W1_PBYTES - 1u)));                    \
		asm volatile("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(at_) : "r"(lo), "n"(M_K), "r"(page_));          \
		out = lds32(at_);                                                                                       \
	} while (0)
					uint32_t ea, eb, spec; /* ea / eb alternate: no register copy between a load and the branch behind it */
					W1_SHIFT_LOOKUP(e, ea);
					for (;;) {
						W1_SHIFT_LOOKUP(ea, eb);
						R = __dp4a(ea, 1u, R);
						if (ea & W1_DONE) {
							flags |= ea;
							spec = eb;
							break;
						}
						W1_SHIFT_LOOKUP(eb, ea);
						R = __dp4a(eb, 1u, R);
						if (eb & W1_DONE) {
							flags |= eb;
							spec = ea;
							break;
						}
					}
					w = ring4(R);
					/* the lookup too many is "used" (selector-page entries never carry W1_BAD), and only here,
					 * when it has long landed: the compiler would otherwise sink it below the branch */
					flags |= spec & W1_BAD;
				}
			}
			clean = !(flags & W1_BAD);
			if (clean && R + 2u <= limit) {
				status = SCAN_OK; /* 128 columns, every read inside the stream */
				ncols = COLS;
				pend = R + 2u;
			} else {
				/* bad selector, or the stream ended inside the block: the reference's verdicts */
				BitReader br;
				br.init(reinterpret_cast<const uint32_t *>(a.blob + d.base_off), d.file_end);
				uint16_t *off = reinterpret_cast<uint16_t *>(smem_raw + (W1_A_OFF - sbase));
				const ScanResult sc = scan_block(br, P, limit, (uint32_t)COLS, (uint32_t)ROWS, off, P, a.tables->kind, a.tables->k8);
				status = sc.status;
				ncols = sc.ncols;
				pend = sc.end;
				val = (uint32_t)sc.val;
			}
		}
		status = __shfl_sync(0xFFFFFFFFu, status, 0);
		pend = __shfl_sync(0xFFFFFFFFu, pend, 0);
		__syncwarp();
		{
			uint2 v;
			asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(W1_A_OFF + 8u * (uint32_t)lane) : "memory");
			reinterpret_cast<uint2 *>(g.coff16 + (gs.rec_base + b) * (uint64_t)COLS)[lane] = v;
		}
		if (lane == 0) {
			uint4 *rp = reinterpret_cast<uint4 *>(g.rec + gs.rec_base + b);
			rp[0] = make_uint4(P, pend, val, (uint32_t)status);
			rp[1] = make_uint4(ncols, 0u, b, g.epoch);
		}
		__syncwarp();
		if (status != SCAN_OK) {
			b++;
			break; /* the stream ends with this block */
		}
		P = pend;
	}
	if (lane == 0) {
		g.nscan[0] = b;
		sts32_volatile(A_QUIT, 1u);
#ifdef W1_DEBUG
		unsigned long long t1n;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1n));
		printf("walk1: %u blocks, %llu cycles (%llu waiting for the ring), %llu ns\n", b - b0, clock64() - t0c, waitc, t1n - t0n);
#endif
	}
}

/* ================================================================== unpack */

constexpr int UB = 16;              /* blocks per tile */
constexpr int U_THREADS = 256;
constexpr int U_COLS = UB * COLS;   /* 2048 columns per tile */
constexpr int U_PER = U_COLS / U_THREADS;
/* work classes in the order they are processed: the longest first.  Prefix-coded columns by payload
 * length (lanes of a warp then run about the same number of table steps), linear columns by
 * their width (the field positions are compile-time constants), radix-coded ones by code. */
enum {
	C_K7 = 0, C_K0 = 7,   /* prefix codes, eight length classes, longest = 0 */
	C_WIDE = 8,           /* linear, 8 .. 16 bits: int16 side array */
	C_T2 = 9, C_T0 = 11,  /* t37, t27, t15 */
	C_L7 = 12, C_L3 = 16, /* linear, 7 .. 3 bits */
	NCLS = 17, C_NONE = 31
};
constexpr int U_LIST = U_COLS + 32 * NCLS; /* every class padded to whole warps */

struct SmemUnpack {
	uint64_t k8w[ACM_K8_SIZE];
	uint16_t t[ACM_T_SIZE];
	uint8_t kind[32];
	uint32_t pos[U_COLS];        /* P of the column's payload */
	uint16_t list[U_LIST];       /* the tile's columns, sorted by class: column | aux << 11; 0xFFFF = nothing */
	uint8_t chunk_cls[U_LIST / 32];
	uint32_t cnt[NCLS + 1], base[NCLS + 1];
	uint32_t next_chunk;
	uint32_t wmask[UB][4];
	const uint32_t *bbase[UB];   /* per block of the tile: stream base */
	uint32_t bP[UB], bend[UB], bfe[UB], bstream[UB], bno[UB], bcheck[UB], bflags[UB];
};
enum { BF_OK = 1, BF_NEAR = 2 }; /* block decodes; block (and the unpackers' read-ahead) reaches the end of its file */

/* word i of a stream.  near = the block reaches the end of its file: bits at and past file_end read
 * as zero (decode.c:57-61), and memory past the word that holds the file's last bit is not touched */
__device__ __forceinline__ uint32_t ldw(const uint32_t *w, uint32_t i, uint32_t fe, bool near)
{
	if (!near)
		return __ldg(w + i);
	const uint32_t fe_word = fe >> 5, fe_tail = fe & 31u;
	if (i < fe_word)
		return __ldg(w + i);
	if (i == fe_word && fe_tail)
		return __ldg(w + i) & ((1u << fe_tail) - 1u);
	return 0u;
}
struct StreamWords {
	const uint32_t *w;
	uint32_t fe;
	bool near;
	__device__ __forceinline__ uint32_t word(uint32_t i) const { return ldw(w, i, fe, near); }
};

/* eight 4-bit two's complement values -> eight signed bytes, times 2^QS */
__device__ __forceinline__ void nib_to_bytes(uint32_t a, uint32_t &b0, uint32_t &b1)
{
	const uint32_t lo4 = a & 0x0F0F0F0Fu, hi4 = (a >> 4) & 0x0F0F0F0Fu;
	const uint32_t n0 = __byte_perm(lo4, hi4, 0x5140), n1 = __byte_perm(lo4, hi4, 0x7362);
	/* sign fill: bit 3 of the nibble, times (0xF0 << QS & 0xFF) / 8 */
	constexpr uint32_t FILL = ((0xF0u << QS) & 0xFFu) >> 3;
	b0 = (n0 << QS) | ((n0 & 0x08080808u) * FILL);
	b1 = (n1 << QS) | ((n1 & 0x08080808u) * FILL);
}

/* f_linear (decode.c:196-206) for a column of IND-bit codes, IND + QS <= 8: W = the 128 stream bits
 * from the column's payload on; the sixteen (code - 2^(IND-1)) << QS as signed bytes.  Every field
 * position is a constant. */
template <int IND>
__device__ __forceinline__ uint4 lin_narrow(const uint32_t (&W)[4])
{
	constexpr uint32_t MASK = (1u << IND) - 1u, MIDQ = (1u << (IND - 1)) << QS;
	constexpr uint32_t SB = MIDQ * 0x01010101u;
	constexpr uint32_t FILLM = ((0xFFu << (IND + QS)) & 0xFFu) >> (IND - 1 + QS);
	uint32_t q[4];
#pragma unroll
	for (int j = 0; j < 4; j++) {
		uint32_t acc = 0u;
#pragma unroll
		for (int r = 0; r < 4; r++) {
			const int bit = (4 * j + r) * IND, k = bit >> 5, sft = bit & 31;
			uint32_t v = sft + IND <= 32 ? W[k] >> sft : __funnelshift_r(W[k], W[k + 1 < 4 ? k + 1 : 3], sft);
			if (sft + IND != 32)
				v &= MASK;
			acc += v << (8 * r + QS);
		}
		/* code ^ mid = the index in IND-bit two's complement; fill the bits above its sign */
		const uint32_t y = acc ^ SB;
		q[j] = y | ((y & SB) * FILLM);
	}
	return make_uint4(q[0], q[1], q[2], q[3]);
}

__global__ void __launch_bounds__(U_THREADS) acm_unpack_kernel(KernelArgs a, SplitArgs g)
{
	__shared__ SmemUnpack sm;
	const int tid = threadIdx.x, lane = tid & 31;
	for (int i = tid; i < ACM_K8_SIZE; i += U_THREADS)
		sm.k8w[i] = a.tables->k8w[i];
	for (int i = tid; i < ACM_T_SIZE; i += U_THREADS)
		sm.t[i] = a.tables->t[i];
	if (tid < 32)
		sm.kind[tid] = a.tables->kind[tid];

	for (uint64_t tile = blockIdx.x; tile * UB < g.n_blocks; tile += gridDim.x) {
		const uint64_t g0 = tile * UB;
		__syncthreads(); /* the previous tile is done with the shared state */
		/* ---- the tile's blocks */
		if (tid < UB) {
			const uint64_t gb = g0 + tid;
			uint32_t check = 0, flags = 0;
			if (gb < g.n_blocks) {
				const uint4 r0 = __ldcg(reinterpret_cast<const uint4 *>(g.rec + gb));
				const uint4 r1 = __ldcg(reinterpret_cast<const uint4 *>(g.rec + gb) + 1);
				if (r1.w == g.epoch) { /* walked in this run */
					const int status = (int)r0.w;
					const DevStream d = a.streams[r1.y];
					flags = status == SCAN_OK ? BF_OK : 0u;
					/* column ncols is included when its payload ran past the limit: a radix code
					 * that still fits may be out of range first (decode.c:412/:438/:464) */
					check = flags ? (uint32_t)COLS : r1.x + (status == -7 ? 1u : 0u);
					/* the unpackers read up to 160 bits past a column's selector */
					if (!flags || r0.y + 224u > d.file_end)
						flags |= BF_NEAR;
					sm.bP[tid] = r0.x;
					sm.bend[tid] = r0.y - r0.x;
					sm.bfe[tid] = d.file_end;
					sm.bstream[tid] = r1.y;
					sm.bno[tid] = r1.z;
					sm.bbase[tid] = reinterpret_cast<const uint32_t *>(a.blob + d.base_off);
				}
			}
			sm.bcheck[tid] = check;
			sm.bflags[tid] = flags;
		}
		if (tid < NCLS + 1)
			sm.cnt[tid] = 0u;
		if (tid < UB * 4)
			(&sm.wmask[0][0])[tid] = 0u;
		if (tid == 0)
			sm.next_chunk = 0u;
		__syncthreads();

		/* ---- classify: where the column's payload sits, its class, the class's running count.  A warp
		 * looks at 32 consecutive columns of one block. */
		uint32_t mine[U_PER]; /* class << 24 | aux << 16 | rank */
#pragma unroll
		for (int k = 0; k < U_PER; k++) {
			const uint32_t cid = (uint32_t)tid + (uint32_t)k * U_THREADS;
			const uint32_t bi = cid >> 7, col = cid & 127u;
			const uint32_t ncheck = sm.bcheck[bi], flags = sm.bflags[bi];
			const uint16_t *coff = g.coff16 + (g0 + bi) * (uint64_t)COLS;
			const uint32_t off = col < ncheck ? coff[col] : 0xFFFFu;
			/* the next column's offset: the neighbour lane has it */
			uint32_t nxt = __shfl_down_sync(0xFFFFFFFFu, off, 1);
			if (lane == 31)
				nxt = col + 1u < ncheck ? coff[col + 1u] : sm.bend[bi];
			uint32_t cls = C_NONE, aux = 0u;
			if (col < ncheck) {
				const uint32_t Pc = sm.bP[bi] + off;
				const bool near = (flags & BF_NEAR) != 0u;
				const uint32_t *w = sm.bbase[bi];
				const uint32_t iw = Pc >> 5;
				const uint32_t w0 = ldw(w, iw, sm.bfe[bi], near), w1 = ldw(w, iw + 1u, sm.bfe[bi], near);
				const uint32_t ind = fsr(w0, w1, Pc) & 31u;
				const uint32_t kind = sm.kind[ind];
				const uint32_t c = kind & 7u, sub = kind >> 3;
				sm.pos[cid] = Pc + 5u;
				if (c == ACM_CLS_T) {
					cls = C_T0 - sub;
					aux = flags & BF_OK;
				} else if (flags & BF_OK) {
					if (c == ACM_CLS_ZERO) {
						/* f_zero decode.c:181-188 */
						*reinterpret_cast<uint4 *>(g.inter + (g0 + bi) * (uint64_t)BLEN + (col & 31u) * 64u + (col >> 5) * 16u) =
							make_uint4(0u, 0u, 0u, 0u);
					} else if (c == ACM_CLS_LINEAR) {
						if (ind + QS <= 8u) {
							cls = C_L3 - (ind - 3u);
						} else {
							cls = C_WIDE;
							aux = ind - 8u;
						}
					} else if (c == ACM_CLS_K) {
						uint32_t len = nxt - off - 5u; /* 8 .. 80 bits */
						len = len > 80u ? 80u : len;
						const uint32_t lc = len > 16u ? (len - 9u) >> 3 : 0u; /* 0 .. 8 */
						cls = C_K0 - (lc > 7u ? 7u : lc);
						aux = sub;
					}
				}
			}
			/* rank within the class: one shared-memory atomic per class and warp */
			const unsigned peers = __match_any_sync(0xFFFFFFFFu, cls);
			uint32_t base = 0;
			const int leader = __ffs((int)peers) - 1;
			if (lane == leader && cls != C_NONE)
				base = atomicAdd(&sm.cnt[cls], (uint32_t)__popc(peers));
			base = __shfl_sync(0xFFFFFFFFu, base, leader);
			mine[k] = (cls << 24) | (aux << 16) | (base + (uint32_t)__popc(peers & ((1u << lane) - 1u)));
		}
		__syncthreads();
		if (tid == 0) {
			uint32_t acc = 0;
			for (int c = 0; c < NCLS; c++) {
				sm.base[c] = acc;
				acc += (sm.cnt[c] + 31u) & ~31u;
			}
			sm.base[NCLS] = acc;
		}
		__syncthreads();
#pragma unroll
		for (int k = 0; k < U_PER; k++) {
			const uint32_t cid = (uint32_t)tid + (uint32_t)k * U_THREADS;
			const uint32_t cls = mine[k] >> 24, aux = (mine[k] >> 16) & 31u, rank = mine[k] & 0xFFFFu;
			if (cls != C_NONE)
				sm.list[sm.base[cls] + rank] = (uint16_t)(cid | (aux << 11));
		}
		if (tid < NCLS) {
			/* the holes at the end of every class, and the class of every chunk of 32 entries */
			const uint32_t b0 = sm.base[tid], n = sm.cnt[tid], b1 = sm.base[tid + 1];
			for (uint32_t i = b0 + n; i < b1; i++)
				sm.list[i] = 0xFFFFu;
			for (uint32_t ch = b0 >> 5; ch < b1 >> 5; ch++)
				sm.chunk_cls[ch] = (uint8_t)tid;
		}
		__syncthreads();

		/* ---- unpack: a warp takes the next 32 columns of the sorted list; they are of one class */
		const uint32_t nchunks = sm.base[NCLS] >> 5;
		for (;;) {
			uint32_t ch = 0;
			if (lane == 0)
				ch = atomicAdd(&sm.next_chunk, 1u);
			ch = __shfl_sync(0xFFFFFFFFu, ch, 0);
			if (ch >= nchunks)
				break;
			const uint32_t cls = sm.chunk_cls[ch];
			const uint32_t ent = sm.list[32u * ch + (uint32_t)lane];
			const bool have = ent != 0xFFFFu;
			const uint32_t cid = have ? ent & 2047u : 0u, aux = ent >> 11;
			const uint32_t bi = cid >> 7, col = cid & 127u;
			const uint32_t PP = sm.pos[cid];
			const uint64_t gb = g0 + bi;
			const uint32_t *w = sm.bbase[bi];
			const uint32_t fe = sm.bfe[bi];
			const bool near = (sm.bflags[bi] & BF_NEAR) != 0u;
			const uint32_t iw = PP >> 5;
			uint4 *const dst = reinterpret_cast<uint4 *>(g.inter + gb * (uint64_t)BLEN + (col & 31u) * 64u + (col >> 5) * 16u);
			if (cls <= C_K0) {
				if (have) {
					const uint32_t w0 = ldw(w, iw, fe, near), w1 = ldw(w, iw + 1u, fe, near), w2 = ldw(w, iw + 2u, fe, near),
						       w3 = ldw(w, iw + 3u, fe, near);
					const uint32_t lo = fsr(w0, w1, PP), mid = fsr(w1, w2, PP), hi = fsr(w2, w3, PP);
					uint32_t a0, a1;
					fast2::unpack_k(lo, mid, hi, aux, sm.k8w, a0, a1);
					uint4 v;
					nib_to_bytes(a0, v.x, v.y);
					nib_to_bytes(a1, v.z, v.w);
					*dst = v;
				}
			} else if (cls >= C_T2 && cls <= C_T0) {
				if (have) {
					const uint32_t w0 = ldw(w, iw, fe, near), w1 = ldw(w, iw + 1u, fe, near), w2 = ldw(w, iw + 2u, fe, near);
					const uint32_t lo = fsr(w0, w1, PP), mid = fsr(w1, w2, PP);
					uint32_t a0, a1;
					const int bad = fast2::unpack_t(lo, mid, PP, fe + 8u, (uint32_t)C_T0 - cls, sm.t, a0, a1);
					if (bad)
						atomicMin(g.first_bad + sm.bstream[bi], sm.bno[bi]);
					if (aux & 1u) { /* the block decodes */
						uint4 v;
						nib_to_bytes(a0, v.x, v.y);
						nib_to_bytes(a1, v.z, v.w);
						*dst = v;
					}
				}
			} else if (cls >= C_L7) {
				if (have) {
					const uint32_t w0 = ldw(w, iw, fe, near), w1 = ldw(w, iw + 1u, fe, near), w2 = ldw(w, iw + 2u, fe, near),
						       w3 = ldw(w, iw + 3u, fe, near), w4 = ldw(w, iw + 4u, fe, near);
					const uint32_t W[4] = { fsr(w0, w1, PP), fsr(w1, w2, PP), fsr(w2, w3, PP), fsr(w3, w4, PP) };
					uint4 v;
					switch (cls) { /* uniform */
					case C_L3: v = lin_narrow<3>(W); break;
					case C_L3 - 1: v = lin_narrow<4>(W); break;
					case C_L3 - 2: v = lin_narrow<5>(W); break;
					case C_L3 - 3: v = lin_narrow<6>(W); break;
					default: v = lin_narrow<8 - QS>(W); break;
					}
					*dst = v;
				}
			} else { /* C_WIDE */
				if (have) {
					/* f_linear decode.c:196-206: code - 2^(ind-1), as int16 */
					StreamWords sw;
					sw.w = w;
					sw.fe = fe;
					sw.near = near;
					uint32_t v[ROWS];
					fast2::unpack_linear(sw, PP, aux + 8u, 1, v);
					uint32_t q[8];
#pragma unroll
					for (int j = 0; j < 8; j++)
						q[j] = __byte_perm(v[2 * j], v[2 * j + 1], 0x5410);
					uint4 *wd = reinterpret_cast<uint4 *>(g.wide + (gb * COLS + col) * (uint64_t)ROWS);
					wd[0] = make_uint4(q[0], q[1], q[2], q[3]);
					wd[1] = make_uint4(q[4], q[5], q[6], q[7]);
					atomicOr(&sm.wmask[bi][col >> 5], 1u << (col & 31u));
				}
			}
		}
		__syncthreads();
		if (tid < UB * 4) {
			const uint64_t gb = g0 + (tid >> 2);
			if (gb < g.n_blocks && (sm.bflags[tid >> 2] & BF_OK))
				g.wmask[gb * 4u + (tid & 3)] = sm.wmask[tid >> 2][tid & 3];
		}
	}
}

/* ================================================================== lift */

constexpr int L_WARPS = 8;
constexpr int L_THREADS = 32 * L_WARPS;
constexpr int XPRE = 68;                     /* chunk -1: the previous block's last 64 stage-2 words (+4 pad) */
constexpr int XWORDS = XPRE + BLEN + 4 * 32; /* transpose layout: 4 pad words per 64 */
constexpr int LSTAGE_W = BLEN / 4 + 16;      /* the block's index bytes, its record (8 words) and wide mask (4 words) */
constexpr int LB_WORDS = XWORDS + LSTAGE_W;  /* per warp */
constexpr size_t LIFT_SMEM = (size_t)L_WARPS * LB_WORDS * 4;

static_assert((XWORDS * 4) % 16 == 0 && (LB_WORDS * 4) % 16 == 0, "16-byte aligned staging");

__device__ __forceinline__ uint32_t lift(uint32_t a, uint32_t p1, uint32_t p2, bool odd)
{
	/* decode.c:518-519 */
	const uint32_t s = a + p2;
	return odd ? 2u * p1 - s : 2u * p1 + s;
}

/* two results -> one 32-bit word of 16-bit PCM: bytes 1, 2 of a and of b (OSH = 8) */
__device__ __forceinline__ uint32_t pack2(uint32_t a, uint32_t b, uint32_t sel, uint32_t flip)
{
	return __byte_perm(a, b, sel) ^ flip;
}

/* what a block needs of its predecessor, per lane (the reference's wrapbuf, decode.c:803, in the
 * basis of the flat form): the last four stage-1 inputs, the last two stage-1 and stage-2 outputs */
struct Hist {
	uint32_t hx[4], hy[2], hz[2];
};

/* byte `j` (0..3) of w times val, as a two-way dot product: lo/hi picks the byte pair, the
 * 16-bit operand (val, 0) or (0, val) the byte of the pair.  val is unsigned, the byte signed. */
template <int J>
__device__ __forceinline__ uint32_t byte_mul(uint32_t w, uint32_t v_lo, uint32_t v_hi)
{
	uint32_t d;
	if (J == 0)
		asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(v_lo), "r"(w), "r"(0));
	else if (J == 1)
		asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(v_hi), "r"(w), "r"(0));
	else if (J == 2)
		asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(v_lo), "r"(w), "r"(0));
	else
		asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(v_hi), "r"(w), "r"(0));
	return d;
}

/*
 * One 16-word piece (window words t = 16 K .. 16 K + 15) of lifting stages 3..7.  Stage 3 runs
 * from K = 2, stages 4-6 (S46) from K = 3, stage 7 and the output (OUT) from K = 4; ODD = K & 1 is
 * the row parity of stage 3.  P1 = the stage-3 inputs of piece K - 1, A3..A6 = what the later
 * stages need of piece K - 1; all are replaced by this piece's.  The caller runs the pieces in
 * pairs (even, odd): within a pair nothing is copied and the stage-3 sign is a constant.
 */
template <int ODD, bool S46, bool OUT, bool CKS>
__device__ __forceinline__ void s37_piece(const int K, const uint32_t *win, uint32_t (&P1)[16], uint32_t (&A3)[16],
					  uint32_t (&A4)[8], uint32_t (&A5)[4], uint32_t (&A6)[2], uint4 *dst, bool full,
					  uint32_t n, int lane, uint32_t pos0, uint32_t sel, uint32_t flip, uint32_t bias,
					  unsigned long long &cks)
{
	/* word t of the window lives at t + (t >= 64 ? 4 : 0) */
	const uint4 *p0 = reinterpret_cast<const uint4 *>(win + 16 * K + (K >= 4 ? 4 : 0));
	const uint4 *p2 = reinterpret_cast<const uint4 *>(win + 16 * (K - 2) + (K >= 6 ? 4 : 0));
	uint32_t a3[16], c0[16];
#pragma unroll
	for (int q = 0; q < 4; q++) {
		const uint4 v0 = p0[q], v2 = p2[q];
		c0[4 * q] = v0.x; c0[4 * q + 1] = v0.y; c0[4 * q + 2] = v0.z; c0[4 * q + 3] = v0.w;
		/* C = 16: 2*in[t-16] +- (in[t] + in[t-32]); row parity = K & 1 */
		a3[4 * q + 0] = lift(v0.x, P1[4 * q + 0], v2.x, ODD);
		a3[4 * q + 1] = lift(v0.y, P1[4 * q + 1], v2.y, ODD);
		a3[4 * q + 2] = lift(v0.z, P1[4 * q + 2], v2.z, ODD);
		a3[4 * q + 3] = lift(v0.w, P1[4 * q + 3], v2.w, ODD);
	}
	if (S46) {
		uint32_t a4[16], a5[16], a6[16];
#pragma unroll
		for (int j = 0; j < 16; j++) /* C = 8 */
			a4[j] = lift(a3[j], j >= 8 ? a3[j - 8] : A3[j + 8], A3[j], (j >> 3) & 1);
#pragma unroll
		for (int j = 0; j < 16; j++) /* C = 4 */
			a5[j] = lift(a4[j], j >= 4 ? a4[j - 4] : A4[j + 4], j >= 8 ? a4[j - 8] : A4[j], (j >> 2) & 1);
#pragma unroll
		for (int j = 0; j < 16; j++) /* C = 2 */
			a6[j] = lift(a5[j], j >= 2 ? a5[j - 2] : A5[j + 2], j >= 4 ? a5[j - 4] : A5[j], (j >> 1) & 1);
		if (OUT) {
			uint32_t pk[8];
#pragma unroll
			for (int j = 0; j < 16; j += 2) { /* C = 1 */
				const uint32_t v0 = lift(a6[j], j >= 1 ? a6[j - 1] : A6[1], j >= 2 ? a6[j - 2] : A6[0], 0);
				const uint32_t v1 = lift(a6[j + 1], a6[j], j >= 1 ? a6[j - 1] : A6[1], 1);
				pk[j >> 1] = pack2(v0, v1, sel, flip);
				if (CKS) {
					/* u_i as an unsigned 16-bit value, independent of byte order */
					const uint32_t m = pos0 + 64u * lane + 16u * (uint32_t)(K - 4) + (uint32_t)j;
					const uint32_t w0 = ((v0 >> OSH) + bias) & 0xFFFFu;
					const uint32_t w1 = ((v1 >> OSH) + bias) & 0xFFFFu;
					if (m - pos0 < n)
						cks += (unsigned long long)(m + 1u) * (w0 + 1ull);
					if (m + 1u - pos0 < n)
						cks += (unsigned long long)(m + 2u) * (w1 + 1ull);
				}
			}
			const int q = 2 * (K - 4);
			if (full) {
				__stcs(dst + q, make_uint4(pk[0], pk[1], pk[2], pk[3]));
				__stcs(dst + q + 1, make_uint4(pk[4], pk[5], pk[6], pk[7]));
			} else {
				/* last block of a stream: word-granular tail */
				uint16_t *d16 = reinterpret_cast<uint16_t *>(dst + q);
#pragma unroll 1
				for (int e = 0; e < 16; e++) {
					const uint32_t m = 64u * lane + 8u * q + e;
					const uint32_t w = e < 8 ? (e < 4 ? (e < 2 ? pk[0] : pk[1]) : (e < 6 ? pk[2] : pk[3]))
								 : (e < 12 ? (e < 10 ? pk[4] : pk[5]) : (e < 14 ? pk[6] : pk[7]));
					if (m < n)
						d16[e] = (uint16_t)(w >> (16 * (e & 1)));
				}
			}
		}
#pragma unroll
		for (int j = 0; j < 8; j++)
			A4[j] = a4[8 + j];
#pragma unroll
		for (int j = 0; j < 4; j++)
			A5[j] = a5[12 + j];
		A6[0] = a6[14];
		A6[1] = a6[15];
	}
#pragma unroll
	for (int j = 0; j < 16; j++) {
		A3[j] = a3[j];
		P1[j] = c0[j];
	}
}

/* stage 1 (C = 64) output i of the lane: m-64 -> i-2, m-128 -> i-4; row parity = (i >> 1) & 1; the
 * "+1" of decode.c:561-564 (m % 64 == 0: lane 0, even i) rides in the first add */
__device__ __forceinline__ uint32_t stage1(uint32_t xi, uint32_t p1, uint32_t p2, int i, uint32_t one0)
{
	const bool odd = (i >> 1) & 1;
	const uint32_t sum = (i & 1) ? xi + p2 : odd ? xi + p2 - one0 : xi + p2 + one0;
	return odd ? 2u * p1 - sum : 2u * p1 + sum;
}

/*
 * Transform + output of one block by one warp.  x[i] = the dequantised word m = 32 i + lane of the
 * block (row i / 4, column 32 (i % 4) + lane); xs0 = the warp's transpose buffer; h = the history
 * left by the previous block, replaced by this block's.  n = words to emit (<= 2048; 0: history
 * only).  Returns this lane's checksum contribution.
 */
template <bool CKS>
__device__ __forceinline__ unsigned long long juggle_and_store(uint32_t (&x)[64], Hist &h, uint32_t *xs0, int lane,
							       uint8_t *out, uint32_t pos0, uint32_t n, const Format fmt)
{
	uint32_t *xs = xs0 + XPRE;
	unsigned long long cks = 0ull;
	const uint32_t one0 = lane == 0 ? 1u << QS : 0u;
	uint32_t y[64];
#pragma unroll
	for (int i = 0; i < 64; i++) {
		const uint32_t p1 = i >= 2 ? x[i - 2] : h.hx[i + 2];
		const uint32_t p2 = i >= 4 ? x[i - 4] : h.hx[i];
		y[i] = stage1(x[i], p1, p2, i, one0);
	}
#pragma unroll
	for (int k = 0; k < 4; k++)
		h.hx[k] = x[60 + k];
	/* chunk -1 = the previous block's last 64 stage-2 words */
	xs[-XPRE + lane] = h.hz[0];
	xs[-XPRE + 32 + lane] = h.hz[1];
#pragma unroll
	for (int i = 0; i < 64; i++) {
		/* C = 32: m-32 -> i-1, m-64 -> i-2; row parity = i & 1 */
		const uint32_t p1 = i >= 1 ? y[i - 1] : h.hy[1];
		const uint32_t p2 = i >= 2 ? y[i - 2] : h.hy[i];
		const uint32_t z = lift(y[i], p1, p2, i & 1);
		/* transpose layout: word m lives at m + 4 * (m / 64); m / 64 = i / 2 for every lane */
		xs[32 * i + lane + 4 * (i >> 1)] = z;
		if (i == 62)
			h.hz[0] = z;
		if (i == 63)
			h.hz[1] = z;
	}
	h.hy[0] = y[62];
	h.hy[1] = y[63];
	__syncwarp();

	/* ---- stages 3..7: lane owns m in [64 lane, 64 lane + 64) and walks the 128-word window made
	 * of the 64 words before it (the halo: the previous lane's chunk, chunk -1 for lane 0) and its
	 * own, in 16-word pieces k = 2..7.  A stage is only run where its inputs are complete: stage 3
	 * from t = 32, stages 4-6 from t = 48, stage 7 (= the output) from t = 64. */
	const uint32_t sel = fmt.be ? 0x5612u : 0x6521u;
	const uint32_t flip = fmt.bias ? (fmt.be ? 0x00800080u : 0x80008000u) : 0u;
	const bool full = (uint32_t)(64 * lane + 64) <= n;
	uint4 *dst = reinterpret_cast<uint4 *>(out + ((size_t)pos0 + 64u * lane) * 2u);
	const uint32_t *win = xs + 68 * (lane - 1);
	if (n) {
		uint32_t A3[16], A4[8], A5[4], A6[2], P1[16];
#pragma unroll
		for (int j = 0; j < 16; j++)
			A3[j] = 0u;
#pragma unroll
		for (int j = 0; j < 8; j++)
			A4[j] = 0u;
		A5[0] = A5[1] = A5[2] = A5[3] = 0u;
		A6[0] = A6[1] = 0u;
#pragma unroll
		for (int q = 0; q < 4; q++) {
			const uint4 v = reinterpret_cast<const uint4 *>(win + 16)[q];
			P1[4 * q] = v.x; P1[4 * q + 1] = v.y; P1[4 * q + 2] = v.z; P1[4 * q + 3] = v.w;
		}
		s37_piece<0, false, false, CKS>(2, win, P1, A3, A4, A5, A6, dst, full, n, lane, pos0, sel, flip, fmt.bias, cks);
		s37_piece<1, true, false, CKS>(3, win, P1, A3, A4, A5, A6, dst, full, n, lane, pos0, sel, flip, fmt.bias, cks);
#pragma unroll 1
		for (int k = 4; k < 8; k += 2) {
			s37_piece<0, true, true, CKS>(k, win, P1, A3, A4, A5, A6, dst, full, n, lane, pos0, sel, flip, fmt.bias, cks);
			s37_piece<1, true, true, CKS>(k + 1, win, P1, A3, A4, A5, A6, dst, full, n, lane, pos0, sel, flip, fmt.bias, cks);
		}
	}
	__syncwarp(); /* all shared-memory reads of this block are done */
	return cks;
}

/* the lane's 64 dequantised words of a block from its staged index bytes (column p = 32 p + lane
 * is the p-th 16 bytes: row r is byte r), val = the block's multiplier; the columns flagged in wm
 * come from the side array as int16 */
__device__ __forceinline__ void dequant(uint32_t (&x)[64], const uint4 *bytes, uint32_t val, const uint4 wm,
					const uint16_t *wide_blk, int lane)
{
	const uint32_t v_lo = val & 0xFFFFu, v_hi = val << 16;
#pragma unroll
	for (int p = 0; p < 4; p++) {
		const uint4 q = bytes[p];
		const uint32_t w[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
		for (int j = 0; j < 4; j++) {
			x[4 * (4 * j + 0) + p] = byte_mul<0>(w[j], v_lo, v_hi);
			x[4 * (4 * j + 1) + p] = byte_mul<1>(w[j], v_lo, v_hi);
			x[4 * (4 * j + 2) + p] = byte_mul<2>(w[j], v_lo, v_hi);
			x[4 * (4 * j + 3) + p] = byte_mul<3>(w[j], v_lo, v_hi);
		}
	}
	const uint32_t any = wm.x | wm.y | wm.z | wm.w;
	if (any) { /* uniform */
		const uint32_t vq = val << QS;
#pragma unroll
		for (int p = 0; p < 4; p++) {
			const uint32_t m = p == 0 ? wm.x : p == 1 ? wm.y : p == 2 ? wm.z : wm.w;
			if (m) { /* uniform: a real branch per pass with a wide column */
				const bool mine = (m >> lane) & 1u;
				const uint4 *src = reinterpret_cast<const uint4 *>(wide_blk + (size_t)(32 * p + lane) * ROWS);
				uint4 q0 = make_uint4(0u, 0u, 0u, 0u), q1 = q0;
				if (mine) {
					q0 = __ldcg(src);
					q1 = __ldcg(src + 1);
				}
				const uint32_t w[8] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w };
#pragma unroll
				for (int j = 0; j < 8; j++) {
					const uint32_t e0 = (uint32_t)((int32_t)(int16_t)(w[j] & 0xFFFFu)) * vq;
					const uint32_t e1 = (uint32_t)((int32_t)w[j] >> 16) * vq;
					x[4 * (2 * j) + p] = mine ? e0 : x[4 * (2 * j) + p];
					x[4 * (2 * j + 1) + p] = mine ? e1 : x[4 * (2 * j + 1) + p];
				}
			}
		}
	}
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *gp)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gp) : "memory");
}

/* asks for block gb's index bytes (every lane its own 64), record and wide mask */
__device__ __forceinline__ void stage_block(const SplitArgs &g, uint64_t gb, uint32_t *stg, int lane)
{
	const uint32_t sa = (uint32_t)__cvta_generic_to_shared(stg);
	const uint8_t *src = g.inter + gb * (uint64_t)BLEN + (uint32_t)lane * 64u;
#pragma unroll
	for (int p = 0; p < 4; p++)
		cp_async16(sa + (uint32_t)lane * 64u + 16u * p, src + 16 * p);
	if (lane < 2)
		cp_async16(sa + BLEN + 16u * lane, reinterpret_cast<const uint8_t *>(g.rec + gb) + 16 * lane);
	if (lane == 2)
		cp_async16(sa + BLEN + 32u, g.wmask + gb * 4u);
	cp_async_commit();
}

template <bool CKS>
__global__ void __launch_bounds__(L_THREADS, 2) acm_lift_kernel(KernelArgs a, SplitArgs g)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint32_t *wb = reinterpret_cast<uint32_t *>(smem_raw) + (size_t)warp * LB_WORDS;
	uint32_t *stg = wb + XWORDS;

	for (;;) {
		uint32_t it = 0;
		if (lane == 0)
			it = atomicAdd(g.item_counter, 1u);
		it = __shfl_sync(0xFFFFFFFFu, it, 0);
		if (it >= g.n_items)
			break;
		const Gen2Item item = g.items[it];
		const DevStream d = a.streams[item.stream];
		const Gen2Stream gs = g.gs[item.stream];
		const uint32_t nscan = g.nscan[item.stream];
		uint32_t bend = item.b0 + item.nb;
		bend = bend < nscan ? bend : nscan;
		if (item.b0 >= bend)
			continue;
		uint8_t *out = a.out + d.out_off;
		Hist h;
#pragma unroll
		for (int k = 0; k < 4; k++)
			h.hx[k] = 0u;
		h.hy[0] = h.hy[1] = h.hz[0] = h.hz[1] = 0u; /* decode.c:812: a stream starts from zero history */
		__syncwarp();
		if (item.b0) {
			/* rebuild the history from the previous block's last two rows of indices (it was walked
			 * and unpacked: a stream's records end at its first failing block) */
			const uint64_t gp = gs.rec_base + item.b0 - 1u;
			stage_block(g, gp, stg, lane);
			cp_async_wait_all();
			__syncwarp();
			uint32_t x[64];
			const uint4 wm = *reinterpret_cast<const uint4 *>(stg + BLEN / 4 + 8);
			dequant(x, reinterpret_cast<const uint4 *>(stg) + 4 * lane, stg[BLEN / 4 + 2], wm,
				g.wide + gp * (uint64_t)BLEN, lane);
			const uint32_t one0 = lane == 0 ? 1u << QS : 0u;
			uint32_t y[4];
#pragma unroll
			for (int i = 60; i < 64; i++)
				y[i - 60] = stage1(x[i], x[i - 2], x[i - 4], i, one0);
			h.hz[0] = lift(y[2], y[1], y[0], 0);
			h.hz[1] = lift(y[3], y[2], y[1], 1);
			h.hy[0] = y[2];
			h.hy[1] = y[3];
#pragma unroll
			for (int k = 0; k < 4; k++)
				h.hx[k] = x[60 + k];
			__syncwarp();
		}
		stage_block(g, gs.rec_base + item.b0, stg, lane);
		for (uint32_t b = item.b0; b < bend; b++) {
			const uint64_t gb = gs.rec_base + b;
			cp_async_wait_all();
			__syncwarp();
			const int status = (int)stg[BLEN / 4 + 3];
			if (status != SCAN_OK)
				break; /* the stream's last record: nothing to deliver */
			uint32_t x[64];
			{
				const uint4 wm = *reinterpret_cast<const uint4 *>(stg + BLEN / 4 + 8);
				dequant(x, reinterpret_cast<const uint4 *>(stg) + 4 * lane, stg[BLEN / 4 + 2], wm,
					g.wide + gb * (uint64_t)BLEN, lane);
			}
			__syncwarp(); /* the staged bytes have been read */
			if (b + 1u < bend)
				stage_block(g, gb + 1u, stg, lane); /* travels while this block is transformed */
			const uint64_t pos = (uint64_t)b * BLEN;
			const uint32_t n = pos < d.words_limit ? (d.words_limit - pos < (uint64_t)BLEN ? (uint32_t)(d.words_limit - pos) : (uint32_t)BLEN) : 0u;
			unsigned long long c2 = juggle_and_store<CKS>(x, h, wb, lane, out, (uint32_t)pos, n, a.fmt);
			if (CKS) {
				for (int o = 16; o; o >>= 1)
					c2 += __shfl_xor_sync(0xFFFFFFFFu, c2, o);
				if (lane == 0)
					g.cks_blk[gb] = c2;
			}
		}
		cp_async_wait_all();
		__syncwarp();
	}
}

} // namespace split

/* ================================================================== launch */

bool split_shape(uint32_t level, uint32_t rows) { return level == split::LEVEL && rows == (uint32_t)split::ROWS; }

size_t split_bytes_per_block()
{
	return sizeof(BlockRec) + 8 + 2 * split::COLS + split::BLEN + 2 * split::BLEN + 16;
}

static cudaError_t split_configure()
{
	static bool configured[64] = {};
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess)
		return e;
	if (!configured[dev & 63]) {
		e = cudaFuncSetAttribute(split::acm_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
					 (int)sizeof(walk::SmemWalk));
		if (e == cudaSuccess)
			e = cudaFuncSetAttribute(split::acm_walk1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
						 (int)split::W1_SMEM);
		if (e == cudaSuccess)
			e = cudaFuncSetAttribute(split::acm_lift_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
						 (int)split::LIFT_SMEM);
		if (e == cudaSuccess)
			e = cudaFuncSetAttribute(split::acm_lift_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
						 (int)split::LIFT_SMEM);
		if (e != cudaSuccess)
			return e;
		configured[dev & 63] = true;
	}
	return cudaSuccess;
}

/* unpack of the records [u0, u0 + n) of g's arrays */
static void launch_unpack(const KernelArgs &a, const SplitArgs &g, uint64_t u0, uint64_t n, int sms, cudaStream_t st)
{
	SplitArgs gu = g;
	gu.rec += u0;
	gu.coff16 += u0 * split::COLS;
	gu.inter += u0 * split::BLEN;
	gu.wide += u0 * split::BLEN;
	gu.wmask += u0 * 4u;
	gu.n_blocks = n;
	const uint64_t tiles = (n + split::UB - 1) / split::UB;
	const uint32_t ugrid = (uint32_t)(tiles < (uint64_t)sms * 5u ? tiles : (uint64_t)sms * 5u);
	if (ugrid)
		split::acm_unpack_kernel<<<ugrid, split::U_THREADS, 0, st>>>(a, gu);
}

static void launch_lift(const KernelArgs &a, const SplitArgs &g, int sms, cudaStream_t st)
{
	if (!g.n_items)
		return;
	const uint32_t want = (g.n_items + split::L_WARPS - 1) / split::L_WARPS;
	const uint32_t lgrid = want < 2u * (uint32_t)sms ? want : 2u * (uint32_t)sms;
	if (a.fmt.checksums)
		split::acm_lift_kernel<true><<<lgrid, split::L_THREADS, split::LIFT_SMEM, st>>>(a, g);
	else
		split::acm_lift_kernel<false><<<lgrid, split::L_THREADS, split::LIFT_SMEM, st>>>(a, g);
}

/*
 * One stream (a.count == 1, g.gs[0].rec_base == 0): walk blocks [b0, b0 + nb) from bit position P0,
 * unpack them (which also finds out-of-range radix codes: g.first_bad) and, if lift is set,
 * transform them: g.items / g.n_items = the lift work items of exactly these blocks,
 * g.item_counter zeroed by the caller.  The block before b0 is unpacked again (its index bytes
 * rebuild the transform history).
 */
cudaError_t launch_split_range(const KernelArgs &a, const SplitArgs &g, uint32_t b0, uint32_t nb, uint32_t P0,
			       int lift, int sms, cudaStream_t st)
{
	cudaError_t e = split_configure();
	if (e != cudaSuccess)
		return e;
	split::acm_walk1_kernel<<<1, split::W1_THREADS, split::W1_SMEM, st>>>(a, g, b0, nb, P0);
	const uint32_t u0 = b0 ? b0 - 1u : 0u;
	launch_unpack(a, g, u0, (uint64_t)b0 + nb - u0, sms, st);
	if (lift)
		launch_lift(a, g, sms, st);
	return cudaGetLastError();
}

cudaError_t launch_split(const KernelArgs &a, const SplitArgs &g, int sms, cudaStream_t st)
{
	if (a.count == 0)
		return cudaSuccess;
	cudaError_t e = split_configure();
	if (e != cudaSuccess)
		return e;
	/* walk: one stream per lane; few streams are spread thin (the walk is latency bound: a warp
	 * that has its sub-partition to itself steps faster), many fill 8 warps per SM */
	const uint32_t warps = (a.count + 31u) / 32u;
	uint32_t nw = (warps + (uint32_t)sms - 1u) / (uint32_t)sms;
	nw = nw < 1u ? 1u : nw > (uint32_t)walk::SW ? (uint32_t)walk::SW : nw;
	uint32_t ctas = (warps + nw - 1u) / nw;
	ctas = ctas > (uint32_t)sms ? (uint32_t)sms : ctas;
	split::acm_walk_kernel<<<ctas, 32 * nw, sizeof(walk::SmemWalk), st>>>(a, g);
	launch_unpack(a, g, 0, g.n_blocks, sms, st);
	launch_lift(a, g, sms, st);
	return launch_gen2_finish(a, g, st);
}

} // namespace acm
