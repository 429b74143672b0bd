/*
 * acm_fast.cu -- the throughput kernel for the common block shape: level 7
 * (128 columns), 16 rows, 2048 words per block (the shape of BASELINE configs 1, 2, 4).
 *
 * Why it looks the way it does (DESIGN.md section 4 has the numbers):
 *
 *  - A stream's bitstream is serial: where column c+1 starts is only known once column
 *    c has been walked.  Only the LENGTH walk is serial, though.  One
 *    "scan" warp per CTA runs it with one stream per lane (S streams in flight per
 *    CTA) as a flat, divergence-free state machine: every iteration each lane either
 *    reads a column selector (and, for prefix codes, the first table step out of the
 *    same 32-bit peek, via the 13-bit sel13 table) or takes one more multi-symbol table
 *    step (kstep table, row cap folded in).  Each lane sees its stream through a 96-bit
 *    shift register fed from a private 256-byte shared-memory ring; the ring is topped up
 *    by all lanes AT THE SAME TIME every 8 iterations with 128-bit global loads whose data
 *    is only stored to the ring at the following top-up, so no instruction of the walk
 *    ever waits on global memory (register scoreboards are per warp: with 32 independent
 *    streams, any load issued at a lane-specific moment stalls all lanes -- a TMA ring
 *    with per-lane mbarriers and a register look-ahead were both tried and measured,
 *    profiles/r01_ncu_v4_summary.md, r01_ncu_v7_summary.md).  The EOF rule (one zero byte,
 *    decode.c:57-61) is applied while storing to the ring.  The walk is a
 *    dependent chain (~450 steps per block), so a CTA runs TWO scan warps = 64 stream
 *    slots to cover the latency with streams.  The 128 column offsets of each block are
 *    published through shared memory.
 *  - Everything else is parallel inside a block.  W "worker" warps each own S/W of the
 *    CTA's stream slots.  Per block a worker warp
 *      stage    copies the block's compressed bytes (<= 4.2 KB, known from the scan)
 *               into shared memory with coalesced 128-bit loads;
 *      unpack   lane = column (4 passes of 32): filler dispatch, table decode, store of
 *               the 16-bit index X0[row*128+col] (two lanes per 32-bit word, no bank
 *               conflict);
 *      juggle   dequantise (idx*val) on load; stages 1-2 (C=64,32) in registers: lane j
 *               owns every word m = j mod 32, which is exactly what it just unpacked;
 *               one transpose through shared memory to contiguous ownership (lane j
 *               owns m in [64j,64j+64)) and stages 3-7 (C=16..1) in registers over a
 *               62-word halo that is recomputed instead of exchanged;
 *      output   >>7, low 16 bits, byte order / sign bias folded into one PRMT (+LOP),
 *               eight 128-bit stores per lane: the block leaves as 4 KiB of PCM.
 *    The reference's wrapbuf (decode.c:803, 2*cols-2 = 254 words) becomes 256 words of
 *    per-slot history (last 128 X0, 64 X1, 64 X2 words) kept in an L2-resident global
 *    array, which lets any worker warp take any slot (dynamic slot queue per round).
 *  - Scan and workers are double buffered: in round r the scan warp walks block r of
 *    every slot while the workers decode block r-1; one __syncthreads per round.
 *  - Streams are handed out by an atomic cursor in longest-first order; a slot that
 *    finishes its stream takes the next one, so mixed lengths do not idle lanes.
 *
 * Bit-exactness: same arithmetic as the generic kernel (uint32 wrap-around, arithmetic
 * shift, truncation), same table-driven symbol decode, same status rules.
 */
#include "acm_kernels.cuh"

namespace acm {

namespace fast {

constexpr int LEVEL = 7;
constexpr int COLS = 128;
constexpr int ROWS = 16;
constexpr int BLEN = COLS * ROWS;        /* 2048 */
constexpr int NSCAN = 2;                 /* scan warps */
constexpr int W = 14;                    /* worker warps; 16 warps = 512 threads -> 128 regs/thread */
constexpr int S = 32 * NSCAN;            /* stream slots per CTA */
constexpr int THREADS = 32 * (W + NSCAN);
constexpr int OFF_PITCH = S + 2;         /* u16 per column row of the offset table (bank spread) */
constexpr int XPRE = 68;                 /* chunk -1: the previous block's last 64 X2 words (+4 pad) */
constexpr int XWORDS = XPRE + BLEN + 4 * 32; /* transpose layout: 4 pad words per 64 */
constexpr int X0_WORDS = BLEN / 2;       /* int16 indices */
constexpr int STAGE_BYTES = (BLEN + 4 * 32 - X0_WORDS) * 4; /* 4608: a whole block (<= 4179 B) + slack */
constexpr int STAGE_CHUNKS = STAGE_BYTES / 16;
constexpr int RING_WORDS = 64;           /* 16 chunks of 16 bytes per scan lane */
constexpr int RING_LEAD = 8;             /* chunks kept ahead of the read position */
constexpr int SCAN_UNROLL = 4;           /* unrolled walk steps (keeps the loop body inside the L0 I-cache) */
constexpr int SCAN_PERIOD = 8;           /* walk steps between two ring top-ups (<= 32 bytes consumed) */
constexpr int HIST_WORDS = 256;          /* per slot: X0 tail [0,128) X1 tail [128,192) X2 tail [192,256) */

enum { ENT_IDLE = -100 };

struct Entry {
	uint32_t pblock; /* P of the block header */
	uint32_t pend;   /* P where the scan stopped (block end when status == SCAN_OK) */
	uint32_t desc;   /* index into the kernel's descriptor slice */
	uint32_t blk;    /* block number, bit 31 = last attempt of the stream */
	int32_t status;  /* SCAN_OK / SCAN_EOF / ACM_ERR_* / ENT_IDLE */
	uint32_t ncols;
	int32_t val;
	uint32_t pad;
};

struct Smem {
	uint64_t k8[ACM_K8_SIZE];
	uint16_t sel13[8192];
	uint8_t kstep[8 * 8 * 256];
	uint32_t x[W][XWORDS];       /* per worker: chunk -1 | int16 X0 + staged bytes, later transposed X2 */
	uint32_t ring[S][RING_WORDS]; /* per scan lane: compressed window */
	uint16_t coloff[2][COLS * OFF_PITCH];
	Entry ent[2][S];
	unsigned long long cks[S];
	uint32_t pos[S];
	uint32_t dead[2][S];         /* [round parity][slot]: desc+1 of a stream a worker finalised early */
	uint32_t info[32];
	uint16_t t[ACM_T_SIZE];
	int more[2][NSCAN];
	uint32_t next_slot[2];
};

/* ------------------------------------------------------------------ PTX helpers */

__device__ __forceinline__ uint4 ldg_nc_v4(const void *p)
{
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
		     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
		     : "l"(p));
	return r;
}
__device__ __forceinline__ void prefetch_l1(const void *p)
{
	asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

/* ------------------------------------------------------------------ bit readers */

/*
 * Scan-lane reader: a 96-bit shift register (lo, mid, hi) over a private shared-memory
 * ring.  `lo` always holds the next 32 stream bits, so a table lookup needs no funnel
 * shift by the bit position; consuming `step` <= 32 bits is three funnel shifts, only the
 * first of which (lo) is on the walk's dependent chain.  At least 64 bits are valid at
 * the top of every iteration; when fewer remain, ring word `widx` (read unconditionally
 * at the top of the iteration, next to the table lookups) is spliced in above them.
 *
 * topup() runs for all lanes together every SCAN_PERIOD iterations: it stores the (at most
 * two) 16-byte chunks requested by the previous top-up into the ring -- zeroing everything
 * at and past the end of the file, which is the reference's "one zero byte, then nothing"
 * (decode.c:57-61) -- and requests the next ones so that RING_LEAD chunks stay ahead of
 * the read position.  SCAN_PERIOD iterations consume at most 32 bytes = the two chunks a
 * top-up can add, so the ring never runs dry; the ~1500+ cycles between two top-ups cover a
 * scattered load that misses to DRAM, so nothing waits on a load either.
 */
struct ScanReader {
	const uint4 *base16;  /* stream base (16-byte aligned) */
	uint32_t room16;      /* 16-byte chunks readable at base16 */
	uint32_t fe_word, fe_tail;
	uint32_t *ring;       /* this lane's RING_WORDS words of shared memory */
	uint32_t lo, mid, hi, avail; /* bits [0, avail) of hi:mid:lo are the stream at the read position */
	uint32_t widx;        /* stream word that will be spliced in next */
	uint32_t fill;        /* chunks stored so far = index of the next chunk to store */
	uint32_t npend;       /* chunks requested at the last top-up (0..2) */
	uint4 pa, pb;

	__device__ __forceinline__ uint4 load_chunk(uint32_t c) const
	{
		uint4 v = make_uint4(0u, 0u, 0u, 0u);
		if (c < room16)
			v = ldg_nc_v4(base16 + c);
		return v;
	}
	__device__ __forceinline__ void store_chunk(uint32_t c, uint4 v)
	{
		if (4u * c + 3u >= fe_word) { /* touches the end of the file: rare */
			uint32_t q[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
			for (int j = 0; j < 4; j++) {
				const uint32_t k = 4u * c + j;
				if (k > fe_word || (k == fe_word && !fe_tail))
					q[j] = 0u;
				else if (k == fe_word)
					q[j] &= (1u << fe_tail) - 1u;
			}
			v = make_uint4(q[0], q[1], q[2], q[3]);
		}
		reinterpret_cast<uint4 *>(ring)[c & (RING_WORDS / 4 - 1)] = v;
	}
	__device__ __forceinline__ void reset(uint32_t *r)
	{
		ring = r;
		base16 = nullptr;
		room16 = 0;
		fe_word = fe_tail = 0;
		lo = mid = hi = widx = 0;
		avail = 96;
		fill = 0x40000000u; /* never asks for data */
		npend = 0;
		pa = pb = make_uint4(0u, 0u, 0u, 0u);
	}
	/* new stream: fill the ring synchronously once, position the window on bit P0 */
	__device__ __forceinline__ void start(const uint8_t *src, uint64_t room, uint32_t file_end, uint32_t P0)
	{
		base16 = reinterpret_cast<const uint4 *>(src);
		room16 = (uint32_t)(room >> 4);
		fe_word = file_end >> 5;
		fe_tail = file_end & 31u;
		const uint32_t i = P0 >> 5, sh = P0 & 31u, c0 = i >> 2;
#pragma unroll
		for (int j = 0; j < RING_LEAD; j++)
			store_chunk(c0 + j, load_chunk(c0 + j));
		fill = c0 + RING_LEAD;
		npend = 0;
		const uint32_t a = ring[i & (RING_WORDS - 1)], b = ring[(i + 1) & (RING_WORDS - 1)],
			       c = ring[(i + 2) & (RING_WORDS - 1)];
		lo = __funnelshift_r(a, b, sh);
		mid = __funnelshift_r(b, c, sh);
		hi = c >> sh;
		avail = 96u - sh;
		widx = i + 3;
	}
	__device__ __forceinline__ void topup()
	{
		if (npend >= 1)
			store_chunk(fill, pa);
		if (npend >= 2)
			store_chunk(fill + 1, pb);
		fill += npend;
		const int want = (int)((widx >> 2) + RING_LEAD) - (int)fill;
		npend = want <= 0 ? 0u : (want >= 2 ? 2u : 1u);
		if (npend >= 1)
			pa = load_chunk(fill);
		if (npend >= 2)
			pb = load_chunk(fill + 1);
	}
	__device__ __forceinline__ uint32_t peek() const { return lo; }
	__device__ __forceinline__ uint32_t next_word() const { return ring[widx & (RING_WORDS - 1)]; }
	/* consume step <= 32 bits; cand = next_word() read earlier in the iteration */
	__device__ __forceinline__ void consume(uint32_t step, uint32_t cand)
	{
		lo = __funnelshift_rc(lo, mid, step);
		mid = __funnelshift_rc(mid, hi, step);
		hi = __funnelshift_rc(hi, 0u, step);
		avail -= step;
		const bool refill = avail <= 64u; /* then avail is in (32, 64] and hi holds nothing */
		const unsigned long long t = (unsigned long long)cand << ((avail - 32u) & 63u);
		mid = refill ? (mid | (uint32_t)t) : mid;
		hi = refill ? (uint32_t)(t >> 32) : hi;
		avail = refill ? avail + 32u : avail;
		widx = refill ? widx + 1u : widx;
	}
};

/* Worker view of the staged block: stream words [w_lo, ...) in shared memory, EOF already
 * patched in at staging time.  No bounds check: the staging area (STAGE_BYTES = 4608) always
 * holds the whole block (<= 33 428 bits) plus the 96 bits the decoders read ahead, and for a
 * block that failed in the scan the walk stopped at most 261 bits past the stream's limit. */
struct StageReader {
	const uint32_t *st;
	uint32_t w_lo;
	__device__ __forceinline__ uint32_t word(uint32_t i) const { return st[i - w_lo]; }
};

/* ------------------------------------------------------------------ scan */

/*
 * Per-selector facts for this block shape (16 rows), one 32-bit word each:
 *   bit 16 prefix-coded (k) filler, bit 17 bad selector, bit 18 t filler, bit 19 linear
 *   bits 20..23 sub-type: k8 table number / t table number
 */
enum { INF_K = 1u << 16, INF_BAD = 1u << 17, INF_T = 1u << 18, INF_LIN = 1u << 19 };

__device__ __forceinline__ uint32_t make_info(uint32_t kind)
{
	const uint32_t cls = kind & 7u, sub = kind >> 3;
	uint32_t v = sub << 20;
	if (cls == ACM_CLS_LINEAR)
		v |= INF_LIN;
	else if (cls == ACM_CLS_T)
		v |= INF_T;
	else if (cls == ACM_CLS_K)
		v |= INF_K;
	else if (cls == ACM_CLS_BAD)
		v |= INF_BAD;
	return v;
}

/*
 * Flat walk over one block for every lane of a scan warp at once (replaces the
 * per-column loop nest of scan_block for this shape; same verdicts).  Per iteration a
 * lane is in one of three states:
 *   at a selector (rem == 0, pend == 0): one sel13 lookup gives the whole advance of a
 *       fixed-size column, or the selector plus the first prefix-code step and its table;
 *   inside a prefix-coded column (rem rows to come): one kstep lookup (row cap folded in);
 *   inside a fixed-size payload (pend bits to come): skip, at most 32 bits per iteration
 *       (the shift-register window consumes at most one word per step).
 * The body is straight-line code (every update is a select on the lane's state, including
 * the reference's end-of-file and corruption verdicts), unrolled SCAN_PERIOD times between
 * two ring top-ups, and the loop condition is a warp vote: the 32 lanes execute ONE
 * instruction stream however their column types differ.
 */
enum { SCAN_RUNNING = 2 }; /* besides SCAN_OK (1), SCAN_EOF (0) and the negative error codes */

struct ScanState {
	uint32_t P, rem, pend, col, kbase;
	int st; /* SCAN_RUNNING while walking, then the verdict */
};

/* NEAR = false: no lane of the warp can reach its limit within this period (a step is <= 32
 * bits), so the end-of-file verdicts are compiled out; a bad selector is still caught. */
template <bool NEAR>
__device__ __forceinline__ void scan_step(ScanReader &br, ScanState &s, uint32_t limit, uint16_t *&cp,
					  uint32_t pblock, const uint16_t *sel13, const uint8_t *kstep)
{
	const uint32_t w = br.peek();
	const uint32_t cand = br.next_word();
	const bool at_sel = (s.rem | s.pend) == 0u;
	const uint32_t es = sel13[w & 0x1FFFu];
	const uint32_t ek = kstep[s.kbase + umin32(s.rem, 7u) * 256u + (w & 255u)];
	const uint32_t adv_s = es & 511u, hi7 = es >> 9, rem_s = hi7 & 15u;
	const bool run = s.st == SCAN_RUNNING;
	/* selector: GET_BITS_EXPECT_EOF decode.c:496, then f_bad decode.c:190-194 */
	const bool sel_try = run && at_sel;
	const bool eof = NEAR && sel_try && (s.P + 5u > limit);
	const bool bad = sel_try && hi7 == 0x70u;
	s.st = eof ? (int)SCAN_EOF : (bad ? -6 : s.st);
	const bool go = s.st == SCAN_RUNNING;
	const bool sel = go && at_sel;
	if (sel)
		*cp = (uint16_t)(s.P - pblock);
	cp = sel ? cp + OFF_PITCH : cp;
	s.col = sel ? s.col + 1u : s.col;
	s.kbase = sel ? (es >> 13) * 2048u : s.kbase;
	const bool is_k = at_sel ? rem_s != 0u : s.rem != 0u;
	const uint32_t tot = at_sel ? adv_s : s.pend; /* fixed-size payload still to skip */
	uint32_t step = is_k ? (at_sel ? adv_s : (ek & 15u)) : umin32(tot, 32u);
	step = go ? step : 0u;
	s.pend = go ? (is_k ? 0u : tot - step) : s.pend;
	s.rem = go ? (at_sel ? rem_s : s.rem - (ek >> 4)) : s.rem; /* kstep[.][0][.] = 0 */
	s.P += step;
	br.consume(step, cand);
	const bool fin = go && (s.rem | s.pend) == 0u;
	const bool over = NEAR && fin && s.P > limit; /* a GET_BITS inside the payload ran dry: decode.c:146-152 */
	s.col = over ? s.col - 1u : s.col;    /* its selector was consumed, its payload did not complete */
	s.st = over ? -7 : ((fin && s.col == (uint32_t)COLS) ? (int)SCAN_OK : s.st);
}

__device__ __forceinline__ ScanResult scan_block_flat(ScanReader &br, uint32_t P, uint32_t limit,
						      uint16_t *coloff, const uint16_t *sel13,
						      const uint8_t *kstep, bool active)
{
	ScanState s;
	ScanResult r;
	uint16_t *cp = coloff;
	s.P = P;
	s.rem = s.pend = s.col = s.kbase = 0;
	s.st = active ? (int)SCAN_RUNNING : (int)SCAN_OK;
	r.val = 0;
	if (active) {
		if (P + 20 > limit) { /* pwr(4) / val(16): GET_BITS_EXPECT_EOF decode.c:588-589 */
			s.st = SCAN_EOF;
		} else {
			r.val = (int)((br.peek() >> 4) & 0xFFFFu);
			s.P = P + 20;
			br.consume(20u, br.next_word());
		}
	}
	while (__any_sync(0xFFFFFFFFu, s.st == SCAN_RUNNING)) {
		br.topup();
		/* SCAN_PERIOD steps move at most 32*SCAN_PERIOD bits; +13 for the lookup window */
		const bool near = s.st == SCAN_RUNNING && s.P + 32u * SCAN_PERIOD + 16u > limit;
		if (__any_sync(0xFFFFFFFFu, near)) {
#pragma unroll 1
			for (int h = 0; h < SCAN_PERIOD / SCAN_UNROLL; h++) {
#pragma unroll
				for (int k = 0; k < SCAN_UNROLL; k++)
					scan_step<true>(br, s, limit, cp, P, sel13, kstep);
			}
		} else {
#pragma unroll 1
			for (int h = 0; h < SCAN_PERIOD / SCAN_UNROLL; h++) {
#pragma unroll
				for (int k = 0; k < SCAN_UNROLL; k++)
					scan_step<false>(br, s, limit, cp, P, sel13, kstep);
			}
		}
	}
	r.status = s.st;
	r.ncols = s.col;
	r.end = s.P;
	return r;
}

/* ------------------------------------------------------------------ unpack */

/*
 * Decode one column (lane = column) into 16-bit indices x0c[r*128], r < 16.  One
 * straight-line routine per filler class.  Prefix- and radix-coded columns only produce
 * values in -5..5, so their 16 rows are accumulated as 4-bit fields in two registers and
 * stored once at the end (the first version spent 60 % of its step loop on predicated
 * per-value stores).  Returns non-zero if a t-code that the reference gets to read is out
 * of range (decode.c:412/:438/:464).
 */
__device__ __forceinline__ int unpack_column(const StageReader &sr, uint32_t P, uint32_t limit, uint32_t ind,
					     uint32_t inf, int16_t *x0c, const uint64_t *k8,
					     const uint16_t *tt)
{
	int bad = 0;
	if (inf & (INF_K | INF_T)) {
		uint32_t a0 = 0u, a1 = 0u; /* rows 0-7 / 8-15, one nibble each */
		if (inf & INF_K) {
			/* prefix codes (decode.c:208-403): up to 7 values per table step */
			const uint64_t *tab = k8 + (inf >> 20) * 256u;
			uint32_t i = P >> 5, lo = sr.word(i), hi = sr.word(i + 1), r = 0;
			while (r < (uint32_t)ROWS) {
				const uint32_t ni = P >> 5;
				if (ni != i) { /* a step consumes <= 8 bits: at most one word further */
					lo = hi;
					hi = sr.word(ni + 1);
					i = ni;
				}
				const uint64_t e64 = tab[__funnelshift_r(lo, hi, P & 31u) & 255u];
				const uint32_t e = (uint32_t)e64, hv = (uint32_t)(e64 >> 32);
				const uint32_t k = umin32(e & 15u, (uint32_t)ROWS - r);
				P += (e >> (4u * k)) & 15u;
				const unsigned long long vv = (unsigned long long)(hv & ((1u << (4u * k)) - 1u)) << (4u * r);
				a0 |= (uint32_t)vv;
				a1 |= (uint32_t)(vv >> 32);
				r += k;
			}
		} else {
			/* f_t15 / f_t27 / f_t37 (decode.c:405-476): all codes sit in one 64-bit window */
			const uint32_t sub = inf >> 20;
			const uint32_t width = sub == 0 ? 5u : 7u, per = sub == 2 ? 2u : 3u;
			const uint32_t ncodes = sub == 2 ? 8u : 6u, cmask = (1u << width) - 1u;
			const uint32_t vmask = sub == 2 ? 0xFFu : 0xFFFu;
			const uint16_t *tab = tt + sub * 128u;
			const uint32_t i = P >> 5, sh = P & 31u;
			const uint32_t w0 = sr.word(i), w1 = sr.word(i + 1), w2 = sr.word(i + 2);
			const unsigned long long win =
				(unsigned long long)__funnelshift_r(w0, w1, sh) |
				((unsigned long long)__funnelshift_r(w1, w2, sh) << 32);
			/* codes the reference gets to read before the stream runs dry: all of them,
			 * except in the last block of a truncated stream */
			const uint32_t nread = P + ncodes * width <= limit ? ncodes : (limit > P ? (limit - P) / width : 0u);
			uint32_t seen = 0u;
#pragma unroll
			for (int q = 0; q < 8; q++) {
				if ((uint32_t)q < ncodes) {
					const uint32_t e = tab[(uint32_t)(win >> (q * width)) & cmask];
					seen |= (uint32_t)q < nread ? e : 0u;
					const unsigned long long vv = (unsigned long long)(e & vmask) << (4u * q * per);
					a0 |= (uint32_t)vv;
					a1 |= (uint32_t)(vv >> 32);
				}
			}
			bad = (seen & 0x8000u) != 0u;
		}
#pragma unroll
		for (int r = 0; r < ROWS; r++)
			x0c[r * COLS] = (int16_t)nib_s(r < 8 ? a0 : a1, r & 7);
	} else if (inf & INF_LIN) {
		/* f_linear (decode.c:196-206): sliding 64-bit window, branch-free refill */
		const uint32_t mask = (1u << ind) - 1u;
		const int mid = 1 << (ind - 1);
		const uint32_t i = P >> 5, sh = P & 31u;
		unsigned long long win = (((unsigned long long)sr.word(i + 1) << 32) | sr.word(i)) >> sh;
		uint32_t avail = 64u - sh, nx = i + 2;
#pragma unroll
		for (int r = 0; r < ROWS; r++) {
			x0c[r * COLS] = (int16_t)((int)((uint32_t)win & mask) - mid);
			win >>= ind;
			avail -= ind;
			if (avail <= 32u) {
				win |= (unsigned long long)sr.word(nx) << avail;
				avail += 32u;
				nx++;
			}
		}
	} else {
		/* f_zero decode.c:181-188 */
#pragma unroll
		for (int r = 0; r < ROWS; r++)
			x0c[r * COLS] = 0;
	}
	return bad;
}

/* ------------------------------------------------------------------ transform + output */

__device__ __forceinline__ uint32_t lift(uint32_t a, uint32_t p1, uint32_t p2, bool odd)
{
	/* decode.c:518-519 */
	uint32_t s = a + p2;
	return odd ? 2u * p1 - s : 2u * p1 + s;
}

/* pack two results into one 32-bit word of 16-bit PCM: (v >> 7) low 16 bits each */
__device__ __forceinline__ uint32_t pack2(uint32_t a, uint32_t b, uint32_t sel, uint32_t flip)
{
	uint32_t lo = (uint32_t)((int32_t)a >> LEVEL); /* bytes 0,1 wanted */
	uint32_t hi = b << (16 - LEVEL);               /* bytes 2,3 wanted */
	return __byte_perm(lo, hi, sel) ^ flip;
}

/*
 * Transform + output of one block by one warp.  xs[0..1024) holds the 16-bit indices
 * X0[row*128+col]; val is the block's multiplier; gh is the slot's history in global
 * memory (first == true: all-zero history, decode.c:812).  n = words to emit (<= 2048).
 * Returns this lane's checksum contribution.
 */
template <bool CKS>
__device__ __forceinline__ unsigned long long
juggle_and_store(uint32_t *xs, uint32_t *gh, bool first, int lane, int val, uint8_t *out, uint32_t pos0,
		 uint32_t n, const Format fmt)
{
	const int16_t *x0 = reinterpret_cast<const int16_t *>(xs);
	uint32_t x[64];
	unsigned long long cks = 0ull;

	/* history of the previous block (L2 resident; .cg: never a stale L1 line) */
	uint32_t hx[4], hy[2], hz[2];
#pragma unroll
	for (int k = 0; k < 4; k++)
		hx[k] = first ? 0u : __ldcg(gh + 32 * k + lane);       /* X0[-128 + 32k + lane] */
	hy[0] = first ? 0u : __ldcg(gh + 128 + lane);               /* X1[-64 + lane] */
	hy[1] = first ? 0u : __ldcg(gh + 160 + lane);               /* X1[-32 + lane] */
	hz[0] = first ? 0u : __ldcg(gh + 192 + lane);               /* X2[-64 + lane] */
	hz[1] = first ? 0u : __ldcg(gh + 224 + lane);               /* X2[-32 + lane] */

	/* ---- dequantise (decode.c:174-177, :591-600) and stages 1, 2 in registers:
	 * lane owns m = 32*i + lane */
#pragma unroll
	for (int i = 0; i < 64; i++)
		x[i] = (uint32_t)((int)x0[32 * i + lane] * val);
#pragma unroll
	for (int k = 0; k < 4; k++)
		__stcg(gh + 32 * k + lane, x[60 + k]);
	const uint32_t one0 = lane == 0 ? 1u : 0u; /* decode.c:561-564: +1 where m % 64 == 0 */
	uint32_t y[64];
#pragma unroll
	for (int i = 0; i < 64; i++) {
		/* C = 64: m-64 -> i-2, m-128 -> i-4; row parity = (m/64)&1 = (i>>1)&1 */
		uint32_t p1 = i >= 2 ? x[i - 2] : hx[i + 2];
		uint32_t p2 = i >= 4 ? x[i - 4] : hx[i];
		y[i] = lift(x[i], p1, p2, (i >> 1) & 1);
		if ((i & 1) == 0)
			y[i] += one0;
	}
	__stcg(gh + 128 + lane, y[62]);
	__stcg(gh + 160 + lane, y[63]);
	__syncwarp(); /* every lane has read its X0 indices: the buffer can be overwritten */
	/* chunk -1 = the previous block's last 64 X2 words */
	xs[-XPRE + lane] = hz[0];
	xs[-XPRE + 32 + lane] = hz[1];
#pragma unroll
	for (int i = 0; i < 64; i++) {
		/* C = 32: m-32 -> i-1, m-64 -> i-2; row parity = i&1 */
		uint32_t p1 = i >= 1 ? y[i - 1] : hy[1];
		uint32_t p2 = i >= 2 ? y[i - 2] : hy[i];
		uint32_t z = lift(y[i], p1, p2, i & 1);
		/* transpose layout: word m lives at m + 4*(m/64); m/64 = i/2 for every lane */
		xs[32 * i + lane + 4 * (i >> 1)] = z;
		if (i == 62)
			__stcg(gh + 192 + lane, z);
		if (i == 63)
			__stcg(gh + 224 + lane, z);
	}
	__syncwarp();

	/* ---- stages 3..7 in registers: lane owns m in [64*lane, 64*lane+64), halo = the 64
	 * words before it (previous lane's chunk; chunk -1 for lane 0: 68*(lane-1) = -XPRE) */
	const uint4 *own = reinterpret_cast<const uint4 *>(xs + 68 * lane);
	const uint4 *prev = reinterpret_cast<const uint4 *>(xs + 68 * (lane - 1));
	uint32_t u[128], a3[128], a4[128], a5[128], a6[128], a7[2];
	const uint32_t sel = fmt.be ? 0x6701u : 0x7610u;
	const uint32_t flip = fmt.bias ? (fmt.be ? 0x00800080u : 0x80008000u) : 0u;
	const bool full = (uint32_t)(64 * lane + 64) <= n;
	uint4 *dst = reinterpret_cast<uint4 *>(out + ((size_t)pos0 + 64u * lane) * 2u);
	uint32_t pk[4];
#pragma unroll
	for (int t = 0; t < 36; t += 4) {
		uint4 q = prev[t >> 2];
		u[t] = q.x; u[t + 1] = q.y; u[t + 2] = q.z; u[t + 3] = q.w;
	}
#pragma unroll
	for (int t = 34; t < 128; t++) {
		if ((t & 3) == 0 && t >= 36) {
			/* just-in-time halo / chunk load keeps ~70 words live instead of 128 */
			uint4 q = t < 64 ? prev[t >> 2] : own[(t - 64) >> 2];
			u[t] = q.x; u[t + 1] = q.y; u[t + 2] = q.z; u[t + 3] = q.w;
		}
		a3[t] = lift(u[t], u[t - 16], u[t - 32], (t >> 4) & 1);          /* C = 16 */
		if (t >= 50)
			a4[t] = lift(a3[t], a3[t - 8], a3[t - 16], (t >> 3) & 1); /* C = 8 */
		if (t >= 58)
			a5[t] = lift(a4[t], a4[t - 4], a4[t - 8], (t >> 2) & 1);  /* C = 4 */
		if (t >= 62)
			a6[t] = lift(a5[t], a5[t - 2], a5[t - 4], (t >> 1) & 1);  /* C = 2 */
		if (t >= 64) {
			a7[t & 1] = lift(a6[t], a6[t - 1], a6[t - 2], t & 1);     /* C = 1 */
			if (t & 1) {
				const int w = ((t - 64) >> 1) & 3;
				pk[w] = pack2(a7[0], a7[1], sel, flip);
				if (CKS) {
					/* u_i as an unsigned 16-bit value, independent of byte order */
					uint32_t m = pos0 + 64u * lane + (uint32_t)(t - 64);
					uint32_t w0 = (((uint32_t)((int32_t)a7[0] >> LEVEL)) + fmt.bias) & 0xFFFFu;
					uint32_t w1 = (((uint32_t)((int32_t)a7[1] >> LEVEL)) + fmt.bias) & 0xFFFFu;
					if (m - 1u - pos0 < n)
						cks += (unsigned long long)m * (w0 + 1ull);
					if (m - pos0 < n)
						cks += (unsigned long long)(m + 1u) * (w1 + 1ull);
				}
				if (w == 3) {
					const int q = (t - 64) >> 3;
					if (full) {
						/* streaming store: PCM is written once and must not evict the
						 * L2-resident history and the compressed bytes still to be staged */
						__stcs(dst + q, make_uint4(pk[0], pk[1], pk[2], pk[3]));
					} else {
						/* last block of a stream: word-granular tail */
						uint16_t *d16 = reinterpret_cast<uint16_t *>(dst + q);
#pragma unroll
						for (int e = 0; e < 8; e++) {
							uint32_t m = 64u * lane + 8u * q + e;
							if (m < n)
								d16[e] = (uint16_t)(pk[e >> 1] >> (16 * (e & 1)));
						}
					}
				}
			}
		}
	}
	__syncwarp(); /* all shared-memory reads of this block are done */
	return cks;
}

/* ------------------------------------------------------------------ kernel */

template <bool CKS>
__global__ void __launch_bounds__(THREADS, 1) acm_decode_fast_kernel(KernelArgs a)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

	for (int i = tid; i < ACM_K8_SIZE; i += THREADS)
		sm.k8[i] = a.tables->k8[i];
	for (int i = tid; i < 8192; i += THREADS)
		sm.sel13[i] = a.tables->sel13_r16[i];
	for (int i = tid; i < 8 * 8 * 256 / 4; i += THREADS)
		reinterpret_cast<uint32_t *>(sm.kstep)[i] = reinterpret_cast<const uint32_t *>(a.tables->kstep)[i];
	for (int i = tid; i < ACM_T_SIZE; i += THREADS)
		sm.t[i] = a.tables->t[i];
	if (tid < 32)
		sm.info[tid] = make_info(a.tables->kind[tid]);
	for (int i = tid; i < S; i += THREADS) {
		sm.dead[0][i] = 0;
		sm.dead[1][i] = 0;
		sm.pos[i] = 0;
		sm.cks[i] = 0ull;
	}
	if (tid < 2 * NSCAN)
		(&sm.more[0][0])[tid] = 0;
	if (tid < 2)
		sm.next_slot[tid] = 0;
	__syncthreads();

	/* scan-lane state (meaningful in the scan warps only) */
	const bool is_scan = warp >= W;
	const int myslot = is_scan ? 32 * (warp - W) + lane : 0;
	bool active = false;
	uint32_t cur = 0, P = 0, blk = 0, limit = 0, n_attempt = 0;
	ScanReader sbr;
	sbr.reset(sm.ring[myslot]);
	uint32_t *const cta_hist = a.hist + (size_t)blockIdx.x * S * HIST_WORDS;

	for (int round = 0;; round++) {
		const int buf = round & 1;
		if (is_scan) {
			/* ================= scan warps: lane = stream slot ================= */
			Entry e;
			e.status = ENT_IDLE;
			e.pblock = 0; e.pend = 0; e.desc = 0; e.blk = 0; e.ncols = 0; e.val = 0; e.pad = 0;
			if (warp == W && lane == 0)
				sm.next_slot[buf] = 0; /* the workers drain this queue next round */
			/* written by a worker in the previous round (other parity: no concurrent writer) */
			if (active && sm.dead[buf ^ 1][myslot] == cur + 1u)
				active = false; /* a worker found a corrupt t-code: abandon the stream */
			if (!active) {
				uint32_t idx = atomicAdd(a.counter, 1u);
				if (idx < a.count) {
					const DevStream d = a.streams[idx];
					cur = idx;
					P = d.bit0;
					blk = 0;
					limit = d.file_end + 8u;
					n_attempt = d.n_attempt;
					sbr.start(a.blob + d.base_off, a.blob_room > d.base_off ? a.blob_room - d.base_off : 0,
						  d.file_end, P);
					active = true;
				}
			}
			bool walk = false;
			if (active) {
				e.desc = cur;
				e.pblock = P;
				e.blk = blk;
				if (blk >= n_attempt) {
					/* nothing (more) to attempt: clean end */
					e.status = SCAN_EOF;
					e.pend = P;
					e.blk |= 0x80000000u;
					active = false;
				} else {
					walk = true;
				}
			}
			{
				ScanResult sc = scan_block_flat(sbr, P, limit, sm.coloff[buf] + myslot, sm.sel13,
								sm.kstep, walk);
				if (walk) {
					e.status = sc.status;
					e.ncols = sc.ncols;
					e.val = sc.val;
					e.pend = sc.end;
					P = sc.end;
					blk++;
					if (sc.status != SCAN_OK || blk >= n_attempt) {
						e.blk |= 0x80000000u;
						active = false;
					}
				}
			}
			sm.ent[buf][myslot] = e;
			const int any = __any_sync(0xFFFFFFFFu, e.status != ENT_IDLE);
			if (lane == 0)
				sm.more[buf][warp - W] = any;
		} else if (round > 0) {
			/* ================= worker warps: take slots from the round's queue ================= */
			const int pb = buf ^ 1;
			uint32_t *xs = sm.x[warp] + XPRE;
			int16_t *x0 = reinterpret_cast<int16_t *>(xs);
			uint32_t *stage = xs + X0_WORDS;
			for (;;) {
				int slot = 0;
				if (lane == 0)
					slot = (int)atomicAdd(&sm.next_slot[pb], 1u);
				slot = __shfl_sync(0xFFFFFFFFu, slot, 0);
				if (slot >= S)
					break;
				const Entry e = sm.ent[pb][slot];
				if (e.status == ENT_IDLE)
					continue;
				const DevStream d = a.streams[e.desc];
				const uint32_t bno = e.blk & 0x7FFFFFFFu;
				const bool last = (e.blk >> 31) != 0;
				if (bno == 0) {
					if (lane == 0) {
						sm.pos[slot] = 0u;
						sm.cks[slot] = 0ull;
					}
					__syncwarp();
				}
				if (sm.dead[0][slot] == e.desc + 1u || sm.dead[1][slot] == e.desc + 1u)
					continue; /* stream already finalised by a corrupt code (descriptor ids are unique) */
				const uint32_t limit_w = d.file_end + 8u;
				const bool ok = e.status == SCAN_OK;
				const uint32_t ncheck = ok ? (uint32_t)COLS : e.ncols + (e.status == -7 ? 1u : 0u);
				int bad = 0;
				if (ncheck) {
					/* ---- stage the block's bytes: 16-byte chunks [c_lo, c_hi) of the stream,
					 * with the EOF rule applied (bits at and past file_end read as zero) */
					const uint32_t c_lo = e.pblock >> 7;
					uint32_t c_hi = (e.pend + 96u + 127u) >> 7;
					if (c_hi > c_lo + (uint32_t)STAGE_CHUNKS)
						c_hi = c_lo + (uint32_t)STAGE_CHUNKS;
					const uint8_t *src = a.blob + d.base_off;
					const uint64_t room = a.blob_room > d.base_off ? a.blob_room - d.base_off : 0;
					const uint32_t fe_word = d.file_end >> 5, fe_tail = d.file_end & 31u;
					for (uint32_t c = c_lo + lane; c < c_hi; c += 32) {
						uint4 v = make_uint4(0u, 0u, 0u, 0u);
						if ((uint64_t)c * 16u + 16u <= room)
							v = ldg_nc_v4(src + (size_t)c * 16u);
						if (4u * c + 3u >= fe_word) {
							uint32_t q[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
							for (int j = 0; j < 4; j++) {
								const uint32_t k = 4u * c + j;
								if (k > fe_word || (k == fe_word && !fe_tail))
									q[j] = 0u;
								else if (k == fe_word)
									q[j] &= (1u << fe_tail) - 1u;
							}
							v = make_uint4(q[0], q[1], q[2], q[3]);
						}
						reinterpret_cast<uint4 *>(stage)[c - c_lo] = v;
					}
					__syncwarp();
					StageReader sr;
					sr.st = stage;
					sr.w_lo = c_lo * 4u;
					/* ---- unpack: lane = column */
					const uint16_t *offs = sm.coloff[pb] + slot;
#pragma unroll 1
					for (int p = 0; p < 4; p++) {
						const uint32_t c = 32u * p + lane;
						if (c < ncheck) {
							const uint32_t Pc = e.pblock + offs[c * OFF_PITCH];
							const uint32_t i = Pc >> 5;
							const uint32_t ind =
								__funnelshift_r(sr.word(i), sr.word(i + 1), Pc & 31u) & 31u;
							bad |= unpack_column(sr, Pc + 5u, limit_w, ind, sm.info[ind], x0 + c, sm.k8,
									     sm.t);
						}
					}
				}
				bad = __any_sync(0xFFFFFFFFu, bad);
				__syncwarp();
				uint32_t pos = sm.pos[slot];
				int st = 0;
				if (bad)
					st = -6;
				else if (!ok)
					st = e.status == SCAN_EOF ? 0 : e.status;
				if (ok && !bad) {
					uint32_t n = d.words_limit - pos;
					if (n > (uint32_t)BLEN)
						n = BLEN;
					uint8_t *out = a.out + d.out_off;
					unsigned long long c2 = juggle_and_store<CKS>(xs, cta_hist + slot * HIST_WORDS, bno == 0,
										       lane, e.val, out, pos, n, a.fmt);
					pos += n;
					if (CKS) {
						for (int o = 16; o; o >>= 1)
							c2 += __shfl_xor_sync(0xFFFFFFFFu, c2, o);
						if (lane == 0)
							sm.cks[slot] += c2;
					}
					if (lane == 0)
						sm.pos[slot] = pos;
				}
				if (!ok || bad || last) {
					/* finalise: results + zero padding of the undelivered tail */
					uint8_t *p0 = a.out + d.out_off + (size_t)pos * a.fmt.wordlen;
					/* up to the 16-byte boundary that ends this stream's slot (out_off is 16-byte
					 * aligned), so that alignment gaps never carry stale bytes */
					size_t nb = d.pad_words >= pos && d.pad_words
							    ? (((size_t)d.pad_words * a.fmt.wordlen + 15u) & ~(size_t)15u) -
								      (size_t)pos * a.fmt.wordlen
							    : 0;
					for (size_t i = lane; i < nb; i += 32)
						p0[i] = 0;
					__syncwarp();
					if (lane == 0) {
						a.status[d.index] = st;
						a.words[d.index] = pos;
						a.cks[d.index] = a.fmt.checksums ? sm.cks[slot] : 0ull;
						sm.dead[buf][slot] = e.desc + 1u;
					}
				}
				__syncwarp();
			}
		}
		__syncthreads();
		int more = 0;
#pragma unroll
		for (int k = 0; k < NSCAN; k++)
			more |= sm.more[buf][k];
		if (!more)
			break; /* the scan warps produced nothing this round: all streams are done */
	}
}

} // namespace fast

bool fast_shape(uint32_t level, uint32_t rows) { return level == fast::LEVEL && rows == fast::ROWS; }

size_t fast_smem_bytes() { return sizeof(fast::Smem); }

int fast_slots_per_cta() { return fast::S; }

size_t fast_hist_words_per_cta() { return (size_t)fast::S * fast::HIST_WORDS; }

cudaError_t launch_fast(const KernelArgs &a, int n_ctas, cudaStream_t st)
{
	if (a.count == 0)
		return cudaSuccess;
	/* the opt-in shared-memory size is a per-device function attribute */
	static bool configured[2][64] = {};
	const size_t smem = sizeof(fast::Smem);
	const int v = a.fmt.checksums ? 1 : 0;
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess)
		return e;
	if (!configured[v][dev & 63]) {
		e = v ? cudaFuncSetAttribute(fast::acm_decode_fast_kernel<true>,
					     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
		      : cudaFuncSetAttribute(fast::acm_decode_fast_kernel<false>,
					     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess)
			return e;
		configured[v][dev & 63] = true;
	}
	if (v)
		fast::acm_decode_fast_kernel<true><<<n_ctas, fast::THREADS, smem, st>>>(a);
	else
		fast::acm_decode_fast_kernel<false><<<n_ctas, fast::THREADS, smem, st>>>(a);
	return cudaGetLastError();
}

} // namespace acm
