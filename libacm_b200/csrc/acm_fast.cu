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
 *    same 32-bit peek) or takes one more multi-symbol table step.  Each lane reads its
 *    stream through a 1 KiB shared-memory ring that is refilled 256 bytes at a time by
 *    TMA bulk copies (cp.async.bulk + mbarrier), two quarters ahead of the read
 *    position, so the walk never waits on HBM.  The 128 column offsets of each block
 *    are published through shared memory.
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
 *    The reference's wrapbuf (decode.c:803, 2*cols-2 = 254 words) becomes per-slot
 *    history in shared memory: last 128 X0 words, last 64 X1 words, last 64 X2 words.
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
constexpr int BLEN = COLS * ROWS;      /* 2048 */
constexpr int W = 15;                  /* worker warps (+1 scan warp = 512 threads, 128 regs) */
constexpr int S = 2 * W;               /* stream slots per CTA = active lanes of the scan warp */
constexpr int THREADS = 32 * (W + 1);
constexpr int SLOTS_PER_WORKER = S / W;
constexpr int OFF_PITCH = 33;          /* u16 per column row of the offset table (bank spread) */
constexpr int XWORDS = BLEN + 4 * 32;  /* transpose layout: 4 pad words per 64 */
constexpr int X0_BYTES = BLEN * 2;     /* int16 indices */
constexpr int STAGE_BYTES = XWORDS * 4 - X0_BYTES; /* 4608: a whole block (<= 4179 B) + slack */
constexpr int STAGE_CHUNKS = STAGE_BYTES / 16;
constexpr int RING_WORDS = 256;        /* per-slot compressed window: 4 quarters of 256 B */
constexpr int QWORDS = 64;
constexpr uint32_t SPIN_LIMIT = 1u << 24;

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
	uint32_t x[W][XWORDS];       /* per worker: int16 X0 + staged bytes, later transposed X2 */
	uint32_t ring[S][RING_WORDS];
	uint32_t hist0[S][128];      /* last 128 X0 words: [k*32+lane] = x[60+k] of that lane */
	uint32_t hist1[S][64];       /* last 64 X1 words: [k*32+lane] = y[62+k] */
	uint32_t hist2[S][64];       /* last 64 X2 words, flat order */
	unsigned long long bar[S][4];
	unsigned long long cks[S];
	Entry ent[2][S];
	uint32_t pos[S];
	uint32_t dead[S];
	uint16_t coloff[2][COLS * OFF_PITCH];
	uint32_t info[32];
	uint16_t t[ACM_T_SIZE];
	uint8_t kind[32];
	int more[2];
};

/* ------------------------------------------------------------------ PTX helpers */

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
		     : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, uint32_t parity)
{
	uint32_t ok;
	asm volatile("{\n\t.reg .pred p;\n\t"
		     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		     "selp.u32 %0, 1, 0, p;\n\t}"
		     : "=r"(ok)
		     : "r"(smem_u32(bar)), "r"(parity)
		     : "memory");
	return ok != 0;
}
/* TMA 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP) */
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
			     smem_u32(dst)),
		     "l"(src), "r"(bytes), "r"(smem_u32(bar))
		     : "memory");
}
__device__ __forceinline__ uint4 ldg_nc_v4(const void *p)
{
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
		     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
		     : "l"(p));
	return r;
}

/* ------------------------------------------------------------------ bit readers */

/*
 * Scan-lane reader: the stream is seen through a 4-quarter shared-memory ring.  Quarter
 * numbers G are global per lane (they keep counting across streams) so that ring slot
 * G&3 and mbarrier parity (G>>2)&1 stay consistent.  The ring always holds the quarter
 * before the read position and up to two ahead.
 */
struct RingReader {
	const uint32_t *ring;
	unsigned long long *bars;
	const uint8_t *src;  /* stream base in global memory (16-byte aligned) */
	uint64_t room;       /* bytes readable at src */
	uint32_t file_end;
	int32_t g0;          /* G of the current stream's quarter 0 */
	int32_t g_rd;        /* highest G waited for */
	int32_t g_is;        /* next G to issue */
	uint32_t widx, w0, w1;
	uint32_t *errflag;

	__device__ __forceinline__ void issue(int32_t G)
	{
		const uint64_t off = (uint64_t)(uint32_t)(G - g0) * (QWORDS * 4);
		uint32_t bytes = 0;
		if (off < room)
			bytes = room - off < (uint64_t)(QWORDS * 4) ? (uint32_t)(room - off) : (uint32_t)(QWORDS * 4);
		unsigned long long *bar = bars + (G & 3);
		/* order this lane's earlier generic-proxy reads of the slot before the async write */
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		mbar_expect_tx(bar, bytes);
		if (bytes)
			bulk_g2s((void *)(ring + (G & 3) * QWORDS), src + off, bytes, bar);
	}
	__device__ __forceinline__ void wait(int32_t G)
	{
		unsigned long long *bar = bars + (G & 3);
		const uint32_t parity = ((uint32_t)G >> 2) & 1u;
		uint32_t spins = 0;
		while (!mbar_try_wait(bar, parity)) {
			if (++spins > SPIN_LIMIT) {
				atomicExch(errflag, 1u);
				break;
			}
		}
	}
	__device__ __forceinline__ void drain()
	{
		while (g_rd + 1 < g_is) {
			g_rd++;
			wait(g_rd);
		}
	}
	__device__ __forceinline__ void reset()
	{
		g0 = 0;
		g_rd = -1;
		g_is = 0;
		widx = 0xFFFFFFF0u;
		w0 = w1 = 0;
		src = nullptr;
		room = 0;
		file_end = 0;
	}
	__device__ __forceinline__ void start(const uint8_t *s, uint64_t r, uint32_t fe)
	{
		drain();             /* nothing of the previous stream may still be landing */
		g0 = g_is;
		g_rd = g_is - 1;
		src = s;
		room = r;
		file_end = fe;
		widx = 0xFFFFFFF0u;
		issue(g_is++);
		issue(g_is++);
		issue(g_is++);
	}
	/* make word i readable: wait for its quarter, keep two quarters in flight behind it */
	__device__ __forceinline__ void ensure(uint32_t i)
	{
		const int32_t G = g0 + (int32_t)(i / QWORDS);
		while (g_rd < G) {
			g_rd++;
			wait(g_rd);
			while (g_is <= g_rd + 2)
				issue(g_is++);
		}
	}
	__device__ __forceinline__ uint32_t word(uint32_t i) const
	{
		const uint32_t last = file_end >> 5, tail = file_end & 31u;
		if (i > last || (i == last && !tail))
			return 0u; /* decode.c:57-61: one zero byte, then nothing */
		uint32_t v = ring[(((uint32_t)g0 + i / QWORDS) & 3u) * QWORDS + (i % QWORDS)];
		if (i == last)
			v &= (1u << tail) - 1u;
		return v;
	}
	__device__ __forceinline__ uint32_t peek(uint32_t P)
	{
		const uint32_t i = P >> 5, s = P & 31u;
		if (i != widx) {
			ensure(i + 1);
			if (i == widx + 1) {
				w0 = w1;
				w1 = word(i + 1);
			} else {
				w0 = word(i);
				w1 = word(i + 1);
			}
			widx = i;
		}
		return __funnelshift_r(w0, w1, s);
	}
};

/* Worker reader: the block's bytes staged in shared memory, words [w_lo, w_lo + n). */
struct StageReader {
	const uint32_t *st;
	uint32_t w_lo, n, file_end;
	uint32_t widx, w0, w1;

	__device__ __forceinline__ void init(const uint32_t *s, uint32_t lo, uint32_t cnt, uint32_t fe)
	{
		st = s;
		w_lo = lo;
		n = cnt;
		file_end = fe;
		widx = 0xFFFFFFF0u;
		w0 = w1 = 0;
	}
	__device__ __forceinline__ uint32_t word(uint32_t i) const
	{
		const uint32_t last = file_end >> 5, tail = file_end & 31u;
		if (i > last || (i == last && !tail) || i - w_lo >= n)
			return 0u;
		uint32_t v = st[i - w_lo];
		if (i == last)
			v &= (1u << tail) - 1u;
		return v;
	}
	__device__ __forceinline__ uint32_t peek(uint32_t P)
	{
		const uint32_t i = P >> 5, s = P & 31u;
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
		return __funnelshift_r(w0, w1, s);
	}
};

/* ------------------------------------------------------------------ scan */

/*
 * Per-selector facts for this block shape (16 rows), one 32-bit word each:
 *   bits 0..15  payload bits of a fixed-size filler (zero 0, linear 16*ind, t15 6*5,
 *               t27 6*7, t37 8*7)
 *   bit 16 prefix-coded (k) filler, bit 17 bad selector, bit 18 t filler, bit 19 linear
 *   bits 20..23 sub-type: k8 table number / t table number
 */
enum { INF_K = 1u << 16, INF_BAD = 1u << 17, INF_T = 1u << 18, INF_LIN = 1u << 19 };

__device__ __forceinline__ uint32_t make_info(uint32_t ind, uint32_t kind)
{
	const uint32_t cls = kind & 7u, sub = kind >> 3;
	uint32_t v = sub << 20;
	if (cls == ACM_CLS_LINEAR)
		v |= INF_LIN | ((uint32_t)ROWS * ind);
	else if (cls == ACM_CLS_T)
		v |= INF_T | (sub == 0 ? 30u : (sub == 1 ? 42u : 56u));
	else if (cls == ACM_CLS_K)
		v |= INF_K;
	else if (cls == ACM_CLS_BAD)
		v |= INF_BAD;
	return v;
}

/*
 * Flat walk over one block for every lane of the scan warp at once (replaces the
 * per-column loop nest of scan_block for this shape; same verdicts).  Each iteration a
 * lane is either AT A SELECTOR (rem == 0) or INSIDE a prefix-coded column (rem rows
 * still to come).  The body is written with selects instead of branches and the loop
 * condition is a warp vote, so the 32 lanes execute ONE instruction stream however
 * their column types differ (the first version let the compiler rebuild nested loops:
 * 4-8 active lanes per instruction, profiles/r01_ncu_v2_summary.md).
 */
__device__ __forceinline__ ScanResult scan_block_flat(RingReader &br, uint32_t P, uint32_t limit,
						      uint16_t *coloff, const uint32_t *info,
						      const uint64_t *k8, bool active)
{
	const uint32_t *k8lo = reinterpret_cast<const uint32_t *>(k8); /* low halves: nv + cum */
	const uint32_t pblock = P;
	ScanResult s;
	s.status = SCAN_OK;
	s.ncols = 0;
	s.val = 0;
	bool done = !active;
	uint32_t col = 0, rem = 0, ksub = 0;
	if (!done) {
		if (P + 20 > limit) { /* pwr(4) / val(16): GET_BITS_EXPECT_EOF decode.c:588-589 */
			s.status = SCAN_EOF;
			done = true;
		} else {
			s.val = (int)((br.peek(P) >> 4) & 0xFFFFu);
			P += 20;
		}
	}
	while (__any_sync(0xFFFFFFFFu, !done)) {
		if (!done) {
			const uint32_t w = br.peek(P);
			const bool at_sel = rem == 0;
			const uint32_t inf = info[w & 31u];
			/* selector: GET_BITS_EXPECT_EOF decode.c:496; f_bad decode.c:190-194 */
			const bool sel_eof = at_sel && (P + 5u > limit);
			const bool sel_bad = at_sel && (inf & INF_BAD) != 0u;
			if (at_sel && !sel_eof)
				coloff[col * OFF_PITCH] = (uint16_t)(P - pblock);
			const bool isk = at_sel ? (inf & INF_K) != 0u : true;
			ksub = at_sel ? (inf >> 20) * 256u : ksub;
			/* the selector's own peek already holds the first 8 payload bits */
			const uint32_t sym = (at_sel ? (w >> 5) : w) & 255u;
			const uint32_t e = k8lo[2u * (ksub + sym)];
			const uint32_t remk = at_sel ? (uint32_t)ROWS : rem;
			const uint32_t kk = umin32(e & 15u, remk);
			const uint32_t klen = (e >> (4u * kk)) & 15u;
			P += (at_sel ? 5u : 0u) + (isk ? klen : (inf & 0xFFFFu));
			rem = isk ? remk - kk : 0u;
			if (sel_eof) {
				s.status = SCAN_EOF;
				done = true;
			} else if (sel_bad) {
				s.status = -6;
				done = true;
			} else if (rem == 0u) {
				if (P > limit) { /* a GET_BITS inside the payload ran dry: decode.c:146-152 */
					s.status = -7;
					done = true;
				} else {
					col++;
					done = col == (uint32_t)COLS;
				}
			}
		}
	}
	s.ncols = col;
	s.end = P;
	return s;
}

/* ------------------------------------------------------------------ unpack */

/*
 * Decode one column (lane = column) into 16-bit indices x0c[r*128], r < 16.  One
 * straight-line routine per filler class, fully unrolled with predicated stores, so a
 * warp whose 32 columns mix classes pays each routine once per pass instead of
 * diverging inside data-dependent inner loops.  Returns non-zero if a t-code that the
 * reference gets to read is out of range (decode.c:412/:438/:464).
 */
__device__ __forceinline__ int unpack_column(StageReader &br, uint32_t P, uint32_t limit, uint32_t ind,
					     uint32_t inf, int16_t *x0c, const uint64_t *k8,
					     const uint16_t *tt)
{
	int bad = 0;
	if (inf & INF_K) {
		/* prefix codes (decode.c:208-403): up to 7 values per table step */
		const uint64_t *tab = k8 + (inf >> 20) * 256u;
		uint32_t r = 0;
		while (r < (uint32_t)ROWS) {
			const uint64_t e64 = tab[br.peek(P) & 255u];
			const uint32_t e = (uint32_t)e64, hi = (uint32_t)(e64 >> 32);
			const uint32_t k = umin32(e & 15u, (uint32_t)ROWS - r);
			P += (e >> (4u * k)) & 15u;
			int16_t *d = x0c + r * COLS;
#pragma unroll
			for (int j = 0; j < 7; j++)
				if ((uint32_t)j < k)
					d[j * COLS] = (int16_t)nib_s(hi, j);
			r += k;
		}
	} else if (inf & INF_LIN) {
		/* f_linear decode.c:196-206 */
		const uint32_t mask = (1u << ind) - 1u;
		const int mid = 1 << (ind - 1);
#pragma unroll
		for (int r = 0; r < ROWS; r++) {
			x0c[r * COLS] = (int16_t)((int)(br.peek(P) & mask) - mid);
			P += ind;
		}
	} else if (inf & INF_T) {
		/* f_t15 / f_t27 / f_t37 decode.c:405-476 */
		const uint32_t sub = inf >> 20;
		const uint32_t width = sub == 0 ? 5u : 7u, per = sub == 2 ? 2u : 3u;
		const uint32_t ncodes = sub == 2 ? 8u : 6u, mask = (1u << width) - 1u;
		const uint16_t *tab = tt + sub * 128u;
#pragma unroll
		for (int q = 0; q < 8; q++) {
			if ((uint32_t)q < ncodes) {
				const uint32_t e = tab[br.peek(P) & mask];
				if (P + width <= limit && (e & 0x8000u))
					bad = 1;
				P += width;
				const uint32_t r0 = (uint32_t)q * per;
#pragma unroll
				for (int j = 0; j < 3; j++)
					if ((uint32_t)j < per && r0 + j < (uint32_t)ROWS)
						x0c[(r0 + j) * COLS] = (int16_t)nib_s(e, j);
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
 * Transform + output of one block by one warp.  xs holds the 16-bit indices
 * X0[row*128+col]; val is the block's multiplier.  n = words to emit (<= 2048).
 * Returns this lane's checksum contribution.
 */
template <bool CKS>
__device__ __forceinline__ unsigned long long
juggle_and_store(Smem &sm, uint32_t *xs, int slot, int lane, int val, uint8_t *out, uint32_t pos0,
		 uint32_t n, const Format fmt)
{
	uint32_t *h0 = sm.hist0[slot], *h1 = sm.hist1[slot], *h2 = sm.hist2[slot];
	const int16_t *x0 = reinterpret_cast<const int16_t *>(xs);
	uint32_t x[64];
	unsigned long long cks = 0ull;

	/* ---- dequantise (decode.c:174-177, :591-600) and stages 1, 2 in registers:
	 * lane owns m = 32*i + lane */
#pragma unroll
	for (int i = 0; i < 64; i++)
		x[i] = (uint32_t)((int)x0[32 * i + lane] * val);
	uint32_t hx[4], hy[2];
#pragma unroll
	for (int k = 0; k < 4; k++)
		hx[k] = h0[32 * k + lane];
	hy[0] = h1[lane];
	hy[1] = h1[32 + lane];
#pragma unroll
	for (int k = 0; k < 4; k++)
		h0[32 * k + lane] = x[60 + k];
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
	h1[lane] = y[62];
	h1[32 + lane] = y[63];
	__syncwarp(); /* every lane has read its X0 indices: the buffer can be overwritten */
#pragma unroll
	for (int i = 0; i < 64; i++) {
		/* C = 32: m-32 -> i-1, m-64 -> i-2; row parity = i&1 */
		uint32_t p1 = i >= 1 ? y[i - 1] : hy[1];
		uint32_t p2 = i >= 2 ? y[i - 2] : hy[i];
		uint32_t z = lift(y[i], p1, p2, i & 1);
		/* transpose layout: word m lives at m + 4*(m/64); m/64 = i/2 for every lane */
		xs[32 * i + lane + 4 * (i >> 1)] = z;
	}
	__syncwarp();

	/* ---- stages 3..7 in registers: lane owns m in [64*lane, 64*lane+64), halo = the 64
	 * words before it (previous lane's chunk, or the previous block's tail for lane 0) */
	const uint4 *own = reinterpret_cast<const uint4 *>(xs + 68 * lane);
	const uint4 *prev = reinterpret_cast<const uint4 *>(lane ? xs + 68 * (lane - 1) : h2);
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
						dst[q] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
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
	__syncwarp(); /* all halo reads done before the tail of this block becomes history */
	if (lane < 16)
		reinterpret_cast<uint4 *>(h2)[lane] = reinterpret_cast<const uint4 *>(xs + 68 * 31)[lane];
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
	for (int i = tid; i < ACM_T_SIZE; i += THREADS)
		sm.t[i] = a.tables->t[i];
	if (tid < 32) {
		sm.kind[tid] = a.tables->kind[tid];
		sm.info[tid] = make_info((uint32_t)tid, a.tables->kind[tid]);
	}
	if (tid < S) {
		sm.dead[tid] = 0;
		sm.pos[tid] = 0;
		sm.cks[tid] = 0ull;
	}
	if (tid < S * 4)
		mbar_init(&sm.bar[0][0] + tid, 1u);
	if (tid < 2)
		sm.more[tid] = 0;
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncthreads();

	/* scan-lane state (meaningful in the scan warp only) */
	const bool has_slot = warp == W && lane < S;
	const int myslot = lane < S ? lane : 0;
	bool active = false;
	uint32_t cur = 0, P = 0, blk = 0, limit = 0, n_attempt = 0;
	RingReader sbr;
	sbr.reset();
	sbr.ring = sm.ring[myslot];
	sbr.bars = sm.bar[myslot];
	sbr.errflag = a.errflag;

	for (int round = 0;; round++) {
		const int buf = round & 1;
		if (warp == W) {
			/* ================= scan warp: lane = stream slot ================= */
			Entry e;
			e.status = ENT_IDLE;
			e.pblock = 0; e.pend = 0; e.desc = 0; e.blk = 0; e.ncols = 0; e.val = 0; e.pad = 0;
			if (active && sm.dead[myslot] == cur + 1u)
				active = false; /* a worker found a corrupt t-code: abandon the stream */
			if (!active && has_slot) {
				uint32_t idx = atomicAdd(a.counter, 1u);
				if (idx < a.count) {
					const DevStream d = a.streams[idx];
					cur = idx;
					P = d.bit0;
					blk = 0;
					limit = d.file_end + 8u;
					n_attempt = d.n_attempt;
					sbr.start(a.blob + d.base_off,
						  a.blob_room > d.base_off ? a.blob_room - d.base_off : 0, d.file_end);
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
				ScanResult sc = scan_block_flat(sbr, P, limit, sm.coloff[buf] + myslot, sm.info,
								sm.k8, walk);
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
			if (has_slot)
				sm.ent[buf][lane] = e;
			const int any = __any_sync(0xFFFFFFFFu, e.status != ENT_IDLE);
			if (lane == 0)
				sm.more[buf] = any;
		} else if (round > 0) {
			/* ================= worker warps ================= */
			const int pb = buf ^ 1;
			uint32_t *xs = sm.x[warp];
			int16_t *x0 = reinterpret_cast<int16_t *>(xs);
			uint32_t *stage = xs + X0_BYTES / 4;
			for (int k = 0; k < SLOTS_PER_WORKER; k++) {
				const int slot = warp + k * W;
				const Entry e = sm.ent[pb][slot];
				if (e.status == ENT_IDLE)
					continue;
				const DevStream d = a.streams[e.desc];
				const uint32_t bno = e.blk & 0x7FFFFFFFu;
				const bool last = (e.blk >> 31) != 0;
				if (bno == 0) {
					/* new stream in this slot: zero history (decode.c:812) */
#pragma unroll
					for (int q = 0; q < 4; q++)
						sm.hist0[slot][32 * q + lane] = 0u;
					sm.hist1[slot][lane] = 0u; sm.hist1[slot][32 + lane] = 0u;
					sm.hist2[slot][lane] = 0u; sm.hist2[slot][32 + lane] = 0u;
					if (lane == 0) {
						sm.pos[slot] = 0u;
						sm.cks[slot] = 0ull;
						sm.dead[slot] = 0u;
					}
					__syncwarp();
				}
				if (sm.dead[slot] == e.desc + 1u)
					continue; /* stream already finalised by a corrupt code */
				const uint32_t limit_w = d.file_end + 8u;
				const bool ok = e.status == SCAN_OK;
				const uint32_t ncheck = ok ? (uint32_t)COLS : e.ncols + (e.status == -7 ? 1u : 0u);
				int bad = 0;
				if (ncheck) {
					/* ---- stage the block's bytes: 16-byte chunks [c_lo, c_hi) of the stream */
					const uint32_t c_lo = e.pblock >> 7;
					uint32_t c_hi = (e.pend + 32u + 127u) >> 7;
					if (c_hi > c_lo + (uint32_t)STAGE_CHUNKS)
						c_hi = c_lo + (uint32_t)STAGE_CHUNKS;
					const uint8_t *src = a.blob + d.base_off;
					const uint64_t room = a.blob_room > d.base_off ? a.blob_room - d.base_off : 0;
					for (uint32_t c = c_lo + lane; c < c_hi; c += 32) {
						uint4 v = make_uint4(0u, 0u, 0u, 0u);
						if ((uint64_t)c * 16u + 16u <= room)
							v = ldg_nc_v4(src + (size_t)c * 16u);
						reinterpret_cast<uint4 *>(stage)[c - c_lo] = v;
					}
					__syncwarp();
					StageReader br;
					br.init(stage, c_lo * 4u, (c_hi - c_lo) * 4u, d.file_end);
					/* ---- unpack: lane = column */
					const uint16_t *offs = sm.coloff[pb] + slot;
#pragma unroll 1
					for (int p = 0; p < 4; p++) {
						const uint32_t c = 32u * p + lane;
						if (c < ncheck) {
							const uint32_t Pc = e.pblock + offs[c * OFF_PITCH];
							const uint32_t ind = br.peek(Pc) & 31u;
							bad |= unpack_column(br, Pc + 5u, limit_w, ind, sm.info[ind], x0 + c,
									     sm.k8, sm.t);
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
					unsigned long long c2 =
						juggle_and_store<CKS>(sm, xs, slot, lane, e.val, out, pos, n, a.fmt);
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
					size_t nb = d.pad_words > pos ? (size_t)(d.pad_words - pos) * a.fmt.wordlen : 0;
					for (size_t i = lane; i < nb; i += 32)
						p0[i] = 0;
					__syncwarp();
					if (lane == 0) {
						a.status[d.index] = st;
						a.words[d.index] = pos;
						a.cks[d.index] = a.fmt.checksums ? sm.cks[slot] : 0ull;
						sm.dead[slot] = e.desc + 1u;
					}
				}
				__syncwarp();
			}
		}
		__syncthreads();
		if (!sm.more[buf])
			break; /* the scan warp produced nothing this round: all streams are done */
	}
	if (warp == W)
		sbr.drain(); /* no bulk copy may still be landing when the CTA exits */
}

} // namespace fast

bool fast_shape(uint32_t level, uint32_t rows) { return level == fast::LEVEL && rows == fast::ROWS; }

size_t fast_smem_bytes() { return sizeof(fast::Smem); }

int fast_slots_per_cta() { return fast::S; }

cudaError_t launch_fast(const KernelArgs &a, int n_ctas, cudaStream_t st)
{
	if (a.count == 0)
		return cudaSuccess;
	static bool configured[2] = { false, false };
	const size_t smem = sizeof(fast::Smem);
	if (a.fmt.checksums) {
		if (!configured[1]) {
			cudaError_t e = cudaFuncSetAttribute(fast::acm_decode_fast_kernel<true>,
							     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			if (e != cudaSuccess)
				return e;
			configured[1] = true;
		}
		fast::acm_decode_fast_kernel<true><<<n_ctas, fast::THREADS, smem, st>>>(a);
	} else {
		if (!configured[0]) {
			cudaError_t e = cudaFuncSetAttribute(fast::acm_decode_fast_kernel<false>,
							     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			if (e != cudaSuccess)
				return e;
			configured[0] = true;
		}
		fast::acm_decode_fast_kernel<false><<<n_ctas, fast::THREADS, smem, st>>>(a);
	}
	return cudaGetLastError();
}

} // namespace acm
