/*
 * acm_fast2.cu -- the throughput kernel for the common block shape: level 7 (128 columns),
 * 16 rows, 2048 words per block (the shape of BASELINE configs 1, 2, 4).
 *
 * What bounds this path: a stream's bitstream is serial -- where column
 * c+1 starts is only known once column c has been walked -- so a batch cannot
 * finish before (blocks of its longest stream) x (walk steps per block) x (latency of one
 * step), and the walk's throughput is lanes / (steps x latency).  Everything else is parallel
 * and has to keep up with the walk.  One persistent launch, two kinds of CTA:
 *
 *  - SCAN CTAs (the first n_scan of the grid, one per SM, nothing else on that SM) walk column
 *    LENGTHS only, one stream per lane (SW warps x 32 stream slots per CTA).  A step is: fetch
 *    32 stream bits from the lane's shared-memory ring (two conflict-free LDS + one funnel
 *    shift), ONE table lookup (uni16: the whole walk as a state machine -- page 0 is "at a column
 *    selector", indexed by selector + first payload byte; page (type, rows to come) is "inside
 *    a prefix-coded column"; an entry is bits-to-advance | next page), position += advance, and
 *    at a selector a predicated 2-byte store of the column's offset.  No branches on the data:
 *    the 32 lanes run one instruction stream however their column types differ, finished and
 *    idle lanes sit on a HALT page.  A warp issues in order and the walk is one dependent
 *    chain: table LDS -> dp4a (position += advance byte) -> LOP3 (ring address) -> ring LDS ->
 *    funnel shift -> LOP3 (table address).  To get there the lane keeps Q = position - 1 (the
 *    32 bits at Q, masked, are the index already doubled for 16-bit entries, and page | index
 *    is the same LOP3), 32 * Q alongside it (the ring row is a mask of it), and the ring is laid
 *    out [word][warp][lane] with power-of-two rows.  Beyond the chain the step time is queueing
 *    in the SM's load/store pipe: sharing an SM with decode warps doubled it (measured), hence
 *    the dedicated SMs, and anything that keeps lanes out of that pipe pays.  The ring (64
 *    words per lane) is topped up for all lanes together every SCAN_PERIOD steps, LEAD chunks
 *    of 16 bytes ahead of the position: two chunks per lane as 16-byte loads into registers
 *    (stored to the ring at the next top-up), bursts beyond that as 4-byte cp.async copies whose
 *    src-size zero-fill is the end-of-file rule (one zero byte, decode.c:57-61).  End-of-file
 *    verdicts are not part of the walk: a block whose walk ends inside the stream cannot have
 *    read past its end; the rare other block is re-walked with the reference's verdicts
 *    (scan_block).
 *  - Every walked block becomes a 288-byte RECORD (128 column offsets + header facts) in a
 *    per-slot ring of RING_D records in global memory (L2 resident), announced through a
 *    per-slot counter.  Scan lanes run up to RING_D blocks ahead of the decode.
 *  - DECODE CTAs (the rest of the grid) own the slots round-robin.  A worker warp claims an
 *    owned slot that has records pending and decodes its blocks in order:
 *      stage    the block's compressed bytes (<= 4.2 KB) into shared memory, coalesced;
 *      sort     the 128 columns by filler class (ballot + popc) into per-class work lists, so
 *               that every unpack pass runs one straight-line routine on 32 busy lanes;
 *      unpack   prefix-coded columns: 96-bit shift register, one 64-bit table entry per step
 *               (up to 8 values + bit and row counts), 16 nibbles accumulated in two registers,
 *               no row cap (rows past the 16th shift out); radix-coded columns: all codes from
 *               one 64-bit window; linear columns: sliding window.  A column leaves as sixteen
 *               int16 indices in two 128-bit shared-memory stores (halves swizzled:
 *               conflict-free for the transform's 128-bit loads);
 *      juggle   dequantise (idx*val) on load; lifting stages 1-2 (C=64,32) in registers with
 *               lane j owning every word m = j mod 32; one transpose through shared memory;
 *               stages 3-7 (C=16..1) in registers over a recomputed halo, as a rolled loop
 *               (the unrolled 40 KB version thrashed the instruction caches);
 *      output   >>7, low 16 bits, byte order / sign bias folded into one PRMT (+LOP), 128-bit
 *               streaming stores: the block leaves as 4 KiB of PCM.
 *    The reference's wrapbuf (decode.c:803, 2*cols-2 = 254 words) is 256 words of per-slot
 *    history (last 128 X0, 64 X1, 64 X2 words) in an L2-resident global array.
 *  - Streams are handed out to scan lanes by an atomic cursor in longest-first order.
 *
 * Bit-exactness: same arithmetic as the generic kernel (uint32 wrap-around, arithmetic shift,
 * truncation), same table-driven symbol decode, same status rules.
 */
#include <cstdlib>
#include "acm_fast2_core.cuh"
#include "acm_kernels.cuh"

namespace acm {

namespace fast2 {

constexpr int LEVEL = 7;
constexpr int COLS = 128;
constexpr int BLEN = COLS * ROWS;        /* 2048 */
#ifndef F2_W
#define F2_W 16        /* worker warps of a decode CTA */
#endif
#ifndef F2_SW
#define F2_SW 8        /* scan warps of a scan CTA */
#endif
#ifndef F2_RING_D
#define F2_RING_D 32
#endif
#ifndef F2_RW
#define F2_RW 64
#endif
#ifndef F2_PERIOD
#define F2_PERIOD 12   /* walk steps between two top-ups of the scan rings (measured, 1 B200, config 2 / config-4 shape:
			* 8: 4.45 / 18.2 ms, 10: 4.49 / 18.2, 12: 4.22 / 18.15, 14: 4.49 / 19.4, 16: 4.33 / 19.4, 24: 4.46 / 19.7,
			* 32: 4.64 / 20.2 -- lanes that outrun their ring wait for the next top-up) */
#endif
#ifndef F2_STEP_UNROLL
#define F2_STEP_UNROLL 4
#endif
#ifndef F2_HYST
#define F2_HYST 1      /* scan warp sleeps while every lane is at least RING_D/2 records ahead */
#endif
constexpr int W = F2_W;                  /* worker warps per decode CTA */
constexpr int SW = F2_SW;                /* scan warps per scan CTA */
constexpr int SLOTS = 32 * SW;           /* stream slots per scan CTA */
constexpr int THREADS = 32 * W;
constexpr int MAXOWN = 128;              /* slots a decode CTA can own */
constexpr int RING_D = F2_RING_D;              /* block records a scan lane may be ahead of the decode */
constexpr int REC_BYTES = 288;           /* 128 x u16 column offsets + 32-byte Rec */
constexpr int RW = F2_RW;               /* ring words per scan lane (+1 duplicate of word 0) */
constexpr int RROW = 32 * SW;            /* words per ring row: one word of every scan lane of the CTA */
#ifndef F2_PRED_NOTE
#define F2_PRED_NOTE 1
#endif
#ifndef F2_CARRY
#define F2_CARRY 1
#endif
#ifndef F2_KEEP_L2
#define F2_KEEP_L2 1 /* L2 eviction hints on the two reads of the compressed bytes */
#endif
#ifndef F2_UNPACK4
#define F2_UNPACK4 0 /* 1: a lane walks its four columns together (four independent chains) */
#endif
#ifndef F2_S37_UNROLL
#define F2_S37_UNROLL 1 /* lifting stages 3..7 in pairs of pieces (0: a loop of single pieces) */
#endif
#ifndef F2_NHOLD
#define F2_NHOLD 2
#endif
constexpr int NHOLD = F2_NHOLD;          /* 16-byte chunks per lane and period that travel through registers */
constexpr int LEAD = RW / 4 - 2;               /* 16-byte chunks requested ahead of the read position */
constexpr int SCAN_PERIOD = F2_PERIOD;         /* walk steps between two top-ups */
constexpr int STEP_UNROLL = F2_STEP_UNROLL;    /* ... unrolled by */
constexpr int XPRE = 68;                 /* chunk -1: the previous block's last 64 X2 words (+4 pad) */
constexpr int XWORDS = XPRE + BLEN + 4 * 32; /* transpose layout: 4 pad words per 64 */
constexpr int STAGE_W = 1088;            /* 4352 bytes: a whole block (<= 4179 B) + alignment + read-ahead */
constexpr int WB_WORDS = XWORDS + STAGE_W; /* per worker warp: transpose buffer, then the staged block bytes */
constexpr int HIST_WORDS = 256;          /* per slot: X0 tail [0,128) X1 tail [128,192) X2 tail [192,256) */
#ifndef F2_KMAX
#define F2_KMAX 8
#endif
constexpr int KMAX = F2_KMAX;            /* blocks a worker decodes per visit of a slot */

static_assert((XWORDS * 4) % 16 == 0 && (WB_WORDS * 4) % 16 == 0, "bulk copies need 16-byte aligned staging");
static_assert(MAXOWN == 128, "the slot scan reads four groups of 32 lock bits");
static_assert((RW & (RW - 1)) == 0 && RW % 4 == 0, "the scan ring is a power of two of words, whole 16-byte chunks");
static_assert(LEAD + 2 <= RW / 4, "the chunk being read and the one after it are never re-requested");
static_assert(SW <= W, "scan CTAs are launched with the decode CTAs' thread count");


struct Rec {
	uint32_t pblock; /* P of the block header */
	uint32_t pend;   /* P where the scan stopped (block end when status == SCAN_OK) */
	uint32_t desc;   /* index into the kernel's descriptor slice */
	uint32_t blk;    /* block number, bit 31 = last record of the stream */
	int32_t status;  /* SCAN_OK / SCAN_EOF / ACM_ERR_* */
	uint32_t ncols;
	int32_t val;
	uint32_t pad;
};

/* per-slot control words in global memory, shared by the slot's scan lane and the decode CTA
 * that owns the slot */
struct SlotCtl {
	uint32_t prod; /* records published (scan lane writes) */
	uint32_t rem;  /* blocks the stream being walked still has to go (scan lane writes, with prod) */
	uint32_t cons; /* records consumed (decode warp writes) */
	uint32_t dead; /* desc+1 of a stream the decode side finalised early */
	/* second half: the decode side's running state of the slot, handed from one worker warp to the
	 * next under the slot's lock (a bit of SmemWork::busy) */
	uint32_t pos;  /* words delivered so far of the slot's current stream */
	uint32_t pad;
	unsigned long long cks;
};

constexpr int OFFP = 2 * COLS + 8; /* bytes per lane of the column-offset staging (+8: bank spread, 8-byte rows) */

struct SmemScan {
	uint16_t uni16[ACM_UNI_PAGES * ACM_UNI_PSIZE];
	uint32_t ring[RW + 1][SW * 32];   /* [word][scan warp][lane] */
	/* column offsets of the block a lane is walking, copied to the block's record with
	 * coalesced stores at the end of the round (scattered 2-byte global stores from 32 lanes
	 * are 32 partial-sector writes per instruction) */
	unsigned char off[SW][32 * OFFP];
};

struct SmemWork {
	uint64_t k8w[ACM_K8_SIZE];
	uint32_t wb[W][WB_WORDS];
	unsigned long long mbar[W]; /* one transaction barrier per worker warp (bulk-copy staging) */
	uint32_t busy[MAXOWN / 32]; /* bit j: local slot j is being decoded by some worker warp */
	uint32_t info[32];
	uint16_t t[ACM_T_SIZE];
};

constexpr size_t SMEM_BYTES = sizeof(SmemScan) > sizeof(SmemWork) ? sizeof(SmemScan) : sizeof(SmemWork);

#ifndef F2_PROF
#define F2_PROF 0
#endif
#if F2_PROF
#define PROF_DECL unsigned long long prof_t0 = clock64(), prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PROF_MARK(k) do { const unsigned long long t_ = clock64(); prof_acc[k] += t_ - prof_t0; prof_t0 = t_; } while (0)
#define PROF_FLUSH(base) do { if (lane == 0) for (int k_ = 0; k_ < 8; k_++) atomicAdd(a.prof + (base) + k_, prof_acc[k_]); } while (0)
#else
#define PROF_DECL do { } while (0)
#define PROF_MARK(k) do { } while (0)
#define PROF_FLUSH(base) do { } while (0)
#endif

static_assert(SMEM_BYTES <= 232448, "shared memory exceeds the 227 KB a CTA can opt in to");

/* ------------------------------------------------------------------ helpers */

__device__ __forceinline__ uint4 ldg_nc_v4(const void *p)
{
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
		     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
		     : "l"(p));
	return r;
}
/*
 * The compressed bytes are read twice: by the scan lane that walks them, and a few blocks later by
 * the decode warp that stages them.  In between, the batch's PCM (4.5 x as many bytes) streams
 * through the L2; the scan's loads therefore ask the L2 to keep their lines (evict_last), so that
 * the second read is a hit and not a second trip to HBM.
 */
__device__ __forceinline__ uint64_t l2_keep_policy()
{
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void *g, uint64_t pol)
{
#if F2_KEEP_L2
	asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"(saddr), "l"(g), "l"(pol) : "memory");
#else
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(g) : "memory");
#endif
}
__device__ __forceinline__ void cp_async4z(uint32_t saddr, const void *g, uint32_t n, uint64_t pol)
{
#if F2_KEEP_L2
	asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2, %3;" ::"r"(saddr), "l"(g), "r"(n), "l"(pol) : "memory");
#else
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(saddr), "l"(g), "r"(n) : "memory");
#endif
}
__device__ __forceinline__ uint4 ldg_keep_v4(const void *p, uint64_t pol)
{
	uint4 r;
#if F2_KEEP_L2
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
		     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
#else
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
		     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
#endif
	return r;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ uint32_t vol_ld(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ void vol_st(uint32_t *p, uint32_t v) { *reinterpret_cast<volatile uint32_t *>(p) = v; }

/* transaction barrier + bulk copy (TMA): how a worker warp stages a block's bytes.  One lane arms
 * the barrier with the byte count and issues the copy; the copy engine writes shared memory and
 * completes the barrier's phase, every lane waits for that phase. */
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar)
{
#if F2_KEEP_L2
	/* the second and last read of these bytes (the scan lane kept them in the L2): let them go */
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
		     ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar), "l"(pol) : "memory");
#else
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		     ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
#endif
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity)
{
	uint32_t done;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			     "selp.u32 %0, 1, 0, p;\n\t}"
			     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
	} while (!done);
}
/* generic-proxy accesses to shared memory are ordered before the copy engine's next write */
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

/*
 * Watchdog of the two waiting loops.  The scan CTAs and the decode CTAs of a launch wait for each
 * other, so all of them have to be resident; plan_create sizes the grid for that (one CTA per SM,
 * at most the SM count), but a GPU shared with another process could still hold some back.
 * Every scan round and every slot claim bumps a heartbeat word (scan_done[1]); a side that has
 * waited WATCHDOG_NS while the heartbeat stood still raises the error flag (reported by
 * acm_gpu_plan_fetch as an internal failure) and leaves, instead of hanging the device.
 */
constexpr unsigned long long WATCHDOG_NS = 4000000000ull;

__device__ __forceinline__ unsigned long long now_ns()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}

/* ------------------------------------------------------------------ scan */

/*
 * A scan lane's view of its stream: RW ring words in shared memory, word i of the stream
 * (32-bit words from the 16-byte aligned stream base) at ring[i % RW][32 * warp + lane], plus a
 * copy of ring word 0 in row RW so that the pair (i, i+1) is always (row, row + 1).  With a
 * power-of-two row size the address of the lane's word is (P << 5 & mask) | lane bits: two
 * instructions on the walk's dependent chain.
 * Chunks (16 bytes) [.., fill) have been requested; words [.., ready_w) have landed, and
 * a 32-bit fetch at P is safe while P < ready_p.
 */
/* the lane's ring word that holds bit P.  ring0 = row 0 of the CTA's ring, lane4 = byte offset
 * of the lane within a row; row offset and lane offset share no bits when a row is a power of
 * two bytes, so the sum is one LOP3 (and | or) and the array base folds into the LDS. */
__device__ __forceinline__ const uint32_t *ring_word(const uint32_t *ring0, uint32_t lane4, uint32_t P32)
{
	constexpr uint32_t ROWB = 4u * RROW;
	uint32_t off;
	if (ROWB == 1024u)
		off = (P32 & ((RW - 1u) * ROWB)) | lane4; /* P32 = 32 * position: (P >> 5) * 1024 without the low bits */
	else
		off = ((P32 >> 10) & (RW - 1u)) * ROWB + lane4;
	return reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(ring0) + off);
}

struct ScanRing {
	uint32_t saddr;       /* shared-space address of this lane's ring word 0 */
	const uint32_t *rw;   /* the same, generic */
	const uint8_t *base;  /* stream base (16-byte aligned) */
	uint32_t room16;      /* 16-byte chunks readable at base */
	uint32_t full16;      /* chunks [0, full16) lie entirely inside the file and the blob */
	uint32_t fe_byte;     /* bytes of the stream that exist (relative to base) */
	uint32_t fill, fill_prev, ready_w, ready_p;
	/* the first NHOLD chunks a lane asks for in a period travel as one 16-byte load into
	 * registers and are written to the ring at the next top-up: a quarter of the load/store
	 * pipe's work of four 4-byte async copies (the scan SM's pipe is its busiest unit);
	 * whatever a lane wants beyond that (bursts of wide columns) goes by cp.async */
	uint4 hold[NHOLD];
	uint32_t hold_c0, hold_n;
	uint64_t pol; /* L2 eviction policy of this lane's loads */

	/* no stream: the lane's walk sits on the HALT page at P = 0 over two zero ring words, so that
	 * all idle lanes of a warp look up the same table word (a broadcast, not a bank conflict) */
	__device__ __forceinline__ void idle()
	{
		cp_async_wait_all();
		const_cast<uint32_t *>(rw)[0] = 0u;
		const_cast<uint32_t *>(rw)[RROW] = 0u;
		base = nullptr;
		hold_n = 0;
		room16 = 0;
		full16 = 0;
		fe_byte = 0;
		fill = fill_prev = 0x0FFFFFF0u; /* never asks for data */
		ready_w = 0;
		ready_p = 0;
	}
	__device__ __forceinline__ void start(const uint8_t *src, uint64_t room, uint32_t file_end, uint32_t P0)
	{
		cp_async_wait_all(); /* copies of the slot's previous stream must not land after this one's */
		base = src;
		hold_n = 0;
		room16 = (uint32_t)(room >> 4);
		fe_byte = file_end >> 3;
		full16 = fe_byte >> 4 < room16 ? fe_byte >> 4 : room16;
		fill = P0 >> 7;
		/* the first chunks are fetched synchronously (once per stream): the walk starts at once */
		for (int j = 0; j < LEAD; j++)
			request(fill + j);
		fill += LEAD;
		cp_async_commit();
		cp_async_wait_all();
		fill_prev = fill;
		ready_w = fill * 4u;
		ready_p = (ready_w - 1u) * 32u;
	}
	__device__ __forceinline__ void request(uint32_t c)
	{
		const uint32_t sa = saddr + (c & (RW / 4 - 1)) * (16u * RROW);
		const uint8_t *g = base + (size_t)c * 16u;
		if (c < full16) {
			cp_async4(sa, g, pol);
			cp_async4(sa + 4 * RROW, g + 4, pol);
			cp_async4(sa + 8 * RROW, g + 8, pol);
			cp_async4(sa + 12 * RROW, g + 12, pol);
			if ((c & (RW / 4 - 1)) == 0)
				cp_async4(saddr + RW * 4 * RROW, g, pol);
		} else {
			/* touches the end of the file (or of the blob): bytes at and past it read as zero,
			 * which is the reference's "one zero byte, then nothing" (decode.c:57-61) */
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const uint32_t at = c * 16u + 4u * k;
				uint32_t n = 0;
				if (c < room16 && at < fe_byte)
					n = fe_byte - at < 4u ? fe_byte - at : 4u;
				const void *src = n ? (const void *)(g + 4 * k) : (const void *)base;
				cp_async4z(sa + 4 * RROW * k, src, n, pol);
				if (k == 0 && (c & (RW / 4 - 1)) == 0)
					cp_async4z(saddr + RW * 4 * RROW, src, n, pol);
			}
		}
	}
	/* all lanes together, every SCAN_PERIOD steps */
	__device__ __forceinline__ void topup(uint32_t P)
	{
		/* last period's register chunks go into the ring (the lane's own column of it) */
#pragma unroll
		for (int k = 0; k < NHOLD; k++) {
			if ((uint32_t)k < hold_n) {
				const uint32_t c = hold_c0 + k;
				uint32_t *row = const_cast<uint32_t *>(rw) + (c & (RW / 4 - 1)) * (4u * RROW);
				row[0] = hold[k].x;
				row[RROW] = hold[k].y;
				row[2 * RROW] = hold[k].z;
				row[3 * RROW] = hold[k].w;
				if ((c & (RW / 4 - 1)) == 0)
					const_cast<uint32_t *>(rw)[RW * RROW] = hold[k].x;
			}
		}
		/* after the position has jumped (a block re-walked from global memory), the chunks
		 * behind it are never read: asking for them would wrap the ring */
		const uint32_t f0 = fill < (P >> 7) ? (P >> 7) : fill;
		const int want = (int)((P >> 7) + LEAD) - (int)f0;
		int nreg = want < NHOLD ? want : NHOLD;
		const int inside = (int)full16 - (int)f0; /* chunks from f0 that lie entirely inside the stream */
		nreg = nreg < inside ? nreg : inside;
		nreg = nreg > 0 ? nreg : 0;
#pragma unroll
		for (int k = 0; k < NHOLD; k++)
			if (k < nreg)
				hold[k] = ldg_keep_v4(base + (size_t)(f0 + k) * 16u, pol);
		hold_c0 = f0;
		hold_n = (uint32_t)nreg;
#pragma unroll 1
		for (int j = nreg; j < want; j++)
			request(f0 + j);
		fill = want > 0 ? f0 + want : f0;
		cp_async_commit();
		cp_async_wait1(); /* everything but the group just committed has landed */
		ready_w = base ? fill_prev * 4u : 0u;
		ready_p = ready_w ? (ready_w - 1u) * 32u : 0u;
		fill_prev = fill;
	}
};

/*
 * One walk step for all lanes of a scan warp: fetch the 32 stream bits at P from the ring, note
 * the column offset if the lane is at a selector, one uni16 lookup, apply it.  After the 128th
 * column the lane moves to the HALT page (entries advance 0 bits and stay), which is also where
 * finished and idle lanes sit: every lane runs the same instructions, no branch (a branch here
 * costs a convergence barrier per step).  A lane whose bits have not landed yet (P >= ready_p)
 * does nothing this step.  No end-of-file checks here: bits past the end read as zero, and a
 * block whose walk ends at or before the stream's limit cannot have read past it (the caller
 * re-walks the rare other case with the reference's verdicts).
 * A warp issues in order and the walk is one dependent chain through two shared-memory round
 * trips; measured variants (a register window that keeps the ring fetch off the chain, per-period
 * instead of per-step data checks with rollback) are in profiles/r01_ncu_fast2.md.
 */
__device__ __forceinline__ void fast_step(Walk &s, uint32_t &cp, uint32_t cpend, uint32_t qblock,
					  const uint32_t *ring0, uint32_t lane4, uint32_t ready_p,
					  const unsigned char *uni)
{
	const uint32_t *rp = ring_word(ring0, lane4, s.Q32);
	const uint32_t w1 = fsr(rp[0], rp[RROW], s.Q);
	const uint32_t e = *reinterpret_cast<const uint16_t *>(uni + walk_index(s, w1));
	const bool have = s.Q + 1u <= ready_p;
	const bool note = have && s.msk == MSK_SEL;
#if F2_PRED_NOTE
	/* SmemScan::off (cp = shared-space address): a predicated store, not a branch; lanes that are
	 * not at a selector (two steps in three) stay out of the load/store pipe */
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u16 [%0], %1;\n\t}"
		     :: "r"(cp), "h"((unsigned short)(s.Q - qblock)), "r"((uint32_t)note) : "memory");
#else
	asm volatile("st.shared.u16 [%0], %1;" :: "r"(note ? cp : cpend), "h"((unsigned short)(s.Q - qblock)) : "memory");
#endif
	cp += note ? 2u : 0u;
	const bool at_sel = walk_next_if(s, e, have); /* not landed: advance 0, same page */
	const bool done = at_sel && cp == cpend;
	s.s8 = done ? UNI_HALT8 : s.s8;
	s.msk = done ? MSK_K : s.msk;
}

/* ------------------------------------------------------------------ transform + output */

__device__ __forceinline__ uint32_t lift(uint32_t a, uint32_t p1, uint32_t p2, bool odd)
{
	/* decode.c:518-519 */
	uint32_t s = a + p2;
	return odd ? 2u * p1 - s : 2u * p1 + s;
}

/*
 * Everything the transform touches is carried times 2^QS (QS = 16 - LEVEL): the dequantisation
 * multiplies by val << QS, the stage-1 "+1" is 1 << QS.  The lifting is linear modulo 2^32
 * (decode.c:512: unsigned arithmetic), so a result is the reference's times 2^QS modulo 2^32,
 * and the 16 bits the output wants -- bits LEVEL..LEVEL+15 of the reference's word, decode.c:620 --
 * are its top half: packing two results is ONE byte permute instead of two shifts and a permute.
 */
constexpr int QS = 16 - LEVEL;

/* pack two results into one 32-bit word of 16-bit PCM: the top halves of a and b */
__device__ __forceinline__ uint32_t pack2(uint32_t a, uint32_t b, uint32_t sel, uint32_t flip)
{
	return __byte_perm(a, b, sel) ^ flip;
}

/*
 * One 16-word piece (window words t = 16 K .. 16 K + 15) of lifting stages 3..7.  Stage 3 runs
 * from K = 2, stages 4-6 (S46) from K = 3, stage 7 and the output (OUT) from K = 4; ODD = K & 1 is
 * the row parity of stage 3.  P1 = the stage-3 inputs of piece K - 1, A3..A6 = what the later
 * stages need of piece K - 1; all are replaced by this piece's.  The caller runs the pieces in
 * pairs (even, odd): within a pair nothing is copied, the stage-3 sign is a constant, and only
 * one set of carried values crosses the loop's back edge (a loop of single pieces copies 46
 * registers per piece; six pieces in a row are 16 KB of code and fall out of the instruction
 * cache: measured, profiles/r02_decode.md).
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
		/* C = 16 (decode.c:518-519): 2*in[t-16] +- (in[t] + in[t-32]); row parity = K & 1 */
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
					const uint32_t w0 = ((v0 >> 16) + bias) & 0xFFFFu;
					const uint32_t w1 = ((v1 >> 16) + bias) & 0xFFFFu;
					if (m - pos0 < n)
						cks += (unsigned long long)(m + 1u) * (w0 + 1ull);
					if (m + 1u - pos0 < n)
						cks += (unsigned long long)(m + 2u) * (w1 + 1ull);
				}
			}
			const int q = 2 * (K - 4);
			if (full) {
				/* streaming stores: PCM is written once and must not evict the L2-resident
				 * history, records and compressed bytes */
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

/*
 * Transform + output of one block by one warp.  x[i] = the dequantised word m = 32 i + lane of
 * the block (row i / 4, column 32 (i % 4) + lane: the lane's own four columns, straight from
 * the unpackers); xs = the warp's transpose buffer; h = the history of the slot's previous block
 * (all zero for a stream's first block, decode.c:812), gh = where this block's goes.  n = words
 * to emit (<= 2048).  Returns this lane's checksum contribution.
 */
struct Hist {
	uint32_t hx[4], hy[2], hz[2];
};

/* history of the slot's previous block (L2 resident; .cg: never a stale L1 line); issued before
 * the unpack, so that the transform does not wait for it */
__device__ __forceinline__ void hist_load(Hist &h, const uint32_t *gh, bool first, int lane)
{
#pragma unroll
	for (int k = 0; k < 4; k++)
		h.hx[k] = first ? 0u : __ldcg(gh + 32 * k + lane);       /* X0[-128 + 32k + lane] */
	h.hy[0] = first ? 0u : __ldcg(gh + 128 + lane);               /* X1[-64 + lane] */
	h.hy[1] = first ? 0u : __ldcg(gh + 160 + lane);               /* X1[-32 + lane] */
	h.hz[0] = first ? 0u : __ldcg(gh + 192 + lane);               /* X2[-64 + lane] */
	h.hz[1] = first ? 0u : __ldcg(gh + 224 + lane);               /* X2[-32 + lane] */
}

template <bool CKS>
__device__ __forceinline__ unsigned long long
juggle_and_store(uint32_t (&x)[64], const Hist &h, uint32_t *xs0, uint32_t *gh, int lane, uint8_t *out,
		 uint32_t pos0, uint32_t n, const Format fmt)
{
	uint32_t *xs = xs0 + XPRE;
	unsigned long long cks = 0ull;
	const uint32_t (&hx)[4] = h.hx;
	const uint32_t (&hy)[2] = h.hy;
	const uint32_t (&hz)[2] = h.hz;

	/* ---- stages 1, 2 in registers: lane owns m = 32*i + lane */
#pragma unroll
	for (int k = 0; k < 4; k++)
		__stcg(gh + 32 * k + lane, x[60 + k]);
	/* decode.c:561-564: +1 where m % 64 == 0, i.e. lane 0, even i; it rides in the lift's first add */
	const uint32_t one0 = lane == 0 ? 1u << QS : 0u;
	uint32_t y[64];
#pragma unroll
	for (int i = 0; i < 64; i++) {
		/* C = 64: m-64 -> i-2, m-128 -> i-4; row parity = (m/64)&1 = (i>>1)&1 */
		const uint32_t p1 = i >= 2 ? x[i - 2] : hx[i + 2];
		const uint32_t p2 = i >= 4 ? x[i - 4] : hx[i];
		const bool odd = (i >> 1) & 1;
		const uint32_t sum = (i & 1) ? x[i] + p2 : odd ? x[i] + p2 - one0 : x[i] + p2 + one0;
		y[i] = odd ? 2u * p1 - sum : 2u * p1 + sum;
	}
	__stcg(gh + 128 + lane, y[62]);
	__stcg(gh + 160 + lane, y[63]);
	__syncwarp(); /* every lane has picked up its parked linear columns: the buffer can be overwritten */
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

	/* ---- stages 3..7 in registers: lane owns m in [64*lane, 64*lane+64) and walks the 128-word
	 * window t = 0..127 made of the 64 words before it (the halo: the previous lane's chunk,
	 * chunk -1 for lane 0: 68*(lane-1) = -XPRE) and its own 64 words, in 16-word pieces
	 * k = 2..7 (t = 16k..16k+15).  A ROLLED loop: the body is ~4 KB of code that stays in the
	 * instruction caches (the fully unrolled version was 40 KB and every worker warp streamed it
	 * from L2, profiles/r01_ncu_fast2.md).  The stage-3 inputs of a piece come from shared
	 * memory; what later stages need of the previous piece is carried in registers
	 * (A3: its 16 stage-3 outputs, A4: its last 8 stage-4 outputs, A5: 4, A6: 2).  A stage is
	 * only run where its inputs are complete: stage 3 from t = 32, stages 4-6 from t = 48,
	 * stage 7 (= the output) from t = 64. */
	const uint32_t sel = fmt.be ? 0x6723u : 0x7632u;
	const uint32_t flip = fmt.bias ? (fmt.be ? 0x00800080u : 0x80008000u) : 0u;
	const bool full = (uint32_t)(64 * lane + 64) <= n;
	uint4 *dst = reinterpret_cast<uint4 *>(out + ((size_t)pos0 + 64u * lane) * 2u);
	const uint32_t *win = xs + 68 * (lane - 1);
	uint32_t A3[16], A4[8], A5[4], A6[2];
#pragma unroll
	for (int j = 0; j < 16; j++)
		A3[j] = 0u;
#pragma unroll
	for (int j = 0; j < 8; j++)
		A4[j] = 0u;
	A5[0] = A5[1] = A5[2] = A5[3] = 0u;
	A6[0] = A6[1] = 0u;
#if F2_S37_UNROLL
	{
		uint32_t P1[16]; /* the inputs of piece k - 1 */
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
#else
#if F2_CARRY >= 1
	uint32_t P1[16]; /* the inputs of piece k - 1 */
#pragma unroll
	for (int q = 0; q < 4; q++) {
		const uint4 v = reinterpret_cast<const uint4 *>(win + 16)[q];
		P1[4 * q] = v.x; P1[4 * q + 1] = v.y; P1[4 * q + 2] = v.z; P1[4 * q + 3] = v.w;
	}
#endif
#if F2_CARRY >= 2
	uint32_t P2[16]; /* the inputs of piece k - 2 */
#pragma unroll
	for (int q = 0; q < 4; q++) {
		const uint4 v = reinterpret_cast<const uint4 *>(win)[q];
		P2[4 * q] = v.x; P2[4 * q + 1] = v.y; P2[4 * q + 2] = v.z; P2[4 * q + 3] = v.w;
	}
#endif
#pragma unroll 1
	for (int k = 2; k < 8; k++) {
		/* word t of the window lives at t + (t >= 64 ? 4 : 0) */
		const uint4 *p0 = reinterpret_cast<const uint4 *>(win + 16 * k + (k >= 4 ? 4 : 0));
		const uint4 *p1 = reinterpret_cast<const uint4 *>(win + 16 * (k - 1) + (k >= 5 ? 4 : 0));
		const uint4 *p2 = reinterpret_cast<const uint4 *>(win + 16 * (k - 2) + (k >= 6 ? 4 : 0));
		const uint32_t sgn = (k & 1) ? 0xFFFFFFFFu : 1u; /* row parity of stage 3 = (t >> 4) & 1 */
		uint32_t a3[16];
#pragma unroll
		for (int q = 0; q < 4; q++) {
#if F2_CARRY >= 2
			const uint4 c0 = p0[q];
			const uint4 c1 = make_uint4(P1[4 * q], P1[4 * q + 1], P1[4 * q + 2], P1[4 * q + 3]);
			const uint4 c2 = make_uint4(P2[4 * q], P2[4 * q + 1], P2[4 * q + 2], P2[4 * q + 3]);
			P2[4 * q] = c1.x; P2[4 * q + 1] = c1.y; P2[4 * q + 2] = c1.z; P2[4 * q + 3] = c1.w;
			P1[4 * q] = c0.x; P1[4 * q + 1] = c0.y; P1[4 * q + 2] = c0.z; P1[4 * q + 3] = c0.w;
#elif F2_CARRY == 1
			/* the piece before this one was this loop's c0 a turn ago: kept in registers, one
			 * 128-bit shared-memory load less per four words (these loads are a fifth of the
			 * kernel's shared-memory traffic) */
			const uint4 c0 = p0[q], c2 = p2[q];
			const uint4 c1 = make_uint4(P1[4 * q], P1[4 * q + 1], P1[4 * q + 2], P1[4 * q + 3]);
			P1[4 * q] = c0.x; P1[4 * q + 1] = c0.y; P1[4 * q + 2] = c0.z; P1[4 * q + 3] = c0.w;
#else
			const uint4 c0 = p0[q], c1 = p1[q], c2 = p2[q];
#endif
			/* C = 16 (decode.c:518-519): 2*in[t-16] +- (in[t] + in[t-32]) */
			a3[4 * q + 0] = 2u * c1.x + sgn * (c0.x + c2.x);
			a3[4 * q + 1] = 2u * c1.y + sgn * (c0.y + c2.y);
			a3[4 * q + 2] = 2u * c1.z + sgn * (c0.z + c2.z);
			a3[4 * q + 3] = 2u * c1.w + sgn * (c0.w + c2.w);
		}
		if (k >= 3) {
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
			if (k >= 4) {
				uint32_t pk[8];
#pragma unroll
				for (int j = 0; j < 16; j += 2) { /* C = 1 */
					const uint32_t v0 = lift(a6[j], j >= 1 ? a6[j - 1] : A6[1], j >= 2 ? a6[j - 2] : A6[0], 0);
					const uint32_t v1 = lift(a6[j + 1], a6[j], j >= 1 ? a6[j - 1] : A6[1], 1);
					pk[j >> 1] = pack2(v0, v1, sel, flip);
					if (CKS) {
						/* u_i as an unsigned 16-bit value, independent of byte order */
						const uint32_t m = pos0 + 64u * lane + 16u * (uint32_t)(k - 4) + (uint32_t)j;
						const uint32_t w0 = ((v0 >> 16) + fmt.bias) & 0xFFFFu;
						const uint32_t w1 = ((v1 >> 16) + fmt.bias) & 0xFFFFu;
						if (m - pos0 < n)
							cks += (unsigned long long)(m + 1u) * (w0 + 1ull);
						if (m + 1u - pos0 < n)
							cks += (unsigned long long)(m + 2u) * (w1 + 1ull);
					}
				}
				const int q = 2 * (k - 4);
				if (full) {
					/* streaming stores: PCM is written once and must not evict the L2-resident
					 * history, records and compressed bytes */
					__stcs(dst + q, make_uint4(pk[0], pk[1], pk[2], pk[3]));
					__stcs(dst + q + 1, make_uint4(pk[4], pk[5], pk[6], pk[7]));
				} else {
					/* last block of a stream: word-granular tail */
					uint16_t *d16 = reinterpret_cast<uint16_t *>(dst + q);
#pragma unroll
					for (int e = 0; e < 16; e++) {
						const uint32_t m = 64u * lane + 8u * q + e;
						if (m < n)
							d16[e] = (uint16_t)(pk[e >> 1] >> (16 * (e & 1)));
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
		for (int j = 0; j < 16; j++)
			A3[j] = a3[j];
	}
#endif
	__syncwarp(); /* all shared-memory reads of this block are done */
	return cks;
}

/* ------------------------------------------------------------------ worker: one block record */

/* the stream facts a worker needs, cached in the lane that holds the slot */
struct Stage {
	const uint32_t *st;
	uint32_t w_lo;
	__device__ __forceinline__ uint32_t word(uint32_t i) const { return st[i - w_lo]; }
};

__device__ __forceinline__ uint2 vol_ld2(const uint32_t *p)
{
	uint2 v;
	asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void vol_st2(uint32_t *p, uint32_t x, uint32_t y)
{
	asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p)
{
	uint32_t v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release(uint32_t *p, uint32_t v)
{
	asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void rec_load(Rec &e, const uint8_t *recbase)
{
	const uint4 r0 = __ldcg(reinterpret_cast<const uint4 *>(recbase + 256));
	const uint4 r1 = __ldcg(reinterpret_cast<const uint4 *>(recbase + 272));
	e.pblock = r0.x; e.pend = r0.y; e.desc = r0.z; e.blk = r0.w;
	e.status = (int32_t)r1.x; e.ncols = r1.y; e.val = (int32_t)r1.z; e.pad = r1.w;
}

/* 16-byte chunks [c_lo, c_hi) of the stream hold the block (and the unpackers' read-ahead) */
__device__ __forceinline__ void stage_range(const Rec &e, uint32_t &c_lo, uint32_t &c_hi)
{
	c_lo = e.pblock >> 7;
	c_hi = (e.pend + 160u + 127u) >> 7;
	if (c_hi > c_lo + (uint32_t)(STAGE_W / 4))
		c_hi = c_lo + (uint32_t)(STAGE_W / 4);
}

/* Issues the bulk copy of a block's bytes into the warp's staging area (which nobody reads any
 * more: the caller has synchronised the warp).  Returns the bytes under way (0: nothing to wait
 * for). */
__device__ __forceinline__ uint32_t stage_issue(const KernelArgs &a, const DevStream &d, const Rec &e, uint32_t *stage,
					       uint32_t mbar, int lane)
{
	uint32_t c_lo, c_hi;
	stage_range(e, c_lo, c_hi);
	const uint64_t room16 = (a.blob_room > d.base_off ? a.blob_room - d.base_off : 0) >> 4;
	const uint32_t c_cp = (uint64_t)c_hi < room16 ? c_hi : (uint32_t)room16; /* chunks that exist in the blob */
	const uint32_t nbytes = c_cp > c_lo ? (c_cp - c_lo) * 16u : 0u;
	if (nbytes && lane == 0) {
		fence_proxy_async();
		mbar_expect_tx(mbar, nbytes);
		bulk_g2s((uint32_t)__cvta_generic_to_shared(stage), a.blob + d.base_off + (size_t)c_lo * 16u, nbytes, mbar);
	}
	return nbytes;
}

/*
 * A visit: up to KMAX consecutive records (= blocks) of one slot, in order.
 * Staging: a block's bytes arrive by ONE bulk copy (TMA) on the warp's transaction barrier; the
 * copy for block i+1 is issued as soon as block i is unpacked, so it travels while block i is
 * transformed.  The end-of-file rule (bits at and past file_end read as zero, decode.c:57-61) is
 * applied afterwards to the rare block that reaches the end of its file.
 * Unpack: lane j owns columns j, j+32, j+64, j+96 (pass p = 0..3), i.e. exactly the words
 * m = j mod 32 that lifting stages 1-2 want in its registers: the values go from the code tables
 * to the transform without touching shared memory (linear columns are parked in the lane's own
 * words of the transpose buffer).
 */
template <bool CKS>
__device__ __forceinline__ void decode_visit(SmemWork &sm, const KernelArgs &a, uint32_t *wb, uint32_t mbar,
					     uint32_t &mphase, int lane, const uint8_t *slot_ring, uint32_t c0,
					     uint32_t nrec, uint32_t *gh, SlotCtl *ctl, uint32_t &pos,
					     unsigned long long &cks, uint32_t &dead)
{
	uint32_t *stage = wb + XWORDS;
	DevStream d;
	uint32_t d_id = 0xFFFFFFFFu;
	uint32_t under_way = 0u;   /* bytes of a staging copy that has been issued and not waited for */
	uint32_t staged_for = 0u;  /* ... and the record (index + 1) it belongs to */
	PROF_DECL; /* 24: record + descriptor, 25: staging wait, 26: unpack, 27: next copy + dequantise, 28: transform, 29: rest */
	for (uint32_t i = 0; i < nrec; i++) {
		PROF_MARK(5);
		const uint32_t c = c0 + i;
		const uint8_t *recbase = slot_ring + (size_t)(c % RING_D) * REC_BYTES;
		Rec e;
		rec_load(e, recbase);
		uint32_t offs[4]; /* selector positions of the lane's four columns, relative to the block */
#pragma unroll
		for (int p = 0; p < 4; p++)
			offs[p] = __ldcg(reinterpret_cast<const uint16_t *>(recbase) + 32 * p + lane);
		if (e.desc != d_id) { /* consecutive records of a slot mostly belong to one stream */
			d = a.streams[e.desc];
			d_id = e.desc;
		}
		const uint32_t bno = e.blk & 0x7FFFFFFFu;
		const bool last = (e.blk >> 31) != 0;
		if (bno == 0) {
			pos = 0u;
			cks = 0ull;
		}
		const bool skip = dead == e.desc + 1u; /* stream already finalised by a corrupt code (ids are unique) */
		const uint32_t limit_w = d.file_end + 8u;
		const bool ok = e.status == SCAN_OK;
		const uint32_t ncheck = skip ? 0u : ok ? (uint32_t)COLS : e.ncols + (e.status == -7 ? 1u : 0u);
		PROF_MARK(0);
		/* ---- the block's bytes */
		if (staged_for != c + 1u) {
			if (under_way) { /* cannot happen (a copy is only issued for the next record); keep the phase right */
				mbar_wait(mbar, mphase);
				mphase ^= 1u;
				under_way = 0u;
			}
			__syncwarp(); /* nobody still reads the previous block's bytes */
			if (ncheck)
				under_way = stage_issue(a, d, e, stage, mbar, lane);
		}
		if (under_way) {
			mbar_wait(mbar, mphase);
			mphase ^= 1u;
			under_way = 0u;
		}
		staged_for = 0u;
		if (skip) {
			__syncwarp();
			continue;
		}
		PROF_MARK(1);
		int bad = 0;
		uint32_t A0[4], A1[4];              /* per pass: the column's sixteen nibbles */
		uint32_t linmask = 0u, linany = 0u; /* bit p: this lane's / some lane's column of pass p is linear */
		Hist h;
		hist_load(h, gh, bno == 0, lane);
		/* the record after this one (it exists: nrec was counted behind an acquire): asked for now,
		 * needed after the unpack */
		const bool more = i + 1u < nrec && ok && !last;
		Rec en;
		if (more)
			rec_load(en, slot_ring + (size_t)((c + 1u) % RING_D) * REC_BYTES);
		if (ncheck) {
			uint32_t c_lo, c_hi;
			stage_range(e, c_lo, c_hi);
			const uint32_t fe_word = d.file_end >> 5, fe_tail = d.file_end & 31u;
			if (c_hi * 4u > fe_word) {
				/* the staged range reaches the end of the file: zero what lies at and past it
				 * (and the chunks past the end of the blob, which were not copied) */
				const uint32_t k0 = fe_word > c_lo * 4u ? fe_word : c_lo * 4u;
				for (uint32_t k = k0 + lane; k < c_hi * 4u; k += 32) {
					uint32_t v = 0u;
					if (k == fe_word && fe_tail)
						v = stage[k - c_lo * 4u] & ((1u << fe_tail) - 1u);
					stage[k - c_lo * 4u] = v;
				}
				__syncwarp();
			}
			Stage sr;
			sr.st = stage;
			sr.w_lo = c_lo * 4u;
#if F2_UNPACK4
			/* ---- unpack the lane's four columns (32 p + lane).  Prefix- and radix-coded (and zero)
			 * columns become sixteen nibbles in two registers (A0[p] / A1[p]), all four columns
			 * advancing together (independent chains); the sixteen words of a linear column are
			 * parked in the lane's own words of the transpose buffer (word 32 i + lane, i = 4 r + p:
			 * conflict-free, nobody else touches them). */
			uint32_t lo[4], mid[4], hi[4], cls[4], sub[4], Pp[4], ind[4];
#pragma unroll
			for (int p = 0; p < 4; p++) {
				const uint32_t col = 32u * (uint32_t)p + (uint32_t)lane;
				const uint32_t Pc = e.pblock + offs[p], iw = Pc >> 5;
				Pp[p] = Pc + 5u; /* payload */
				cls[p] = ACM_CLS_ZERO;
				sub[p] = 0u;
				ind[p] = 0u;
				lo[p] = mid[p] = hi[p] = 0u;
				A0[p] = A1[p] = 0u; /* f_zero decode.c:181-188: all rows zero */
				if (col < ncheck) {
					const uint32_t w0 = sr.word(iw), w1 = sr.word(iw + 1), w2 = sr.word(iw + 2),
						       w3 = sr.word(iw + 3), w4 = sr.word(iw + 4);
					ind[p] = __funnelshift_r(w0, w1, Pc) & 31u;
					const uint32_t kind = sm.info[ind[p]];
					cls[p] = kind & 7u;
					sub[p] = kind >> 3;
					const bool up = (Pc & 31u) + 5u >= 32u; /* the payload starts in the next word */
					const uint32_t v0 = up ? w1 : w0, v1 = up ? w2 : w1, v2 = up ? w3 : w2, v3 = up ? w4 : w3;
					lo[p] = __funnelshift_r(v0, v1, Pp[p]);
					mid[p] = __funnelshift_r(v1, v2, Pp[p]);
					hi[p] = __funnelshift_r(v2, v3, Pp[p]);
				}
			}
			bad |= unpack_t4(lo, mid, Pp, limit_w, cls, sub, sm.t, A0, A1);
			if (ok) {
				unpack_k4(lo, mid, hi, cls, sub, sm.k8w, A0, A1);
#pragma unroll
				for (int p = 0; p < 4; p++) {
					const bool isl = cls[p] == ACM_CLS_LINEAR;
#pragma unroll 1
					for (uint32_t t = __any_sync(0xFFFFFFFFu, isl) ? 1u : 0u; t; t--) { /* a real branch */
						if (isl) {
							uint32_t v[ROWS];
							unpack_linear(sr, Pp[p], ind[p], e.val << QS, v);
#pragma unroll
							for (int r = 0; r < ROWS; r++)
								wb[32 * (4 * r + p) + lane] = v[r];
						}
						__syncwarp();
						linmask |= (isl ? 1u : 0u) << p;
						linany |= 1u << p;
					}
				}
			}
#else
			/* ---- unpack: four passes of 32 columns (pass p: column 32 p + lane).  Prefix- and
			 * radix-coded (and zero) columns leave the pass as sixteen nibbles in two registers
			 * (A0[p] / A1[p]); the sixteen words of a linear column are parked in the lane's own
			 * words of the transpose buffer (word 32 i + lane, i = 4 r + p: conflict-free, nobody
			 * else touches them).  One copy of the code: the loop is not unrolled.  (Walking the
			 * lane's four columns together, F2_UNPACK4, costs more instructions than it hides
			 * latency: the decode SMs are bound by instruction issue, profiles/r02_decode.md.) */
#pragma unroll 1
			for (int p = 0; p < 4; p++) {
				const uint32_t col = 32u * (uint32_t)p + (uint32_t)lane;
				const uint32_t off = p == 0 ? offs[0] : p == 1 ? offs[1] : p == 2 ? offs[2] : offs[3];
				uint32_t cls = ACM_CLS_ZERO, sub = 0u, ind = 0u, lo = 0u, mid = 0u, hi = 0u;
				const uint32_t P = e.pblock + off + 5u; /* payload */
				if (col < ncheck) {
					const uint32_t Pc = e.pblock + off, iw = Pc >> 5;
					const uint32_t w0 = sr.word(iw), w1 = sr.word(iw + 1), w2 = sr.word(iw + 2),
						       w3 = sr.word(iw + 3), w4 = sr.word(iw + 4);
					ind = __funnelshift_r(w0, w1, Pc) & 31u;
					const uint32_t kind = sm.info[ind];
					cls = kind & 7u;
					sub = kind >> 3;
					const bool up = (Pc & 31u) + 5u >= 32u; /* the payload starts in the next word */
					const uint32_t v0 = up ? w1 : w0, v1 = up ? w2 : w1, v2 = up ? w3 : w2, v3 = up ? w4 : w3;
					lo = __funnelshift_r(v0, v1, P);
					mid = __funnelshift_r(v1, v2, P);
					hi = __funnelshift_r(v2, v3, P);
				}
				const bool isk = cls == ACM_CLS_K, ist = cls == ACM_CLS_T, isl = cls == ACM_CLS_LINEAR;
				uint32_t a0 = 0u, a1 = 0u; /* f_zero decode.c:181-188: all rows zero */
				if (ok && __any_sync(0xFFFFFFFFu, isk)) {
					if (isk)
						unpack_k(lo, mid, hi, sub, sm.k8w, a0, a1);
					__syncwarp();
				}
				if (__any_sync(0xFFFFFFFFu, ist)) {
					if (ist)
						bad |= unpack_t(lo, mid, P, limit_w, sub, sm.t, a0, a1);
					__syncwarp();
				}
				if (ok && __any_sync(0xFFFFFFFFu, isl)) {
					if (isl) {
						uint32_t v[ROWS];
						unpack_linear(sr, P, ind, e.val << QS, v);
#pragma unroll
						for (int r = 0; r < ROWS; r++)
							wb[32 * (4 * r + p) + lane] = v[r];
					}
					__syncwarp();
					linmask |= (isl ? 1u : 0u) << p;
					linany |= 1u << p;
				}
				if (p == 0) {
					A0[0] = a0; A1[0] = a1;
				} else if (p == 1) {
					A0[1] = a0; A1[1] = a1;
				} else if (p == 2) {
					A0[2] = a0; A1[2] = a1;
				} else {
					A0[3] = a0; A1[3] = a1;
				}
			}
#endif
		}
		bad = __any_sync(0xFFFFFFFFu, bad);
		PROF_MARK(2);
		/* ---- the next block's bytes travel while this one is transformed */
		__syncwarp(); /* the staged bytes have been read */
		if (more && !bad) {
			if (en.desc == e.desc && en.status == SCAN_OK) {
				under_way = stage_issue(a, d, en, stage, mbar, lane);
				staged_for = c + 2u;
			}
		}
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
			/* dequantise (set_pos + midbuf, decode.c:174-177, :591-600), times 2^QS: x[i] = word m = 32 i + lane */
			uint32_t x[64];
#pragma unroll
			for (int p = 0; p < 4; p++) {
#pragma unroll
				for (int r = 0; r < ROWS; r++)
					x[4 * r + p] = nib_val(r < 8 ? A0[p] : A1[p], r & 7, e.val << QS);
			}
			if (linany) {
				/* pick up the parked linear columns; real branches (a loop of one turn per pass with
				 * a linear column): predicated, this would be 128 instructions for every block */
#pragma unroll
				for (int p = 0; p < 4; p++) {
					const bool isl = (linmask >> p) & 1u;
#pragma unroll 1
					for (uint32_t t = (linany >> p) & 1u; t; t--) {
#pragma unroll
						for (int r = 0; r < ROWS; r++) {
							const uint32_t lv = wb[32 * (4 * r + p) + lane];
							x[4 * r + p] = isl ? lv : x[4 * r + p];
						}
					}
				}
			}
			PROF_MARK(3);
			unsigned long long c2 = juggle_and_store<CKS>(x, h, wb, gh, lane, out, pos, n, a.fmt);
			PROF_MARK(4);
			pos += n;
			if (CKS) {
				for (int o = 16; o; o >>= 1)
					c2 += __shfl_xor_sync(0xFFFFFFFFu, c2, o);
				cks += c2;
			}
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
			for (size_t k = lane; k < nb; k += 32)
				p0[k] = 0;
			__syncwarp();
			if (lane == 0) {
				a.status[d.index] = st;
				a.words[d.index] = pos;
				a.cks[d.index] = a.fmt.checksums ? cks : 0ull;
				vol_st(&ctl->dead, e.desc + 1u); /* the slot's scan lane stops walking this stream */
			}
			dead = e.desc + 1u;
		}
		__syncwarp();
	}
	if (under_way) { /* not reached: a copy is only issued when the visit goes on */
		mbar_wait(mbar, mphase);
		mphase ^= 1u;
	}
	PROF_MARK(5);
	PROF_FLUSH(24);
}

/* ------------------------------------------------------------------ kernel */

/*
 * Scan CTA: SW scan warps, lane = stream slot (global slot id g = blockIdx * SLOTS + 32 * warp +
 * lane, slots >= a.n_slots stay empty).  A scan CTA has no decode work on its SM: the walk is a
 * dependent chain of shared-memory lookups, and sharing the SM's load/store pipe and issue slots
 * with decode warps doubled its step time (profiles/r01_ncu_fast2.md).
 */
__device__ __forceinline__ void scan_cta(const KernelArgs &a, SmemScan &sm, int warp, int lane)
{
	const uint32_t g = (uint32_t)blockIdx.x * 32u * a.scan_warps + 32u * (uint32_t)warp + (uint32_t)lane;
	const bool enabled = g < a.n_slots;
	SlotCtl *const ctl = reinterpret_cast<SlotCtl *>(a.slotctl) + g;
	uint8_t *const slot_ring = a.ring + (size_t)g * RING_D * REC_BYTES;
	ScanRing ring;
	const uint32_t lane4 = 4u * (uint32_t)(warp * 32 + lane);
	ring.rw = &sm.ring[0][warp * 32 + lane];
	ring.saddr = (uint32_t)__cvta_generic_to_shared(ring.rw);
	ring.pol = l2_keep_policy();
	ring.idle();
	bool active = false, exhausted = !enabled, first = true;
	uint32_t cur = 0, P = 0, blk = 0, limit = 0, n_attempt = 0, prodn = 0;
	unsigned long long waiting_since = 0ull;
	uint32_t seen_hb = 0, wait_ns = 250u;
	uint4 c4n = make_uint4(0u, 0u, 0u, 0u);
	bool fresh = false;
	PROF_DECL;
#if F2_PROF
	uint32_t prof_rounds = 0, prof_periods = 0;
#endif
	for (;;) {
		PROF_MARK(0); /* 0: round tail (publish) */
		uint32_t cons = 0, dead = 0;
		if (enabled) {
			/* never a stale line; after a walk, the copy fetched while walking */
			const uint4 c4 = fresh ? c4n : __ldcv(reinterpret_cast<const uint4 *>(ctl));
			cons = c4.z;
			dead = c4.w;
		}
		fresh = false;
		if (active && dead == cur + 1u) {
			active = false; /* the decode side found a corrupt t-code: abandon the stream */
			ring.idle();
			P = 0;
		}
		if (!active && !exhausted) {
			/* a slot's first stream is queue entry g (the host decides which streams share a warp and
			 * which warps share a sub-partition: plan_create); the rest of the queue goes to whoever is
			 * free first */
			const uint32_t idx = first ? g : a.n_slots + atomicAdd(a.counter, 1u);
			first = false;
			if (idx < a.count) {
				const DevStream d = a.streams[idx];
				cur = idx;
				P = d.bit0;
				blk = 0;
				limit = d.file_end + 8u;
				n_attempt = d.n_attempt;
				ring.start(a.blob + d.base_off, a.blob_room > d.base_off ? a.blob_room - d.base_off : 0,
					   d.file_end, P);
				active = true;
			} else {
				exhausted = true;
			}
		}
		const uint32_t lead = prodn - cons;
		const bool can = active && lead < (uint32_t)RING_D;
		{
			/* nothing to walk: every lane is out of streams (done), or every lane with a stream is
			 * RING_D records ahead; nobody close to starving the decode: let lanes bunch up */
			const bool none = !__any_sync(0xFFFFFFFFu, can);
			if (none && !__any_sync(0xFFFFFFFFu, active))
				break;
			if (none || (F2_HYST && !__any_sync(0xFFFFFFFFu, can && lead < (uint32_t)(RING_D / 2)))) {
				/* back off: while the decode side is the bottleneck a scan warp spends most of its time
				 * here, and a short sleep makes that a quarter of the kernel's executed instructions */
				__nanosleep(wait_ns);
				wait_ns = wait_ns < 2000u ? wait_ns * 2u : wait_ns;
				PROF_MARK(1); /* 1: waiting for the decode side */
				const uint32_t hb = *reinterpret_cast<volatile uint32_t *>(a.scan_done + 1);
				if (waiting_since == 0ull || hb != seen_hb) {
					waiting_since = now_ns();
					seen_hb = hb;
				} else if (now_ns() - waiting_since > WATCHDOG_NS) {
					if (lane == 0)
						atomicExch(a.errflag, 2u); /* the decode side does not consume anything */
					break;
				}
				continue;
			}
		}
		waiting_since = 0ull;
		wait_ns = 250u;
		if (lane == 0)
			atomicAdd(a.scan_done + 1, 1u); /* heartbeat */
		if (enabled) {
			/* the control words for the next round: an L2 round trip that the walk hides (what it
			 * may miss is one more consumed record or a stop request: both wait a round) */
			c4n = __ldcv(reinterpret_cast<const uint4 *>(ctl));
			fresh = true;
		}
		PROF_MARK(3); /* 3: round head (retire / acquire) */
#if F2_PROF
		prof_rounds++;
#endif
		/* ---- one record per lane that can produce */
		uint8_t *const recbase = slot_ring + (size_t)(prodn % RING_D) * REC_BYTES;
		Rec e;
		e.pblock = P; e.pend = P; e.desc = cur; e.blk = blk; e.status = SCAN_EOF; e.ncols = 0; e.val = 0; e.pad = 0;
		Walk s;
		s.Q = active ? P - 1u : 0u; /* P = 0 is a position like any other (Q wraps); idle lanes: ring rows 0 and 1 (zeros) */
		s.Q32 = s.Q << 5;
		s.s8 = UNI_HALT8;
		s.msk = MSK_K;
		int mode = 0; /* 0 not walking (any more), 1 block header pending, 2 walking */
		bool walk = false;
		if (can) {
			if (blk >= n_attempt) {
				/* nothing (more) to attempt: clean end */
				e.blk |= 0x80000000u;
			} else {
				walk = true;
				mode = 1;
			}
		}
		uint16_t *const off0 = reinterpret_cast<uint16_t *>(&sm.off[warp][lane * OFFP]);
		uint32_t cp = (uint32_t)__cvta_generic_to_shared(off0); /* shared-space address of the next offset to note */
		const uint32_t cpend = cp + 2u * COLS;
		bool hdr_eof = false;
		while (__any_sync(0xFFFFFFFFu, mode != 0)) {
			PROF_MARK(4); /* 4: walk steps */
#if F2_PROF
			prof_periods++;
#endif
			ring.topup(s.Q + 1u);
			if (mode == 1) {
				/* pwr(4) / val(16): GET_BITS_EXPECT_EOF decode.c:588-589 */
				if (s.Q + 21u > limit) {
					hdr_eof = true;
					mode = 0;
				} else if (s.Q + 1u <= ring.ready_p) {
					const uint32_t *rp = ring_word(&sm.ring[0][0], lane4, s.Q32);
					const uint32_t w1 = fsr(rp[0], rp[RROW], s.Q);
					e.val = (int)((w1 >> 5) & 0xFFFFu);
					s.Q += 20u;
					s.Q32 = s.Q << 5;
					s.s8 = 0u;
					s.msk = MSK_SEL;
					mode = 2;
				}
			}
			PROF_MARK(5); /* 5: top-up + header */
#pragma unroll STEP_UNROLL
			for (int k = 0; k < SCAN_PERIOD; k++)
				fast_step(s, cp, cpend, P - 1u, &sm.ring[0][0], lane4, ring.ready_p,
					  reinterpret_cast<const unsigned char *>(sm.uni16));
			if (mode == 2 && (s.s8 == UNI_HALT8 || s.s8 == UNI_BAD8))
				mode = 0;
		}
		PROF_MARK(4);
		if (walk) {
			if (hdr_eof) {
				e.status = SCAN_EOF;
			} else if (s.s8 == UNI_HALT8 && s.Q + 1u <= limit) {
				/* 128 columns, every read inside the stream */
				e.status = SCAN_OK;
				e.ncols = COLS;
				e.pend = s.Q + 1u;
			} else {
				/* bad selector, or the stream ended inside the block: walk it again with the
				 * reference's verdicts (rare: at most once per stream) */
				const DevStream d = a.streams[cur];
				BitReader br;
				atomicAdd(a.prof + 32, 1ull); /* always counted (tests: a healthy stream never gets here) */
				br.init(reinterpret_cast<const uint32_t *>(a.blob + d.base_off), d.file_end);
				const ScanResult sc = scan_block(br, P, limit, (uint32_t)COLS, (uint32_t)ROWS, off0, P,
								 a.tables->kind, a.tables->k8);
				e.status = sc.status;
				e.ncols = sc.ncols;
				e.pend = sc.end;
				e.val = sc.val;
				s.Q = sc.end - 1u;
				s.Q32 = s.Q << 5;
			}
			P = s.Q + 1u;
			blk++;
			if (e.status != SCAN_OK || blk >= n_attempt)
				e.blk |= 0x80000000u;
		}
		/* ---- the walked blocks' column offsets leave as 256-byte rows: thread t copies columns
		 * 4t..4t+3 of lane i's block */
		{
			const unsigned wm = __ballot_sync(0xFFFFFFFFu, walk);
			const unsigned long long rec64 = (unsigned long long)(uintptr_t)recbase;
			__syncwarp();
			for (unsigned mm = wm; mm; mm &= mm - 1u) {
				const int i = __ffs((int)mm) - 1;
				const unsigned long long r = __shfl_sync(0xFFFFFFFFu, rec64, i);
				const uint2 v = *reinterpret_cast<const uint2 *>(&sm.off[warp][i * OFFP + 8 * lane]);
				reinterpret_cast<uint2 *>((uintptr_t)r)[lane] = v;
			}
			__syncwarp();
		}
		if (can) {
			uint4 *rp = reinterpret_cast<uint4 *>(recbase + 256);
			rp[0] = make_uint4(e.pblock, e.pend, e.desc, e.blk);
			rp[1] = make_uint4((uint32_t)e.status, e.ncols, (uint32_t)e.val, 0u);
			if (e.blk >> 31) {
				active = false;
				ring.idle();
				P = 0;
			}
		}
		/* the records are in L2 before they are announced to the decode CTA (another SM) */
		__threadfence();
		if (can) {
			/* one 8-byte store: the count, and how far the stream still has to go -- the decode
			 * warp serves the slot with the longest way to go first, so that the batch's longest
			 * streams (its critical path) never wait for ring space */
			const uint32_t rem = active ? n_attempt - blk : 0u;
			vol_st2(&ctl->prod, ++prodn, rem);
		}
	}
	__threadfence();
	PROF_FLUSH(0);
#if F2_PROF
	if (lane == 0) {
		/* 16: busiest scan warp (cycles in steps + top-up), 17: its rounds, 18: all rounds, 19: all periods */
		atomicMax(a.prof + 16, prof_acc[4] + prof_acc[5]);
		atomicMax(a.prof + 17, (unsigned long long)prof_rounds);
		atomicAdd(a.prof + 18, (unsigned long long)prof_rounds);
		atomicAdd(a.prof + 19, (unsigned long long)prof_periods);
	}
#endif
	if (lane == 0)
		atomicAdd(a.scan_done, 1u);
}

/*
 * Decode CTA number w of n_work owns the slots g = w, w + n_work, w + 2 n_work, ... (local index
 * j < MAXOWN); any of its worker warps may decode any of them, one warp per slot at a time (a
 * slot's blocks depend on each other through the transform history).  Looking for work is one
 * 16-byte load per lane and owned slot group (prod and cons sit side by side in SlotCtl) and one
 * shared-memory atomic for the 128 lock bits; the slot's running state (words delivered,
 * checksum) lives in the second half of its SlotCtl in global memory.  A worker takes the free
 * slot with the largest backlog and decodes up to KMAX of its records in order.
 */
template <bool CKS>
__device__ __forceinline__ void work_cta(const KernelArgs &a, SmemWork &sm, int warp, int lane)
{
	const uint32_t w = (uint32_t)blockIdx.x - a.n_scan, n_work = (uint32_t)gridDim.x - a.n_scan;
	const uint32_t J = a.n_slots > w ? (a.n_slots - w + n_work - 1u) / n_work : 0u; /* the CTA's slots */
	SlotCtl *const ctl = reinterpret_cast<SlotCtl *>(a.slotctl);
	uint32_t *wb = sm.wb[warp];
	const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&sm.mbar[warp]);
	uint32_t mphase = 0u;
	uint32_t nap = 64u;
	bool confirmed = false;
	unsigned long long waiting_since = 0ull;
	uint32_t seen_hb = 0;
	PROF_DECL;
	if (J == 0u)
		return;
	for (;;) {
		PROF_MARK(0); /* 8+0: decode */
		const bool done = *reinterpret_cast<volatile uint32_t *>(a.scan_done) == a.n_scan * a.scan_warps;
		/* the free owned slot with the largest backlog: best = backlog << 8 | local index (0: none).
		 * Largest first: the slots of the longest streams are the ones that fall behind while
		 * decoding is the bottleneck, and whatever backlog they have when the scan ends is the
		 * launch's tail. */
		const uint32_t bw = atomicOr(&sm.busy[lane & (MAXOWN / 32 - 1)], 0u);
		uint32_t best = 0u;
#pragma unroll
		for (int k = 0; k < MAXOWN / 32; k++) {
			const uint32_t j = (uint32_t)lane + 32u * k;
			const uint32_t busyk = __shfl_sync(0xFFFFFFFFu, bw, k);
			if (j < J && !((busyk >> lane) & 1u)) {
				const uint4 c4 = __ldcv(reinterpret_cast<const uint4 *>(&ctl[w + n_work * j]));
				const uint32_t backlog = c4.x - c4.z;
				const uint32_t key = backlog ? ((backlog < 0xFFFFu ? backlog : 0xFFFFu) << 8) | j : 0u;
				best = key > best ? key : best;
			}
		}
		best = __reduce_max_sync(0xFFFFFFFFu, best);
		if (!best) {
			if (done) {
				/* every scan warp has finished: look once more behind a fence (acquire: all
				 * their announcements are visible), then leave */
				if (confirmed) {
					PROF_FLUSH(8);
					break;
				}
				__threadfence();
				confirmed = true;
				continue;
			}
			__nanosleep(nap); /* idle: back off */
			nap = nap < 1024u ? nap * 2u : nap;
			PROF_MARK(1); /* 8+1: idle */
			const uint32_t hb = *reinterpret_cast<volatile uint32_t *>(a.scan_done + 1);
			if (waiting_since == 0ull || hb != seen_hb) {
				waiting_since = now_ns();
				seen_hb = hb;
			} else if (now_ns() - waiting_since > WATCHDOG_NS) {
				if (lane == 0)
					atomicExch(a.errflag, 3u); /* no record, no end of scan, nobody moving */
				break;
			}
			continue;
		}
		nap = 64u;
		confirmed = false;
		waiting_since = 0ull;
		const uint32_t j = best & 255u, bit = 1u << (j & 31u);
		uint32_t got = 0;
		if (lane == 0)
			got = (atomicOr(&sm.busy[j >> 5], bit) & bit) == 0u;
		got = __shfl_sync(0xFFFFFFFFu, got, 0);
		if (!got)
			continue; /* another warp was quicker */
		PROF_MARK(2); /* 8+2: claim */
		const uint32_t g = w + n_work * j;
		/* acquire: the slot's state as the previous holder left it, and the records behind prod */
		uint32_t c = ld_acquire(&ctl[g].cons);
		uint32_t nrec = ld_acquire(&ctl[g].prod) - c;
		uint32_t ps = __ldcv(&ctl[g].pos);
		uint32_t dd = __ldcv(&ctl[g].dead);
		unsigned long long ck = __ldcv(&ctl[g].cks);
		if (nrec > (uint32_t)KMAX)
			nrec = KMAX;
		if (lane == 0)
			atomicAdd(a.scan_done + 1, 1u); /* heartbeat */
		decode_visit<CKS>(sm, a, wb, mbar, mphase, lane, a.ring + (size_t)g * RING_D * REC_BYTES, c, nrec,
				  a.hist + (size_t)g * HIST_WORDS, ctl + g, ps, ck, dd);
		c += nrec;
		__syncwarp();
		if (lane == 0) {
			ctl[g].pos = ps;
			ctl[g].cks = ck;
			/* release: the state above, and the records have been read before the scan lane may
			 * overwrite them */
			st_release(&ctl[g].cons, c);
			__threadfence_block();
			atomicAnd(&sm.busy[j >> 5], ~bit);
		}
		__syncwarp();
	}
}

template <bool CKS>
__global__ void __launch_bounds__(THREADS, 1) acm_decode_fast2_kernel(KernelArgs a)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

	if (blockIdx.x < a.n_scan) {
		SmemScan &sm = *reinterpret_cast<SmemScan *>(smem_raw);
		for (int i = tid; i < ACM_UNI_PAGES * ACM_UNI_PSIZE / 2; i += THREADS)
			reinterpret_cast<uint32_t *>(sm.uni16)[i] = reinterpret_cast<const uint32_t *>(a.tables->uni16)[i];
		for (int i = tid; i < SW * (RW + 1) * 32; i += THREADS)
			(&sm.ring[0][0])[i] = 0u;
		__syncthreads();
		if (warp < (int)a.scan_warps)
			scan_cta(a, sm, warp, lane);
	} else {
		SmemWork &sm = *reinterpret_cast<SmemWork *>(smem_raw);
		for (int i = tid; i < ACM_K8_SIZE; i += THREADS)
			sm.k8w[i] = a.tables->k8w[i];
		for (int i = tid; i < ACM_T_SIZE; i += THREADS)
			sm.t[i] = a.tables->t[i];
		if (tid < 32)
			sm.info[tid] = a.tables->kind[tid];
		if (lane == 0)
			mbar_init((uint32_t)__cvta_generic_to_shared(&sm.mbar[warp]), 1u);
		if (tid < MAXOWN / 32)
			sm.busy[tid] = 0u;
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		__syncthreads();
#ifdef F2_TEST_NO_DECODE
		return; /* watchdog test: the scan side must give up on its own */
#endif
		work_cta<CKS>(a, sm, warp, lane);
	}
}

} // namespace fast2

bool fast_shape(uint32_t level, uint32_t rows) { return level == fast2::LEVEL && rows == (uint32_t)fast2::ROWS; }

size_t fast2_smem_bytes() { return fast2::SMEM_BYTES; }

/*
 * Grid geometry for `count` streams on a device with `sms` SMs, using at most max_ctas of them:
 * n_scan scan CTAs (fast2::SLOTS stream slots each) followed by n_work decode CTAs (one CTA per
 * SM: all CTAs of a launch must be resident together, decode CTAs wait for records; the scan CTAs
 * have the lowest block indices, so they are placed first).  n_slots = slots in use.
 */
void fast2_geometry(uint64_t count, int sms, int max_ctas, uint32_t *n_scan, uint32_t *n_work, uint32_t *n_slots,
		    int walk_bound, uint32_t *scan_warps)
{
#ifndef F2_SCAN_PCT
#define F2_SCAN_PCT 22
#endif
	int total = max_ctas < sms ? max_ctas : sms;
	if (total < 2)
		total = 2;
	uint64_t want_scan = (count + fast2::SLOTS - 1) / fast2::SLOTS;
	uint32_t sw = (uint32_t)fast2::SW;
	/* A batch whose longest stream takes as long to WALK as the whole batch takes to decode (few, long
	 * streams: BASELINE configs[1]) gets more scan CTAs: with a lane for every stream from the start no
	 * lane is refilled, the scan warps thin out as their shorter streams end, and a round of a warp with
	 * few busy lanes is up to 30 % shorter -- which is what the longest streams, the launch's critical
	 * path, see (4.67 -> 4.30 ms on config 2).  A decode-bound batch (many short streams) keeps the
	 * smaller share: there every SM taken from the decode side costs throughput (19.1 -> 19.8 ms). */
#ifndef F2_SCAN_PCT_WALK
#define F2_SCAN_PCT_WALK 28
#endif
	int pct_walk = F2_SCAN_PCT_WALK;
	if (const char *e = getenv("ACM_B200_SCAN_PCT_WALK")) /* tuning */
		pct_walk = atoi(e);
	uint32_t cap_scan = (uint32_t)((total * (walk_bound ? pct_walk : F2_SCAN_PCT) + 50) / 100);
	if (cap_scan < 1)
		cap_scan = 1;
	/* A walk-bound batch small enough for it gets its lanes on ONE scan warp per SM sub-partition (the lower
	 * half of twice as many scan CTAs): a scan warp that shares its sub-partition steps a third slower, and
	 * here the SMs are there for the taking (a segment of the host path, a small resident batch). */
	static const bool sparse_ok = !(getenv("ACM_B200_SCAN_SPARSE") && atoi(getenv("ACM_B200_SCAN_SPARSE")) == 0);
	if (walk_bound && sparse_ok && scan_warps) {
		const uint64_t want_half = (count + fast2::SLOTS / 2 - 1) / (fast2::SLOTS / 2);
		if (want_half <= cap_scan) {
			sw = (uint32_t)fast2::SW / 2u;
			want_scan = want_half;
		}
	}
	if (scan_warps)
		*scan_warps = sw;
	uint32_t ns = want_scan < cap_scan ? (uint32_t)want_scan : cap_scan;
	if (ns < 1)
		ns = 1;
	uint64_t slots = (uint64_t)ns * 32u * sw;
	if (slots > count)
		slots = (count + 31) / 32 * 32; /* whole warps */
	/* one decode CTA per ~16 slots, at least one, at most what is left of the budget */
	const uint64_t want_work = (slots + 15) / 16;
	const uint32_t cap_work = (uint32_t)total - ns;
	uint32_t nw = want_work < cap_work ? (uint32_t)want_work : cap_work;
	if (nw < 1)
		nw = 1;
	/* decode CTA w owns the slots w, w + nw, w + 2 nw, ..., and the host deals the longest streams
	 * of a batch to lane 0 of every scan warp (slots 0, 32, 64, ...): with an even nw those land on
	 * nw / gcd(32, nw) decode CTAs only (measured: 112 decode CTAs, 7 of them busy to the end, +25 %).
	 * An odd nw spreads them over all decode CTAs; the odd CTA out stays unused. */
	if (nw > 1 && (nw & 1u) == 0u)
		nw -= 1;
	if (slots > (uint64_t)nw * fast2::MAXOWN)
		slots = (uint64_t)nw * fast2::MAXOWN;
	*n_scan = ns;
	*n_work = nw;
	*n_slots = (uint32_t)slots;
}

/* Is a launch bound by the walk of its longest stream rather than by decode throughput?  Measured rates:
 * a lock-step scan round (one block per lane) takes ~37 us, a decode CTA finishes ~1.3 blocks per us. */
int fast2_walk_bound(uint64_t longest_blocks, uint64_t total_blocks, int sms, int max_ctas)
{
	const int total = max_ctas < sms ? max_ctas : sms;
	const double n_work = total * (100 - F2_SCAN_PCT) / 100.0;
	const double t_walk = (double)longest_blocks * 37.0, t_decode = (double)total_blocks / (n_work * 1.3);
	return t_walk * 1.25 > t_decode;
}

size_t fast2_hist_words_per_slot() { return fast2::HIST_WORDS; }

int fast2_scan_warps() { return fast2::SW; }

size_t fast2_ring_bytes_per_slot() { return (size_t)fast2::RING_D * fast2::REC_BYTES; }

size_t fast2_ctl_bytes_per_slot() { return sizeof(fast2::SlotCtl); }

cudaError_t launch_fast2(const KernelArgs &a, int n_ctas, cudaStream_t st)
{
	if (a.count == 0)
		return cudaSuccess;
	/* the opt-in shared-memory size is a per-device function attribute */
	static bool configured[2][64] = {};
	static int resident[2][64] = {};
	const size_t smem = fast2::SMEM_BYTES;
	const int v = a.fmt.checksums ? 1 : 0;
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess)
		return e;
	if (!configured[v][dev & 63]) {
		e = v ? cudaFuncSetAttribute(fast2::acm_decode_fast2_kernel<true>,
					     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
		      : cudaFuncSetAttribute(fast2::acm_decode_fast2_kernel<false>,
					     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess)
			return e;
		/* the scan CTAs and the decode CTAs of a launch wait for each other: the whole grid has to
		 * be resident at once.  Ask the runtime instead of assuming (one CTA of this size per SM);
		 * the in-kernel watchdog stays as the guard against a GPU shared with other work. */
		int per_sm = 0, sms = 0;
		e = v ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fast2::acm_decode_fast2_kernel<true>,
								    fast2::THREADS, smem)
		      : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fast2::acm_decode_fast2_kernel<false>,
								    fast2::THREADS, smem);
		if (e != cudaSuccess)
			return e;
		e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if (e != cudaSuccess)
			return e;
		resident[v][dev & 63] = per_sm * sms;
		configured[v][dev & 63] = true;
	}
	if (n_ctas > resident[v][dev & 63])
		return cudaErrorCooperativeLaunchTooLarge; /* would deadlock: not all CTAs can be resident */
	if (v)
		fast2::acm_decode_fast2_kernel<true><<<n_ctas, fast2::THREADS, smem, st>>>(a);
	else
		fast2::acm_decode_fast2_kernel<false><<<n_ctas, fast2::THREADS, smem, st>>>(a);
	return cudaGetLastError();
}

} // namespace acm
