/*
 * acm_walk.cuh -- the column-length walk of 16-row blocks as a lock-step warp routine: one stream
 * per lane, the lane's next stream bits in a shared-memory ring, the whole walk one table-driven
 * state machine (uni16, acm_tables.h).  Used by the walk kernel of the split path (acm_split.cu).
 *
 * What is walked is fill_block's control flow (decode.c:491-502) without its data flow: where
 * every column's 5-bit selector sits.  The walk is serial per stream (where column c+1 starts is
 * only known once column c has been walked); everything after it is parallel.
 *
 * A step = fetch the 32 stream bits at the lane's position from its ring (two LDS + one funnel
 * shift), ONE uni16 lookup (page 0 = "at a selector", indexed by selector + first payload byte;
 * page (type, rows to come) = "inside a prefix-coded column"; entry = bits to advance | next
 * page), position += advance, and at a selector a predicated 2-byte store of the column's offset.
 * No branch on the data: lanes that are done, idle or corrupt sit on pages whose entries advance
 * 0 bits and stay.  The dependent chain is table LDS -> dp4a -> LOP3 -> ring LDS -> SHF -> LOP3:
 * the lane keeps Q = position - 1 (the 32 bits at Q, masked, are the entry's byte offset), 32 Q
 * alongside (the ring row is a mask of it; rows are 1 KB: [word][8 warps][32 lanes]).
 */
#pragma once

#include "acm_fast2_core.cuh"

namespace acm {
namespace walk {

using fast2::Walk;
using fast2::fsr;
using fast2::walk_index;
using fast2::walk_next_if;
using fast2::UNI_HALT8;
using fast2::UNI_BAD8;
using fast2::MSK_SEL;
using fast2::MSK_K;

constexpr int SW = 8;               /* lane groups (warps) a ring row has room for */
constexpr int RW = 64;              /* ring words per lane (+1 duplicate of word 0) */
constexpr int RROW = 32 * SW;       /* words per ring row: one word of every lane of the CTA */
#ifndef WALK_NHOLD
#define WALK_NHOLD 8
#endif
#ifndef WALK_PERIOD
#define WALK_PERIOD 32
#endif
constexpr int NHOLD = WALK_NHOLD;   /* 16-byte chunks a lane can take in per period */
constexpr int NSMALL = 2;           /* ... and how many of them a top-up handles without asking whether any lane wants more */
static_assert(NSMALL <= NHOLD, "top-up tiers");
constexpr int NSTART = 4;           /* chunks a lane loads synchronously when it takes a new stream */
constexpr int LEAD = RW / 4 - 2;    /* 16-byte chunks requested ahead of the read position */
constexpr int PERIOD = WALK_PERIOD; /* walk steps between two top-ups */
constexpr int MAXCOLS = 128;        /* columns whose offsets are staged per block */
constexpr int OFFP = 2 * MAXCOLS + 8; /* bytes per lane of the column-offset staging (+8: bank spread) */

static_assert(4 * RROW == 1024, "ring_word() masks 32 * position: rows are 1 KB");

struct SmemWalk {
	uint16_t uni16[ACM_UNI_PAGES * ACM_UNI_PSIZE];
	uint32_t ring[RW + 1 + 4][RROW]; /* [word][warp][lane]; row RW = copy of row 0, then four spare rows */
	unsigned char off[SW][32 * OFFP]; /* column offsets of the block a lane is walking (u16, relative to the block) */
};

__device__ __forceinline__ uint64_t l2_keep_policy()
{
	/* the compressed bytes are read again by the unpack kernel: ask the L2 to keep the lines */
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
__device__ __forceinline__ uint4 ldg_keep_v4(const void *p, uint64_t pol)
{
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
		     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
	return r;
}
__device__ __forceinline__ void prefetch_l2(const void *p)
{
	asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p));
}

/* the lane's ring word that holds the bit whose position times 32 is P32 */
__device__ __forceinline__ const uint32_t *ring_word(const uint32_t *ring0, uint32_t lane4, uint32_t P32)
{
	const uint32_t off = (P32 & ((RW - 1u) * 1024u)) | lane4;
	return reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(ring0) + off);
}

/*
 * A lane's view of its stream: word i (32-bit words from the 16-byte aligned stream base) at
 * ring[i % RW][32 * warp + lane], plus a copy of ring word 0 in row RW so that the pair (i, i+1)
 * is always (row, row + 1).  The ring is topped up for all lanes together every PERIOD steps, up
 * to LEAD chunks of 16 bytes ahead of the position and at most NHOLD chunks per period: a chunk
 * travels as ONE 16-byte load into registers and is written to the ring (four 4-byte rows) at the
 * next top-up, a period later, when the load has long landed; a 32-bit fetch at P is safe while
 * P < ready_p.  The line a kilobyte ahead is prefetched into the L2, so that the loads are L2 hits
 * even when a warp has its SM sub-partition to itself and a period is shorter than a trip to HBM.
 * The end-of-file rule (one zero byte, then nothing: decode.c:57-61) is applied to the rare chunk
 * that touches the end of the file: bytes at and past it are zeroed before they reach the ring.
 */
struct Ring {
	const uint32_t *rw;   /* this lane's ring word 0 */
	const uint8_t *base;  /* stream base (16-byte aligned) */
	const uint8_t *safe;  /* 16 readable bytes: what a lane that wants nothing loads */
	uint32_t room16;      /* 16-byte chunks readable at base */
	uint32_t full16;      /* chunks [0, full16) lie entirely inside the file and the blob */
	uint32_t fe_byte;     /* bytes of the stream that exist (relative to base) */
	uint32_t fill;        /* chunks [.., fill) have been asked for */
	uint32_t ready_p;
	uint4 hold[NHOLD];
	uint32_t hold_c0, hold_n;
	uint64_t pol;

	/* the end-of-file rule for a chunk that is not entirely inside the file */
	__device__ __forceinline__ uint4 trim(uint32_t c, uint4 v) const
	{
		if (c < full16)
			return v;
		if (!(c < room16 && c * 16u < fe_byte))
			return make_uint4(0u, 0u, 0u, 0u);
		const uint32_t n = fe_byte - c * 16u; /* 1 .. 15 bytes of the chunk exist */
		uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const uint32_t nb = n > 4u * k ? n - 4u * k : 0u;
			w[k] = nb >= 4u ? w[k] : nb ? w[k] & ((1u << (8u * nb)) - 1u) : 0u;
		}
		return make_uint4(w[0], w[1], w[2], w[3]);
	}
	/* no stream: the lane sits on the HALT page at P = 0 over two zero ring words */
	__device__ __forceinline__ void idle()
	{
		const_cast<uint32_t *>(rw)[0] = 0u;
		const_cast<uint32_t *>(rw)[RROW] = 0u;
		base = nullptr;
		hold_n = 0;
		room16 = 0;
		full16 = 0;
		fe_byte = 0;
		fill = 0x0FFFFFF0u;
		ready_p = 0;
	}
	/* all lanes together, every PERIOD steps; P = the lane's position.  No branches: a lane that has
	 * nothing to store writes the ring's spare rows, a lane that wants nothing loads `safe`. */
	__device__ __forceinline__ void topup(uint32_t P)
	{
		/* last period's loads go into the ring; nothing touches them before (a register copy right
		 * after the loads would wait out the memory latency in every period) */
		if (__any_sync(0xFFFFFFFFu, hold_n > 0u && hold_c0 + hold_n > full16)) {
			/* some lane is at the end of its file (once per stream) */
#pragma unroll
			for (int k = 0; k < NHOLD; k++)
				if ((uint32_t)k < hold_n)
					hold[k] = trim(hold_c0 + k, hold[k]);
		}
		/* Most periods no lane of the warp moves more than NSMALL chunks (a lane that walks short columns
		 * reads a few hundred bits per period): the chunks beyond them are skipped for the whole warp, one
		 * vote and one uniform branch instead of the stores and loads of NHOLD - NSMALL chunks nobody wants
		 * (the top-up was an eighth of the instructions of a warp whose walk bounds a launch) */
		const bool many_st = __any_sync(0xFFFFFFFFu, hold_n > (uint32_t)NSMALL);
#pragma unroll
		for (int k = 0; k < NHOLD; k++) {
			if (k >= NSMALL && !many_st)
				break;
			const uint32_t c = hold_c0 + k;
			const bool on = (uint32_t)k < hold_n;
			uint32_t *row = const_cast<uint32_t *>(rw) + (on ? (c & (RW / 4 - 1)) * (4u * RROW) : (RW + 1u) * RROW);
			row[0] = hold[k].x;
			row[RROW] = hold[k].y;
			row[2 * RROW] = hold[k].z;
			row[3 * RROW] = hold[k].w;
			if (on && (c & (RW / 4 - 1)) == 0)
				const_cast<uint32_t *>(rw)[RW * RROW] = hold[k].x;
		}
		ready_p = base && fill ? (fill * 4u - 1u) * 32u : 0u; /* everything asked for so far is in the ring */
		/* after the position has jumped past what was asked for (a long column), the chunks behind
		 * it are never read */
		const uint32_t c0 = P >> 7, f0 = fill < c0 ? c0 : fill;
		int n = (int)(c0 + LEAD) - (int)f0;
		n = n < NHOLD ? n : NHOLD;
		n = base && n > 0 ? n : 0;
		const bool many_ld = __any_sync(0xFFFFFFFFu, n > NSMALL);
#pragma unroll
		for (int k = 0; k < NHOLD; k++) {
			if (k >= NSMALL && !many_ld)
				break;
			hold[k] = ldg_keep_v4(k < n && f0 + (uint32_t)k < room16 ? base + (size_t)(f0 + k) * 16u : safe, pol);
		}
		prefetch_l2(base && f0 + 64u < room16 ? base + (size_t)(f0 + 64u) * 16u : safe);
		hold_c0 = f0;
		hold_n = (uint32_t)n;
		fill = f0 + (uint32_t)n;
	}
	/* a new stream: the first chunks synchronously (once per stream), the walk starts at once.
	 * (Leaving them to the top-ups -- the lane idles a period instead of the warp waiting for the
	 * loads -- decoded streams whose first data bit sits on a 128-bit boundary wrongly; not understood,
	 * not used.) */
	__device__ __forceinline__ void start(const uint8_t *src, uint64_t room, uint32_t file_end, uint32_t P0)
	{
		base = src;
		room16 = (uint32_t)(room >> 4);
		fe_byte = file_end >> 3;
		full16 = fe_byte >> 4 < room16 ? fe_byte >> 4 : room16;
		fill = P0 >> 7;
		hold_n = 0;
		uint4 v[NSTART];
#pragma unroll
		for (int k = 0; k < NSTART; k++)
			v[k] = ldg_keep_v4(fill + k < room16 ? base + (size_t)(fill + k) * 16u : safe, pol);
#pragma unroll
		for (int k = 0; k < NSTART; k++) {
			const uint32_t c = fill + k;
			const uint4 t = trim(c, v[k]);
			uint32_t *row = const_cast<uint32_t *>(rw) + (c & (RW / 4 - 1)) * (4u * RROW);
			row[0] = t.x;
			row[RROW] = t.y;
			row[2 * RROW] = t.z;
			row[3 * RROW] = t.w;
			if ((c & (RW / 4 - 1)) == 0)
				const_cast<uint32_t *>(rw)[RW * RROW] = t.x;
		}
		fill += NSTART;
		ready_p = (fill * 4u - 1u) * 32u;
	}
};

/*
 * One walk step for all lanes of a warp.  cp = shared-space address of the next column offset to
 * note, cpend = where the block's last one goes; qblock = Q of the block start (offsets are
 * relative to the block).  After the last column the lane moves to the HALT page.  A lane whose
 * bits have not landed yet (position >= ready_p) does nothing this step.  No end-of-file checks
 * here: bits past the end read as zero, and a block whose walk ends at or before the stream's
 * limit cannot have read past it (the caller re-walks the rare other case with the reference's
 * verdicts).
 */
__device__ __forceinline__ void step(Walk &s, uint32_t &cp, uint32_t cpend, uint32_t qblock, const uint32_t *ring0,
				     uint32_t lane4, uint32_t ready_p, const unsigned char *uni)
{
	const uint32_t *rp = ring_word(ring0, lane4, s.Q32);
	const uint32_t w1 = fsr(rp[0], rp[RROW], s.Q);
	const uint32_t e = *reinterpret_cast<const uint16_t *>(uni + walk_index(s, w1));
	const bool have = s.Q + 1u <= ready_p;
	const bool note = have && s.msk == MSK_SEL;
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u16 [%0], %1;\n\t}"
		     :: "r"(cp), "h"((unsigned short)(s.Q - qblock)), "r"((uint32_t)note) : "memory");
	cp += note ? 2u : 0u;
	const bool at_sel = walk_next_if(s, e, have);
	const bool done = at_sel && cp == cpend;
	s.s8 = done ? UNI_HALT8 : s.s8;
	s.msk = done ? MSK_K : s.msk;
}

} // namespace walk
} // namespace acm
