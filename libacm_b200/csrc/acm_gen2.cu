/*
 * acm_gen2.cu -- the GENERAL decode path: any level (cols = 1 << level), any row count, any
 * output format, as three kernels around a table of block records:
 *
 *   scan      acm_scan_kernel: one LANE per stream, which walks the stream block by
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
 *
 * Streams of level <= 10 (cols <= 1024: every shape the reference's own files have) take the
 * THROUGHPUT form of the decode stage, two kernels instead of acm_blocks_kernel:
 *
 *   unpack    acm_unpack_any_kernel: with the column positions known every COLUMN of every block
 *             is an independent unit of work; a CTA takes a run of blocks and spreads their columns
 *             over its threads (a 16-column block does not leave 7/8 of a CTA idle).  A column
 *             leaves as int16 quantiser indices (decode.c:174-177 without the midbuf multiply) in
 *             stream order: word m of the stream at inter16[word_base + m].
 *   lift      acm_lift_tile_kernel: juggle_block (decode.c:528-577) in flat form does not know
 *             about blocks at all: the stream is ONE sequence of words run through `level` 3-tap
 *             stages of delay C = cols/2 ... 1 (row parity = (m / C) & 1 and the "+1" of
 *             decode.c:561-564 = "m % (cols/2) == 0 in the first stage" hold for stream positions,
 *             because a block is a whole number of row pairs of every stage).  So the unit of work
 *             is a TILE of 4096 consecutive words of a stream, whatever its rows and cols, plus the
 *             2 * cols words before it that its history depends on: int16
 *             index x the block's multiplier (decode.c:591-600) into shared memory, the stages
 *             ping-pong between two shared buffers four words per thread (128-bit accesses), the
 *             last stage's words leave through output_values' shift / format (decode.c:617-677) as
 *             coalesced stores.  Tiles of one stream are independent: a long stream is transformed
 *             by many CTAs at once.
 */
#include "acm_walk.cuh"
#include "acm_kernels.cuh"

namespace acm {

namespace {

constexpr int G2_THREADS = 128;   /* decode CTA */
constexpr uint32_t G3_MAXLEVEL = GEN3_MAX_LEVEL; /* cols <= 1024: halo 2048 words */
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

/*
 * All lanes of a warp step their streams together: a step is "read the 32 bits at the lane's
 * position and act on the lane's state" -- block header, column selector, or one table step inside
 * a prefix-coded column (up to 7 rows per 8 bits, k8) -- so a warp's 32 walks cost what the slowest
 * lane's does, not their sum (one stream per lane with scan_block's loops, lanes diverging, was
 * measured at 265 us per block and lane).  No end-of-file checks on the way: bits past the end read
 * as zero (BitReader), and a block whose walk ends at or before the limit cannot have read past it;
 * the rare other block -- and one with a bad selector -- is walked again by scan_block, which has
 * the reference's verdicts (decode.c:108-135, :190-194).  Lanes take streams from a queue.
 */
struct SmemScan {
	TablesSmem tab;
	/* the first table step of a prefix-coded column, indexed by the 13 stream bits selector + first
	 * payload byte (the k8 entry of the selector's type; 0 for the other selectors): the step at a
	 * selector takes the column's first symbols along, without a second dependent lookup */
	uint32_t sel13[8192];
	uint32_t ring[walk::RW + 1 + 4][walk::RROW]; /* acm_walk.cuh: [word][warp][lane] */
};
constexpr int G2_SCAN_PERIOD = 16; /* steps between two top-ups of the rings */

__global__ void __launch_bounds__(32 * G2_SCAN_WARPS) acm_scan_kernel(KernelArgs a, Gen2Args g)
{
	extern __shared__ __align__(128) unsigned char scan_smem[];
	SmemScan &sm = *reinterpret_cast<SmemScan *>(scan_smem);
	TablesSmem &tab = sm.tab;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	load_tables(tab, a.tables, tid, 32 * G2_SCAN_WARPS);
	for (int i = tid; i < (walk::RW + 5) * walk::RROW; i += 32 * G2_SCAN_WARPS)
		(&sm.ring[0][0])[i] = 0u;
	__syncthreads();
	for (int i = tid; i < 8192; i += 32 * G2_SCAN_WARPS) {
		const uint32_t kind = tab.kind[i & 31];
		sm.sel13[i] = (kind & 7u) == ACM_CLS_K ? (uint32_t)tab.k8[(kind >> 3) * 256u + ((uint32_t)i >> 5)] : 0u;
	}
	__syncthreads();
	/* the lane's stream bits come from its shared-memory ring, topped up for all lanes together with
	 * 16-byte loads well ahead of the position (acm_walk.cuh): a warp-wide load straight from global
	 * memory waits for the slowest of 32 unrelated cache lines at every step */
	walk::Ring ring;
	const uint32_t lane4 = 4u * (uint32_t)(warp * 32 + lane);
	ring.rw = &sm.ring[0][warp * 32 + lane];
	ring.pol = walk::l2_keep_policy();
	ring.safe = a.blob;
	ring.hold_c0 = 0u;
	ring.idle();
	/* M_OVER / M_REWALK: the lane has reached the rare end of its walk and waits (at most a period) for the
	 * one place that deals with it, outside the unrolled steps */
	enum { M_NONE = 0, M_HDR = 1, M_SEL = 2, M_K = 3, M_OVER = 4, M_REWALK = 5 };
	/* A ring row has room for the lanes of eight warps (acm_walk.cuh), a scan CTA has four: the upper half of
	 * rows 0..31 is the lanes' selector tables, entry i of a lane in row i at the lane's own bank */
	static_assert(G2_SCAN_WARPS * 32 * 2 <= walk::RROW && walk::RW >= 32, "selector tables in the ring's unused half");
	uint32_t *const seltab = &sm.ring[0][0] + walk::RROW / 2 + warp * 32 + lane;
	constexpr uint32_t SELTAB_BYTE0 = (uint32_t)(walk::RROW / 2) * 4u; /* byte offset of that half within a row */
	int mode = M_NONE;
	bool exhausted = false;
	uint32_t si = 0, cols = 0, rows = 0, limit = 0, nmax = 0, P = 0, Pblk = 0, b = 0, c = 0, rem = 0, ktab = 0, val = 0;
	uint32_t t15_bits = 0, t27_bits = 0, t37_bits = 0;
	uint64_t rec_base = 0, coff_base = 0;
	uint32_t *cp = g.coff;
	BlockRec *recp = g.rec;
	BitReader br;
	br.init(nullptr, 0);
	for (;;) {
		/* ---- the rare end of a walk: the stream is over (all blocks walked, or no header left to read), a
		 * bad selector, a block that ends past the end of the stream.  Kept out of the steps below: four
		 * inlined copies of scan_block between them cost a branch per step and most of the instruction cache */
		if (mode >= M_OVER) {
			if (mode == M_OVER) {
				if (b >= nmax) {
					g.nscan[si] = b;
				} else {
					/* pwr / val cannot be read: GET_BITS_EXPECT_EOF, decode.c:588-589 */
					BlockRec r;
					r.P = P; r.end = P; r.val = 0; r.status = SCAN_EOF; r.ncols = 0; r.pad0 = r.pad1 = r.pad2 = 0u;
					g.rec[rec_base + b] = r;
					g.nscan[si] = b + 1u;
				}
				mode = M_NONE;
			} else {
				/* the stream's last block: the reference's verdict */
				const ScanResult sc = scan_block(br, Pblk, limit, cols, rows, g.coff + coff_base + (size_t)b * cols, 0u, tab.kind, tab.k8);
				BlockRec r;
				r.P = Pblk; r.end = sc.end; r.val = sc.val; r.status = sc.status; r.ncols = sc.ncols; r.pad0 = r.pad1 = r.pad2 = 0u;
				g.rec[rec_base + b] = r;
				if (sc.status == SCAN_OK) {
					/* cannot happen (the walk found the block's end past the limit, or a bad selector) */
					b++;
					P = sc.end;
					cp = g.coff + coff_base + (size_t)b * cols;
					recp = g.rec + rec_base + b;
					mode = M_HDR;
				} else {
					g.nscan[si] = b + 1u;
					mode = M_NONE;
				}
			}
			if (mode == M_NONE)
				ring.idle();
		}
		if (mode == M_NONE && !exhausted) {
			const uint32_t idx = atomicAdd(g.g3_counters + 2, 1u);
			if (idx < a.count) {
				const DevStream d = a.streams[idx];
				const Gen2Stream gs = g.gs[idx];
				si = idx;
				cols = 1u << d.level;
				rows = d.rows;
				t15_bits = ((rows + 2u) / 3u) * 5u; /* f_t15 / f_t27: 3 rows per code, f_t37: 2 (decode.c:405-476) */
				t27_bits = ((rows + 2u) / 3u) * 7u;
				t37_bits = ((rows + 1u) / 2u) * 7u;
				limit = d.file_end + 8u;
				nmax = d.n_attempt < gs.max_blocks ? d.n_attempt : gs.max_blocks;
				rec_base = gs.rec_base;
				coff_base = gs.coff_base;
				cp = g.coff + coff_base;
				recp = g.rec + rec_base;
				br.init((const uint32_t *)(a.blob + d.base_off), d.file_end);
				P = d.bit0;
				b = 0;
				ring.start(a.blob + d.base_off, a.blob_room > d.base_off ? a.blob_room - d.base_off : 0, d.file_end, P);
				/* the lane's selector table (see the step): what a selector advances the position by for THIS
				 * stream's row count -- f_zero 5 bits, f_linear 5 + rows * ind (decode.c:196-206), radix codes
				 * whole (3 rows per 5- or 7-bit code, 2 for f_t37: decode.c:405-476), prefix codes and bad
				 * selectors the 5 selector bits -- with the filler's class and sub-type above it */
				for (uint32_t i = 0; i < 32u; i++) {
					const uint32_t kind = tab.kind[i], cls = kind & 7u, sub = kind >> 3;
					uint32_t adv = 5u;
					if (cls == ACM_CLS_LINEAR)
						adv += rows * i;
					if (cls == ACM_CLS_T)
						adv += sub == 0u ? t15_bits : sub == 1u ? t27_bits : t37_bits;
					seltab[i * walk::RROW] = adv | (sub << 20) | (cls << 24);
				}
				mode = M_HDR;
			} else {
				exhausted = true;
			}
		}
		if (!__any_sync(0xFFFFFFFFu, mode != M_NONE))
			break;
#pragma unroll 4
		for (int step = 0; step < G2_SCAN_PERIOD; step++) {
			/* a lane whose bits have not landed yet does nothing this step, nor does one that waits for
			 * the rare path */
			const bool have = (uint32_t)(mode - M_HDR) <= (uint32_t)(M_K - M_HDR) && P <= ring.ready_p;
			const uint32_t *rp = walk::ring_word(&sm.ring[0][0], lane4, P << 5);
			const uint32_t w = walk::fsr(rp[0], rp[walk::RROW], P);
			/* ---- the common step, at a selector or inside a prefix-coded column: ONE straight-line
			 * instruction stream for both states (two table loads side by side, selects, a predicated
			 * store) -- with one warp per sub-partition every branch costs as much as four dependent
			 * ALU instructions, and lanes in different states would take turns.  What depends on "this
			 * lane takes part" is applied through masks the compiler cannot see through (it otherwise
			 * wraps the advance's arithmetic -- the head of the dependent chain -- in a branch) */
			/* a block header that can be read (pwr 4 bits, val 16: decode.c:588-589) is part of the common
			 * step too; the one that cannot -- the stream is over -- is the rare path's */
			const bool at_hdr = have && mode == M_HDR;
			const bool hdr_ok = at_hdr && b < nmax && P + 20u <= limit;
			const bool hdr = at_hdr && !hdr_ok;
			const bool act = have && mode >= M_SEL;
			const bool in_k = mode == M_K;
			uint32_t am = 0u - (uint32_t)act, hm = 0u - (uint32_t)hdr_ok;
			asm volatile("" : "+r"(am), "+r"(hm));
			const uint32_t e2 = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(&sm.ring[0][0]) +
									      (((w & 31u) * (4u * walk::RROW)) | (SELTAB_BYTE0 + lane4)));
			const uint32_t cls = e2 >> 24;
			const uint32_t e_k = reinterpret_cast<const uint32_t *>(tab.k8)[2u * (ktab + (w & 255u))];
			const uint32_t e_s = sm.sel13[w & 0x1FFFu];
			{
				/* column c's selector position (fill_block's "ind = get_bits(5)", decode.c:496); the
				 * offsets of a stream's blocks are one array: a running pointer */
				asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.u32 [%0], %1;\n\t}"
					     ::"l"(cp), "r"(P), "r"((uint32_t)(act && !in_k)) : "memory");
			}
			/* selector: the advance over the selector and a fixed-size payload, from the lane's table */
			const uint32_t adv_sel = e2 & 0xFFFFFu;
			/* prefix codes: whole symbols of the next 8 bits, up to the rows that remain -- inside a column
			 * from its type's table, at a selector (the column's first step) from the 13-bit table */
			const uint32_t e = in_k ? e_k : e_s;
			const uint32_t rem_in = in_k ? rem : rows;
			const uint32_t k = umin32(e & 15u, rem_in);
			const uint32_t adv_k = (e >> (4u * k)) & 15u;
			const bool is_k = in_k || cls == ACM_CLS_K;
			const uint32_t nrem = rem_in - k;
			const bool enter_k = !in_k && cls == ACM_CLS_K && nrem != 0u;
			const bool col_done = is_k ? nrem == 0u : true;
			const bool bad = act && !in_k && cls == ACM_CLS_BAD;
			const uint32_t adv_col = in_k ? adv_k : adv_sel + (is_k ? adv_k : 0u);
			const uint32_t next_mode = (enter_k || (in_k && nrem != 0u)) ? (uint32_t)M_K : (uint32_t)M_SEL;
			const uint32_t cd = (uint32_t)col_done & am;
			val = hdr_ok ? (w >> 4) & 0xFFFFu : val;
			Pblk = hdr_ok ? P : Pblk;
			c = hdr_ok ? 0u : c;
			P += (adv_col & am) | (20u & hm);
			c += cd;
			cp += cd;
			mode = (int)((next_mode & am) | ((uint32_t)M_SEL & hm) | ((uint32_t)mode & ~(am | hm)));
			ktab = (act && enter_k) ? (e2 >> 12) & 0x700u : ktab; /* sub-type * 256 */
			rem = (act && is_k) ? nrem : rem;
			/* the end of a block that ends inside the stream: its record (two predicated 16-byte stores) */
			const bool endblk = act && mode == M_SEL && c == cols;
			const bool end_ok = endblk && !bad && P <= limit;
			{
				const uint32_t on = end_ok ? 1u : 0u;
				asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %9, 0;\n\t"
					     "@p st.global.v4.u32 [%0], {%1,%2,%3,%4};\n\t"
					     "@p st.global.v4.u32 [%0+16], {%5,%6,%7,%8};\n\t}"
					     ::"l"(recp), "r"(Pblk), "r"(P), "r"(val), "r"((uint32_t)SCAN_OK), "r"(cols), "r"(0u), "r"(0u), "r"(0u), "r"(on)
					     : "memory");
			}
			b += end_ok ? 1u : 0u;
			recp += end_ok ? 1 : 0;
			mode = end_ok ? M_HDR : mode;
			/* the rare end of the walk: the lane stops here and is dealt with at the top of the loop */
			mode = hdr ? M_OVER : (bad || (endblk && !end_ok)) ? M_REWALK : mode;
		}
		ring.topup(P);
	}
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
		__syncthreads(); /* everyone has read it before the next one is written */
		if (it >= g.n_items)
			break;
		const Gen2Item item = g.items[it];
		const uint32_t si = item.stream;
		const DevStream d = a.streams[si];
		if (d.level <= G3_MAXLEVEL)
			continue; /* the unpack + tile-lift kernels' (uniform: decided from the item) */
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

/* ------------------------------------------------------------------ level <= 10: unpack + tile lift */

constexpr int G3_THREADS = 256;
constexpr uint32_t G3_TILE = GEN3_TILE_WORDS; /* output words per lift work item */
constexpr uint32_t G3_BUF = G3_TILE + 2048;
constexpr size_t G3_LIFT_SMEM = 2 * (size_t)G3_BUF * 4;

constexpr uint32_t G3_CHUNK = 2048;                 /* columns sorted together */
constexpr int G3_PER = G3_CHUNK / G3_THREADS;       /* ... per thread */
constexpr int G3_NCLS = 4;                          /* zero, linear, prefix-coded, radix-coded */
constexpr uint32_t G3_LIST = G3_CHUNK + 32 * G3_NCLS; /* every class padded to whole warps */

struct SmemUnpackAny {
	TablesSmem tab;
	uint32_t list[G3_LIST];   /* the chunk's columns sorted by filler class: Pc of the selector; 0xFFFFFFFF = nothing */
	uint16_t col[G3_LIST];    /* ... and the column's number within the run */
	uint32_t cnt[G3_NCLS], base[G3_NCLS + 1];
	uint32_t next;
	uint32_t item;
};

__global__ void __launch_bounds__(G3_THREADS) acm_unpack_any_kernel(KernelArgs a, Gen2Args g)
{
	int16_t *const inter16 = g.inter16;
	__shared__ SmemUnpackAny sm;
	TablesSmem &tab = sm.tab;
	const int tid = threadIdx.x, lane = tid & 31;
	load_tables(tab, a.tables, tid, G3_THREADS);
	__syncthreads();
	for (;;) {
		if (tid == 0)
			sm.item = atomicAdd(g.g3_counters, 1u);
		__syncthreads();
		const uint32_t it = sm.item;
		__syncthreads();
		if (it >= g.n_items)
			break;
		const Gen2Item item = g.items[it];
		const uint32_t si = item.stream;
		const DevStream d = a.streams[si];
		if (d.level > G3_MAXLEVEL)
			continue; /* acm_blocks_kernel's */
		const Gen2Stream gs = g.gs[si];
		const uint32_t level = d.level, cols = 1u << level, rows = d.rows;
		const uint32_t blen = rows * cols, limit = d.file_end + 8u;
		const uint32_t nscan = g.nscan[si];
		const uint32_t bend = item.b0 + item.nb < nscan ? item.b0 + item.nb : nscan;
		if (item.b0 >= bend)
			continue;
		BitReader br;
		br.init((const uint32_t *)(a.blob + d.base_off), d.file_end);
		int16_t *dst0 = inter16 + gs.word_base + (size_t)item.b0 * blen;
		const uint32_t ncol_run = (bend - item.b0) << level;
		/* The run's columns in chunks of 2048, each chunk sorted by filler class (counting sort in
		 * shared memory) so that a warp decodes 32 columns of ONE class: neighbouring columns are of
		 * unrelated types, and a warp that takes them as they come runs every filler's loop in turn
		 * (measured on the stress corpus: 7.9 of 32 lanes per instruction). */
		for (uint32_t c0 = 0; c0 < ncol_run; c0 += G3_CHUNK) {
			if (tid < G3_NCLS)
				sm.cnt[tid] = 0u;
			if (tid == 0)
				sm.next = 0u;
			__syncthreads();
			uint32_t mine[G3_PER], pcs[G3_PER]; /* class << 16 | rank; selector position */
#pragma unroll
			for (int k = 0; k < G3_PER; k++) {
				const uint32_t idx = c0 + (uint32_t)tid + (uint32_t)k * G3_THREADS;
				uint32_t cls = 7u, Pc = 0u;
				if (idx < ncol_run) {
					const uint32_t b = item.b0 + (idx >> level), c = idx & (cols - 1u);
					const BlockRec *rp = g.rec + gs.rec_base + b;
					const int status = __ldg(&rp->status);
					/* column ncols is included when its payload ran past the limit: a radix code that still
					 * fits may be out of range first (decode.c:412/:438/:464) */
					const uint32_t ncheck = __ldg(&rp->ncols) + (status == -7 ? 1u : 0u);
					if (c < ncheck) {
						Pc = __ldg(g.coff + gs.coff_base + (size_t)b * cols + c);
						const uint32_t kc = tab.kind[br.peek(Pc) & 31u] & 7u;
						cls = kc == ACM_CLS_ZERO ? 0u : kc == ACM_CLS_LINEAR ? 1u : kc == ACM_CLS_K ? 2u : kc == ACM_CLS_T ? 3u : 7u;
					}
				}
				/* rank within the class: one shared-memory atomic per class and warp */
				const unsigned peers = __match_any_sync(0xFFFFFFFFu, cls);
				const int leader = __ffs((int)peers) - 1;
				uint32_t base = 0;
				if (lane == leader && cls != 7u)
					base = atomicAdd(&sm.cnt[cls], (uint32_t)__popc(peers));
				base = __shfl_sync(0xFFFFFFFFu, base, leader);
				mine[k] = (cls << 16) | (base + (uint32_t)__popc(peers & ((1u << lane) - 1u)));
				pcs[k] = Pc;
			}
			__syncthreads();
			if (tid == 0) {
				uint32_t acc = 0;
				for (int c = 0; c < G3_NCLS; c++) {
					sm.base[c] = acc;
					acc += (sm.cnt[c] + 31u) & ~31u;
				}
				sm.base[G3_NCLS] = acc;
			}
			__syncthreads();
#pragma unroll
			for (int k = 0; k < G3_PER; k++) {
				const uint32_t cls = mine[k] >> 16, rank = mine[k] & 0xFFFFu;
				if (cls != 7u) {
					const uint32_t at = sm.base[cls] + rank;
					sm.list[at] = pcs[k];
					sm.col[at] = (uint16_t)(tid + k * G3_THREADS);
				}
			}
			if (tid < G3_NCLS) /* the holes at the end of every class */
				for (uint32_t i = sm.base[tid] + sm.cnt[tid]; i < sm.base[tid + 1]; i++)
					sm.list[i] = 0xFFFFFFFFu;
			__syncthreads();
			/* ---- a warp takes the next 32 columns of the sorted list */
			const uint32_t nwork = sm.base[G3_NCLS] >> 5;
			for (;;) {
				uint32_t w = 0;
				if (lane == 0)
					w = atomicAdd(&sm.next, 1u);
				w = __shfl_sync(0xFFFFFFFFu, w, 0);
				if (w >= nwork)
					break;
				const uint32_t Pc = sm.list[32u * w + (uint32_t)lane];
				if (Pc != 0xFFFFFFFFu) {
					const uint32_t idx = c0 + sm.col[32u * w + (uint32_t)lane];
					const uint32_t b = idx >> level, c = idx & (cols - 1u); /* b relative to the run */
					const uint32_t ind = br.peek(Pc) & 31u;
					const int r = decode_column(br, Pc + 5u, limit, ind, tab.kind[ind], rows, 1, dst0 + (size_t)b * blen + c, cols,
								    tab.k8, tab.t);
					if (r < 0)
						atomicMin(g.first_bad + si, item.b0 + b);
				}
			}
			__syncthreads(); /* the lists are free */
		}
	}
}

/* four consecutive outputs of a stage with C < 4 from the eight inputs v[0..7] = words j-4 .. j+3
 * (j = 4 * q: row parity and the "+1" follow from the word's position alone) */
template <int C>
__device__ __forceinline__ uint4 lift_small(const uint32_t (&v)[8], bool first_stage)
{
	uint32_t o[4];
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const uint32_t s = v[4 + k] + v[4 + k - 2 * C];
		uint32_t r = ((k / C) & 1) ? 2u * v[4 + k - C] - s : 2u * v[4 + k - C] + s;
		if (first_stage && (k % C) == 0)
			r += 1u; /* decode.c:561-564 */
		o[k] = r;
	}
	return make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(G3_THREADS) acm_lift_tile_kernel(KernelArgs a, Gen2Args g)
{
	const int16_t *const inter16 = g.inter16;
	const Gen2Item *const tiles = g.tiles;
	const uint32_t n_tiles = g.n_tiles;
	uint32_t *const tile_counter = g.g3_counters + 1;
	extern __shared__ __align__(16) uint32_t g3_smem[];
	__shared__ uint32_t s_item;
	uint32_t *bufA = g3_smem, *bufB = g3_smem + G3_BUF;
	const int tid = threadIdx.x;
	for (;;) {
		if (tid == 0)
			s_item = atomicAdd(tile_counter, 1u);
		__syncthreads();
		const uint32_t it = s_item;
		__syncthreads();
		if (it >= n_tiles)
			break;
		const Gen2Item item = tiles[it];
		const uint32_t si = item.stream;
		const DevStream d = a.streams[si];
		const Gen2Stream gs = g.gs[si];
		const uint32_t level = d.level, cols = 1u << level, blen = d.rows << level;
		/* words that are delivered: blocks up to the first one that fails (scan verdict or an
		 * out-of-range radix code), clipped to what the read loop delivers */
		const uint32_t ns = g.nscan[si], fb = g.first_bad[si];
		uint32_t nok = 0;
		if (ns) {
			const int last = __ldg(&g.rec[gs.rec_base + ns - 1].status);
			nok = last == SCAN_OK ? ns : ns - 1u;
		}
		nok = fb < nok ? fb : nok;
		const uint64_t w64 = (uint64_t)nok * blen;
		const uint32_t wend = w64 < d.words_limit ? (uint32_t)w64 : d.words_limit;
		const uint32_t t0 = item.b0 * G3_TILE; /* first output word of the tile */
		if (t0 >= wend)
			continue;
		const uint32_t t1 = t0 + G3_TILE < wend ? t0 + G3_TILE : wend;
		const uint32_t halo = level ? (2u * cols < t0 ? 2u * cols : t0) : 0u; /* a multiple of 4 (t0 is one of 4096) */
		const uint32_t m0 = t0 - halo;                 /* stream word at buffer position 0 */
		const uint32_t n = (t1 - m0 + 3u) & ~3u;       /* buffer words in use */
		const int16_t *src = inter16 + gs.word_base;
		const uint64_t cap = (uint64_t)gs.max_blocks * blen; /* words the intermediate holds */

		/* ---- load: index x the block's multiplier (midbuf, decode.c:591-600) */
		for (uint32_t j = 4u * tid; j < n; j += 4u * G3_THREADS) {
			const uint32_t m = m0 + j;
			uint32_t b = m / blen;
			uint32_t bnext = (b + 1u) * blen; /* first word of the next block */
			uint32_t val = b < ns ? (uint32_t)__ldg(&g.rec[gs.rec_base + b].val) : 0u;
			uint32_t x[4];
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const uint32_t mk = m + (uint32_t)k;
				if (mk >= bnext) {
					b++;
					bnext += blen;
					val = b < ns ? (uint32_t)__ldg(&g.rec[gs.rec_base + b].val) : 0u;
				}
				const int16_t idx = (uint64_t)mk < cap ? src[mk] : (int16_t)0;
				x[k] = (uint32_t)((int32_t)idx * (int32_t)val);
			}
			*reinterpret_cast<uint4 *>(bufA + j) = make_uint4(x[0], x[1], x[2], x[3]);
		}
		__syncthreads();

		/* ---- the stages: positions before the buffer read as zero -- exact for a tile at the start of
		 * its stream (decode.c:812: zero history), and only reaches the halo's own words otherwise */
		uint32_t *cur = bufA, *nxt = bufB;
		for (uint32_t l = 1; l <= level; l++) {
			const uint32_t C = cols >> l;
			if (C >= 4u) {
				const uint32_t sh = level - l; /* log2 C */
				for (uint32_t j = 4u * tid; j < n; j += 4u * G3_THREADS) {
					const uint4 x0 = *reinterpret_cast<const uint4 *>(cur + j);
					const uint4 z = make_uint4(0u, 0u, 0u, 0u);
					const uint4 x1 = j >= C ? *reinterpret_cast<const uint4 *>(cur + j - C) : z;
					const uint4 x2 = j >= 2u * C ? *reinterpret_cast<const uint4 *>(cur + j - 2u * C) : z;
					const uint32_t m = m0 + j;
					const bool odd = (m >> sh) & 1u;
					uint4 o;
					if (odd) {
						o.x = 2u * x1.x - (x0.x + x2.x); o.y = 2u * x1.y - (x0.y + x2.y);
						o.z = 2u * x1.z - (x0.z + x2.z); o.w = 2u * x1.w - (x0.w + x2.w);
					} else {
						o.x = 2u * x1.x + (x0.x + x2.x); o.y = 2u * x1.y + (x0.y + x2.y);
						o.z = 2u * x1.z + (x0.z + x2.z); o.w = 2u * x1.w + (x0.w + x2.w);
					}
					if (l == 1u && (m & (C - 1u)) == 0u)
						o.x += 1u; /* decode.c:561-564 */
					*reinterpret_cast<uint4 *>(nxt + j) = o;
				}
			} else {
				for (uint32_t j = 4u * tid; j < n; j += 4u * G3_THREADS) {
					const uint4 x0 = *reinterpret_cast<const uint4 *>(cur + j);
					const uint4 xp = j >= 4u ? *reinterpret_cast<const uint4 *>(cur + j - 4u) : make_uint4(0u, 0u, 0u, 0u);
					const uint32_t v[8] = { xp.x, xp.y, xp.z, xp.w, x0.x, x0.y, x0.z, x0.w };
					*reinterpret_cast<uint4 *>(nxt + j) = C == 2u ? lift_small<2>(v, l == 1u) : lift_small<1>(v, l == 1u);
				}
			}
			__syncthreads();
			uint32_t *t = cur; cur = nxt; nxt = t;
		}

		/* ---- output (decode.c:617-677), words [t0, t1) */
		uint8_t *out = a.out + d.out_off;
		unsigned long long cks = 0ull;
		for (uint32_t j = halo + 4u * tid; j < t1 - m0; j += 4u * G3_THREADS) {
			const uint4 q = *reinterpret_cast<const uint4 *>(cur + j);
			const uint32_t m = m0 + j;
			const uint32_t w[4] = { q.x, q.y, q.z, q.w };
			if (a.fmt.wordlen == 2 && m + 4u <= t1 && !a.fmt.checksums) {
				uint32_t u[4];
#pragma unroll
				for (int k = 0; k < 4; k++) {
					u[k] = ((uint32_t)((int32_t)w[k] >> level) + a.fmt.bias) & 0xFFFFu;
					if (a.fmt.be)
						u[k] = ((u[k] >> 8) | (u[k] << 8)) & 0xFFFFu;
				}
				/* out_off is 16-byte aligned and m a multiple of 4: an 8-byte aligned store */
				*reinterpret_cast<uint2 *>(out + (size_t)m * 2u) = make_uint2(u[0] | (u[1] << 16), u[2] | (u[3] << 16));
			} else {
#pragma unroll
				for (int k = 0; k < 4; k++) {
					if (m + (uint32_t)k < t1) {
						const uint32_t u = emit_word(out + (size_t)(m + k) * a.fmt.wordlen, (int32_t)w[k] >> level, a.fmt);
						if (a.fmt.checksums)
							cks += (unsigned long long)(m + (uint32_t)k + 1u) * (unsigned long long)(u + 1ull);
					}
				}
			}
		}
		if (a.fmt.checksums) {
			/* the finalise kernel adds up the blocks that were delivered: a tile's sum goes to its first
			 * block's slot (every block it touches is delivered, or none of its words were emitted) */
			for (int o = 16; o; o >>= 1)
				cks += __shfl_xor_sync(0xFFFFFFFFu, cks, o);
			if ((tid & 31) == 0 && cks)
				atomicAdd(&g.cks_blk[gs.rec_base + t0 / blen], cks);
		}
		__syncthreads(); /* the buffers are free */
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
	/* scan: one stream per LANE (a lane walks its stream on its own: the code is the scalar walk, the
	 * warp diverges, but 32 streams are in flight per warp instead of one) */
	static bool configured[64] = {};
	int dev = 0, sms = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess)
		return e;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	if (sms < 1)
		sms = 1;
	if (!configured[dev & 63]) {
		e = cudaFuncSetAttribute(acm_lift_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G3_LIFT_SMEM);
		if (e == cudaSuccess)
			e = cudaFuncSetAttribute(acm_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemScan));
		if (e != cudaSuccess)
			return e;
		configured[dev & 63] = true;
	}
	{
		/* the walk is latency bound: few streams are spread thin, many fill 16 warps per SM */
		unsigned scan_grid = (a.count + 32 * G2_SCAN_WARPS - 1) / (32 * G2_SCAN_WARPS);
		if (scan_grid > (unsigned)sms * 2u)
			scan_grid = (unsigned)sms * 2u; /* two CTAs' rings fit an SM */
		acm_scan_kernel<<<scan_grid, 32 * G2_SCAN_WARPS, sizeof(SmemScan), st>>>(a, g);
	}
	if (g.n_tiles) {
		const unsigned ugrid = g.n_items < (unsigned)sms * 8u ? g.n_items : (unsigned)sms * 8u;
		acm_unpack_any_kernel<<<ugrid, G3_THREADS, 0, st>>>(a, g);
		const unsigned lgrid = g.n_tiles < (unsigned)sms * 4u ? g.n_tiles : (unsigned)sms * 4u;
		acm_lift_tile_kernel<<<lgrid, G3_THREADS, G3_LIFT_SMEM, st>>>(a, g);
	}
	if (g.n_items && g.n_deep)
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
