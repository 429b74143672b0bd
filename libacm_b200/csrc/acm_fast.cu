/*
 * acm_fast.cu -- the throughput kernel for the common block shape: level 7
 * (128 columns), 16 rows, 2048 words per block (the shape of BASELINE configs 1, 2, 4).
 *
 * Why it looks the way it does (DESIGN.md section 4 has the numbers):
 *
 *  - A stream's bitstream is serial: where column c+1 starts is only known once column
 *    c has been walked.  Only the LENGTH walk is serial, though.  One
 *    "scan" warp per CTA runs it with one stream per lane (32 streams in flight per
 *    CTA), using the multi-symbol k8 table so a k-coded column costs ~3 steps, and
 *    publishes the 128 column offsets of each block through shared memory.
 *  - Everything else is parallel inside a block.  W "worker" warps each own S/W of the
 *    CTA's stream slots.  Per block a worker warp
 *      unpack   lane = column (4 passes of 32): filler dispatch, table decode,
 *               idx*val, store X0[row*128+col] -- the bank is the lane, conflict free;
 *      juggle   stages 1-2 (C=64,32) in registers: lane j owns every word m = j mod 32,
 *               which is exactly what it just unpacked; then one transpose through
 *               shared memory to contiguous ownership (lane j owns m in [64j,64j+64))
 *               and stages 3-7 (C=16..1) in registers over a 62-word halo that is
 *               recomputed instead of exchanged;
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

enum { ENT_IDLE = -100 };

struct Entry {
	uint32_t pblock; /* P of the block header */
	uint32_t desc;   /* index into the kernel's descriptor slice */
	uint32_t blk;    /* block number, bit 31 = last attempt of the stream */
	int32_t status;  /* SCAN_OK / SCAN_EOF / ACM_ERR_* / ENT_IDLE */
	uint32_t ncols;
	int32_t val;
};

struct Smem {
	uint64_t k8[ACM_K8_SIZE];
	uint32_t x[W][XWORDS];
	uint32_t hist0[S][128]; /* last 128 X0 words: [k*32+lane] = x[60+k] of that lane */
	uint32_t hist1[S][64];  /* last 64 X1 words: [k*32+lane] = y[62+k] */
	uint32_t hist2[S][64];  /* last 64 X2 words, flat order */
	unsigned long long cks[S];
	uint32_t pos[S];
	uint32_t dead[S];
	Entry ent[2][S];
	uint16_t coloff[2][COLS * OFF_PITCH];
	uint16_t t[ACM_T_SIZE];
	uint8_t kind[32];
	int more[2];
};

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
 * Transform + output of one block held in xs (X0, row-major [row*128+col]) by one warp.
 * n = words to emit (<= 2048).  Returns this lane's checksum contribution.
 */
template <bool CKS>
__device__ __forceinline__ unsigned long long
juggle_and_store(Smem &sm, uint32_t *xs, int slot, int lane, uint8_t *out, uint32_t pos0,
		 uint32_t n, const Format fmt)
{
	uint32_t *h0 = sm.hist0[slot], *h1 = sm.hist1[slot], *h2 = sm.hist2[slot];
	uint32_t x[64];
	unsigned long long cks = 0ull;

	/* ---- stages 1 and 2 in registers: lane owns m = 32*i + lane */
#pragma unroll
	for (int i = 0; i < 64; i++)
		x[i] = xs[32 * i + lane];
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
	__syncwarp(); /* every lane has read its X0 words: the buffer can be overwritten */
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

template <bool CKS>
__global__ void __launch_bounds__(THREADS, 1) acm_decode_fast_kernel(KernelArgs a)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

	for (int i = tid; i < ACM_K8_SIZE; i += THREADS)
		sm.k8[i] = a.tables->k8[i];
	for (int i = tid; i < ACM_T_SIZE; i += THREADS)
		sm.t[i] = a.tables->t[i];
	if (tid < 32)
		sm.kind[tid] = a.tables->kind[tid];
	if (tid < S) {
		sm.dead[tid] = 0;
		sm.pos[tid] = 0;
		sm.cks[tid] = 0ull;
	}
	if (tid < 2)
		sm.more[tid] = 0;
	__syncthreads();

	/* scan-lane state (meaningful in the scan warp only) */
	bool active = false;
	uint32_t cur = 0, P = 0, blk = 0, limit = 0, n_attempt = 0;
	BitReader sbr;
	sbr.init(nullptr, 0);

	for (int round = 0;; round++) {
		const int buf = round & 1;
		if (warp == W) {
			/* ================= scan warp: lane = stream slot ================= */
			Entry e;
			e.status = ENT_IDLE;
			e.pblock = 0; e.desc = 0; e.blk = 0; e.ncols = 0; e.val = 0;
			const bool has_slot = lane < S;
			if (active && sm.dead[lane < S ? lane : 0] == cur + 1u)
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
					sbr.init((const uint32_t *)(a.blob + d.base_off), d.file_end);
					active = true;
				}
			}
			if (active) {
				e.desc = cur;
				e.pblock = P;
				e.blk = blk;
				if (blk >= n_attempt) {
					/* nothing (more) to attempt: clean end */
					e.status = SCAN_EOF;
					e.blk |= 0x80000000u;
					active = false;
				} else {
					ScanResult sc = scan_block(sbr, P, limit, (uint32_t)COLS, (uint32_t)ROWS,
								   sm.coloff[buf] + lane, P, sm.kind, sm.k8,
								   OFF_PITCH);
					e.status = sc.status;
					e.ncols = sc.ncols;
					e.val = sc.val;
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
				BitReader br;
				br.init((const uint32_t *)(a.blob + d.base_off), d.file_end);
				int bad = 0;
				const uint16_t *offs = sm.coloff[pb] + slot;
#pragma unroll 1
				for (int p = 0; p < 4; p++) {
					const uint32_t c = 32u * p + lane;
					if (c < ncheck) {
						const uint32_t Pc = e.pblock + offs[c * OFF_PITCH];
						const uint32_t ind = br.peek(Pc) & 31u;
						int r = decode_column(br, Pc + 5u, limit_w, ind, sm.kind[ind], (uint32_t)ROWS,
								      e.val, xs + c, (uint32_t)COLS, sm.k8, sm.t);
						bad |= (r < 0);
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
					unsigned long long c2 = juggle_and_store<CKS>(sm, xs, slot, lane, out, pos, n, a.fmt);
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
}

} // namespace fast

bool fast_shape(uint32_t level, uint32_t rows) { return level == fast::LEVEL && rows == fast::ROWS; }

size_t fast_smem_bytes() { return sizeof(fast::Smem); }

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
