/*
 * acm_stream.cu -- the libacm.h drop-in surface (reference src/libacm.h:120-170) on top
 * of the CUDA decoder.
 *
 * Same signatures, argument meaning, return codes and bookkeeping as the reference's
 * acm_open_decoder / acm_read / acm_close (decode.c:758-893) and util.c (acm_open_file,
 * the info getters, acm_seek_pcm / acm_seek_time, acm_read_loop, acm_strerror).  What is
 * different is where decode_block (decode.c:580-611) runs: blocks are decoded on the
 * GPU in look-ahead chunks (resumable generic kernel: bit position + per-stage history
 * carried from chunk to chunk) into a host-side PCM cache that acm_read serves from.
 *
 *  - the image is pulled through read_func as the reference pulls it: (1, 65536) requests
 *    (decode.c:41-67; short reads are fine), one window at open for the header, and then
 *    ahead of every chunk as many windows as that chunk can consume at most (a block is at
 *    most 20 + cols * (5 + 16 * rows) bits).  What has been pulled stays cached on the host
 *    and on the device; after a backward seek -- one seek_func call to the start of the data,
 *    as util.c:223-228 -- the source is read again from there and the cached prefix skipped.
 *  - acm_read keeps the reference's contract to the letter: never crosses a block
 *    boundary, clips to total_values, rounds down to whole frames, dst == NULL decodes
 *    and discards (decode.c:826-876).  The PCM format is a per-call argument, so the
 *    cache remembers the format it holds and a chunk is simply decoded again if a
 *    caller switches formats mid-stream.
 *  - acm_seek_pcm follows util.c:214-253 (one seek_func call on a backward seek,
 *    ACM_ERR_NOT_SEEKABLE without it, forward skipping by acm_read(NULL)), but every
 *    chunk boundary visited so far is an index entry (block number, bit position,
 *    history snapshot), so a seek re-decodes at most one chunk instead of the whole
 *    prefix.
 *
 * Deviations, all outside what the reference defines: ACM_ERR_CORRUPT is sticky (the
 * reference has no resync and re-enters decode_block at an undefined position, SURVEY
 * Q17; after ACM_ERR_UNEXPECTED_EOF it reports a clean end of stream, and so do we);
 * acm_raw_tell reports the bytes consumed up to the end of the decoded look-ahead;
 * wordlen 3 and 4 are accepted (the reference returns ACM_ERR_BADFMT).
 * There is no CPU decode path: without a usable CUDA device acm_open_decoder fails
 * with ACM_ERR_OTHER.
 */
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

#include "acm_gpu.h"
#include "acm_host.h"
#include "acm_kernels.cuh"
#include "libacm.h"

using namespace acm;

namespace {

struct SavedState {
	uint32_t block;             /* state at the START of this block */
	uint32_t P;                 /* bit position (P coordinates) */
	std::vector<uint32_t> hist; /* 2*cols words; empty = all zero (block 0) */
};

struct GpuState {
	int device = 0;
	std::vector<uint8_t> file;   /* the image from byte 0, as far as it has been pulled */
	uint64_t src_pos = 0;        /* file offset the data source stands at */
	bool src_eof = false;        /* read_func has returned 0 */
	bool complete = false;       /* ... at the end of the cache: the whole image is here */
	size_t uploaded = 0;         /* bytes of `file` that are in d_blob */
	size_t blob_cap = 0;         /* bytes d_blob can hold */
	size_t len_hint = 0;         /* get_length_func's answer (0: unknown) */
	acm_header hdr{};
	DevStream base{};            /* descriptor of the whole stream (block 0 start) */
	uint32_t n_attempt_total = 0, words_limit = 0, chunk_blocks = 1, cols = 1;
	bool read_error = false;

	/* device side */
	uint8_t *d_blob = nullptr;
	uint64_t blob_room = 0;
	DevStream *d_desc = nullptr;
	int32_t *d_status = nullptr;
	uint32_t *d_words = nullptr;
	unsigned long long *d_cks = nullptr;
	uint32_t *d_counter = nullptr;
	uint32_t *d_endpos = nullptr;
	uint32_t *d_hist = nullptr;
	acm_tables *d_tables = nullptr;
	GenericScratch scratch{};
	uint8_t *d_pcm = nullptr;
	size_t pcm_cap = 0;
	cudaStream_t stream = nullptr;

	/* host chunk cache */
	std::vector<uint8_t> pcm;
	bool c_valid = false;
	int c_fmt = -1;
	uint32_t c_block0 = 0, c_attempted = 0, c_nok = 0, c_words = 0, c_endP = 0;
	int c_status = 0;            /* verdict of attempt c_block0 + c_nok when c_nok < c_attempted */

	/* split path (acm_split.cu; level 7, 16 rows, 16-bit formats): one lane walks the stream, the
	 * blocks of a chunk are unpacked and transformed in parallel.  Per-block intermediates for the
	 * whole stream: the records double as the block index (bit position of every walked block). */
	bool split = false;
	SplitArgs sp{};
	uint8_t *d_split = nullptr;
	Gen2Item *d_items = nullptr;
	int sm_count = 0;
	/* two chunk slots (device PCM, pinned host PCM / records / walk state): while acm_read serves one
	 * chunk, the next one is walked, decoded and copied behind it (look-ahead) */
	uint8_t *d_chunk[2] = { nullptr, nullptr };
	uint8_t *h_chunk[2] = { nullptr, nullptr };
	BlockRec *h_recs[2] = { nullptr, nullptr };
	uint32_t *h_state[2] = { nullptr, nullptr };
	int cur_slot = 0;
	cudaEvent_t slot_done[2] = { nullptr, nullptr }; /* recorded behind a chunk's last copy */
	struct Ahead {
		bool active = false;
		uint32_t b0 = 0, nb = 0, P0 = 0;
		int key = 0, slot = 0;
	} ahead;
	const uint8_t *pcm_ptr = nullptr; /* the cached chunk's PCM: pcm.data(), or a pinned chunk slot */

	std::vector<SavedState> index;  /* sorted by block; [0] is block 0 */
	uint32_t next_block = 0;        /* block acm_read decodes next when !block_ready */
	uint32_t cur_block = 0;         /* block the current block_pos refers to */
	uint32_t consumed_P = 0;        /* for acm_raw_tell */
	bool drained = false;           /* an UNEXPECTED_EOF was reported: the reference's bit reader is
					   empty after it, so every later decode_block is a clean EOF */
};

inline int fmt_key(int be, int wordlen, int sgned) { return (be ? 1 : 0) | (wordlen << 1) | (sgned ? 16 : 0); }

#define CUS(call)                                                                   \
	do {                                                                        \
		cudaError_t e_ = (call);                                            \
		if (e_ != cudaSuccess) {                                            \
			acm_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call,   \
				      cudaGetErrorString(e_));                      \
			return ACM_ERR_OTHER;                                       \
		}                                                                   \
	} while (0)


/*
 * Device and pinned-host buffers of closed streams are kept for the next open (cudaMalloc,
 * cudaHostAlloc and cudaFree are milliseconds each and synchronise the device: a stream's open and
 * close would cost more than decoding it).  Bounded; acm_gpu_release_workspace() empties it.
 */
struct PoolBuf {
	void *p;
	size_t cap;
	int dev;
	bool pinned;
};
std::mutex g_pool_mu;
std::vector<PoolBuf> g_pool_free;
std::unordered_map<void *, PoolBuf> g_pool_live;
size_t g_pool_bytes = 0;
constexpr size_t POOL_MAX_BYTES = (size_t)1 << 30;

cudaError_t pool_alloc(void **out, size_t size, bool pinned, int dev)
{
	const size_t want = (size + 255u) & ~(size_t)255u;
	{
		std::lock_guard<std::mutex> lk(g_pool_mu);
		size_t best = g_pool_free.size();
		for (size_t i = 0; i < g_pool_free.size(); i++) {
			const PoolBuf &b = g_pool_free[i];
			if (b.dev == dev && b.pinned == pinned && b.cap >= want && b.cap <= 4 * want + ((size_t)1 << 16) &&
			    (best == g_pool_free.size() || b.cap < g_pool_free[best].cap))
				best = i;
		}
		if (best != g_pool_free.size()) {
			const PoolBuf b = g_pool_free[best];
			g_pool_free.erase(g_pool_free.begin() + (long)best);
			g_pool_bytes -= b.cap;
			g_pool_live[b.p] = b;
			*out = b.p;
			return cudaSuccess;
		}
	}
	void *p = nullptr;
	const cudaError_t e = pinned ? cudaHostAlloc(&p, want, cudaHostAllocDefault) : cudaMalloc(&p, want);
	if (e != cudaSuccess)
		return e;
	std::lock_guard<std::mutex> lk(g_pool_mu);
	g_pool_live[p] = PoolBuf{ p, want, dev, pinned };
	*out = p;
	return cudaSuccess;
}

void pool_free(void *p)
{
	if (!p)
		return;
	PoolBuf b;
	{
		std::lock_guard<std::mutex> lk(g_pool_mu);
		auto it = g_pool_live.find(p);
		if (it == g_pool_live.end())
			return;
		b = it->second;
		g_pool_live.erase(it);
		if (g_pool_bytes + b.cap <= POOL_MAX_BYTES && g_pool_free.size() < 256) {
			g_pool_free.push_back(b);
			g_pool_bytes += b.cap;
			return;
		}
	}
	if (b.pinned)
		cudaFreeHost(b.p);
	else
		cudaFree(b.p);
}

} // namespace

/* empties the pool (acm_gpu_release_workspace) */
void acm_stream_release_pool()
{
	std::vector<PoolBuf> take;
	{
		std::lock_guard<std::mutex> lk(g_pool_mu);
		take.swap(g_pool_free);
		g_pool_bytes = 0;
	}
	for (const PoolBuf &b : take) {
		if (b.pinned)
			cudaFreeHost(b.p);
		else
			cudaFree(b.p);
	}
}

namespace {

/* the code tables on a device: built and uploaded once per process and device */
const acm_tables *device_tables(int dev)
{
	static std::mutex mu;
	static acm_tables *d_tab[64] = {};
	static acm_tables host;
	static bool built = false;
	std::lock_guard<std::mutex> lk(mu);
	if (dev < 0 || dev >= 64)
		return nullptr;
	if (!d_tab[dev]) {
		if (!built) {
			acm_tables_build(&host);
			built = true;
		}
		acm_tables *p = nullptr;
		if (cudaMalloc(&p, sizeof(acm_tables)) != cudaSuccess)
			return nullptr;
		if (cudaMemcpy(p, &host, sizeof(host), cudaMemcpyHostToDevice) != cudaSuccess) {
			cudaFree(p);
			return nullptr;
		}
		d_tab[dev] = p;
	}
	return d_tab[dev];
}

void gpu_free(GpuState *g)
{
	if (!g)
		return;
	cudaSetDevice(g->device);
	if (g->stream)
		cudaStreamSynchronize(g->stream); /* a look-ahead may still be writing the buffers that go back to the pool */
	pool_free(g->d_blob);
	pool_free(g->d_desc);
	pool_free(g->d_status);
	pool_free(g->d_words);
	pool_free(g->d_cks);
	pool_free(g->d_counter);
	pool_free(g->d_endpos);
	pool_free(g->d_hist);
	pool_free(g->scratch.buf);
	pool_free(g->d_pcm);
	pool_free(g->d_split);
	for (int k = 0; k < 2; k++) {
		pool_free(g->d_chunk[k]);
		pool_free(g->h_chunk[k]);
		pool_free(g->h_recs[k]);
		pool_free(g->h_state[k]);
	}
	for (int k = 0; k < 2; k++)
		if (g->slot_done[k])
			cudaEventDestroy(g->slot_done[k]);
	if (g->stream)
		cudaStreamDestroy(g->stream);
	delete g;
}

int gpu_setup(GpuState *g)
{
	const uint32_t blen = g->hdr.rows << g->hdr.level;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
		acm_set_error("no CUDA device: libacm_b200 has no CPU decode path");
		return ACM_ERR_OTHER;
	}
	CUS(cudaGetDevice(&g->device));
	CUS(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
	g->blob_room = 0;
	CUS(pool_alloc((void **)&g->d_desc, sizeof(DevStream), false, g->device));
	CUS(pool_alloc((void **)&g->d_status, 16, false, g->device));
	CUS(pool_alloc((void **)&g->d_words, 16, false, g->device));
	CUS(pool_alloc((void **)&g->d_cks, 16, false, g->device));
	CUS(pool_alloc((void **)&g->d_counter, 64, false, g->device));
	CUS(pool_alloc((void **)&g->d_endpos, 16, false, g->device));
	CUS(pool_alloc((void **)&g->d_hist, (size_t)2 * g->cols * 4 + 16, false, g->device));
	g->d_tables = const_cast<acm_tables *>(device_tables(g->device));
	if (!g->d_tables) {
		acm_set_error("code tables: device allocation failed");
		return ACM_ERR_OTHER;
	}
	g->scratch.stride = generic_scratch_words(blen, g->cols);
	g->scratch.max_blen = blen;
	g->scratch.max_cols = g->cols;
	CUS(pool_alloc((void **)&g->scratch.buf, g->scratch.stride * 4, false, g->device));
	g->pcm_cap = (size_t)g->chunk_blocks * blen * 4 + 64;
	CUS(pool_alloc((void **)&g->d_pcm, g->pcm_cap, false, g->device));
	CUS(cudaDeviceGetAttribute(&g->sm_count, cudaDevAttrMultiProcessorCount, g->device));
	g->split = split_shape(g->hdr.level, g->hdr.rows) && g->n_attempt_total > 0 &&
		   (uint64_t)g->n_attempt_total * split_bytes_per_block() < ((uint64_t)8 << 30);
	if (g->split) {
		/* one allocation: stream table, state, work items (runs of SP_RUN blocks), per-block arrays */
		static std::atomic<uint32_t> epoch_counter{0x40000000u};
		constexpr uint32_t SP_RUN = 8;
		const uint64_t nb = (uint64_t)g->n_attempt_total + 1;
		const size_t n_items = (size_t)((nb + SP_RUN - 1) / SP_RUN);
		size_t total = 0;
		auto carve = [&total](size_t bytes) {
			const size_t at = total;
			total += (bytes + 255u) & ~(size_t)255u;
			return at;
		};
		const size_t o_gs = carve(sizeof(Gen2Stream)), o_state = carve(64), o_items = carve(n_items * sizeof(Gen2Item));
		const size_t o_rec = carve(nb * sizeof(BlockRec)), o_cks = carve(nb * 8), o_coff = carve(nb * 256);
		const size_t o_wmask = carve(nb * 16), o_inter = carve(nb * 2048), o_wide = carve(nb * 4096);
		CUS(pool_alloc((void **)&g->d_split, total, false, g->device));
		uint8_t *base = g->d_split;
		Gen2Stream gs;
		memset(&gs, 0, sizeof(gs));
		gs.max_blocks = g->n_attempt_total;
		std::vector<Gen2Item> items(n_items);
		for (size_t i = 0; i < n_items; i++) {
			items[i].stream = 0;
			items[i].b0 = (uint32_t)(i * SP_RUN);
			items[i].nb = SP_RUN;
			items[i].warm = 0;
		}
		CUS(cudaMemcpyAsync(base + o_gs, &gs, sizeof(gs), cudaMemcpyHostToDevice, g->stream));
		CUS(cudaMemcpyAsync(base + o_items, items.data(), n_items * sizeof(Gen2Item), cudaMemcpyHostToDevice, g->stream));
		CUS(cudaMemsetAsync(base + o_state, 0, 4, g->stream));        /* nscan */
		CUS(cudaMemsetAsync(base + o_state + 4, 0xFF, 4, g->stream)); /* first_bad */
		CUS(cudaMemsetAsync(base + o_rec, 0, nb * sizeof(BlockRec), g->stream));
		CUS(cudaStreamSynchronize(g->stream)); /* `items` goes out of scope */
		memset(&g->sp, 0, sizeof(g->sp));
		g->sp.gs = reinterpret_cast<const Gen2Stream *>(base + o_gs);
		g->sp.nscan = reinterpret_cast<uint32_t *>(base + o_state);
		g->sp.first_bad = g->sp.nscan + 1;
		g->sp.item_counter = g->sp.nscan + 2;
		g->d_items = reinterpret_cast<Gen2Item *>(base + o_items);
		g->sp.rec = reinterpret_cast<BlockRec *>(base + o_rec);
		g->sp.cks_blk = reinterpret_cast<unsigned long long *>(base + o_cks);
		g->sp.coff16 = reinterpret_cast<uint16_t *>(base + o_coff);
		g->sp.wmask = reinterpret_cast<uint32_t *>(base + o_wmask);
		g->sp.inter = base + o_inter;
		g->sp.wide = reinterpret_cast<uint16_t *>(base + o_wide);
		g->sp.n_blocks = g->n_attempt_total;
		g->sp.epoch = epoch_counter.fetch_add(1);
		for (int k = 0; k < 2; k++) {
			const size_t cb = (size_t)g->chunk_blocks * blen * 2 + 64;
			CUS(pool_alloc((void **)&g->d_chunk[k], cb, false, g->device));
			CUS(pool_alloc((void **)&g->h_chunk[k], cb, true, g->device));
			CUS(pool_alloc((void **)&g->h_recs[k], (size_t)g->chunk_blocks * sizeof(BlockRec), true, g->device));
			CUS(pool_alloc((void **)&g->h_state[k], 16, true, g->device));
			CUS(cudaEventCreateWithFlags(&g->slot_done[k], cudaEventDisableTiming));
		}
	}
	return ACM_OK;
}

/*
 * Pull the image through read_func until `want` bytes of it are cached (or the source ends), the
 * way the reference's load_buf asks: (ptr, 1, 65536) (decode.c:50-52).  After a backward seek the
 * source stands before the end of the cache: what comes again is skipped.
 */
int pull(ACMStream *acm, GpuState *g, size_t want)
{
	std::vector<uint8_t> chunk;
	while (g->file.size() < want && !g->complete && !g->src_eof && !g->read_error) {
		if (chunk.empty())
			chunk.resize(65536);
		const int got = acm->io.read_func ? acm->io.read_func(chunk.data(), 1, (int)chunk.size(), acm->io_arg) : 0;
		if (got < 0) {
			g->read_error = true;
			break;
		}
		if (got == 0) {
			g->src_eof = true;
			if (g->src_pos >= g->file.size())
				g->complete = true;
			break;
		}
		const uint64_t end = g->src_pos + (uint64_t)got;
		if (end > g->file.size()) {
			const size_t skip = g->src_pos < g->file.size() ? (size_t)(g->file.size() - g->src_pos) : 0;
			g->file.insert(g->file.end(), chunk.begin() + (long)skip, chunk.begin() + got);
		}
		g->src_pos = end;
	}
	return ACM_OK;
}

/* bring the device copy of the image up to date (grown geometrically, or sized from
 * get_length_func's answer when there is one) */
int upload(GpuState *g)
{
	const size_t have = g->file.size();
	if (have + 64 > g->blob_cap) {
		size_t cap = g->blob_cap ? g->blob_cap * 2 : ((size_t)1 << 20);
		if (cap < have + 64)
			cap = have + 64;
		if (g->len_hint + 64 > cap && g->len_hint >= have)
			cap = g->len_hint + 64;
		cap = (cap + 255u) & ~(size_t)255u;
		uint8_t *nb = nullptr;
		CUS(pool_alloc((void **)&nb, cap, false, g->device));
		CUS(cudaMemsetAsync(nb, 0, cap, g->stream));
		if (g->uploaded)
			CUS(cudaMemcpyAsync(nb, g->d_blob, g->uploaded, cudaMemcpyDeviceToDevice, g->stream));
		CUS(cudaStreamSynchronize(g->stream));
		pool_free(g->d_blob);
		g->d_blob = nb;
		g->blob_cap = cap;
	}
	if (have > g->uploaded) {
		CUS(cudaMemcpyAsync(g->d_blob + g->uploaded, g->file.data() + g->uploaded, have - g->uploaded,
				    cudaMemcpyHostToDevice, g->stream));
		g->uploaded = have;
	}
	g->blob_room = (have + 15u) & ~(uint64_t)15u;
	return ACM_OK;
}

/*
 * Decode the chunk that starts at index entry `si` (blocks [b0, b0 + nb)) in format key
 * `fmt` into the host cache.  On full success the end state becomes a new index entry.
 */
int decode_chunk_split(ACMStream *acm, GpuState *g, size_t si, uint32_t nb, DevStream d, int be, int sgned, bool lift);

/* what every chunk decode starts with: pull and upload what the chunk can consume at most, and the
 * stream descriptor with the end of the data as it is known by then */
int chunk_prologue(ACMStream *acm, GpuState *g, size_t si, uint32_t *nb_out, DevStream *d_out)
{
	const SavedState &s0 = g->index[si];
	const uint32_t b0 = s0.block;
	uint32_t nb = g->n_attempt_total - b0;
	if (nb > g->chunk_blocks)
		nb = g->chunk_blocks;
	DevStream d = g->base;
	CUS(cudaSetDevice(g->device));
	{
		/* everything this chunk can consume at most, or the end of the source */
		const unsigned long long colbits = 5ull + 16ull * g->hdr.rows;
		const unsigned long long blkbytes = (20ull + (unsigned long long)g->cols * colbits + 7ull) / 8ull;
		const unsigned long long at = g->base.base_off + (s0.P >> 3);
		unsigned long long want = at + (unsigned long long)nb * blkbytes + 64ull;
		if (want > ((unsigned long long)1 << 40))
			want = (unsigned long long)1 << 40;
		int perr = pull(acm, g, (size_t)want);
		if (perr < 0)
			return perr;
		perr = upload(g);
		if (perr < 0)
			return perr;
		const uint32_t hdr = g->hdr.header_len;
		const unsigned long long data_len = g->file.size() > hdr ? g->file.size() - hdr : 0;
		if (data_len * 8ull + 256ull + (1ull << 20) >= 0xFFFFFFFFull) {
			acm_set_error("stream image of 512 MiB or more: not supported by the device decoder");
			return ACM_ERR_OTHER;
		}
		/* file_end: the real end once the source has run dry (the reference's one zero byte follows
		 * it, decode.c:57-61); until then the end of what has been pulled, which no block of this
		 * chunk can reach */
		d.file_end = g->base.bit0 + (uint32_t)data_len * 8u;
		g->base.file_end = d.file_end;
	}
	*nb_out = nb;
	*d_out = d;
	return ACM_OK;
}

int decode_chunk(ACMStream *acm, GpuState *g, size_t si, int be, int wordlen, int sgned, bool lift = true)
{
	const SavedState &s0 = g->index[si];
	const uint32_t blen = acm->block_len, b0 = s0.block;
	uint32_t nb = 0;
	DevStream d;
	KernelArgs a;
	uint32_t endpos[2] = { 0, 0 }, words = 0;
	int32_t status = 0;
	{
		const int perr = chunk_prologue(acm, g, si, &nb, &d);
		if (perr < 0)
			return perr;
	}
	d.bit0 = s0.P;
	d.out_off = 0;
	d.index = 0;
	d.pad_words = 0;
	d.n_attempt = nb;
	d.resume = s0.hist.empty() ? 0u : 1u;
	{
		/* words the read loop may still deliver from block b0 on, capped to this chunk */
		uint64_t done = (uint64_t)b0 * blen;
		uint64_t left = g->words_limit > done ? g->words_limit - done : 0;
		uint64_t cap = (uint64_t)nb * blen;
		d.words_limit = (uint32_t)(left < cap ? left : cap);
	}
	if (g->split && wordlen == 2)
		return decode_chunk_split(acm, g, si, nb, d, be, sgned, lift);
	if (!lift)
		return ACM_OK; /* only the split path can skip ahead */
	CUS(cudaMemcpyAsync(g->d_desc, &d, sizeof(d), cudaMemcpyHostToDevice, g->stream));
	if (d.resume)
		CUS(cudaMemcpyAsync(g->d_hist, s0.hist.data(), s0.hist.size() * 4, cudaMemcpyHostToDevice, g->stream));
	CUS(cudaMemsetAsync(g->d_counter, 0, 64, g->stream));
	memset(&a, 0, sizeof(a));
	a.blob = g->d_blob;
	a.blob_room = g->blob_room;
	a.out = g->d_pcm;
	a.streams = g->d_desc;
	a.count = 1;
	a.status = g->d_status;
	a.words = g->d_words;
	a.cks = g->d_cks;
	a.tables = g->d_tables;
	a.counter = g->d_counter;
	a.errflag = g->d_counter + 2;
	a.fmt.wordlen = wordlen;
	a.fmt.be = be ? 1 : 0;
	a.fmt.bias = sgned ? 0u : (1u << (8 * wordlen - 1));
	a.fmt.checksums = 0;
	a.resume_hist = g->d_hist;
	a.resume_stride = 2 * g->cols;
	a.end_pos = g->d_endpos;
	CUS(launch_generic(a, g->scratch, 1, g->stream));
	CUS(cudaMemcpyAsync(&status, g->d_status, 4, cudaMemcpyDeviceToHost, g->stream));
	CUS(cudaMemcpyAsync(&words, g->d_words, 4, cudaMemcpyDeviceToHost, g->stream));
	CUS(cudaMemcpyAsync(endpos, g->d_endpos, 8, cudaMemcpyDeviceToHost, g->stream));
	CUS(cudaStreamSynchronize(g->stream));
	g->pcm.resize((size_t)words * wordlen + 16);
	if (words)
		CUS(cudaMemcpy(g->pcm.data(), g->d_pcm, (size_t)words * wordlen, cudaMemcpyDeviceToHost));
	g->pcm_ptr = g->pcm.data();
	g->c_valid = true;
	g->c_fmt = fmt_key(be, wordlen, sgned);
	g->c_block0 = b0;
	g->c_attempted = nb;
	g->c_nok = endpos[1];
	g->c_words = words;
	g->c_endP = endpos[0];
	g->c_status = status;
	if (g->c_endP > g->consumed_P)
		g->consumed_P = g->c_endP;
	if (g->c_nok == nb && b0 + nb < g->n_attempt_total && si + 1 == g->index.size()) {
		SavedState n;
		n.block = b0 + nb;
		n.P = g->c_endP;
		n.hist.resize((size_t)2 * g->cols);
		CUS(cudaMemcpy(n.hist.data(), g->d_hist, n.hist.size() * 4, cudaMemcpyDeviceToHost));
		g->index.push_back(std::move(n));
	}
	return ACM_OK;
}

/*
 * The same for a stream on the split path: one lane walks blocks [b0, b0 + nb) (acm_walk1_kernel),
 * the blocks are unpacked in parallel and -- unless lift is false: a seek skipping ahead, which only
 * needs to know that the blocks decode -- transformed in parallel, PCM into a pinned chunk slot.
 * split_launch queues all of it on the stream; split_finish, after the stream has been waited for,
 * turns the records into what the reference's read loop would see.
 */
int split_launch(ACMStream *acm, GpuState *g, uint32_t b0, uint32_t nb, uint32_t P0, DevStream d, int be, int sgned,
		 bool lift, int slot)
{
	const uint32_t blen = acm->block_len;
	KernelArgs a;
	SplitArgs sp = g->sp;
	constexpr uint32_t SP_RUN = 8;

	d.bit0 = g->base.bit0;
	d.n_attempt = g->n_attempt_total;
	d.words_limit = g->words_limit;
	d.resume = 0;
	CUS(cudaMemcpyAsync(g->d_desc, &d, sizeof(d), cudaMemcpyHostToDevice, g->stream));
	CUS(cudaMemsetAsync(sp.item_counter, 0, 4, g->stream));
	memset(&a, 0, sizeof(a));
	a.blob = g->d_blob;
	a.blob_room = g->blob_room;
	/* the lift kernel writes block b at word b * block_len of the stream's PCM: block b0 = the start of
	 * the chunk buffer */
	a.out = g->d_chunk[slot] - (size_t)b0 * blen * 2u;
	a.streams = g->d_desc;
	a.count = 1;
	a.status = g->d_status;
	a.words = g->d_words;
	a.cks = g->d_cks;
	a.tables = g->d_tables;
	a.counter = g->d_counter;
	a.errflag = g->d_counter + 2;
	a.fmt.wordlen = 2;
	a.fmt.be = be ? 1 : 0;
	a.fmt.bias = sgned ? 0u : 0x8000u;
	a.fmt.checksums = 0;
	sp.items = g->d_items + b0 / SP_RUN; /* chunks start at multiples of chunk_blocks, a multiple of SP_RUN */
	sp.n_items = (nb + SP_RUN - 1) / SP_RUN;
	CUS(launch_split_range(a, sp, b0, nb, P0, lift ? 1 : 0, g->sm_count, g->stream));
	CUS(cudaMemcpyAsync(g->h_recs[slot], sp.rec + b0, (size_t)nb * sizeof(BlockRec), cudaMemcpyDeviceToHost, g->stream));
	CUS(cudaMemcpyAsync(g->h_state[slot], sp.nscan, 8, cudaMemcpyDeviceToHost, g->stream));
	if (lift) {
		const uint64_t done = (uint64_t)b0 * blen, left = g->words_limit > done ? g->words_limit - done : 0;
		const uint64_t cap = (uint64_t)nb * blen, words = left < cap ? left : cap;
		if (words)
			CUS(cudaMemcpyAsync(g->h_chunk[slot], g->d_chunk[slot], (size_t)words * 2, cudaMemcpyDeviceToHost, g->stream));
	}
	CUS(cudaEventRecord(g->slot_done[slot], g->stream));
	return ACM_OK;
}

void split_finish(ACMStream *acm, GpuState *g, uint32_t b0, uint32_t nb, uint32_t P0, int key, bool lift, int slot)
{
	const uint32_t blen = acm->block_len;
	const BlockRec *recs = g->h_recs[slot];
	const uint32_t nscan = g->h_state[slot][0], first_bad = g->h_state[slot][1];
	/* blocks decode up to the first one whose walk fails, or that holds an out-of-range radix code */
	const uint32_t walked = nscan > b0 ? nscan - b0 : 0u;
	uint32_t nok = 0;
	int status = 0;
	while (nok < nb && nok < walked && recs[nok].status == SCAN_OK)
		nok++;
	if (nok < nb && nok < walked)
		status = recs[nok].status == SCAN_EOF ? 0 : recs[nok].status;
	if (first_bad < b0 + nok || (first_bad == b0 + nok && nok < nb)) {
		nok = first_bad > b0 ? first_bad - b0 : 0u;
		status = ACM_ERR_CORRUPT;
	}
	if (lift) {
		const uint64_t done = (uint64_t)b0 * blen, left = g->words_limit > done ? g->words_limit - done : 0;
		const uint64_t cap = (uint64_t)nok * blen;
		g->pcm_ptr = g->h_chunk[slot];
		g->cur_slot = slot;
		g->c_valid = true;
		g->c_fmt = key;
		g->c_block0 = b0;
		g->c_attempted = nb;
		g->c_nok = nok;
		g->c_words = (uint32_t)(left < cap ? left : cap);
		g->c_endP = nok ? recs[nok - 1].end : P0;
		g->c_status = status;
		if (g->c_endP > g->consumed_P)
			g->consumed_P = g->c_endP;
	}
	if (nok == nb && b0 + nb < g->n_attempt_total && g->index.back().block == b0) {
		SavedState n;
		n.block = b0 + nb;
		n.P = recs[nb - 1].end;
		g->index.push_back(std::move(n));
	}
}

int chunk_prologue(ACMStream *acm, GpuState *g, size_t si, uint32_t *nb_out, DevStream *d_out);

/* queue the chunk after the cached one, if it is known to exist and nothing is under way */
void split_look_ahead(ACMStream *acm, GpuState *g, int be, int sgned)
{
	if (!g->split || !g->c_valid || g->c_nok != g->c_attempted || g->ahead.active)
		return;
	const uint32_t nextb = g->c_block0 + g->c_attempted;
	if (nextb >= g->n_attempt_total)
		return;
	size_t si = g->index.size();
	while (si-- > 0)
		if (g->index[si].block <= nextb)
			break;
	if (si == (size_t)-1 || g->index[si].block != nextb)
		return;
	uint32_t nb = 0;
	DevStream d;
	if (chunk_prologue(acm, g, si, &nb, &d) < 0)
		return;
	const int slot = 1 - g->cur_slot;
	if (split_launch(acm, g, nextb, nb, g->index[si].P, d, be, sgned, true, slot) < 0)
		return;
	g->ahead.active = true;
	g->ahead.b0 = nextb;
	g->ahead.nb = nb;
	g->ahead.P0 = g->index[si].P;
	g->ahead.key = fmt_key(be, 2, sgned);
	g->ahead.slot = slot;
}

int decode_chunk_split(ACMStream *acm, GpuState *g, size_t si, uint32_t nb, DevStream d, int be, int sgned, bool lift)
{
	const uint32_t b0 = g->index[si].block, P0 = g->index[si].P;
	/* a chunk slot that a look-ahead may still be filling is not reused before the stream is idle */
	const int slot = g->ahead.active ? 1 - g->ahead.slot : g->cur_slot;
	int err = split_launch(acm, g, b0, nb, P0, d, be, sgned, lift, slot);
	if (err < 0)
		return err;
	CUS(cudaStreamSynchronize(g->stream));
	g->ahead.active = false; /* whatever was under way has landed too; it is not what was asked for */
	split_finish(acm, g, b0, nb, P0, fmt_key(be, 2, sgned), lift, slot);
	if (lift)
		split_look_ahead(acm, g, be, sgned);
	return ACM_OK;
}

/*
 * A seek skipping ahead on the split path: walk (and unpack, which finds out-of-range radix codes)
 * whole chunks until the index reaches the chunk that holds block tb; nothing is transformed or
 * copied.  Two chunks are in flight: chunk k + 1 is queued -- its walk takes its start position from
 * chunk k's last record on the device -- before the host waits for chunk k, so the walker, the one
 * serial resource, never waits for the host.
 */
int split_skip_ahead(ACMStream *acm, GpuState *g, uint32_t tb)
{
	struct Pend {
		bool on = false;
		uint32_t b0 = 0, nb = 0;
		int slot = 0;
	} pend;
	const unsigned long long colbits = 5ull + 16ull * g->hdr.rows;
	const unsigned long long blkbytes = (20ull + (unsigned long long)g->cols * colbits + 7ull) / 8ull;
	CUS(cudaSetDevice(g->device));
	if (g->ahead.active) {
		CUS(cudaStreamSynchronize(g->stream)); /* a look-ahead of the read path: dropped */
		g->ahead.active = false;
	}
	uint32_t nextb = g->index.back().block;
	/* where the next chunk can start at most (bytes from the start of the image) */
	unsigned long long at_upper = g->base.base_off + (g->index.back().P >> 3);
	for (;;) {
		const bool more = nextb + g->chunk_blocks <= tb && nextb + g->chunk_blocks < g->n_attempt_total;
		int slot = 0;
		if (more) {
			DevStream d = g->base;
			at_upper += (unsigned long long)g->chunk_blocks * blkbytes;
			unsigned long long want = at_upper + 64ull;
			if (want > ((unsigned long long)1 << 40))
				want = (unsigned long long)1 << 40;
			int perr = pull(acm, g, (size_t)want);
			if (perr < 0)
				return perr;
			perr = upload(g);
			if (perr < 0)
				return perr;
			const uint32_t hdr = g->hdr.header_len;
			const unsigned long long data_len = g->file.size() > hdr ? g->file.size() - hdr : 0;
			if (data_len * 8ull + 256ull + (1ull << 20) >= 0xFFFFFFFFull) {
				acm_set_error("stream image of 512 MiB or more: not supported by the device decoder");
				return ACM_ERR_OTHER;
			}
			d.file_end = g->base.bit0 + (uint32_t)data_len * 8u;
			g->base.file_end = d.file_end;
			slot = pend.on ? 1 - pend.slot : 1 - g->cur_slot;
			/* P0 = 0: continue from the previous chunk's last record (on the device) */
			const int err = split_launch(acm, g, nextb, g->chunk_blocks, pend.on ? 0u : g->index.back().P, d, 0, 1, false, slot);
			if (err < 0)
				return err;
		}
		if (pend.on) {
			CUS(cudaEventSynchronize(g->slot_done[pend.slot]));
			const size_t before = g->index.size();
			split_finish(acm, g, pend.b0, pend.nb, 0u, fmt_key(0, 2, 1), false, pend.slot);
			if (g->index.size() == before) {
				/* a block on the way does not decode: the read loop reports it */
				if (more)
					CUS(cudaStreamSynchronize(g->stream));
				return ACM_OK;
			}
			at_upper = g->base.base_off + (g->index.back().P >> 3) + (more ? (unsigned long long)g->chunk_blocks * blkbytes : 0ull);
		}
		if (!more)
			break;
		pend.on = true;
		pend.b0 = nextb;
		pend.nb = g->chunk_blocks;
		pend.slot = slot;
		nextb += g->chunk_blocks;
	}
	return ACM_OK;
}

/*
 * Make block b available in the cache (in the given format).  Returns 1 if the block
 * decoded, 0 for a clean end of stream at that block (decode.c:842-843), <0 for the
 * error the reference's decode_block would return.
 */
int ensure_block(ACMStream *acm, GpuState *g, uint32_t b, int be, int wordlen, int sgned)
{
	const int key = fmt_key(be, wordlen, sgned);
	if (g->split && wordlen != 2) {
		/* 24 / 32-bit words are the generic kernel's: its index entries carry the transform history,
		 * the split path's do not -- start the index over */
		g->split = false;
		g->index.resize(1);
		g->c_valid = false;
	}
	for (;;) {
		if (g->split && g->ahead.active && g->ahead.key == key && b >= g->ahead.b0 && b < g->ahead.b0 + g->ahead.nb &&
		    !(g->c_valid && g->c_fmt == key && b >= g->c_block0 && b < g->c_block0 + g->c_attempted)) {
			/* the chunk that was queued behind the previous one */
			if (cudaStreamSynchronize(g->stream) != cudaSuccess)
				return ACM_ERR_OTHER;
			g->ahead.active = false;
			split_finish(acm, g, g->ahead.b0, g->ahead.nb, g->ahead.P0, key, true, g->ahead.slot);
			split_look_ahead(acm, g, be, sgned);
			continue;
		}
		if (g->c_valid && g->c_fmt == key && b >= g->c_block0 && b < g->c_block0 + g->c_attempted) {
			if (b < g->c_block0 + g->c_nok)
				return 1;
			/* b is the failing attempt (later blocks are never reached); if the data source
			 * itself failed, running out of bits is a read error (decode.c:54-55) */
			if (g->read_error && (g->c_status == 0 || g->c_status == ACM_ERR_UNEXPECTED_EOF))
				return ACM_ERR_READ_ERR;
			return g->c_status;
		}
		if (b >= g->n_attempt_total)
			return 0;
		/* index entry with the largest block <= b */
		size_t si = g->index.size() - 1;
		while (si > 0 && g->index[si].block > b)
			si--;
		/* entries are chunk starts; if b lies past the chunk of the last entry, walk forward */
		int err = decode_chunk(acm, g, si, be, wordlen, sgned);
		if (err < 0)
			return err;
		if (b >= g->c_block0 + g->c_attempted) {
			if (g->c_nok < g->c_attempted) {
				/* an earlier block fails: b is unreachable, report that failure */
				if (g->read_error && (g->c_status == 0 || g->c_status == ACM_ERR_UNEXPECTED_EOF))
					return ACM_ERR_READ_ERR;
				return g->c_status;
			}
			continue; /* a new index entry was appended: decode the next chunk */
		}
	}
}

} // namespace

/* ------------------------------------------------------------------ open / close */

extern "C" int acm_open_decoder(ACMStream **res, void *io_arg, acm_io_callbacks io, int force_chans)
{
	ACMStream *acm = (ACMStream *)calloc(1, sizeof(*acm));
	GpuState *g = nullptr;
	int err = ACM_ERR_OTHER;
	if (!acm)
		return ACM_ERR_OTHER;
	acm->io_arg = io_arg;
	acm->io = io;
	acm->data_len = io.get_length_func ? (unsigned)io.get_length_func(io_arg) : 0; /* decode.c:771-775 */
	g = new (std::nothrow) GpuState();
	if (!g)
		goto fail;
	/* the first window: the header lies in it (the reference's first load_buf, decode.c:41-67;
	 * a source that hands out a few bytes per call is asked until 42 bytes or its end are there) */
	g->len_hint = acm->data_len;
	pull(acm, g, 64);
	err = ACM_ERR_NOT_ACM; /* decode.c:783-785: every header problem */
	if (acm_parse_header(g->file.data(), g->file.size(), force_chans, &g->hdr) < 0)
		goto fail;
	acm->info.channels = g->hdr.channels;
	acm->info.rate = g->hdr.rate;
	acm->info.acm_id = ACM_ID;
	acm->info.acm_version = 1;
	acm->info.acm_channels = g->hdr.acm_channels;
	acm->info.acm_level = g->hdr.level;
	acm->info.acm_cols = 1u << g->hdr.level;
	acm->info.acm_rows = g->hdr.rows;
	acm->total_values = g->hdr.total_values;
	acm->wavc_file = g->hdr.wavc ? 1 : 0;
	acm->block_len = g->hdr.rows << g->hdr.level;           /* decode.c:802-804 */
	acm->wrapbuf_len = 2 * acm->info.acm_cols - 2;
	acm->buf_start_ofs = g->hdr.header_len;
	g->cols = acm->info.acm_cols;
	{
		acm_gpu_stream s;
		memset(&s, 0, sizeof(s));
		s.in_off = 0;
		s.in_len = (uint32_t)(g->len_hint > g->file.size() ? g->len_hint : g->file.size());
		s.total_values = g->hdr.total_values;
		s.channels = g->hdr.channels;
		s.level = g->hdr.level;
		s.rows = g->hdr.rows;
		s.wavc = g->hdr.wavc;
		err = acm_make_devstream(&s, 0, 0, &g->base);
		if (err < 0)
			goto fail;
		err = ACM_ERR_OTHER;
	}
	g->n_attempt_total = g->base.n_attempt;
	g->words_limit = g->base.words_limit;
	{
		/* look-ahead chunk: about 256 K words, 1..256 blocks */
		uint32_t cb = 262144u / (acm->block_len ? acm->block_len : 1u);
		g->chunk_blocks = cb < 1 ? 1 : (cb > 256 ? 256 : cb);
	}
	if (gpu_setup(g) < 0)
		goto fail;
	{
		SavedState s0;
		s0.block = 0;
		s0.P = g->base.bit0;
		g->index.push_back(std::move(s0));
	}
	g->consumed_P = g->base.bit0;
	acm->gpu = g;
	*res = acm;
	return ACM_OK;
fail:
	/* decode.c:817-823: the callbacks are dropped, close_func is NOT called */
	gpu_free(g);
	free(acm);
	return err;
}

extern "C" void acm_close(ACMStream *acm)
{
	if (!acm)
		return; /* decode.c:880-881 */
	if (acm->io.close_func)
		acm->io.close_func(acm->io_arg);
	gpu_free((GpuState *)acm->gpu);
	free(acm);
}

/* ------------------------------------------------------------------ read */

extern "C" int acm_read(ACMStream *acm, void *dst, unsigned numbytes, int bigendianp, int wordlen, int sgned)
{
	GpuState *g = (GpuState *)acm->gpu;
	int numwords, avail, gotbytes;

	if (wordlen < 2 || wordlen > 4)
		return ACM_ERR_BADFMT; /* decode.c:832-835 (2 only there) */
	numwords = (int)(numbytes / (unsigned)wordlen);
	if (acm->stream_pos >= acm->total_values)
		return 0; /* decode.c:837-838 */

	if (!acm->block_ready) {
		if (g->drained)
			return 0;
		int err = ensure_block(acm, g, g->next_block, bigendianp, wordlen, sgned);
		if (err == 0)
			return 0; /* EXPECTED_EOF, decode.c:842-843 */
		if (err == ACM_ERR_UNEXPECTED_EOF)
			g->drained = true; /* fewer bits are left than a block header needs */
		if (err < 0)
			return err;
		g->cur_block = g->next_block;
		acm->block_ready = 1;
		acm->block_pos = 0;
	} else if (dst) {
		/* same block, but the cache may hold another chunk or another format by now */
		int err = ensure_block(acm, g, g->cur_block, bigendianp, wordlen, sgned);
		if (err <= 0)
			return err < 0 ? err : ACM_ERR_OTHER;
	}

	/* decode.c:849-857 */
	avail = (int)(acm->block_len - acm->block_pos);
	if (avail < numwords)
		numwords = avail;
	if (acm->stream_pos + (unsigned)numwords > acm->total_values)
		numwords = (int)(acm->total_values - acm->stream_pos);
	if (acm->info.channels > 1)
		numwords -= numwords % (int)acm->info.channels;

	gotbytes = numwords * wordlen;
	if (dst && numwords > 0) {
		size_t off = ((size_t)(g->cur_block - g->c_block0) * acm->block_len + acm->block_pos) * (size_t)wordlen;
		if (off + (size_t)gotbytes > (size_t)g->c_words * wordlen)
			return ACM_ERR_OTHER; /* cannot happen: the block decoded and holds these words */
		memcpy(dst, g->pcm_ptr + off, (size_t)gotbytes);
	}
	/* decode.c:868-873 */
	acm->stream_pos += (unsigned)numwords;
	acm->block_pos += (unsigned)numwords;
	if (acm->block_pos == acm->block_len) {
		acm->block_ready = 0;
		g->next_block = g->cur_block + 1;
	}
	return gotbytes;
}

/* util.c:258-277 */
extern "C" int acm_read_loop(ACMStream *acm, void *dst, unsigned bytes, int bigendianp, int wordlen, int sgned)
{
	unsigned char *dstp = (unsigned char *)dst;
	int res, got = 0;
	while (bytes > 0) {
		res = acm_read(acm, dstp, bytes, bigendianp, wordlen, sgned);
		if (res > 0) {
			if (dstp)
				dstp += res;
			got += res;
			bytes -= (unsigned)res;
		} else {
			if (res < 0 && got == 0)
				return res;
			break;
		}
	}
	return got;
}

/* ------------------------------------------------------------------ seek */

/* util.c:214-253, with the chunk index standing in for "decode again from the start" */
extern "C" int acm_seek_pcm(ACMStream *acm, unsigned pcm_pos)
{
	GpuState *g = (GpuState *)acm->gpu;
	unsigned word_pos = pcm_pos * acm->info.channels;

	if (word_pos < acm->stream_pos) {
		if (acm->io.seek_func == NULL)
			return ACM_ERR_NOT_SEEKABLE;
		if (acm->io.seek_func(acm->io_arg, (int)g->hdr.header_len, SEEK_SET) < 0)
			return ACM_ERR_NOT_SEEKABLE;
		g->src_pos = g->hdr.header_len; /* where the source stands now; the cache keeps what it has */
		g->src_eof = false;
		acm->stream_pos = 0;
		acm->block_pos = 0;
		acm->block_ready = 0;
		acm->buf_start_ofs = 14; /* util.c:239 (Q12: 14 even for WAVC) */
		g->next_block = 0;
		g->cur_block = 0;
		g->drained = false; /* util.c:230-234: the reader starts over */
	}
	if (g->split && word_pos > acm->stream_pos && acm->block_len % acm->info.channels == 0) {
		/* split path: the walk alone (plus the unpack, which finds out-of-range codes) extends the
		 * index up to the chunk that holds the target; nothing is transformed or copied on the way */
		/* whatever goes wrong on the way is reported by the read loop below, as in the reference
		 * (util.c:243-251: the loop breaks, the position reached is returned) */
		(void)split_skip_ahead(acm, g, word_pos / acm->block_len);
	}
	while (acm->stream_pos < word_pos) {
		int step = 2048, res;
		/* at a block boundary, jump over whole blocks that are known to decode: to the start
		 * of the last visited chunk at or before the target (only when blocks hold whole
		 * frames; otherwise the read loop stalls inside block 0 anyway, Q3) */
		if (!acm->block_ready && acm->block_len % acm->info.channels == 0) {
			const uint32_t tb = word_pos / acm->block_len;
			for (size_t i = g->index.size(); i-- > 0;) {
				const SavedState &s = g->index[i];
				if (s.block <= g->next_block)
					break;
				if (s.block <= tb) {
					g->next_block = s.block;
					acm->stream_pos = s.block * acm->block_len;
					break;
				}
			}
			if (acm->stream_pos >= word_pos)
				break;
		}
		if (acm->stream_pos + (unsigned)step > word_pos)
			step = (int)(word_pos - acm->stream_pos);
		res = acm_read(acm, NULL, (unsigned)step * 2, 0, 2, 1);
		if (res < 1)
			break;
	}
	return (int)(acm->stream_pos / acm->info.channels);
}

static unsigned pcm2time(ACMStream *acm, unsigned long long pcm) { return (unsigned)(pcm * 1000 / acm->info.rate); }
static unsigned time2pcm(ACMStream *acm, unsigned long long ms) { return (unsigned)(ms * acm->info.rate / 1000); }

/* util.c:206-212 */
extern "C" int acm_seek_time(ACMStream *acm, unsigned time_ms)
{
	int res = acm_seek_pcm(acm, time2pcm(acm, time_ms));
	if (res <= 0)
		return res;
	return (int)pcm2time(acm, (unsigned)res);
}

/* ------------------------------------------------------------------ util.c getters */

extern "C" const ACMInfo *acm_info(ACMStream *acm) { return &acm->info; }
extern "C" unsigned acm_rate(ACMStream *acm) { return acm->info.rate; }
extern "C" unsigned acm_channels(ACMStream *acm) { return acm->info.channels; }
extern "C" int acm_seekable(ACMStream *acm) { return acm->data_len > 0; } /* util.c:152-155 */
extern "C" unsigned acm_pcm_tell(ACMStream *acm) { return acm->stream_pos / acm->info.channels; }
extern "C" unsigned acm_pcm_total(ACMStream *acm) { return acm->total_values / acm->info.channels; }
extern "C" unsigned acm_time_tell(ACMStream *acm) { return pcm2time(acm, acm_pcm_tell(acm)); }
extern "C" unsigned acm_time_total(ACMStream *acm) { return pcm2time(acm, acm_pcm_total(acm)); }
extern "C" unsigned acm_raw_total(ACMStream *acm) { return acm->data_len; }

extern "C" unsigned acm_raw_tell(ACMStream *acm)
{
	/* the reference reports bytes pulled into its bit accumulator, 4 at a time (decode.c:117-121);
	 * here: the same rounding applied to the end of the decoded look-ahead */
	GpuState *g = (GpuState *)acm->gpu;
	unsigned long long bits = (unsigned long long)g->hdr.header_len * 8 + (g->consumed_P - g->base.bit0);
	unsigned long long bytes = 4 * ((bits + 31) / 32);
	unsigned long long cap = (unsigned long long)g->file.size() + 1;
	return (unsigned)(bytes < cap ? bytes : cap);
}

/* util.c:157-170 */
extern "C" unsigned acm_bitrate(ACMStream *acm)
{
	unsigned long long bits, time, bitrate = 0;
	if (acm_raw_total(acm) == 0)
		return 13000;
	time = acm_time_total(acm);
	if (time > 0) {
		bits = 8ull * acm_raw_total(acm);
		bitrate = 1000 * bits / time;
	}
	return (unsigned)bitrate;
}

/* util.c:34-52 */
extern "C" const char *acm_strerror(int err)
{
	static const char *const msgs[] = { "No error", "ACM error", "Cannot open file", "Not an ACM file",
					    "Read error", "Bad format", "Corrupt file", "Unexcpected EOF",
					    "Stream not seekable" };
	const int n = (int)(sizeof(msgs) / sizeof(msgs[0]));
	if (-err < 0 || -err >= n)
		return "Unknown error";
	return msgs[-err];
}

/* ------------------------------------------------------------------ stdio front end (util.c:58-115) */

static int file_read(void *ptr, int size, int n, void *arg) { return (int)fread(ptr, (size_t)size, (size_t)n, (FILE *)arg); }
static int file_close(void *arg) { return fclose((FILE *)arg); }
static int file_seek(void *arg, int offset, int whence) { return fseek((FILE *)arg, offset, whence); }
static int file_length(void *arg)
{
	FILE *f = (FILE *)arg;
	long pos = ftell(f), len = -1;
	if (pos < 0)
		return -1;
	if (fseek(f, 0, SEEK_END) >= 0) {
		len = ftell(f);
		fseek(f, pos, SEEK_SET);
	}
	return (int)len;
}

extern "C" int acm_open_file(ACMStream **res, const char *filename, int force_chans)
{
	acm_io_callbacks io;
	ACMStream *acm = NULL;
	FILE *f = fopen(filename, "rb");
	int err;
	if (!f)
		return ACM_ERR_OPEN;
	memset(&io, 0, sizeof(io));
	io.read_func = file_read;
	io.seek_func = file_seek;
	io.close_func = file_close;
	io.get_length_func = file_length;
	if ((err = acm_open_decoder(&acm, f, io, force_chans)) < 0) {
		fclose(f); /* a failed open leaves the data source with the caller (util.c:109-111) */
		return err;
	}
	*res = acm;
	return ACM_OK;
}
