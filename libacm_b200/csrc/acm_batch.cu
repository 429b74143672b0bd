/*
 * acm_batch.cu -- host side of the batched decoder: header parsing, descriptor
 * tables, kernel launches, and the C ABI declared in include/acm_gpu.h.
 *
 * Host work per stream is limited to what the reference does once per stream in
 * acm_open_decoder (decode.c:758-824): read_header / read_wavc_header
 * (decode.c:679-752) and the force_chans rule (decode.c:795-799).  Everything the
 * reference does per block runs on the device.
 */
#include <algorithm>
#include <chrono>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "acm_gpu.h"
#include "acm_host.h"
#include "acm_kernels.cuh"
#include "libacm.h"

using namespace acm;

/* ------------------------------------------------------------------ errors */

static thread_local char g_err[512];

void acm_set_error(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

extern "C" const char *acm_gpu_last_error(void) { return g_err; }
extern "C" int acm_gpu_abi_version(void) { return ACM_GPU_ABI_VERSION; }

#define CU(call)                                                                      \
	do {                                                                          \
		cudaError_t e_ = (call);                                              \
		if (e_ != cudaSuccess) {                                              \
			acm_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call,     \
				      cudaGetErrorString(e_));                        \
			goto fail;                                                    \
		}                                                                     \
	} while (0)

/* ------------------------------------------------------------------ opts */

extern "C" void acm_gpu_opts_init(acm_gpu_opts *o)
{
	memset(o, 0, sizeof(*o));
	o->device = -1;
	o->wordlen = 2;
	o->sgned = 1;
	o->pad_tail = 1;
}

static int use_device(const acm_gpu_opts *o)
{
	if (o && o->device >= 0) {
		cudaError_t e = cudaSetDevice(o->device);
		if (e != cudaSuccess) {
			acm_set_error("cudaSetDevice(%d): %s", o->device, cudaGetErrorString(e));
			return ACM_ERR_OTHER;
		}
	}
	return ACM_OK;
}

/* ------------------------------------------------------------------ probe */

extern "C" int64_t acm_gpu_probe(const void *blob, uint64_t blob_len, int blob_on_device,
				 acm_gpu_stream *s, uint64_t n, const acm_gpu_opts *opts)
{
	acm_gpu_opts defaults;
	std::vector<uint8_t> hdrs;
	uint64_t *d_off = nullptr;
	uint32_t *d_len = nullptr;
	uint8_t *d_hdr = nullptr;
	int64_t ok = 0;
	g_err[0] = 0;
	if (!opts) {
		acm_gpu_opts_init(&defaults);
		opts = &defaults;
	}
	if (blob_on_device) {
		if (use_device(opts) < 0)
			return ACM_ERR_OTHER;
		std::vector<uint64_t> off(n);
		std::vector<uint32_t> len(n);
		for (uint64_t i = 0; i < n; i++) {
			off[i] = s[i].in_off;
			len[i] = s[i].in_len;
		}
		hdrs.resize(n * 48);
		CU(cudaMalloc(&d_off, n * 8 + 8));
		CU(cudaMalloc(&d_len, n * 4 + 8));
		CU(cudaMalloc(&d_hdr, n * 48 + 8));
		CU(cudaMemcpy(d_off, off.data(), n * 8, cudaMemcpyHostToDevice));
		CU(cudaMemcpy(d_len, len.data(), n * 4, cudaMemcpyHostToDevice));
		CU(launch_gather_headers((const uint8_t *)blob, blob_len, d_off, d_len, d_hdr, n, 0));
		CU(cudaMemcpy(hdrs.data(), d_hdr, n * 48, cudaMemcpyDeviceToHost));
		cudaFree(d_off);
		cudaFree(d_len);
		cudaFree(d_hdr);
		d_off = nullptr; d_len = nullptr; d_hdr = nullptr;
	}
	for (uint64_t i = 0; i < n; i++) {
		acm_header h;
		const uint8_t *p;
		uint64_t len = s[i].in_len;
		if (blob_on_device) {
			p = hdrs.data() + i * 48;
		} else {
			if (s[i].in_off > blob_len)
				len = 0;
			else if (s[i].in_off + len > blob_len)
				len = blob_len - s[i].in_off;
			p = (const uint8_t *)blob + (len ? s[i].in_off : 0);
		}
		int err = acm_parse_header(p, len, opts->force_chans, &h);
		s[i].total_values = err < 0 ? 0 : h.total_values;
		s[i].channels = h.channels;
		s[i].acm_channels = h.acm_channels;
		s[i].rate = h.rate;
		s[i].level = h.level;
		s[i].rows = h.rows;
		s[i].wavc = h.wavc;
		s[i].status = err;
		s[i].words = 0;
		s[i].checksum = 0;
		if (err == ACM_OK)
			ok++;
	}
	return ok;
fail:
	cudaFree(d_off);
	cudaFree(d_len);
	cudaFree(d_hdr);
	return ACM_ERR_OTHER;
}

extern "C" uint64_t acm_gpu_layout(acm_gpu_stream *s, uint64_t n, int wordlen)
{
	uint64_t at = 0;
	for (uint64_t i = 0; i < n; i++) {
		s[i].out_off = at;
		at += ((uint64_t)s[i].total_values * (uint64_t)wordlen + 15u) & ~(uint64_t)15u;
	}
	return at;
}

/* ------------------------------------------------------------------ plan */

/*
 * A plan may be cut into SEGMENTS: contiguous ranges of the caller's stream array whose
 * images and PCM occupy contiguous byte ranges of blob / out.  Every segment has its own
 * slices of the descriptor table and its own work-queue cursors, so the host path of
 * acm_gpu_decode_batch can copy segment g+1 in while segment g decodes and segment g-1
 * copies out.  The resident path (acm_gpu_plan_create) uses one segment.
 */
struct Segment {
	uint64_t s_first, s_count;        /* caller's streams [s_first, s_first + s_count) */
	uint64_t fast_first, n_fast;      /* slices of d_streams */
	uint64_t gen_first, n_gen;
	uint64_t blob_lo, blob_hi;        /* byte ranges touched (16-byte granular) */
	uint64_t out_lo, out_hi;
	int fast_ctas, gen_ctas;          /* grid sizes; segments may run concurrently, so each has */
	size_t hist_off, scratch_off;     /* its own history / scratch region (offsets in words) */
	size_t ring_off;                  /* block-record rings of the fast kernel (offset in bytes) */
	size_t ctl_off;                   /* slot control words (offset in bytes) */
	uint32_t n_scan, n_slots;         /* fast kernel geometry: scan CTAs, slots in use */
	uint32_t scan_warps;              /* ... scan warps in use per scan CTA */
	bool sparse;                      /* the kernels do not write every byte of [out_lo, out_hi) */
	int walk_bound;                   /* fast kernel: the longest stream's walk bounds the launch (more scan CTAs) */
	uint64_t item_first, n_items;     /* general path: this segment's decode work items */
	uint64_t g2_first;                /* ... and its slice of the per-stream arrays */
	uint64_t tile_first, n_tiles;     /* level <= 10: lift tiles */
	uint32_t n_deep;                  /* generic streams of level > 10 */
	uint64_t split_first, n_split;    /* split path (acm_split.cu): slice of d_streams, */
	uint64_t sp_item_first, sp_n_items; /* lift work items, */
	uint64_t sp_first;                /* slice of the per-stream arrays, */
	uint64_t sp_block0, sp_blocks;    /* and of the per-block arrays */
};

/*
 * Groups of the general path (single-segment plans: the resident path).  The scan is one serial walk
 * per stream, so a launch over the whole batch lasts as long as its longest walk while most lanes
 * have long since finished; unpack and lift are throughput kernels that only need the records of their
 * own streams.  The streams are therefore sorted by expected walk time and cut into groups, each on a
 * CUDA stream of its own: the short groups are unpacked and transformed while the long ones are still
 * being walked.  A group is a contiguous slice of the segment's streams, work items and tiles; its
 * items number streams from the group's first.
 */
struct GenGroup {
	uint64_t k0, n;                  /* streams [k0, k0 + n) of the segment's general streams */
	uint64_t item_first, n_items;
	uint64_t tile_first, n_tiles;
};
constexpr unsigned GEN_GROUPS_MAX = 8;
/* counters: per segment 4 queue words + 4 general-path words, 4 shared words, and the general-path words
 * of the groups after the first */
static inline size_t counter_words(size_t nseg) { return 8 * nseg + 4 + 4 * (GEN_GROUPS_MAX - 1); }

struct acm_gpu_plan {
	int device;
	uint64_t n;          /* caller's stream count */
	uint64_t n_dev;      /* streams that reach the device */
	uint64_t n_fast, n_generic;
	uint64_t blob_room;  /* bytes the kernels may read: end of the last image, rounded up to 16 */
	Format fmt;
	std::vector<int32_t> host_status; /* statuses decided on the host (NOT_ACM, OTHER) */
	std::vector<Segment> seg;
	DevStream *d_streams;
	int32_t *d_status;
	uint32_t *d_words;
	unsigned long long *d_cks;
	acm_tables *d_tables;
	uint32_t *d_counters; /* per segment g: [4g] fast queue, [4g+1] generic queue, [4g+2] finished scan warps, [4g+3] heartbeat; [4*nseg] error flag;
			       * [4*nseg+4+4g ..] general path: unpack queue, tile queue, scan queue */
	GenericScratch scratch;
	Gen2Args g2;         /* general path: block records, column offsets, work items (device pointers) */
	size_t g2_state_bytes; /* nscan + first_bad: reset before every run */
	size_t g2_cks_bytes;   /* per-block checksum slots: zeroed before a run with checksums (tiles add to them) */
	SplitArgs sp;        /* split path: the same for its streams, plus the intermediates between its kernels */
	size_t sp_state_bytes;
	uint64_t n_split;
	uint32_t *d_hist;    /* fast kernel history, 256 words per stream slot */
	uint8_t *d_ring;     /* fast kernel block-record rings */
	uint8_t *d_pool;     /* the one device allocation all of the above point into (null: arena) */
	uint8_t *d_ctl;      /* fast kernel slot control words */
	size_t ctl_bytes;
	unsigned long long *d_prof; /* 64 counters of -DF2_PROF tuning builds */
	int fast_ctas;
	int generic_ctas;
	int sm_count;
	cudaEvent_t ev0, ev1;
	bool timed;
	std::vector<GenGroup> grp;                 /* more than one entry: see GenGroup */
	cudaStream_t grp_st[GEN_GROUPS_MAX] = {};  /* group j > 0 runs on grp_st[j] */
	cudaEvent_t grp_ev[GEN_GROUPS_MAX] = {};   /* [0]: fork, [j]: group j done */
};

static void plan_free(acm_gpu_plan *p)
{
	if (!p)
		return;
	cudaFree(p->d_pool); /* null when the plan lives in a caller's arena */
	for (unsigned j = 0; j < GEN_GROUPS_MAX; j++) {
		if (p->grp_st[j])
			cudaStreamDestroy(p->grp_st[j]);
		if (p->grp_ev[j])
			cudaEventDestroy(p->grp_ev[j]);
	}
	if (p->ev0)
		cudaEventDestroy(p->ev0);
	if (p->ev1)
		cudaEventDestroy(p->ev1);
	delete p;
}

/* the code tables, built once per process */
static const acm_tables *host_tables()
{
	static acm_tables tab;
	static std::once_flag once;
	std::call_once(once, [] { acm_tables_build(&tab); });
	return &tab;
}

static bool fast_eligible(const acm_gpu_stream &g, const acm_gpu_opts *o)
{
	/* the register/shared-memory kernel covers the common block shape and the
	 * reference's own 16-bit formats; everything else takes the generic kernel */
	return o->wordlen == 2 && fast_shape(g.level, g.rows);
}

/* device memory kept between calls by the host path (grown, never shrunk) */
struct DevArena {
	uint8_t *base = nullptr;
	size_t cap = 0;
};

static acm_gpu_plan *plan_create(const acm_gpu_stream *s, uint64_t n, const acm_gpu_opts *opts,
				 unsigned nseg, DevArena *arena, int *err_out)
{
	acm_gpu_opts defaults;
	acm_gpu_plan *p = nullptr;
	std::vector<DevStream> all;
	std::vector<Gen2Stream> g2_streams; /* parallel to the generic streams, in `all` order */
	std::vector<Gen2Item> g2_items;
	std::vector<Gen2Item> g3_tiles;     /* lift work items of the level <= 10 streams */
	uint64_t g3_words = 0;              /* words of the int16 intermediate */
	std::vector<Gen2Stream> sp_streams; /* the same for the split path's streams */
	std::vector<Gen2Item> sp_items;
	uint64_t g2_blocks = 0, g2_coffs = 0, sp_blocks = 0;
	constexpr uint32_t G2_RUN = 16;
	constexpr uint32_t SP_RUN = 8; /* blocks per lift work item */
	int sm_count = 0, max_ctas = 0;
	int err = ACM_ERR_OTHER, dev = 0;
	uint32_t max_blen = 1, max_cols = 1;
	uint64_t max_fast = 0, max_gen = 0;

	g_err[0] = 0;
	if (!opts) {
		acm_gpu_opts_init(&defaults);
		opts = &defaults;
	}
	if (opts->wordlen < 2 || opts->wordlen > 4) {
		acm_set_error("wordlen %d not supported (2, 3 or 4)", opts->wordlen);
		err = ACM_ERR_BADFMT;
		goto fail;
	}
	if (use_device(opts) < 0)
		goto fail;
	CU(cudaGetDevice(&dev));
	CU(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));

	p = new acm_gpu_plan();
	memset(&p->scratch, 0, sizeof(p->scratch));
	memset(&p->g2, 0, sizeof(p->g2));
	p->g2_state_bytes = 0;
	p->g2_cks_bytes = 0;
	memset(&p->sp, 0, sizeof(p->sp));
	p->sp_state_bytes = 0;
	p->n_split = 0;
	p->d_streams = nullptr; p->d_status = nullptr; p->d_words = nullptr; p->d_cks = nullptr;
	p->d_tables = nullptr; p->d_counters = nullptr; p->ev0 = nullptr; p->ev1 = nullptr;
	p->d_hist = nullptr; p->fast_ctas = 0; p->generic_ctas = 0;
	p->d_ring = nullptr;
	p->d_pool = nullptr;
	p->d_ctl = nullptr;
	p->ctl_bytes = 0;
	p->d_prof = nullptr;
	p->device = dev;
	p->n = n;
	p->n_fast = p->n_generic = 0;
	p->sm_count = sm_count;
	p->timed = false;
	p->fmt.wordlen = opts->wordlen;
	p->fmt.be = opts->bigendianp ? 1 : 0;
	p->fmt.bias = opts->sgned ? 0u : (1u << (8 * opts->wordlen - 1));
	p->fmt.checksums = opts->want_checksums ? 1 : 0;
	p->host_status.assign(n, 0);
	p->blob_room = 0;

	if (nseg < 1)
		nseg = 1;
	if (nseg > n)
		nseg = n ? (unsigned)n : 1;
	/* segments of the host path run concurrently (copy-in / decode / copy-out pipeline): a
	 * segment's launch is latency-bound by its longest stream, so each gets a share of the SMs
	 * and up to four of them are resident together */
	max_ctas = nseg > 1 ? sm_count / (int)(nseg < 4 ? nseg : 4) : sm_count;
	{
		/* segment boundaries by output words.  The pipeline's step is the first segment's latency
		 * (copy-in + a decode that lasts as long as the walk of its longest stream, whatever its size)
		 * plus the PCM's trip to the host, during which the copy engine should never wait for a decode:
		 * the first segments are small (the copy-out starts early, and their copies are short enough for
		 * the next decode to be done in time), the later ones make up for it: shares 1 : 2 : 3 : 4 : 5 : 6 : 6 ... */
		uint64_t total = 0, acc = 0, first = 0;
		double wsum = 0.0, wacc = 0.0;
		static const bool ramp = !(getenv("ACM_B200_SEG_RAMP") && atoi(getenv("ACM_B200_SEG_RAMP")) == 0);
		auto seg_weight = [nseg](unsigned g) { return ramp && nseg >= 4 ? (g < 5 ? 1.0 + g : 6.0) : 1.0; };
		for (uint64_t i = 0; i < n; i++)
			total += s[i].total_values;
		for (unsigned g = 0; g < nseg; g++)
			wsum += seg_weight(g);
		for (unsigned g = 0; g < nseg; g++) {
			Segment sg;
			memset(&sg, 0, sizeof(sg));
			sg.s_first = first;
			wacc += seg_weight(g);
			uint64_t target = (uint64_t)((double)total * (wacc / wsum));
			uint64_t i = first;
			if (g + 1 == nseg) {
				i = n;
			} else {
				while (i < n && (acc < target || i == first)) {
					acc += s[i].total_values;
					i++;
				}
				if (n - i < nseg - 1 - g) /* leave at least one stream per remaining segment */
					i = n - (nseg - 1 - g);
			}
			sg.s_count = i - first;
			first = i;
			p->seg.push_back(sg);
		}
	}
	for (Segment &sg : p->seg) {
		std::vector<DevStream> fast, gen, spl;
		uint64_t slot_bytes = 0;
		sg.blob_lo = sg.out_lo = ~(uint64_t)0;
		sg.sparse = !opts->pad_tail;
		for (uint64_t i = sg.s_first; i < sg.s_first + sg.s_count; i++) {
			const acm_gpu_stream &g = s[i];
			if (acm_stream_accepted(&g) && (g.out_off & 15u)) {
				acm_set_error("stream %llu: out_off must be a multiple of 16", (unsigned long long)i);
				goto fail;
			}
			DevStream d;
			int hs = acm_make_devstream(&g, (uint32_t)i, opts->pad_tail, &d);
			if (hs < 0) {
				p->host_status[i] = hs;
				if (g.total_values)
					sg.sparse = true; /* a refused stream's slot stays unwritten */
				continue;
			}
			uint32_t blen = g.rows << g.level;
			uint64_t in_hi = (g.in_off + g.in_len + 15u) & ~(uint64_t)15u;
			uint64_t o_hi = g.out_off + (((uint64_t)g.total_values * opts->wordlen + 15u) & ~(uint64_t)15u);
			p->blob_room = std::max(p->blob_room, in_hi);
			sg.blob_lo = std::min(sg.blob_lo, g.in_off & ~(uint64_t)15u);
			sg.blob_hi = std::max(sg.blob_hi, in_hi);
			sg.out_lo = std::min(sg.out_lo, g.out_off);
			sg.out_hi = std::max(sg.out_hi, o_hi);
			slot_bytes += o_hi - g.out_off;
			if (opts->kernel == 2 && opts->wordlen == 2 && split_shape(g.level, g.rows)) {
				spl.push_back(d);
			} else if (opts->kernel != 1 && fast_eligible(g, opts)) {
				fast.push_back(d);
			} else {
				gen.push_back(d);
				max_blen = std::max(max_blen, blen);
				max_cols = std::max(max_cols, 1u << g.level);
			}
		}
		/* A small group of level-7 / 16-row streams inside a batch that is mostly other shapes: the
		 * fused kernel's launch would last as long as the walk of its longest stream (37 us per block)
		 * whatever the group's size, while the general path walks them alongside its own streams for
		 * free -- they go with the majority. */
		if (opts->kernel == 0 && !fast.empty() && !gen.empty()) {
			uint64_t wf = 0, wg = 0;
			for (const DevStream &d : fast)
				wf += (uint64_t)d.n_attempt * (d.rows << d.level);
			for (const DevStream &d : gen)
				wg += (uint64_t)d.n_attempt * (d.rows << d.level);
			if (wf * 4u < wg) {
				for (const DevStream &d : fast) {
					gen.push_back(d);
					max_blen = std::max(max_blen, d.rows << d.level);
					max_cols = std::max(max_cols, 1u << d.level);
				}
				fast.clear();
			}
		}
		if (sg.blob_lo > sg.blob_hi)
			sg.blob_lo = sg.blob_hi = 0;
		if (sg.out_lo > sg.out_hi)
			sg.out_lo = sg.out_hi = 0;
		if (slot_bytes != sg.out_hi - sg.out_lo)
			sg.sparse = true; /* gaps (or overlaps) in a caller-chosen layout */
		/* longest first: the work queues hand out streams in this order (LPT) */
		auto by_len = [](const DevStream &a, const DevStream &b) {
			if (a.n_attempt != b.n_attempt)
				return a.n_attempt > b.n_attempt;
			return a.index < b.index;
		};
		std::sort(fast.begin(), fast.end(), by_len);
		if (!fast.empty()) {
			/* The streams that start together (one per stream slot) are dealt out to the scan
			 * warps like cards: every warp -- 32 consecutive queue entries -- gets one of the very
			 * longest streams, one of the next longest, and so on down.  A scan warp walks its 32
			 * lanes in lockstep, a round lasts as long as its slowest lane; with the longest
			 * streams spread out, the warps that finish last (the launch is latency-bound on them)
			 * run their final rounds with one or two busy lanes instead of 32.  The rest of the
			 * queue stays longest-first. */
			uint32_t n_scan = 0, n_work = 0, n_slots = 0;
			uint64_t blocks = 0;
			for (const DevStream &d : fast)
				blocks += d.n_attempt;
			sg.walk_bound = fast2_walk_bound(fast[0].n_attempt, blocks, sm_count, max_ctas);
			uint32_t scan_warps = 0;
			fast2_geometry(fast.size(), sm_count, max_ctas, &n_scan, &n_work, &n_slots, sg.walk_bound, &scan_warps);
			const size_t head = std::min<size_t>(n_slots, fast.size()) / 32 * 32, warps = head / 32;
			int deal = sg.walk_bound && scan_warps == (uint32_t)fast2_scan_warps() ? 1 : 0; /* no partners in a sparse launch */
			if (const char *e = getenv("ACM_B200_DEAL"))
				deal = atoi(e);
			if (warps > 1 && deal == 1) {
				/* A walk-bound launch (few, long streams): what counts is when the longest walks end, and
				 * a scan warp that has its SM sub-partition to itself steps about a third faster than one
				 * that shares it (the scan SMs are bound by instruction issue).  Slot g is lane g % 32 of
				 * warp (g / 32) % SW of scan CTA g / SLOTS, and warps w and w + SW / 2 of a CTA share a
				 * sub-partition: the lower warp of every pair gets 32 of the longest streams, its partner 32
				 * of the shortest -- the longest of all are paired with the shortest of all -- so that the
				 * partner is done, and gone, when the long walks have most of their way still to go. */
				const size_t sw = (size_t)fast2_scan_warps(), half = sw / 2;
				std::vector<size_t> wa, wb; /* warp numbers of the lower / upper halves, in pair order */
				for (size_t w = 0; w < warps; w++)
					((w % sw) < half ? wa : wb).push_back(w);
				std::vector<DevStream> dealt(head);
				size_t r = 0;
				for (size_t j = 0; j < wa.size(); j++, r += 32)
					std::copy(fast.begin() + r, fast.begin() + r + 32, dealt.begin() + wa[j] * 32);
				/* what is left, shortest first, to the partners in the same pair order */
				for (size_t j = 0; j < wb.size(); j++) {
					const size_t from = head - 32 * (j + 1);
					std::copy(fast.begin() + from, fast.begin() + from + 32, dealt.begin() + wb[j] * 32);
				}
				std::copy(dealt.begin(), dealt.end(), fast.begin());
			} else if (warps > 1) {
				std::vector<DevStream> dealt(head);
				for (size_t r = 0; r < head; r++)
					dealt[(r % warps) * 32 + r / warps] = fast[r];
				std::copy(dealt.begin(), dealt.end(), fast.begin());
			}
		}
		/* general path: longest WALK first.  Expected walk steps of a block: one selector step per
		 * column plus, for a prefix-coded column (8 of the 28 valid selectors), a table step per ~4 rows:
		 * cols * (1 + rows / 14)
		 * -- or, for a stream of long columns (f_linear of many rows), what its ring can take in: 1024 bits
		 * per top-up period of 16 steps (acm_walk.cuh).  In units of 1 / (14 * 64) step */
		auto walk_cost = [](const DevStream &d) {
			const uint64_t steps = (uint64_t)d.n_attempt * ((uint64_t)(14u + d.rows) << d.level) * 64u;
			const uint64_t bits = d.file_end > d.bit0 ? (uint64_t)(d.file_end - d.bit0) * 14u : 0u;
			return steps > bits ? steps : bits;
		};
		auto by_walk = [&walk_cost](const DevStream &a, const DevStream &b) {
			const uint64_t wa = walk_cost(a), wb = walk_cost(b);
			if (wa != wb)
				return wa > wb;
			return a.index < b.index;
		};
		std::sort(gen.begin(), gen.end(), by_walk);
		std::vector<uint64_t> item_at, tile_at; /* first work item / tile of every general stream */
		std::sort(spl.begin(), spl.end(), by_len);
		sg.fast_first = all.size();
		sg.n_fast = fast.size();
		all.insert(all.end(), fast.begin(), fast.end());
		sg.split_first = all.size();
		sg.n_split = spl.size();
		all.insert(all.end(), spl.begin(), spl.end());
		/* split path: a record per block the image can hold; lift work items = runs of SP_RUN
		 * consecutive blocks (a run rebuilds the transform history from the block before it) */
		sg.sp_first = sp_streams.size();
		sg.sp_item_first = sp_items.size();
		sg.sp_block0 = sp_blocks;
		for (size_t k = 0; k < spl.size(); k++) {
			const DevStream &d = spl[k];
			Gen2Stream gs;
			memset(&gs, 0, sizeof(gs));
			gs.rec_base = sp_blocks - sg.sp_block0;
			gs.max_blocks = (uint32_t)gen2_max_blocks(d.n_attempt, d.file_end > d.bit0 ? d.file_end - d.bit0 : 0, d.level);
			sp_blocks += gs.max_blocks;
			sp_streams.push_back(gs);
		}
		/* items block-range major: the first runs of all streams, then the second runs, ...: neighbours in
		 * the queue read neighbouring records, and the longest streams' last runs are the last items */
		{
			uint32_t longest = 0;
			for (size_t k = 0; k < spl.size(); k++)
				longest = std::max(longest, sp_streams[sg.sp_first + k].max_blocks);
			for (uint32_t b0 = 0; b0 < longest; b0 += SP_RUN)
				for (size_t k = 0; k < spl.size(); k++) {
					const uint32_t mb = sp_streams[sg.sp_first + k].max_blocks;
					if (b0 >= mb)
						continue;
					Gen2Item it;
					it.stream = (uint32_t)k;
					it.b0 = b0;
					it.nb = std::min<uint32_t>(SP_RUN, mb - b0);
					it.warm = 0;
					sp_items.push_back(it);
				}
		}
		sg.sp_n_items = sp_items.size() - sg.sp_item_first;
		sg.sp_blocks = sp_blocks - sg.sp_block0;
		p->n_split += sg.n_split;
		sg.gen_first = all.size();
		sg.n_gen = gen.size();
		all.insert(all.end(), gen.begin(), gen.end());
		/* general path: block records for every block the image can hold, and the decode work items:
		 * runs of G2_RUN consecutive blocks, each with the warm-up blocks that rebuild the transform
		 * history (the look-back is 2*cols-2 words) */
		sg.item_first = g2_items.size();
		sg.g2_first = g2_streams.size();
		sg.tile_first = g3_tiles.size();
		sg.n_deep = 0;
		for (size_t k = 0; k < gen.size(); k++) {
			const DevStream &d = gen[k];
			Gen2Stream gs;
			memset(&gs, 0, sizeof(gs));
			const uint32_t cols = 1u << d.level, blen = d.rows << d.level;
			item_at.push_back(g2_items.size());
			tile_at.push_back(g3_tiles.size());
			gs.rec_base = g2_blocks;
			gs.coff_base = g2_coffs;
			gs.max_blocks = (uint32_t)gen2_max_blocks(d.n_attempt, d.file_end > d.bit0 ? d.file_end - d.bit0 : 0, d.level);
			g2_blocks += gs.max_blocks;
			g2_coffs += (uint64_t)gs.max_blocks * cols;
			if (d.level <= GEN3_MAX_LEVEL) {
				/* unpack -> int16 intermediate -> tile lift; tiles cover what the read loop can deliver */
				const uint64_t wcap = std::min<uint64_t>((uint64_t)gs.max_blocks * blen, d.words_limit);
				gs.word_base = g3_words;
				g3_words += ((uint64_t)gs.max_blocks * blen + 7u) & ~(uint64_t)7u;
				for (uint64_t t0 = 0; t0 < wcap; t0 += GEN3_TILE_WORDS) {
					Gen2Item it;
					it.stream = (uint32_t)k;
					it.b0 = (uint32_t)(t0 / GEN3_TILE_WORDS);
					it.nb = 1;
					it.warm = 0;
					g3_tiles.push_back(it);
				}
			} else {
				sg.n_deep++;
			}
			g2_streams.push_back(gs);
			const uint32_t look = 2u * cols - 2u;
			const uint32_t warm_full = look ? (look + blen - 1u) / blen : 0u;
			for (uint32_t b0 = 0; b0 < gs.max_blocks; b0 += G2_RUN) {
				Gen2Item it;
				it.stream = (uint32_t)k;
				it.b0 = b0;
				it.nb = std::min<uint32_t>(G2_RUN, gs.max_blocks - b0);
				it.warm = std::min(warm_full, b0);
				g2_items.push_back(it);
			}
		}
		sg.n_items = g2_items.size() - sg.item_first;
		sg.n_tiles = g3_tiles.size() - sg.tile_first;
		item_at.push_back(g2_items.size());
		tile_at.push_back(g3_tiles.size());
		size_t group_min = 1024;
		if (const char *e = getenv("ACM_B200_GEN_GROUP_MIN"))
			group_min = (size_t)atol(e);
		if (nseg == 1 && sg.n_deep == 0 && !gen.empty() && gen.size() >= group_min) {
			/* groups (GenGroup): equal slices of the longest stream's expected walk time */
			unsigned ng = 8;
			if (const char *e = getenv("ACM_B200_GEN_GROUPS"))
				ng = (unsigned)atoi(e);
			ng = ng < 1 ? 1 : ng > GEN_GROUPS_MAX ? GEN_GROUPS_MAX : ng;
			const uint64_t c0 = walk_cost(gen[0]);
			size_t k = 0;
			for (unsigned j = 0; j < ng && k < gen.size(); j++) {
				/* group j ends at (1 - (j + 1) / ng) ^ 1.5 of the longest walk (measured on the filler-stress
				 * corpus, 8 groups: exponent 0.6 27.0 ms, 0.8 24.8, 1.0 24.3, 1.3 - 2.0 24.0, 2.6 24.3; 5 groups
				 * 25.0 - 25.9): the groups of short streams, whose unpack and lift run beside everybody
				 * else's walk, are the thin ones */
				static const double gpow = getenv("ACM_B200_GEN_GROUP_POW") ? atof(getenv("ACM_B200_GEN_GROUP_POW")) : 1.5; /* tuning */
				const uint64_t lo = j + 1 == ng ? 0 : (uint64_t)((double)c0 * pow(1.0 - (double)(j + 1) / ng, gpow) * (1.0 - 0.4 / ng));
				GenGroup gg;
				gg.k0 = k;
				while (k < gen.size() && (walk_cost(gen[k]) > lo || j + 1 == ng))
					k++;
				gg.n = k - gg.k0;
				if (!gg.n)
					continue;
				gg.item_first = item_at[gg.k0];
				gg.n_items = item_at[k] - gg.item_first;
				gg.tile_first = tile_at[gg.k0];
				gg.n_tiles = tile_at[k] - gg.tile_first;
				for (uint64_t i = gg.item_first; i < gg.item_first + gg.n_items; i++)
					g2_items[i].stream -= (uint32_t)gg.k0;
				for (uint64_t i = gg.tile_first; i < gg.tile_first + gg.n_tiles; i++)
					g3_tiles[i].stream -= (uint32_t)gg.k0;
				p->grp.push_back(gg);
			}
			if (p->grp.size() < 2)
				p->grp.clear(); /* one group: numbering unchanged (k0 = 0) */
		}
		p->n_fast += sg.n_fast;
		p->n_generic += sg.n_gen;
		max_fast = std::max(max_fast, sg.n_fast);
		max_gen = std::max(max_gen, sg.n_gen);
	}
	p->n_dev = all.size();

	CU(cudaEventCreate(&p->ev0));
	CU(cudaEventCreate(&p->ev1));
	for (size_t j = 0; j < p->grp.size(); j++) {
		CU(cudaEventCreateWithFlags(&p->grp_ev[j], cudaEventDisableTiming));
		if (j)
			CU(cudaStreamCreateWithFlags(&p->grp_st[j], cudaStreamNonBlocking));
	}
	{
		/* ---- geometry of every segment, then ONE device allocation for the whole plan (or a
		 * slice of the caller's arena: the host path of acm_gpu_decode_batch keeps one
		 * between calls, cudaMalloc / cudaFree synchronise the device) */
		const size_t stride = generic_scratch_words(max_blen, max_cols);
		const size_t budget = ((size_t)4 << 30) / p->seg.size(); /* scratch bytes per segment */
		size_t hist_words = 0, scratch_words = 0, ring_bytes = 0, ctl_bytes = 0;
		for (Segment &sg : p->seg) {
			sg.hist_off = hist_words;
			sg.ring_off = ring_bytes;
			sg.ctl_off = ctl_bytes;
			sg.n_scan = sg.n_slots = 0;
			sg.fast_ctas = 0;
			if (sg.n_fast) {
				uint32_t n_work = 0;
				fast2_geometry(sg.n_fast, p->sm_count, max_ctas, &sg.n_scan, &n_work, &sg.n_slots, sg.walk_bound, &sg.scan_warps);
				sg.fast_ctas = (int)(sg.n_scan + n_work);
			}
			hist_words += (size_t)sg.n_slots * fast2_hist_words_per_slot();
			ring_bytes += (size_t)sg.n_slots * fast2_ring_bytes_per_slot();
			ctl_bytes += ((size_t)sg.n_slots * fast2_ctl_bytes_per_slot() + 255u) & ~(size_t)255u;
			int ctas = p->sm_count * gen2_ctas_per_sm();
			if ((uint64_t)ctas > sg.n_items)
				ctas = (int)sg.n_items;
			if (ctas < 1 && sg.n_gen)
				ctas = 1;
			while (ctas > 1 && (size_t)ctas * stride * 4 > budget)
				ctas /= 2;
			sg.gen_ctas = ctas;
			sg.scratch_off = scratch_words;
			scratch_words += (size_t)ctas * stride;
			p->fast_ctas = std::max(p->fast_ctas, sg.fast_ctas);
			p->generic_ctas = std::max(p->generic_ctas, sg.gen_ctas);
		}
		size_t total = 0;
		auto carve = [&total](size_t bytes) {
			const size_t at = total;
			total += (bytes + 255u) & ~(size_t)255u;
			return at;
		};
		const size_t o_streams = carve((p->n_dev + 1) * sizeof(DevStream));
		const size_t o_results = carve((n + 1) * 16); /* status (4) | words (4) | checksums (8) */
		const size_t o_tables = carve(sizeof(acm_tables));
		const size_t o_counters = carve(counter_words(p->seg.size()) * sizeof(uint32_t));
		const size_t o_prof = carve(64 * sizeof(unsigned long long));
		const size_t o_ctl = carve(ctl_bytes);
		const size_t o_hist = carve(hist_words * 4 + 16);
		const size_t o_ring = carve(ring_bytes + 16);
		const size_t o_scratch = carve(max_gen ? scratch_words * 4 + 16 : 0);
		const size_t n_g2 = g2_streams.size();
		if ((g2_blocks * (sizeof(BlockRec) + 8) + g2_coffs * 4) >> 35) {
			acm_set_error("general path: %llu blocks need more than 32 GB of block records", (unsigned long long)g2_blocks);
			goto fail;
		}
		const size_t o_g2s = carve((n_g2 + 1) * sizeof(Gen2Stream));
		const size_t o_g2items = carve((g2_items.size() + 1) * sizeof(Gen2Item));
		const size_t o_g2state = carve((n_g2 + 1) * 8); /* nscan | first_bad */
		const size_t o_g2rec = carve((g2_blocks + 1) * sizeof(BlockRec));
		const size_t o_g2cks = carve((g2_blocks + 1) * 8);
		const size_t o_g2coff = carve((g2_coffs + 4) * 4);
		const size_t o_g3tiles = carve((g3_tiles.size() + 1) * sizeof(Gen2Item));
		const size_t o_g3inter = carve(g3_words ? (g3_words + 8) * 2 : 0);
		if (g3_words >> 35) {
			acm_set_error("general path: %llu words of intermediate", (unsigned long long)g3_words);
			goto fail;
		}
		/* split path: per-stream tables and state, work items, and per block: record, checksum,
		 * column offsets, index bytes, wide mask, side array */
		const size_t n_sp = sp_streams.size();
		if ((sp_blocks * split_bytes_per_block()) >> 37) {
			acm_set_error("split path: %llu blocks need more than 128 GB of intermediates", (unsigned long long)sp_blocks);
			goto fail;
		}
		const size_t o_sps = carve((n_sp + 1) * sizeof(Gen2Stream));
		const size_t o_spitems = carve((sp_items.size() + 1) * sizeof(Gen2Item));
		const size_t o_spstate = carve((n_sp + 1) * 8);
		const size_t o_sprec = carve((sp_blocks + 1) * sizeof(BlockRec));
		const size_t o_spcks = carve((sp_blocks + 1) * 8);
		const size_t o_spcoff = carve((sp_blocks + 1) * 256);
		const size_t o_spwmask = carve((sp_blocks + 1) * 16);
		const size_t o_spinter = carve(n_sp ? (sp_blocks + 1) * 2048 : 0);
		const size_t o_spwide = carve(n_sp ? (sp_blocks + 1) * 4096 : 0);
		uint8_t *base = nullptr;
		if (arena) {
			if (arena->cap < total) {
				cudaFree(arena->base);
				arena->base = nullptr;
				arena->cap = 0;
				CU(cudaMalloc(&arena->base, total + total / 4));
				arena->cap = total + total / 4;
			}
			base = arena->base;
		} else {
			CU(cudaMalloc(&p->d_pool, total));
			base = p->d_pool;
		}
		p->d_streams = reinterpret_cast<DevStream *>(base + o_streams);
		p->d_cks = reinterpret_cast<unsigned long long *>(base + o_results);
		p->d_status = reinterpret_cast<int32_t *>(base + o_results + (n + 1) * 8);
		p->d_words = reinterpret_cast<uint32_t *>(base + o_results + (n + 1) * 12);
		p->d_tables = reinterpret_cast<acm_tables *>(base + o_tables);
		p->d_counters = reinterpret_cast<uint32_t *>(base + o_counters);
		p->d_prof = reinterpret_cast<unsigned long long *>(base + o_prof);
		p->d_ctl = base + o_ctl;
		p->ctl_bytes = ctl_bytes;
		p->d_hist = reinterpret_cast<uint32_t *>(base + o_hist);
		p->d_ring = base + o_ring;
		if (max_gen) {
			p->scratch.stride = stride;
			p->scratch.max_blen = max_blen;
			p->scratch.max_cols = max_cols;
			p->scratch.buf = reinterpret_cast<uint32_t *>(base + o_scratch);
		}
		p->g2.gs = reinterpret_cast<const Gen2Stream *>(base + o_g2s);
		p->g2.items = reinterpret_cast<const Gen2Item *>(base + o_g2items);
		p->g2.nscan = reinterpret_cast<uint32_t *>(base + o_g2state);
		p->g2.first_bad = p->g2.nscan + (n_g2 + 1);
		p->g2_state_bytes = (n_g2 + 1) * 8;
		p->g2.rec = reinterpret_cast<BlockRec *>(base + o_g2rec);
		p->g2.cks_blk = reinterpret_cast<unsigned long long *>(base + o_g2cks);
		p->g2.coff = reinterpret_cast<uint32_t *>(base + o_g2coff);
		p->g2.tiles = reinterpret_cast<const Gen2Item *>(base + o_g3tiles);
		p->g2.inter16 = reinterpret_cast<int16_t *>(base + o_g3inter);
		p->g2_cks_bytes = (g2_blocks + 1) * 8;
		if (!g3_tiles.empty())
			CU(cudaMemcpy(base + o_g3tiles, g3_tiles.data(), g3_tiles.size() * sizeof(Gen2Item), cudaMemcpyHostToDevice));
		p->sp.gs = reinterpret_cast<const Gen2Stream *>(base + o_sps);
		p->sp.items = reinterpret_cast<const Gen2Item *>(base + o_spitems);
		p->sp.nscan = reinterpret_cast<uint32_t *>(base + o_spstate);
		p->sp.first_bad = p->sp.nscan + (n_sp + 1);
		p->sp_state_bytes = (n_sp + 1) * 8;
		p->sp.rec = reinterpret_cast<BlockRec *>(base + o_sprec);
		p->sp.cks_blk = reinterpret_cast<unsigned long long *>(base + o_spcks);
		p->sp.coff16 = reinterpret_cast<uint16_t *>(base + o_spcoff);
		p->sp.wmask = reinterpret_cast<uint32_t *>(base + o_spwmask);
		p->sp.inter = base + o_spinter;
		p->sp.wide = reinterpret_cast<uint16_t *>(base + o_spwide);
		if (n_sp)
			CU(cudaMemcpy(base + o_sps, sp_streams.data(), n_sp * sizeof(Gen2Stream), cudaMemcpyHostToDevice));
		if (!sp_items.empty())
			CU(cudaMemcpy(base + o_spitems, sp_items.data(), sp_items.size() * sizeof(Gen2Item), cudaMemcpyHostToDevice));
		if (n_g2)
			CU(cudaMemcpy(base + o_g2s, g2_streams.data(), n_g2 * sizeof(Gen2Stream), cudaMemcpyHostToDevice));
		if (!g2_items.empty())
			CU(cudaMemcpy(base + o_g2items, g2_items.data(), g2_items.size() * sizeof(Gen2Item), cudaMemcpyHostToDevice));
		if (p->n_dev)
			CU(cudaMemcpy(p->d_streams, all.data(), p->n_dev * sizeof(DevStream), cudaMemcpyHostToDevice));
		CU(cudaMemset(base + o_results, 0, (n + 1) * 16));
		CU(cudaMemset(p->d_prof, 0, 64 * sizeof(unsigned long long)));
		CU(cudaMemcpy(p->d_tables, host_tables(), sizeof(acm_tables), cudaMemcpyHostToDevice));
		/* the kernels run on non-blocking streams (or the caller's), which do not wait for the
		 * legacy default stream: the descriptor / table uploads and the memsets above must have
		 * landed before this function returns (a pageable cudaMemcpy may return before its DMA) */
		CU(cudaStreamSynchronize(cudaStreamLegacy));
	}
	if (err_out)
		*err_out = ACM_OK;
	return p;
fail:
	plan_free(p);
	if (err_out)
		*err_out = err;
	return nullptr;
}

extern "C" acm_gpu_plan *acm_gpu_plan_create(const acm_gpu_stream *s, uint64_t n,
					     const acm_gpu_opts *opts, int *err_out)
{
	return plan_create(s, n, opts, 1, nullptr, err_out);
}

/* launch the kernels of one segment on `st`; the cursors must have been zeroed on that stream */
static int plan_run_segment(acm_gpu_plan *p, size_t g, const void *d_blob, void *d_out, cudaStream_t st)
{
	const Segment &sg = p->seg[g];
	KernelArgs a;
	a.blob = (const uint8_t *)d_blob;
	a.blob_room = p->blob_room;
	a.errflag = p->d_counters + 4 * p->seg.size();
	a.out = (uint8_t *)d_out;
	a.status = p->d_status;
	a.words = p->d_words;
	a.cks = p->d_cks;
	a.tables = p->d_tables;
	a.fmt = p->fmt;
	a.hist = p->d_hist ? p->d_hist + sg.hist_off : nullptr;
	a.ring = p->d_ring ? p->d_ring + sg.ring_off : nullptr;
	a.prof = p->d_prof;
	a.slotctl = p->d_ctl ? p->d_ctl + sg.ctl_off : nullptr;
	a.scan_done = p->d_counters + 4 * g + 2;
	a.n_scan = sg.n_scan;
	a.scan_warps = sg.scan_warps;
	a.n_slots = sg.n_slots;
	a.resume_hist = nullptr;
	a.resume_stride = 0;
	a.end_pos = nullptr;
	if (sg.n_fast) {
		a.streams = p->d_streams + sg.fast_first;
		a.count = (uint32_t)sg.n_fast;
		a.counter = p->d_counters + 4 * g;
		CU(launch_fast2(a, sg.fast_ctas, st));
	}
	if (sg.n_split) {
		/* records of one run are told from stale ones by the run's epoch: unique per process */
		static std::atomic<uint32_t> epoch_counter{1};
		SplitArgs sp = p->sp;
		a.streams = p->d_streams + sg.split_first;
		a.count = (uint32_t)sg.n_split;
		a.counter = p->d_counters + 4 * g;
		sp.gs += sg.sp_first;
		sp.nscan += sg.sp_first;
		sp.first_bad += sg.sp_first;
		sp.items += sg.sp_item_first;
		sp.n_items = (uint32_t)sg.sp_n_items;
		sp.item_counter = p->d_counters + 4 * g + 2;
		sp.rec += sg.sp_block0;
		sp.cks_blk += sg.sp_block0;
		sp.coff16 += sg.sp_block0 * 128u;
		sp.wmask += sg.sp_block0 * 4u;
		sp.inter += sg.sp_block0 * 2048u;
		sp.wide += sg.sp_block0 * 2048u;
		sp.n_blocks = sg.sp_blocks;
		sp.epoch = epoch_counter.fetch_add(1);
		CU(launch_split(a, sp, p->sm_count, st));
	}
	if (sg.n_gen) {
		GenericScratch sc = p->scratch;
		sc.buf += sg.scratch_off;
		a.streams = p->d_streams + sg.gen_first;
		a.count = (uint32_t)sg.n_gen;
		a.counter = p->d_counters + 4 * g + 1;
		Gen2Args g2 = p->g2;
		g2.gs += sg.g2_first;
		g2.nscan += sg.g2_first;
		g2.first_bad += sg.g2_first;
		g2.items += sg.item_first;
		g2.n_items = (uint32_t)sg.n_items;
		g2.item_counter = p->d_counters + 4 * g + 1;
		g2.tiles += sg.tile_first;
		g2.n_tiles = (uint32_t)sg.n_tiles;
		g2.n_deep = sg.n_deep;
		g2.g3_counters = p->d_counters + 4 * p->seg.size() + 4 + 4 * g;
		if (p->grp.size() > 1) {
			/* one CUDA stream per group, forked from and joined to st (GenGroup) */
			const Gen2Args all = g2;
			const DevStream *const streams = a.streams;
			CU(cudaEventRecord(p->grp_ev[0], st));
			for (size_t j = 0; j < p->grp.size(); j++) {
				const GenGroup &gg = p->grp[j];
				cudaStream_t gst = j ? p->grp_st[j] : st;
				if (j)
					CU(cudaStreamWaitEvent(gst, p->grp_ev[0], 0));
				a.streams = streams + gg.k0;
				a.count = (uint32_t)gg.n;
				g2 = all;
				g2.gs += gg.k0;
				g2.nscan += gg.k0;
				g2.first_bad += gg.k0;
				g2.items = all.items + gg.item_first;
				g2.n_items = (uint32_t)gg.n_items;
				g2.tiles = all.tiles + gg.tile_first;
				g2.n_tiles = (uint32_t)gg.n_tiles;
				g2.g3_counters = all.g3_counters + 4 * j;
				CU(launch_gen2(a, g2, sc, sg.gen_ctas, gst));
				if (j)
					CU(cudaEventRecord(p->grp_ev[j], gst));
			}
			for (size_t j = 1; j < p->grp.size(); j++)
				CU(cudaStreamWaitEvent(st, p->grp_ev[j], 0));
		} else {
			CU(launch_gen2(a, g2, sc, sg.gen_ctas, st));
		}
	}
	return ACM_OK;
fail:
	return ACM_ERR_OTHER;
}

extern "C" int acm_gpu_plan_run(acm_gpu_plan *p, const void *d_blob, void *d_out, void *cuda_stream)
{
	cudaStream_t st = (cudaStream_t)cuda_stream;
	g_err[0] = 0;
	if (((uintptr_t)d_blob & 15u) || ((uintptr_t)d_out & 15u)) {
		acm_set_error("acm_gpu_plan_run: blob and out must be 16-byte aligned");
		return ACM_ERR_OTHER;
	}
	CU(cudaSetDevice(p->device));
	CU(cudaMemsetAsync(p->d_counters, 0, counter_words(p->seg.size()) * sizeof(uint32_t), st));
	if (p->d_ctl)
		CU(cudaMemsetAsync(p->d_ctl, 0, p->ctl_bytes, st));
	if (p->n_generic) {
		CU(cudaMemsetAsync(p->g2.nscan, 0, p->g2_state_bytes / 2, st));
		CU(cudaMemsetAsync(p->g2.first_bad, 0xFF, p->g2_state_bytes / 2, st));
		if (p->fmt.checksums)
			CU(cudaMemsetAsync(p->g2.cks_blk, 0, p->g2_cks_bytes, st));
	}
	if (p->n_split) {
		CU(cudaMemsetAsync(p->sp.nscan, 0, p->sp_state_bytes / 2, st));
		CU(cudaMemsetAsync(p->sp.first_bad, 0xFF, p->sp_state_bytes / 2, st));
	}
	CU(cudaEventRecord(p->ev0, st));
	for (size_t g = 0; g < p->seg.size(); g++)
		if (plan_run_segment(p, g, d_blob, d_out, st) < 0)
			return ACM_ERR_OTHER;
	CU(cudaEventRecord(p->ev1, st));
	p->timed = true;
	return ACM_OK;
fail:
	return ACM_ERR_OTHER;
}

extern "C" int acm_gpu_plan_fetch(acm_gpu_plan *p, acm_gpu_stream *s, void *cuda_stream)
{
	cudaStream_t st = (cudaStream_t)cuda_stream;
	std::vector<int32_t> status(p->n);
	std::vector<uint32_t> words(p->n);
	std::vector<unsigned long long> cks(p->n);
	uint32_t flag = 0;
	g_err[0] = 0;
	CU(cudaSetDevice(p->device));
	CU(cudaStreamSynchronize(st));
	CU(cudaMemcpy(&flag, p->d_counters + 4 * p->seg.size(), sizeof(flag), cudaMemcpyDeviceToHost));
	if (flag) {
		acm_set_error("decode kernel reported internal failure %u", flag);
		return ACM_ERR_OTHER;
	}
	if (p->n) {
		CU(cudaMemcpy(status.data(), p->d_status, p->n * sizeof(int32_t), cudaMemcpyDeviceToHost));
		CU(cudaMemcpy(words.data(), p->d_words, p->n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
		CU(cudaMemcpy(cks.data(), p->d_cks, p->n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
	}
	for (uint64_t i = 0; i < p->n; i++) {
		if (p->host_status[i] < 0) {
			s[i].status = p->host_status[i];
			s[i].words = 0;
			s[i].checksum = 0;
		} else {
			s[i].status = status[i];
			s[i].words = words[i];
			s[i].checksum = cks[i];
		}
	}
	return ACM_OK;
fail:
	return ACM_ERR_OTHER;
}

extern "C" int acm_gpu_plan_launches(const acm_gpu_plan *p)
{
	int k = 0;
	for (const Segment &sg : p->seg) {
		k += (sg.n_fast ? 1 : 0) + (sg.n_split ? 3 + (sg.sp_n_items ? 1 : 0) : 0); /* split path: walk, unpack, lift, finish */
		if (!sg.n_gen)
			continue;
		/* general path (launch_gen2): scan, unpack + tile lift (levels <= 10), blocks (levels > 10), finish -- per group */
		if (p->grp.size() > 1)
			for (const GenGroup &gg : p->grp)
				k += 2 + (gg.n_tiles ? 2 : 0);
		else
			k += 2 + (sg.n_tiles ? 2 : 0) + (sg.n_items && sg.n_deep ? 1 : 0);
	}
	return k;
}

extern "C" void acm_gpu_plan_routes(const acm_gpu_plan *p, uint64_t *out4)
{
	uint64_t deep = 0;
	for (const Segment &sg : p->seg)
		deep += sg.n_deep;
	out4[0] = p->n_fast;
	out4[1] = p->n_split;
	out4[2] = p->n_generic - deep;
	out4[3] = deep;
}

extern "C" int acm_gpu_plan_gen_groups(const acm_gpu_plan *p)
{
	return p->grp.empty() ? 1 : (int)p->grp.size();
}

extern "C" void acm_gpu_plan_split(const acm_gpu_plan *p, uint64_t *n_fast, uint64_t *n_generic)
{
	if (n_fast)
		*n_fast = p->n_fast + p->n_split;
	if (n_generic)
		*n_generic = p->n_generic;
}

extern "C" float acm_gpu_plan_last_ms(acm_gpu_plan *p)
{
	float ms = -1.0f;
	if (!p->timed)
		return ms;
	if (cudaEventSynchronize(p->ev1) != cudaSuccess)
		return -1.0f;
	if (cudaEventElapsedTime(&ms, p->ev0, p->ev1) != cudaSuccess)
		return -1.0f;
	return ms;
}

extern "C" void acm_gpu_plan_destroy(acm_gpu_plan *p) { plan_free(p); }

/* the grid the level-7 / 16-row kernel would get for n streams on a device with `sms` SMs when it may
 * use at most max_ctas of them: out3 = { scan CTAs, decode CTAs, stream slots }.  Pure host logic
 * (no device needed): lets the CPU-only test tier check the co-residency rules. */
extern "C" void acm_gpu_debug_geometry(uint64_t n, int sms, int max_ctas, uint32_t *out3)
{
	fast2_geometry(n, sms, max_ctas, &out3[0], &out3[1], &out3[2], 0);
}

extern "C" void acm_gpu_debug_geometry_walk(uint64_t n, int sms, int max_ctas, uint32_t *out4)
{
	fast2_geometry(n, sms, max_ctas, &out4[0], &out4[1], &out4[2], 1, &out4[3]);
}

/* the plan's 64 in-kernel counters, accumulated over its runs ([32]: blocks re-walked by the generic scan,
 * always counted; the cycle counters of -DF2_PROF tuning builds are zero otherwise) */
extern "C" int acm_gpu_plan_debug_counters(acm_gpu_plan *p, unsigned long long *out64)
{
	if (!p || !p->d_prof)
		return ACM_ERR_OTHER;
	return cudaMemcpy(out64, p->d_prof, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess
		       ? ACM_OK
		       : ACM_ERR_OTHER;
}

/* ------------------------------------------------------------------ one-shot */

namespace {

/* device staging buffers and streams of the host path, kept between calls (one set per device) */
constexpr unsigned MAX_SEG = 8;

struct Workspace {
	uint8_t *d_blob = nullptr, *d_out = nullptr;
	size_t blob_cap = 0, out_cap = 0;
	cudaStream_t s_in = nullptr, s_out = nullptr;
	cudaStream_t s_k[MAX_SEG] = {};
	DevArena arena; /* the plans' descriptor tables, rings, history */
};
/* one workspace and one lock per device: callers on different GPUs never wait for each other;
 * two callers on the same GPU take turns (they would share its copy engines anyway) */
constexpr int MAX_DEV = 64;
std::mutex g_ws_mutex[MAX_DEV];
Workspace g_ws[MAX_DEV];

int ws_reserve(Workspace &w, size_t blob_bytes, size_t out_bytes)
{
	if (!w.s_in) {
		CU(cudaStreamCreateWithFlags(&w.s_in, cudaStreamNonBlocking));
		CU(cudaStreamCreateWithFlags(&w.s_out, cudaStreamNonBlocking));
		for (unsigned k = 0; k < MAX_SEG; k++)
			CU(cudaStreamCreateWithFlags(&w.s_k[k], cudaStreamNonBlocking));
	}
	if (w.blob_cap < blob_bytes) {
		cudaFree(w.d_blob);
		w.d_blob = nullptr;
		w.blob_cap = 0;
		CU(cudaMalloc(&w.d_blob, blob_bytes));
		w.blob_cap = blob_bytes;
	}
	if (w.out_cap < out_bytes) {
		cudaFree(w.d_out);
		w.d_out = nullptr;
		w.out_cap = 0;
		CU(cudaMalloc(&w.d_out, out_bytes));
		w.out_cap = out_bytes;
	}
	return ACM_OK;
fail:
	return ACM_ERR_OTHER;
}

} // namespace

extern "C" void acm_gpu_release_workspace(void)
{
	int cur = 0;
	cudaGetDevice(&cur);
	acm_stream_release_pool();
	for (int d = 0; d < MAX_DEV; d++) {
		std::lock_guard<std::mutex> lock(g_ws_mutex[d]);
		Workspace &w = g_ws[d];
		if (!w.s_in && !w.d_blob && !w.d_out && !w.arena.base)
			continue;
		cudaSetDevice(d);
		cudaFree(w.d_blob);
		cudaFree(w.d_out);
		cudaFree(w.arena.base);
		if (w.s_in) {
			cudaStreamDestroy(w.s_in);
			cudaStreamDestroy(w.s_out);
			for (unsigned k = 0; k < MAX_SEG; k++)
				if (w.s_k[k])
					cudaStreamDestroy(w.s_k[k]);
		}
		w = Workspace();
	}
	cudaSetDevice(cur);
}

#define CUR(call)                                                                     \
	do {                                                                          \
		cudaError_t e_ = (call);                                              \
		if (e_ != cudaSuccess) {                                              \
			acm_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call,     \
				      cudaGetErrorString(e_));                        \
			return ACM_ERR_OTHER;                                         \
		}                                                                     \
	} while (0)

/* copy-in / decode / copy-out of every segment, overlapped on the workspace's three streams */
static int run_segments(acm_gpu_plan *plan, const acm_gpu_batch *b, Workspace &w,
			std::vector<cudaEvent_t> &ev_in, std::vector<cudaEvent_t> &ev_k)
{
	const uint8_t *blob_dev = (const uint8_t *)b->blob;
	uint8_t *out_dev = (uint8_t *)b->out;
	const size_t ns = plan->seg.size();
	const auto t_enter = std::chrono::steady_clock::now();
	if (ws_reserve(w, b->blob_on_device ? 0 : ((b->blob_len + 15u) & ~(uint64_t)15u) + 64,
		       b->out_on_device ? 0 : b->out_len + 64) < 0)
		return ACM_ERR_OTHER;
	if (!b->blob_on_device)
		blob_dev = w.d_blob;
	if (!b->out_on_device)
		out_dev = w.d_out;
	if (((uintptr_t)blob_dev & 15u) || ((uintptr_t)out_dev & 15u)) {
		acm_set_error("acm_gpu_decode_batch: device blob and out must be 16-byte aligned");
		return ACM_ERR_OTHER;
	}
	ev_in.assign(ns, nullptr);
	ev_k.assign(ns, nullptr);
	for (size_t g = 0; g < ns; g++) {
		CUR(cudaEventCreateWithFlags(&ev_in[g], cudaEventDisableTiming));
		CUR(cudaEventCreateWithFlags(&ev_k[g], cudaEventDisableTiming));
	}
	/* the cursors are zeroed before anything else is queued: the copy-in stream waits for it */
	CUR(cudaMemsetAsync(plan->d_counters, 0, counter_words(ns) * sizeof(uint32_t), w.s_in));
	if (plan->d_ctl)
		CUR(cudaMemsetAsync(plan->d_ctl, 0, plan->ctl_bytes, w.s_in));
	if (plan->n_generic) {
		CUR(cudaMemsetAsync(plan->g2.nscan, 0, plan->g2_state_bytes / 2, w.s_in));
		CUR(cudaMemsetAsync(plan->g2.first_bad, 0xFF, plan->g2_state_bytes / 2, w.s_in));
		if (plan->fmt.checksums)
			CUR(cudaMemsetAsync(plan->g2.cks_blk, 0, plan->g2_cks_bytes, w.s_in));
	}
	if (plan->n_split) {
		CUR(cudaMemsetAsync(plan->sp.nscan, 0, plan->sp_state_bytes / 2, w.s_in));
		CUR(cudaMemsetAsync(plan->sp.first_bad, 0xFF, plan->sp_state_bytes / 2, w.s_in));
	}
	for (size_t g = 0; g < ns; g++) {
		const Segment &sg = plan->seg[g];
		cudaStream_t sk = w.s_k[g % MAX_SEG];
		if (!b->blob_on_device && sg.blob_hi > sg.blob_lo) {
			uint64_t hi = sg.blob_hi < b->blob_len ? sg.blob_hi : b->blob_len;
			CUR(cudaMemcpyAsync(w.d_blob + sg.blob_lo, (const uint8_t *)b->blob + sg.blob_lo,
					    hi - sg.blob_lo, cudaMemcpyHostToDevice, w.s_in));
		}
		CUR(cudaEventRecord(ev_in[g], w.s_in));
		CUR(cudaStreamWaitEvent(sk, ev_in[g], 0));
		/* the PCM staging buffer is reused between calls: where the kernels will not write every
		 * byte of the segment's range (streams refused on the host, no tail padding, gaps in a
		 * caller-chosen layout) it is cleared first, so that no stale byte reaches the caller */
		if (!b->out_on_device && sg.out_hi > sg.out_lo && sg.sparse)
			CUR(cudaMemsetAsync(w.d_out + sg.out_lo, 0, sg.out_hi - sg.out_lo, sk));
		if (plan_run_segment(plan, g, blob_dev, out_dev, sk) < 0)
			return ACM_ERR_OTHER;
		CUR(cudaEventRecord(ev_k[g], sk));
		if (!b->out_on_device && sg.out_hi > sg.out_lo) {
			uint64_t hi = sg.out_hi < b->out_len ? sg.out_hi : b->out_len;
			CUR(cudaStreamWaitEvent(w.s_out, ev_k[g], 0));
			CUR(cudaMemcpyAsync((uint8_t *)b->out + sg.out_lo, w.d_out + sg.out_lo, hi - sg.out_lo,
					    cudaMemcpyDeviceToHost, w.s_out));
		}
	}
	const bool trace = getenv("ACM_B200_TRACE") != nullptr;
	if (trace)
		fprintf(stderr, "[acm trace] queued %.2f ms after run_segments was entered\n",
			std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_enter).count());
	const auto t_q = std::chrono::steady_clock::now();
	auto since = [&t_q]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_q).count(); };
	for (size_t g = 0; g < ns; g++) {
		CUR(cudaEventSynchronize(ev_k[g]));
		if (trace)
			fprintf(stderr, "[acm trace] segment %zu decoded %.2f ms after the last launch was queued\n", g, since());
	}
	if (acm_gpu_plan_fetch(plan, b->streams, w.s_in) < 0)
		return ACM_ERR_OTHER;
	if (trace)
		fprintf(stderr, "[acm trace] results fetched %.2f ms\n", since());
	CUR(cudaStreamSynchronize(w.s_out));
	if (trace)
		fprintf(stderr, "[acm trace] last copy-out done %.2f ms\n", since());
	CUR(cudaGetLastError());
	return ACM_OK;
}

static int decode_batch_one_device(const acm_gpu_batch *b, const acm_gpu_opts *opts)
{
	acm_gpu_plan *plan = nullptr;
	std::vector<cudaEvent_t> ev_in, ev_k;
	int err = ACM_ERR_OTHER, perr = 0, dev = 0;
	bool need_probe = false, host_io, monotonic = true;
	unsigned nseg = 1;
	const auto t_call = std::chrono::steady_clock::now();

	if (use_device(opts) < 0)
		return ACM_ERR_OTHER;
	for (uint64_t i = 0; i < b->n; i++)
		if (b->streams[i].rows == 0 && b->streams[i].total_values == 0 && b->streams[i].status == 0)
			need_probe = true; /* never probed (a rejected header keeps its ACM_ERR_NOT_ACM) */
	if (need_probe &&
	    acm_gpu_probe(b->blob, b->blob_len, b->blob_on_device, b->streams, b->n, opts) < 0)
		return ACM_ERR_OTHER;
	for (uint64_t i = 0; i < b->n; i++) {
		const acm_gpu_stream &g = b->streams[i];
		/* a stream's slot ends at the next 16-byte boundary: the kernels zero-fill up to there, so
		 * that is what `out` has to hold (acm_gpu_layout's figure) */
		uint64_t need = g.out_off + (((uint64_t)g.total_values * (uint64_t)opts->wordlen + 15u) & ~(uint64_t)15u);
		if (acm_stream_accepted(&g) && (need > b->out_len || g.in_off + g.in_len > b->blob_len)) {
			acm_set_error("stream %llu does not fit its blob/out range", (unsigned long long)i);
			return ACM_ERR_OTHER;
		}
		if (i && (g.in_off < b->streams[i - 1].in_off || g.out_off < b->streams[i - 1].out_off))
			monotonic = false;
	}
	host_io = !b->blob_on_device || !b->out_on_device;
	/* host buffers: cut the batch into segments and pipeline copy-in / decode / copy-out, provided the
	 * caller's order is also the byte order of blob and out (acm_gpu_layout's order) */
	if (host_io && monotonic && b->out_len >= ((uint64_t)32 << 20) && b->n >= 64)
		nseg = MAX_SEG;
	if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) {
		acm_set_error("cudaGetDevice failed (or device ordinal >= %d)", MAX_DEV);
		return ACM_ERR_OTHER;
	}
	{
		std::lock_guard<std::mutex> lock(g_ws_mutex[dev]);
		Workspace &w = g_ws[dev];
		const auto t0 = std::chrono::steady_clock::now();
		plan = plan_create(b->streams, b->n, opts, nseg, &w.arena, &perr);
		if (!plan)
			return perr ? perr : ACM_ERR_OTHER;
		if (getenv("ACM_B200_TRACE"))
			fprintf(stderr, "[acm trace] plan for %llu streams in %u segments built in %.2f ms\n", (unsigned long long)b->n, nseg,
				std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
		err = run_segments(plan, b, w, ev_in, ev_k);
		if (w.s_in) {
			cudaStreamSynchronize(w.s_in);
			cudaStreamSynchronize(w.s_out);
			for (unsigned k = 0; k < MAX_SEG; k++)
				cudaStreamSynchronize(w.s_k[k]);
		}
	}
	for (cudaEvent_t e : ev_in)
		if (e)
			cudaEventDestroy(e);
	for (cudaEvent_t e : ev_k)
		if (e)
			cudaEventDestroy(e);
	acm_gpu_plan_destroy(plan);
	if (getenv("ACM_B200_TRACE"))
		fprintf(stderr, "[acm trace] call done %.2f ms after it was entered\n",
			std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count());
	return err;
}

/*
 * Streams are independent, so a batch shards over the GPUs of a box by
 * stream, with no collective: opts->device_mask names the CUDA ordinals, the caller's stream
 * array is cut into one contiguous range per device with (nearly) equal shares of the output
 * words, and every range is decoded by its own host thread through the single-device path
 * (own workspace, own lock, own CUDA streams).  Host buffers only: a device pointer belongs to
 * one GPU.  The per-stream results land in the caller's array as usual.
 */
extern "C" int acm_gpu_decode_batch(const acm_gpu_batch *b, const acm_gpu_opts *opts)
{
	acm_gpu_opts defaults;
	g_err[0] = 0;
	if (!b || (!b->streams && b->n)) {
		acm_set_error("acm_gpu_decode_batch: null batch");
		return ACM_ERR_OTHER;
	}
	if (!opts) {
		acm_gpu_opts_init(&defaults);
		opts = &defaults;
	}
	std::vector<int> devs;
	for (int d = 0; d < 32; d++)
		if (opts->device_mask & (1u << d))
			devs.push_back(d);
	if (devs.size() <= 1) {
		if (devs.size() == 1) {
			acm_gpu_opts o = *opts;
			o.device = devs[0];
			o.device_mask = 0;
			return decode_batch_one_device(b, &o);
		}
		return decode_batch_one_device(b, opts);
	}
	if (b->blob_on_device || b->out_on_device) {
		acm_set_error("acm_gpu_decode_batch: device_mask with more than one GPU needs host buffers");
		return ACM_ERR_OTHER;
	}
	int have = 0;
	if (cudaGetDeviceCount(&have) != cudaSuccess || devs.back() >= have) {
		acm_set_error("acm_gpu_decode_batch: device_mask names GPU %d, the box has %d", devs.back(), have);
		return ACM_ERR_OTHER;
	}
	{
		/* headers first (host work, once), so that the split can weigh the streams */
		bool need_probe = false;
		for (uint64_t i = 0; i < b->n; i++)
			if (b->streams[i].rows == 0 && b->streams[i].total_values == 0 && b->streams[i].status == 0)
				need_probe = true;
		if (need_probe && acm_gpu_probe(b->blob, b->blob_len, 0, b->streams, b->n, opts) < 0)
			return ACM_ERR_OTHER;
	}
	const size_t nd = devs.size();
	std::vector<uint64_t> cut(nd + 1, 0);
	{
		uint64_t total = 0, acc = 0, i = 0;
		for (uint64_t k = 0; k < b->n; k++)
			total += b->streams[k].total_values;
		for (size_t d = 0; d < nd; d++) {
			const uint64_t target = total / nd * (d + 1);
			while (i < b->n && (d + 1 == nd || acc < target))
				acc += b->streams[i++].total_values;
			cut[d + 1] = i;
		}
		cut[nd] = b->n;
	}
	std::vector<int> rc(nd, ACM_OK);
	std::vector<std::string> msg(nd);
	std::vector<std::thread> th;
	for (size_t d = 0; d < nd; d++) {
		th.emplace_back([&, d] {
			acm_gpu_batch sb = *b;
			acm_gpu_opts o = *opts;
			o.device = devs[d];
			o.device_mask = 0;
			sb.streams = b->streams + cut[d];
			sb.n = cut[d + 1] - cut[d];
			if (sb.n)
				rc[d] = decode_batch_one_device(&sb, &o);
			msg[d] = g_err; /* the error text is per thread */
		});
	}
	for (std::thread &t : th)
		t.join();
	for (size_t d = 0; d < nd; d++)
		if (rc[d] != ACM_OK) {
			acm_set_error("GPU %d: %s", devs[d], msg[d].c_str());
			return rc[d];
		}
	return ACM_OK;
}
