/*
 * acm_gpu.h -- batched ACM decode on a B200 (the entry point the reference does
 * not have).  Plain C ABI: pointers and sizes only.
 *
 * One call decodes thousands to millions of INDEPENDENT ACM streams.  It
 * replaces, for every stream at once, the loop a reference caller writes around
 *   acm_open_decoder (decode.c:758) -> acm_read_loop (util.c:258) -> acm_close
 * and produces, per stream, exactly what that loop produces:
 *   PCM      the words acm_read delivered (decode.c:826-876), in the format
 *            (bigendianp, wordlen, sgned) of output_values (decode.c:657-677);
 *            when pad_tail is set the rest of the stream's total_values words
 *            are zero bytes, which is byte-for-byte `acmtool -d -r` (acmtool.c:293-310)
 *   status   what the LAST acm_read returned: 0, ACM_ERR_CORRUPT (-6) or
 *            ACM_ERR_UNEXPECTED_EOF (-7); ACM_ERR_NOT_ACM (-3) if the header is
 *            rejected (decode.c:712-752); ACM_ERR_OTHER (-1) if the image is too
 *            large for the device decoder (>= 512 MiB)
 *   words    acm->stream_pos at that moment
 *
 * Streams are "file images": the bytes of a .acm / WAVC file, header included,
 * concatenated in one blob (any alignment; 16-byte aligned images are fastest).
 */
#ifndef ACM_GPU_H
#define ACM_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACM_GPU_ABI_VERSION 2

typedef struct acm_gpu_stream {
	/* ---- caller fills */
	uint64_t in_off;       /* byte offset of the file image inside the blob */
	uint32_t in_len;       /* image length in bytes */
	uint32_t reserved0;
	uint64_t out_off;      /* byte offset of this stream's PCM inside `out`; multiple of 16
				  (acm_gpu_layout assigns it) */
	/* ---- acm_gpu_probe fills (header fields, read_header decode.c:712-752) */
	uint32_t total_values; /* PCM words, all channels */
	uint32_t channels;     /* effective, after force_chans (decode.c:795-799) */
	uint32_t acm_channels; /* as written in the header */
	uint32_t rate;
	uint32_t level;
	uint32_t rows;
	uint32_t wavc;         /* 1 if the 28-byte WAVC pre-header is present */
	/* ---- decode fills */
	int32_t status;
	uint32_t words;
	uint32_t reserved1;
	uint64_t checksum;     /* sum over emitted words i of (i+1)*(u_i+1) mod 2^64, u_i the word as
				  an unsigned wordlen-byte integer; 0 unless want_checksums */
} acm_gpu_stream;

typedef struct acm_gpu_opts {
	int32_t device;         /* CUDA ordinal; -1 = the calling thread's current device */
	int32_t bigendianp;     /* as acm_read: 0 little endian, 1 big endian */
	int32_t wordlen;        /* 2 (reference), 3 or 4 (extension) */
	int32_t sgned;          /* 1 signed, 0 unsigned (adds the sign bit, decode.c:640) */
	int32_t force_chans;    /* as acm_open_decoder */
	int32_t want_checksums; /* compute acm_gpu_stream.checksum on the device */
	int32_t pad_tail;       /* zero-fill words [words, total_values) of every stream */
	int32_t kernel;         /* 0 auto; 1 force the generic kernel (testing) */
	uint32_t device_mask;   /* acm_gpu_decode_batch only: bit d = use CUDA device d.  More than one
				   bit shards the batch by stream over those GPUs (contiguous ranges of the
				   stream array with equal shares of the output, one host thread per GPU, no
				   collective: the streams are independent); host buffers only.  0 = `device` */
	int32_t reserved[7];
} acm_gpu_opts;

typedef struct acm_gpu_batch {
	const void *blob;      /* concatenated file images */
	uint64_t blob_len;
	int32_t blob_on_device; /* 0: host memory, 1: device memory of opts->device */
	int32_t out_on_device;
	void *out;             /* PCM destination */
	uint64_t out_len;
	acm_gpu_stream *streams; /* host array, n entries */
	uint64_t n;
} acm_gpu_batch;

/* defaults: current device, s16le signed, trust header, no checksums, pad_tail=1 */
void acm_gpu_opts_init(acm_gpu_opts *o);

/*
 * Parse the n headers (host or device blob) and fill the header fields of
 * every stream; a rejected header gets status = ACM_ERR_NOT_ACM and
 * total_values = 0.  Returns the number of accepted streams or ACM_ERR_*.
 */
int64_t acm_gpu_probe(const void *blob, uint64_t blob_len, int blob_on_device,
		      acm_gpu_stream *streams, uint64_t n, const acm_gpu_opts *opts);

/* Assign out_off back to back (16-byte aligned); returns the bytes `out` must hold. */
uint64_t acm_gpu_layout(acm_gpu_stream *streams, uint64_t n, int wordlen);

/*
 * One-shot decode.  Host blobs/outputs are copied inside the call: the batch is cut
 * into segments whose copy-in, decode and copy-out overlap on three CUDA streams (pass
 * pinned host memory for full PCIe speed; the streams must be listed in blob/out byte
 * order, which is what acm_gpu_layout produces).  Device pointers are used in place.
 * Streams must have been probed.  Returns ACM_OK, or ACM_ERR_OTHER for a CUDA/runtime
 * failure (per-stream problems are reported in streams[i].status).
 */
int acm_gpu_decode_batch(const acm_gpu_batch *batch, const acm_gpu_opts *opts);

/*
 * Resident path: build the device-side descriptor tables once, then launch the
 * decode kernels any number of times on device-resident blob/out pointers.
 */
typedef struct acm_gpu_plan acm_gpu_plan;

acm_gpu_plan *acm_gpu_plan_create(const acm_gpu_stream *streams, uint64_t n,
				  const acm_gpu_opts *opts, int *err);
/* asynchronous on `cuda_stream` (a cudaStream_t, NULL = default stream).  d_out must hold the
 * acm_gpu_layout() byte count: every stream's slot ends on a 16-byte boundary and is written
 * (PCM, then zeros) up to there */
int acm_gpu_plan_run(acm_gpu_plan *plan, const void *d_blob, void *d_out, void *cuda_stream);
/* waits for cuda_stream and copies status / words / checksum into streams[] */
int acm_gpu_plan_fetch(acm_gpu_plan *plan, acm_gpu_stream *streams, void *cuda_stream);
/* kernels launched by one acm_gpu_plan_run */
int acm_gpu_plan_launches(const acm_gpu_plan *plan);
/* streams routed to the (fast, generic) kernels */
void acm_gpu_plan_split(const acm_gpu_plan *plan, uint64_t *n_fast, uint64_t *n_generic);
/* the same by decode route: out4 = { the fused level-7 / 16-row kernel, the split path (opts.kernel = 2),
 * the general throughput path (any rows, level <= 10: scan -> unpack -> tile lift), the block-at-a-time
 * backstop (level 11..15) } */
void acm_gpu_plan_routes(const acm_gpu_plan *plan, uint64_t *out4);
/* general path of a resident plan: its streams are sorted by expected walk time and cut into this many
 * groups, each scanned / unpacked / transformed on a CUDA stream of its own, so that the short groups are
 * decoded while the long ones are still being walked (1 = one group: small batches, levels above 10;
 * the ACM_B200_GEN_GROUPS environment variable overrides the default of 8, ACM_B200_GEN_GROUP_MIN the
 * batch size of 1024 streams from which a plan is grouped) */
int acm_gpu_plan_gen_groups(const acm_gpu_plan *plan);
/* average device time of the kernels of the last run on that plan, ms (CUDA events on cuda_stream) */
float acm_gpu_plan_last_ms(acm_gpu_plan *plan);
void acm_gpu_plan_destroy(acm_gpu_plan *plan);
/* copies the plan's 64 in-kernel counters, accumulated over its runs.  Always counted: [32] = blocks
 * the level-7 / 16-row kernel had to re-walk with the generic block scan (bad selector, end of a
 * truncated stream; never for a healthy stream).  The cycle counters [0..19] are those of tuning
 * builds (-DF2_PROF, tools/build_variants.py) and stay zero in a normal build */
int acm_gpu_plan_debug_counters(acm_gpu_plan *plan, unsigned long long *out64);
/* grid geometry of the level-7 / 16-row kernel for n streams: out3 = { scan CTAs, decode CTAs, stream
 * slots }; host logic only (tests) */
void acm_gpu_debug_geometry(uint64_t n, int sms, int max_ctas, uint32_t *out3);
/* the same for a launch that is bound by the walk of its longest stream (few, long streams): out4 =
 * { scan CTAs, decode CTAs, stream slots, scan warps in use per scan CTA (all 8, or one per SM
 * sub-partition when the batch is small enough for twice as many scan CTAs) } */
void acm_gpu_debug_geometry_walk(uint64_t n, int sms, int max_ctas, uint32_t *out4);

/*
 * On-GPU corpus generation (test / benchmark input side, not part
 * of the reference's API).  The reference ships neither an encoder nor sample files; its test
 * inputs are synthetic (libacm_b200/csrc/acmgen.c).  This runs the same generator one thread per
 * stream on the device and writes the images into d_blob back-to-back on 16-byte boundaries --
 * byte-identical to the host generator -- so that a million-stream corpus never crosses PCIe.
 *   params: n `acmgen_params` (libacm_b200/csrc/acmgen.h) in host memory; level <= 10
 *   d_blob/cap: device buffer; NULL = sizing call (offs, lens, *used are still filled in)
 *   offs, lens: host arrays of n entries (placement of every image); *used = bytes of d_blob used
 *   device: CUDA device ordinal, -1 = current
 * Returns ACM_OK or ACM_ERR_OTHER (acm_gpu_last_error() has the text).
 */
int acm_gpu_generate(const void *params, uint64_t n, void *d_blob, uint64_t cap, uint64_t *offs,
		     uint32_t *lens, uint64_t *used, int device);

/* acm_gpu_decode_batch keeps its device staging buffers and streams between calls (host-buffer
 * path); this frees them */
void acm_gpu_release_workspace(void);

/* last CUDA / runtime error text of the calling thread ("" if none) */
const char *acm_gpu_last_error(void);
int acm_gpu_abi_version(void);

#ifdef __cplusplus
}
#endif

#endif
