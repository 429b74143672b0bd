/*
 * acm_kernels.cuh -- launch interface between the host glue (acm_batch.cu) and the
 * decode kernels (acm_kernels.cu, acm_fast.cu).
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "acm_device.cuh"

namespace acm {

struct KernelArgs {
	const uint8_t *blob;      /* 16-byte aligned */
	uint64_t blob_room;       /* readable bytes at blob: blob_len rounded up to 16 */
	uint8_t *out;
	const DevStream *streams; /* this kernel's slice of the descriptor table */
	uint32_t count;
	int32_t *status;          /* result arrays, indexed by DevStream::index */
	uint32_t *words;
	unsigned long long *cks;
	const acm_tables *tables; /* device copy */
	uint32_t *counter;        /* work-queue cursor (zeroed before launch) */
	uint32_t *errflag;        /* set non-zero on an internal failure (e.g. copy timeout) */
	uint32_t *hist;           /* fast kernel: per-slot transform history (256 words each) */
	uint8_t *ring;            /* fast kernel: per-slot ring of block records */
	uint8_t *slotctl;         /* fast kernel: per-slot control words (zeroed before launch) */
	uint32_t *scan_done;      /* fast kernel: [0] scan warps that have finished, [1] heartbeat (zeroed before launch) */
	uint32_t n_scan;          /* fast kernel: scan CTAs (the first n_scan of the grid) */
	uint32_t n_slots;         /* fast kernel: stream slots in use */
	uint32_t scan_warps;      /* fast kernel: scan warps in use per scan CTA (all of them, or one per SM sub-partition) */
	unsigned long long *prof; /* 64 counters, only written by -DF2_PROF tuning builds */
	/* generic kernel, resumable decode (acm_stream.cu): when resume_hist != NULL stream i of the
	 * slice starts from / leaves its per-stage history at resume_hist + i * resume_stride (2*cols
	 * words); a stream whose DevStream::resume is 0 still starts from zero history.
	 * end_pos[2i] receives the bit position after the last block that decoded, end_pos[2i+1]
	 * the number of blocks that decoded. */
	uint32_t *resume_hist;
	uint32_t resume_stride;
	uint32_t *end_pos;
	Format fmt;
};

struct GenericScratch {
	uint32_t *buf;      /* n_ctas * stride words */
	size_t stride;      /* words per CTA */
	uint32_t max_blen;  /* largest rows*cols among the streams */
	uint32_t max_cols;
};

/* words of scratch one CTA of the generic kernel needs */
inline size_t generic_scratch_words(uint32_t max_blen, uint32_t max_cols)
{
	return 2 * (size_t)max_blen + 4 * (size_t)max_cols + 64; /* two block buffers, history, two offset buffers */
}

cudaError_t launch_generic(const KernelArgs &a, const GenericScratch &s, int n_ctas,
			   cudaStream_t st);
int generic_ctas_per_sm();

/* ---- the general path as scan -> block records -> block-parallel decode -> finalise (acm_gen2.cu) */

struct BlockRec {
	uint32_t P, end;  /* bit position of the block header / just past the block (or where the scan stopped) */
	int32_t val;      /* block multiplier (decode.c:589) */
	int32_t status;   /* SCAN_OK / SCAN_EOF / ACM_ERR_* */
	uint32_t ncols;   /* columns whose payload was scanned completely */
	uint32_t pad0, pad1, pad2;
};

struct Gen2Stream {
	uint64_t rec_base;   /* index of the stream's first BlockRec (and per-block checksum) */
	uint64_t coff_base;  /* index of its first column offset: block b's are at coff_base + b * cols */
	uint32_t max_blocks; /* records reserved: min(n_attempt, what the image can hold) */
	uint32_t pad;
	uint64_t word_base;  /* level <= 10: index of the stream's first word in the int16 intermediate */
	uint64_t pad2;
};

struct Gen2Item {
	uint32_t stream; /* index into the kernel's descriptor slice */
	uint32_t b0, nb; /* blocks [b0, b0 + nb) are this item's output */
	uint32_t warm;   /* blocks before b0 decoded only to rebuild the transform history */
};

struct Gen2Args {
	const Gen2Stream *gs;
	BlockRec *rec;
	uint32_t *coff;              /* P of every column selector */
	unsigned long long *cks_blk; /* per-block checksum contributions */
	uint32_t *nscan;             /* per stream: records written by the scan */
	uint32_t *first_bad;         /* per stream: first block with an out-of-range radix code (0xFFFFFFFF: none) */
	const Gen2Item *items;
	uint32_t n_items;
	uint32_t *item_counter;      /* zeroed before launch */
	/* streams of level <= 10: unpack -> int16 intermediate -> tile lift (acm_gen2.cu) */
	int16_t *inter16;            /* quantiser indices in stream order */
	const Gen2Item *tiles;       /* lift work items: stream, b0 = tile number (4096 words each) */
	uint32_t n_tiles;
	uint32_t n_deep;             /* streams of level > 10: acm_blocks_kernel's */
	uint32_t *g3_counters;       /* [0] unpack queue, [1] tile queue, [2] scan queue (zeroed before launch) */
};
constexpr uint32_t GEN3_TILE_WORDS = 4096;
constexpr uint32_t GEN3_MAX_LEVEL = 10;

cudaError_t launch_gen2(const KernelArgs &a, const Gen2Args &g, const GenericScratch &s, int n_ctas, cudaStream_t st);
/* the finalise kernel alone (also the last stage of the split path) */
cudaError_t launch_gen2_finish(const KernelArgs &a, const Gen2Args &g, cudaStream_t st);
int gen2_ctas_per_sm();
/* blocks an image of data_bits bits can hold at most (a block is at least 20 + 5 * cols bits) */
inline uint64_t gen2_max_blocks(uint64_t n_attempt, uint64_t data_bits, uint32_t level)
{
	const uint64_t minb = 20u + 5u * ((uint64_t)1 << level);
	const uint64_t fit = (data_bits + 8u) / minb + 2u;
	return n_attempt < fit ? n_attempt : fit;
}

/* ---- the throughput path for 16-row blocks of 128 columns: walk -> unpack -> lift -> finalise (acm_split.cu).
 * Shares the general path's records (BlockRec: pad0 = stream, pad1 = block number, pad2 = the run's epoch),
 * per-stream tables and work items; items carry no warm-up blocks (a run rebuilds its history from the
 * previous block's last two rows of indices). */
struct SplitArgs : Gen2Args {
	uint16_t *coff16;   /* [block][128]: selector positions relative to the block's P */
	uint8_t *inter;     /* [block][2048]: quantiser indices times 2 as signed bytes, lane-major (lane l = columns l, l+32, l+64, l+96: 16 bytes each) */
	uint16_t *wide;     /* [block][128][16]: int16 indices of the columns flagged in wmask */
	uint32_t *wmask;    /* [block][4]: bit l of word p = column 32 p + l is in `wide` */
	uint64_t n_blocks;  /* records of this launch: [0, n_blocks) */
	uint32_t epoch;     /* marks the records written by this run */
	uint32_t pad;
};
bool split_shape(uint32_t level, uint32_t rows);
size_t split_bytes_per_block();
cudaError_t launch_split(const KernelArgs &a, const SplitArgs &g, int sms, cudaStream_t st);
/* one stream, a range of its blocks (streaming API): see acm_split.cu */
cudaError_t launch_split_range(const KernelArgs &a, const SplitArgs &g, uint32_t b0, uint32_t nb, uint32_t P0,
			       int lift, int sms, cudaStream_t st);

/* level-7 / 16-row kernel (acm_fast2.cu): scan CTAs + decode CTAs; 16-bit output formats only */
bool fast_shape(uint32_t level, uint32_t rows);
size_t fast2_smem_bytes();
/* walk_bound: the batch's longest stream takes about as long to walk as the batch to decode (see fast2_walk_bound) */
void fast2_geometry(uint64_t count, int sms, int max_ctas, uint32_t *n_scan, uint32_t *n_work, uint32_t *n_slots,
		    int walk_bound, uint32_t *scan_warps = nullptr);
int fast2_walk_bound(uint64_t longest_blocks, uint64_t total_blocks, int sms, int max_ctas);
size_t fast2_hist_words_per_slot();
int fast2_scan_warps(); /* scan warps per scan CTA: warps w and w + half of them share an SM sub-partition */
size_t fast2_ring_bytes_per_slot();
size_t fast2_ctl_bytes_per_slot();
cudaError_t launch_fast2(const KernelArgs &a, int n_ctas, cudaStream_t st);

/* gathers the first 48 bytes of every image (header parse on the host) */
cudaError_t launch_gather_headers(const uint8_t *blob, uint64_t blob_len, const uint64_t *in_off,
				  const uint32_t *in_len, uint8_t *dst, uint64_t n, cudaStream_t st);

} // namespace acm
