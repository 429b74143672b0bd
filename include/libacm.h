/*
 * libacm - Interplay ACM audio decoder.
 *
 * Copyright (c) 2004-2010, Marko Kreen
 *
 * Permission to use, copy, modify, and/or distribute this software for any
 * purpose with or without fee is hereby granted, provided that the above
 * copyright notice and this permission notice appear in all copies.
 *
 * THE SOFTWARE IS PROVIDED "AS IS" AND THE AUTHOR DISCLAIMS ALL WARRANTIES
 * WITH REGARD TO THIS SOFTWARE INCLUDING ALL IMPLIED WARRANTIES OF
 * MERCHANTABILITY AND FITNESS. IN NO EVENT SHALL THE AUTHOR BE LIABLE FOR
 * ANY SPECIAL, DIRECT, INDIRECT, OR CONSEQUENTIAL DAMAGES OR ANY DAMAGES
 * WHATSOEVER RESULTING FROM LOSS OF USE, DATA OR PROFITS, WHETHER IN AN
 * ACTION OF CONTRACT, NEGLIGENCE OR OTHER TORTIOUS ACTION, ARISING OUT OF
 * OR IN CONNECTION WITH THE USE OR PERFORMANCE OF THIS SOFTWARE.
 */

/*
 * The declarations below reproduce the public interface of libacm 1.3
 * (constants, ACMInfo, acm_io_callbacks, struct ACMStream and the prototypes of
 * reference src/libacm.h:26-170), which is why the ISC notice above is kept.
 */

/*
 * libacm.h -- drop-in C surface of the B200 ACM decoder.
 *
 * Source- and ABI-compatible with markokr/libacm 1.3's public header
 * (reference src/libacm.h:26-170): the same constants, the same PUBLIC
 * `struct ACMStream` field layout (callers such as acmtool.c:52 and
 * plugin-gstreamer.c:357/:675 read its fields directly), the same 20 entry
 * points with the same argument meaning and return conventions.  A program that
 * was built against the reference header links against libacm_b200.so unchanged.
 *
 * What differs is behind the surface: decode_block's three stages (bit unpack +
 * filler dispatch, the juggle transform, PCM packing) run as CUDA kernels on an
 * sm_100a device; there is no CPU decode path, and acm_open_decoder fails with
 * ACM_ERR_OTHER when no CUDA device is usable.  The batched entry point lives in
 * acm_gpu.h.
 *
 * Extension over the reference: acm_read / acm_read_loop also accept wordlen 3
 * and 4 (the reference returns ACM_ERR_BADFMT, decode.c:832-835).
 */
#ifndef __LIBACM_H
#define __LIBACM_H

#ifdef __cplusplus
extern "C" {
#endif

#define LIBACM_VERSION "1.3"
#define LIBACM_B200 1

/* stream signature and native sample size (reference libacm.h:28-29) */
#define ACM_ID   0x032897
#define ACM_WORD 2

/* return codes (reference libacm.h:31-39) */
#define ACM_OK                  0
#define ACM_ERR_OTHER          -1
#define ACM_ERR_OPEN           -2
#define ACM_ERR_NOT_ACM        -3
#define ACM_ERR_READ_ERR       -4
#define ACM_ERR_BADFMT         -5
#define ACM_ERR_CORRUPT        -6
#define ACM_ERR_UNEXPECTED_EOF -7
#define ACM_ERR_NOT_SEEKABLE   -8

/* reference libacm.h:41-50 */
typedef struct ACMInfo {
	unsigned channels;     /* effective channel count (after force_chans) */
	unsigned rate;         /* sample rate, Hz */
	unsigned acm_id;
	unsigned acm_version;
	unsigned acm_channels; /* channel count as written in the header */
	unsigned acm_level;    /* number of transform levels */
	unsigned acm_cols;     /* 1 << acm_level */
	unsigned acm_rows;
} ACMInfo;

/*
 * I/O callbacks, passed BY VALUE to acm_open_decoder (reference libacm.h:52-69).
 *   read_func        fread-like: returns items read, 0 at EOF, <0 on error
 *   seek_func        optional; only ever asked for (header_len, SEEK_SET)
 *   close_func       optional; called by acm_close, NOT on a failed open
 *   get_length_func  optional; >0 makes acm_seekable() true
 */
typedef struct {
	int (*read_func)(void *ptr, int size, int n, void *datasrc);
	int (*seek_func)(void *datasrc, int offset, int whence);
	int (*close_func)(void *datasrc);
	int (*get_length_func)(void *datasrc);
} acm_io_callbacks;

/*
 * Public stream state; field order and types match reference libacm.h:71-100 so
 * that offsets agree.  Fields the reference used for its CPU working set
 * (buf/block/wrapbuf/ampbuf/midbuf) are kept for layout and left NULL: the
 * working set lives in device memory behind `gpu`.
 */
struct ACMStream {
	ACMInfo info;
	unsigned total_values; /* PCM words in the stream, all channels */

	void *io_arg;
	acm_io_callbacks io;
	unsigned data_len;

	unsigned char *buf;
	unsigned buf_max, buf_size, buf_pos, bit_avail;
	unsigned bit_data;
	unsigned buf_start_ofs;

	unsigned block_len;    /* words per block: acm_rows * acm_cols */
	unsigned wrapbuf_len;  /* 2 * acm_cols - 2 */
	int *block;
	int *wrapbuf;
	int *ampbuf;
	int *midbuf;

	unsigned block_ready:1;
	unsigned file_eof:1;
	unsigned wavc_file:1;
	unsigned stream_pos;   /* words delivered so far (absolute) */
	unsigned block_pos;    /* words delivered from the current block */

	void *gpu;             /* private: device-side decoder state (appended) */
};
typedef struct ACMStream ACMStream;

/*
 * force_chans (reference libacm.h:105-118, decode.c:795-799):
 *   > 0  use that channel count;  0  trust the header;
 *   -1   plain ACM files are taken as stereo, WAVC files keep their header.
 * It changes only bookkeeping (frame rounding, pcm totals) -- never sample values.
 */
int acm_open_decoder(ACMStream **res, void *io_arg, acm_io_callbacks io, int force_chans);
int acm_open_file(ACMStream **acm, const char *filename, int force_chans);
void acm_close(ACMStream *acm);

/*
 * Returns bytes written (>0), 0 at end of stream, or ACM_ERR_* (<0).  Never
 * crosses a block boundary, clips to total_values, rounds down to whole frames
 * (reference decode.c:826-876).  buf == NULL decodes and discards.
 */
int acm_read(ACMStream *acm, void *buf, unsigned nbytes,
	     int bigendianp, int wordlen, int sgned);
/* acm_read until `len` bytes are filled or the stream ends (reference util.c:258-277) */
int acm_read_loop(ACMStream *acm, void *dst, unsigned len,
		  int bigendianp, int wordlen, int sgned);

/* returns the new position or ACM_ERR_NOT_SEEKABLE (reference util.c:206-253) */
int acm_seek_pcm(ACMStream *acm, unsigned pcm_pos);
int acm_seek_time(ACMStream *acm, unsigned pos_ms);

const ACMInfo *acm_info(ACMStream *acm);
int acm_seekable(ACMStream *acm);
unsigned acm_bitrate(ACMStream *acm);
unsigned acm_rate(ACMStream *acm);
unsigned acm_channels(ACMStream *acm);
unsigned acm_raw_total(ACMStream *acm);
unsigned acm_raw_tell(ACMStream *acm);
unsigned acm_pcm_total(ACMStream *acm);
unsigned acm_pcm_tell(ACMStream *acm);
unsigned acm_time_total(ACMStream *acm);
unsigned acm_time_tell(ACMStream *acm);
const char *acm_strerror(int err);

#ifdef __cplusplus
}
#endif

#endif
