/*
 * oracle/ref_driver.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin driver that is compiled TOGETHER with the unmodified reference sources
 * (/root/reference/src/decode.c + util.c, where they lie) into
 * oracle/_ref/libacm_ref.so by oracle/Makefile.  It contains no decoding logic of
 * its own: it feeds an in-memory file image to the reference's public API
 * (acm_open_decoder / acm_read, libacm.h:120-136) through memory callbacks and
 * records exactly what the reference returns.
 *
 * Result contract (the same one acm_gpu_decode_batch and oracle/acm_oracle.c use):
 *   out[0 .. words_out*2)  PCM words in the requested 16-bit format
 *   status                 what the LAST acm_read call returned when the loop ended:
 *                          0 (EOF / all total_values delivered) or ACM_ERR_* (<0)
 *   words_out              acm->stream_pos at that moment
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "libacm.h" /* the reference's own header, via -I/root/reference/src */

typedef struct {
	const unsigned char *data;
	size_t len, pos;
	int chunk; /* >0: cap every read at this many bytes (short-read tests) */
	int seekable;
	int n_seek, n_close, n_read;
} mem_src;

static int mem_read(void *ptr, int size, int n, void *arg)
{
	mem_src *m = arg;
	size_t want = (size_t)size * (size_t)n, left = m->len - m->pos;
	m->n_read++;
	if (m->chunk > 0 && want > (size_t)m->chunk)
		want = m->chunk;
	if (want > left)
		want = left;
	want -= want % (size_t)size;
	memcpy(ptr, m->data + m->pos, want);
	m->pos += want;
	return (int)(want / (size_t)size);
}

static int mem_seek(void *arg, int off, int whence)
{
	mem_src *m = arg;
	m->n_seek++;
	if (!m->seekable || whence != SEEK_SET || off < 0 || (size_t)off > m->len)
		return -1;
	m->pos = off;
	return 0;
}

static int mem_close(void *arg)
{
	mem_src *m = arg;
	m->n_close++;
	return 0;
}

static int mem_len(void *arg)
{
	mem_src *m = arg;
	return (int)m->len;
}

typedef struct {
	unsigned channels, rate, acm_channels, acm_level, acm_cols, acm_rows;
	unsigned total_values, block_len, wavc;
} ref_info;

/* open + decode everything; returns the acm_open_decoder code (0 or <0). */
int ref_decode(const unsigned char *file, size_t len, int force_chans,
	       int bigendianp, int sgned, unsigned char *out, size_t out_cap,
	       unsigned *words_out, int *status, ref_info *info)
{
	mem_src src;
	acm_io_callbacks io;
	ACMStream *acm = NULL;
	int err, res = 0;
	size_t done = 0;

	memset(&src, 0, sizeof(src));
	src.data = file;
	src.len = len;
	src.seekable = 1;
	memset(&io, 0, sizeof(io));
	io.read_func = mem_read;
	io.seek_func = mem_seek;
	io.close_func = mem_close;
	io.get_length_func = mem_len;

	*words_out = 0;
	*status = 0;
	err = acm_open_decoder(&acm, &src, io, force_chans);
	if (err < 0)
		return err;
	if (info) {
		info->channels = acm->info.channels;
		info->rate = acm->info.rate;
		info->acm_channels = acm->info.acm_channels;
		info->acm_level = acm->info.acm_level;
		info->acm_cols = acm->info.acm_cols;
		info->acm_rows = acm->info.acm_rows;
		info->total_values = acm->total_values;
		info->block_len = acm->block_len;
		info->wavc = acm->wavc_file;
	}
	/* same loop shape as acm_read_loop (util.c:258-277) but keeping the raw code */
	while (done < out_cap) {
		size_t want = out_cap - done;
		if (want > 0x40000000u)
			want = 0x40000000u;
		res = acm_read(acm, out + done, (unsigned)want, bigendianp, 2, sgned);
		if (res <= 0)
			break;
		done += res;
	}
	if (res > 0)
		res = 0; /* buffer full: caller gave exactly total*2 bytes */
	*status = res;
	*words_out = acm->stream_pos;
	acm_close(acm);
	return 0;
}

/* decode-only timing loop for the CPU baseline: `reps` full decodes of one image,
 * the way `acmtool -d -n` does it (16 KiB buffer, 8 KiB requests: acmtool.c:269-285). */
double ref_time_decode(const unsigned char *file, size_t len, int reps, unsigned long long *words)
{
	struct timespec t0, t1;
	unsigned long long tot = 0;
	int r;
	static __thread unsigned char buf[16 * 1024];

	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (r = 0; r < reps; r++) {
		mem_src src;
		acm_io_callbacks io;
		ACMStream *acm = NULL;
		memset(&src, 0, sizeof(src));
		src.data = file;
		src.len = len;
		memset(&io, 0, sizeof(io));
		io.read_func = mem_read;
		io.get_length_func = mem_len;
		if (acm_open_decoder(&acm, &src, io, 0) < 0)
			return -1.0;
		for (;;) {
			int res = acm_read_loop(acm, buf, sizeof(buf) / 2, 0, 2, 1);
			if (res <= 0)
				break;
		}
		tot += acm->stream_pos;
		acm_close(acm);
	}
	clock_gettime(CLOCK_MONOTONIC, &t1);
	*words = tot;
	return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* ---- handle-style access for the API-parity tests (streaming + seek) ---- */

typedef struct {
	mem_src src;
	ACMStream *acm;
} ref_handle;

ref_handle *ref_open(const unsigned char *file, size_t len, int force_chans,
		     int seekable, int with_seek_func, int chunk, int *err,
		     int *closed_on_fail)
{
	ref_handle *h = calloc(1, sizeof(*h));
	acm_io_callbacks io;
	h->src.data = file;
	h->src.len = len;
	h->src.seekable = seekable;
	h->src.chunk = chunk;
	memset(&io, 0, sizeof(io));
	io.read_func = mem_read;
	if (with_seek_func)
		io.seek_func = mem_seek;
	io.close_func = mem_close;
	if (seekable)
		io.get_length_func = mem_len;
	*err = acm_open_decoder(&h->acm, &h->src, io, force_chans);
	if (*err < 0) {
		*closed_on_fail = h->src.n_close; /* Q10: must stay 0 (decode.c:817-823) */
		free(h);
		return NULL;
	}
	return h;
}

ACMStream *ref_stream(ref_handle *h) { return h->acm; }
int ref_n_seek(ref_handle *h) { return h->src.n_seek; }
int ref_n_close(ref_handle *h) { return h->src.n_close; }

void ref_close(ref_handle *h)
{
	acm_close(h->acm);
	free(h);
}
