/*
 * acm_hostlogic.cpp -- CUDA-free host logic of the batched decoder: header parsing,
 * the reference's read-loop bookkeeping, and device descriptor construction.  Kept
 * free of CUDA calls so the CPU-only test suite can link and exercise it.
 */
#include <cstring>

#include "acm_device.cuh"
#include "acm_gpu.h"
#include "acm_host.h"
#include "libacm.h"

/* ------------------------------------------------------------------ header parse */

namespace {

/* LSB-first reader over at most 48 header bytes of an image of `len` bytes, with the
 * reference's EOF rule: the image is followed by one zero byte (decode.c:57-61) */
struct HdrBits {
	const uint8_t *p;
	uint64_t have;  /* bytes of p that are valid */
	uint64_t limit; /* readable bits: 8*len + 8 */
	uint64_t pos;
	int get(unsigned n)
	{
		if (pos + n > limit)
			return -1;
		uint32_t v = 0;
		for (unsigned i = 0; i < n; i++) {
			uint64_t q = pos + i, byte = q >> 3;
			if (byte < have)
				v |= (uint32_t)((p[byte] >> (q & 7)) & 1u) << i;
		}
		pos += n;
		return (int)v;
	}
};

} // namespace

/* read_header + read_wavc_header (decode.c:679-752); every failure is NOT_ACM
 * (decode.c:783-785) */
int acm_parse_header(const uint8_t *p, uint64_t len, int force_chans, acm_header *h)
{
	HdrBits b;
	int t;
	memset(h, 0, sizeof(*h));
	b.p = p;
	b.have = len < 48 ? len : 48;
	b.limit = len * 8 + 8;
	b.pos = 0;
	if ((t = b.get(24)) < 0)
		return ACM_ERR_NOT_ACM;
	if (t == 0x564157) { /* 'WAV' decode.c:685 */
		int w[12];
		if (b.get(8) != 'C')
			return ACM_ERR_NOT_ACM;
		for (int i = 0; i < 12; i++)
			if ((w[i] = b.get(16)) < 0)
				return ACM_ERR_NOT_ACM;
		/* only "V1.0" and the magic 28 are checked (decode.c:699-706) */
		if (w[0] != 0x3156 || w[1] != 0x302E || w[6] != 28)
			return ACM_ERR_NOT_ACM;
		h->wavc = 1;
		if ((t = b.get(24)) < 0)
			return ACM_ERR_NOT_ACM;
	}
	if (t != ACM_ID)
		return ACM_ERR_NOT_ACM;
	if (b.get(8) != 1)
		return ACM_ERR_NOT_ACM;
	int lo = b.get(16), hi = b.get(16);
	if (lo < 0 || hi < 0)
		return ACM_ERR_NOT_ACM;
	h->total_values = (uint32_t)lo + ((uint32_t)hi << 16);
	if (h->total_values == 0)
		return ACM_ERR_NOT_ACM;
	t = b.get(16);
	if (t < 1 || t > 2)
		return ACM_ERR_NOT_ACM;
	h->acm_channels = h->channels = (uint32_t)t;
	t = b.get(16);
	if (t < 4096)
		return ACM_ERR_NOT_ACM;
	h->rate = (uint32_t)t;
	if ((t = b.get(4)) < 0)
		return ACM_ERR_NOT_ACM;
	h->level = (uint32_t)t;
	if ((t = b.get(12)) <= 0)
		return ACM_ERR_NOT_ACM;
	h->rows = (uint32_t)t;
	h->header_len = h->wavc ? 42 : 14;
	/* decode.c:795-799 */
	if (force_chans > 0)
		h->channels = (uint32_t)force_chans;
	else if (force_chans == -1 && !h->wavc && h->channels < 2)
		h->channels = 2;
	return ACM_OK;
}

/* what the reference's read loop will do with this stream: words it can deliver
 * and blocks it will try to decode (acm_read decode.c:837-857, util.c:258-277) */
void acm_read_plan(uint32_t total, uint32_t blen, uint32_t channels, uint32_t *words_limit,
		   uint32_t *n_attempt)
{
	uint32_t c = channels ? channels : 1;
	uint32_t whole = total - total % c;
	if (blen % c == 0) {
		*words_limit = whole;
		*n_attempt = (uint32_t)(((uint64_t)total + blen - 1) / blen);
	} else {
		/* Q3: the tail of the first block is smaller than a frame, so the loop stalls */
		uint32_t first = blen - blen % c;
		*words_limit = whole < first ? whole : first;
		*n_attempt = 1;
	}
}


bool acm_stream_accepted(const acm_gpu_stream *g) { return g->total_values != 0 && g->rows != 0; }

/*
 * Device descriptor of one probed stream.  Returns ACM_OK, or the status the stream
 * gets without reaching the device.
 */
int acm_make_devstream(const acm_gpu_stream *g, uint32_t index, int pad_tail, acm::DevStream *d)
{
	uint32_t hdr = g->wavc ? 42u : 14u; /* util.c:28-29 */
	uint64_t data_off = g->in_off + hdr;
	uint64_t data_len = g->in_len > hdr ? g->in_len - hdr : 0;
	uint32_t blen = g->rows << g->level;

	/* the verdict of an earlier DECODE (-6, -7) must not keep a stream from being decoded again:
	 * status is an output */
	if (!acm_stream_accepted(g))
		return ACM_ERR_NOT_ACM;
	if (g->level > 15 || g->rows > 4095)
		return ACM_ERR_NOT_ACM; /* not representable in the 4+12-bit header field */
	/* bit positions are 32-bit on the device: images of 512 MiB and more are refused */
	if (data_len * 8 + 256 + (1u << 20) >= 0xFFFFFFFFull)
		return ACM_ERR_OTHER;
	memset(d, 0, sizeof(*d));
	d->base_off = data_off & ~(uint64_t)15;
	d->bit0 = (uint32_t)(data_off & 15u) * 8u;
	d->file_end = d->bit0 + (uint32_t)data_len * 8u;
	d->out_off = g->out_off;
	d->rows = g->rows;
	d->level = g->level;
	d->index = index;
	d->pad_words = pad_tail ? g->total_values : 0;
	acm_read_plan(g->total_values, blen, g->channels, &d->words_limit, &d->n_attempt);
	if (g->in_len < hdr)
		d->n_attempt = 0; /* the header ended inside the EOF zero byte: no data bits */
	return ACM_OK;
}
