/* acm_host.h -- host helpers shared by acm_batch.cu and acm_stream.cu (internal). */
#ifndef ACM_HOST_H
#define ACM_HOST_H

#include <stdint.h>

typedef struct acm_header {
	uint32_t total_values, channels, acm_channels, rate, level, rows, wavc, header_len;
} acm_header;

/* read_header + read_wavc_header + force_chans (reference decode.c:679-752, :795-799).
 * p points at the first bytes of the image (at least min(len, 48) readable). */
int acm_parse_header(const uint8_t *p, uint64_t len, int force_chans, acm_header *h);

/* how far the reference's acm_read loop gets on a healthy stream (decode.c:837-857) */
void acm_read_plan(uint32_t total, uint32_t blen, uint32_t channels, uint32_t *words_limit,
		   uint32_t *n_attempt);

void acm_set_error(const char *fmt, ...);
#ifdef __cplusplus
/* acm_stream.cu keeps the buffers of closed streams for the next open; this frees them */
void acm_stream_release_pool();
#endif

#ifdef __cplusplus
struct acm_gpu_stream;
/* acm_gpu_stream::status is an OUTPUT of every decode (and of acm_gpu_probe); what tells a probed,
 * accepted stream from a rejected or never-probed one is its header fields: acm_gpu_probe leaves
 * total_values = 0 in a rejected stream, and rows = 0 only in one it has never seen */
bool acm_stream_accepted(const acm_gpu_stream *g);
namespace acm { struct DevStream; }
int acm_make_devstream(const acm_gpu_stream *g, uint32_t index, int pad_tail, acm::DevStream *d);
#endif

#endif
