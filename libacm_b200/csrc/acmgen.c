/*
 * acmgen.c -- synthetic InterPlay ACM stream generator.
 *
 * markokr/libacm ships no encoder and no sample files, so every input this
 * repository decodes is produced here.  The generator emits *valid* bitstreams
 * (verified against the reference decoder) that exercise
 * every filler of the reference's table (decode.c:480-489), every level/row
 * shape and both container variants (plain 14-byte header, decode.c:712-752;
 * 28-byte WAVC pre-header, decode.c:687-710), and obeys the constraints that
 * keep the reference's behaviour defined (Appendix D): pwr >= need(ind) so the
 * dequantisation table covers every index (decode.c:591-600), t-codes below
 * their radix limit, no bad selector unless a negative test asks for one.
 *
 * It draws code SYMBOLS, not PCM: any valid symbol sequence is a valid stream,
 * and the decoder under test (reference, oracle or GPU) defines the PCM.
 *
 * Built both as a shared library (ctypes: acmgen_bound / acmgen_write /
 * acmgen_write_many) and, with -DACMGEN_MAIN, as a CLI.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "acmgen.h"
#include "acmgen_core.h"

size_t acmgen_bound(const acmgen_params *p) { return acmgen_bound_core(p); }

size_t acmgen_write(const acmgen_params *p, uint8_t *out, size_t cap)
{
	uint8_t *inds = malloc((size_t)1 << p->level);
	size_t n;
	if (!inds)
		return 0;
	n = acmgen_write_core(p, out, cap, inds, 0);
	free(inds);
	return n;
}

/*
 * Batch helper: n streams, stream i uses params[i]; images are packed
 * back-to-back, each starting on a 16-byte boundary, into blob[0..cap).
 * offs[i]/lens[i] receive the placement.  Returns bytes used or 0 on overflow.
 */
size_t acmgen_write_many(const acmgen_params *params, size_t n, uint8_t *blob, size_t cap,
			 uint64_t *offs, uint32_t *lens)
{
	size_t i, at = 0;
	for (i = 0; i < n; i++) {
		size_t got;
		at = (at + 15) & ~(size_t)15;
		if (at >= cap)
			return 0;
		got = acmgen_write(&params[i], blob + at, cap - at);
		if (!got)
			return 0;
		offs[i] = at;
		lens[i] = (uint32_t)got;
		at += got;
	}
	return at;
}

#ifdef ACMGEN_MAIN
int main(int argc, char **argv)
{
	acmgen_params p;
	uint8_t *buf;
	size_t cap, n;
	FILE *f;
	if (argc < 8) {
		fprintf(stderr, "usage: acmgen OUT level rows channels rate total_values seed [dist 0|1|2] [wavc] [single_ind]\n");
		return 2;
	}
	memset(&p, 0, sizeof(p));
	p.level = atoi(argv[2]);
	p.rows = atoi(argv[3]);
	p.channels = atoi(argv[4]);
	p.rate = atoi(argv[5]);
	p.total_values = strtoul(argv[6], NULL, 0);
	p.seed = strtoull(argv[7], NULL, 0);
	p.dist = argc > 8 ? atoi(argv[8]) : ACMGEN_DIST_FALLOUT;
	p.wavc = argc > 9 ? atoi(argv[9]) : 0;
	p.single_ind = argc > 10 ? atoi(argv[10]) : 0;
	p.pzero = 128;
	cap = acmgen_bound(&p);
	buf = malloc(cap);
	n = acmgen_write(&p, buf, cap);
	if (!n || !(f = fopen(argv[1], "wb")))
		return 1;
	fwrite(buf, 1, n, f);
	fclose(f);
	return 0;
}
#endif
