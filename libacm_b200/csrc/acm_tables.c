/* acm_tables.c -- builds the lookup tables described in acm_tables.h. */
#include <string.h>

#include "acm_tables.h"

static const int8_t sel_of_kt[8] = { 17, 18, 20, 21, 23, 24, 26, 27 };

/*
 * Decode ONE prefix-code symbol of selector `sel` from bits b[pos..8).
 * Returns its length in bits, or 0 if it does not fit completely in the window.
 * *count = values produced (2 for the "0 = two zeros" symbol), *value = the value.
 * Symbol grammar: SURVEY.md Appendix A.4 / reference decode.c:208-403.
 */
static int k_symbol(int sel, unsigned b, int pos, int *count, int *value)
{
	static const int near2[4] = { -2, -1, 1, 2 }, far2[4] = { -3, -2, 2, 3 };
	static const int m3[8] = { -4, -3, -2, -1, 1, 2, 3, 4 };
	int pair = (sel == 17 || sel == 20 || sel == 23 || sel == 26);
	int p = pos, len;
#define BIT(i) ((b >> (i)) & 1u)
#define FIELD(i, n) ((b >> (i)) & ((1u << (n)) - 1))
	*count = 1;
	*value = 0;
	if (p + 1 > 8)
		return 0;
	if (BIT(p) == 0) { /* "0" */
		*count = pair ? 2 : 1;
		return 1;
	}
	p++;
	if (pair) { /* "1 0" = one zero */
		if (p + 1 > 8)
			return 0;
		if (BIT(p) == 0)
			return 2;
		p++;
	}
	switch (sel) {
	case 17: case 18: len = 1; break;
	case 20: case 21: len = 2; break;
	case 26: case 27: len = 3; break;
	default: /* 23: 1 1 [0 x | 1 xx]   24: 1 [0 x | 1 xx] */
		if (p + 1 > 8)
			return 0;
		len = BIT(p) ? 2 : 1;
		p++;
		if (p + len > 8)
			return 0;
		*value = len == 1 ? (FIELD(p, 1) ? 1 : -1) : far2[FIELD(p, 2)];
		return p + len - pos;
	}
	if (p + len > 8)
		return 0;
	if (len == 1)
		*value = FIELD(p, 1) ? 1 : -1;
	else if (len == 2)
		*value = near2[FIELD(p, 2)];
	else
		*value = m3[FIELD(p, 3)];
	return p + len - pos;
#undef BIT
#undef FIELD
}

void acm_tables_build(acm_tables *t)
{
	int kt, sel;
	unsigned b;

	memset(t, 0, sizeof(*t));
	for (kt = 0; kt < 8; kt++) {
		for (b = 0; b < 256; b++) {
			uint64_t e = 0;
			int pos = 0, nv = 0;
			for (;;) {
				int count, value, len, k;
				len = k_symbol(sel_of_kt[kt], b, pos, &count, &value);
				if (!len || nv + count > 7)
					break;
				pos += len;
				for (k = 0; k < count; k++, nv++) {
					e |= (uint64_t)pos << (4 + 4 * nv);
					e |= (uint64_t)(value & 15) << (32 + 4 * nv);
				}
			}
			/* every symbol is <= 5 bits, so nv >= 1 always */
			t->k8[kt * 256 + b] = e | (uint64_t)nv;
		}
	}
	for (kt = 0; kt < 8; kt++) {
		for (b = 0; b < 256; b++) {
			uint64_t vals = 0;
			int pos = 0, nv = 0;
			for (;;) {
				int count, value, len, k;
				len = k_symbol(sel_of_kt[kt], b, pos, &count, &value);
				if (!len || nv + count > 8)
					break;
				pos += len;
				for (k = 0; k < count; k++, nv++)
					vals |= (uint64_t)(value & 15) << (4 * nv);
			}
			t->k8w[kt * 256 + b] = vals | ((uint64_t)pos << 32) | ((uint64_t)(4 * nv) << 40);
		}
	}
	/* the scan walk as one state machine */
	for (b = 0; b < ACM_UNI_PSIZE; b++) {
		t->uni16[ACM_UNI_HALT * ACM_UNI_PSIZE + b] = (uint16_t)(ACM_UNI_HALT << 8);
		t->uni16[ACM_UNI_BAD * ACM_UNI_PSIZE + b] = (uint16_t)(ACM_UNI_BAD << 8);
		t->uni16[ACM_UNI_SKIP6 * ACM_UNI_PSIZE + b] = (uint16_t)(6u | (ACM_UNI_SEL << 8));
	}
	for (kt = 0; kt < 8; kt++) {
		int rem;
		for (rem = 1; rem <= 16; rem++) {
			/* rem < 16: a prefix-code page, ACM_UNI_KBITS index bits.  rem == 16: the first step
			 * of a column, taken from the selector page, 8 payload bits */
			const unsigned nb = rem < 16 ? ACM_UNI_KBITS : 8;
			for (b = 0; b < (1u << nb); b++) {
				/* whole symbols of the nb bits b (k_symbol refuses symbols that reach bit 8, so
				 * the window is shifted up to end at bit 8), until rem rows are covered */
				int pos = 8 - (int)nb, rows = 0, next;
				for (;;) {
					int count, value, len;
					len = k_symbol(sel_of_kt[kt], b << (8 - nb), pos, &count, &value);
					if (!len)
						break;
					pos += len;
					rows += count;
					if (rows >= rem)
						break;
				}
				pos -= 8 - (int)nb;
				next = rows >= rem ? ACM_UNI_SEL : ACM_UNI_K0 + kt * 15 + (rem - rows - 1);
				if (rem < 16) {
					t->uni16[(ACM_UNI_K0 + kt * 15 + rem - 1) * ACM_UNI_PSIZE + b] = (uint16_t)(pos | (next << 8));
				} else {
					unsigned idx = ((unsigned)b << 5) | (unsigned)sel_of_kt[kt];
					t->uni16[idx] = (uint16_t)((5 + pos) | (next << 8));
				}
			}
		}
	}
	for (b = 0; b < 8192; b++) {
		unsigned ind = b & 31u, adv, next = ACM_UNI_SEL;
		if (ind == 0)
			adv = 5;
		else if (ind >= 3 && ind <= 16)
			adv = 5 + 16 * ind;
		else if (ind == 19)
			adv = 5 + 30;
		else if (ind == 22)
			adv = 5 + 42;
		else if (ind == 29)
			adv = 5 + 56;
		else if (ind == 17 || ind == 18 || ind == 20 || ind == 21 || ind == 23 || ind == 24 || ind == 26 ||
			 ind == 27)
			continue; /* filled above */
		else {
			adv = 0;
			next = ACM_UNI_BAD;
		}
		if (adv > 255) {
			adv -= 6;
			next = ACM_UNI_SKIP6;
		}
		t->uni16[b] = (uint16_t)(adv | (next << 8));
	}
	for (b = 0; b < 256; b++) {
		int lo = (int)(b & 15u), hi = (int)(b >> 4);
		lo = lo >= 8 ? lo - 16 : lo;
		hi = hi >= 8 ? hi - 16 : hi;
		t->nib2w[b] = ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16);
	}
	for (b = 0; b < 128; b++) {
		unsigned v;
		/* t15 (selector 19): b < 27, digits base 3 minus 1 */
		v = b < 27 ? (((b % 3 - 1) & 15) | (((b / 3) % 3 - 1) & 15) << 4 | ((b / 9 - 1) & 15) << 8)
			   : 0x8000u;
		if (b < 32)
			t->t[0 * 128 + b] = (uint16_t)v;
		/* t27 (selector 22): b < 125, digits base 5 minus 2 */
		v = b < 125 ? (((b % 5 - 2) & 15) | (((b / 5) % 5 - 2) & 15) << 4 | ((b / 25 - 2) & 15) << 8)
			    : 0x8000u;
		t->t[1 * 128 + b] = (uint16_t)v;
		/* t37 (selector 29): b < 121, digits base 11 minus 5 */
		v = b < 121 ? (((b % 11 - 5) & 15) | ((b / 11 - 5) & 15) << 4) : 0x8000u;
		t->t[2 * 128 + b] = (uint16_t)v;
	}
	/* selector classes: the reference's filler_list (decode.c:480-489) */
	for (sel = 0; sel < 32; sel++) {
		uint8_t k = ACM_CLS_BAD;
		if (sel == 0)
			k = ACM_CLS_ZERO;
		else if (sel >= 3 && sel <= 16)
			k = ACM_CLS_LINEAR;
		else if (sel == 19)
			k = ACM_CLS_T | (0 << 3);
		else if (sel == 22)
			k = ACM_CLS_T | (1 << 3);
		else if (sel == 29)
			k = ACM_CLS_T | (2 << 3);
		else
			for (kt = 0; kt < 8; kt++)
				if (sel_of_kt[kt] == sel)
					k = (uint8_t)(ACM_CLS_K | (kt << 3));
		t->kind[sel] = k;
	}
}
