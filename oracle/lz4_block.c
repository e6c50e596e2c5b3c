/*
 * lz4_block.c — oracle restatement of the one LZ4 1.8.1 block encoder / decoder instance the
 * reference's filemap reaches.  TEST INFRASTRUCTURE ONLY (see ef_oracle.h).
 *
 * Written from the behavioural description in SURVEY.md Appendix A and checked line by line
 * against /root/reference/cachemap/lz4.c; positions are integer offsets into the input, the
 * probe table holds positions, and there is no wild copying: only dst[0,ret) is ever written.
 */
#include <string.h>
#include "ef_oracle.h"

enum {
	MIN_MATCH = 4,          /* lz4.c:293 */
	TAIL_LITERALS = 5,      /* lz4.c:296 LASTLITERALS */
	MATCH_FIND_MARGIN = 12, /* lz4.c:297 MFLIMIT */
	MIN_INPUT = 13,         /* lz4.c:298 LZ4_minLength */
	NARROW_LIMIT = 65536 + 11, /* lz4.c:446 LZ4_64Klimit */
	FAR = 65535,            /* lz4.c:304-305 MAX_DISTANCE */
	SKIP_SHIFT = 6          /* lz4.c:447 LZ4_skipTrigger */
};

static inline uint32_t le32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t le64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

/* lz4.c:475-498.  narrow (byU16): 13-bit hash of 4 bytes; wide (byU32 on 64-bit): 12-bit
 * hash of the low 5 bytes of an 8-byte little-endian read. */
static inline uint32_t
probe_hash(const uint8_t *src, uint32_t pos, int wide)
{
	if (!wide)
		return (le32(src + pos) * 2654435761u) >> 19;
	return (uint32_t)(((le64(src + pos) << 24) * 889523592379ULL) >> 52);
}

/* lz4.c:415-439: length of the common prefix of src[a..] and src[b..], a side capped at lim. */
static inline uint32_t
common_len(const uint8_t *src, uint32_t a, uint32_t b, uint32_t lim)
{
	uint32_t n = 0;
	while (a + n < lim && src[a + n] == src[b + n])
		n++;
	return n;
}

static inline uint32_t
put_run_len(uint8_t *dst, uint32_t op, uint32_t rem)
{
	while (rem >= 255) { dst[op++] = 255; rem -= 255; }
	dst[op++] = (uint8_t)rem;
	return op;
}

int
ef_lz4_bound(int n)
{
	return n + n / 255 + 16;
}

int
ef_lz4_encode(const uint8_t *src, int n_in, uint8_t *dst, int accel)
{
	static __thread uint32_t table[8192];
	const uint32_t n = (uint32_t)n_in;
	const int wide = n_in >= NARROW_LIMIT;      /* lz4.c:742-746 */
	uint32_t op = 0, anchor = 0, ip, match, tok;

	if (n_in < 0)
		return 0;
	if (accel < 1)
		accel = 1;                           /* lz4.c:740 */
	memset(table, 0, sizeof(table));             /* lz4.c:739 */

	if (n < MIN_INPUT)                           /* lz4.c:580 */
		goto tail;

	const uint32_t mflimit = n - MATCH_FIND_MARGIN; /* lz4.c:553 */
	const uint32_t mlimit = n - TAIL_LITERALS;      /* lz4.c:554 */

	table[probe_hash(src, 0, wide)] = 0;         /* lz4.c:583 */
	ip = 1;
	uint32_t next_h = probe_hash(src, 1, wide);  /* lz4.c:584 */

	for (;;) {
		/* search, lz4.c:593-619 */
		uint32_t fwd = ip, step = 1, nb = (uint32_t)accel << SKIP_SHIFT;
		for (;;) {
			uint32_t h = next_h;
			ip = fwd;
			fwd += step;
			step = nb++ >> SKIP_SHIFT;
			if (fwd > mflimit)
				goto tail;
			match = table[h];
			next_h = probe_hash(src, fwd, wide);
			table[h] = ip;
			if (wide && match + FAR < ip)
				continue;
			if (le32(src + match) == le32(src + ip))
				break;
		}
		/* catch-up, lz4.c:622 */
		while (ip > anchor && match > 0 && src[ip - 1] == src[match - 1]) {
			ip--;
			match--;
		}
		/* literal run, lz4.c:625-641 */
		{
			uint32_t lit = ip - anchor;
			tok = op++;
			if (lit >= 15) {
				dst[tok] = 0xF0;
				op = put_run_len(dst, op, lit - 15);
			} else {
				dst[tok] = (uint8_t)(lit << 4);
			}
			memcpy(dst + op, src + anchor, lit);
			op += lit;
		}
		for (;;) {
			/* offset + match length, lz4.c:643-683 */
			uint32_t off = ip - match;
			dst[op++] = (uint8_t)off;
			dst[op++] = (uint8_t)(off >> 8);
			uint32_t mc = common_len(src, ip + MIN_MATCH, match + MIN_MATCH, mlimit);
			ip += MIN_MATCH + mc;
			if (mc >= 15) {
				dst[tok] += 15;
				op = put_run_len(dst, op, mc - 15);
			} else {
				dst[tok] += (uint8_t)mc;
			}
			anchor = ip;
			if (ip > mflimit)                /* lz4.c:688 */
				goto tail;
			/* lz4.c:691-707 */
			table[probe_hash(src, ip - 2, wide)] = ip - 2;
			uint32_t h = probe_hash(src, ip, wide);
			match = table[h];
			table[h] = ip;
			if (match + FAR >= ip && le32(src + match) == le32(src + ip)) {
				tok = op++;
				dst[tok] = 0;
				continue;
			}
			break;
		}
		next_h = probe_hash(src, ++ip, wide);    /* lz4.c:710 */
	}

tail:   /* lz4.c:713-729 */
	{
		uint32_t run = n - anchor;
		if (run >= 15) {
			dst[op++] = 0xF0;
			op = put_run_len(dst, op, run - 15);
		} else {
			dst[op++] = (uint8_t)(run << 4);
		}
		memcpy(dst + op, src + anchor, run);
		op += run;
	}
	return (int)op;
}

int
ef_lz4_decode(const uint8_t *src, int src_cap, uint8_t *dst, int n_out)
{
	const uint32_t n = (uint32_t)n_out, cap = (uint32_t)src_cap;
	uint32_t ip = 0, op = 0;

	if (n_out <= 0 || src_cap <= 0)
		return -1;
	for (;;) {
		if (ip >= cap) return -1;
		uint32_t token = src[ip++];                 /* lz4.c:1211 */
		uint32_t len = token >> 4;
		if (len == 15) {                            /* lz4.c:1231-1239 */
			uint32_t s;
			do {
				if (ip >= cap) return -1;
				s = src[ip++];
				len += s;
			} while (s == 255);
		}
		if (len > n - op || len > cap - ip) return -1;
		/* lz4.c:1242-1256: a literal run reaching past n-8 must end the block exactly */
		if (op + len + 8 > n) {
			if (op + len != n) return -1;
			memcpy(dst + op, src + ip, len);
			ip += len;
			return (int)ip;                     /* lz4.c:1339 */
		}
		memcpy(dst + op, src + ip, len);
		ip += len;
		op += len;
		if (ip + 2 > cap) return -1;
		uint32_t off = src[ip] | ((uint32_t)src[ip + 1] << 8);  /* lz4.c:1261 */
		ip += 2;
		if (off == 0 || off > op) return -1;
		len = token & 15;
		if (len == 15) {                            /* lz4.c:1267-1276 */
			uint32_t s;
			do {
				if (ip >= cap) return -1;
				s = src[ip++];
				len += s;
			} while (s == 255);
		}
		len += MIN_MATCH;
		if (op + len + TAIL_LITERALS > n) return -1;    /* lz4.c:1319 */
		for (uint32_t i = 0; i < len; i++)          /* overlap-safe byte copy */
			dst[op + i] = dst[op + i - off];
		op += len;
	}
}
