/*
 * fingerprint.c — CPU statement of the EF128 content fingerprint.
 * TEST INFRASTRUCTURE ONLY (see ef_oracle.h).  PARITY UNPINNED: the reference has no content
 * hash; this file and the CUDA kernel are two independent implementations of
 * the same written spec and are compared with each other.
 *
 * Spec.  32 lanes, lane l owns accumulators (a,b) and four secrets S[4l..4l+3] where
 * S[i] = output number i (from 0) of splitmix64 seeded with state 0x4544474546555345,
 * i.e. mix(0x4544474546555345 + (i+1) * 0x9E3779B97F4A7C15); a,b start as S2,S3.  Input is consumed in
 * 512-byte stripes; in stripe s lane l reads x0 = LE64(data[512s+16l]), x1 = LE64(+8), bytes at
 * or beyond n reading as zero.  Per stripe:  d0 = x0^S0, d1 = x1^S1,
 *   a += lo32(d0)*hi32(d0) + x1 ;  b += lo32(d1)*hi32(d1) + x0.
 * After every 16th stripe:  a = ((a ^ a>>47) ^ S2) * 0x9E3779B1 ; b = ((b ^ b>>47) ^ S3) * 0x85EBCA77.
 * Finish: u_l = fold(a^S0, b^S1), v_l = fold(a^S3, b^S2) with fold(x,y) = lo64(x*y) ^ hi64(x*y);
 * U = n*P1 + sum u_l, V = ~(n*P2) + sum v_l (mod 2^64); lo = av(U), hi = av(V),
 * av(h): h ^= h>>37; h *= 0x165667919E3779F9; h ^= h>>32.
 */
#include <string.h>
#include "ef_oracle.h"

#define P1 0x9E3779B185EBCA87ULL
#define P2 0xC2B2AE3D27D4EB4FULL

static uint64_t
secret(unsigned i)
{
	uint64_t z = 0x4544474546555345ULL + (uint64_t)(i + 1) * 0x9E3779B97F4A7C15ULL;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

static uint64_t
fold(uint64_t x, uint64_t y)
{
	__uint128_t p = (__uint128_t)x * y;
	return (uint64_t)p ^ (uint64_t)(p >> 64);
}

static uint64_t
av(uint64_t h)
{
	h ^= h >> 37;
	h *= 0x165667919E3779F9ULL;
	h ^= h >> 32;
	return h;
}

static uint64_t
le64_padded(const uint8_t *data, size_t n, size_t pos)
{
	uint8_t t[8] = {0};
	if (pos < n)
		memcpy(t, data + pos, n - pos < 8 ? n - pos : 8);
	uint64_t v;
	memcpy(&v, t, 8);
	return v;
}

void
ef_fingerprint128(const uint8_t *data, size_t n, uint64_t out[2])
{
	uint64_t a[32], b[32], S[128];
	for (unsigned i = 0; i < 128; i++)
		S[i] = secret(i);
	for (unsigned l = 0; l < 32; l++) {
		a[l] = S[4 * l + 2];
		b[l] = S[4 * l + 3];
	}
	size_t stripes = (n + 511) / 512;
	for (size_t s = 0; s < stripes; s++) {
		for (unsigned l = 0; l < 32; l++) {
			uint64_t x0 = le64_padded(data, n, 512 * s + 16 * l);
			uint64_t x1 = le64_padded(data, n, 512 * s + 16 * l + 8);
			uint64_t d0 = x0 ^ S[4 * l], d1 = x1 ^ S[4 * l + 1];
			a[l] += (uint64_t)(uint32_t)d0 * (d0 >> 32) + x1;
			b[l] += (uint64_t)(uint32_t)d1 * (d1 >> 32) + x0;
			if ((s & 15) == 15) {
				a[l] = ((a[l] ^ (a[l] >> 47)) ^ S[4 * l + 2]) * 0x9E3779B1ULL;
				b[l] = ((b[l] ^ (b[l] >> 47)) ^ S[4 * l + 3]) * 0x85EBCA77ULL;
			}
		}
	}
	uint64_t U = (uint64_t)n * P1, V = ~((uint64_t)n * P2);
	for (unsigned l = 0; l < 32; l++) {
		U += fold(a[l] ^ S[4 * l], b[l] ^ S[4 * l + 1]);
		V += fold(a[l] ^ S[4 * l + 3], b[l] ^ S[4 * l + 2]);
	}
	out[0] = av(V);   /* hi */
	out[1] = av(U);   /* lo */
}
