/*
 * uint128.h — 16-byte page address and the FNV-1a-64 helper edgefs.c calls directly.
 *
 * Drop-in for the reference's cachemap/uint128.h (type at :4, FNV_hash at :6-21): edgefs.c:209
 * and edgefs.c:1911 hash object names and bucket paths with FNV_hash to build nhid_small, so the
 * name, signature and result must not change.
 */
#ifndef UINT128_H
#define UINT128_H

#include <stdint.h>

/* u first, l second, both little-endian in memory: the 16 bytes that get FNV-hashed into the
 * store key. */
typedef struct { uint64_t u; uint64_t l; } uint128_t;

/* FNV-1a, 64 bit: offset basis 0xcbf29ce484222325, prime 0x100000001b3, one byte per step. */
static inline void
FNV_hash(const void *key, int length, uint64_t *out)
{
	const unsigned char *byte = (const unsigned char *)key;
	const unsigned char *end = byte + (length > 0 ? length : 0);
	uint64_t acc = 0xcbf29ce484222325ULL;

	while (byte != end)
		acc = (acc ^ *byte++) * 0x100000001b3ULL;
	*out = acc;
}

#endif
