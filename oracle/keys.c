/*
 * keys.c — oracle restatement of address composition, FNV-1a-64 keying and the record header.
 * TEST INFRASTRUCTURE ONLY (see ef_oracle.h).
 */
#include <string.h>
#include "ef_oracle.h"

/* cachemap/uint128.h:6-21 */
uint64_t
ef_fnv1a64(const void *buf, size_t len)
{
	const uint8_t *p = (const uint8_t *)buf;
	uint64_t h = 14695981039346656037ULL;
	for (size_t i = 0; i < len; i++)
		h = (h ^ p[i]) * 0x100000001b3ULL;
	return h;
}

/* cachemap/cachemap.c:151-166 (PNUM_SHIFT 44) */
int
ef_addr_compose(uint64_t offset, uint64_t nhid_small, uint32_t genid, int pshift, ef_addr_t *out)
{
	uint64_t page = offset >> pshift;
	if (page >> 44)
		return -1;
	out->l = page | ((uint64_t)genid << 44);
	out->u = nhid_small;
	return 0;
}

/* cachemap/filemap.c:18-24: FNV over the in-memory bytes of {u,l} (little-endian host) */
uint64_t
ef_addr_key(const ef_addr_t *a)
{
	uint8_t raw[16];
	memcpy(raw, &a->u, 8);
	memcpy(raw + 8, &a->l, 8);
	return ef_fnv1a64(raw, 16);
}

/* cachemap/filemap.c:30, cachemap/filemap.h:7 */
int
ef_key_shard(uint64_t key)
{
	return (int)(key & 31);
}

/* cachemap/filemap.c:9-12,140-147 */
void
ef_record_prefix(const ef_addr_t *a, int32_t compressed_length, uint8_t out[24])
{
	memset(out, 0, 24);
	memcpy(out, &a->u, 8);
	memcpy(out + 8, &a->l, 8);
	memcpy(out + 16, &compressed_length, 4);
}

/* edgefs.c:192-203 (cachemap_cache_check): cache only requests whose two ends are page-aligned.
 * The reference reads the globals cachemap_pshift / cachemap_obj; they are arguments here. */
int
ef_cache_check(int have_cache, int pshift, uint64_t off, uint64_t size, uint64_t *page_size_out,
    uint64_t *aligned_off_out)
{
	uint64_t page_size = 1ULL << pshift;
	uint64_t unaligned_start = off & (page_size - 1);
	uint64_t unaligned_end = (off + size) & (page_size - 1);
	*page_size_out = page_size;
	*aligned_off_out = off - unaligned_start;
	return have_cache && !unaligned_start && !unaligned_end;
}

/* edgefs.c:205-212 (cachemap_build_nhid) with bhid_small from edgefs.c:1911 (FNV of the url path) */
uint64_t
ef_build_nhid(const char *name, const char *bucket_path)
{
	return ef_fnv1a64(name, strlen(name)) ^ ef_fnv1a64(bucket_path, strlen(bucket_path));
}
