/*
 * edgefs_glue.h — the caller-side helpers of the page-cache path.
 *
 * In the reference these are static functions inside edgefs.c, between the FUSE callbacks and
 * cachemap_get/put:
 *   edgefs.c:192-203   cachemap_cache_check  — a request is cached only if it starts and ends on
 *                                              page boundaries
 *   edgefs.c:205-212   cachemap_build_nhid   — nhid_small = FNV(object name) ^ FNV(bucket path)
 *   edgefs.c:1911      parse_url tail        — bhid_small = FNV(url path)
 * They are restated here, without the reference's globals (cachemap_pshift, cachemap_obj become
 * arguments), so that a caller of cachemap_read_range/_write_range (cachemap.h) and the parity
 * tests use one definition.  Header-only, plain C.
 */
#ifndef EDGEFS_GLUE_H
#define EDGEFS_GLUE_H

#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include "uint128.h"

/* edgefs.c:192-203.  Returns non-zero when [off, off+size) may go through the cache: a cache
 * exists, and both ends are multiples of the page size.  The outputs are always written. */
static inline int
edgefs_cache_check(int have_cache, int pshift, uint64_t off, size_t size, uint64_t *page_size_out,
    uint64_t *aligned_off_out)
{
	const uint64_t page = 1ULL << pshift;
	const uint64_t head = off & (page - 1);
	const uint64_t tail = (off + (uint64_t)size) & (page - 1);

	*page_size_out = page;
	*aligned_off_out = off - head;
	return have_cache && head == 0 && tail == 0;
}

/* edgefs.c:1911: the bucket id is the FNV-1a-64 of the url path as parse_url left it. */
static inline uint64_t
edgefs_bucket_hid(const char *path)
{
	uint64_t h;
	FNV_hash(path, (int)strlen(path), &h);
	return h;
}

/* edgefs.c:205-212 */
static inline uint64_t
edgefs_build_nhid(const char *name, uint64_t bhid_small)
{
	uint64_t h;
	FNV_hash(name, (int)strlen(name), &h);
	return h ^ bhid_small;
}

#endif
