/* Exposes the static inline helpers of include/edgefs_glue.h to the Python tests. */
#include "edgefs_glue.h"

int shim_cache_check(int have_cache, int pshift, uint64_t off, uint64_t size, uint64_t *ps, uint64_t *ao)
{
	return edgefs_cache_check(have_cache, pshift, off, (size_t)size, ps, ao);
}

uint64_t shim_build_nhid(const char *name, const char *bucket_path)
{
	return edgefs_build_nhid(name, edgefs_bucket_hid(bucket_path));
}
