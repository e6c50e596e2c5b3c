/*
 * oracle_port_abi.c — gives the oracle port the same call shapes as the reference's LZ4 entry
 * points so cpu_bench.c can time either one.  TEST INFRASTRUCTURE ONLY (see ef_oracle.h).
 */
#include "ef_oracle.h"

int
ef_port_compress_fast(const char *src, char *dst, int n, int cap, int accel)
{
	(void)cap;
	return ef_lz4_encode((const uint8_t *)src, n, (uint8_t *)dst, accel);
}

int
ef_port_decompress_fast(const char *src, char *dst, int n)
{
	return ef_lz4_decode((const uint8_t *)src, n + n / 255 + 16, (uint8_t *)dst, n);
}
