/*
 * ef_oracle.h — CPU oracle for the edge-fuse cachemap hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and there only as the checker or the timed CPU baseline.  The product library
 * (edge_fuse_b200/csrc) never links, includes or calls anything in this directory.
 *
 * Every function is an independent restatement of the algorithm the reference runs on the
 * path; the citation after each prototype is the reference code it follows (paths relative
 * to /root/reference).  The restatement is pinned against the reference's own sources
 * compiled as oracle/_ref/libcachemap_ref.so (see oracle/Makefile, tests/test_oracle_pin.py)
 * and against the committed vectors in tests/golden/ that were produced with that library.
 *
 * Exception — ef_fingerprint128(): the reference has no content fingerprint,
 * so EF128 is a new definition and this is its CPU statement:
 * PARITY UNPINNED for that one function (self-consistency KATs only).
 */
#ifndef EF_ORACLE_H
#define EF_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 16-byte page address, memory order u then l (cachemap/uint128.h:4). */
typedef struct { uint64_t u; uint64_t l; } ef_addr_t;

/* Byte-serial FNV-1a 64 (cachemap/uint128.h:6-21). */
uint64_t ef_fnv1a64(const void *buf, size_t len);

/* Address composition (cachemap/cachemap.c:151-166): returns -1 when offset>>pshift does not
 * fit 44 bits, else fills {u = nhid_small, l = page | (uint64)genid << 44} and returns 0. */
int ef_addr_compose(uint64_t offset, uint64_t nhid_small, uint32_t genid, int pshift,
    ef_addr_t *out);

/* Store key = FNV-1a-64 of the 16 address bytes; shard = key & 31
 * (cachemap/filemap.c:18-33, cachemap/filemap.h:7). */
uint64_t ef_addr_key(const ef_addr_t *a);
int ef_key_shard(uint64_t key);

/* LZ4 1.8.1 block encoder exactly as filemap_set reaches it:
 * LZ4_compress_fast(src,dst,n,n+1024,accel) -> LZ4_compress_generic<notLimited,
 * byU16 if n < 65547 else byU32, noDict, noDictIssue> (cachemap/lz4.c:532-733,736-771;
 * hashes lz4.c:475-498; constants lz4.c:293-310,446-447; call site cachemap/filemap.c:124-128).
 * dst needs n + n/255 + 16 bytes.  Returns the block length; only dst[0,ret) is defined. */
int ef_lz4_encode(const uint8_t *src, int n, uint8_t *dst, int accel);

/* LZ4 block decoder with LZ4_decompress_fast semantics (cachemap/lz4.c:1169-1344,1360-1363;
 * call site cachemap/filemap.c:243-248): decodes exactly n output bytes and returns the number
 * of compressed bytes consumed; <0 on a malformed block.  Bounds-checked (src_cap = bytes
 * readable at src), unlike the reference's trusting variant. */
int ef_lz4_decode(const uint8_t *src, int src_cap, uint8_t *dst, int n);

/* LZ4_compressBound (cachemap/lz4.h:157). */
int ef_lz4_bound(int n);

/* Store record header, 24 bytes: {u64 u; u64 l; i32 compressed_length; 4 pad}
 * (cachemap/filemap.c:9-12,140-147).  Pad bytes are unspecified in the reference
 * (stack garbage); this oracle and the product write zeros. */
void ef_record_prefix(const ef_addr_t *a, int32_t compressed_length, uint8_t out[24]);

/* EF128 content fingerprint (NEW definition; parity unpinned). out = {hi, lo}. */
void ef_fingerprint128(const uint8_t *data, size_t n, uint64_t out[2]);

#ifdef __cplusplus
}
#endif
/* Caller-side gate and object id (edgefs.c:192-212,1911). */
int ef_cache_check(int have_cache, int pshift, uint64_t off, uint64_t size, uint64_t *page_size_out,
    uint64_t *aligned_off_out);
uint64_t ef_build_nhid(const char *name, const char *bucket_path);

#endif
