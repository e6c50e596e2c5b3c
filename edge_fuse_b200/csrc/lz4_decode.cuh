// lz4_decode.cuh — LZ4 block decoder with LZ4_decompress_fast semantics, one warp per chunk.
//
// Replaces filemap_get's LZ4_decompress_fast(block, page, bsize) (cachemap/filemap.c:243-248 ->
// cachemap/lz4.c:1360-1363,1169-1344): decodes exactly n bytes and returns the number of
// compressed bytes consumed, which filemap_get compares with the stored compressed_length.
// Unlike the reference's trusting variant this one is bounds-checked and returns <0 on a
// malformed block (SURVEY.md App. A "Decoder freedom": any correct decoder yields the same page).
//
// The token chain is serial; literal and match copies are warp-parallel.  Match bytes are read
// back from the page being written (global memory, coherent loads, __syncwarp between
// sequences); an overlapping match (offset < length) is expanded as out[i] = out[base + i % offset]
// so that every lane only reads bytes finished before this sequence began.
#pragma once
#include "common.cuh"

namespace cmb {

// Reads an LZ4 length extension at blk[ip..cap): sum of bytes up to and including the first
// non-0xFF one, 32 bytes per step.  Returns false when the block ends first.
__device__ __forceinline__ bool lz4_read_ext(const uint8_t *blk, uint32_t cap, uint32_t &ip,
    uint32_t &len, int lane) {
	for (;;) {
		bool in = ip + lane < cap;
		uint32_t b = in ? ldg8(blk + ip + lane) : 0u;
		uint32_t stop = __ballot_sync(CMB_FULL, !in || b != 255u);
		if (stop) {
			int f = __ffs(stop) - 1;
			if (ip + f >= cap) return false;
			len += 255u * f + __shfl_sync(CMB_FULL, b, f);
			ip += f + 1;
			return true;
		}
		len += 255u * 32u;
		ip += 32;
	}
}

// Returns bytes consumed (>0) or a negative error; uniform across the warp.
__device__ int lz4_decode_warp(const uint8_t *__restrict__ blk, uint32_t cap, uint8_t *out,
    uint32_t n, int lane) {
	uint32_t ip = 0, op = 0;
	if (n == 0 || cap == 0) return -1;
	for (;;) {
		if (ip >= cap) return -1;
		uint32_t token = ldg8(blk + ip++);
		uint32_t len = token >> 4;
		if (len == 15u && !lz4_read_ext(blk, cap, ip, len, lane)) return -1;
		if (len > n - op || len > cap - ip) return -1;
		if (op + len + 8u > n) {                     // lz4.c:1242-1256: last literals
			if (op + len != n) return -1;
			warp_copy_ro(out + op, blk + ip, len, lane);
			return (int)(ip + len);              // lz4.c:1339
		}
		warp_copy_ro(out + op, blk + ip, len, lane);
		ip += len; op += len;
		if (ip + 2u > cap) return -1;
		uint32_t off = ldg8(blk + ip) | (ldg8(blk + ip + 1) << 8);
		ip += 2;
		if (off == 0u || off > op) return -1;
		len = token & 15u;
		if (len == 15u && !lz4_read_ext(blk, cap, ip, len, lane)) return -1;
		len += 4u;
		if (op + len + 5u > n) return -1;            // lz4.c:1319
		__syncwarp();                                // literals above are now visible to all lanes
		const uint8_t *from = out + op - off;
		if (off >= len) {
			if (len <= 32u) { if ((uint32_t)lane < len) out[op + lane] = from[lane]; }
			else warp_copy_rw(out + op, from, len, lane);
		} else {
			for (uint32_t i = lane; i < len; i += 32) out[op + i] = from[i % off];
		}
		op += len;
		__syncwarp();
	}
}

}  // namespace cmb
