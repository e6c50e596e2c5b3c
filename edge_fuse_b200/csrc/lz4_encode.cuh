// lz4_encode.cuh — byte-exact LZ4 1.8.1 block encoder, one warp per chunk (sm_100a).
//
// Emits exactly the bytes the reference's filemap_set stores:
//   LZ4_compress_fast(page, dst, n, n+1024, accel)            cachemap/filemap.c:124-128
//   -> LZ4_compress_generic<notLimited, byU16|byU32, noDict>   cachemap/lz4.c:532-733,736-771
// (byU16 + 13-bit hash4 for n < 65547, byU32 + 12-bit hash5 + MAX_DISTANCE test otherwise).
//
// The greedy parse is a serial dependency chain per chunk (every probe reads then writes the
// position table), so the parallelism is (a) one independent chunk per warp, with only the 16 KiB
// position table in shared memory so that 12-14 chunks are resident per SM, and (b) inside a
// chunk the warp executes the reference's loop speculatively 32 probes at a time:
//   * the probe positions of a search are a closed form of the probe index (lz4.c:594-600:
//     +1, then +accel for 64 probes, +accel+1 for the next 64, ...), so lane i hashes probe
//     kbase+i;
//   * a later lane that hashes to the same slot as an earlier lane of the batch must see the
//     earlier lane's position (what the serial loop would have stored), resolved with
//     __match_any_sync;
//   * the first hit (ballot) wins and only lanes up to the winner commit their table writes,
//     the highest lane per slot last (last-writer-wins, as in the serial loop);
//   * catch-up, match length, literal copy and length-byte emission are warp-parallel.
// The page itself is read straight from HBM through the read-only L1 path (it is immutable);
// staging the 64 KiB window in shared memory was rejected because it caps residency at two
// chunks per SM for a kernel whose throughput is chunks-in-flight / latency.
#pragma once
#include "common.cuh"

namespace cmb {

constexpr uint32_t LZ4_MIN_MATCH = 4;
constexpr uint32_t LZ4_TAIL_LITERALS = 5;      // lz4.c:296
constexpr uint32_t LZ4_MATCH_FIND_MARGIN = 12; // lz4.c:297
constexpr uint32_t LZ4_MIN_INPUT = 13;         // lz4.c:298
constexpr uint32_t LZ4_NARROW_LIMIT = 65536 + 11;  // lz4.c:446
constexpr uint32_t LZ4_FAR = 65535;            // lz4.c:304-305
constexpr uint32_t LZ4_TABLE_BYTES = 16384;    // lz4.h:120

__device__ __forceinline__ uint32_t lz4_hash4(uint32_t v) { return (v * 2654435761u) >> 19; }
__device__ __forceinline__ uint32_t lz4_hash5(uint64_t v) {
	return (uint32_t)(((v << 24) * 889523592379ULL) >> 52);
}

// Offset of probe k of a search from its first probe position (lz4.c:594-600).
__device__ __forceinline__ uint32_t lz4_probe_off(uint32_t k, uint32_t accel) {
	if (k == 0) return 0;
	uint32_t m = k - 1, q = m >> 6, r = m & 63u;
	return 1u + accel * m + 32u * q * (q - 1u) + q * r;
}

template <bool WIDE> struct Lz4Table;
template <> struct Lz4Table<false> {
	uint16_t *t;
	__device__ __forceinline__ uint32_t get(uint32_t h) const { return t[h]; }
	__device__ __forceinline__ void put(uint32_t h, uint32_t pos) const { t[h] = (uint16_t)pos; }
};
template <> struct Lz4Table<true> {
	uint32_t *t;
	__device__ __forceinline__ uint32_t get(uint32_t h) const { return t[h]; }
	__device__ __forceinline__ void put(uint32_t h, uint32_t pos) const { t[h] = pos; }
};

template <bool WIDE>
__device__ __forceinline__ uint32_t lz4_hash_at(const uint8_t *src, uint32_t pos, uint32_t lim4, uint32_t &seq) {
	if (WIDE) {
		uint64_t v = read64u(src, pos, lim4);
		seq = (uint32_t)v;
		return lz4_hash5(v);
	}
	seq = read32u(src, pos, lim4);
	return lz4_hash4(seq);
}

// Emits `count` as LZ4 length-extension bytes at dst[op..): count/255 bytes of 0xFF then count%255.
__device__ __forceinline__ uint32_t lz4_emit_len(uint8_t *dst, uint32_t op, uint32_t count, int lane) {
	uint32_t nff = count / 255u;
	for (uint32_t i = lane; i < nff; i += 32) dst[op + i] = 0xFF;
	if (lane == 0) dst[op + nff] = (uint8_t)(count - nff * 255u);
	return op + nff + 1;
}

// Common prefix length of src[a..) and src[b..), the a side capped at `lim` (lz4.c:415-439).
__device__ __forceinline__ uint32_t lz4_warp_count(const uint8_t *src, uint32_t a, uint32_t b,
    uint32_t lim, uint32_t lim4, int lane) {
	uint32_t total = 0;
	for (;;) {
		uint32_t pa = a + total + 4u * lane;
		uint32_t avail = pa < lim ? min(4u, lim - pa) : 0u;
		uint32_t n = 0;
		if (avail) {
			uint32_t x = read32u(src, pa, lim4) ^ read32u(src, b + total + 4u * lane, lim4);
			n = x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
			n = min(n, avail);
		}
		uint32_t stop = __ballot_sync(CMB_FULL, n < 4u);
		if (stop) {
			int f = __ffs(stop) - 1;
			return total + 4u * f + __shfl_sync(CMB_FULL, n, f);
		}
		total += 128u;
	}
}

// Encodes src[0,n) into dst; returns the block length (uniform across the warp).
// `tab_raw` is this warp's 16 KiB of shared memory.  src must be 4-byte aligned.
template <bool WIDE>
__device__ uint32_t lz4_encode_warp(const uint8_t *__restrict__ src, uint32_t n, uint8_t *__restrict__ dst,
    uint32_t accel, void *tab_raw, int lane) {
	Lz4Table<WIDE> tab;
	tab.t = reinterpret_cast<decltype(tab.t)>(tab_raw);
	const uint32_t lim4 = (n + 3u) & ~3u;
	uint32_t op = 0, anchor = 0;

	// lz4.c:739 — table cleared per call: an untouched slot aliases position 0.
	{
		uint4 z = make_uint4(0, 0, 0, 0);
		uint4 *t4 = reinterpret_cast<uint4 *>(tab_raw);
		for (uint32_t i = lane; i < LZ4_TABLE_BYTES / 16; i += 32) t4[i] = z;
	}
	__syncwarp();

	if (n >= LZ4_MIN_INPUT) {
		const uint32_t mflimit = n - LZ4_MATCH_FIND_MARGIN;
		const uint32_t mlimit = n - LZ4_TAIL_LITERALS;
		// lz4.c:583 stores position 0 under hash(0): a no-op on the cleared table, so skipped.
		uint32_t p0 = 1;                                                    // lz4.c:584
		for (;;) {
			// ---- search (lz4.c:593-619), 32 probes per step ----
			uint32_t ip = 0, match = 0;
			bool found = false;
			for (uint32_t kbase = 0;; kbase += 32) {
				uint32_t k = kbase + lane;
				uint32_t pos = p0 + lz4_probe_off(k, accel);
				uint32_t nxt = p0 + lz4_probe_off(k + 1, accel);
				bool valid = nxt <= mflimit;
				uint32_t h = 0x10000u + lane, pseq = 0, cand = 0;
				if (valid) {
					h = lz4_hash_at<WIDE>(src, pos, lim4, pseq);
					cand = tab.get(h);
				}
				uint32_t peers = __match_any_sync(CMB_FULL, h);
				uint32_t lower = peers & ((1u << lane) - 1u);
				int prev = lower ? 31 - __clz(lower) : lane;
				uint32_t prev_pos = __shfl_sync(CMB_FULL, pos, prev);
				if (lower) cand = prev_pos;
				bool hit = false;
				if (valid && (!WIDE || cand + LZ4_FAR >= pos))
					hit = read32u(src, cand, lim4) == pseq;
				uint32_t hits = __ballot_sync(CMB_FULL, hit);
				uint32_t vmask = __ballot_sync(CMB_FULL, valid);
				int w = hits ? __ffs(hits) - 1 : 31;
				uint32_t commit = hits ? (0xffffffffu >> (31 - w)) : vmask;
				if ((commit >> lane) & 1u) {
					uint32_t pc = peers & commit;
					if (31 - __clz(pc) == lane) tab.put(h, pos);
				}
				__syncwarp();
				if (hits) {
					ip = __shfl_sync(CMB_FULL, pos, w);
					match = __shfl_sync(CMB_FULL, cand, w);
					found = true;
					break;
				}
				if (vmask != CMB_FULL) break;
			}
			if (!found) break;   // -> last literals

			// ---- catch-up (lz4.c:622) ----
			for (;;) {
				uint32_t k = lane + 1;
				bool ok = ip >= anchor + k && match >= k &&
				    ldg8(src + ip - k) == ldg8(src + match - k);
				uint32_t fail = __ballot_sync(CMB_FULL, !ok);
				uint32_t back = fail ? (uint32_t)(__ffs(fail) - 1) : 32u;
				ip -= back; match -= back;
				if (back < 32u) break;
			}

			// ---- literal run (lz4.c:625-641) ----
			uint32_t lit = ip - anchor;
			uint32_t tok = op++;
			uint32_t tokval = min(lit, 15u) << 4;
			if (lit >= 15u) op = lz4_emit_len(dst, op, lit - 15u, lane);
			warp_copy_ro(dst + op, src + anchor, lit, lane);
			op += lit;

			bool done = false;
			for (;;) {
				// ---- offset + match length (lz4.c:643-683) ----
				uint32_t off = ip - match;
				if (lane == 0) { dst[op] = (uint8_t)off; dst[op + 1] = (uint8_t)(off >> 8); }
				op += 2;
				uint32_t mc = lz4_warp_count(src, ip + LZ4_MIN_MATCH, match + LZ4_MIN_MATCH, mlimit, lim4, lane);
				ip += LZ4_MIN_MATCH + mc;
				if (lane == 0) dst[tok] = (uint8_t)(tokval | min(mc, 15u));
				if (mc >= 15u) op = lz4_emit_len(dst, op, mc - 15u, lane);
				anchor = ip;
				if (ip > mflimit) { done = true; break; }          // lz4.c:688
				// ---- lz4.c:691-707: refill ip-2, test ip ----
				uint32_t s2, s0;
				uint32_t h2 = lz4_hash_at<WIDE>(src, ip - 2, lim4, s2);
				uint32_t h0 = lz4_hash_at<WIDE>(src, ip, lim4, s0);
				if (lane == 0) tab.put(h2, ip - 2);
				__syncwarp();
				uint32_t m = tab.get(h0);
				__syncwarp();
				if (lane == 0) tab.put(h0, ip);
				__syncwarp();
				if (m + LZ4_FAR >= ip && read32u(src, m, lim4) == s0) {
					match = m;
					tok = op++;
					tokval = 0;
					continue;
				}
				break;
			}
			if (done) break;
			p0 = ip + 1;                                           // lz4.c:710
		}
	}

	// ---- last literals (lz4.c:713-729) ----
	uint32_t run = n - anchor;
	if (lane == 0) dst[op] = (uint8_t)(min(run, 15u) << 4);
	op++;
	if (run >= 15u) op = lz4_emit_len(dst, op, run - 15u, lane);
	warp_copy_ro(dst + op, src + anchor, run, lane);
	op += run;
	return op;
}

}  // namespace cmb
