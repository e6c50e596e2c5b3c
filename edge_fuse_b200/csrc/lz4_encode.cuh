// lz4_encode.cuh — byte-exact LZ4 1.8.1 block encoder, one warp per chunk (sm_100a).
//
// Emits exactly the bytes the reference's filemap_set stores:
//   LZ4_compress_fast(page, dst, n, n+1024, accel)            cachemap/filemap.c:124-128
//   -> LZ4_compress_generic<notLimited, byU16|byU32, noDict>   cachemap/lz4.c:532-733,736-771
// (byU16 + 13-bit hash4 for n < 65547, byU32 + 12-bit hash5 + MAX_DISTANCE test otherwise).
//
// The greedy parse is a serial dependency chain per chunk (every probe reads then writes the
// position table), so throughput = chunks in flight / latency per LZ4 sequence:
//   * one independent chunk per warp; per warp only the 16 KiB position table and a small sliding
//     window of the page (RING bytes: ~1 KiB ahead of the parse position, the rest behind it) live
//     in shared memory, so 9-11 chunks are resident per SM.  The window is filled 512 bytes at a
//     time with cp.async (global -> shared, no registers) one block ahead of need; probe reads,
//     literal bytes and most match candidates (LZ4 matches are mostly recent) are then
//     shared-memory reads.  Candidates older than the window fall back to ld.global.nc.
//     Staging the whole 64 KiB page would cap residency at two chunks per SM.
//   * inside a chunk the warp runs the reference's loop speculatively, one LZ4 sequence per
//     iteration and two dependent memory round trips per sequence:
//       step 1 "unified batch": lane 0 replays the table refill of position end-2 (lz4.c:691),
//         lane 1 the immediate re-test at `end` (lz4.c:694-707), lanes 2.. the first 30 probes of
//         the following search (lz4.c:593-619; probe positions are a closed form of the probe
//         index: +1, then +accel for 64 probes, +accel+1 for the next 64, ...).  All are "read
//         slot, write slot, compare 4 bytes" in program order.  Every lane stores its position
//         speculatively and reads the slot back: if all 32 see their own value no two lanes share
//         a slot, program order is irrelevant, the first hit (ballot) wins and lanes past the
//         winner put the old value back.  A clash (two lanes, one slot) or a search that needs
//         more than 30 probes goes to lz4_search_slow, which resolves program order with
//         __match_any_sync.
//       step 2 "extend": lanes 0-15 count the match forward (lz4.c:415-439) while lanes 16-31
//         catch up backward (lz4.c:622), one ballot for both.
//   * the hot loop is kept small on purpose (the profile of the first version showed a third of
//     the stall samples waiting on instruction fetch): rare paths are __noinline__.
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

// ---- sliding window -----------------------------------------------------------------------

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
	uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
	asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Window of the page in shared memory: byte p of the page lives at ring[p & (RING-1)] for
// lo <= p < done.  Blocks of 512 bytes are appended with cp.async one block ahead of need; all
// but the newest block are complete after a refill.  State is three scalars kept in registers.
constexpr uint32_t LZ4_AHEAD = 448;   // bytes past p0 one sequence may touch (accel <= 12): 30 probes + 64-byte count

struct Lz4Win { uint32_t lo, hi, done; };

// Makes [p0 - 3, p0 + LZ4_AHEAD) resident.  Normally appends one block; after a long match it
// restarts the window around p0, keeping as much history as fits.  Out of line: runs once per
// ~512 bytes of progress.  Returns lo | hi << 21 | done << 42.
template <uint32_t RING>
__device__ __noinline__ uint64_t lz4_window_refill(const uint8_t *src, uint8_t *ring, uint32_t n16,
    uint32_t lo, uint32_t hi, uint32_t p0, int lane) {
	const uint32_t want = (p0 + LZ4_AHEAD + 511u) & ~511u;
	if (p0 >= hi + 3u || p0 < lo + 3u) {                              // nothing useful resident
		// at most RING/512 blocks may be in flight at once: two copies into one ring slot would race
		const uint32_t back = (p0 - 3u) & ~511u, hist = RING - 1536u;
		lo = back > hist ? back - hist : 0u;
		hi = lo;
		__syncwarp();
	}
	const uint32_t last = ((n16 + 511u) & ~511u) + 512u;             // blocks past the page are empty
	while (hi < want + 512u && hi < last) {
		const uint32_t p = hi + 16u * lane;
		if (p < n16) cp_async16(ring + (p & (RING - 1)), src + p);
		cp_async_commit();
		hi += 512u;
		if (hi - lo > RING) lo = hi - RING;
	}
	uint32_t done;
	if (hi >= want + 512u) { cp_async_wait<1>(); done = hi - 512u; } else { cp_async_wait<0>(); done = hi; }
	__syncwarp();
	return (uint64_t)lo | ((uint64_t)hi << 21) | ((uint64_t)done << 42);
}

template <uint32_t RING>
__device__ __forceinline__ uint32_t ring32(const uint8_t *ring, uint32_t p) {
	if (RING == 0) return 0;
	const uint32_t o = p & (RING - 1) & ~3u;
	const uint32_t w0 = *reinterpret_cast<const uint32_t *>(ring + o);
	const uint32_t w1 = *reinterpret_cast<const uint32_t *>(ring + ((o + 4u) & (RING - 1)));
	return __funnelshift_r(w0, w1, (p & 3u) * 8u);
}
// Unaligned 4 bytes from the page in global memory; the word after the last one is readable
// (page buffers are padded), so no bounds predicate.
__device__ __forceinline__ uint32_t glob32(const uint8_t *src, uint32_t p) {
	const uint32_t *q = reinterpret_cast<const uint32_t *>(src + (p & ~3u));
	return __funnelshift_r(__ldg(q), __ldg(q + 1), (p & 3u) * 8u);
}

// ---- rare paths, kept out of line ------------------------------------------------------------

// Emits `count` as LZ4 length-extension bytes at dst[op..): count/255 bytes of 0xFF then count%255.
__device__ __noinline__ uint32_t lz4_emit_len(uint8_t *dst, uint32_t op, uint32_t count, int lane) {
	uint32_t nff = count / 255u;
	for (uint32_t i = lane; i < nff; i += 32) dst[op + i] = 0xFF;
	if (lane == 0) dst[op + nff] = (uint8_t)(count - nff * 255u);
	return op + nff + 1;
}

__device__ __noinline__ void lz4_copy_literals(uint8_t *dst, const uint8_t *src, uint32_t len, int lane) {
	warp_copy_ro(dst, src, len, lane);
}

// Common prefix length of src[a..) and src[b..), the a side capped at `lim` (lz4.c:415-439);
// 512 bytes per step for the long matches of repetitive pages.
__device__ __noinline__ uint32_t lz4_count_long(const uint8_t *src, uint32_t a, uint32_t b, uint32_t lim,
    uint32_t lim4, int lane) {
	uint32_t total = 0;
	for (;;) {
		const uint32_t pa = a + total + 16u * lane;
		uint32_t nb = 0;                                  // equal bytes in this lane's 16
		if (pa < lim) {
			const uint32_t avail = min(16u, lim - pa);
			const uint32_t pb = b + total + 16u * lane;
#pragma unroll
			for (uint32_t j = 0; j < 4; j++) {
				if (nb == 4u * j && 4u * j < avail) {
					uint32_t x = read32u(src, pa + 4u * j, lim4) ^ read32u(src, pb + 4u * j, lim4);
					nb += x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
				}
			}
			nb = min(nb, avail);
		}
		const uint32_t stop = __ballot_sync(CMB_FULL, nb < 16u);
		if (stop) {
			int f = __ffs(stop) - 1;
			return total + 16u * f + __shfl_sync(CMB_FULL, nb, f);
		}
		total += 512u;
	}
}

// Backward extension (lz4.c:622) continuing from (ip, match): returns extra steps.
__device__ __noinline__ uint32_t lz4_catchup_long(const uint8_t *src, uint32_t ip, uint32_t match,
    uint32_t anchor, int lane) {
	uint32_t total = 0;
	for (;;) {
		uint32_t k = total + lane + 1;
		bool ok = ip >= anchor + k && match >= k && ldg8(src + ip - k) == ldg8(src + match - k);
		uint32_t fail = __ballot_sync(CMB_FULL, !ok);
		if (fail) return total + (uint32_t)(__ffs(fail) - 1);
		total += 32;
	}
}

// Result of a search, packed so that the out-of-line function returns in registers:
// bit 63 found, bit 62 hit was the re-test lane, bits 32..61 ip, bits 0..31 match.
__device__ __forceinline__ uint64_t lz4_pack(bool found, bool retest, uint32_t ip, uint32_t match) {
	return ((uint64_t)found << 63) | ((uint64_t)retest << 62) | ((uint64_t)ip << 32) | match;
}

// The general search (any number of probes, any hash clashes), reading the page from global
// memory.  Lanes below `shift` of the first batch are the refill / re-test lanes.  Starts at
// batch g0 (0, or 32 when the inlined first batch found nothing).
template <bool WIDE>
__device__ __noinline__ uint64_t lz4_search_slow(const uint8_t *src, uint32_t lim4, Lz4Table<WIDE> tab,
    uint32_t anchor, uint32_t shift, uint32_t accel, uint32_t mflimit, uint32_t g0, int lane) {
	const uint32_t p0 = anchor + 1;
	for (;; g0 += 32) {
		const uint32_t g = g0 + lane;
		const bool special = g < shift;
		const uint32_t k = g - shift;
		uint32_t pos = p0 + lz4_probe_off(k, accel);
		const uint32_t nxt = p0 + lz4_probe_off(k + 1, accel);
		bool en = nxt <= mflimit;
		if (special) { pos = anchor - 2u + 2u * g; en = true; }
		uint32_t h = 0x10000u + lane, pseq = 0, cand = 0;
		if (en) {
			if (WIDE) { uint64_t v = read64u(src, pos, lim4); pseq = (uint32_t)v; h = lz4_hash5(v); }
			else { pseq = read32u(src, pos, lim4); h = lz4_hash4(pseq); }
			cand = tab.get(h);
		}
		const uint32_t peers = __match_any_sync(CMB_FULL, h);
		const uint32_t lower = peers & ((1u << lane) - 1u);
		const uint32_t prev_pos = __shfl_sync(CMB_FULL, pos, lower ? 31 - __clz(lower) : lane);
		if (lower) cand = prev_pos;          // what the serial loop would have stored by then
		bool hit = false;
		if (en && !(special && g == 0) && cand + LZ4_FAR >= pos)
			hit = read32u(src, cand, lim4) == pseq;
		const uint32_t hits = __ballot_sync(CMB_FULL, hit);
		const uint32_t enmask = __ballot_sync(CMB_FULL, en);
		const int w = hits ? __ffs(hits) - 1 : 31;
		const uint32_t commit = hits ? (0xffffffffu >> (31 - w)) : enmask;
		if ((commit >> lane) & 1u) {
			const uint32_t pc = peers & commit;
			if (31 - __clz(pc) == lane) tab.put(h, pos);     // last writer per slot wins
		}
		__syncwarp();
		if (hits)
			return lz4_pack(true, g0 + (uint32_t)w < shift, __shfl_sync(CMB_FULL, pos, w),
			    __shfl_sync(CMB_FULL, cand, w));
		if (enmask != CMB_FULL) return 0;
	}
}

// ---- the encoder -----------------------------------------------------------------------------

// Everything after a found match that does not fit the straight-line emitter: long literal runs,
// length bytes beyond one.  Returns the new output offset.
__device__ __noinline__ uint32_t lz4_emit_general(uint8_t *dst, uint32_t op, const uint8_t *src, uint32_t anchor,
    uint32_t lit, uint32_t off, uint32_t mc, int lane) {
	if (lane == 0) dst[op] = (uint8_t)((min(lit, 15u) << 4) | min(mc, 15u));
	op++;
	if (lit >= 15u) op = lz4_emit_len(dst, op, lit - 15u, lane);
	warp_copy_ro(dst + op, src + anchor, lit, lane);
	op += lit;
	if (lane == 0) { dst[op] = (uint8_t)off; dst[op + 1] = (uint8_t)(off >> 8); }
	op += 2;
	if (mc >= 15u) op = lz4_emit_len(dst, op, mc - 15u, lane);
	return op;
}

// Encodes src[0,n) into dst; returns the block length (uniform across the warp).
// smem: LZ4_TABLE_BYTES of table followed by RING bytes of window, 16-byte aligned.
// src must be 16-byte aligned and readable up to 16 bytes past src+n.
// DIRECT (accel > 12: probes of one batch span more than the look-ahead) reads the page from
// global memory only.
template <bool WIDE, uint32_t RING, bool DIRECT>
__device__ uint32_t lz4_encode_warp(const uint8_t *__restrict__ src, uint32_t n, uint8_t *__restrict__ dst,
    uint32_t accel, uint8_t *smem, int lane) {
	Lz4Table<WIDE> tab;
	tab.t = reinterpret_cast<decltype(tab.t)>(smem);
	uint8_t *ring = smem + LZ4_TABLE_BYTES;
	const uint32_t lim4 = (n + 3u) & ~3u;
	const uint32_t n16 = (n + 15u) & ~15u;
	uint32_t op = 0, anchor = 0;

	// lz4.c:739 — table cleared per call: an untouched slot aliases position 0.
	{
		uint4 z = make_uint4(0, 0, 0, 0);
		uint4 *t4 = reinterpret_cast<uint4 *>(smem);
#pragma unroll 4
		for (uint32_t i = lane; i < LZ4_TABLE_BYTES / 16; i += 32) t4[i] = z;
	}
	Lz4Win win = {0, 0, 0};
	__syncwarp();

	if (n >= LZ4_MIN_INPUT) {
		const uint32_t mflimit = n - LZ4_MATCH_FIND_MARGIN;
		const uint32_t mlimit = n - LZ4_TAIL_LITERALS;
		// lz4.c:583 stores position 0 under hash(0): a no-op on the cleared table, so skipped.
		uint32_t shift = 0;          // 2 once a match has ended: lanes 0,1 replay lz4.c:691-707
		for (;;) {
			const uint32_t p0 = anchor + 1;       // first probe of the search (lz4.c:584,710)
			if (!DIRECT && (p0 + LZ4_AHEAD > win.done || p0 < win.lo + 3u)) {
				uint64_t r = lz4_window_refill<RING>(src, ring, n16, win.lo, win.hi, shift ? p0 : 3u, lane);
				win.lo = (uint32_t)r & 0x1fffffu; win.hi = (uint32_t)(r >> 21) & 0x1fffffu; win.done = (uint32_t)(r >> 42);
			}
			// speculative literal byte: src[anchor + lane] (used when the run is <= 32 bytes)
			const uint32_t litbyte = DIRECT ? ldg8(src + min(anchor + lane, n - 1u)) : (uint32_t)ring[(anchor + lane) & (RING - 1)];

			// ---- step 1: unified batch ----
			const bool special = (uint32_t)lane < shift;
			const uint32_t k = (uint32_t)lane - shift;
			uint32_t pos = special ? anchor - 2u + 2u * lane : p0 + (k ? 1u + accel * (k - 1u) : 0u);
			const bool en = special || p0 + 1u + accel * k <= mflimit;
			pos = en ? pos : 0u;                                   // keep disabled lanes' reads in range
			uint32_t pseq, h;
			if (WIDE) {
				const uint64_t v = DIRECT ? read64u(src, pos, lim4)
				    : ((uint64_t)ring32<RING>(ring, pos) | ((uint64_t)ring32<RING>(ring, pos + 4u) << 32));
				pseq = (uint32_t)v; h = lz4_hash5(v);
			} else {
				pseq = DIRECT ? glob32(src, pos) : ring32<RING>(ring, pos);
				h = lz4_hash4(pseq);
			}
			uint32_t cand = tab.get(h);
			__syncwarp();
			if (en) tab.put(h, pos);                                // speculative commit
			__syncwarp();
			const uint32_t cseq = glob32(src, cand);                // latency overlaps the read-back
			const bool clash = en && tab.get(h) != (WIDE ? pos : (pos & 0xffffu));
			const bool hit = en && !(special && lane == 0) && cand + LZ4_FAR >= pos && cseq == pseq;
			const uint32_t clashes = __ballot_sync(CMB_FULL, clash);
			const uint32_t hits = __ballot_sync(CMB_FULL, hit);
			uint32_t ip, match;
			bool retest_hit;
			if (clashes == 0u && hits != 0u) {
				const int w = __ffs(hits) - 1;
				if (en && lane > w) tab.put(h, cand);                // undo past the winner
				__syncwarp();
				ip = __shfl_sync(CMB_FULL, pos, w);
				match = __shfl_sync(CMB_FULL, cand, w);
				retest_hit = (uint32_t)w < shift;
			} else {
				uint64_t res = 0;
				const uint32_t enmask = __ballot_sync(CMB_FULL, en);
				if (clashes) {                                       // two lanes, one slot: redo in order
					if (en) tab.put(h, cand);
					__syncwarp();
					res = lz4_search_slow<WIDE>(src, lim4, tab, anchor, shift, accel, mflimit, 0, lane);
				} else if (enmask == CMB_FULL) {                     // 30 probes were not enough
					res = lz4_search_slow<WIDE>(src, lim4, tab, anchor, shift, accel, mflimit, 32, lane);
				}
				if (!(res >> 63)) break;                             // -> last literals
				retest_hit = (res >> 62) & 1u;
				ip = (uint32_t)(res >> 32) & 0x3fffffffu;
				match = (uint32_t)res;
				if (!DIRECT && (ip + LZ4_AHEAD > win.done)) {         // found far ahead: move the window there
					uint64_t r = lz4_window_refill<RING>(src, ring, n16, win.lo, win.hi, ip, lane);
					win.lo = (uint32_t)r & 0x1fffffu; win.hi = (uint32_t)(r >> 21) & 0x1fffffu; win.done = (uint32_t)(r >> 42);
				}
			}

			// ---- step 2: extend forward (lanes 0-15, lz4.c:415-439) and backward (lanes 16-31,
			// lz4.c:622) with one instruction stream: lane compares 4 bytes at ip+d against match+d
			uint32_t fwd, back;
			{
				const bool fw = lane < 16;
				const uint32_t kb = (uint32_t)lane - 15u;                       // backward step of lanes 16..31
				const uint32_t pa = fw ? ip + LZ4_MIN_MATCH + 4u * lane : ip - kb;
				const uint32_t pb = fw ? match + LZ4_MIN_MATCH + 4u * lane : match - kb;
				const bool bw_ok = !fw && !retest_hit && ip >= anchor + kb && match >= kb;
				const bool deep = bw_ok && !DIRECT && pa < win.lo;               // behind the window: rare
				const bool live = fw ? pa < mlimit : (bw_ok && !deep);
				const uint32_t pas = live ? pa : ip, pbs = live ? pb : match;    // keep dead lanes in range
				const uint32_t x = (DIRECT ? glob32(src, pas) : ring32<RING>(ring, pas)) ^ glob32(src, pbs);
				uint32_t nf = x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
				nf = live ? min(nf, fw ? mlimit - pa : 1u) : 0u;
				const bool flag = fw ? nf < 4u : nf == 0u;
				const uint32_t bal = __ballot_sync(CMB_FULL, flag);
				const uint32_t f = bal & 0xffffu, b = bal >> 16;
				if (f) {
					const int fl = __ffs(f) - 1;
					fwd = 4u * fl + __shfl_sync(CMB_FULL, nf, fl);
				} else {
					fwd = 64u + lz4_count_long(src, ip + LZ4_MIN_MATCH + 64u, match + LZ4_MIN_MATCH + 64u, mlimit, lim4, lane);
				}
				if (__any_sync(CMB_FULL, deep)) back = lz4_catchup_long(src, ip, match, anchor, lane);
				else if (b) back = (uint32_t)(__ffs(b) - 1);
				else back = 16u + lz4_catchup_long(src, ip - 16u, match - 16u, anchor, lane);
			}
			const uint32_t off = ip - match;
			const uint32_t mc = back + fwd;               // lz4.c:660 matchCode
			const uint32_t lit = ip - back - anchor;
			const uint32_t end = ip + LZ4_MIN_MATCH + fwd;

			// ---- emit: token, literal run (lz4.c:625-641), offset + match length (lz4.c:643-683) ----
			if (lit <= 32u && mc < 15u + 255u) {
				uint8_t *o = dst + op;
				const uint32_t hl = 1u + (lit >= 15u);
				const uint32_t mext = mc >= 15u;
				if (lane == 0) o[0] = (uint8_t)((min(lit, 15u) << 4) | min(mc, 15u));
				if (lane == 1 && lit >= 15u) o[1] = (uint8_t)(lit - 15u);
				if ((uint32_t)lane < lit) o[hl + lane] = (uint8_t)litbyte;
				if (lane == 31) { o[hl + lit] = (uint8_t)off; o[hl + lit + 1] = (uint8_t)(off >> 8); }
				if (lane == 30 && mext) o[hl + lit + 2] = (uint8_t)(mc - 15u);
				op += hl + lit + 2u + mext;
			} else {
				op = lz4_emit_general(dst, op, src, anchor, lit, off, mc, lane);
			}

			anchor = end;
			shift = 2;
			if (end > mflimit) break;                     // lz4.c:688
		}
	}
	if (!DIRECT) cp_async_wait<0>();

	// ---- last literals (lz4.c:713-729) ----
	uint32_t run = n - anchor;
	if (lane == 0) dst[op] = (uint8_t)(min(run, 15u) << 4);
	op++;
	if (run >= 15u) op = lz4_emit_len(dst, op, run - 15u, lane);
	lz4_copy_literals(dst + op, src + anchor, run, lane);
	op += run;
	return op;
}

}  // namespace cmb
