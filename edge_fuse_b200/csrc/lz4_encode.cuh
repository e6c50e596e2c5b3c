// lz4_encode.cuh — byte-exact LZ4 1.8.1 block encoder, one warp per chunk (sm_100a).
//
// Emits exactly the bytes the reference's filemap_set stores:
//   LZ4_compress_fast(page, dst, n, n+1024, accel)            cachemap/filemap.c:124-128
//   -> LZ4_compress_generic<notLimited, byU16|byU32, noDict>   cachemap/lz4.c:532-733,736-771
// (byU16 + 13-bit hash4 for n < 65547, byU32 + 12-bit hash5 + MAX_DISTANCE test otherwise).
//
// The greedy parse is a serial dependency chain per chunk (every probe reads then writes the
// position table), so throughput = chunks in flight / latency per LZ4 sequence.
//   * Chunks in flight: one independent chunk per warp; only the 16 KiB position table lives in
//     shared memory, 14 chunks per SM.  The page is read from HBM through the read-only L1 path
//     (ld.global.nc): it is immutable, probes walk it forward and LZ4 candidates are mostly
//     recent, so the 128-byte lines get reused.  Staging the whole 64 KiB page in shared memory
//     would cap residency at two chunks per SM; a 2-8 KiB cp.async sliding window per warp was
//     built and measured slower than spending the same shared memory on more resident chunks
//     (DESIGN.md §4).
//   * Latency per sequence: the warp runs the reference's loop speculatively, one LZ4 sequence
//     per iteration with ONE table round trip and ONE page round trip:
//       "unified batch": lane 0 replays the table refill of position end-2 (lz4.c:691), lane 1
//         the immediate re-test at `end` (lz4.c:694-707), lanes 2.. the first 30 probes of the
//         following search (lz4.c:593-619; probe positions are a closed form of the probe index:
//         +1, then +accel for 64 probes, +accel+1 for the next 64, ...).  All are "read slot,
//         write slot, compare 4 bytes" in program order.  Every lane stores its position
//         speculatively and reads the slot back; lanes that see a foreign value share a slot
//         with another lane.  Below the lowest such lane program order is irrelevant, so if the
//         first hit (ballot) lies there it wins and the later lanes put the old values back.
//         Otherwise (a true intra-batch dependency, or 30 probes were not enough) the general
//         search lz4_search_slow resolves program order with __match_any_sync.
//       Each lane fetches 12 bytes around its probe and around its candidate in that same round
//         trip, so the winner already knows the match extension up to 4 bytes forward
//         (lz4.c:415-439) and 4 bytes backward (lz4.c:622); longer ones go out of line.
//   * The hot loop is kept small on purpose (the profile of the first version showed a third of
//     the stall samples waiting on instruction fetch): rare paths are __noinline__.
#pragma once
#include "common.cuh"
#include "fingerprint.cuh"

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

// The 12 bytes [p-4, p+8) of the page as three little-endian words {before, at, next}: four
// aligned loads (page buffers are padded past their end; the word before the page start is never
// needed because backward extension is capped by the position itself).
struct Lz4Around { uint32_t before, at, next; };
#ifndef CMB_LZ4_HINT_PROBE
#define CMB_LZ4_HINT_PROBE 0
#endif
#ifndef CMB_LZ4_HINT_CAND
#define CMB_LZ4_HINT_CAND 0
#endif
// HINT: 0 = ld.global.nc, 1 = + L1::evict_last, 2 = + L1::no_allocate, 3 = .cg, 4 = .cs, 5 = .lu,
// 6 = .nc + L1::evict_first (tuning, profiles/r1_encode_notes.md)
template <int HINT> __device__ __forceinline__ uint32_t lz4_ldw(const uint32_t *q) {
	uint32_t v;
	if (HINT == 1) asm("ld.global.nc.L1::evict_last.b32 %0, [%1];" : "=r"(v) : "l"(q));
	else if (HINT == 2) asm("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(q));
	else if (HINT == 3) v = __ldcg(q);
	else if (HINT == 4) v = __ldcs(q);
	else if (HINT == 5) v = __ldlu(q);
	else if (HINT == 6) asm("ld.global.nc.L1::evict_first.b32 %0, [%1];" : "=r"(v) : "l"(q));
	else v = __ldg(q);
	return v;
}
template <int HINT = 0>
__device__ __forceinline__ Lz4Around lz4_around(const uint8_t *src, uint32_t p) {
	const uint32_t a = p & ~3u, sh = (p & 3u) * 8u;
	const uint32_t *q = reinterpret_cast<const uint32_t *>(src + a);
	// p < 4: the word before the page does not exist and is not needed (backward extension is capped
	// by the position), so the first word is read twice instead of branching
	const uint32_t w0 = lz4_ldw<HINT>(q - (a != 0u));
	const uint32_t w1 = lz4_ldw<HINT>(q), w2 = lz4_ldw<HINT>(q + 1), w3 = lz4_ldw<HINT>(q + 2);
	Lz4Around r;
	r.before = __funnelshift_r(w0, w1, sh);
	r.at = __funnelshift_r(w1, w2, sh);
	r.next = __funnelshift_r(w2, w3, sh);
	return r;
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
	// most matches that outgrow the neighbourhoods end within the next few bytes: one byte per lane first
	{
		const uint32_t pa = a + (uint32_t)lane;
		const bool same = pa < lim && ldg8(src + pa) == ldg8(src + b + (uint32_t)lane);
		const uint32_t stop = __ballot_sync(CMB_FULL, !same);
		if (stop) return (uint32_t)(__ffs(stop) - 1);
	}
	uint32_t total = 32;
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

// The general search (any number of probes, any hash clashes).  Slot g of the search is the
// refill (g = 0) / re-test (g = 1) when g < shift, else probe g - shift; starts at slot g0.
// specials = false: the two special slots exist in the numbering but do nothing (first search of
// a chunk in the ring encoder, which keeps one lane layout for every batch).
template <bool WIDE>
__device__ __noinline__ uint64_t lz4_search_slow(const uint8_t *src, uint32_t lim4, Lz4Table<WIDE> tab,
    uint32_t anchor, uint32_t shift, uint32_t accel, uint32_t mflimit, uint32_t g0, int lane, bool specials = true) {
	const uint32_t p0 = anchor + 1;
	for (;; g0 += 32) {
		const uint32_t g = g0 + lane;
		const bool special = g < shift;
		const uint32_t k = g - shift;
		uint32_t pos = p0 + lz4_probe_off(k, accel);
		const uint32_t nxt = p0 + lz4_probe_off(k + 1, accel);
		bool en = nxt <= mflimit;
		if (special) { pos = anchor - 2u + 2u * g; en = specials; }
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
		const uint32_t commit = hits ? (0xffffffffu >> (31 - w)) & enmask : enmask;
		if ((commit >> lane) & 1u) {
			const uint32_t pc = peers & commit;
			if (31 - __clz(pc) == lane) tab.put(h, pos);     // last writer per slot wins
		}
		__syncwarp();
		if (hits)
			return lz4_pack(true, g0 + (uint32_t)w < shift, __shfl_sync(CMB_FULL, pos, w),
			    __shfl_sync(CMB_FULL, cand, w));
		if (__ballot_sync(CMB_FULL, en || special) != CMB_FULL) return 0;   // a probe ran into the end margin
	}
}

// Everything after a found match that does not fit the straight-line emitter: long literal runs,
// length bytes beyond one.  Returns the new output offset.
__device__ __noinline__ uint32_t lz4_emit_general(uint8_t *dst, uint32_t op, const uint8_t *src, uint32_t anchor,
    uint32_t lit, uint32_t off, uint32_t mc, int lane) {
	if (lane == 0) dst[op] = (uint8_t)((min(lit, 15u) << 4) | min(mc, 15u));
	op++;
	if (lit >= 15u) op = lz4_emit_len(dst, op, lit - 15u, lane);
	if (lit <= 256u) {           // the usual case here is a run of 65..200 bytes: bytes, no alignment work
		for (uint32_t i = lane; i < lit; i += 32) st_out8(dst + op + i, ldg8(src + anchor + i));
	} else {
		warp_copy_ro(dst + op, src + anchor, lit, lane);
	}
	op += lit;
	if (lane == 0) { dst[op] = (uint8_t)off; dst[op + 1] = (uint8_t)(off >> 8); }
	op += 2;
	if (mc >= 15u) op = lz4_emit_len(dst, op, mc - 15u, lane);
	return op;
}

// ---- the encoder -----------------------------------------------------------------------------

// Encodes src[0,n) into dst; returns the block length (uniform across the warp).
// `smem` is this warp's LZ4_TABLE_BYTES of shared memory.  src must be 4-byte aligned and
// readable up to 16 bytes past src+n (the library's page buffers are contiguous and padded).
// With FP the EF128 fingerprint of the page is computed along the way (EfFrontier): the parse and
// the fingerprint then read the page from HBM once, and the stripe loads prefetch the parse.
template <bool WIDE, bool FP, bool FP_NOALLOC = false>
__device__ uint32_t lz4_encode_warp(const uint8_t *__restrict__ src, uint32_t n, uint8_t *__restrict__ dst,
    uint32_t accel, uint8_t *smem, int lane, uint64_t &fp_hi, uint64_t &fp_lo) {
	Lz4Table<WIDE> tab;
	tab.t = reinterpret_cast<decltype(tab.t)>(smem);
	const uint32_t lim4 = (n + 3u) & ~3u;
	uint32_t op = 0, anchor = 0;
	EfFrontierT<FP_NOALLOC> fp;
	if (FP) fp.start(src, n, lane);

	// lz4.c:739 — table cleared per call: an untouched slot aliases position 0.
	{
		uint4 z = make_uint4(0, 0, 0, 0);
		uint4 *t4 = reinterpret_cast<uint4 *>(smem);
#pragma unroll 4
		for (uint32_t i = lane; i < LZ4_TABLE_BYTES / 16; i += 32) t4[i] = z;
	}
	__syncwarp();

	if (n >= LZ4_MIN_INPUT) {
		const uint32_t mflimit = n - LZ4_MATCH_FIND_MARGIN;
		const uint32_t mlimit = n - LZ4_TAIL_LITERALS;
		// lz4.c:583 stores position 0 under hash(0): a no-op on the cleared table, so skipped.
		// Per-lane constants of the batch that follows a match ending at `end` (lz4.c:691-710):
		// lane 0 refills end-2, lane 1 re-tests end, lane j >= 2 is probe k = j-2 of the search
		// starting at end+1, which runs only while the probe after it stays <= mflimit.
		const uint32_t kk = (uint32_t)lane - 2u;
		const int32_t delta2 = lane < 2 ? 2 * lane - 2 : (int32_t)(1u + (kk ? 1u + accel * (kk - 1u) : 0u));
		const uint32_t need2 = lane < 2 ? 0u : 2u + accel * kk;            // end + need2 <= mflimit
		// first batch of the chunk: plain search from position 1 (lz4.c:584), no refill / re-test
		uint32_t shift = 0;          // 2 once a match has ended
		for (;;) {
			if (FP) fp.upto(src, anchor + 512u, lane);     // stripes ahead of this batch's probes
			const bool special = (uint32_t)lane < shift;
			uint32_t pos;
			bool en;
			if (shift) {
				pos = anchor + (uint32_t)delta2;
				en = anchor + need2 <= mflimit;
			} else {
				pos = 1u + (lane ? 1u + accel * ((uint32_t)lane - 1u) : 0u);
				en = 2u + accel * (uint32_t)lane <= mflimit;
			}
			pos = en ? pos : 0u;                                   // keep disabled lanes' reads in range
			// speculative literal bytes: src[anchor + lane], src[anchor + 32 + lane] (used when the run is <= 64 bytes)
			const uint32_t litbyte = ldg8(src + min(anchor + lane, n - 1u));
			const uint32_t litbyte2 = ldg8(src + min(anchor + 32u + lane, n - 1u));

			// ---- unified batch ----
			const Lz4Around ai = lz4_around<CMB_LZ4_HINT_PROBE>(src, pos);
			const uint32_t pseq = ai.at;
			const uint32_t h = WIDE ? lz4_hash5((uint64_t)ai.at | ((uint64_t)ai.next << 32)) : lz4_hash4(ai.at);
			const uint32_t cand = tab.get(h);
			__syncwarp();
			if (en) tab.put(h, pos);                                // speculative commit
			__syncwarp();
			const Lz4Around ac = lz4_around<CMB_LZ4_HINT_CAND>(src, cand);   // latency overlaps the read-back
			// (Lanes that share a slot store to it in the same instruction: CUDA guarantees that one
			// of those stores lands; racecheck reports the write-write conflict, it is the mechanism.)
			const uint32_t seen = tab.get(h);
			__syncwarp();                                           // read-backs done before any undo store
			const bool foreign = en && seen != (WIDE ? pos : (pos & 0xffffu));
			const bool hit = en && !(special && lane == 0) && cand + LZ4_FAR >= pos && ac.at == pseq;
			const uint32_t foreigns = __ballot_sync(CMB_FULL, foreign);
			const uint32_t hits = __ballot_sync(CMB_FULL, hit);
			// match extension known to this lane: up to 4 bytes forward, 4 backward
			uint32_t nf, nb;
			{
				const uint32_t xf = ai.next ^ ac.next;
				nf = xf ? (uint32_t)(__ffs(xf) - 1) >> 3 : 4u;
				nf = min(nf, mlimit - min(pos + LZ4_MIN_MATCH, mlimit));
				const uint32_t xb = ai.before ^ ac.before;
				nb = xb ? (uint32_t)__clz(xb) >> 3 : 4u;
				nb = min(nb, min(pos - min(anchor, pos), cand));
				if (special) nb = 0;                               // the re-test starts a sequence as is
			}
			// lanes below the lowest lane that met a foreign value form a dependency-free prefix
			// (first hit below first foreign lane <=> lowest set bit of `hits` below that of `foreigns`)
			const uint32_t low_hit = hits & (0u - hits), low_for = foreigns & (0u - foreigns);
			uint32_t ip, match, fwd, back;
			bool retest_hit;
			// one compare: with no hit low_hit - 1 is 0xffffffff (never smaller), with no foreign lane
			// low_for - 1 is 0xffffffff (any hit is smaller)
			if (low_hit - 1u < low_for - 1u) {
				const int w = __ffs(hits) - 1;
				// put the old value back past the winner, unless the slot now holds the position
				// of a lane at or before the winner (a committed write that must stay)
				const uint32_t pos_w = __shfl_sync(CMB_FULL, pos, w);
				if (en && lane > w && !(foreign && seen <= (WIDE ? pos_w : (pos_w & 0xffffu)))) tab.put(h, cand);
				__syncwarp();
				ip = pos_w;
				match = __shfl_sync(CMB_FULL, cand, w);
				fwd = __shfl_sync(CMB_FULL, nf, w);
				back = __shfl_sync(CMB_FULL, nb, w);
				retest_hit = (uint32_t)w < shift;
				if (fwd == 4u || back == 4u) {                      // longer than the neighbourhoods show: rare
					if (fwd == 4u) fwd = 4u + lz4_count_long(src, ip + 8u, match + 8u, mlimit, lim4, lane);
					if (back == 4u && ip >= anchor + 5u && match >= 5u)
						back = 4u + lz4_catchup_long(src, ip - 4u, match - 4u, anchor, lane);
				}
			} else {
				uint64_t res = 0;
				const uint32_t enmask = __ballot_sync(CMB_FULL, en);
				if (foreigns) {
					// a lane at or before the first hit depends on an earlier lane of the batch:
					// undo everything and redo the search in program order
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
				fwd = lz4_count_long(src, ip + LZ4_MIN_MATCH, match + LZ4_MIN_MATCH, mlimit, lim4, lane);
				back = retest_hit ? 0u : lz4_catchup_long(src, ip, match, anchor, lane);
			}
			const uint32_t off = ip - match;
			const uint32_t mc = back + fwd;               // lz4.c:660 matchCode
			const uint32_t lit = ip - back - anchor;
			const uint32_t end = ip + LZ4_MIN_MATCH + fwd;

			// ---- emit: token, literal run (lz4.c:625-641), offset + match length (lz4.c:643-683) ----
			if (lit <= 64u && mc < 15u + 255u) {
				uint8_t *o = dst + op;
				const uint32_t lext = lit >= 15u, mext = mc >= 15u;
				const uint32_t hl = 1u + lext;
				if ((uint32_t)lane < lit) o[hl + lane] = (uint8_t)litbyte;
				if ((uint32_t)lane + 32u < lit) o[hl + 32u + lane] = (uint8_t)litbyte2;
				// lanes 0-4 each own one of the bytes around the literals: token, literal length
				// byte, offset low, offset high, match length byte (no branches: byte `lane` of a
				// packed word, written if the lane's bit of `owners` is set)
				const uint32_t tail = hl + lit;
				const uint32_t head4 = (min(lit, 15u) << 4) | min(mc, 15u) | (((lit - 15u) & 0xffu) << 8) | (off << 16);
				const uint32_t val = lane < 4 ? head4 >> (8u * (uint32_t)lane) : mc - 15u;
				const uint32_t at = lane < 2 ? (uint32_t)lane : tail + (uint32_t)lane - 2u;
				const uint32_t owners = 0x0du | (lext << 1) | (mext << 4);
				if ((owners >> lane) & 1u) o[at] = (uint8_t)val;
				op += tail + 2u + mext;
			} else {
				op = lz4_emit_general(dst, op, src, anchor, lit, off, mc, lane);
			}

			anchor = end;
			shift = 2;
			if (end > mflimit) break;                     // lz4.c:688
		}
	}

	// ---- last literals (lz4.c:713-729) ----
	uint32_t run = n - anchor;
	if (lane == 0) dst[op] = (uint8_t)(min(run, 15u) << 4);
	op++;
	if (run >= 15u) op = lz4_emit_len(dst, op, run - 15u, lane);
	lz4_copy_literals(dst + op, src + anchor, run, lane);
	op += run;
	if (FP) fp.finish(src, n, lane, fp_hi, fp_lo);
	return op;
}

}  // namespace cmb
