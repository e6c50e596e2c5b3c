// lz4_encode_ring.cuh — the warp-per-chunk LZ4 1.8.1 encoder of lz4_encode.cuh with the page's
// sliding window staged in shared memory by TMA (sm_100a: cp.async.bulk + mbarrier).
//
// Same output bytes as lz4_encode_warp (cachemap/lz4.c:532-733 behind filemap.c:124-128); what
// changes is where the parse frontier reads the page from.  Profile of the plain kernel (round 1,
// T-class pages): 46 % of the stall samples are long-scoreboard waits on the two dependent page
// reads of a batch — the probe neighbourhoods (30 lanes x 12 bytes spread over ~350 bytes ahead of
// the anchor) and the candidate neighbourhoods (anywhere earlier in the page).  Both go through an
// L1 of ~28 KB that 14 warps share and that the fingerprint's streaming loads keep flushing, so a
// warp-wide load almost always waits for an L2 round trip (a load is as slow as its slowest lane).
//
// Here every warp owns a 1 KiB ring of the page around its parse frontier:
//   * four 256-byte buffers; buffer g & 3 holds page bytes [256 g, 256 g + 256);
//   * ONE lane issues cp.async.bulk (global -> shared, 256 B, no registers, no L1 allocation) for
//     the buffers ahead of the frontier and arms the buffer's mbarrier with the byte count; the warp
//     waits on the mbarrier (try_wait.parity) only when the frontier first enters a buffer, i.e. once
//     per 256 bytes of parse (~10 LZ4 sequences on text-like pages), by which time the copy issued
//     256-768 bytes earlier has long landed;
//   * a batch (refill + re-test + 30 probes) reads bytes [anchor - 4, anchor + 376): at most three
//     buffers, so the fourth is always free to prefetch into;
//   * the probe neighbourhoods and the speculative literal bytes then come from shared memory
//     (29-cycle LDS, conflict-free: lanes are 12 bytes = 3 banks apart), which removes the first of
//     the two page round trips from the per-sequence chain and leaves the L1 to the candidate reads
//     (the fingerprint frontier loads with L1::no_allocate).
// The ring needs 1 KiB + 4 mbarriers per warp next to the 16 KiB position table: 13 chunks per SM
// instead of 14 (one CTA of 13 warps).  It serves accel <= 12 (the reference's setting,
// edgefs.c:168; a larger accel spreads 30 probes over more than the ring holds) — other
// accelerations use the plain kernel.
#pragma once
#include "lz4_encode.cuh"

namespace cmb {

constexpr uint32_t RING_BYTES = 1024, RING_BUF = 256, RING_BUFS = 4;
constexpr uint32_t RING_AHEAD = 376;      // a batch reads page bytes [anchor - 4, anchor + RING_AHEAD)
constexpr uint32_t RING_MAX_ACCEL = 12;   // 2 + accel * 29 + 12 <= RING_AHEAD, first batch 2 + accel * 30 + 12
constexpr uint32_t RING_MBAR_BYTES = 64;  // 4 x 8-byte mbarriers, padded
constexpr uint32_t RING_WARP_SMEM = LZ4_TABLE_BYTES + RING_BYTES + RING_MBAR_BYTES;

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t arrivals) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on `bar` (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
	    "l"(src), "r"(bytes), "r"(bar)
	    : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
	uint32_t ok;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
	    : "=r"(ok)
	    : "r"(bar), "r"(parity)
	    : "memory");
	return ok != 0;
}

// One warp's window of the page it is encoding.  All members are warp-uniform and stay in
// registers: the cold part (ring_advance) is an out-of-line function that takes and returns them
// by value, so nothing of the ring state lives in local memory.
struct PageRing {
	const uint8_t *bytes;  // RING_BYTES of shared memory, 128-byte aligned
	uint32_t s_bytes;      // the same as a shared-space address
	uint32_t s_bar;        // shared-space address of the RING_BUFS mbarriers
	uint32_t issued;       // page buffers [.., issued) have been requested from TMA
	uint32_t ready;        // page buffers [.., ready) have landed and been waited for
	uint32_t parity;       // bit r: phase of ring buffer r's mbarrier that the next wait expects
};

__device__ __forceinline__ void ring_wait_buf(uint32_t s_bar, uint32_t &parity, uint32_t g) {
	const uint32_t r = g & (RING_BUFS - 1u);
	while (!mbar_try_wait(s_bar + 8u * r, (parity >> r) & 1u)) {}
	parity ^= 1u << r;
}

// once per warp, before the first page
__device__ __forceinline__ void ring_setup(PageRing &ring, uint8_t *ring_smem, uint8_t *bar_smem, int lane) {
	ring.bytes = ring_smem;
	ring.s_bytes = smem_addr(ring_smem);
	ring.s_bar = smem_addr(bar_smem);
	ring.issued = ring.ready = 0;
	ring.parity = 0;
	if (lane == 0) {
#pragma unroll
		for (uint32_t r = 0; r < RING_BUFS; r++) mbar_init(ring.s_bar + 8u * r, 1u);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncwarp();
}

__device__ __forceinline__ bool ring_needs_advance(const PageRing &ring, uint32_t g_lo, uint32_t g_hi, uint32_t nbufs) {
	return g_hi >= ring.ready || (ring.issued < nbufs && ring.issued < g_lo + RING_BUFS);
}

// Makes page buffers [g_lo, g_hi] resident (g_hi - g_lo <= 2) and requests the ones that follow,
// up to the ring's capacity.  Every buffer that was requested is waited for exactly once and in
// order, so a ring buffer is never handed to TMA again while an earlier copy into it is in flight.
// state = issued | ready << 16 | parity << 32, in and out (page sizes up to 2^20: < 2^16 buffers).
__device__ __noinline__ uint64_t ring_advance(uint32_t s_bytes, uint32_t s_bar, uint64_t state, const uint8_t *src,
    uint32_t n, uint32_t nbufs, uint32_t g_lo, uint32_t g_hi, int lane) {
	uint32_t issued = (uint32_t)state & 0xffffu, ready = (uint32_t)(state >> 16) & 0xffffu, parity = (uint32_t)(state >> 32);
	// (1) requested buffers that are needed, or that the parse has already left behind (a long match
	//     jumped over them), must have landed before their ring buffer can be reused
	const uint32_t upto = min(issued, g_hi + 1u);
	while (ready < upto) { ring_wait_buf(s_bar, parity, ready); ready++; }
	// (2) the jump went past everything requested so far: restart at g_lo
	if (issued < g_lo) { issued = g_lo; ready = g_lo; }
	// (3) request what the ring has room for: buffers below g_lo are dead, so [g_lo, g_lo + 4) fit
	const uint32_t want = min(g_lo + RING_BUFS, nbufs);
	__syncwarp();                                   // every lane is done reading the buffers being replaced
	if (lane == 0 && issued < want) {
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads before async-proxy writes
		for (uint32_t g = issued; g < want; g++) {
			const uint32_t r = g & (RING_BUFS - 1u);
			// whole 16-byte units; the last buffer of a ragged page reads < 16 bytes past its end
			// (page buffers are padded, lz4_encode_warp's contract)
			const uint32_t nb = min(RING_BUF, (n - g * RING_BUF + 15u) & ~15u);
			mbar_expect_tx(s_bar + 8u * r, nb);
			tma_load_1d(s_bytes + r * RING_BUF, src + (size_t)g * RING_BUF, nb, s_bar + 8u * r);
		}
	}
	if (issued < want) issued = want;
	// (4) what this batch reads
	while (ready <= g_hi) { ring_wait_buf(s_bar, parity, ready); ready++; }
	return (uint64_t)issued | ((uint64_t)ready << 16) | ((uint64_t)parity << 32);
}
__device__ __forceinline__ void ring_step(PageRing &ring, const uint8_t *src, uint32_t n, uint32_t nbufs, uint32_t g_lo,
    uint32_t g_hi, int lane) {
	const uint64_t st = ring_advance(ring.s_bytes, ring.s_bar,
	    (uint64_t)ring.issued | ((uint64_t)ring.ready << 16) | ((uint64_t)ring.parity << 32), src, n, nbufs, g_lo, g_hi, lane);
	ring.issued = (uint32_t)st & 0xffffu; ring.ready = (uint32_t)(st >> 16) & 0xffffu; ring.parity = (uint32_t)(st >> 32);
}
// nothing may be in flight when the page (or the kernel) ends
__device__ __forceinline__ void ring_drain(PageRing &ring) {
	while (ring.ready < ring.issued) { ring_wait_buf(ring.s_bar, ring.parity, ring.ready); ring.ready++; }
}

// The 12 page bytes [p-4, p+8) from the ring (p inside the resident window), as lz4_around.
__device__ __forceinline__ Lz4Around ring_around(const uint8_t *ring, uint32_t p) {
	const uint32_t a = p & ~3u, sh = (p & 3u) * 8u;
	const uint32_t b = a - ((a != 0u) ? 4u : 0u);        // p < 4: the word before the page is never needed
	const uint32_t w0 = *reinterpret_cast<const uint32_t *>(ring + (b & (RING_BYTES - 1u)));
	const uint32_t w1 = *reinterpret_cast<const uint32_t *>(ring + (a & (RING_BYTES - 1u)));
	const uint32_t w2 = *reinterpret_cast<const uint32_t *>(ring + ((a + 4u) & (RING_BYTES - 1u)));
	const uint32_t w3 = *reinterpret_cast<const uint32_t *>(ring + ((a + 8u) & (RING_BYTES - 1u)));
	Lz4Around r;
	r.before = __funnelshift_r(w0, w1, sh);
	r.at = __funnelshift_r(w1, w2, sh);
	r.next = __funnelshift_r(w2, w3, sh);
	return r;
}

// Encodes src[0,n) into dst; returns the block length (uniform across the warp).  Same contract as
// lz4_encode_warp, plus: accel <= RING_MAX_ACCEL, src 16-byte aligned, `ring` set up by this warp.
// One lane layout for every batch: lane 0 refills the slot of end-2 (lz4.c:691), lane 1 re-tests
// `end` (lz4.c:694-707), lane j >= 2 is probe j-2 of the search that starts at end+1
// (lz4.c:593-619).  The first search of a page (lz4.c:583-584: from position 1, nothing before it)
// is the same batch with end = 0 and the two special lanes switched off.
template <bool WIDE, bool FP, bool FP_NOALLOC>
__device__ uint32_t lz4_encode_ring(const uint8_t *__restrict__ src, uint32_t n, uint8_t *__restrict__ dst,
    uint32_t accel, uint8_t *tab_smem, PageRing &ring, int lane, uint64_t &fp_hi, uint64_t &fp_lo) {
	Lz4Table<WIDE> tab;
	tab.t = reinterpret_cast<decltype(tab.t)>(tab_smem);
	const uint32_t lim4 = (n + 3u) & ~3u;
	uint32_t op = 0, anchor = 0;
	EfFrontierT<FP_NOALLOC> fp;
	if (FP) fp.start(src, n, lane);

	// lz4.c:739 — table cleared per call: an untouched slot aliases position 0.
	{
		uint4 z = make_uint4(0, 0, 0, 0);
		uint4 *t4 = reinterpret_cast<uint4 *>(tab_smem);
#pragma unroll 4
		for (uint32_t i = lane; i < LZ4_TABLE_BYTES / 16; i += 32) t4[i] = z;
	}
	__syncwarp();

	if (n >= LZ4_MIN_INPUT) {
		const uint32_t mflimit = n - LZ4_MATCH_FIND_MARGIN;
		const uint32_t mlimit = n - LZ4_TAIL_LITERALS;
		const uint32_t nbufs = (n + RING_BUF - 1u) / RING_BUF;
		const uint8_t *const rb = ring.bytes;
		ring.issued = ring.ready = 0;
		const bool special = lane < 2;
		const uint32_t kk = (uint32_t)lane - 2u;
		const uint32_t delta2 = special ? 2u * (uint32_t)lane - 2u : 1u + (kk ? 1u + accel * (kk - 1u) : 0u);
		const uint32_t need2 = special ? 0u : 2u + accel * kk;             // enabled while end + need2 <= mflimit
		bool started = false;                                              // a match has ended (uniform)
		for (;;) {
			if (FP) fp.upto(src, anchor + 512u, lane);
			{
				const uint32_t g_lo = (max(anchor, 4u) - 4u) / RING_BUF;
				const uint32_t g_hi = min((anchor + RING_AHEAD - 1u) / RING_BUF, nbufs - 1u);
				if (ring_needs_advance(ring, g_lo, g_hi, nbufs)) ring_step(ring, src, n, nbufs, g_lo, g_hi, lane);
			}
			const bool en = special ? started : (anchor + need2 <= mflimit);
			const uint32_t pos = en ? anchor + delta2 : 0u;        // disabled lanes read (and ignore) whatever sits at ring offset 0
			// speculative literal bytes: src[anchor + lane], src[anchor + 32 + lane] (used when the run is <= 64 bytes)
			const uint32_t litbyte = rb[min(anchor + (uint32_t)lane, n - 1u) & (RING_BYTES - 1u)];
			const uint32_t litbyte2 = rb[min(anchor + 32u + (uint32_t)lane, n - 1u) & (RING_BYTES - 1u)];

			// ---- unified batch ----
			const Lz4Around ai = ring_around(rb, pos);
			const uint32_t pseq = ai.at;
			const uint32_t h = WIDE ? lz4_hash5((uint64_t)ai.at | ((uint64_t)ai.next << 32)) : lz4_hash4(ai.at);
			const uint32_t cand = tab.get(h);
			__syncwarp();
			if (en) tab.put(h, pos);                                // speculative commit
			__syncwarp();
			const Lz4Around ac = lz4_around<CMB_LZ4_HINT_CAND>(src, cand);   // latency overlaps the read-back
			const uint32_t seen = tab.get(h);
			__syncwarp();                                           // read-backs done before any undo store
			const bool foreign = en && seen != (WIDE ? pos : (pos & 0xffffu));
			const bool hit = en && lane != 0 && cand + LZ4_FAR >= pos && ac.at == pseq;
			const uint32_t foreigns = __ballot_sync(CMB_FULL, foreign);
			const uint32_t hits = __ballot_sync(CMB_FULL, hit);
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
			const uint32_t low_hit = hits & (0u - hits), low_for = foreigns & (0u - foreigns);
			uint32_t ip, match, fwd, back;
			bool retest_hit;
			if (low_hit - 1u < low_for - 1u) {
				const int w = __ffs(hits) - 1;
				const uint32_t pos_w = __shfl_sync(CMB_FULL, pos, w);
				if (en && lane > w && !(foreign && seen <= (WIDE ? pos_w : (pos_w & 0xffffu)))) tab.put(h, cand);
				__syncwarp();
				ip = pos_w;
				match = __shfl_sync(CMB_FULL, cand, w);
				fwd = __shfl_sync(CMB_FULL, nf, w);
				back = __shfl_sync(CMB_FULL, nb, w);
				retest_hit = w == 1;
				if (fwd == 4u || back == 4u) {                      // longer than the neighbourhoods show: rare
					if (fwd == 4u) fwd = 4u + lz4_count_long(src, ip + 8u, match + 8u, mlimit, lim4, lane);
					if (back == 4u && ip >= anchor + 5u && match >= 5u)
						back = 4u + lz4_catchup_long(src, ip - 4u, match - 4u, anchor, lane);
				}
			} else {
				uint64_t res = 0;
				const uint32_t enmask = __ballot_sync(CMB_FULL, en || special);
				if (foreigns) {
					if (en) tab.put(h, cand);
					__syncwarp();
					res = lz4_search_slow<WIDE>(src, lim4, tab, anchor, 2u, accel, mflimit, 0, lane, started);
				} else if (enmask == CMB_FULL) {                     // 30 probes were not enough
					res = lz4_search_slow<WIDE>(src, lim4, tab, anchor, 2u, accel, mflimit, 32, lane, started);
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
			started = true;
			if (end > mflimit) break;                     // lz4.c:688
		}
		ring_drain(ring);
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
