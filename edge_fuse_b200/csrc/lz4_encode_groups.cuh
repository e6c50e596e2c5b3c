// lz4_encode_groups.cuh — byte-exact LZ4 1.8.1 block encoder, one 8-lane GROUP per chunk.
//
// Same output as lz4_encode.cuh (the reference's LZ4_compress_generic<notLimited, byU16|byU32,
// noDict>, cachemap/lz4.c:532-733); this is the throughput organisation of it.
//
// What the measurements of the warp-per-chunk encoder said (profiles/r1_encode_notes.md):
//   * with the 16 KiB position table in shared memory only 14 chunks fit an SM, and throughput is
//     14 / (latency of one LZ4 sequence);
//   * moving the tables to global memory (L2) lifts the residency limit, but a 32-lane batch
//     issues ~370 L2 sector requests per sequence (32 lanes x (8 page words + 3.5 table
//     accesses)), and the L2 request rate then caps the SM at ~25 GiB/s however many warps run;
//   * with only 8 of the 32 lanes active the same kernel reached 48 GiB/s: the requests of the
//     lanes past the winner are pure waste, and text-like data finds its match within the first
//     few probes.
// So: 8 lanes per chunk, four chunks per warp, tables in global memory, 24 warps per SM, and the
// bound becomes the L2 request rate at ~1/4 of the requests per sequence.
//
// One loop iteration of a group = one batch of 8 consecutive table operations of the reference
// loop, in program order: slot g of a search is the refill of end-2 (g = 0, lz4.c:691) and the
// re-test of end (g = 1, lz4.c:694-707) when a match has just ended, then probe g-2 (or g) of the
// search (lz4.c:593-619; closed-form probe positions).  Lane j takes slot g0 + j.  Program order
// inside the batch is resolved exactly with __match_any_sync on the hash: a lane whose slot was
// written by a lower lane of the batch uses that lane's position as its candidate; the first hit
// wins; lanes up to the winner commit, the highest lane per slot last.  No hit and every slot
// valid -> next batch (g0 += 8); an invalid slot without a hit before it -> last literals.
// Every lane loads the 12 bytes around its probe and its candidate, so the winner knows the
// extension up to 4 bytes each way; longer ones are finished by the group (128 bytes per step).
// Groups of a warp run the same code every iteration whatever their chunk is doing (new chunk /
// batch / tail are states of one flat loop), so divergence is bounded by one iteration.
#pragma once
#include "common.cuh"
#include "lz4_encode.cuh"

namespace cmb {

constexpr int LZ4_G = 8;                          // lanes per chunk

enum : uint32_t { GS_NEW = 0, GS_RUN = 1, GS_TAIL = 2, GS_OUT = 3 };

struct GroupCtx {
	uint32_t gmask;      // lanes of this group
	int gl;              // lane within the group
	int gbase;           // first lane of the group
};

__device__ __forceinline__ uint32_t grp_ballot(const GroupCtx &g, bool p) {
	return (__ballot_sync(g.gmask, p) >> g.gbase) & 0xffu;
}
template <class T> __device__ __forceinline__ T grp_shfl(const GroupCtx &g, T v, int src) {
	return __shfl_sync(g.gmask, v, src, LZ4_G);
}

__device__ __forceinline__ uint32_t lane_word_ro(const uint8_t *src, uint32_t p) {
	const uint32_t *q = reinterpret_cast<const uint32_t *>(src + (p & ~3u));
	return __funnelshift_r(__ldg(q), __ldg(q + 1), (p & 3u) * 8u);
}

// group-cooperative copy, source immutable (page) or written by this group (stage)
template <bool RO>
__device__ __forceinline__ void grp_copy(const GroupCtx &g, uint8_t *dst, const uint8_t *src, uint32_t len) {
	uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
	if (head > len) head = len;
	for (uint32_t i = g.gl; i < head; i += LZ4_G) dst[i] = RO ? (uint8_t)ldg8(src + i) : src[i];
	dst += head; src += head; len -= head;
	const uint32_t nvec = len >> 4;
	const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u) * 8u;
	const uint8_t *s4 = src - (reinterpret_cast<uintptr_t>(src) & 3u);
	for (uint32_t i = g.gl; i < nvec; i += LZ4_G) {
		const uint32_t *q = reinterpret_cast<const uint32_t *>(s4 + (size_t)i * 16);
		uint32_t w0, w1, w2, w3, w4;
		if (RO) { w0 = __ldg(q); w1 = __ldg(q + 1); w2 = __ldg(q + 2); w3 = __ldg(q + 3); w4 = sh ? __ldg(q + 4) : 0u; }
		else { w0 = q[0]; w1 = q[1]; w2 = q[2]; w3 = q[3]; w4 = sh ? q[4] : 0u; }
		uint4 v;
		v.x = __funnelshift_r(w0, w1, sh); v.y = __funnelshift_r(w1, w2, sh);
		v.z = __funnelshift_r(w2, w3, sh); v.w = __funnelshift_r(w3, w4, sh);
		*reinterpret_cast<uint4 *>(dst + (size_t)i * 16) = v;
	}
	const uint32_t rem = len & 15u;
	for (uint32_t i = g.gl; i < rem; i += LZ4_G)
		dst[nvec * 16 + i] = RO ? (uint8_t)ldg8(src + nvec * 16 + i) : src[nvec * 16 + i];
}

// count/255 bytes of 0xFF then count%255 (LZ4 length extension)
__device__ __forceinline__ uint32_t grp_emit_len(const GroupCtx &g, uint8_t *dst, uint32_t op, uint32_t count) {
	const uint32_t nff = count / 255u;
	for (uint32_t i = g.gl; i < nff; i += LZ4_G) dst[op + i] = 0xFF;
	if (g.gl == 0) dst[op + nff] = (uint8_t)(count - nff * 255u);
	return op + nff + 1;
}

// common prefix of src[a..) and src[b..), a side capped at lim (lz4.c:415-439), 128 bytes per step
__device__ __forceinline__ uint32_t grp_count(const GroupCtx &g, const uint8_t *src, uint32_t a, uint32_t b,
    uint32_t lim, uint32_t lim4) {
	uint32_t total = 0;
	for (;;) {
		const uint32_t pa = a + total + 16u * g.gl;
		uint32_t nb = 0;
		if (pa < lim) {
			const uint32_t avail = min(16u, lim - pa);
			const uint32_t pb = b + total + 16u * g.gl;
#pragma unroll
			for (uint32_t j = 0; j < 4; j++) {
				if (nb == 4u * j && 4u * j < avail) {
					uint32_t x = read32u(src, pa + 4u * j, lim4) ^ read32u(src, pb + 4u * j, lim4);
					nb += x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
				}
			}
			nb = min(nb, avail);
		}
		const uint32_t stop = grp_ballot(g, nb < 16u);
		if (stop) {
			const int f = __ffs(stop) - 1;
			return total + 16u * f + grp_shfl(g, nb, f);
		}
		total += 16u * LZ4_G;
	}
}

// backward extension (lz4.c:622) from (ip, match): extra steps
__device__ __forceinline__ uint32_t grp_catchup(const GroupCtx &g, const uint8_t *src, uint32_t ip, uint32_t match,
    uint32_t anchor) {
	uint32_t total = 0;
	for (;;) {
		const uint32_t k = total + g.gl + 1;
		const bool ok = ip >= anchor + k && match >= k && ldg8(src + ip - k) == ldg8(src + match - k);
		const uint32_t fail = grp_ballot(g, !ok);
		if (fail) return total + (uint32_t)(__ffs(fail) - 1);
		total += LZ4_G;
	}
}

}  // namespace cmb
