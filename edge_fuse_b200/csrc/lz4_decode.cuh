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
// The kernel is issue-bound (64 warps per SM, profiles/r2_decode_blend_summary.txt), so the loop is
// written for few instructions per sequence: the common sequence — up to 32 literals with at most one
// length byte, a match of up to 32 bytes that does not overlap itself — is one predicated byte load
// and store per lane for the literals and one for the match; the token of the NEXT sequence and the
// byte after it are requested together with the offset bytes (their position is known as soon as
// the literal length is), so a sequence costs one block round trip, not two.  Reads past `cap` are
// harmless prefetches (blocks lie in padded buffers) and every use is bounds-checked.
__device__ int lz4_decode_warp(const uint8_t *__restrict__ blk, uint32_t cap, uint8_t *out,
    uint32_t n, int lane) {
	uint32_t ip = 0, op = 0;
	if (n == 0 || cap == 0) return -1;
	uint32_t tok = ldg8(blk), b0 = ldg8(blk + 1);
	for (;;) {
		if (ip >= cap) return -1;
		uint32_t len = tok >> 4, mlen = tok & 15u;
		uint32_t lit_src = ip + 1u;
		if (len == 15u) {
			if (b0 != 255u && ip + 1u < cap) { len += b0; lit_src = ip + 2u; }
			else { uint32_t q = ip + 1u; if (!lz4_read_ext(blk, cap, q, len, lane)) return -1; lit_src = q; }
		}
		if (len > n - op || len > cap - lit_src) return -1;
		const bool last = op + len + 8u > n;                 // lz4.c:1242-1256: last literals
		if (last && op + len != n) return -1;
		if (len <= 32u) { if ((uint32_t)lane < len) out[op + lane] = (uint8_t)ldg8(blk + lit_src + lane); }
		else warp_copy_ro(out + op, blk + lit_src, len, lane);
		if (last) return (int)(lit_src + len);              // lz4.c:1339
		ip = lit_src + len; op += len;
		if (ip + 2u > cap) return -1;
		const uint32_t o0 = ldg8(blk + ip), o1 = ldg8(blk + ip + 1u), m0 = ldg8(blk + ip + 2u);
		uint32_t nip = ip + 2u + (mlen == 15u ? 1u : 0u);
		const uint32_t t1 = ldg8(blk + nip), t2 = ldg8(blk + nip + 1u);
		const uint32_t off = o0 | (o1 << 8);
		if (mlen == 15u) {
			if (m0 != 255u && ip + 2u < cap) { mlen += m0; tok = t1; b0 = t2; }
			else {
				uint32_t q = ip + 2u;
				if (!lz4_read_ext(blk, cap, q, mlen, lane)) return -1;
				nip = q; tok = ldg8(blk + nip); b0 = ldg8(blk + nip + 1u);
			}
		} else { tok = t1; b0 = t2; }
		ip = nip;
		mlen += 4u;
		if (off == 0u || off > op || op + mlen + 5u > n) return -1;   // lz4.c:1319
		__syncwarp();                                // literals above are now visible to all lanes
		const uint8_t *from = out + op - off;
		if (off >= mlen) {
			if (mlen <= 32u) { if ((uint32_t)lane < mlen) out[op + lane] = from[lane]; }
			else warp_copy_rw(out + op, from, mlen, lane);
		} else {
			for (uint32_t i = lane; i < mlen; i += 32) out[op + i] = from[i % off];
		}
		op += mlen;
		__syncwarp();
	}
}

}  // namespace cmb

namespace cmb {

// Length extension inside a shared-memory block (same rule as lz4_read_ext).
__device__ __forceinline__ bool lz4_read_ext_smem(const uint8_t *blk, uint32_t cap, uint32_t &ip, uint32_t &len, int lane) {
	for (;;) {
		const bool in = ip + lane < cap;
		const uint32_t b = in ? blk[ip + lane] : 0u;
		const uint32_t stop = __ballot_sync(CMB_FULL, !in || b != 255u);
		if (stop) {
			const int f = __ffs(stop) - 1;
			if (ip + f >= cap) return false;
			len += 255u * f + __shfl_sync(CMB_FULL, b, f);
			ip += f + 1;
			return true;
		}
		len += 255u * 32u;
		ip += 32;
	}
}

// ---- CTA decoder: block AND page in shared memory, three warps in a pipeline (k_get_small) -------
//
// One chunk's LZ4 block is a serial token chain, but only the chain is serial: the literal bytes of
// a sequence come from the block and depend on nothing, and a match copy depends only on output
// that earlier sequences (and the sequence's own literals) produced.  So the CTA splits the work
//   warp 0  PARSER   walks the tokens (two dependent 29-cycle shared-memory reads per sequence) and
//                    publishes {literal source, output position, literal length, match offset|length}
//                    into a 64-entry ring in shared memory;
//   warp 1  LITERALS copies each sequence's literal run block -> page, in order, and counts them done;
//   warp 2  MATCHES  copies each sequence's match page -> page, in order, once that sequence's
//                    literals are done (everything earlier is done by construction);
// and the stages overlap: the per-sequence cost becomes that of the slowest stage instead of their
// sum.  Everything is addressed in the shared state space (32-bit addresses, ld/st.shared): generic
// pointers cost an address-space conversion per access in these loops.  Progress counters are
// volatile shared words; a stage writes its data, __syncwarp()s and then bumps its counter — shared
// memory executes one warp's accesses in issue order, so a warp that has seen the counter sees the
// data — and counters move every few sequences, not every sequence (each warp has a scheduler of
// its own in a 4-warp CTA, so a spinning stage costs nobody an issue slot).
// Same result as LZ4_decompress_fast (lz4.c:1169-1344,1360-1363): exactly n bytes decoded, return
// value = bytes of the block consumed (lz4.c:1339), or negative for a malformed block.
constexpr uint32_t DQ = 64;                       // descriptor ring entries
#ifndef CMB_DQ_PUBLISH
#define CMB_DQ_PUBLISH 4
#endif
#ifndef CMB_DP_SLEEP
#define CMB_DP_SLEEP 0
#endif
constexpr uint32_t DQ_PUBLISH = CMB_DQ_PUBLISH;   // counters move every DQ_PUBLISH sequences
struct DecodePipe {
	uint4 q[DQ];                              // {lit_src, out_pos, lit_len, off | (match_len - 4) << 16}; off 0 = last sequence
	uint32_t parsed;                          // sequences published
	uint32_t lit_done;                        // sequences whose literals are in the page
	uint32_t mat_done;                        // sequences completely done
	int32_t result;                           // consumed (> 0) or error (< 0), written by the parser at the end
	uint32_t abort;                           // nonzero: a stage gave up (which one), everybody stops
	uint32_t lit_by[2];                       // literal warps: sequences of their parity done
	uint32_t lit_ended;                       // a literal warp has handled the last sequence
};

__device__ __forceinline__ uint32_t sld8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sst8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ uint4 sld128(uint32_t a) {
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
	return v;
}
__device__ __forceinline__ void sst128(uint32_t a, uint4 v) {
	asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
// Spin limit of the stage hand-offs: a stage that waits this long for its neighbour gives up and the
// get reports a decode error instead of hanging the CTA (cannot happen with a consistent ring; it
// bounds the damage of a bug or of corrupted shared state).
constexpr uint32_t DP_SPIN_LIMIT = 1u << 26;
__device__ __forceinline__ uint32_t sld_flag(uint32_t a) {
	uint32_t v;
	asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
#if CMB_DP_SLEEP
	__nanosleep(CMB_DP_SLEEP);              // a waiting stage backs off: its polls compete with the working stages for the shared-memory pipe
#endif
	return v;
}
__device__ __forceinline__ void sst_flag(uint32_t a, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// length extension at shared address blk + ip (LZ4 255-run rule); cap = block length
__device__ __forceinline__ bool sld_ext(uint32_t blk, uint32_t cap, uint32_t &ip, uint32_t &len, int lane) {
	for (;;) {
		const bool in = ip + lane < cap;
		const uint32_t b = in ? sld8(blk + ip + lane) : 0u;
		const uint32_t stop = __ballot_sync(CMB_FULL, !in || b != 255u);
		if (stop) {
			const int f = __ffs(stop) - 1;
			if (ip + f >= cap) return false;
			len += 255u * f + __shfl_sync(CMB_FULL, b, f);
			ip += f + 1;
			return true;
		}
		len += 255u * 32u;
		ip += 32;
	}
}

// warp 0.  dp / blk are shared-space addresses.
// One shared-memory round trip per sequence: the token and the byte after it (the literal-length
// extension when the run is 15..269 bytes) arrive together; they fix where the offset, the
// match-length extension and the NEXT token lie, so those five bytes are requested in one go and
// the next iteration starts from registers.  Longer extensions (a 0xFF byte) take the general path.
__device__ void lz4_pipe_parse(uint32_t dp, uint32_t blk, uint32_t cap, uint32_t n, int lane) {
	const uint32_t a_parsed = dp + DQ * 16u, a_mat = a_parsed + 8u, a_res = a_parsed + 12u, a_abort = a_parsed + 16u;
	uint32_t ip = 0, op = 0, s = 0, freed = 0;      // freed: sequences the match stage is known to be done with
	int32_t res = 0;
	if (n == 0 || cap == 0) res = -1;
	// the block buffer is padded, so reading a few bytes past `cap` is safe; every use is bounds-checked
	uint32_t tok = res ? 0u : sld8(blk), b0 = res ? 0u : sld8(blk + 1u);
	while (res == 0) {
		// ---- the common sequence, straight line: literal and match lengths with at most one extension
		// byte each, everything inside the block and the page, a free ring slot.  Anything else — longer
		// extensions, the last sequence, a malformed block, a full ring — is ONE rarely taken branch to
		// the general code below, which re-derives the sequence from (tok, b0, ip, op).
		{
			const uint32_t l4 = tok >> 4, m4 = tok & 15u;
			const uint32_t lx = l4 == 15u ? 1u : 0u, mx = m4 == 15u ? 1u : 0u;
			const uint32_t flen = l4 + (lx ? b0 : 0u);
			const uint32_t fsrc = ip + 1u + lx;
			const uint32_t ip2 = fsrc + flen, op2 = op + flen;           // offset bytes; output after the literals
			const uint32_t a = blk + min(ip2, cap);                       // (clamped: the rare path rejects what lies outside)
			const uint32_t o0 = sld8(a), o1 = sld8(a + 1u), m0 = sld8(a + 2u);
			const uint32_t nip = ip2 + 2u + mx;
			const uint32_t an = blk + min(nip, cap);
			const uint32_t t1 = sld8(an), t2 = sld8(an + 1u);
			const uint32_t off = o0 | (o1 << 8);
			const uint32_t fm = m4 + (mx ? m0 : 0u);
			const uint32_t op3 = op2 + fm + 4u;
			const bool rare = ip + 2u >= cap || (lx && b0 == 255u) || (mx && m0 == 255u) || nip > cap || op2 + 8u > n ||
			    off == 0u || off > op2 || op3 + 5u > n || s >= freed + DQ;
			if (!rare) {
				if (lane == 0) sst128(dp + (s % DQ) * 16u, make_uint4(fsrc, op, flen, off | (fm << 16)));
				s++;
				if (lane == 0 && (s % DQ_PUBLISH) == 0u) sst_flag(a_parsed, s);   // same lane wrote the entries: ordered
				ip = nip; op = op3; tok = t1; b0 = t2;
				continue;
			}
		}
		if (ip >= cap) { res = -1; break; }
		uint32_t len = tok >> 4, mlen = tok & 15u;
		uint32_t lit_src = ip + 1u;
		if (len == 15u) {
			if (b0 != 255u && ip + 1u < cap) { len += b0; lit_src = ip + 2u; }
			else { uint32_t q = ip + 1u; if (!sld_ext(blk, cap, q, len, lane)) { res = -1; break; } lit_src = q; }
		}
		const bool last = op + len + 8u > n;                 // lz4.c:1242-1256: last literals
		if (len > n - op || len > cap - lit_src || (last && op + len != n)) { res = -1; break; }
		const uint32_t out_pos = op;
		uint32_t off = 0;
		ip = lit_src + len; op += len;
		if (!last) {
			if (ip + 2u > cap) { res = -1; break; }
			// offset, match-length extension, and the next token with its follower: one round trip
			const uint32_t o0 = sld8(blk + ip), o1 = sld8(blk + ip + 1u), m0 = sld8(blk + ip + 2u);
			uint32_t nip = ip + 2u + (mlen == 15u ? 1u : 0u);
			const uint32_t t1 = sld8(blk + nip), t2 = sld8(blk + nip + 1u);
			off = o0 | (o1 << 8);
			if (mlen == 15u) {
				if (m0 != 255u && ip + 2u < cap) { mlen += m0; tok = t1; b0 = t2; }
				else {
					uint32_t q = ip + 2u;
					if (!sld_ext(blk, cap, q, mlen, lane)) { res = -1; break; }
					nip = q; tok = sld8(blk + nip); b0 = sld8(blk + nip + 1u);
				}
			} else { tok = t1; b0 = t2; }
			ip = nip;
			if (off == 0u || off > op || op + mlen + 9u > n) { res = -1; break; }   // lz4.c:1319: op + (mlen + 4) + 5 > n
			op += mlen + 4u;
		} else {
			mlen = 0;
		}
		// the ring slot is free once the match stage is done with the sequence DQ entries back
		for (uint32_t spin = 0; s >= freed + DQ; spin++) {
			freed = sld_flag(a_mat);
			if (spin > DP_SPIN_LIMIT || sld_flag(a_abort)) { res = -101; break; }
		}
		if (res) break;
		if (lane == 0) sst128(dp + (s % DQ) * 16u, make_uint4(lit_src, out_pos, len, off | (mlen << 16)));
		s++;
		if (last) { res = (int32_t)ip; break; }              // lz4.c:1339: bytes consumed
		if ((s % DQ_PUBLISH) == 0u) { __syncwarp(); if (lane == 0) sst_flag(a_parsed, s); }
	}
	if (res < 0) {
		// malformed: a terminating entry stops the other stages (and the abort word, should they be stuck)
		if (lane == 0) sst_flag(a_abort, 1u);
		for (uint32_t spin = 0; s >= freed + DQ && spin < 1024u; spin++) freed = sld_flag(a_mat);
		if (lane == 0) sst128(dp + (s % DQ) * 16u, make_uint4(0, 0, 0, 0));
		s++;
	}
	__syncwarp();
	if (lane == 0) { sst_flag(a_res, (uint32_t)res); sst_flag(a_parsed, s); }
}

// warps 1 and 3: literal runs of the even / odd sequences (which = 0 / 1); progress counter per warp
// at a_lit[which] = number of ITS sequences done.
__device__ void lz4_pipe_literals(uint32_t dp, uint32_t blk, uint32_t out, const uint8_t *blk_g, uint8_t *out_g,
    uint32_t which, int lane) {
	const uint32_t a_parsed = dp + DQ * 16u, a_lit = a_parsed + 20u + 4u * which, a_abort = a_parsed + 16u;
	uint32_t avail = 0;
	for (uint32_t s = which, mine = 0;; s += 2u, mine++) {
		bool ended = false;
		for (uint32_t spin = 0; avail <= s; spin++) {
			avail = sld_flag(a_parsed);
			// the other literal warp may have seen the last sequence: then nothing more is coming
			// (look at the count again AFTER the flag: the final count was published before it)
			if (avail <= s && sld_flag(a_parsed + 28u)) { avail = sld_flag(a_parsed); if (avail <= s) { ended = true; break; } }
			if (spin > DP_SPIN_LIMIT) { if (lane == 0) sst_flag(a_abort, 2u); return; }
		}
		if (ended) return;
		const uint4 d = sld128(dp + (s % DQ) * 16u);
#ifndef CMB_DP_SKIP_LIT      /* diagnostic builds only: the stage consumes its descriptors without copying */
		if (d.z <= 32u) {
			if ((uint32_t)lane < d.z) sst8(out + d.y + (uint32_t)lane, sld8(blk + d.x + (uint32_t)lane));
		} else if (d.z < 256u) {
			for (uint32_t k = lane; k < d.z; k += 32u) sst8(out + d.y + k, sld8(blk + d.x + k));
		} else {
			warp_copy_rw(out_g + d.y, blk_g + d.x, d.z, lane);   // long runs: 16 bytes per lane per step
		}
#endif
		const bool last = (d.w & 0xffffu) == 0u;
		__syncwarp();
		if (lane == 0) { sst_flag(a_lit, mine + 1u); if (last) sst_flag(a_parsed + 28u, 1u); }
		if (last) return;                                    // last sequence (or the parser's stop entry)
	}
}

// warp 2
__device__ void lz4_pipe_matches(uint32_t dp, uint32_t out, int lane) {
	const uint32_t a_parsed = dp + DQ * 16u, a_lit = a_parsed + 20u, a_mat = a_parsed + 8u, a_abort = a_parsed + 16u;
	uint32_t avail = 0, lits[2] = {0, 0};
	for (uint32_t s = 0;; s++) {
		for (uint32_t spin = 0; avail <= s; spin++) {
			avail = sld_flag(a_parsed);
			if (spin > DP_SPIN_LIMIT) { if (lane == 0) sst_flag(a_abort, 3u); return; }
		}
		const uint4 d = sld128(dp + (s % DQ) * 16u);
		const uint32_t off = d.w & 0xffffu, len = (d.w >> 16) + 4u;
		// this sequence's literals and all earlier ones are in the page: both literal warps have passed it
		const uint32_t need0 = (s >> 1) + 1u, need1 = (s + 1u) >> 1;      // even sequences <= s, odd sequences <= s
		for (uint32_t spin = 0; lits[0] < need0 || lits[1] < need1; spin++) {
			lits[0] = sld_flag(a_lit); lits[1] = sld_flag(a_lit + 4u);
			if (spin > DP_SPIN_LIMIT || sld_flag(a_abort)) { if (lane == 0) sst_flag(a_abort, 4u); return; }
		}
		if (off == 0u) { if (lane == 0) sst_flag(a_mat, s + 1u); return; }
		const uint32_t to = out + d.y + d.z, from = to - off;
#ifndef CMB_DP_SKIP_MATCH    /* diagnostic builds only */
		if (off >= len) {
			if (len <= 32u) { if ((uint32_t)lane < len) sst8(to + (uint32_t)lane, sld8(from + (uint32_t)lane)); }
			else for (uint32_t k = lane; k < len; k += 32u) sst8(to + k, sld8(from + k));
		} else {
			// overlapping match = periodic extension of the off bytes before it: lanes only read bytes
			// that were complete before this sequence began.  k mod off is carried along (k grows by 32).
			uint32_t r = (uint32_t)lane % off;
			const uint32_t step = 32u % off;
			for (uint32_t k = lane; k < len; k += 32u) {
				sst8(to + k, sld8(from + r));
				r += step;
				if (r >= off) r -= off;
			}
		}
#endif
		__syncwarp();                                        // the next match may read what this one wrote
		if (((s + 1u) % DQ_PUBLISH) == 0u && lane == 0) sst_flag(a_mat, s + 1u);
	}
}

}  // namespace cmb
