// kernels.cu — sm_100a kernels of the cachemap hot path and their launchers.
//
//   k_encode      fingerprint + LZ4 block encode + arena commit, one warp per chunk  (HBM: §4)
//   k_decode      table record -> LZ4 decode -> page, one warp per request
//   k_fingerprint EF128 alone
//   k_compose / k_upsert / k_lookup / k_unset / k_sample   the HBM key table
//   k_streamgen   synthetic benchmark input
#include <stdio.h>
#include <stdlib.h>
#include "kernels.h"
#include "common.cuh"
#include "fingerprint.cuh"
#include "lz4_encode.cuh"
#include "lz4_encode_ring.cuh"
#include <type_traits>
#include "lz4_decode.cuh"
#include "lz4_decode_cta.cuh"
#include "streamgen.cuh"

namespace cmb {

// ------------------------------------------------------------------------------------------
// key table primitives
// ------------------------------------------------------------------------------------------

// FNV-1a-64 over the 16 in-memory bytes of {u,l} (cachemap/uint128.h:6-21, filemap.c:18-24).
__host__ __device__ __forceinline__ unsigned long long fnv_addr(unsigned long long u, unsigned long long l) {
	unsigned long long h = 14695981039346656037ULL;
#pragma unroll
	for (int i = 0; i < 8; i++) { h = (h ^ ((u >> (8 * i)) & 0xFF)) * 0x100000001b3ULL; }
#pragma unroll
	for (int i = 0; i < 8; i++) { h = (h ^ ((l >> (8 * i)) & 0xFF)) * 0x100000001b3ULL; }
	return h;
}

// Home slot: the low bits of an FNV key are its weakest, so remix before masking.
__device__ __forceinline__ uint64_t home_slot(unsigned long long key, uint64_t cap) {
	unsigned long long z = key;
	z = (z ^ (z >> 32)) * 0xD6E8FEB86659FD93ULL;
	z ^= z >> 32;
	return z & (cap - 1);
}

__device__ __forceinline__ unsigned long long ld_key(const Slot *s) {
	return *reinterpret_cast<const volatile unsigned long long *>(&s->key);
}

// Finds the slot holding `key`, claiming one if absent.  Returns the slot index, or 0xffffffff
// when the table is full.  A deleted slot (tombstone) met on the way is reused, but only after the
// whole probe chain has been searched for the key, and by CAS, so that concurrent claimers of the
// same key inside one kernel converge on one slot: they walk the same chain, so they either agree
// on the first tombstone, or the loser of a CAS rescans and finds the winner's entry.
__device__ uint32_t table_find_or_claim(const TableView &t, unsigned long long key) {
	if (key == KEY_EMPTY) return (uint32_t)t.cap;
	if (key == KEY_TOMB) return (uint32_t)t.cap + 1;
	for (int attempt = 0; attempt < 64; attempt++) {
		uint64_t i = home_slot(key, t.cap);
		uint64_t tomb = ~0ull;
		bool retry = false;
		for (uint64_t n = 0; n < t.cap; n++, i = (i + 1) & (t.cap - 1)) {
			unsigned long long cur = ld_key(&t.slots[i]);
			if (cur == key) return (uint32_t)i;
			if (cur == KEY_TOMB) { if (tomb == ~0ull) tomb = i; continue; }
			if (cur == KEY_EMPTY) {
				if (tomb != ~0ull) {
					unsigned long long old = atomicCAS(&t.slots[tomb].key, KEY_TOMB, key);
					if (old == KEY_TOMB) { atomicAdd(t.tombs, (unsigned long long)-1ll); return (uint32_t)tomb; }
					if (old == key) return (uint32_t)tomb;
					retry = true;           // someone else took that tombstone: rescan
					break;
				}
				unsigned long long old = atomicCAS(&t.slots[i].key, KEY_EMPTY, key);
				if (old == KEY_EMPTY || old == key) return (uint32_t)i;
				// lost the empty slot to another key: keep walking from here
			}
		}
		if (retry) continue;
		if (tomb != ~0ull) {        // chain wrapped without an empty slot: still may reuse a tombstone
			unsigned long long old = atomicCAS(&t.slots[tomb].key, KEY_TOMB, key);
			if (old == KEY_TOMB) { atomicAdd(t.tombs, (unsigned long long)-1ll); return (uint32_t)tomb; }
			if (old == key) return (uint32_t)tomb;
			continue;
		}
		break;
	}
	return 0xffffffffu;
}

__device__ uint32_t table_find(const TableView &t, unsigned long long key) {
	if (key == KEY_EMPTY) return (uint32_t)t.cap;
	if (key == KEY_TOMB) return (uint32_t)t.cap + 1;
	uint64_t i = home_slot(key, t.cap);
	for (uint64_t n = 0; n < t.cap; n++, i = (i + 1) & (t.cap - 1)) {
		unsigned long long cur = ld_key(&t.slots[i]);
		if (cur == key) return (uint32_t)i;
		if (cur == KEY_EMPTY) break;
	}
	return 0xffffffffu;
}

// cachemap/cachemap.c:151-166 + filemap.c:18-24, one thread per request.
__global__ void k_compose(const uint64_t *offset, const uint64_t *nhid, const uint32_t *genid, int pshift,
    uint32_t n, unsigned long long *addr, uint8_t *valid, unsigned long long *key) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	unsigned long long page = offset[i] >> pshift;
	bool ok = (page >> 44) == 0;
	unsigned long long l = page | ((unsigned long long)genid[i] << 44);
	unsigned long long u = nhid[i];
	addr[2 * i] = u;
	addr[2 * i + 1] = l;
	valid[i] = ok;
	if (key) key[i] = fnv_addr(u, l);
}

// Claims the slot of every chunk of a put batch and records stream order: the chunk with the
// highest sequence per key is the one whose record survives (sequential last-writer-wins,
// SURVEY.md App. B rule 4).
__global__ void k_upsert(TableView t, const unsigned long long *addr, const uint8_t *valid, uint32_t n,
    unsigned long long seq0, unsigned long long seq_stride, uint32_t *slot_idx) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t idx = 0xffffffffu;
	if (!valid || valid[i]) {
		unsigned long long key = fnv_addr(addr[2 * i], addr[2 * i + 1]);
		idx = table_find_or_claim(t, key);
		if (idx != 0xffffffffu) atomicMax(&t.slots[idx].seq, seq0 + seq_stride * i);
	}
	slot_idx[i] = idx;
}

__global__ void k_lookup(TableView t, const unsigned long long *addr, const uint8_t *valid, uint32_t n,
    int32_t *status, uint64_t *rec_off, uint32_t *vlen, unsigned long long *ts_out) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int32_t st = ST_MISS;
	uint64_t off = 0;
	uint32_t vl = 0;
	unsigned long long ts = 0;
	if (valid && !valid[i]) {
		st = ST_INVALID;
	} else {
		unsigned long long u = addr[2 * i], l = addr[2 * i + 1];
		uint32_t idx = table_find(t, fnv_addr(u, l));
		if (idx != 0xffffffffu) {
			const Slot &s = t.slots[idx];
			if (s.vlen != 0) {
				if (s.addr_u == u && s.addr_l == l) {
					st = ST_HIT; off = s.rec_off; vl = s.vlen; ts = s.ts;
				} else {
					st = ST_BAD_ENTRY;
				}
			} else if (s.owner != 0 && s.addr_u == u && s.addr_l == l) {
				st = ST_REMOTE; off = s.owner - 1;
			}
		}
	}
	status[i] = st;
	if (rec_off) rec_off[i] = off;
	if (vlen) vlen[i] = vl;
	if (ts_out) ts_out[i] = ts;
}

// filemap_unset (filemap.c:188-215): delete by key, whatever address the record holds.
__global__ void k_unset(TableView t, ArenaView a, const unsigned long long *addr, uint32_t n) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t idx = table_find(t, fnv_addr(addr[2 * i], addr[2 * i + 1]));
	if (idx == 0xffffffffu) return;
	Slot &s = t.slots[idx];
	// Two requests of one batch may name the same key: let exactly one retire the record.
	uint32_t old = atomicExch(&s.vlen, 0u);
	if (old == 0) return;
	atomicAdd(t.entries, (unsigned long long)-1ll);
	atomicAdd(a.garbage, (unsigned long long)s.alloc);
	s.alloc = 0;
	s.owner = 0;
	if (idx < t.cap) {
		s.key = KEY_TOMB;
		atomicAdd(t.tombs, 1ull);
	}
}

// filemap_get_rand (filemap.c:264-314) picks the first key at or after a random 64-bit draw; the
// policy-equivalent here is the first live slot at or after a random slot.  One thread per draw
// walks at most SAMPLE_WALK slots (a table at its eviction threshold is >= 1/8 full, so this
// almost always ends within a few slots) and redraws a few times; whatever is still unresolved
// (a nearly empty table) is finished by k_sample_scan, one CTA per draw, 256 slots per step.
constexpr uint32_t SAMPLE_WALK = 128, SAMPLE_REDRAWS = 8;
__device__ __forceinline__ bool slot_live(const Slot &s) { return s.vlen != 0; }
__global__ void k_sample(TableView t, const unsigned long long *r, uint32_t n, unsigned long long *addr_out,
    unsigned long long *ts_out, int32_t *ok) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	unsigned long long draw = r[i];
	int32_t found = -1;                                   // -1 = left to k_sample_scan
	for (uint32_t a = 0; a < SAMPLE_REDRAWS && found < 0; a++) {
		const uint64_t start = home_slot(draw, t.cap);
		for (uint32_t k = 0; k < SAMPLE_WALK; k++) {
			const uint64_t j = (start + k) % (t.cap + 2);
			const Slot &s = t.slots[j];
			if (slot_live(s)) {
				addr_out[2 * i] = s.addr_u; addr_out[2 * i + 1] = s.addr_l; ts_out[i] = s.ts;
				found = 1;
				break;
			}
		}
		draw = draw * 0x9E3779B97F4A7C15ULL + 0xD1B54A32D192ED03ULL;
	}
	ok[i] = found;
}
__global__ void __launch_bounds__(256) k_sample_scan(TableView t, const unsigned long long *r, uint32_t n,
    unsigned long long *addr_out, unsigned long long *ts_out, int32_t *ok) {
	__shared__ unsigned long long first;
	for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
		if (ok[i] >= 0) continue;                         // uniform across the CTA
		const uint64_t total = t.cap + 2, start = home_slot(r[i], t.cap);
		if (threadIdx.x == 0) first = ~0ull;
		__syncthreads();
		for (uint64_t base = 0; base < total; base += blockDim.x) {
			const uint64_t k = base + threadIdx.x;
			if (k < total && slot_live(t.slots[(start + k) % total])) atomicMin(&first, (unsigned long long)k);
			__syncthreads();
			if (first != ~0ull) break;
			__syncthreads();
		}
		if (threadIdx.x == 0) {
			if (first != ~0ull) {
				const Slot &s = t.slots[(start + first) % total];
				addr_out[2 * i] = s.addr_u; addr_out[2 * i + 1] = s.addr_l; ts_out[i] = s.ts;
				ok[i] = 1;
			} else {
				ok[i] = 0;
			}
		}
		__syncthreads();
	}
}

// ------------------------------------------------------------------------------------------
// fused fingerprint -> LZ4 encode -> arena commit
// ------------------------------------------------------------------------------------------

// Points the slot at a finished record (one thread).  Order: location first, then the length that
// makes the slot valid; readers on other streams take the length and the address from the record's
// own prefix and only the location from the slot (k_get_small).
__device__ __forceinline__ void slot_publish(const EncodeJob &job, Slot &s, uint32_t i, uint32_t idx, unsigned long long off,
    uint32_t need, uint32_t clen, unsigned long long au, unsigned long long al, uint64_t fp_hi, uint64_t fp_lo) {
	// whatever checkpoints the slot has describe the record it is leaving (ckpt_store renews them)
	if (job.table.ckpt) *reinterpret_cast<volatile uint32_t *>(&job.table.ckpt[(size_t)idx * CKPT_WORDS]) = 0u;
	if (s.owner) { atomicAdd(job.table.remote, (unsigned long long)-1ll); s.owner = 0; }   // now newest here (alloc held the remote length)
	else if (s.alloc) atomicAdd(job.arena.garbage, (unsigned long long)s.alloc);         // the record this one replaces
	s.addr_u = au; s.addr_l = al;
	s.ts = job.ts ? job.ts[i] : 0;
	if (job.table.fp) { job.table.fp[2 * (size_t)idx] = fp_hi; job.table.fp[2 * (size_t)idx + 1] = fp_lo; }
	s.alloc = need;
	*reinterpret_cast<volatile unsigned long long *>(&s.rec_off) = off;
	__threadfence();
	if (s.vlen == 0) atomicAdd(job.table.entries, 1ull);
	*reinterpret_cast<volatile uint32_t *>(&s.vlen) = clen + 1u;
	if (job.rec_out) job.rec_out[i] = off;
}

// Parse checkpoints of the record just published in slot idx (whole warp; lane k holds word k, see
// lz4_encode_lean).  slot_publish has zeroed the tag and fenced; the words go in, then the tag that
// names this record version — a reader takes the words only between two equal reads of that tag.
__device__ __forceinline__ void ckpt_store(const EncodeJob &job, uint32_t idx, unsigned long long off, uint32_t clen,
    uint32_t ck, int lane) {
	if (!job.table.ckpt) return;
	uint32_t *w = job.table.ckpt + (size_t)idx * CKPT_WORDS;
	__syncwarp();
	if (lane >= 1 && lane < (int)CKPT_WORDS) *reinterpret_cast<volatile uint32_t *>(w + lane) = ck;
	__threadfence();
	__syncwarp();
	if (lane == 0) *reinterpret_cast<volatile uint32_t *>(w) = ckpt_tag(off, clen);
}

// Stores the finished block as a filemap record {data_prefix, block} (filemap.c:140-147) and
// publishes it in the key table.  Called by the whole warp; lane 0 owns the bookkeeping.
__device__ unsigned long long commit_record(const EncodeJob &job, uint32_t i, uint32_t idx, const uint8_t *payload,
    uint32_t plen, int32_t clen, bool payload_ro, uint64_t fp_hi, uint64_t fp_lo, int lane) {
	Slot &s = job.table.slots[idx];
	const uint32_t need = (24u + plen + 15u) & ~15u;
	unsigned long long off = 0;
	int ok = 1;
	if (lane == 0) {
		// Records are immutable once published and a rewrite never reuses the old record's bytes:
		// a get that runs concurrently on another stream (k_get_small) decodes either the old or the
		// new record, never a torn one (LMDB gives the reference's readers a snapshot, filemap.c:223).
		// The old bytes become garbage until the arena is compacted.
		off = atomicAdd(job.arena.head, (unsigned long long)need);
		if (off + need > job.arena.size) {
			// arena full: the put is dropped silently, as a full LMDB map drops it
			// (filemap.c:143-145,154-157).  The bump pointer is never rolled back (a rollback
			// races with allocations that succeeded in between and would hand their bytes out
			// twice): it stays saturated until cmb200_compact resets it.
			// The slot is left as it is: a record the key already has stays readable, as the
			// reference's store keeps the old value when mdb_put fails.
			atomicAdd(job.arena.dropped, 1ull);
			ok = 0;
		}
	}
	ok = __shfl_sync(CMB_FULL, ok, 0);
	if (!ok) { if (lane == 0 && job.rec_out) job.rec_out[i] = ~0ull; return ~0ull; }
	off = __shfl_sync(CMB_FULL, off, 0);
	uint8_t *rec = job.arena.base + off;
	const unsigned long long au = job.addr[2 * i], al = job.addr[2 * i + 1];
	if (lane < 6) {
		uint32_t w;
		switch (lane) {
		case 0: w = (uint32_t)au; break;
		case 1: w = (uint32_t)(au >> 32); break;
		case 2: w = (uint32_t)al; break;
		case 3: w = (uint32_t)(al >> 32); break;
		case 4: w = (uint32_t)clen; break;
		default: w = 0; break;                   // the reference leaves these 4 pad bytes unspecified
		}
		reinterpret_cast<uint32_t *>(rec)[lane] = w;
	}
	if (payload_ro) warp_copy_ro(rec + 24, payload, plen, lane);
	else warp_copy_rw(rec + 24, payload, plen, lane);
	__threadfence();                                 // the record is complete before the slot points to it
	__syncwarp();
	if (lane == 0) slot_publish(job, s, i, idx, off, need, (uint32_t)clen, au, al, fp_hi, fp_lo);
	return off;                                      // arena offset of the record (~0: dropped)
}

// Direct variant of commit_record: the block already sits in the arena at `base + 24` (this warp's
// segment cursor) and stays there.  Returns the bytes of the segment consumed.
__device__ uint32_t commit_direct(const EncodeJob &job, uint32_t i, uint32_t idx, unsigned long long base,
    uint32_t clen, uint64_t fp_hi, uint64_t fp_lo, int lane) {
	Slot &s = job.table.slots[idx];
	const uint32_t need = (24u + clen + 15u) & ~15u;
	uint8_t *r = job.arena.base + base;
	const unsigned long long au = job.addr[2 * i], al = job.addr[2 * i + 1];
	if (lane < 6) {
		const uint32_t w = lane == 0 ? (uint32_t)au : lane == 1 ? (uint32_t)(au >> 32) : lane == 2 ? (uint32_t)al
		    : lane == 3 ? (uint32_t)(al >> 32) : lane == 4 ? clen : 0u;
		reinterpret_cast<uint32_t *>(r)[lane] = w;
	}
	__threadfence();                                 // block (written by all lanes) and prefix before the slot
	__syncwarp();
	if (lane == 0) slot_publish(job, s, i, idx, base, need, clen, au, al, fp_hi, fp_lo);
	return need;
}

// ENC 0: page read through the L1.  ENC 1: parse frontier staged in a per-warp shared-memory ring
// by TMA (lz4_encode_ring.cuh); shared memory = tables | rings | mbarriers.
// FPNA: the fingerprint's streaming loads do not allocate in the L1.
// Launch bounds = the real launch shapes (2 CTAs x 7 warps, or 1 CTA x 13 warps with the ring), so
// that the register allocator may use what the SM has (146 / 157 registers per thread) instead of
// rematerialising loop invariants inside the parse loop.
constexpr int ENC_PLAIN_WARPS = 7, ENC_RING_WARPS = 13;
template <bool WIDE, int ENC, bool FPNA>
__global__ void __launch_bounds__(ENC == 1 ? ENC_RING_WARPS * 32 : ENC_PLAIN_WARPS * 32, ENC == 1 ? 1 : 2) k_encode(EncodeJob job) {
	extern __shared__ __align__(128) uint8_t smem[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t nwarps = blockDim.x >> 5;
	uint8_t *wsm = smem + (size_t)warp * LZ4_TABLE_BYTES;
	PageRing ring;
	if (ENC == 1)
		ring_setup(ring, smem + (size_t)nwarps * LZ4_TABLE_BYTES + (size_t)warp * RING_ALLOC,
		    smem + (size_t)nwarps * (LZ4_TABLE_BYTES + RING_ALLOC) + (size_t)warp * RING_MBAR_BYTES, lane);
	const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;         // resident warp slot
	const bool direct = job.slot_idx != nullptr && job.arena.seg_bytes != 0u && job.accel != 0u && gw < ARENA_SEG_SLOTS;
	// room one chunk may need while it is being encoded: prefix + a stage row (filemap.c:120 dest[bsize+1024])
	const uint32_t worst = (uint32_t)((24u + job.stage_stride + 15u) & ~15ull);
	unsigned long long seg_cur = 0, seg_end = 0;                         // lane 0's copy is the truth
	if (direct && lane == 0) { seg_cur = job.arena.seg[2 * gw]; seg_end = job.arena.seg[2 * gw + 1]; }
	for (;;) {
		uint32_t i = 0;
		if (lane == 0) i = atomicAdd(job.work, 1u);
		i = __shfl_sync(CMB_FULL, i, 0);
		if (i >= job.n) break;
		const bool store = job.slot_idx != nullptr;
		uint32_t idx = 0;
		if (store) {
			idx = job.slot_idx[i];
			// invalid address, or a later chunk of this batch rewrites the same key
			bool live = idx != 0xffffffffu && job.table.slots[idx].seq == job.seq0 + job.seq_stride * i;
			if (!live) { if (lane == 0) job.lens[i] = -1; continue; }
		}
		const uint8_t *src = job.pages + (size_t)i * job.page_stride;
		uint64_t fp_hi = 0, fp_lo = 0;
		if (job.accel == 0) {           // comp_accel == 0: raw page, compressed_length 0 (filemap.c:129-133)
			if (job.fps) {
				warp_fingerprint128(src, job.nbytes, lane, fp_hi, fp_lo);
				if (lane == 0) { job.fps[2 * (size_t)i] = fp_hi; job.fps[2 * (size_t)i + 1] = fp_lo; }
			}
			if (lane == 0) job.lens[i] = 0;
			if (store) commit_record(job, i, idx, src, job.nbytes, 0, true, fp_hi, fp_lo, lane);
			continue;
		}
		// store mode: the block only passes through the stage on its way into the arena, so each warp
		// reuses ONE stage row instead of a row per chunk; with a large arena it does not even do
		// that: the block is encoded straight into this warp's arena segment
		uint8_t *dst = job.stage + (size_t)(store ? gw : i) * job.stage_stride;
		unsigned long long base = 0;
		int in_arena = 0;
		if (direct) {
			if (lane == 0) {
				if (seg_cur + worst > seg_end) {             // segment exhausted: take the next one
					if (seg_end > seg_cur) atomicAdd(job.arena.garbage, seg_end - seg_cur);
					const unsigned long long off = atomicAdd(job.arena.head, (unsigned long long)job.arena.seg_bytes);
					if (off + job.arena.seg_bytes <= job.arena.size) { seg_cur = off; seg_end = off + job.arena.seg_bytes; }
					else { seg_cur = seg_end = 0; }       // no rollback (see commit_record): saturated until compaction
				}
				in_arena = seg_cur + worst <= seg_end;
				base = seg_cur;
			}
			in_arena = __shfl_sync(CMB_FULL, in_arena, 0);
			base = __shfl_sync(CMB_FULL, base, 0);
			if (in_arena) dst = job.arena.base + base + 24;
		}
		uint32_t clen, ck = 0xffffffffu;
		if (job.fps) {                  // fingerprint along the parse frontier: the page is read once
			clen = lz4_encode_lean<WIDE, true, FPNA, ENC == 1>(src, job.nbytes, dst, job.accel, wsm, ring, lane, fp_hi, fp_lo, ck);
			if (lane == 0) { job.fps[2 * (size_t)i] = fp_hi; job.fps[2 * (size_t)i + 1] = fp_lo; }
		} else {
			clen = lz4_encode_lean<WIDE, false, false, ENC == 1>(src, job.nbytes, dst, job.accel, wsm, ring, lane, fp_hi, fp_lo, ck);
		}
		if (lane == 0) job.lens[i] = (int32_t)clen;
		if (store) {
			__syncwarp();
			if (in_arena) {
				const uint32_t used = commit_direct(job, i, idx, base, clen, fp_hi, fp_lo, lane);
				if (lane == 0) seg_cur += used;
				ckpt_store(job, idx, base, clen, ck, lane);
			} else {
				const unsigned long long at = commit_record(job, i, idx, dst, clen, (int32_t)clen, false, fp_hi, fp_lo, lane);
				if (at != ~0ull) ckpt_store(job, idx, at, clen, ck, lane);
			}
		}
	}
	if (direct && lane == 0) { job.arena.seg[2 * gw] = seg_cur; job.arena.seg[2 * gw + 1] = seg_end; }
}

static int g_sm_count = 0;
int sm_count() {
	if (!g_sm_count) {
		int dev = 0;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
		if (g_sm_count <= 0) g_sm_count = 148;
	}
	return g_sm_count;
}

static int env_int(const char *name, int dflt, int lo, int hi) {
	const char *v = getenv(name);
	if (!v || !*v) return dflt;
	int x = atoi(v);
	return x < lo ? lo : x > hi ? hi : x;
}

template <class K>
static int launch_encode_kernel(K kern, const EncodeJob &job, int warps, int ctas_per_sm, size_t smem, cudaStream_t st) {
	CMB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	// what the tables leave of the 256 KiB per SM is L1 for the page reads; -1 = driver's choice
	static int carve = env_int("CMB200_ENC_CARVEOUT", -1, -1, 100);
	if (carve >= 0) CMB_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
	uint32_t grid = (uint32_t)(sm_count() * ctas_per_sm);
	uint32_t need = (job.n + warps - 1) / warps;
	if (grid > need) grid = need;
	kern<<<grid, warps * 32, smem, st>>>(job);
	CMB_CHECK(cudaGetLastError());
	return 0;
}

// Plain organisation: residency is bounded by shared memory, one 16 KiB position table per chunk,
// 14 of them in the 227 KiB of an SM (2 CTAs x 7 warps); chunks handed out dynamically.
static int launch_encode_warps(const EncodeJob &job, cudaStream_t st, bool fpna) {
	static int warps = env_int("CMB200_ENC_WARPS", ENC_PLAIN_WARPS, 1, ENC_PLAIN_WARPS);
	static int ctas = env_int("CMB200_ENC_CTAS_PER_SM", 2, 1, 2);
	const size_t smem = (size_t)warps * LZ4_TABLE_BYTES;
	const bool wide = job.nbytes >= LZ4_NARROW_LIMIT;
	if (wide) return fpna ? launch_encode_kernel(k_encode<true, 0, true>, job, warps, ctas, smem, st)
	                      : launch_encode_kernel(k_encode<true, 0, false>, job, warps, ctas, smem, st);
	return fpna ? launch_encode_kernel(k_encode<false, 0, true>, job, warps, ctas, smem, st)
	            : launch_encode_kernel(k_encode<false, 0, false>, job, warps, ctas, smem, st);
}

// Ring organisation (lz4_encode_ring.cuh): table + 1 KiB TMA ring + mbarriers per warp, 13 chunks
// per SM in one CTA.
static int launch_encode_ring(const EncodeJob &job, cudaStream_t st, bool fpna) {
	static int warps = env_int("CMB200_RING_WARPS", ENC_RING_WARPS, 1, ENC_RING_WARPS);
	const size_t smem = (size_t)warps * RING_WARP_SMEM;
	const bool wide = job.nbytes >= LZ4_NARROW_LIMIT;
	if (wide) return fpna ? launch_encode_kernel(k_encode<true, 1, true>, job, warps, 1, smem, st)
	                      : launch_encode_kernel(k_encode<true, 1, false>, job, warps, 1, smem, st);
	return fpna ? launch_encode_kernel(k_encode<false, 1, true>, job, warps, 1, smem, st)
	            : launch_encode_kernel(k_encode<false, 1, false>, job, warps, 1, smem, st);
}

int launch_encode(const EncodeJob &job_in, cudaStream_t st) {
	if (job_in.n == 0) return 0;
	// One encoder loop (lz4_encode_lean), two data paths with identical output:
	//   2 "ring" (default): the parse frontier staged in a per-warp shared-memory ring by TMA
	//               (lz4_encode_ring.cuh), 13 chunks per SM in one CTA;
	//   0 "plain":  the page read through the L1, 14 chunks per SM — also what accelerations above
	//               12 and unaligned page buffers use.
	static int mode = env_int("CMB200_ENC_MODE", 2, 0, 2);
	static int fpna = env_int("CMB200_FP_NOALLOC", 1, 0, 1);
	EncodeJob job = job_in;
	CMB_CHECK(cudaMemsetAsync(job.work, 0, sizeof(unsigned int), st));
	// the ring holds what 30 probes at accel <= 12 reach and TMA wants 16-byte aligned pages
	const bool ring_ok = job.accel >= 1 && job.accel <= RING_MAX_ACCEL && job.nbytes < (1u << 24) &&
	    (reinterpret_cast<uintptr_t>(job.pages) & 15u) == 0 && (job.page_stride & 15u) == 0;
	if (mode == 2 && ring_ok) return launch_encode_ring(job, st, fpna != 0);
	return launch_encode_warps(job, st, fpna != 0);
}

// ------------------------------------------------------------------------------------------
// decode
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256, 8) k_decode(DecodeJob job) {
	const int lane = threadIdx.x & 31;
	const uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (i >= job.n) return;
	uint8_t *out = job.pages + (size_t)i * job.nbytes;
	if (job.rec_off) {                                  // store mode
		if (job.status[i] != ST_HIT) return;
		const uint8_t *rec = job.arena + job.rec_off[i];
		uint32_t clen = job.vlen[i] - 1u;
		if (clen == 0) {                            // raw page (filemap.c:249-251)
			warp_copy_ro(out, rec + 24, job.nbytes, lane);
			return;
		}
		int used = lz4_decode_warp(rec + 24, clen, out, job.nbytes, lane);
		if (used != (int)clen && lane == 0) job.status[i] = ST_BAD_DECODE;   // filemap.c:244-248
	} else {
		int used = lz4_decode_warp(job.blocks + (size_t)i * job.block_stride, (uint32_t)job.lens[i], out,
		    job.nbytes, lane);
		if (lane == 0) job.status[i] = used;
	}
}

int launch_decode(const DecodeJob &job, cudaStream_t st) {
	if (job.n == 0) return 0;
	const int warps = 8;
	k_decode<<<(job.n + warps - 1) / warps, warps * 32, 0, st>>>(job);
	CMB_CHECK(cudaGetLastError());
	return 0;
}

// ------------------------------------------------------------------------------------------
// fused small-batch get: lookup + record staging (TMA) + decode in shared memory + page out
// ------------------------------------------------------------------------------------------

constexpr uint32_t GS_THREADS = DC_THREADS;       // 16 warps: one per parse section (lz4_decode_cta.cuh)
constexpr uint32_t GS_CTRL = 128 + 1152;          // control block + decoder state at the start of the shared memory
static_assert(sizeof(DecodeCta) <= 1152, "decoder state fits its slot");
constexpr uint32_t GS_MAX_PAGE = 65536;           // record buffer + page buffer must fit 227 KiB
struct GetShared {
	unsigned long long bar;                   // mbarrier of the record copy
	unsigned long long off;                   // arena offset of the record
	int32_t st;
	uint32_t clen;                            // expected compressed_length (0 = raw page)
	uint32_t owner;                           // rank + 1 when the record is in a peer's arena
	uint32_t idx;                             // slot of the key
	uint32_t region;                          // scratch region this CTA holds (~0: none)
	uint32_t sections;                        // 1 = one warp walks the block, 16 = the record's checkpoints are used
	uint32_t ck[CKPT_WORDS];
};
static_assert(sizeof(GetShared) <= 128, "control block");
__host__ __device__ inline uint32_t gs_recbuf(uint32_t nbytes) { return (24u + nbytes + 1024u + 31u) & ~15u; }
bool get_small_supports(uint32_t nbytes) { return nbytes >= 64u && nbytes <= GS_MAX_PAGE && (nbytes & 15u) == 0; }
size_t get_small_smem(uint32_t nbytes) { return GS_CTRL + gs_recbuf(nbytes) + nbytes; }
uint32_t get_small_region_entries(uint32_t nbytes) { return dc_region(nbytes); }

__device__ __forceinline__ unsigned long long ldv64(const unsigned long long *p) { return *reinterpret_cast<const volatile unsigned long long *>(p); }
__device__ __forceinline__ uint32_t ldv32(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }

// Reads the slot of {u, l}.  Readers never block writers: the location is a single 8-byte load and
// everything else about the record (address, length) is taken from the record's own prefix later.
__device__ void gs_lookup(const GetJob &job, unsigned long long u, unsigned long long l, GetShared *sh) {
	int32_t st = ST_MISS;
	uint32_t clen = 0, owner = 0;
	unsigned long long off = 0;
	const uint32_t idx = table_find(job.table, fnv_addr(u, l));
	if (idx != 0xffffffffu) {
		const Slot *s = &job.table.slots[idx];
		const uint32_t vlen = ldv32(&s->vlen);
		const unsigned long long au = ldv64(&s->addr_u), al = ldv64(&s->addr_l);
		if (vlen != 0u) {
			if (au == u && al == l) { st = ST_HIT; __threadfence(); off = ldv64(&s->rec_off); clen = vlen - 1u; }
			else st = ST_BAD_ENTRY;                              // filemap.c:236-240
		} else {
			const unsigned long long ow = ldv64(&s->owner);
			if (ow != 0ull && au == u && al == l) {
				st = ST_REMOTE; owner = (uint32_t)ow;
				__threadfence();
				off = ldv64(&s->rec_off);
				const uint32_t len1 = ldv32(&s->alloc);             // remote stored length + 1, 0 = unknown
				clen = len1 ? len1 - 1u : 0xffffffffu;
			}
		}
	}
	sh->st = st; sh->clen = clen; sh->off = off; sh->owner = owner; sh->idx = idx;
}

// Scratch regions for the sequence descriptors: one per resident CTA, handed out through a bitmap
// (the pool has as many regions as CTAs of this kernel can be resident, so a free one exists).
__device__ uint32_t gs_region_take(const GetJob &job) {
	uint32_t smid;
	asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
	const uint32_t words = (job.pool_n + 31u) / 32u;
	for (uint32_t probe = 0; probe < (1u << 22); probe++) {
		const uint32_t w = (smid + probe) % words;
		const uint32_t live = w + 1u == words && (job.pool_n & 31u) ? (1u << (job.pool_n & 31u)) - 1u : 0xffffffffu;
		const uint32_t vacant = ~ldv32(&job.pool_bits[w]) & live;
		if (!vacant) continue;
		const uint32_t b = (uint32_t)__ffs(vacant) - 1u;
		if (!((atomicOr(&job.pool_bits[w], 1u << b) >> b) & 1u)) return w * 32u + b;
	}
	return 0xffffffffu;
}
__device__ void gs_region_give(const GetJob &job, uint32_t r) { atomicAnd(&job.pool_bits[r / 32u], ~(1u << (r & 31u))); }

// The record's checkpoints (ckpt_store) -> section table.  Seqlock: tag, words, tag again; the tag
// names the record (arena offset + length), so words of another record version never pass.
__device__ void gs_sections(const GetJob &job, GetShared *sh, DecodeCta *dc, uint32_t clen, bool local) {
	const uint32_t n = job.nbytes, S = n / DC_CHAINS;
	bool use = false;
	if (local && job.table.ckpt && CMB_GET_CKPT) {
		const uint32_t *ck = job.table.ckpt + (size_t)sh->idx * CKPT_WORDS;
		const uint32_t want = ckpt_tag(sh->off, clen);
		if (ldv32(ck) == want) {
			__threadfence();
			for (uint32_t k = 1; k < CKPT_WORDS; k++) sh->ck[k] = ldv32(ck + k);
			__threadfence();
			use = ldv32(ck) == want;
		}
	}
	for (uint32_t c = 0; c < DC_CHAINS; c++) { dc->ip0[c] = 0xffffffffu; dc->op0[c] = 0; dc->op_end[c] = 0; dc->cnt[c] = 0; dc->ip1[c] = 0; dc->op1[c] = 0; }
	dc->err = 0;
	dc->ip0[0] = 0; dc->op0[0] = 0;
	uint32_t prev = 0;
	if (use) {
		for (uint32_t k = 1; k < DC_CHAINS; k++) {
			const uint32_t v = sh->ck[k];
			if (v == 0xffffffffu) continue;                      // no sequence starts in this section
			const uint32_t ip = v >> CKPT_POS_BITS, op = k * S + (v & ((1u << CKPT_POS_BITS) - 1u));
			if (ip >= clen || op >= n || op >= (k + 1u) * S || ip <= dc->ip0[prev]) { use = false; break; }
			dc->ip0[k] = ip; dc->op0[k] = op;
			dc->op_end[prev] = op;
			prev = k;
		}
	}
	if (!use) {
		for (uint32_t c = 1; c < DC_CHAINS; c++) dc->ip0[c] = 0xffffffffu;
		prev = 0;
	}
	dc->op_end[prev] = n;
	sh->sections = use ? DC_CHAINS : 1u;
}

// Do the sections add up to the serial parse?  (one thread)
__device__ bool gs_sections_fit(const DecodeCta *dc, uint32_t clen, uint32_t n) {
	if (dc->err) return false;
	uint32_t prev = 0;
	for (uint32_t c = 1; c < DC_CHAINS; c++) {
		if (dc->ip0[c] == 0xffffffffu) continue;
		if (dc->ip1[prev] != dc->ip0[c] || dc->op1[prev] != dc->op0[c]) return false;
		prev = c;
	}
	return dc->ip1[prev] == clen && dc->op1[prev] == n;          // filemap.c:244-248: consumed == compressed_length
}

#ifdef CMB_GS_TRACE            /* diagnostic builds only: cycle stamps of the phases, printed by CTA 0 */
#define GS_STAMP(k) do { if (tid == 0) stamp[k] = clock64(); } while (0)
#else
#define GS_STAMP(k) do { } while (0)
#endif
__global__ void __launch_bounds__(GS_THREADS, 1) k_get_small(GetJob job) {
	extern __shared__ __align__(128) uint8_t smem[];
#ifdef CMB_GS_TRACE
	long long stamp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
	GetShared *sh = reinterpret_cast<GetShared *>(smem);
	DecodeCta *dc = reinterpret_cast<DecodeCta *>(smem + 128);
	uint8_t *rec = smem + GS_CTRL;
	uint8_t *page = rec + gs_recbuf(job.nbytes);
	const uint32_t i = blockIdx.x, tid = threadIdx.x;
	const int lane = tid & 31;
	const uint32_t warp = tid >> 5;
	const unsigned long long u = job.addr[2 * (size_t)i], l = job.addr[2 * (size_t)i + 1];
	uint8_t *out = job.out + (size_t)i * job.nbytes;
	const uint32_t s_bar = smem_addr(&sh->bar);
	if (job.valid && !job.valid[i]) { if (tid == 0) job.status[i] = ST_INVALID; return; }
	GS_STAMP(0);
	if (tid == 0) {
		sh->region = 0xffffffffu;
		mbar_init(s_bar, 1u);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	uint32_t phase = 0;
	int32_t result = ST_MISS;
	for (int attempt = 0; attempt < 4; attempt++) {
		if (tid == 0) gs_lookup(job, u, l, sh);
		__syncthreads();
		const int32_t st = sh->st;
		const uint32_t clen = sh->clen, owner = sh->owner;
		const unsigned long long off = sh->off;
		result = st;
		if (st != ST_HIT && st != ST_REMOTE) break;
		const uint32_t plen = clen ? clen : job.nbytes;           // payload bytes (raw page when compressed_length is 0)
		const uint32_t tx = (24u + plen + 15u) & ~15u;
		const uint8_t *base = job.arena;
		uint64_t limit = job.arena_size + 256u;                  // every arena is allocated with 256 bytes of slack
		bool ok = clen <= job.nbytes + 1024u && (off & 15u) == 0;
		if (st == ST_REMOTE) {
			base = owner - 1u < GET_MAX_PEERS ? job.peer[owner - 1u] : nullptr;
			limit = owner - 1u < GET_MAX_PEERS ? job.peer_size[owner - 1u] + 256u : 0;
			ok = ok && base != nullptr && clen != 0xffffffffu;
			if (!ok) break;                                      // no path to the owner's arena: REMOTE is the answer
		}
		if (!ok || off + tx > limit) { result = ST_MISS; break; }
		if (st == ST_HIT) {
			if (tid == 0) { mbar_expect_tx(s_bar, tx); tma_load_1d(smem_addr(rec), base + off, tx, s_bar); }
			while (!mbar_try_wait(s_bar, phase)) {}
			phase ^= 1u;
		} else {
			// the owner's arena over NVLink: plain 16-byte loads, no local L2 (peer lines are not cached there)
			const uint4 *src = reinterpret_cast<const uint4 *>(base + off);
			for (uint32_t k = tid; k < tx / 16u; k += GS_THREADS) reinterpret_cast<uint4 *>(rec)[k] = __ldcg(src + k);
			__syncthreads();
		}
		// the record's own prefix decides (filemap.c:9-12): address and compressed_length
		const unsigned long long pu = *reinterpret_cast<const unsigned long long *>(rec);
		const unsigned long long pl = *reinterpret_cast<const unsigned long long *>(rec + 8);
		const uint32_t pclen = *reinterpret_cast<const uint32_t *>(rec + 16);
		if (pu != u || pl != l || pclen != clen) {
			// the slot moved on between the two reads (a put or a compaction on another stream / GPU):
			// look again; a remote location that no longer holds the record is a miss
			result = ST_MISS;
			__syncthreads();
			if (st == ST_REMOTE) break;
			continue;
		}
		if (clen == 0u) {
			// raw page (filemap.c:249-251); 8-byte granularity: rec + 24 is not 16-byte aligned
			for (uint32_t k = tid; k < job.nbytes / 8u; k += GS_THREADS)
				reinterpret_cast<unsigned long long *>(out)[k] = reinterpret_cast<const unsigned long long *>(rec + 24)[k];
			result = ST_HIT;
			break;
		}
		// ---- LZ4 block -> page, both in shared memory (lz4_decode_cta.cuh) ----
		GS_STAMP(1);
		if (tid == 0) {
			if (sh->region == 0xffffffffu) sh->region = gs_region_take(job);
			gs_sections(job, sh, dc, clen, st == ST_HIT);
		}
		__syncthreads();
		if (sh->region == 0xffffffffu) { result = ST_BAD_DECODE; break; }           // cannot happen with a pool sized to residency
		uint4 *desc = job.scratch + (size_t)sh->region * job.region_entries;
		const uint32_t stride = dc_stride(job.nbytes);
		const uint32_t blk_s = smem_addr(rec + 24), page_s = smem_addr(page);
		bool good = false;
		GS_STAMP(2);
		for (int pass = 0; pass < 2 && !good; pass++) {
			const bool many = sh->sections > 1u;
			if (dc->ip0[warp] != 0xffffffffu)
				dc_parse_chain(dc, warp, blk_s, clen, job.nbytes, desc + (size_t)warp * stride,
				    many ? stride : job.region_entries, lane);
			__syncthreads();
			good = gs_sections_fit(dc, clen, job.nbytes);
			if (good || !many) break;
			// checkpoints that do not describe this block (never seen; the record is what counts): one walk
			__syncthreads();
			if (tid == 0) gs_sections(job, sh, dc, clen, false);
			__syncthreads();
		}
		if (!good) { result = ST_BAD_DECODE; break; }             // filemap.c:244-248
		GS_STAMP(3);
#ifndef CMB_DC_SKIP_LIT       /* diagnostic builds only: what a phase costs */
		dc_literals(dc, desc, stride, blk_s, page_s, rec + 24, page, warp, lane);
#endif
		__syncthreads();
		GS_STAMP(4);
#ifndef CMB_DC_SKIP_MATCH
		if (warp == 0) dc_matches(dc, desc, stride, page_s, lane);
#endif
		__syncthreads();
		GS_STAMP(5);
		for (uint32_t k = tid; k < job.nbytes / 16u; k += GS_THREADS)
			reinterpret_cast<uint4 *>(out)[k] = reinterpret_cast<const uint4 *>(page)[k];
		GS_STAMP(6);
		result = ST_HIT;
		break;
	}
#ifdef CMB_GS_TRACE
	if (tid == 0 && blockIdx.x == 0 && stamp[6])
		printf("gs_trace clen %u sections %u: stage %lld sections %lld parse %lld literals %lld matches %lld out %lld (cycles)\n", sh->clen, sh->sections,
		    stamp[1] - stamp[0], stamp[2] - stamp[1], stamp[3] - stamp[2], stamp[4] - stamp[3], stamp[5] - stamp[4], stamp[6] - stamp[5]);
#endif
	// status may live in page-locked host memory that the caller polls: the page first, then the status
	// (every thread's stores happen before the barrier, thread 0's system-wide fence after it is cumulative)
	__syncthreads();
	if (tid == 0) {
		if (sh->region != 0xffffffffu) gs_region_give(job, sh->region);
		__threadfence_system();
		*reinterpret_cast<volatile int32_t *>(&job.status[i]) = result;
	}
}

int launch_get_small(const GetJob &job, cudaStream_t st) {
	if (job.n == 0) return 0;
	const size_t smem = get_small_smem(job.nbytes);
	static size_t configured = 0;
	if (smem > configured) {
		CMB_CHECK(cudaFuncSetAttribute(k_get_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		configured = smem;
	}
	k_get_small<<<job.n, GS_THREADS, smem, st>>>(job);
	CMB_CHECK(cudaGetLastError());
	return 0;
}

// CTAs of k_get_small that can be resident on the device at once (= scratch regions needed)
int get_small_residency(uint32_t nbytes) {
	const size_t smem = get_small_smem(nbytes);
	if (cudaFuncSetAttribute(k_get_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
	int per_sm = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_get_small, (int)GS_THREADS, smem) != cudaSuccess) return -1;
	return per_sm * sm_count();
}

// ------------------------------------------------------------------------------------------
// fingerprint alone, stream generator, small launchers
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) k_fingerprint(const uint8_t *pages, uint64_t stride, uint32_t nbytes,
    uint32_t n, uint64_t *fps) {
	const int lane = threadIdx.x & 31;
	const uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (i >= n) return;
	uint64_t hi, lo;
	warp_fingerprint128(pages + (size_t)i * stride, nbytes, lane, hi, lo);
	if (lane == 0) { fps[2 * (size_t)i] = hi; fps[2 * (size_t)i + 1] = lo; }
}

int launch_fingerprint(const uint8_t *pages, uint64_t stride, uint32_t nbytes, uint32_t n, uint64_t *fps,
    cudaStream_t st) {
	if (n == 0) return 0;
	const int warps = 8;
	k_fingerprint<<<(n + warps - 1) / warps, warps * 32, 0, st>>>(pages, stride, nbytes, n, fps);
	CMB_CHECK(cudaGetLastError());
	return 0;
}

__global__ void k_streamgen(const uint64_t *cids, uint32_t n, uint64_t seed, uint32_t bsize, uint8_t *out) {
	const uint32_t words = bsize / 8;
	const uint64_t total = (uint64_t)n * words;
	for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total;
	     g += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t c = (uint32_t)(g / words), w = (uint32_t)(g % words);
		reinterpret_cast<uint64_t *>(out)[g] = sg_chunk_word(seed, cids[c], bsize, w);
	}
}

int launch_streamgen(const uint64_t *cids, uint32_t n, uint64_t seed, uint32_t bsize, uint8_t *out,
    cudaStream_t st) {
	if (n == 0) return 0;
	k_streamgen<<<sm_count() * 8, 256, 0, st>>>(cids, n, seed, bsize, out);
	CMB_CHECK(cudaGetLastError());
	return 0;
}

#define GRID1D(n) (((n) + 255u) / 256u), 256

int launch_compose(const uint64_t *offset, const uint64_t *nhid, const uint32_t *genid, int pshift,
    uint32_t n, unsigned long long *addr, uint8_t *valid, unsigned long long *key, cudaStream_t st) {
	if (n == 0) return 0;
	k_compose<<<GRID1D(n), 0, st>>>(offset, nhid, genid, pshift, n, addr, valid, key);
	CMB_CHECK(cudaGetLastError());
	return 0;
}
// Multi-GPU import, phase 1: claim the slot and record stream order; phase 2: the newest
// sequence per key applies itself (a key may appear several times in one import).
__global__ void k_import_claim(TableView t, const unsigned long long *addr, const unsigned long long *seq,
    uint32_t n, uint32_t *slot_idx) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t idx = table_find_or_claim(t, fnv_addr(addr[2 * i], addr[2 * i + 1]));
	if (idx != 0xffffffffu) atomicMax(&t.slots[idx].seq, seq[i]);
	slot_idx[i] = idx;
}
__global__ void k_import_apply(TableView t, ArenaView a, const unsigned long long *addr, const uint32_t *owner,
    const unsigned long long *seq, const unsigned long long *loc, uint32_t n, const uint32_t *slot_idx) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t idx = slot_idx[i];
	if (idx == 0xffffffffu) return;
	Slot &s = t.slots[idx];
	if (s.seq != seq[i]) return;                 // an even newer put (local or imported) owns the key
	if (s.vlen) {                                // our local record is superseded
		atomicAdd(t.entries, (unsigned long long)-1ll);
		atomicAdd(a.garbage, (unsigned long long)s.alloc);
		s.vlen = 0; s.alloc = 0;
	}
	if (s.owner == 0) atomicAdd(t.remote, 1ull);
	s.addr_u = addr[2 * i]; s.addr_l = addr[2 * i + 1];
	s.rec_off = loc ? xrec_off(loc[i]) : 0ull; s.alloc = loc ? xrec_len1(loc[i]) : 0u;   // 0 = location unknown
	__threadfence();
	s.owner = (unsigned long long)owner[i] + 1;
}
int launch_import(TableView t, ArenaView a, const unsigned long long *addr, const uint32_t *owner,
    const unsigned long long *seq, const unsigned long long *loc, uint32_t n, uint32_t *slot_idx, cudaStream_t st) {
	if (n == 0) return 0;
	k_import_claim<<<GRID1D(n), 0, st>>>(t, addr, seq, n, slot_idx);
	CMB_CHECK(cudaGetLastError());
	k_import_apply<<<GRID1D(n), 0, st>>>(t, a, addr, owner, seq, loc, n, slot_idx);
	CMB_CHECK(cudaGetLastError());
	return 0;
}
// ---- multi-GPU exchange records, device resident ------------------------------------------------
// One 32-byte record per chunk of a put step: {u, l, global stream position, tail}; tail is
// xrec_tail(owner rank, arena offset, stored length) (kernels.h; edge_fuse_b200/sharding.py has the
// same layout).
__global__ void k_pack_records(const unsigned long long *addr, const int32_t *lens, const unsigned long long *rec_off,
    uint32_t n, unsigned long long seq0, unsigned long long stride, uint32_t rank, unsigned long long *out) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(addr + 2 * (size_t)i);
	ulonglong2 b;
	b.x = seq0 + stride * i;
	int32_t len = lens[i];
	unsigned long long off = (len >= 0 && rec_off) ? rec_off[i] : 0ull;
	if (off == ~0ull) { len = -1; off = 0; }        // the put was dropped (arena full)
	b.y = xrec_tail(rank, off, len);
	reinterpret_cast<ulonglong2 *>(out)[2 * (size_t)i] = a;
	reinterpret_cast<ulonglong2 *>(out)[2 * (size_t)i + 1] = b;
}
int launch_pack_records(const unsigned long long *addr, const int32_t *lens, const unsigned long long *rec_off, uint32_t n,
    unsigned long long seq0, unsigned long long stride, uint32_t rank, unsigned long long *out, cudaStream_t st) {
	if (n == 0) return 0;
	k_pack_records<<<GRID1D(n), 0, st>>>(addr, lens, rec_off, n, seq0, stride, rank, out);
	CMB_CHECK(cudaGetLastError());
	return 0;
}
// Import straight from all-gathered records: rows of `my_rank` and rows that stored nothing are skipped.
__global__ void k_import_claim_rec(TableView t, const unsigned long long *rec, uint32_t n, uint32_t my_rank,
    uint32_t *slot_idx) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const unsigned long long tail = rec[4 * (size_t)i + 3];
	uint32_t idx = 0xffffffffu;
	if (xrec_owner(tail) != my_rank && xrec_len1(tail) != 0u) {
		idx = table_find_or_claim(t, fnv_addr(rec[4 * (size_t)i], rec[4 * (size_t)i + 1]));
		if (idx != 0xffffffffu) atomicMax(&t.slots[idx].seq, rec[4 * (size_t)i + 2]);
	}
	slot_idx[i] = idx;
}
__global__ void k_import_apply_rec(TableView t, ArenaView a, const unsigned long long *rec, uint32_t n,
    const uint32_t *slot_idx) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t idx = slot_idx[i];
	if (idx == 0xffffffffu) return;
	Slot &s = t.slots[idx];
	if (s.seq != rec[4 * (size_t)i + 2]) return;  // an even newer put (local or imported) owns the key
	if (s.vlen) {                                 // our local record is superseded
		atomicAdd(t.entries, (unsigned long long)-1ll);
		atomicAdd(a.garbage, (unsigned long long)s.alloc);
		s.vlen = 0; s.alloc = 0;
	}
	if (s.owner == 0) atomicAdd(t.remote, 1ull);
	const unsigned long long tail = rec[4 * (size_t)i + 3];
	s.addr_u = rec[4 * (size_t)i]; s.addr_l = rec[4 * (size_t)i + 1];
	// where the record lies in the owner's arena (vlen stays 0: no local record)
	s.rec_off = xrec_off(tail); s.alloc = xrec_len1(tail);
	__threadfence();
	s.owner = (unsigned long long)xrec_owner(tail) + 1;
}
int launch_import_records(TableView t, ArenaView a, const unsigned long long *rec, uint32_t n, uint32_t my_rank,
    uint32_t *slot_idx, cudaStream_t st) {
	if (n == 0) return 0;
	k_import_claim_rec<<<GRID1D(n), 0, st>>>(t, rec, n, my_rank, slot_idx);
	CMB_CHECK(cudaGetLastError());
	k_import_apply_rec<<<GRID1D(n), 0, st>>>(t, a, rec, n, slot_idx);
	CMB_CHECK(cudaGetLastError());
	return 0;
}

int launch_upsert(TableView t, const unsigned long long *addr, const uint8_t *valid, uint32_t n,
    unsigned long long seq0, unsigned long long seq_stride, uint32_t *slot_idx, cudaStream_t st) {
	if (n == 0) return 0;
	k_upsert<<<GRID1D(n), 0, st>>>(t, addr, valid, n, seq0, seq_stride, slot_idx);
	CMB_CHECK(cudaGetLastError());
	return 0;
}
int launch_lookup(TableView t, const unsigned long long *addr, const uint8_t *valid, uint32_t n,
    int32_t *status, uint64_t *rec_off, uint32_t *vlen, unsigned long long *ts_out, cudaStream_t st) {
	if (n == 0) return 0;
	k_lookup<<<GRID1D(n), 0, st>>>(t, addr, valid, n, status, rec_off, vlen, ts_out);
	CMB_CHECK(cudaGetLastError());
	return 0;
}
int launch_unset(TableView t, ArenaView a, const unsigned long long *addr, uint32_t n, cudaStream_t st) {
	if (n == 0) return 0;
	k_unset<<<GRID1D(n), 0, st>>>(t, a, addr, n);
	CMB_CHECK(cudaGetLastError());
	return 0;
}
__global__ void k_read_fp(TableView t, const unsigned long long *addr, uint32_t n, uint64_t *fp_out, int32_t *ok) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	unsigned long long u = addr[2 * i], l = addr[2 * i + 1];
	uint32_t idx = table_find(t, fnv_addr(u, l));
	int32_t found = 0;
	if (idx != 0xffffffffu && t.slots[idx].vlen != 0 && t.slots[idx].addr_u == u && t.slots[idx].addr_l == l) {
		fp_out[2 * i] = t.fp[2 * (size_t)idx]; fp_out[2 * i + 1] = t.fp[2 * (size_t)idx + 1];
		found = 1;
	}
	ok[i] = found;
}
int launch_read_fp(TableView t, const unsigned long long *addr, uint32_t n, uint64_t *fp_out, int32_t *ok,
    cudaStream_t st) {
	if (n == 0) return 0;
	k_read_fp<<<GRID1D(n), 0, st>>>(t, addr, n, fp_out, ok);
	CMB_CHECK(cudaGetLastError());
	return 0;
}
// ---- snapshot -----------------------------------------------------------------------------
__global__ void k_export_list(TableView t, uint32_t bsize, ExportEntry *out, unsigned long long *count,
    unsigned long long max_out) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= t.cap + 2) return;
	const Slot &s = t.slots[i];
	if (s.vlen == 0 || s.owner != 0) return;        // empty / deleted / the record lives on another GPU
	if (i < t.cap && (s.key == KEY_EMPTY || s.key == KEY_TOMB)) return;
	const unsigned long long j = atomicAdd(count, 1ull);
	if (j >= max_out) return;
	ExportEntry e;
	e.rec_off = s.rec_off; e.ts = s.ts;
	e.fp_hi = t.fp ? t.fp[2 * i] : 0ull; e.fp_lo = t.fp ? t.fp[2 * i + 1] : 0ull;
	e.len = 24u + (s.vlen > 1u ? s.vlen - 1u : bsize);
	e.slot = (uint32_t)i;
	out[j] = e;
}
int launch_export_list(TableView t, uint32_t bsize, ExportEntry *out, unsigned long long *count,
    unsigned long long max_out, cudaStream_t st) {
	const uint64_t n = t.cap + 2;
	k_export_list<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(t, bsize, out, count, max_out);
	CMB_CHECK(cudaGetLastError());
	return 0;
}

__global__ void __launch_bounds__(256) k_restore(EncodeJob job, const uint8_t *blob, const unsigned long long *off,
    const uint64_t *fps, uint32_t bsize) {
	const int lane = threadIdx.x & 31;
	const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (i >= job.n) return;
	const uint32_t idx = job.slot_idx[i];
	if (idx == 0xffffffffu || job.table.slots[idx].seq != job.seq0 + job.seq_stride * i) return;
	const uint8_t *rec = blob + off[i];
	const int32_t clen = *reinterpret_cast<const int32_t *>(rec + 16);   // data_prefix.compressed_length (filemap.c:9-12); off[] is 16-aligned
	const uint32_t plen = clen > 0 ? (uint32_t)clen : bsize;
	commit_record(job, i, idx, rec + 24, plen, clen, true, fps ? fps[2 * i] : 0ull, fps ? fps[2 * i + 1] : 0ull, lane);
}
int launch_restore(const EncodeJob &job, const uint8_t *blob, const unsigned long long *off,
    const uint64_t *fps, uint32_t bsize, cudaStream_t st) {
	if (job.n == 0) return 0;
	k_restore<<<(job.n * 32 + 255) / 256, 256, 0, st>>>(job, blob, off, fps, bsize);
	CMB_CHECK(cudaGetLastError());
	return 0;
}

// ---- arena compaction ---------------------------------------------------------------------
// moves[i].new_off - moves[0].new_off is also the record's place in the bounce buffer (records are
// packed in the same order and with the same 16-byte rounding in both).
__global__ void __launch_bounds__(256) k_compact_gather(ArenaView a, const MoveEntry *moves, uint32_t n, uint8_t *bounce) {
	const int lane = threadIdx.x & 31;
	const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (i >= n) return;
	const MoveEntry m = moves[i];
	warp_copy_rw(bounce + (m.new_off - moves[0].new_off), a.base + m.old_off, m.len, lane);
}
__global__ void __launch_bounds__(256) k_compact_scatter(TableView t, ArenaView a, const MoveEntry *moves, uint32_t n,
    const uint8_t *bounce) {
	const int lane = threadIdx.x & 31;
	const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (i >= n) return;
	const MoveEntry m = moves[i];
	warp_copy_rw(a.base + m.new_off, bounce + (m.new_off - moves[0].new_off), m.len, lane);
	if (lane == 0) {
		Slot &s = t.slots[m.slot];
		s.rec_off = m.new_off;
		s.alloc = (m.len + 15u) & ~15u;
	}
}
int launch_compact_window(TableView t, ArenaView a, const MoveEntry *moves, uint32_t n, uint8_t *bounce,
    cudaStream_t st) {
	if (n == 0) return 0;
	k_compact_gather<<<(n * 32 + 255) / 256, 256, 0, st>>>(a, moves, n, bounce);
	CMB_CHECK(cudaGetLastError());
	k_compact_scatter<<<(n * 32 + 255) / 256, 256, 0, st>>>(t, a, moves, n, bounce);
	CMB_CHECK(cudaGetLastError());
	return 0;
}

// ---- table rebuild ---------------------------------------------------------------------------
// Linear probing never gives a slot back: a deleted key leaves a tombstone and a key whose put was
// dropped leaves a claimed slot without a record, so over a long run the EMPTY slots only shrink and
// miss probes walk ever longer chains.  The rebuild re-inserts what is alive (a local record or a
// remote owner) into a fresh table of the same size; everything else disappears.
__global__ void k_rehash(TableView from, TableView to) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= from.cap + 2) return;
	const Slot s = from.slots[i];
	if (s.vlen == 0 && s.owner == 0) return;
	if (i < from.cap && (s.key == KEY_EMPTY || s.key == KEY_TOMB)) return;
	const unsigned long long key = i < from.cap ? s.key : (i == from.cap ? KEY_EMPTY : KEY_TOMB);
	const uint32_t idx = table_find_or_claim(to, key);
	if (idx == 0xffffffffu) return;                     // cannot happen: same size, fewer keys
	Slot d = s;
	d.key = idx < to.cap ? key : 0ull;
	// the key word was written by the claim; copy the rest field by field so that it is not torn
	Slot &t = to.slots[idx];
	t.addr_u = d.addr_u; t.addr_l = d.addr_l; t.rec_off = d.rec_off; t.vlen = d.vlen; t.alloc = d.alloc;
	t.ts = d.ts; t.seq = d.seq; t.owner = d.owner;
	if (from.fp && to.fp) { to.fp[2 * (size_t)idx] = from.fp[2 * i]; to.fp[2 * (size_t)idx + 1] = from.fp[2 * i + 1]; }
}
int launch_rehash(TableView from, TableView to, cudaStream_t st) {
	const uint64_t n = from.cap + 2;
	k_rehash<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(from, to);
	CMB_CHECK(cudaGetLastError());
	return 0;
}

int launch_sample(TableView t, const unsigned long long *r, uint32_t n, unsigned long long *addr_out,
    unsigned long long *ts_out, int32_t *ok, cudaStream_t st) {
	if (n == 0) return 0;
	k_sample<<<GRID1D(n), 0, st>>>(t, r, n, addr_out, ts_out, ok);
	CMB_CHECK(cudaGetLastError());
	k_sample_scan<<<n < 1024u ? n : 1024u, 256, 0, st>>>(t, r, n, addr_out, ts_out, ok);
	CMB_CHECK(cudaGetLastError());
	return 0;
}

}  // namespace cmb
