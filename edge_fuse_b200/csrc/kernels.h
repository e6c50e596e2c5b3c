// kernels.h — launch interface between the engine (host C++) and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cmb {

// One record of the HBM key table (64 bytes).  Replaces the reference's LMDB B+tree entry:
// key = FNV-1a-64 of the 16 address bytes (cachemap/filemap.c:18-24), one record per key
// (mdb_put_attr overwrite, filemap.c:143), attr = put timestamp (filemap.c:143, cachemap.c:195).
struct Slot {
	unsigned long long key;      // 0 = never used, ~0 = deleted; those two key values live in side slots
	unsigned long long addr_u;   // stored address, verified on get (filemap.c:236-240)
	unsigned long long addr_l;
	unsigned long long rec_off;  // arena offset of {24-byte data_prefix, payload} (filemap.c:140-147)
	uint32_t vlen;               // 0 = no valid record, else compressed_length + 1 (0+1 = raw page)
	uint32_t alloc;              // arena bytes reserved at rec_off
	unsigned long long ts;       // LMDB node attribute of the reference
	unsigned long long seq;      // stream order of the last put that claimed this key
	unsigned long long owner;    // 0 = record (if any) is in this GPU's arena; r+1 = key last written on rank r
};
static_assert(sizeof(Slot) == 64, "slot layout");

constexpr unsigned long long KEY_EMPTY = 0ull;
constexpr unsigned long long KEY_TOMB = ~0ull;

struct TableView {
	Slot *slots;                 // cap + 2 entries; [cap] holds key 0, [cap+1] holds key ~0
	uint64_t cap;                // power of two
	unsigned long long *entries; // live records
	unsigned long long *tombs;   // deleted main-table slots (rebuild trigger)
	unsigned long long *remote;  // keys whose newest record lives on another GPU
	uint64_t *fp;                // optional, 2 x u64 per slot {hi, lo}
	uint32_t *ckpt;              // optional, CKPT_WORDS per slot: parse checkpoints of the slot's record
};

// Parse checkpoints (lz4_decode_cta.cuh): word 0 = tag naming the record version (0 = none), word k
// (1..15) = where the first LZ4 sequence at or after k/16 of the page starts:
// block offset << CKPT_POS_BITS | (output position - k * n/16), or ~0 when no sequence starts in
// that sixteenth.  Written by the encoder (kernels.cu:ckpt_store), read by k_get_small.
constexpr uint32_t CKPT_WORDS = 16;
constexpr uint32_t CKPT_POS_BITS = 13;       // n/16 <= 8192 for pages up to 128 KiB
#ifndef CMB_GET_CKPT
#define CMB_GET_CKPT 1
#endif
__host__ __device__ inline uint32_t ckpt_tag(unsigned long long rec_off, uint32_t clen) {
	return 0x80000000u | (((uint32_t)(rec_off >> 4) ^ (clen * 0x9E3779B1u)) & 0x7fffffffu);
}

#define ARENA_SEG_SLOTS 4096u    // resident encoder warps that can own an arena segment

struct ArenaView {
	uint8_t *base;
	uint64_t size;
	unsigned long long *head;     // bump pointer
	unsigned long long *garbage;  // bytes orphaned by relocated / deleted records
	unsigned long long *dropped;  // puts dropped because the arena was full (filemap.c:154-157 analogue)
	// direct encode (large arenas): every resident encoder warp owns a segment of the arena and
	// writes blocks straight into it; seg[2w] = cursor, seg[2w+1] = end of warp slot w's segment
	unsigned long long *seg;
	uint32_t seg_bytes;           // segment size, 0 = blocks go through the stage buffer instead
};

enum LookupStatus : int32_t {
	ST_MISS = 0,
	ST_HIT = 1,
	ST_INVALID = 2,      // page number overflowed 44 bits: not counted as a request (cachemap.c:173-174)
	ST_BAD_ENTRY = 3,    // key present with another address (filemap.c:236-240)
	ST_BAD_DECODE = 4,   // decoder consumed != stored length (filemap.c:244-248)
	ST_REMOTE = 5,       // multi-GPU: the key's newest record is on another rank (status - 5 is not encoded; see owner_out)
};

struct EncodeJob {
	const uint8_t *pages;    // n chunks, `page_stride` apart, 16-byte aligned
	uint64_t page_stride;
	uint32_t nbytes;         // chunk length
	uint32_t n;
	uint32_t accel;          // 0 = store raw (cachemap.h comp_accel==0)
	uint8_t *stage;          // n x stage_stride scratch for blocks
	uint64_t stage_stride;
	int32_t *lens;           // out: block length, -1 = chunk skipped (superseded inside the batch)
	unsigned long long *rec_out; // out, optional: arena offset of the stored record per chunk (~0 = dropped)
	uint64_t *fps;           // out, optional: 2 x u64 per chunk {hi, lo}
	unsigned int *work;      // dynamic work counter (zeroed by the launcher)
	// store mode (all null/0 for codec-only use)
	const uint32_t *slot_idx; // per chunk, from the upsert kernel; 0xffffffff = invalid address
	const unsigned long long *addr; // per chunk {u,l}
	const unsigned long long *ts;   // per chunk
	unsigned long long seq0;  // sequence of chunk 0
	unsigned long long seq_stride;  // sequence step between chunks (world size when chunks are sharded round-robin)
	TableView table;
	ArenaView arena;
};

int launch_encode(const EncodeJob &job, cudaStream_t st);

struct DecodeJob {
	uint32_t n;
	uint32_t nbytes;
	uint8_t *pages;              // n x nbytes out
	int32_t *status;             // in/out (store mode) or out consumed (codec mode)
	// codec mode
	const uint8_t *blocks;       // n blocks, block_stride apart
	uint64_t block_stride;
	const int32_t *lens;
	// store mode
	const uint64_t *rec_off;     // per request, from lookup
	const uint32_t *vlen;
	const uint8_t *arena;
};
int launch_decode(const DecodeJob &job, cudaStream_t st);

// Fused small-batch get (one CTA per request): key lookup, record staged in shared memory by TMA,
// LZ4 decode shared -> shared, page written out with 16-byte stores (device memory or page-locked
// host memory).  Safe to run on its own stream while puts run on another: records are immutable
// and the record's own prefix is checked against the request.  peer[r] = base of rank r's arena
// mapped into this process (NVLink peer memory), for keys whose newest record lives on rank r.
#define GET_MAX_PEERS 16
struct GetJob {
	TableView table;
	const uint8_t *arena;
	uint64_t arena_size;
	const unsigned long long *addr;   // n x {u, l}
	const uint8_t *valid;             // optional
	uint32_t n;
	uint32_t nbytes;                  // page size, <= 65536 for this kernel
	uint8_t *out;                     // n x nbytes
	int32_t *status;                  // n
	const uint8_t *peer[GET_MAX_PEERS];
	uint64_t peer_size[GET_MAX_PEERS];
	uint4 *scratch;                   // pool_n regions of region_entries sequence descriptors (16 bytes each)
	uint32_t region_entries;
	uint32_t *pool_bits;              // bitmap of the regions in use
	uint32_t pool_n;
};
bool get_small_supports(uint32_t nbytes);
size_t get_small_smem(uint32_t nbytes);
uint32_t get_small_region_entries(uint32_t nbytes);
int get_small_residency(uint32_t nbytes);   // CTAs resident on the device at once, < 0 on error
int launch_get_small(const GetJob &job, cudaStream_t st);

int launch_fingerprint(const uint8_t *pages, uint64_t stride, uint32_t nbytes, uint32_t n,
    uint64_t *fps, cudaStream_t st);

// addr[2i],addr[2i+1] = {u,l}; valid[i] = 0 marks a rejected address (cachemap.c:160-161).
int launch_compose(const uint64_t *offset, const uint64_t *nhid, const uint32_t *genid, int pshift,
    uint32_t n, unsigned long long *addr, uint8_t *valid, unsigned long long *key, cudaStream_t st);

int launch_upsert(TableView t, const unsigned long long *addr, const uint8_t *valid, uint32_t n,
    unsigned long long seq0, unsigned long long seq_stride, uint32_t *slot_idx, cudaStream_t st);

// Multi-GPU index replication: records {addr, owner rank, seq} written on other GPUs.  Newest
// sequence per key wins; a local record that loses is retired.
int launch_import(TableView t, ArenaView a, const unsigned long long *addr, const uint32_t *owner,
    const unsigned long long *seq, const unsigned long long *loc, uint32_t n, uint32_t *slot_idx, cudaStream_t st);

// The same exchange with the records staying on the device: pack one 32-byte record per chunk of a
// put step, import all-gathered records (rows of my_rank / rows that stored nothing are skipped).
// Record = {u, l, global stream position, tail}; tail = owner rank << 56 | arena offset / 16 << 22 |
// stored length + 1 (0 = the chunk stored nothing).  The location lets another GPU read the record
// from the owner's arena over NVLink (k_get_small).
#define XREC_LEN_BITS 22
#define XREC_OFF_BITS 34
__host__ __device__ inline unsigned long long xrec_tail(uint32_t owner, unsigned long long rec_off, int32_t len) {
	return ((unsigned long long)owner << 56) | (((rec_off >> 4) & ((1ull << XREC_OFF_BITS) - 1)) << XREC_LEN_BITS) |
	    (unsigned long long)(len < 0 ? 0u : (uint32_t)len + 1u);
}
__host__ __device__ inline uint32_t xrec_owner(unsigned long long tail) { return (uint32_t)(tail >> 56); }
__host__ __device__ inline unsigned long long xrec_off(unsigned long long tail) {
	return ((tail >> XREC_LEN_BITS) & ((1ull << XREC_OFF_BITS) - 1)) << 4;
}
__host__ __device__ inline uint32_t xrec_len1(unsigned long long tail) { return (uint32_t)(tail & ((1u << XREC_LEN_BITS) - 1)); }
int launch_pack_records(const unsigned long long *addr, const int32_t *lens, const unsigned long long *rec_off, uint32_t n,
    unsigned long long seq0, unsigned long long stride, uint32_t rank, unsigned long long *out, cudaStream_t st);
int launch_import_records(TableView t, ArenaView a, const unsigned long long *rec, uint32_t n, uint32_t my_rank,
    uint32_t *slot_idx, cudaStream_t st);

int launch_lookup(TableView t, const unsigned long long *addr, const uint8_t *valid, uint32_t n,
    int32_t *status, uint64_t *rec_off, uint32_t *vlen, unsigned long long *ts_out, cudaStream_t st);
// After launch_lookup: status ST_REMOTE entries have their owner rank in rec_off[i].

int launch_unset(TableView t, ArenaView a, const unsigned long long *addr, uint32_t n, cudaStream_t st);

// Policy-equivalent of filemap_get_rand (filemap.c:264-314): first live slot at or after r.
int launch_sample(TableView t, const unsigned long long *r, uint32_t n, unsigned long long *addr_out,
    unsigned long long *ts_out, int32_t *ok, cudaStream_t st);

int launch_read_fp(TableView t, const unsigned long long *addr, uint32_t n, uint64_t *fp_out, int32_t *ok,
    cudaStream_t st);

int launch_streamgen(const uint64_t *cids, uint32_t n, uint64_t seed, uint32_t bsize, uint8_t *out,
    cudaStream_t st);

// ---- snapshot of the store (persistence of the cache directory) ----
// One entry per live local record, written by launch_export_list in arbitrary order.
struct ExportEntry {
	unsigned long long rec_off;  // arena offset of {data_prefix, payload}
	unsigned long long ts;
	unsigned long long fp_hi, fp_lo;
	uint32_t len;                // 24 + payload bytes
	uint32_t slot;
};
static_assert(sizeof(ExportEntry) == 40, "export entry layout");
int launch_export_list(TableView t, uint32_t bsize, ExportEntry *out, unsigned long long *count,
    unsigned long long max_out, cudaStream_t st);
// Restores n records {24-byte prefix, payload} lying at blob + off[i]; slot_idx from launch_upsert
// on the records' addresses (job.addr / job.ts / job.table / job.arena / job.seq0 as for a put).
int launch_restore(const EncodeJob &job, const uint8_t *blob, const unsigned long long *off,
    const uint64_t *fps, uint32_t bsize, cudaStream_t st);

// ---- arena compaction (space of deleted / outgrown records comes back) ----
// Moves n records (sorted by old offset, new offset <= old offset) down in two steps per window so
// that no record is overwritten before it has been read: gather into `bounce`, then scatter to the
// new offsets and repoint the slots.
struct MoveEntry { unsigned long long old_off, new_off; uint32_t len, slot; };
static_assert(sizeof(MoveEntry) == 24, "move entry layout");
int launch_compact_window(TableView t, ArenaView a, const MoveEntry *moves, uint32_t n, uint8_t *bounce,
    cudaStream_t st);

// Re-inserts every live slot of `from` into the (zeroed) table `to` of the same geometry.
int launch_rehash(TableView from, TableView to, cudaStream_t st);

int sm_count();

}  // namespace cmb
