/*
 * cachemap_b200.h — C ABI of the B200 cachemap engine (batch extension + kernel-level entries).
 *
 * The drop-in surface is include/cachemap.h + include/filemap.h (same prototypes as the
 * reference's cachemap/cachemap.h:33-47 and cachemap/filemap.h:19-29).  This header is the layer
 * under it: batched put/get over many chunks per call (the reference's API moves one page per
 * call — edgefs.c:1165,1191,1224 — which cannot feed a GPU), plus direct entry
 * points to the individual kernels so that parity tests can compare each one with the oracle.
 *
 * Plain C: pointers and sizes only, no CUDA or torch types.  "host" pointers may be pageable or
 * page-locked (cmb200_host_alloc gives page-locked memory; transfers from it run at full PCIe
 * rate).  "dev" pointers are device addresses in the engine's CUDA device.
 * All functions return 0 on success and -1 on failure unless stated; cmb200_last_error() then
 * describes the failure.  There is no CPU fallback anywhere: without a usable CUDA device every
 * entry point fails.
 */
#ifndef CACHEMAP_B200_H
#define CACHEMAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cmb200_engine cmb200_engine;

/* 16-byte page address, {u = nhid_small, l = page | genid << 44} (cachemap/uint128.h:4,
 * cachemap/cachemap.c:151-166). */
typedef struct { uint64_t u; uint64_t l; } cmb200_addr;

#define CMB200_FINGERPRINT 1u   /* compute + keep the EF128 content fingerprint of every put */

typedef struct cmb200_config {
	int device;             /* CUDA ordinal, -1 = current device */
	int pshift;             /* page shift: chunk = 1 << pshift bytes (edgefs -p, edgefs.c:2043-2048) */
	int accel;              /* LZ4 acceleration, 0 = store raw (cachemap_create comp_accel) */
	uint64_t capacity;      /* entries before eviction starts (cachemap_create capacity) */
	uint64_t arena_bytes;   /* HBM arena for records, 0 = sized from capacity and free memory */
	uint64_t table_slots;   /* key-table slots (power of two), 0 = 4 x capacity rounded up */
	uint32_t max_batch;     /* chunks per kernel launch, 0 = 4096 (host pages are pipelined in
	                         * steps of min(max_batch, CMB200_HOST_BATCH = 4096) chunks) */
	uint32_t flags;
} cmb200_config;

/* per-request result of a get */
enum {
	CMB200_MISS = 0,
	CMB200_HIT = 1,
	CMB200_INVALID = 2,     /* address rejected: not counted as a request (cachemap.c:173-174) */
	CMB200_BAD_ENTRY = 3,   /* key present under another address: a miss (filemap.c:236-240) */
	CMB200_BAD_DECODE = 4,  /* decoded length != stored length: a miss (filemap.c:244-248) */
	CMB200_REMOTE = 5       /* multi-GPU: the newest record of this key lives on another rank */
};

const char *cmb200_last_error(void);
int cmb200_device_count(void);

cmb200_engine *cmb200_engine_create(const cmb200_config *cfg);
void cmb200_engine_destroy(cmb200_engine *e);

void *cmb200_host_alloc(size_t bytes);        /* page-locked host memory */
void cmb200_host_free(void *p);
void *cmb200_dev_alloc(cmb200_engine *e, size_t bytes);
void cmb200_dev_free(cmb200_engine *e, void *p);
int cmb200_memcpy_h2d(cmb200_engine *e, void *dev, const void *host, size_t bytes);
int cmb200_memcpy_d2h(cmb200_engine *e, void *host, const void *dev, size_t bytes);
void *cmb200_stream(cmb200_engine *e);        /* the engine's compute cudaStream_t, for event timing */
int cmb200_sync(cmb200_engine *e);

/* filemap_set for n chunks (cachemap/filemap.c:112-158): pages = n x (1<<pshift) bytes.
 * valid (nullable) = per-chunk flag, 0 skips the chunk (rejected address).  ts (nullable) = the
 * LMDB attribute.  Chunks are applied in array order: a later chunk with the same key wins.
 * lens_out (nullable, host) receives each stored compressed_length, or -1 for skipped chunks. */
int cmb200_put_batch(cmb200_engine *e, size_t n, const cmb200_addr *addr, const uint8_t *valid,
    const void *pages_host, const uint64_t *ts, int32_t *lens_out);
int cmb200_put_batch_dev(cmb200_engine *e, size_t n, const cmb200_addr *addr, const uint8_t *valid,
    const void *pages_dev, const uint64_t *ts, int32_t *lens_out);

/* Write-behind form of cmb200_put_batch, the batch analogue of cachemap_put_async
 * (cachemap/cachemap.c:199-216): returns as soon as addr / valid / pages_host / ts have crossed to
 * the device and may be reused by the caller, like the borrowed page of cachemap_put; the encode of
 * the last sub-batch completes behind *ticket.  Every later call on this engine (gets included)
 * is ordered after the put, so waiting is needed only before reading lens_out, which must then be
 * page-locked (cmb200_host_alloc) and stay valid until cmb200_wait(ticket) has returned.
 * Submitting batch k+1 before waiting for batch k overlaps its host-to-device copy with the tail
 * of batch k's encode. */
int cmb200_put_batch_async(cmb200_engine *e, size_t n, const cmb200_addr *addr, const uint8_t *valid,
    const void *pages_host, const uint64_t *ts, int32_t *lens_out, uint64_t *ticket);
int cmb200_wait(cmb200_engine *e, uint64_t ticket);

/* filemap_get for n requests (cachemap/filemap.c:217-262).  status_out[i] is one of CMB200_*;
 * pages_out receives 1<<pshift bytes per request (untouched for non-hits). */
int cmb200_get_batch(cmb200_engine *e, size_t n, const cmb200_addr *addr, const uint8_t *valid,
    void *pages_out_host, int32_t *status_out);
int cmb200_get_batch_dev(cmb200_engine *e, size_t n, const cmb200_addr *addr, const uint8_t *valid,
    void *pages_out_dev, int32_t *status_out);

/* filemap_unset (filemap.c:188-215), filemap_entries (filemap.c:316-330),
 * filemap_get_rand (filemap.c:264-314; policy-equivalent: first live record at or after r). */
int cmb200_unset_batch(cmb200_engine *e, size_t n, const cmb200_addr *addr);
uint64_t cmb200_entries(cmb200_engine *e);
int cmb200_sample(cmb200_engine *e, size_t n, const uint64_t *r, cmb200_addr *addr_out,
    uint64_t *ts_out, int32_t *ok_out);

/* Copies the stored record of each address — the bytes the reference keeps in LMDB:
 * 24-byte data_prefix + payload (filemap.c:9-12,140-147) — into out (stride bytes apart) and its
 * total length into len_out (-1 = absent). */
int cmb200_read_records(cmb200_engine *e, size_t n, const cmb200_addr *addr, void *out_host,
    size_t stride, int32_t *len_out);
/* EF128 of the stored records (engine created with CMB200_FINGERPRINT): fp_out[2i]=hi, [2i+1]=lo. */
int cmb200_read_fingerprints(cmb200_engine *e, size_t n, const cmb200_addr *addr, uint64_t *fp_out,
    int32_t *ok_out);

typedef struct cmb200_stats {
	uint64_t entries, table_slots, tombstones;
	uint64_t arena_bytes, arena_used, arena_garbage, dropped_puts;
	uint64_t remote_entries;  /* keys whose newest record is on another GPU (multi-GPU index) */
	uint64_t put_chunks, get_requests, get_hits, kernel_launches;
	/* summed CUDA-event durations of the encode / decode kernel launches (last 64 per call) */
	uint64_t encode_kernel_ns, encode_kernel_launches, decode_kernel_ns, decode_kernel_launches;
	uint64_t fingerprint_kernel_ns;
} cmb200_stats;
int cmb200_get_stats(cmb200_engine *e, cmb200_stats *out);

/* One put step of a sharded stream with everything that follows it kept on the device and
 * asynchronous.  Like cmb200_put_batch_async (pages on the host, pages_on_dev = 0), or with the
 * pages in page-locked host memory that the caller leaves untouched until the ticket is done
 * (pages_on_dev = 2: the call returns without waiting for its own copies, so the next step can be
 * queued behind them at once), or device resident (pages_on_dev = 1, same lifetime rule),
 * and additionally writes one 32-byte exchange record per chunk — {u, l, global stream position,
 * rank << 32 | stored length (negative: nothing stored)} — to records_dev_out (device memory,
 * n x 32 bytes) on the engine's stream.  The caller all-gathers those records (NCCL, on that
 * stream) and hands the result to cmb200_import_records_dev, which imports the rows of the other
 * ranks into the index replica, again without a host round trip.  n <= 262144. */
int cmb200_put_step(cmb200_engine *e, size_t n, const cmb200_addr *addr, const uint8_t *valid,
    const void *pages, int pages_on_dev, const uint64_t *ts, uint32_t rank, void *records_dev_out,
    int32_t *lens_out, uint64_t *ticket);
int cmb200_import_records_dev(cmb200_engine *e, size_t n_total, const void *records_dev, uint32_t my_rank);

/* Slides the live records to the start of the arena so that the space of deleted and outgrown
 * records (stats.arena_garbage, and the unused remainders of per-warp segments) can be allocated
 * again; *reclaimed_out = bytes by which stats.arena_used went down.  Blocks the engine while it
 * runs (HBM speed).  The cachemap layer calls it by itself when the arena is about to overflow. */
int cmb200_compact(cmb200_engine *e, uint64_t *reclaimed_out);

/* ---- snapshot: what makes the cache directory persistent --------------------
 * The reference's store is its LMDB files under <cachedir> (cachemap/filemap.c:57,71-72) and so
 * survives a restart.  cmb200_save writes every live local record — byte for byte the LMDB value
 * of the reference, 24-byte data_prefix + payload (filemap.c:140-147), with its timestamp
 * attribute and fingerprint — to one file (written to path.tmp, then renamed); cmb200_load puts
 * the records of such a file into the store as if they had been put in file order (existing keys
 * are overwritten).  The format (engine.cu) is independent of capacity and arena size; the page
 * size must match.  Both return 0 on success, -1 with cmb200_last_error() otherwise. */
int cmb200_save(cmb200_engine *e, const char *path, uint64_t *records_out);
int cmb200_load(cmb200_engine *e, const char *path, uint64_t *records_out);

/* ---- multi-GPU: chunks sharded round-robin over ranks, one replicated key index per GPU ----
 * Each rank puts its own shard with the chunks' GLOBAL stream positions as sequence numbers
 * (next_seq = position of the rank's next chunk, stride = world size), then the ranks all-gather
 * {address, owner rank, sequence} of what they stored (NCCL, done by the caller) and import the
 * others' records: per key the highest sequence wins, exactly as sequential puts would resolve
 * (SURVEY.md §8e "ordering caveat"); a local record that loses is retired. */
int cmb200_set_stream_order(cmb200_engine *e, uint64_t next_seq, uint64_t stride);
/* loc[i] (optional) = word 3 of the exchange record of row i: owner rank << 56 | arena offset / 16
 * << 22 | stored length + 1 — where the record lies in the owner's arena, so that cmb200_get_small
 * can read it over NVLink once cmb200_open_peer has mapped that arena. */
int cmb200_import_remote(cmb200_engine *e, size_t n, const cmb200_addr *addr, const uint32_t *owner,
    const uint64_t *seq, const uint64_t *loc, int arrays_on_device);
/* The other ranks' arenas as NVLink peer memory.  Every rank exports the CUDA IPC handle of its
 * arena (64 bytes; exchange them with one all-gather), and opens the handles of the ranks it wants
 * to read from.  A cmb200_get_small of a key whose newest record lives on rank r then copies the
 * record straight out of rank r's arena (the location travelled with the exchange record) and
 * decodes it locally: replaces cachemap_get's LMDB read (cachemap.c:168-184, filemap.c:217-262) on
 * a box-global index.  Without a mapped peer such a get reports CMB200_REMOTE.  A location that no
 * longer holds the record (the owner compacted its arena since the exchange) is a miss. */
int cmb200_arena_ipc_handle(cmb200_engine *e, void *handle64_out, uint64_t *arena_bytes_out);
int cmb200_open_peer(cmb200_engine *e, uint32_t rank, const void *handle64, uint64_t arena_bytes);
/* Unmaps every peer arena (call on all ranks, then synchronise the ranks, before any of them
 * destroys its engine: an arena must not be freed while another process still maps it). */
int cmb200_close_peers(cmb200_engine *e);

/* Small batches of gets (cachemap_get from FUSE worker threads, n <= a few hundred): ONE fused
 * kernel per call — key lookup, record staged in shared memory by TMA, LZ4 decode shared -> shared,
 * page written with 16-byte stores — on a stream and a lock of its own, so a get neither waits for
 * a put batch in flight nor copies its result a second time: pages_out must be page-locked host
 * memory (cmb200_host_alloc; the kernel writes it directly) or device memory.  Records are
 * immutable and every rewrite goes to fresh arena space, so a get that overlaps a put of the same
 * key returns the old or the new page, never a mix (the reference's LMDB snapshot reads,
 * filemap.c:223-231).  Page sizes above 64 KiB are not served by this call (-2): use
 * cmb200_get_batch.  status_out as cmb200_get_batch. */
int cmb200_get_small(cmb200_engine *e, size_t n, const cmb200_addr *addr, void *pages_out, int32_t *status_out);

/* The same get in two halves, for callers that combine the requests of several threads into one
 * launch (the drop-in's combining queue, cachemap_api.c): begin (n <= 1024) launches and returns;
 * ticket.status[i] — page-locked host memory the kernel writes — holds CMB200_SMALL_PENDING until
 * request i is answered, and page i is complete in pages_out once its status is (read the status
 * with acquire semantics), so every requester can leave when ITS page is there instead of when
 * the slowest page of the batch is.  end waits for the rest, copies the statuses (status_out may
 * be NULL), books the statistics and releases the ticket's engine lane; it may be called from
 * another thread than begin, exactly once per successful begin. */
#define CMB200_SMALL_PENDING (-1)
typedef struct cmb200_small_ticket {
	int lane;                        /* engine lane the launch runs on, -1 = nothing in flight */
	uint32_t n;
	const volatile int32_t *status;
} cmb200_small_ticket;
int cmb200_get_small_begin(cmb200_engine *e, size_t n, const cmb200_addr *addr, void *pages_out, cmb200_small_ticket *ticket);
int cmb200_get_small_end(cmb200_engine *e, cmb200_small_ticket *ticket, int32_t *status_out);

/* Lookup only: status_out[i] in CMB200_{MISS,HIT,BAD_ENTRY,REMOTE}; owner_out[i] = owning rank for
 * CMB200_REMOTE. */
int cmb200_locate_batch(cmb200_engine *e, size_t n, const cmb200_addr *addr, int32_t *status_out,
    uint64_t *owner_out);

/* ---- kernel-level entry points (parity tests, benchmarks); device = CUDA ordinal or -1 ---- */

/* cachemap.c:151-166 + filemap.c:18-24 on the device. */
int cmb200_compose_keys(int device, size_t n, const uint64_t *offset, const uint64_t *nhid,
    const uint32_t *genid, int pshift, cmb200_addr *addr_out, uint8_t *valid_out, uint64_t *key_out);

/* LZ4_compress_fast(page, dst, nbytes, nbytes+1024, accel) for n pages `stride` bytes apart
 * (stride multiple of 16); blocks_out rows are out_stride apart (>= nbytes + nbytes/255 + 16).
 * fp_out nullable: EF128 {hi,lo} per page from the same (fused) kernel. */
int cmb200_lz4_encode_batch(int device, const void *pages_host, size_t n, uint32_t nbytes,
    size_t stride, int accel, void *blocks_out_host, size_t out_stride, int32_t *lens_out,
    uint64_t *fp_out);
/* LZ4_decompress_fast(block, page, nbytes): consumed_out[i] = bytes consumed or < 0. */
int cmb200_lz4_decode_batch(int device, const void *blocks_host, size_t in_stride, const int32_t *lens,
    size_t n, uint32_t nbytes, void *pages_out_host, int32_t *consumed_out);
int cmb200_fingerprint_batch(int device, const void *pages_host, size_t n, uint32_t nbytes,
    size_t stride, uint64_t *fp_out);
/* EF128 of n pages resident in the engine's HBM (1<<pshift bytes each); fp_out on the host. */
int cmb200_fingerprint_dev(cmb200_engine *e, size_t n, const void *pages_dev, uint64_t *fp_out_host);

/* ---- synthetic streams, same definition on host and device ---- */
void cmb200_gen_chunk_host(uint64_t seed, uint64_t cid, uint32_t bsize, void *out);
int cmb200_gen_chunks_dev(cmb200_engine *e, uint64_t seed, const uint64_t *cids_host, size_t n,
    void *out_dev);
/* cid[k] for a stream of n chunks with duplicate fraction dup (same-address repeats);
 * returns the number of distinct chunks. */
uint64_t cmb200_gen_stream_ids(uint64_t seed2, size_t n, double dup, uint64_t first_cid, uint64_t *cid_out);
void cmb200_gen_addr(uint64_t seed, uint64_t cid, int pshift, uint64_t *offset_out, uint64_t *nhid_out);

#ifdef __cplusplus
}
#endif
#endif
