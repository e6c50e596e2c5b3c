// engine.cu — host side of the B200 cachemap engine: HBM layout, batch pipelines, C ABI.
//
// HBM layout per engine (one engine per GPU):
//   key table   (slots+2) x 64 B          replaces the 32 LMDB environments (filemap.c:54-90)
//   arena       bump-allocated records    {24-byte data_prefix, LZ4 block | raw page}
//   page ring   2 x max_batch x bsize     double-buffered landing zone for host pages
//   stage       one (bsize+1024) row per resident encoder warp: block before it is packed into the arena
// A put batch is: H2D copy (copy stream)  ->  k_upsert  ->  k_encode (fingerprint + LZ4 + arena
// commit + table publish), sub-batch k+1's copy overlapping sub-batch k's kernels.
// A get batch is: k_lookup -> k_decode -> D2H copy.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <shared_mutex>
#include <sched.h>
#include <string>
#include <new>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../include/cachemap_b200.h"
#include "kernels.h"
#include "common.cuh"
#include "streamgen.cuh"

static thread_local char g_err[512];

void cmb_set_error(const char *what, cudaError_t e, const char *file, int line) {
	snprintf(g_err, sizeof(g_err), "%s:%d: %s: %s", file, line, what, cudaGetErrorString(e));
}
static void set_error_msg(const char *msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }

using namespace cmb;

struct cmb200_engine {
	int device = 0;
	int pshift = 16;
	uint32_t bsize = 65536;
	int accel = 12;
	uint64_t capacity = 0;
	uint32_t max_batch = 4096;   // chunks per kernel launch (stage buffer size)
	uint32_t host_batch = 4096;  // chunks per H2D/D2H pipeline step (page ring size), <= max_batch
	uint32_t flags = 0;
	cudaStream_t st = nullptr, copy = nullptr;
	// small gets (cmb200_get_small) have their own stream, lock and buffers: they neither queue behind
	// a put batch on `st` nor take `mu`
	// Several small gets may be in flight at once (two leader threads of the combining queue, or any
	// callers of cmb200_get_small): each takes one LANE — a stream plus page-locked request / status
	// words — and the open side of get_gate; what moves records or peer mappings (compaction, peers,
	// destroy) closes get_gate.
#ifndef CMB_GET_LANES
#define CMB_GET_LANES 32
#endif
	static constexpr int GET_LANES = CMB_GET_LANES;
	struct GetLane { std::atomic<int> busy{0}; cudaStream_t st = nullptr; int32_t *h_status = nullptr; cmb200_addr *h_addr = nullptr; };
	GetLane lane[GET_LANES];
	// A small get may be begun by one thread and ended by another (cmb200_get_small_begin / _end), so
	// the "no small get in flight" condition is a counter and a closing flag, not a lock a thread owns.
	struct GetGate {
		std::atomic<int> active{0}, closed{0};
		void enter() {
			for (;;) {
				while (closed.load(std::memory_order_acquire)) sched_yield();
				active.fetch_add(1, std::memory_order_acq_rel);
				if (!closed.load(std::memory_order_acquire)) return;
				active.fetch_sub(1, std::memory_order_acq_rel);
			}
		}
		void leave() { active.fetch_sub(1, std::memory_order_acq_rel); }
		void close() {                        // exclusive: waits for the gets in flight, holds new ones off
			int open = 0;
			while (!closed.compare_exchange_weak(open, 1, std::memory_order_acq_rel)) { open = 0; sched_yield(); }
			while (active.load(std::memory_order_acquire)) sched_yield();
		}
		void reopen() { closed.store(0, std::memory_order_release); }
	} get_gate;
	struct GateClosed {                           // scope guard of the exclusive side
		GetGate &g;
		explicit GateClosed(GetGate &gate) : g(gate) { g.close(); }
		~GateClosed() { g.reopen(); }
	};
	std::atomic<uint32_t> lane_turn{0};
	static constexpr size_t GET_SMALL_MAX = 1024;
	const uint8_t *peer_base[GET_MAX_PEERS] = {};
	uint64_t peer_size[GET_MAX_PEERS] = {};
	std::atomic<uint64_t> small_get_requests{0}, small_get_hits{0}, small_get_launches{0};
	unsigned long long *d_recoff_out = nullptr;  // arena offset per chunk of the current put slice (exchange records)
	// k_get_small's sequence descriptors: one scratch region per CTA that can be resident (kernels.cu:gs_region_take)
	uint4 *d_scratch = nullptr;
	uint32_t *d_pool_bits = nullptr;
	uint32_t pool_n = 0, region_entries = 0;
	cudaEvent_t landed[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
	TableView table{};
	ArenaView arena{};
	unsigned long long *d_counters = nullptr;   // entries, tombs, head, garbage, dropped
	uint8_t *d_pages[2] = {nullptr, nullptr};
	uint8_t *d_stage = nullptr;
	uint64_t stage_stride = 0;
	unsigned long long *d_addr = nullptr, *d_ts = nullptr;
	uint8_t *d_valid = nullptr;
	uint32_t *d_slot = nullptr, *d_vlen = nullptr;
	int32_t *d_lens = nullptr, *d_status = nullptr;
	uint64_t *d_fps = nullptr, *d_recoff = nullptr;
	unsigned int *d_work = nullptr;
	uint32_t *d_import_slot = nullptr;   // slot scratch of cmb200_import_records_dev
	size_t import_slot_cap = 0;
	// page-locked staging for the small per-chunk arrays, so that no copy ever blocks the host
	// thread that is feeding the pipeline (a pageable cudaMemcpyAsync waits for the stream)
	static constexpr size_t META_CAP = 1u << 18;   // chunks per outer slice of a call
	uint8_t *h_meta = nullptr;                     // META_CAP x (16 addr + 8 ts + 4 lens + 4 status + 1 valid)
	unsigned long long seq = 1;          // sequence of the next chunk
	unsigned long long seq_stride = 1;   // > 1 when the global stream is sharded round-robin over ranks
	// per-launch device timing of the dominant kernels (roofline evidence for bench.py)
	static constexpr int RING = 64;
	cudaEvent_t t0[RING] = {}, t1[RING] = {};
	// asynchronous puts (cmb200_put_batch_async): kernel timing events not harvested yet are
	// p0/p1[pend_tail .. pend_head), completion tickets are events on the compute stream
	cudaEvent_t p0[RING] = {}, p1[RING] = {};
	uint64_t pend_head = 0, pend_tail = 0;
	static constexpr int TICKETS = 8;
	cudaEvent_t ticket_ev[TICKETS] = {};
	uint64_t tickets = 0;
	cudaEvent_t meta_done = nullptr, meta_free[2] = {nullptr, nullptr};
	uint64_t ring_pos = 0;               // page ring buffer in turn, kept across calls
	uint64_t meta_pos = 0;               // same for the two copies of d_addr / d_ts / d_valid (host puts)
	size_t meta_cap = 0;                 // entries per copy
	std::mutex mu;
	cmb200_stats stats{};
};

#define ENG_CHECK(expr)                                                 \
	do {                                                            \
		cudaError_t e_ = (expr);                                \
		if (e_ != cudaSuccess) {                                \
			cmb_set_error(#expr, e_, __FILE__, __LINE__);   \
			goto fail;                                      \
		}                                                       \
	} while (0)

static uint64_t next_pow2(uint64_t v) {
	uint64_t p = 1;
	while (p < v) p <<= 1;
	return p;
}

extern "C" const char *cmb200_last_error(void) { return g_err; }

extern "C" int cmb200_device_count(void) {
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) { cmb_set_error("cudaGetDeviceCount", e, __FILE__, __LINE__); return 0; }
	return n;
}

static int select_device(int device) {
	if (device >= 0) CMB_CHECK(cudaSetDevice(device));
	int n = 0;
	CMB_CHECK(cudaGetDeviceCount(&n));
	if (n == 0) { set_error_msg("no CUDA device"); return -1; }
	return 0;
}

extern "C" void cmb200_engine_destroy(cmb200_engine *e) {
	if (!e) return;
	cudaSetDevice(e->device);
	if (e->st) cudaStreamSynchronize(e->st);
	if (e->copy) cudaStreamSynchronize(e->copy);
	cudaFree(e->d_scratch); cudaFree(e->d_pool_bits); cudaFree(e->table.ckpt);
	cudaFree(e->table.slots); cudaFree(e->table.fp); cudaFree(e->arena.base); cudaFree(e->arena.seg); cudaFree(e->d_counters);
	cudaFree(e->d_pages[0]); cudaFree(e->d_pages[1]); cudaFree(e->d_stage);
	cudaFree(e->d_addr); cudaFree(e->d_ts); cudaFree(e->d_valid); cudaFree(e->d_slot); cudaFree(e->d_vlen);
	if (e->h_meta) cudaFreeHost(e->h_meta);
	for (auto &ln : e->lane) {
		if (ln.st) { cudaStreamSynchronize(ln.st); cudaStreamDestroy(ln.st); }
		if (ln.h_status) cudaFreeHost(ln.h_status);
		if (ln.h_addr) cudaFreeHost(ln.h_addr);
	}
	cudaFree(e->d_recoff_out);
	for (int r = 0; r < GET_MAX_PEERS; r++) if (e->peer_base[r]) cudaIpcCloseMemHandle((void *)e->peer_base[r]);
	cudaFree(e->d_lens); cudaFree(e->d_status); cudaFree(e->d_fps); cudaFree(e->d_recoff); cudaFree(e->d_work); cudaFree(e->d_import_slot);
	for (int i = 0; i < 2; i++) {
		if (e->landed[i]) cudaEventDestroy(e->landed[i]);
		if (e->consumed[i]) cudaEventDestroy(e->consumed[i]);
	}
	for (int i = 0; i < cmb200_engine::RING; i++) {
		if (e->t0[i]) cudaEventDestroy(e->t0[i]);
		if (e->t1[i]) cudaEventDestroy(e->t1[i]);
		if (e->p0[i]) cudaEventDestroy(e->p0[i]);
		if (e->p1[i]) cudaEventDestroy(e->p1[i]);
	}
	for (int i = 0; i < cmb200_engine::TICKETS; i++) if (e->ticket_ev[i]) cudaEventDestroy(e->ticket_ev[i]);
	if (e->meta_done) cudaEventDestroy(e->meta_done);
	for (int i = 0; i < 2; i++) if (e->meta_free[i]) cudaEventDestroy(e->meta_free[i]);
	if (e->st) cudaStreamDestroy(e->st);
	if (e->copy) cudaStreamDestroy(e->copy);
	delete e;
}

extern "C" cmb200_engine *cmb200_engine_create(const cmb200_config *cfg) {
	if (!cfg || cfg->pshift < 6 || cfg->pshift > 20) { set_error_msg("bad config: pshift must be 6..20"); return nullptr; }
	if (select_device(cfg->device) != 0) return nullptr;
	cmb200_engine *e = new (std::nothrow) cmb200_engine();
	if (!e) return nullptr;
	cudaGetDevice(&e->device);
	e->pshift = cfg->pshift;
	e->bsize = 1u << cfg->pshift;
	e->accel = cfg->accel < 0 ? 1 : (cfg->accel > (1 << 20) ? (1 << 20) : cfg->accel);   // lz4.c:740
	e->capacity = cfg->capacity;
	e->max_batch = cfg->max_batch ? cfg->max_batch : 4096;
	{
		// host pages are pipelined in smaller steps than a resident batch is launched in: the first
		// copy of a call cannot overlap anything, while a launch wants many chunks per warp
		const char *hb = getenv("CMB200_HOST_BATCH");
		uint32_t v = hb ? (uint32_t)strtoul(hb, nullptr, 0) : 4096u;
		e->host_batch = v && v < e->max_batch ? v : e->max_batch;
	}
	e->flags = cfg->flags;
	const uint64_t B = e->max_batch;
	{
		uint64_t slots = cfg->table_slots ? next_pow2(cfg->table_slots) : next_pow2(4 * (cfg->capacity ? cfg->capacity : 1024));
		if (slots < 1024) slots = 1024;
		const char *cap_env = getenv("CMB200_MAX_TABLE_SLOTS");
		uint64_t max_slots = cap_env ? next_pow2(strtoull(cap_env, nullptr, 0)) : (1ull << 27);
		if (slots > max_slots) slots = max_slots;
		e->table.cap = slots;
		ENG_CHECK(cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking));
		ENG_CHECK(cudaStreamCreateWithFlags(&e->copy, cudaStreamNonBlocking));
		for (auto &ln : e->lane) ENG_CHECK(cudaStreamCreateWithFlags(&ln.st, cudaStreamNonBlocking));
		for (int i = 0; i < 2; i++) {
			ENG_CHECK(cudaEventCreateWithFlags(&e->landed[i], cudaEventDisableTiming));
			ENG_CHECK(cudaEventCreateWithFlags(&e->consumed[i], cudaEventDisableTiming));
		}
		for (int i = 0; i < cmb200_engine::RING; i++) {
			ENG_CHECK(cudaEventCreate(&e->t0[i]));
			ENG_CHECK(cudaEventCreate(&e->t1[i]));
			ENG_CHECK(cudaEventCreate(&e->p0[i]));
			ENG_CHECK(cudaEventCreate(&e->p1[i]));
		}
		for (int i = 0; i < cmb200_engine::TICKETS; i++)
			ENG_CHECK(cudaEventCreateWithFlags(&e->ticket_ev[i], cudaEventDisableTiming));
		ENG_CHECK(cudaEventCreateWithFlags(&e->meta_done, cudaEventDisableTiming));
		ENG_CHECK(cudaEventCreateWithFlags(&e->meta_free[0], cudaEventDisableTiming));
		ENG_CHECK(cudaEventCreateWithFlags(&e->meta_free[1], cudaEventDisableTiming));
		ENG_CHECK(cudaMalloc(&e->table.slots, (slots + 2) * sizeof(Slot)));
		ENG_CHECK(cudaMemsetAsync(e->table.slots, 0, (slots + 2) * sizeof(Slot), e->st));
		if (e->flags & CMB200_FINGERPRINT) {
			ENG_CHECK(cudaMalloc(&e->table.fp, (slots + 2) * 16));
			ENG_CHECK(cudaMemsetAsync(e->table.fp, 0, (slots + 2) * 16, e->st));
		}
		if (get_small_supports(e->bsize)) {
			// parse checkpoints per slot (64 bytes) and the descriptor scratch of the fused single-page get
			const char *ck = getenv("CMB200_CKPT");
			if (!ck || atoi(ck) != 0) {
				ENG_CHECK(cudaMalloc(&e->table.ckpt, (slots + 2) * CKPT_WORDS * 4));
				ENG_CHECK(cudaMemsetAsync(e->table.ckpt, 0, (slots + 2) * CKPT_WORDS * 4, e->st));
			}
			const int resident = get_small_residency(e->bsize);
			if (resident <= 0) { set_error_msg("k_get_small does not fit this device"); goto fail; }
			e->pool_n = (uint32_t)resident;
			e->region_entries = get_small_region_entries(e->bsize);
			ENG_CHECK(cudaMalloc(&e->d_scratch, (size_t)e->pool_n * e->region_entries * sizeof(uint4)));
			ENG_CHECK(cudaMalloc(&e->d_pool_bits, ((size_t)e->pool_n + 31) / 32 * 4));
			ENG_CHECK(cudaMemsetAsync(e->d_pool_bits, 0, ((size_t)e->pool_n + 31) / 32 * 4, e->st));
		}
		ENG_CHECK(cudaMalloc(&e->d_counters, 8 * sizeof(unsigned long long)));
		ENG_CHECK(cudaMemsetAsync(e->d_counters, 0, 8 * sizeof(unsigned long long), e->st));
		e->table.entries = e->d_counters + 0;
		e->table.tombs = e->d_counters + 1;
		e->arena.head = e->d_counters + 2;
		e->arena.garbage = e->d_counters + 3;
		e->arena.dropped = e->d_counters + 4;
		e->table.remote = e->d_counters + 5;

		e->stage_stride = ((uint64_t)e->bsize + 1024 + 15) & ~15ull;        // filemap.c:120 dest[bsize+1024]
		ENG_CHECK(cudaMalloc(&e->d_pages[0], (uint64_t)e->host_batch * e->bsize + 256));
		ENG_CHECK(cudaMalloc(&e->d_pages[1], (uint64_t)e->host_batch * e->bsize + 256));
		// one stage row per resident warp / group of the encode kernels (store mode), not per chunk
		ENG_CHECK(cudaMalloc(&e->d_stage, (uint64_t)16384 * e->stage_stride + 256));
		// small per-chunk arrays are sized for a whole slice of a call (META_CAP chunks) so that
		// they cross PCIe once, outside the page pipeline
		const uint64_t M = cmb200_engine::META_CAP > B ? cmb200_engine::META_CAP : B;
		// two copies: a host put stages the next call's arrays while the kernels of the previous
		// one still read theirs (cmb200_put_batch_async)
		e->meta_cap = M;
		ENG_CHECK(cudaMalloc(&e->d_addr, 2 * M * 16));
		ENG_CHECK(cudaMalloc(&e->d_ts, 2 * M * 8));
		ENG_CHECK(cudaMalloc(&e->d_valid, 2 * M));
		ENG_CHECK(cudaMalloc(&e->d_slot, B * 4));
		ENG_CHECK(cudaMalloc(&e->d_vlen, B * 4));
		ENG_CHECK(cudaMalloc(&e->d_lens, M * 4));
		ENG_CHECK(cudaMalloc(&e->d_status, M * 4));
		ENG_CHECK(cudaMalloc(&e->d_fps, B * 16));
		ENG_CHECK(cudaMalloc(&e->d_recoff, B * 8));
		ENG_CHECK(cudaMalloc(&e->d_work, 64));
		ENG_CHECK(cudaMallocHost(&e->h_meta, cmb200_engine::META_CAP * 33));
		for (auto &ln : e->lane) {
			ENG_CHECK(cudaMallocHost(&ln.h_status, cmb200_engine::GET_SMALL_MAX * 4));
			ENG_CHECK(cudaMallocHost(&ln.h_addr, cmb200_engine::GET_SMALL_MAX * 16));
		}
		ENG_CHECK(cudaMalloc(&e->d_recoff_out, M * 8));

		uint64_t arena = cfg->arena_bytes;
		if (!arena) {
			size_t free_b = 0, total_b = 0;
			ENG_CHECK(cudaMemGetInfo(&free_b, &total_b));
			// capacity worst-case records (incompressible pages: 24 + bsize + bsize/255 + 16) plus 1/8
			// headroom, so that a store at capacity still has garbage worth compacting
			uint64_t want = (cfg->capacity ? cfg->capacity : 1024) * (e->stage_stride + 32);
			want += want / 8;
			uint64_t lim = (uint64_t)(free_b * 0.8);
			arena = want < lim ? want : lim;
		}
		arena = (arena + 255) & ~255ull;
		ENG_CHECK(cudaMalloc(&e->arena.base, arena + 256));
		e->arena.size = arena;
		{
			// direct encode into per-warp arena segments when the arena is large enough that
			// 4096 segments of >= 4 worst-case records stay a small part of it
			// (CMB200_SEG_KB overrides: 0 = always through the stage buffer)
			const uint64_t worst = (24 + e->stage_stride + 15) & ~15ull;
			uint64_t seg = arena / (8ull * 2072ull);
			if (seg > (2ull << 20)) seg = 2ull << 20;
			if (seg < 4 * worst) seg = 0;
			const char *kb = getenv("CMB200_SEG_KB");
			if (kb && *kb) { seg = strtoull(kb, nullptr, 10) << 10; if (seg && seg < worst) seg = worst; }
			seg = (seg + 255) & ~255ull;
			e->arena.seg_bytes = (uint32_t)seg;
			ENG_CHECK(cudaMalloc(&e->arena.seg, ARENA_SEG_SLOTS * 2 * sizeof(unsigned long long)));
			ENG_CHECK(cudaMemsetAsync(e->arena.seg, 0, ARENA_SEG_SLOTS * 2 * sizeof(unsigned long long), e->st));
		}
		ENG_CHECK(cudaStreamSynchronize(e->st));
	}
	return e;
fail:
	cmb200_engine_destroy(e);
	return nullptr;
}

extern "C" void *cmb200_host_alloc(size_t bytes) {
	void *p = nullptr;
	cudaError_t er = cudaMallocHost(&p, bytes);
	if (er != cudaSuccess) { cmb_set_error("cudaMallocHost", er, __FILE__, __LINE__); return nullptr; }
	return p;
}
extern "C" void cmb200_host_free(void *p) { if (p) cudaFreeHost(p); }
extern "C" void *cmb200_dev_alloc(cmb200_engine *e, size_t bytes) {
	void *p = nullptr;
	if (e) cudaSetDevice(e->device);
	cudaError_t er = cudaMalloc(&p, bytes + 256);
	if (er != cudaSuccess) { cmb_set_error("cudaMalloc", er, __FILE__, __LINE__); return nullptr; }
	return p;
}
extern "C" void cmb200_dev_free(cmb200_engine *e, void *p) { if (e) cudaSetDevice(e->device); if (p) cudaFree(p); }
extern "C" int cmb200_memcpy_h2d(cmb200_engine *e, void *dev, const void *host, size_t bytes) {
	cudaSetDevice(e->device);
	CMB_CHECK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, e->st));
	CMB_CHECK(cudaStreamSynchronize(e->st));
	return 0;
}
extern "C" int cmb200_memcpy_d2h(cmb200_engine *e, void *host, const void *dev, size_t bytes) {
	cudaSetDevice(e->device);
	CMB_CHECK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, e->st));
	CMB_CHECK(cudaStreamSynchronize(e->st));
	return 0;
}
extern "C" void *cmb200_stream(cmb200_engine *e) { return (void *)e->st; }
extern "C" int cmb200_sync(cmb200_engine *e) {
	cudaSetDevice(e->device);
	CMB_CHECK(cudaStreamSynchronize(e->copy));
	CMB_CHECK(cudaStreamSynchronize(e->st));
	return 0;
}

// ---- put ---------------------------------------------------------------------------------

// Adds the kernel times of asynchronous puts whose events have completed to the statistics;
// wait = true blocks until every pending one has.
static void harvest_pending(cmb200_engine *e, bool wait) {
	while (e->pend_tail < e->pend_head) {
		const int k = (int)(e->pend_tail % cmb200_engine::RING);
		if (wait) cudaEventSynchronize(e->p1[k]);
		else if (cudaEventQuery(e->p1[k]) != cudaSuccess) { (void)cudaGetLastError(); break; }
		float ms = 0;
		if (cudaEventElapsedTime(&ms, e->p0[k], e->p1[k]) == cudaSuccess) {
			e->stats.encode_kernel_ns += (uint64_t)(ms * 1e6);
			e->stats.encode_kernel_launches++;
		}
		e->pend_tail++;
	}
}

// One slice (<= META_CAP chunks) of a put.  ticket == nullptr: returns when the chunks are stored.
// ticket != nullptr (host pages only): returns as soon as the caller's arrays have crossed to the
// device; the encode of the last sub-batch (and the copy of lens_out, which must then be
// page-locked and stay valid) completes behind the ticket.
// device-resident exchange records of a step; keep_pages: the caller keeps the host pages untouched
// until the ticket is done, so the call need not wait for its own copies
struct StepRecords { uint32_t rank; unsigned long long *out; bool keep_pages; };

static int put_slice(cmb200_engine *e, size_t n, const cmb200_addr *addr, const uint8_t *valid,
    const uint8_t *pages, bool pages_on_dev, const uint64_t *ts, int32_t *lens_out, uint64_t *ticket,
    const StepRecords *recs = nullptr) {
	const size_t B = pages_on_dev ? e->max_batch : e->host_batch;
	const bool deferred = ticket != nullptr;
	const bool copy_meta = !pages_on_dev || deferred;      // small arrays travel on the copy stream
	const unsigned long long seq_first = e->seq;
	// stage the small arrays in page-locked memory once; every copy below is then truly async
	cmb200_addr *h_addr = (cmb200_addr *)e->h_meta;
	uint64_t *h_ts = (uint64_t *)(e->h_meta + cmb200_engine::META_CAP * 16);
	int32_t *h_lens = (int32_t *)(e->h_meta + cmb200_engine::META_CAP * 24);
	uint8_t *h_valid = e->h_meta + cmb200_engine::META_CAP * 32;
	harvest_pending(e, !deferred);
	memcpy(h_addr, addr, n * 16);
	if (ts) memcpy(h_ts, ts, n * 8);
	if (valid) memcpy(h_valid, valid, n);
	// The small arrays go over once, before the page pipeline starts: a small copy issued between
	// two page copies would queue behind the next 256 MiB transfer in the copy engine and stall
	// the kernels that wait for it.  For host pages they travel on the copy stream into the copy of
	// the arrays that the previous call is not using: on the compute stream they would wait for
	// the previous call's last encode, and the page copies issued after them would wait with them
	// in the copy engine's queue.
	const int mb = copy_meta ? (int)(e->meta_pos++ & 1) : 0;
	unsigned long long *d_addr = e->d_addr + (size_t)mb * e->meta_cap * 2;
	unsigned long long *d_ts = e->d_ts + (size_t)mb * e->meta_cap;
	uint8_t *d_valid = e->d_valid + (size_t)mb * e->meta_cap;
	cudaStream_t ms = copy_meta ? e->copy : e->st;
	if (copy_meta) CMB_CHECK(cudaStreamWaitEvent(e->copy, e->meta_free[mb], 0));
	CMB_CHECK(cudaMemcpyAsync(d_addr, h_addr, n * 16, cudaMemcpyHostToDevice, ms));
	if (valid) CMB_CHECK(cudaMemcpyAsync(d_valid, h_valid, n, cudaMemcpyHostToDevice, ms));
	if (ts) CMB_CHECK(cudaMemcpyAsync(d_ts, h_ts, n * 8, cudaMemcpyHostToDevice, ms));
	if (copy_meta) {
		CMB_CHECK(cudaEventRecord(e->meta_done, e->copy));
		CMB_CHECK(cudaStreamWaitEvent(e->st, e->meta_done, 0));
	}
	size_t nb = 0;
	int last_buf = 0;
	static const bool trace = getenv("CMB200_TRACE") != nullptr;
	cudaEvent_t tr[4][16];
	if (trace) for (int i = 0; i < 4; i++) for (int j = 0; j < 16; j++) cudaEventCreate(&tr[i][j]);
	for (size_t at = 0; at < n; at += B, nb++) {
		// (splitting the last step into smaller ones to shorten the un-overlapped tail was tried:
		// launches below ~4096 chunks run below PCIe rate, so the tail got longer, not shorter)
		const uint32_t m = (uint32_t)((n - at < B) ? n - at : B);
		const int buf = (int)(e->ring_pos & 1);
		const uint8_t *d_in;
		if (pages_on_dev) {
			d_in = pages + at * e->bsize;
		} else {
			// land the pages in ring buffer `buf` once the kernels that last read it are done
			e->ring_pos++;
			last_buf = buf;
			CMB_CHECK(cudaStreamWaitEvent(e->copy, e->consumed[buf], 0));
			if (trace && nb < 16) cudaEventRecord(tr[0][nb], e->copy);
			CMB_CHECK(cudaMemcpyAsync(e->d_pages[buf], pages + at * e->bsize, (size_t)m * e->bsize,
			    cudaMemcpyHostToDevice, e->copy));
			if (trace && nb < 16) cudaEventRecord(tr[1][nb], e->copy);
			CMB_CHECK(cudaEventRecord(e->landed[buf], e->copy));
			CMB_CHECK(cudaStreamWaitEvent(e->st, e->landed[buf], 0));
			d_in = e->d_pages[buf];
		}
		if (launch_upsert(e->table, d_addr + 2 * at, valid ? d_valid + at : nullptr, m, e->seq, e->seq_stride, e->d_slot, e->st)) return -1;
		EncodeJob job{};
		job.pages = d_in; job.page_stride = e->bsize; job.nbytes = e->bsize; job.n = m;
		job.accel = (uint32_t)e->accel;
		job.stage = e->d_stage; job.stage_stride = e->stage_stride;
		job.lens = e->d_lens + at;
		job.rec_out = e->d_recoff_out + at;
		job.fps = (e->flags & CMB200_FINGERPRINT) ? e->d_fps : nullptr;
		job.work = e->d_work;
		job.slot_idx = e->d_slot;
		job.addr = d_addr + 2 * at;
		job.ts = ts ? d_ts + at : nullptr;
		job.seq0 = e->seq; job.seq_stride = e->seq_stride;
		job.table = e->table; job.arena = e->arena;
		cudaEvent_t ev0, ev1;
		if (deferred) {
			if (e->pend_head - e->pend_tail >= (uint64_t)cmb200_engine::RING) harvest_pending(e, true);
			ev0 = e->p0[e->pend_head % cmb200_engine::RING]; ev1 = e->p1[e->pend_head % cmb200_engine::RING];
			e->pend_head++;
		} else {
			ev0 = e->t0[nb % e->RING]; ev1 = e->t1[nb % e->RING];
		}
		CMB_CHECK(cudaEventRecord(ev0, e->st));
		if (trace && nb < 16) cudaEventRecord(tr[2][nb], e->st);
		if (launch_encode(job, e->st)) return -1;
		if (trace && nb < 16) cudaEventRecord(tr[3][nb], e->st);
		CMB_CHECK(cudaEventRecord(ev1, e->st));
		if (!pages_on_dev) CMB_CHECK(cudaEventRecord(e->consumed[buf], e->st));
		e->seq += (unsigned long long)m * e->seq_stride;
		e->stats.kernel_launches += 2;
	}
	e->stats.put_chunks += n;
	if (recs && recs->out) {
		if (launch_pack_records(d_addr, e->d_lens, e->d_recoff_out, (uint32_t)n, seq_first, e->seq_stride, recs->rank, recs->out, e->st)) return -1;
		e->stats.kernel_launches++;
	}
	if (copy_meta) CMB_CHECK(cudaEventRecord(e->meta_free[mb], e->st));
	if (deferred) {
		if (lens_out) CMB_CHECK(cudaMemcpyAsync(lens_out, e->d_lens, n * 4, cudaMemcpyDeviceToHost, e->st));
		CMB_CHECK(cudaEventRecord(e->ticket_ev[e->tickets % cmb200_engine::TICKETS], e->st));
		*ticket = ++e->tickets;
		// the caller may reuse addr / valid / ts / pages once they have crossed
		CMB_CHECK(cudaEventSynchronize(e->meta_done));
		if (nb && !pages_on_dev && !(recs && recs->keep_pages)) CMB_CHECK(cudaEventSynchronize(e->landed[last_buf]));
		harvest_pending(e, false);
		return 0;
	}
	if (lens_out) CMB_CHECK(cudaMemcpyAsync(h_lens, e->d_lens, n * 4, cudaMemcpyDeviceToHost, e->st));
	CMB_CHECK(cudaStreamSynchronize(e->st));
	if (trace) {
		for (size_t k = 0; k < nb && k < 16 && !pages_on_dev; k++) {
			float a = 0, b = 0, c = 0, d = 0;
			cudaEventElapsedTime(&a, tr[0][0], tr[0][k]); cudaEventElapsedTime(&b, tr[0][0], tr[1][k]);
			cudaEventElapsedTime(&c, tr[0][0], tr[2][k]); cudaEventElapsedTime(&d, tr[0][0], tr[3][k]);
			fprintf(stderr, "sub-batch %zu: copy %.2f-%.2f ms  encode %.2f-%.2f ms\n", k, a, b, c, d);
		}
		for (int i = 0; i < 4; i++) for (int j = 0; j < 16; j++) cudaEventDestroy(tr[i][j]);
	}
	if (lens_out) memcpy(lens_out, h_lens, n * 4);
	for (size_t k = 0; k < nb && k < (size_t)e->RING; k++) {
		float ms = 0;
		CMB_CHECK(cudaEventElapsedTime(&ms, e->t0[k], e->t1[k]));
		e->stats.encode_kernel_ns += (uint64_t)(ms * 1e6);
		e->stats.encode_kernel_launches++;
	}
	return 0;
}

static int put_impl(cmb200_engine *e, size_t n, const cmb200_addr *addr, const uint8_t *valid,
    const uint8_t *pages, bool pages_on_dev, const uint64_t *ts, int32_t *lens_out, uint64_t *ticket) {
	std::lock_guard<std::mutex> g(e->mu);
	CMB_CHECK(cudaSetDevice(e->device));
	if (ticket) *ticket = e->tickets;       // nothing to wait for unless the last slice is deferred
	for (size_t at = 0; at < n; at += cmb200_engine::META_CAP) {
		size_t m = n - at < cmb200_engine::META_CAP ? n - at : cmb200_engine::META_CAP;
		const bool last = at + m == n;
		if (put_slice(e, m, addr + at, valid ? valid + at : nullptr, pages + at * e->bsize, pages_on_dev,
			ts ? ts + at : nullptr, lens_out ? lens_out + at : nullptr, last ? ticket : nullptr)) return -1;
	}
	return 0;
}

extern "C" int cmb200_put_batch(cmb200_engine *e, size_t n, const cmb200_addr *addr, const uint8_t *valid,
    const void *pages_host, const uint64_t *ts, int32_t *lens_out) {
	return put_impl(e, n, addr, valid, (const uint8_t *)pages_host, false, ts, lens_out, nullptr);
}
extern "C" int cmb200_put_batch_async(cmb200_engine *e, size_t n, const cmb200_addr *addr, const uint8_t *valid,
    const void *pages_host, const uint64_t *ts, int32_t *lens_out, uint64_t *ticket) {
	uint64_t t = 0;
	const int rc = put_impl(e, n, addr, valid, (const uint8_t *)pages_host, false, ts, lens_out, &t);
	if (ticket) *ticket = t;
	return rc;
}
extern "C" int cmb200_put_step(cmb200_engine *e, size_t n, const cmb200_addr *addr, const uint8_t *valid,
    const void *pages, int pages_on_dev, const uint64_t *ts, uint32_t rank, void *records_dev_out,
    int32_t *lens_out, uint64_t *ticket) {
	if (n > cmb200_engine::META_CAP) { set_error_msg("cmb200_put_step: more than 262144 chunks in one step"); return -1; }
	std::lock_guard<std::mutex> g(e->mu);
	CMB_CHECK(cudaSetDevice(e->device));
	uint64_t t = e->tickets;
	StepRecords r{rank, (unsigned long long *)records_dev_out, pages_on_dev == 2};
	const int rc = n ? put_slice(e, n, addr, valid, (const uint8_t *)pages, pages_on_dev == 1, ts, lens_out, &t, &r) : 0;
	if (ticket) *ticket = t;
	return rc;
}

extern "C" int cmb200_import_records_dev(cmb200_engine *e, size_t n_total, const void *records_dev, uint32_t my_rank) {
	std::lock_guard<std::mutex> g(e->mu);
	CMB_CHECK(cudaSetDevice(e->device));
	if (n_total == 0) return 0;
	if (n_total > 0xffffffffull) { set_error_msg("cmb200_import_records_dev: too many records"); return -1; }
	// the whole gathered buffer in one claim + one apply launch (the slot scratch grows on demand)
	if (n_total > e->import_slot_cap) {
		CMB_CHECK(cudaStreamSynchronize(e->st));
		if (e->d_import_slot) cudaFree(e->d_import_slot);
		e->d_import_slot = nullptr; e->import_slot_cap = 0;
		CMB_CHECK(cudaMalloc(&e->d_import_slot, n_total * sizeof(uint32_t)));
		e->import_slot_cap = n_total;
	}
	if (launch_import_records(e->table, e->arena, (const unsigned long long *)records_dev, (uint32_t)n_total, my_rank,
		e->d_import_slot, e->st)) return -1;
	e->stats.kernel_launches += 2;
	return 0;                                               // asynchronous: ordered on the engine's stream
}

extern "C" int cmb200_wait(cmb200_engine *e, uint64_t ticket) {
	cudaEvent_t ev = nullptr;
	{
		std::lock_guard<std::mutex> g(e->mu);
		if (ticket == 0 || ticket > e->tickets) return 0;
		// a ticket whose event has been recorded again waits for the later put: the stream is in order
		ev = e->ticket_ev[(ticket - 1) % cmb200_engine::TICKETS];
		if (cudaSetDevice(e->device) != cudaSuccess) return -1;
	}
	CMB_CHECK(cudaEventSynchronize(ev));
	return 0;
}
extern "C" int cmb200_put_batch_dev(cmb200_engine *e, size_t n, const cmb200_addr *addr, const uint8_t *valid,
    const void *pages_dev, const uint64_t *ts, int32_t *lens_out) {
	return put_impl(e, n, addr, valid, (const uint8_t *)pages_dev, true, ts, lens_out, nullptr);
}

// ---- get ---------------------------------------------------------------------------------

static int get_slice(cmb200_engine *e, size_t n, const cmb200_addr *addr, const uint8_t *valid,
    uint8_t *pages_out, bool out_on_dev, int32_t *status_out) {
	const size_t B = out_on_dev ? e->max_batch : e->host_batch;
	cmb200_addr *h_addr = (cmb200_addr *)e->h_meta;
	int32_t *h_status = (int32_t *)(e->h_meta + cmb200_engine::META_CAP * 28);
	uint8_t *h_valid = e->h_meta + cmb200_engine::META_CAP * 32;
	memcpy(h_addr, addr, n * 16);
	if (valid) memcpy(h_valid, valid, n);
	CMB_CHECK(cudaMemcpyAsync(e->d_addr, h_addr, n * 16, cudaMemcpyHostToDevice, e->st));
	if (valid) CMB_CHECK(cudaMemcpyAsync(e->d_valid, h_valid, n, cudaMemcpyHostToDevice, e->st));
	size_t nb = 0;
	for (size_t at = 0; at < n; at += B, nb++) {
		const uint32_t m = (uint32_t)((n - at < B) ? n - at : B);
		const int buf = (int)(nb & 1);
		uint8_t *d_out = out_on_dev ? pages_out + at * e->bsize : e->d_pages[buf];
		if (!out_on_dev) CMB_CHECK(cudaStreamWaitEvent(e->st, e->consumed[buf], 0));   // D2H of buf finished
		if (launch_lookup(e->table, e->d_addr + 2 * at, valid ? e->d_valid + at : nullptr, m, e->d_status + at, e->d_recoff,
			e->d_vlen, nullptr, e->st)) return -1;
		DecodeJob job{};
		job.n = m; job.nbytes = e->bsize; job.pages = d_out; job.status = e->d_status + at;
		job.rec_off = e->d_recoff; job.vlen = e->d_vlen; job.arena = e->arena.base;
		CMB_CHECK(cudaEventRecord(e->t0[nb % e->RING], e->st));
		if (launch_decode(job, e->st)) return -1;
		CMB_CHECK(cudaEventRecord(e->t1[nb % e->RING], e->st));
		if (!out_on_dev) {
			CMB_CHECK(cudaEventRecord(e->landed[buf], e->st));
			CMB_CHECK(cudaStreamWaitEvent(e->copy, e->landed[buf], 0));
			CMB_CHECK(cudaMemcpyAsync(pages_out + at * e->bsize, d_out, (size_t)m * e->bsize,
			    cudaMemcpyDeviceToHost, e->copy));
			CMB_CHECK(cudaEventRecord(e->consumed[buf], e->copy));
		}
		e->stats.kernel_launches += 2;
	}
	CMB_CHECK(cudaMemcpyAsync(h_status, e->d_status, n * 4, cudaMemcpyDeviceToHost, e->st));
	CMB_CHECK(cudaStreamSynchronize(e->st));
	CMB_CHECK(cudaStreamSynchronize(e->copy));
	memcpy(status_out, h_status, n * 4);
	for (size_t k = 0; k < nb && k < (size_t)e->RING; k++) {
		float ms = 0;
		CMB_CHECK(cudaEventElapsedTime(&ms, e->t0[k], e->t1[k]));
		e->stats.decode_kernel_ns += (uint64_t)(ms * 1e6);
		e->stats.decode_kernel_launches++;
	}
	for (size_t i = 0; i < n; i++) {
		if (status_out[i] != CMB200_INVALID) e->stats.get_requests++;
		if (status_out[i] == CMB200_HIT) e->stats.get_hits++;
	}
	return 0;
}

static int get_impl(cmb200_engine *e, size_t n, const cmb200_addr *addr, const uint8_t *valid,
    uint8_t *pages_out, bool out_on_dev, int32_t *status_out) {
	std::lock_guard<std::mutex> g(e->mu);
	CMB_CHECK(cudaSetDevice(e->device));
	for (size_t at = 0; at < n; at += cmb200_engine::META_CAP) {
		size_t m = n - at < cmb200_engine::META_CAP ? n - at : cmb200_engine::META_CAP;
		if (get_slice(e, m, addr + at, valid ? valid + at : nullptr, pages_out + at * e->bsize, out_on_dev,
			status_out + at)) return -1;
	}
	return 0;
}

extern "C" int cmb200_get_batch(cmb200_engine *e, size_t n, const cmb200_addr *addr, const uint8_t *valid,
    void *pages_out_host, int32_t *status_out) {
	return get_impl(e, n, addr, valid, (uint8_t *)pages_out_host, false, status_out);
}
extern "C" int cmb200_locate_batch(cmb200_engine *e, size_t n, const cmb200_addr *addr, int32_t *status_out,
    uint64_t *owner_out) {
	// lookup only: which requests hit here, miss, or live on another rank (owner_out = rank)
	std::lock_guard<std::mutex> g(e->mu);
	CMB_CHECK(cudaSetDevice(e->device));
	for (size_t at = 0; at < n; at += e->max_batch) {
		uint32_t m = (uint32_t)((n - at < e->max_batch) ? n - at : e->max_batch);
		CMB_CHECK(cudaMemcpyAsync(e->d_addr, addr + at, (size_t)m * 16, cudaMemcpyHostToDevice, e->st));
		if (launch_lookup(e->table, e->d_addr, nullptr, m, e->d_status, e->d_recoff, e->d_vlen, nullptr, e->st)) return -1;
		CMB_CHECK(cudaMemcpyAsync(status_out + at, e->d_status, (size_t)m * 4, cudaMemcpyDeviceToHost, e->st));
		CMB_CHECK(cudaMemcpyAsync(owner_out + at, e->d_recoff, (size_t)m * 8, cudaMemcpyDeviceToHost, e->st));
		e->stats.kernel_launches++;
	}
	CMB_CHECK(cudaStreamSynchronize(e->st));
	return 0;
}
extern "C" int cmb200_get_batch_dev(cmb200_engine *e, size_t n, const cmb200_addr *addr, const uint8_t *valid,
    void *pages_out_dev, int32_t *status_out) {
	return get_impl(e, n, addr, valid, (uint8_t *)pages_out_dev, true, status_out);
}

// ---- multi-GPU index replication -------------------------------------------------------------

extern "C" int cmb200_set_stream_order(cmb200_engine *e, uint64_t next_seq, uint64_t stride) {
	std::lock_guard<std::mutex> g(e->mu);
	if (stride == 0) { set_error_msg("stream order: stride must be >= 1"); return -1; }
	e->seq = next_seq; e->seq_stride = stride;
	return 0;
}

extern "C" int cmb200_import_remote(cmb200_engine *e, size_t n, const cmb200_addr *addr, const uint32_t *owner,
    const uint64_t *seq, const uint64_t *loc, int on_dev) {
	std::lock_guard<std::mutex> g(e->mu);
	CMB_CHECK(cudaSetDevice(e->device));
	for (size_t at = 0; at < n; at += e->max_batch) {
		uint32_t m = (uint32_t)((n - at < e->max_batch) ? n - at : e->max_batch);
		const unsigned long long *d_a; const uint32_t *d_o; const unsigned long long *d_s; const unsigned long long *d_l = nullptr;
		if (on_dev) {
			d_a = (const unsigned long long *)(addr + at); d_o = owner + at; d_s = (const unsigned long long *)(seq + at);
			if (loc) d_l = (const unsigned long long *)(loc + at);
		} else {
			CMB_CHECK(cudaMemcpyAsync(e->d_addr, addr + at, (size_t)m * 16, cudaMemcpyHostToDevice, e->st));
			CMB_CHECK(cudaMemcpyAsync(e->d_vlen, owner + at, (size_t)m * 4, cudaMemcpyHostToDevice, e->st));
			CMB_CHECK(cudaMemcpyAsync(e->d_ts, seq + at, (size_t)m * 8, cudaMemcpyHostToDevice, e->st));
			if (loc) CMB_CHECK(cudaMemcpyAsync(e->d_recoff, loc + at, (size_t)m * 8, cudaMemcpyHostToDevice, e->st));
			d_a = e->d_addr; d_o = e->d_vlen; d_s = e->d_ts; d_l = loc ? (const unsigned long long *)e->d_recoff : nullptr;
		}
		if (launch_import(e->table, e->arena, d_a, d_o, d_s, d_l, m, e->d_slot, e->st)) return -1;
		e->stats.kernel_launches += 2;
	}
	CMB_CHECK(cudaStreamSynchronize(e->st));
	return 0;
}

// ---- small gets: one fused kernel on their own stream ------------------------------------------

// Begin / end halves of a small get.  begin launches the kernel on a free lane and returns at once;
// the answers appear in t->status (page-locked host memory the kernel writes: a page, a system-wide
// fence, then its status word), so every requester of a combined batch can watch ITS word and leave
// as soon as its own page is there — a batch costs each caller its own page's decode, not the
// slowest one's.  end waits for whatever is still pending, books the statistics and frees the lane;
// it may run on another thread than begin.
static const int32_t SMALL_PENDING = -1;

extern "C" int cmb200_get_small_begin(cmb200_engine *e, size_t n, const cmb200_addr *addr, void *pages_out, cmb200_small_ticket *t) {
	if (!t) return -1;
	t->lane = -1; t->n = 0; t->status = nullptr;
	if (n == 0) return 0;
	if (n > cmb200_engine::GET_SMALL_MAX) { set_error_msg("cmb200_get_small_begin: more than 1024 requests"); return -1; }
	if (!get_small_supports(e->bsize)) { set_error_msg("cmb200_get_small: page size not supported by the fused kernel"); return -2; }
	e->get_gate.enter();
	// a free lane if there is one, else wait for one
	cmb200_engine::GetLane *ln = nullptr;
	int li = -1;
	const uint32_t first = e->lane_turn.fetch_add(1, std::memory_order_relaxed);
	for (uint64_t spins = 0; !ln; spins++) {
		for (int k = 0; k < cmb200_engine::GET_LANES && !ln; k++) {
			const int c = (int)((first + k) % cmb200_engine::GET_LANES);
			int idle = 0;
			if (e->lane[c].busy.compare_exchange_strong(idle, 1, std::memory_order_acq_rel)) { ln = &e->lane[c]; li = c; }
		}
		if (!ln) sched_yield();
	}
	if (cudaSetDevice(e->device) != cudaSuccess) { ln->busy.store(0, std::memory_order_release); e->get_gate.leave(); return -1; }
	// Requests and answers travel through page-locked host memory that the kernel reads and writes
	// directly: no copy is queued before or after the launch, and the end is seen by watching the
	// status words flip, which costs a few microseconds where a stream synchronisation costs tens.
	memcpy(ln->h_addr, addr, n * 16);
	volatile int32_t *hs = ln->h_status;
	for (size_t i = 0; i < n; i++) hs[i] = SMALL_PENDING;
	GetJob job{};
	job.table = e->table; job.arena = e->arena.base; job.arena_size = e->arena.size;
	job.addr = (const unsigned long long *)ln->h_addr; job.valid = nullptr; job.n = (uint32_t)n; job.nbytes = e->bsize;
	job.out = (uint8_t *)pages_out;                          // device memory or page-locked host memory (UVA)
	job.status = ln->h_status;
	for (int r = 0; r < GET_MAX_PEERS; r++) { job.peer[r] = e->peer_base[r]; job.peer_size[r] = e->peer_size[r]; }
	job.scratch = e->d_scratch; job.region_entries = e->region_entries; job.pool_bits = e->d_pool_bits; job.pool_n = e->pool_n;
	if (launch_get_small(job, ln->st)) { ln->busy.store(0, std::memory_order_release); e->get_gate.leave(); return -1; }
	t->lane = li; t->n = (uint32_t)n; t->status = ln->h_status;
	return 0;
}

extern "C" int cmb200_get_small_end(cmb200_engine *e, cmb200_small_ticket *t, int32_t *status_out) {
	if (!t || t->lane < 0) return 0;
	cmb200_engine::GetLane *ln = &e->lane[t->lane];
	volatile int32_t *hs = ln->h_status;
	int rc = 0;
	uint32_t done = 0;
	for (uint64_t spins = 0; done < t->n;) {
		if (hs[done] != SMALL_PENDING) { done++; continue; }
#if defined(__x86_64__)
		__builtin_ia32_pause();
#endif
		if (++spins > 20000) {                                // ~1 ms of polling: a large batch, let the driver wait
			if (cudaSetDevice(e->device) != cudaSuccess || cudaStreamSynchronize(ln->st) != cudaSuccess) {
				cmb_set_error("cudaStreamSynchronize(small get)", cudaGetLastError(), __FILE__, __LINE__);
				rc = -1; break;
			}
			spins = 0;
			if (hs[done] == SMALL_PENDING) { set_error_msg("cmb200_get_small: kernel finished without an answer"); rc = -1; break; }
		}
	}
	__atomic_thread_fence(__ATOMIC_ACQUIRE);
	if (rc == 0) {
		uint64_t rq = 0, ht = 0;
		for (uint32_t i = 0; i < t->n; i++) {
			const int32_t st = hs[i];
			if (status_out) status_out[i] = st;
			if (st != CMB200_INVALID) rq++;
			if (st == CMB200_HIT) ht++;
		}
		e->small_get_launches++;
		e->small_get_requests += rq; e->small_get_hits += ht;
	}
	t->lane = -1;
	ln->busy.store(0, std::memory_order_release);
	e->get_gate.leave();
	return rc;
}

extern "C" int cmb200_get_small(cmb200_engine *e, size_t n, const cmb200_addr *addr, void *pages_out, int32_t *status_out) {
	for (size_t at = 0; at < n; at += cmb200_engine::GET_SMALL_MAX) {
		const size_t m = n - at < cmb200_engine::GET_SMALL_MAX ? n - at : cmb200_engine::GET_SMALL_MAX;
		cmb200_small_ticket t;
		const int rc = cmb200_get_small_begin(e, m, addr + at, (uint8_t *)pages_out + at * e->bsize, &t);
		if (rc) return rc;
		if (cmb200_get_small_end(e, &t, status_out + at)) return -1;
	}
	return 0;
}

// ---- peers: the other ranks' arenas, mapped for NVLink reads -------------------------------------

extern "C" int cmb200_close_peers(cmb200_engine *e) {
	cmb200_engine::GateClosed g(e->get_gate);             // no small get in flight
	CMB_CHECK(cudaSetDevice(e->device));
	for (int r = 0; r < GET_MAX_PEERS; r++) {
		if (e->peer_base[r]) cudaIpcCloseMemHandle((void *)e->peer_base[r]);
		e->peer_base[r] = nullptr; e->peer_size[r] = 0;
	}
	return 0;
}

extern "C" int cmb200_arena_ipc_handle(cmb200_engine *e, void *handle64, uint64_t *arena_bytes_out) {
	CMB_CHECK(cudaSetDevice(e->device));
	cudaIpcMemHandle_t h;
	static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
	CMB_CHECK(cudaIpcGetMemHandle(&h, e->arena.base));
	memcpy(handle64, &h, 64);
	if (arena_bytes_out) *arena_bytes_out = e->arena.size;
	return 0;
}

extern "C" int cmb200_open_peer(cmb200_engine *e, uint32_t rank, const void *handle64, uint64_t arena_bytes) {
	if (rank >= GET_MAX_PEERS) { set_error_msg("cmb200_open_peer: rank out of range"); return -1; }
	cmb200_engine::GateClosed g(e->get_gate);
	CMB_CHECK(cudaSetDevice(e->device));
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, 64);
	void *p = nullptr;
	CMB_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
	if (e->peer_base[rank]) cudaIpcCloseMemHandle((void *)e->peer_base[rank]);
	e->peer_base[rank] = (const uint8_t *)p;
	e->peer_size[rank] = arena_bytes;
	return 0;
}

// ---- unset / entries / sample / records ----------------------------------------------------

extern "C" int cmb200_unset_batch(cmb200_engine *e, size_t n, const cmb200_addr *addr) {
	std::lock_guard<std::mutex> g(e->mu);
	CMB_CHECK(cudaSetDevice(e->device));
	for (size_t at = 0; at < n; at += e->max_batch) {
		uint32_t m = (uint32_t)((n - at < e->max_batch) ? n - at : e->max_batch);
		CMB_CHECK(cudaMemcpyAsync(e->d_addr, addr + at, (size_t)m * 16, cudaMemcpyHostToDevice, e->st));
		if (launch_unset(e->table, e->arena, e->d_addr, m, e->st)) return -1;
		e->stats.kernel_launches++;
	}
	CMB_CHECK(cudaStreamSynchronize(e->st));
	return 0;
}

static int read_counters(cmb200_engine *e, unsigned long long out[8]) {
	CMB_CHECK(cudaSetDevice(e->device));
	CMB_CHECK(cudaMemcpyAsync(out, e->d_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->st));
	CMB_CHECK(cudaStreamSynchronize(e->st));
	return 0;
}

extern "C" uint64_t cmb200_entries(cmb200_engine *e) {
	std::lock_guard<std::mutex> g(e->mu);
	unsigned long long c[8];
	if (read_counters(e, c)) return 0;
	return c[0];
}

extern "C" int cmb200_get_stats(cmb200_engine *e, cmb200_stats *out) {
	std::lock_guard<std::mutex> g(e->mu);
	unsigned long long c[8];
	if (read_counters(e, c)) return -1;
	harvest_pending(e, true);
	*out = e->stats;
	out->get_requests += e->small_get_requests.load(); out->get_hits += e->small_get_hits.load();
	out->kernel_launches += e->small_get_launches.load();
	out->entries = c[0]; out->tombstones = c[1];
	out->arena_used = c[2] < e->arena.size ? c[2] : e->arena.size;   // the bump pointer saturates past the end (no rollback)
	out->arena_garbage = c[3];
	out->dropped_puts = c[4]; out->remote_entries = c[5];
	out->table_slots = e->table.cap; out->arena_bytes = e->arena.size;
	return 0;
}

extern "C" int cmb200_sample(cmb200_engine *e, size_t n, const uint64_t *r, cmb200_addr *addr_out,
    uint64_t *ts_out, int32_t *ok_out) {
	std::lock_guard<std::mutex> g(e->mu);
	CMB_CHECK(cudaSetDevice(e->device));
	for (size_t at = 0; at < n; at += e->max_batch) {
		const uint32_t m = (uint32_t)((n - at < e->max_batch) ? n - at : e->max_batch);
		CMB_CHECK(cudaMemcpyAsync(e->d_recoff, r + at, (size_t)m * 8, cudaMemcpyHostToDevice, e->st));
		if (launch_sample(e->table, (const unsigned long long *)e->d_recoff, m, e->d_addr, e->d_ts, e->d_status, e->st)) return -1;
		CMB_CHECK(cudaMemcpyAsync(addr_out + at, e->d_addr, (size_t)m * 16, cudaMemcpyDeviceToHost, e->st));
		CMB_CHECK(cudaMemcpyAsync(ts_out + at, e->d_ts, (size_t)m * 8, cudaMemcpyDeviceToHost, e->st));
		CMB_CHECK(cudaMemcpyAsync(ok_out + at, e->d_status, (size_t)m * 4, cudaMemcpyDeviceToHost, e->st));
		CMB_CHECK(cudaStreamSynchronize(e->st));
		e->stats.kernel_launches++;
	}
	return 0;
}

extern "C" int cmb200_read_records(cmb200_engine *e, size_t n, const cmb200_addr *addr, void *out_host,
    size_t stride, int32_t *len_out) {
	std::lock_guard<std::mutex> g(e->mu);
	CMB_CHECK(cudaSetDevice(e->device));
	std::vector<int32_t> st(e->max_batch);
	std::vector<uint64_t> off(e->max_batch);
	std::vector<uint32_t> vl(e->max_batch);
	for (size_t at = 0; at < n; at += e->max_batch) {
		uint32_t m = (uint32_t)((n - at < e->max_batch) ? n - at : e->max_batch);
		CMB_CHECK(cudaMemcpyAsync(e->d_addr, addr + at, (size_t)m * 16, cudaMemcpyHostToDevice, e->st));
		if (launch_lookup(e->table, e->d_addr, nullptr, m, e->d_status, e->d_recoff, e->d_vlen, nullptr, e->st)) return -1;
		CMB_CHECK(cudaMemcpyAsync(st.data(), e->d_status, m * 4, cudaMemcpyDeviceToHost, e->st));
		CMB_CHECK(cudaMemcpyAsync(off.data(), e->d_recoff, m * 8, cudaMemcpyDeviceToHost, e->st));
		CMB_CHECK(cudaMemcpyAsync(vl.data(), e->d_vlen, m * 4, cudaMemcpyDeviceToHost, e->st));
		CMB_CHECK(cudaStreamSynchronize(e->st));
		for (uint32_t i = 0; i < m; i++) {
			if (st[i] != ST_HIT) { len_out[at + i] = -1; continue; }
			uint32_t clen = vl[i] - 1;
			size_t total = 24 + (clen ? clen : e->bsize);
			if (total > stride) { set_error_msg("cmb200_read_records: stride too small"); return -1; }
			CMB_CHECK(cudaMemcpyAsync((uint8_t *)out_host + (at + i) * stride, e->arena.base + off[i], total,
			    cudaMemcpyDeviceToHost, e->st));
			len_out[at + i] = (int32_t)total;
		}
		CMB_CHECK(cudaStreamSynchronize(e->st));
	}
	return 0;
}

extern "C" int cmb200_read_fingerprints(cmb200_engine *e, size_t n, const cmb200_addr *addr, uint64_t *fp_out,
    int32_t *ok_out) {
	std::lock_guard<std::mutex> g(e->mu);
	CMB_CHECK(cudaSetDevice(e->device));
	if (!e->table.fp) { set_error_msg("engine created without CMB200_FINGERPRINT"); return -1; }
	for (size_t at = 0; at < n; at += e->max_batch) {
		uint32_t m = (uint32_t)((n - at < e->max_batch) ? n - at : e->max_batch);
		CMB_CHECK(cudaMemcpyAsync(e->d_addr, addr + at, (size_t)m * 16, cudaMemcpyHostToDevice, e->st));
		if (launch_read_fp(e->table, e->d_addr, m, e->d_fps, e->d_status, e->st)) return -1;
		CMB_CHECK(cudaMemcpyAsync(fp_out + 2 * at, e->d_fps, (size_t)m * 16, cudaMemcpyDeviceToHost, e->st));
		CMB_CHECK(cudaMemcpyAsync(ok_out + at, e->d_status, (size_t)m * 4, cudaMemcpyDeviceToHost, e->st));
		CMB_CHECK(cudaStreamSynchronize(e->st));
	}
	return 0;
}

struct DevBuf {
	void *p = nullptr;
	~DevBuf() { if (p) cudaFree(p); }
	int alloc(size_t bytes) { CMB_CHECK(cudaMalloc(&p, bytes + 256)); return 0; }
	template <class T> T *as() { return (T *)p; }
};

// ---- snapshot: persistence of the cache directory ---------------------------
//
// The reference's store is persistent because it IS a set of LMDB files under <cachedir>
// (filemap.c:57,71-72).  Here the store lives in HBM, so it is saved to / restored from one file of
// records, each exactly the LMDB value of the reference (24-byte data_prefix + payload,
// filemap.c:140-147) preceded by {ts (the LMDB attribute), length, fingerprint}:
//
//   header   "CMB200S1" | u32 version=1 | u32 pshift | u64 records | u64 payload bytes | u32 flags | pad to 64
//   record   u64 ts | u64 fp_hi | u64 fp_lo | u32 len | u32 0 | len bytes {u, l, compressed_length, pad, payload} | pad to 16
//
// It does not depend on the table geometry or the arena layout, so a snapshot loads into an engine
// of any capacity (records that do not fit are dropped like puts into a full store).
struct SnapHeader {
	char magic[8];
	uint32_t version, pshift;
	uint64_t records, bytes;
	uint32_t flags, pad[7];
};
static_assert(sizeof(SnapHeader) == 64, "snapshot header");
struct SnapRecord { uint64_t ts, fp_hi, fp_lo; uint32_t len, zero; };
static_assert(sizeof(SnapRecord) == 32, "snapshot record header");
static const size_t SNAP_WINDOW = 64u << 20;

extern "C" int cmb200_save(cmb200_engine *e, const char *path, uint64_t *records_out) {
	std::lock_guard<std::mutex> g(e->mu);
	unsigned long long c[8];
	if (read_counters(e, c)) return -1;
	const unsigned long long cap_out = c[0] + 16;
	DevBuf d_list, d_count;
	if (d_list.alloc(cap_out * sizeof(ExportEntry)) || d_count.alloc(8)) return -1;
	CMB_CHECK(cudaMemsetAsync(d_count.p, 0, 8, e->st));
	if (launch_export_list(e->table, e->bsize, d_list.as<ExportEntry>(), d_count.as<unsigned long long>(), cap_out, e->st)) return -1;
	unsigned long long count = 0;
	CMB_CHECK(cudaMemcpyAsync(&count, d_count.p, 8, cudaMemcpyDeviceToHost, e->st));
	CMB_CHECK(cudaStreamSynchronize(e->st));
	if (count > cap_out) count = cap_out;
	std::vector<ExportEntry> list(count);
	if (count) CMB_CHECK(cudaMemcpy(list.data(), d_list.p, count * sizeof(ExportEntry), cudaMemcpyDeviceToHost));
	std::sort(list.begin(), list.end(), [](const ExportEntry &a, const ExportEntry &b) { return a.rec_off < b.rec_off; });

	std::string tmp = std::string(path) + ".tmp";
	FILE *f = fopen(tmp.c_str(), "wb");
	if (!f) { set_error_msg("cmb200_save: cannot create the snapshot file"); return -1; }
	SnapHeader h{};
	memcpy(h.magic, "CMB200S1", 8);
	h.version = 1; h.pshift = (uint32_t)e->pshift; h.records = count; h.flags = e->table.fp ? 1u : 0u;
	for (const ExportEntry &x : list) h.bytes += x.len;
	bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
	uint8_t *win = nullptr;
	if (cudaMallocHost(&win, SNAP_WINDOW) != cudaSuccess) { fclose(f); remove(tmp.c_str()); set_error_msg("cmb200_save: no page-locked window"); return -1; }
	static const uint8_t zeros[16] = {0};
	size_t k = 0;
	while (ok && k < list.size()) {
		// one window of the arena starting at record k; the records wholly inside it are written out
		const unsigned long long w0 = list[k].rec_off;
		unsigned long long w1 = w0 + SNAP_WINDOW;
		if (w1 > e->arena.size + 256) w1 = e->arena.size + 256;
		if (cudaMemcpyAsync(win, e->arena.base + w0, (size_t)(w1 - w0), cudaMemcpyDeviceToHost, e->st) != cudaSuccess ||
		    cudaStreamSynchronize(e->st) != cudaSuccess) { ok = false; break; }
		for (; k < list.size() && list[k].rec_off + list[k].len <= w1; k++) {
			const ExportEntry &x = list[k];
			SnapRecord r{x.ts, x.fp_hi, x.fp_lo, x.len, 0};
			const size_t padn = (16 - (x.len & 15)) & 15;
			ok = ok && fwrite(&r, sizeof(r), 1, f) == 1 && fwrite(win + (x.rec_off - w0), x.len, 1, f) == 1 &&
			    (padn == 0 || fwrite(zeros, padn, 1, f) == 1);
		}
	}
	cudaFreeHost(win);
	ok = ok && fflush(f) == 0;
	ok = (fclose(f) == 0) && ok;
	if (!ok || rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); set_error_msg("cmb200_save: write failed"); return -1; }
	if (records_out) *records_out = count;
	return 0;
}

extern "C" int cmb200_load(cmb200_engine *e, const char *path, uint64_t *records_out) {
	std::lock_guard<std::mutex> g(e->mu);
	CMB_CHECK(cudaSetDevice(e->device));
	if (records_out) *records_out = 0;
	FILE *f = fopen(path, "rb");
	if (!f) { set_error_msg("cmb200_load: no snapshot file"); return -1; }
	SnapHeader h{};
	if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "CMB200S1", 8) != 0 || h.version != 1) {
		fclose(f); set_error_msg("cmb200_load: not a snapshot of this library"); return -1;
	}
	if ((int)h.pshift != e->pshift) { fclose(f); set_error_msg("cmb200_load: snapshot has another page size"); return -1; }
	// batches: <= max_batch records and <= one page-ring buffer of bytes
	const size_t blob_cap = (size_t)e->host_batch * e->bsize;
	const size_t B = e->max_batch < cmb200_engine::META_CAP ? e->max_batch : cmb200_engine::META_CAP;
	uint8_t *blob = nullptr;
	if (cudaMallocHost(&blob, blob_cap) != cudaSuccess) { fclose(f); set_error_msg("cmb200_load: no page-locked buffer"); return -1; }
	std::vector<unsigned long long> off(B), ts(B), fps(2 * B);
	std::vector<cmb200_addr> addr(B);
	DevBuf d_off, d_fp;
	int rc = 0;
	if (d_off.alloc(B * 8) || d_fp.alloc(B * 16)) rc = -1;
	uint64_t done = 0, loaded = 0;
	SnapRecord pending{}; bool have_pending = false;
	while (rc == 0 && done < h.records) {
		size_t m = 0, used = 0;
		while (m < B && done + m < h.records) {
			SnapRecord r;
			if (have_pending) { r = pending; have_pending = false; }
			else if (fread(&r, sizeof(r), 1, f) != 1) { rc = -1; break; }
			const size_t padded = ((size_t)r.len + 15) & ~(size_t)15;
			if (r.len < 24 || r.len > 24u + e->bsize + 1024u) { rc = -1; break; }
			if (used + padded > blob_cap) { pending = r; have_pending = true; break; }
			if (fread(blob + used, padded, 1, f) != 1) { rc = -1; break; }
			{
				// the record must be what filemap_set would have stored (filemap.c:124-147): a foreign or
				// corrupt file must not reach k_restore, which trusts compressed_length
				int32_t clen;
				memcpy(&clen, blob + used + 16, 4);
				if (clen < 0 || (uint32_t)clen > e->bsize + 1024u || r.len != 24u + (clen ? (uint32_t)clen : e->bsize)) { rc = -1; break; }
			}
			memcpy(&addr[m], blob + used, 16);      // data_prefix {u, l}
			off[m] = used; ts[m] = r.ts; fps[2 * m] = r.fp_hi; fps[2 * m + 1] = r.fp_lo;
			used += padded; m++;
		}
		if (rc) { set_error_msg("cmb200_load: truncated or corrupt snapshot"); break; }
		if (m == 0) { rc = -1; set_error_msg("cmb200_load: record larger than the staging buffer"); break; }
		const bool fail =
		    cudaMemcpyAsync(e->d_pages[0], blob, used, cudaMemcpyHostToDevice, e->st) != cudaSuccess ||
		    cudaMemcpyAsync(e->d_addr, addr.data(), m * 16, cudaMemcpyHostToDevice, e->st) != cudaSuccess ||
		    cudaMemcpyAsync(e->d_ts, ts.data(), m * 8, cudaMemcpyHostToDevice, e->st) != cudaSuccess ||
		    cudaMemcpyAsync(d_off.p, off.data(), m * 8, cudaMemcpyHostToDevice, e->st) != cudaSuccess ||
		    cudaMemcpyAsync(d_fp.p, fps.data(), m * 16, cudaMemcpyHostToDevice, e->st) != cudaSuccess;
		if (fail || launch_upsert(e->table, e->d_addr, nullptr, (uint32_t)m, e->seq, e->seq_stride, e->d_slot, e->st)) { rc = -1; break; }
		EncodeJob job{};
		job.n = (uint32_t)m; job.nbytes = e->bsize;
		job.slot_idx = e->d_slot; job.addr = e->d_addr; job.ts = e->d_ts;
		job.seq0 = e->seq; job.seq_stride = e->seq_stride;
		job.table = e->table; job.arena = e->arena;
		if (launch_restore(job, e->d_pages[0], d_off.as<unsigned long long>(), (h.flags & 1u) ? d_fp.as<uint64_t>() : nullptr,
			e->bsize, e->st)) { rc = -1; break; }
		if (cudaStreamSynchronize(e->st) != cudaSuccess) { rc = -1; break; }
		e->seq += (unsigned long long)m * e->seq_stride;
		e->stats.kernel_launches += 2;
		done += m; loaded += m;
	}
	cudaFreeHost(blob);
	fclose(f);
	if (records_out) *records_out = loaded;
	return rc;
}

// ---- arena compaction -------------------------------------------------------------------------
// The arena is a bump allocator: a deleted record, or one that outgrew its place, leaves its bytes
// behind as garbage.  Compaction slides the live records down to the start of the arena (sorted by
// offset, so every record moves to a lower or equal address), window by window through one of the
// page-ring buffers, repoints the slots and resets the bump pointer and the per-warp segments.
// Stop-the-world on the engine's stream, at HBM speed; callers trigger it when the arena is about
// to overflow although a good part of it is garbage (filemap_make_room, cmb200_compact).
extern "C" int cmb200_compact(cmb200_engine *e, uint64_t *reclaimed_out) {
	std::lock_guard<std::mutex> g(e->mu);
	cmb200_engine::GateClosed gg(e->get_gate);          // records move: no small get may be reading the arena
	unsigned long long c[8];
	if (read_counters(e, c)) return -1;
	harvest_pending(e, true);
	const unsigned long long head_before = c[2];
	const unsigned long long cap_out = c[0] + 16;
	DevBuf d_list, d_count, d_moves;
	if (d_list.alloc(cap_out * sizeof(ExportEntry)) || d_count.alloc(8)) return -1;
	CMB_CHECK(cudaMemsetAsync(d_count.p, 0, 8, e->st));
	if (launch_export_list(e->table, e->bsize, d_list.as<ExportEntry>(), d_count.as<unsigned long long>(), cap_out, e->st)) return -1;
	unsigned long long count = 0;
	CMB_CHECK(cudaMemcpyAsync(&count, d_count.p, 8, cudaMemcpyDeviceToHost, e->st));
	CMB_CHECK(cudaStreamSynchronize(e->st));
	if (count > cap_out) { set_error_msg("cmb200_compact: the store changed under the compaction"); return -1; }
	std::vector<ExportEntry> list(count);
	if (count) CMB_CHECK(cudaMemcpy(list.data(), d_list.p, count * sizeof(ExportEntry), cudaMemcpyDeviceToHost));
	std::sort(list.begin(), list.end(), [](const ExportEntry &a, const ExportEntry &b) { return a.rec_off < b.rec_off; });
	std::vector<MoveEntry> moves(count);
	unsigned long long at = 0;
	for (size_t i = 0; i < count; i++) {
		moves[i] = MoveEntry{list[i].rec_off, at, list[i].len, list[i].slot};
		at += ((unsigned long long)list[i].len + 15ull) & ~15ull;
	}
	if (count && d_moves.alloc(count * sizeof(MoveEntry))) return -1;
	if (count) CMB_CHECK(cudaMemcpyAsync(d_moves.p, moves.data(), count * sizeof(MoveEntry), cudaMemcpyHostToDevice, e->st));
	// windows: as many records as fit the bounce buffer (one page-ring buffer)
	const unsigned long long bounce_cap = (unsigned long long)e->host_batch * e->bsize;
	size_t k = 0;
	while (k < count) {
		size_t j = k;
		while (j < count && moves[j].new_off + (((unsigned long long)moves[j].len + 15ull) & ~15ull) - moves[k].new_off <= bounce_cap) j++;
		if (j == k) { set_error_msg("cmb200_compact: record larger than the bounce buffer"); return -1; }
		if (launch_compact_window(e->table, e->arena, d_moves.as<MoveEntry>() + k, (uint32_t)(j - k), e->d_pages[0], e->st)) return -1;
		e->stats.kernel_launches += 2;
		k = j;
	}
	// bump pointer back to the end of the live records, no garbage, no half-used segments
	unsigned long long fresh[2] = {at, 0};
	CMB_CHECK(cudaMemcpyAsync(e->d_counters + 2, fresh, 16, cudaMemcpyHostToDevice, e->st));
	CMB_CHECK(cudaMemsetAsync(e->arena.seg, 0, ARENA_SEG_SLOTS * 2 * sizeof(unsigned long long), e->st));
	CMB_CHECK(cudaStreamSynchronize(e->st));
	if (reclaimed_out) *reclaimed_out = head_before > at ? head_before - at : 0;
	// The table is rebuilt on the same occasion when deleted keys have eaten a good part of its empty
	// slots (tombstones; slots of dropped puts go with them): miss probes end at an EMPTY slot, and
	// linear probing never frees one by itself.
	if (c[1] > e->table.cap / 8) {
		TableView fresh = e->table;
		fresh.slots = nullptr; fresh.fp = nullptr;
		if (cudaMalloc(&fresh.slots, (e->table.cap + 2) * sizeof(Slot)) == cudaSuccess &&
		    (!e->table.fp || cudaMalloc(&fresh.fp, (e->table.cap + 2) * 16) == cudaSuccess)) {
			CMB_CHECK(cudaMemsetAsync(fresh.slots, 0, (e->table.cap + 2) * sizeof(Slot), e->st));
			if (fresh.fp) CMB_CHECK(cudaMemsetAsync(fresh.fp, 0, (e->table.cap + 2) * 16, e->st));
			if (launch_rehash(e->table, fresh, e->st)) return -1;
			// the slots have moved: their parse checkpoints are dropped (those records are walked by one warp)
			if (e->table.ckpt) CMB_CHECK(cudaMemsetAsync(e->table.ckpt, 0, (e->table.cap + 2) * CKPT_WORDS * 4, e->st));
			CMB_CHECK(cudaMemsetAsync(e->d_counters + 1, 0, sizeof(unsigned long long), e->st));   // tombstones
			CMB_CHECK(cudaStreamSynchronize(e->st));
			cudaFree(e->table.slots); cudaFree(e->table.fp);
			e->table.slots = fresh.slots; e->table.fp = fresh.fp;
			e->stats.kernel_launches++;
		} else {
			(void)cudaGetLastError();                    // no room for a second table: keep the old one
			cudaFree(fresh.slots);
		}
	}
	return 0;
}

// ---- kernel-level entry points -------------------------------------------------------------

extern "C" int cmb200_compose_keys(int device, size_t n, const uint64_t *offset, const uint64_t *nhid,
    const uint32_t *genid, int pshift, cmb200_addr *addr_out, uint8_t *valid_out, uint64_t *key_out) {
	if (select_device(device)) return -1;
	DevBuf d_off, d_nh, d_g, d_addr, d_valid, d_key;
	if (d_off.alloc(n * 8) || d_nh.alloc(n * 8) || d_g.alloc(n * 4) || d_addr.alloc(n * 16) || d_valid.alloc(n) || d_key.alloc(n * 8)) return -1;
	CMB_CHECK(cudaMemcpy(d_off.p, offset, n * 8, cudaMemcpyHostToDevice));
	CMB_CHECK(cudaMemcpy(d_nh.p, nhid, n * 8, cudaMemcpyHostToDevice));
	CMB_CHECK(cudaMemcpy(d_g.p, genid, n * 4, cudaMemcpyHostToDevice));
	if (launch_compose(d_off.as<uint64_t>(), d_nh.as<uint64_t>(), d_g.as<uint32_t>(), pshift, (uint32_t)n,
		d_addr.as<unsigned long long>(), d_valid.as<uint8_t>(), d_key.as<unsigned long long>(), 0)) return -1;
	CMB_CHECK(cudaMemcpy(addr_out, d_addr.p, n * 16, cudaMemcpyDeviceToHost));
	CMB_CHECK(cudaMemcpy(valid_out, d_valid.p, n, cudaMemcpyDeviceToHost));
	CMB_CHECK(cudaMemcpy(key_out, d_key.p, n * 8, cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int cmb200_lz4_encode_batch(int device, const void *pages_host, size_t n, uint32_t nbytes,
    size_t stride, int accel, void *blocks_out_host, size_t out_stride, int32_t *lens_out, uint64_t *fp_out) {
	if (select_device(device)) return -1;
	if (stride % 16 || stride < nbytes) { set_error_msg("encode_batch: stride must be a multiple of 16 and >= nbytes"); return -1; }
	size_t bound = (size_t)nbytes + nbytes / 255 + 16;
	size_t sstride = (bound + 15) & ~(size_t)15;
	if (out_stride < bound) { set_error_msg("encode_batch: out_stride below LZ4_compressBound"); return -1; }
	DevBuf d_in, d_stage, d_lens, d_fps, d_work;
	if (d_in.alloc(n * stride) || d_stage.alloc(n * sstride) || d_lens.alloc(n * 4) || d_fps.alloc(n * 16) || d_work.alloc(64)) return -1;
	CMB_CHECK(cudaMemcpy(d_in.p, pages_host, n * stride, cudaMemcpyHostToDevice));
	EncodeJob job{};
	job.pages = d_in.as<uint8_t>(); job.page_stride = stride; job.nbytes = nbytes; job.n = (uint32_t)n;
	job.accel = accel < 0 ? 1u : (uint32_t)(accel > (1 << 20) ? (1 << 20) : accel);
	if (accel == 0) job.accel = 1;   // LZ4_compress_fast(accel<1) -> 1 (lz4.c:740); raw mode is a store-level notion
	job.stage = d_stage.as<uint8_t>(); job.stage_stride = sstride;
	job.lens = d_lens.as<int32_t>();
	job.fps = fp_out ? d_fps.as<uint64_t>() : nullptr;
	job.work = d_work.as<unsigned int>();
	if (launch_encode(job, 0)) return -1;
	CMB_CHECK(cudaMemcpy(lens_out, d_lens.p, n * 4, cudaMemcpyDeviceToHost));
	if (fp_out) CMB_CHECK(cudaMemcpy(fp_out, d_fps.p, n * 16, cudaMemcpyDeviceToHost));
	CMB_CHECK(cudaMemcpy2D(blocks_out_host, out_stride, d_stage.p, sstride, bound < out_stride ? bound : out_stride, n,
	    cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int cmb200_lz4_decode_batch(int device, const void *blocks_host, size_t in_stride, const int32_t *lens,
    size_t n, uint32_t nbytes, void *pages_out_host, int32_t *consumed_out) {
	if (select_device(device)) return -1;
	DevBuf d_blk, d_lens, d_out, d_used;
	if (d_blk.alloc(n * in_stride) || d_lens.alloc(n * 4) || d_out.alloc(n * (size_t)nbytes) || d_used.alloc(n * 4)) return -1;
	CMB_CHECK(cudaMemcpy(d_blk.p, blocks_host, n * in_stride, cudaMemcpyHostToDevice));
	CMB_CHECK(cudaMemcpy(d_lens.p, lens, n * 4, cudaMemcpyHostToDevice));
	CMB_CHECK(cudaMemset(d_out.p, 0, n * (size_t)nbytes));
	DecodeJob job{};
	job.n = (uint32_t)n; job.nbytes = nbytes; job.pages = d_out.as<uint8_t>(); job.status = d_used.as<int32_t>();
	job.blocks = d_blk.as<uint8_t>(); job.block_stride = in_stride; job.lens = d_lens.as<int32_t>();
	if (launch_decode(job, 0)) return -1;
	CMB_CHECK(cudaMemcpy(consumed_out, d_used.p, n * 4, cudaMemcpyDeviceToHost));
	CMB_CHECK(cudaMemcpy(pages_out_host, d_out.p, n * (size_t)nbytes, cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int cmb200_fingerprint_batch(int device, const void *pages_host, size_t n, uint32_t nbytes,
    size_t stride, uint64_t *fp_out) {
	if (select_device(device)) return -1;
	if (stride % 16) { set_error_msg("fingerprint_batch: stride must be a multiple of 16"); return -1; }
	DevBuf d_in, d_fps;
	if (d_in.alloc(n * stride) || d_fps.alloc(n * 16)) return -1;
	CMB_CHECK(cudaMemcpy(d_in.p, pages_host, n * stride, cudaMemcpyHostToDevice));
	if (launch_fingerprint(d_in.as<uint8_t>(), stride, nbytes, (uint32_t)n, d_fps.as<uint64_t>(), 0)) return -1;
	CMB_CHECK(cudaMemcpy(fp_out, d_fps.p, n * 16, cudaMemcpyDeviceToHost));
	return 0;
}

// EF128 of n resident pages (stand-alone form of the pass that k_encode fuses); the kernel's
// CUDA-event time is added to stats.fingerprint_kernel_ns.
extern "C" int cmb200_fingerprint_dev(cmb200_engine *e, size_t n, const void *pages_dev, uint64_t *fp_out_host) {
	std::lock_guard<std::mutex> g(e->mu);
	CMB_CHECK(cudaSetDevice(e->device));
	DevBuf d_fps;
	if (d_fps.alloc(n * 16)) return -1;
	CMB_CHECK(cudaEventRecord(e->t0[0], e->st));
	if (launch_fingerprint((const uint8_t *)pages_dev, e->bsize, e->bsize, (uint32_t)n, d_fps.as<uint64_t>(), e->st)) return -1;
	CMB_CHECK(cudaEventRecord(e->t1[0], e->st));
	CMB_CHECK(cudaMemcpyAsync(fp_out_host, d_fps.p, n * 16, cudaMemcpyDeviceToHost, e->st));
	CMB_CHECK(cudaStreamSynchronize(e->st));
	float ms = 0;
	CMB_CHECK(cudaEventElapsedTime(&ms, e->t0[0], e->t1[0]));
	e->stats.fingerprint_kernel_ns += (uint64_t)(ms * 1e6);
	e->stats.kernel_launches++;
	return 0;
}

// ---- synthetic streams -----------------------------------------------------------------------

extern "C" void cmb200_gen_chunk_host(uint64_t seed, uint64_t cid, uint32_t bsize, void *out) {
	uint64_t *w = (uint64_t *)out;
	for (uint32_t i = 0; i < bsize / 8; i++) w[i] = sg_chunk_word(seed, cid, bsize, i);
}

extern "C" int cmb200_gen_chunks_dev(cmb200_engine *e, uint64_t seed, const uint64_t *cids_host, size_t n, void *out_dev) {
	std::lock_guard<std::mutex> g(e->mu);
	CMB_CHECK(cudaSetDevice(e->device));
	DevBuf d_c;
	if (d_c.alloc(n * 8)) return -1;
	CMB_CHECK(cudaMemcpyAsync(d_c.p, cids_host, n * 8, cudaMemcpyHostToDevice, e->st));
	if (launch_streamgen(d_c.as<uint64_t>(), (uint32_t)n, seed, e->bsize, (uint8_t *)out_dev, e->st)) return -1;
	CMB_CHECK(cudaStreamSynchronize(e->st));
	return 0;
}

extern "C" uint64_t cmb200_gen_stream_ids(uint64_t seed2, size_t n, double dup, uint64_t first_cid, uint64_t *cid_out) {
	uint64_t state = seed2, distinct = 0;
	for (size_t k = 0; k < n; k++) {
		state += SG_GOLDEN;
		uint64_t r = sg_mix(state);
		bool repeat = distinct > 0 && (double)(r >> 11) * (1.0 / 9007199254740992.0) < dup;
		if (repeat) {
			state += SG_GOLDEN;
			cid_out[k] = first_cid + sg_mix(state) % distinct;
		} else {
			cid_out[k] = first_cid + distinct++;
		}
	}
	return distinct;
}

extern "C" void cmb200_gen_addr(uint64_t seed, uint64_t cid, int pshift, uint64_t *offset_out, uint64_t *nhid_out) {
	*offset_out = sg_offset(cid, pshift);
	*nhid_out = sg_nhid(seed, cid);
}
