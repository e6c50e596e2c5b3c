/*
 * cpu_bench.c — multi-threaded CPU timing harness for the reference path.
 * TEST / BASELINE INFRASTRUCTURE ONLY (see ef_oracle.h): used by bench.py's cpu_baseline leg and
 * by `bench.py --impl reference`.  It times function pointers handed in by the caller — the
 * reference's own LZ4_compress_fast / LZ4_decompress_fast / cachemap_put / cachemap_get from
 * oracle/_ref/libcachemap_ref.so (kind "reference"), or the oracle port (kind "port") — on
 * T pthreads over disjoint chunk ranges, wall clock CLOCK_MONOTONIC.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int (*enc_fn)(const char *src, char *dst, int n, int cap, int accel);
typedef int (*dec_fn)(const char *src, char *dst, int n);
typedef void (*put_fn)(void *cm, uint64_t off, uint64_t nhid, uint32_t genid, const void *page);
typedef void *(*get_fn)(void *cm, uint64_t off, uint64_t nhid, uint32_t genid);

static double
now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + ts.tv_nsec * 1e-9;
}

struct job {
	int mode;               /* 0 encode, 1 decode, 2 put, 3 get */
	const uint8_t *data;    /* nchunks * bsize */
	uint8_t *blocks;        /* nchunks * stride (codec modes) */
	int *lens;
	size_t first, count;
	int bsize, accel, stride;
	enc_fn enc; dec_fn dec; put_fn put; get_fn get;
	void *cm;
	const uint64_t *offs, *nhids;
	uint64_t bad;           /* mismatches / misses observed */
	pthread_barrier_t *bar;
	double t0, t1;
};

static void *
worker(void *arg)
{
	struct job *j = arg;
	uint8_t *scratch = malloc((size_t)j->bsize + 64);
	pthread_barrier_wait(j->bar);
	j->t0 = now_s();
	for (size_t i = j->first; i < j->first + j->count; i++) {
		const uint8_t *page = j->data + i * (size_t)j->bsize;
		uint8_t *blk = j->blocks ? j->blocks + i * (size_t)j->stride : NULL;
		switch (j->mode) {
		case 0:
			j->lens[i] = j->enc((const char *)page, (char *)blk, j->bsize, j->stride, j->accel);
			break;
		case 1: {
			int used = j->dec((const char *)blk, (char *)scratch, j->bsize);
			if (used != j->lens[i] || memcmp(scratch, page, j->bsize) != 0)
				j->bad++;
			break; }
		case 2:
			j->put(j->cm, j->offs[i], j->nhids[i], 0, page);
			break;
		case 3: {
			void *p = j->get(j->cm, j->offs[i], j->nhids[i], 0);
			if (!p || memcmp(p, page, j->bsize) != 0)
				j->bad++;
			free(p);
			break; }
		}
	}
	j->t1 = now_s();
	free(scratch);
	return NULL;
}

/* Runs one mode over nchunks chunks on `threads` pthreads; returns wall seconds (first start to
 * last finish) and the number of bad results through *bad. */
static double
run_mode(struct job *proto, size_t nchunks, int threads, uint64_t *bad)
{
	pthread_t tid[256];
	struct job jobs[256];
	pthread_barrier_t bar;
	if (threads < 1) threads = 1;
	if (threads > 256) threads = 256;
	if ((size_t)threads > nchunks) threads = (int)nchunks;
	pthread_barrier_init(&bar, NULL, threads);
	size_t per = nchunks / threads, extra = nchunks % threads, at = 0;
	for (int t = 0; t < threads; t++) {
		jobs[t] = *proto;
		jobs[t].first = at;
		jobs[t].count = per + ((size_t)t < extra);
		at += jobs[t].count;
		jobs[t].bar = &bar;
		jobs[t].bad = 0;
		pthread_create(&tid[t], NULL, worker, &jobs[t]);
	}
	double t0 = 1e300, t1 = 0;
	*bad = 0;
	for (int t = 0; t < threads; t++) {
		pthread_join(tid[t], NULL);
		if (jobs[t].t0 < t0) t0 = jobs[t].t0;
		if (jobs[t].t1 > t1) t1 = jobs[t].t1;
		*bad += jobs[t].bad;
	}
	pthread_barrier_destroy(&bar);
	return t1 - t0;
}

/* Codec-only figure.  out[0]=encode s, out[1]=decode s, out[2]=bad count,
 * out[3]=total compressed bytes. */
void
ef_cpu_bench_codec(enc_fn enc, dec_fn dec, const uint8_t *data, size_t nchunks, int bsize,
    int accel, int threads, double out[4])
{
	int stride = bsize + 1024;
	uint8_t *blocks = malloc(nchunks * (size_t)stride);
	int *lens = calloc(nchunks, sizeof(int));
	memset(blocks, 0, nchunks * (size_t)stride);    /* fault the pages in before the clock starts */
	struct job p;
	uint64_t bad = 0;
	memset(&p, 0, sizeof(p));
	p.data = data; p.blocks = blocks; p.lens = lens; p.bsize = bsize; p.accel = accel;
	p.stride = stride; p.enc = enc; p.dec = dec;
	p.mode = 0;
	out[0] = run_mode(&p, nchunks, threads, &bad);
	p.mode = 1;
	out[1] = run_mode(&p, nchunks, threads, &bad);
	out[2] = (double)bad;
	double tot = 0;
	for (size_t i = 0; i < nchunks; i++) tot += lens[i];
	out[3] = tot;
	free(blocks);
	free(lens);
}

/* Full-path figure: cachemap_put then cachemap_get of every chunk through an
 * already-created reference cachemap.  out[0]=put s, out[1]=get s, out[2]=bad gets. */
void
ef_cpu_bench_store(put_fn put, get_fn get, void *cm, const uint8_t *data, size_t nchunks,
    int bsize, const uint64_t *offs, const uint64_t *nhids, int threads, int do_get, double out[3])
{
	struct job p;
	uint64_t bad = 0;
	memset(&p, 0, sizeof(p));
	p.data = data; p.bsize = bsize; p.put = put; p.get = get; p.cm = cm;
	p.offs = offs; p.nhids = nhids;
	p.mode = 2;
	out[0] = run_mode(&p, nchunks, threads, &bad);
	out[1] = 0; out[2] = 0;
	if (do_get) {
		p.mode = 3;
		out[1] = run_mode(&p, nchunks, threads, &bad);
		out[2] = (double)bad;
	}
}
