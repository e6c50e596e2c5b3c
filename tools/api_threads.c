/* cachemap_put / cachemap_get, one 64 KiB page per call, from T native threads — the calls
 * edgefs_read / edgefs_write make from FUSE worker threads — against any libcachemap.so (the
 * drop-in or the reference's own, loaded by path).  No interpreter between the threads and the
 * library (tools/api_threads_bench.py drives the same calls from Python threads and is bounded by
 * the interpreter lock beyond ~8 threads).
 *   api_threads <libcachemap.so> <cachedir> <pages.bin> <threads> <per-thread>
 * pages.bin = 64 distinct 64 KiB pages (content classes mixed).  Prints one JSON line. */
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#define CH 65536
typedef void *(*create_fn)(const char *, uint64_t, int, int);
typedef int (*put_fn)(void *, uint64_t, uint64_t, uint32_t, void *);
typedef void *(*get_fn)(void *, uint64_t, uint64_t, uint32_t);
typedef void (*free_fn)(void *);

static put_fn f_put;
static get_fn f_get;
static void *cm;
static uint8_t *pages;
static int npages, per_thread, do_get;
static long misses;

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void *worker(void *arg) {
	long t = (long)arg, miss = 0;
	for (int i = 0; i < per_thread; i++) {
		uint64_t k = (uint64_t)t * per_thread + i;
		if (do_get) {
			void *p = f_get(cm, k << 16, 0x77, 0);
			if (!p) miss++;
			else {
				if (memcmp(p, pages + (k % npages) * CH, 64) != 0) miss += 1000000;
				free(p);
			}
		} else {
			f_put(cm, k << 16, 0x77, 0, pages + (k % npages) * CH);
		}
	}
	__atomic_fetch_add(&misses, miss, __ATOMIC_RELAXED);
	return NULL;
}

static double run(int threads) {
	pthread_t th[256];
	double t0 = now();
	for (long t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, (void *)t);
	for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
	return now() - t0;
}

int main(int argc, char **argv) {
	if (argc < 6) { fprintf(stderr, "usage: %s lib cachedir pages.bin threads per-thread\n", argv[0]); return 2; }
	void *h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
	if (!h) { fprintf(stderr, "%s\n", dlerror()); return 1; }
	create_fn f_create = (create_fn)dlsym(h, "cachemap_create");
	free_fn f_free = (free_fn)dlsym(h, "cachemap_free");
	f_put = (put_fn)dlsym(h, "cachemap_put");
	f_get = (get_fn)dlsym(h, "cachemap_get");
	FILE *f = fopen(argv[3], "rb");
	if (!f) { perror(argv[3]); return 1; }
	fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
	npages = (int)(sz / CH);
	pages = malloc((size_t)sz);
	if (fread(pages, 1, (size_t)sz, f) != (size_t)sz) return 1;
	fclose(f);
	int threads = atoi(argv[4]);
	per_thread = atoi(argv[5]);
	if (threads > 256) threads = 256;
	cm = f_create(argv[2], 1 << 16, 12, 16);
	if (!cm) { fprintf(stderr, "cachemap_create failed\n"); return 1; }
	f_put(cm, 1ull << 40, 1, 0, pages);              /* engine start outside the clock */
	do_get = 0;
	double tp = run(threads);
	do_get = 1;
	double tg = run(threads);
	double n = (double)threads * per_thread;
	printf("{\"threads\": %d, \"put_gibs\": %.3f, \"get_gibs\": %.3f, \"put_kops\": %.1f, \"get_kops\": %.1f, \"bad\": %ld}\n", threads,
	    n * CH / tp / (1 << 30), n * CH / tg / (1 << 30), n / tp / 1e3, n / tg / 1e3, misses);
	fflush(stdout);
	/* the reference's cachemap_free joins put threads that never leave their wait (cachemap.c:67-105):
	 * its process just ends; the drop-in shuts down in order */
	if (dlsym(h, "cmb200_engine_create")) f_free(cm);
	_exit(0);
}
