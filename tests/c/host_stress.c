/* Stress of the drop-in's HOST layer (cachemap_api.c) over the CPU stand-in of the engine
 * (mock_engine.c): T threads mix cachemap_put / cachemap_get / cachemap_read_range on keys of their
 * own and on shared keys, and check what the API promises —
 *   - read-your-writes: a get after a put of the same thread returns that page (write-behind ring
 *     first, store after the flush), for keys nobody else writes;
 *   - a page is never torn or somebody else's: every page carries its key and a version in every
 *     64-bit word;
 *   - every call returns (the combining queue loses no request): the run ends, a watchdog aborts it
 *     otherwise;
 *   - requests / hits counters add up.
 * usage: host_stress <cachedir> <threads> <ops per thread> <pshift> [watchdog seconds] [evict] */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../../include/cachemap.h"
#include "../../include/filemap.h"

static struct cachemap *cm;
static int threads, ops, pshift;
static size_t bsize;
static long errors, gets_done, hits_done, range_pages;

static void fill(uint8_t *p, uint64_t key, uint64_t ver) {
	uint64_t *w = (uint64_t *)p;
	for (size_t i = 0; i < bsize / 8; i++) w[i] = key * 0x9E3779B97F4A7C15ull + (ver << 20) + i;
}
/* version of a well-formed page of `key`, or -1 */
static long check(const uint8_t *p, uint64_t key) {
	const uint64_t *w = (const uint64_t *)p;
	const uint64_t ver = (w[0] - key * 0x9E3779B97F4A7C15ull) >> 20;
	for (size_t i = 0; i < bsize / 8; i++)
		if (w[i] != key * 0x9E3779B97F4A7C15ull + (ver << 20) + i) return -1;
	return (long)ver;
}

static void *worker(void *arg) {
	const long t = (long)arg;
	unsigned seed = 12345u + (unsigned)t * 7919u;
	uint8_t *page = malloc(bsize), *buf = malloc(bsize * 8);
	enum { OWN = 48, SHARED = 16 };
	uint64_t own_ver[OWN] = { 0 };
	long err = 0, g = 0, h = 0, rp = 0;
	for (int op = 0; op < ops; op++) {
		const unsigned r = rand_r(&seed);
		if (r % 10 < 4) {                                        /* put on a key of my own */
			const int k = (r >> 8) % OWN;
			const uint64_t key = 1000u * (uint64_t)(t + 1) + (uint64_t)k;
			fill(page, key, ++own_ver[k]);
			if (r & 0x10000) cachemap_put_async(cm, key << pshift, 7, 1, page);
			else cachemap_put(cm, key << pshift, 7, 1, page);
		} else if (r % 10 < 7) {                                 /* get of my own key: must be my latest version */
			const int k = (r >> 8) % OWN;
			const uint64_t key = 1000u * (uint64_t)(t + 1) + (uint64_t)k;
			uint8_t *p = cachemap_get(cm, key << pshift, 7, 1);
			g++;
			if (own_ver[k] == 0) { if (p) err++; }
			else if (!p || check(p, key) != (long)own_ver[k]) err++;
			if (p) h++;
			free(p);
		} else if (r % 10 < 8) {                                 /* shared key: any well-formed version of it, or a miss */
			const uint64_t key = 500000u + (r >> 8) % SHARED;
			if (r & 0x10000) { fill(page, key, (uint64_t)op + 1); cachemap_put(cm, key << pshift, 7, 1, page); }
			else {
				uint8_t *p = cachemap_get(cm, key << pshift, 7, 1);
				g++;
				if (p) { h++; if (check(p, key) < 0) err++; }
				free(p);
			}
		} else if (r % 10 < 9) {                                 /* a key nobody writes: a miss */
			uint8_t *p = cachemap_get(cm, (900000u + (uint64_t)t * 100 + (r >> 8) % 50) << pshift, 7, 1);
			g++;
			if (p) { err++; h++; }
			free(p);
		} else {                                                 /* a range over my keys: what is there must be mine and whole */
			const int k0 = (r >> 8) % (OWN - 8);
			const uint64_t key0 = 1000u * (uint64_t)(t + 1) + (uint64_t)k0;
			const int got = cachemap_read_range(cm, 7, 1, key0 << pshift, bsize * 8, buf);
			int all = 1;
			for (int j = 0; j < 8; j++) all &= own_ver[k0 + j] != 0;
			if (got != all) err++;                                 /* every page of mine that was ever put is there */
			if (got)
				for (int j = 0; j < 8; j++)
					if (check(buf + (size_t)j * bsize, key0 + (uint64_t)j) != (long)own_ver[k0 + j]) err++;
			rp += 8;
		}
	}
	__atomic_fetch_add(&errors, err, __ATOMIC_RELAXED);
	__atomic_fetch_add(&gets_done, g, __ATOMIC_RELAXED);
	__atomic_fetch_add(&hits_done, h, __ATOMIC_RELAXED);
	__atomic_fetch_add(&range_pages, rp, __ATOMIC_RELAXED);
	free(page); free(buf);
	return NULL;
}

/* Eviction mode: far more keys than the capacity.  Nothing can be promised about which keys stay,
 * only that a page that comes back is whole and its key's, and that the store ends up at capacity. */
static void *evict_worker(void *arg) {
	const long t = (long)arg;
	unsigned seed = 777u + (unsigned)t * 104729u;
	uint8_t *page = malloc(bsize);
	long err = 0, g = 0, h = 0;
	for (int op = 0; op < ops; op++) {
		const unsigned r = rand_r(&seed);
		const uint64_t key = 2000000u + (uint64_t)t * 100000u + (r >> 4) % 20000u;
		if (r & 1) { fill(page, key, (uint64_t)op + 1); cachemap_put(cm, key << pshift, 9, 2, page); }
		else {
			uint8_t *p = cachemap_get(cm, key << pshift, 9, 2);
			g++;
			if (p) { h++; if (check(p, key) < 0) err++; }
			free(p);
		}
	}
	__atomic_fetch_add(&errors, err, __ATOMIC_RELAXED);
	__atomic_fetch_add(&gets_done, g, __ATOMIC_RELAXED);
	__atomic_fetch_add(&hits_done, h, __ATOMIC_RELAXED);
	free(page);
	return NULL;
}

static void *watchdog(void *arg) {
	sleep((unsigned)(uintptr_t)arg);
	fprintf(stderr, "host_stress: watchdog — calls did not return in time\n");
	_exit(9);
	return NULL;
}

int main(int argc, char **argv) {
	if (argc < 5) { fprintf(stderr, "usage: %s cachedir threads ops pshift\n", argv[0]); return 2; }
	threads = atoi(argv[2]); ops = atoi(argv[3]); pshift = atoi(argv[4]);
	bsize = (size_t)1 << pshift;
	pthread_t wd;
	pthread_create(&wd, NULL, watchdog, (void *)(uintptr_t)(argc > 5 ? atoi(argv[5]) : 120));
	const int evict = argc > 6 && strcmp(argv[6], "evict") == 0;
	const uint64_t capacity = evict ? 2048 : 1 << 15;
	cm = cachemap_create(argv[1], capacity, 12, pshift);
	if (!cm) { fprintf(stderr, "cachemap_create failed\n"); return 1; }
	pthread_t th[256];
	if (threads > 256) threads = 256;
	for (long t = 0; t < threads; t++) pthread_create(&th[t], NULL, evict ? evict_worker : worker, (void *)t);
	for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
	uint64_t rq = 0, ht = 0;
	cachemap_get_counters(cm, &rq, &ht);
	printf("threads %d ops %d errors %ld gets %ld hits %ld counters %lu/%lu range_pages %ld\n", threads, ops, errors, gets_done, hits_done,
	    (unsigned long)rq, (unsigned long)ht, range_pages);
	int bad = errors != 0;
	/* single gets are counted one by one; the range reads add requests of their own (at most one per page) */
	if ((long)rq < gets_done || (long)rq > gets_done + range_pages || (long)ht < hits_done) { fprintf(stderr, "counters do not add up\n"); bad = 1; }
	if (evict) {
		/* eviction runs before every batch of the flusher: what is left over is at most the capacity
		 * (cachemap.c:17-45), and the cache did fill up */
		const uint64_t left = filemap_entries(*(struct filemap **)cm);     /* `pages` is the first member (cachemap.h:20-21) */
		printf("entries %lu capacity %lu\n", (unsigned long)left, (unsigned long)capacity);
		if (left > capacity || left < capacity / 2) { fprintf(stderr, "entries out of bounds\n"); bad = 1; }
	}
	cachemap_free(cm);
	printf(bad ? "host_stress FAILED\n" : "host_stress ok\n");
	return bad;
}
