/*
 * drop_in_test.c — a C caller of the reference API, linked against libcachemap.so.0.0 the way
 * edgefs is (-lcachemap).  Shape of the reference's only exerciser (cachemap/cachemap_test.c:
 * async inserts, read sweeps, re-put under new genids) but with assertions on the returned bytes.
 * usage: drop_in_test <dir> <pshift> <objects>
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "cachemap.h"
#include "edgefs_glue.h"

static void fill(unsigned char *p, size_t n, unsigned seed) {
	unsigned x = seed * 2654435761u + 1;
	for (size_t i = 0; i < n; i++) {
		x = x * 1664525u + 1013904223u;
		p[i] = (seed & 1) ? (unsigned char)('a' + ((x >> 24) & 3)) : (unsigned char)(x >> 24);
	}
	if (seed % 3 == 0) memset(p + n / 4, 0, n / 2);
}

int main(int argc, char **argv) {
	if (argc != 4) { fprintf(stderr, "usage: %s dir pshift objects\n", argv[0]); return 2; }
	int pshift = atoi(argv[2]), n = atoi(argv[3]);
	size_t ps = (size_t)1 << pshift;
	if (cachemap_create("/nonexistent-dir", 4096, 12, pshift)) return 3;       /* cachemap.c:113-114 */
	if (cachemap_create(argv[1], 1023, 12, pshift)) return 4;                  /* filemap.c:51 */
	struct cachemap *cm = cachemap_create(argv[1], 4096, 12, pshift);
	if (!cm) return 5;
	unsigned char *page = malloc(ps), *want = malloc(ps);
	uint64_t nhid;
	FNV_hash("object-name", 11, &nhid);
	for (int i = 0; i < n; i++) {                                              /* async inserts */
		fill(page, ps, (unsigned)i);
		cachemap_put_async(cm, (uint64_t)i << pshift, nhid, 0, page);           /* page reused at once */
	}
	int hits = 0;
	for (int tries = 0; tries < 200 && hits < n; tries++) {                    /* reference: sleep(1) */
		hits = 0;
		for (int i = 0; i < n; i++) {
			void *p = cachemap_get(cm, (uint64_t)i << pshift, nhid, 0);
			if (p) { fill(want, ps, (unsigned)i); if (memcmp(p, want, ps)) return 6; hits++; free(p); }
		}
		if (hits < n) usleep(20000);
	}
	if (hits != n) return 7;
	for (int i = 0; i < n / 2; i++) {                                          /* synchronous re-put, new genid */
		fill(page, ps, (unsigned)(i + 1000));
		cachemap_put(cm, (uint64_t)i << pshift, nhid, 9, page);
		void *p = cachemap_get(cm, (uint64_t)i << pshift, nhid, 9);
		if (!p || memcmp(p, page, ps)) return 8;
		free(p);
	}
	if (cachemap_get(cm, (uint64_t)(n + 5) << pshift, nhid, 0)) return 9;      /* miss */
	if (cachemap_get(cm, ((uint64_t)1 << 44) << pshift, nhid, 0)) return 10;   /* rejected address */
	/* the page loops of edgefs_read / edgefs_write as one call each (edgefs.c:1150-1228) */
	{
		uint64_t page_size, aligned, rq0, ht0, rq1, ht1;
		uint64_t nh2 = edgefs_build_nhid("other-object", edgefs_bucket_hid("/bk1"));
		size_t req = 4 * ps;
		unsigned char *buf = malloc(req), *back = malloc(req);
		for (size_t i = 0; i < 4; i++) fill(buf + i * ps, ps, (unsigned)(500 + i));
		if (!edgefs_cache_check(1, pshift, 8 * ps, req, &page_size, &aligned) || page_size != ps || aligned != 8 * ps) return 11;
		if (edgefs_cache_check(1, pshift, 8 * ps + 1, req, &page_size, &aligned) || aligned != 8 * ps) return 12;
		cachemap_get_counters(cm, &rq0, &ht0);
		if (cachemap_read_range(cm, nh2, 0, 8 * ps, req, back)) return 13;      /* cold: miss at the first page */
		cachemap_get_counters(cm, &rq1, &ht1);
		if (rq1 != rq0 + 1 || ht1 != ht0) return 14;                           /* the loop stops there */
		cachemap_write_range(cm, nh2, 0, 8 * ps, req, buf);
		cachemap_write_range(cm, nh2, 0, 20 * ps + 1, ps, buf);                 /* unaligned: not cached */
		if (!cachemap_read_range(cm, nh2, 0, 8 * ps, req, back) || memcmp(buf, back, req)) return 15;
		if (cachemap_read_range(cm, nh2, 0, 8 * ps + 1, ps, back)) return 16;   /* unaligned: bypasses the cache */
		if (cachemap_read_range(cm, nh2, 0, 10 * ps, 3 * ps, back)) return 17;  /* third page was never written */
		cachemap_get_counters(cm, &rq0, &ht0);
		if (rq0 != rq1 + 4 + 3 || ht0 != ht1 + 4 + 2) return 18;
		free(buf); free(back);
	}
	uint64_t rq, ht;
	cachemap_get_counters(cm, &rq, &ht);
	cachemap_print_stats(cm);
	if (cachemap_checkpoint(cm) != 0) return 19;
	cachemap_free(cm);
	/* the cache directory outlives the process image: a new cachemap on it serves the pages */
	cm = cachemap_create(argv[1], 4096, 12, pshift);
	if (!cm) return 20;
	for (int i = n / 2; i < n; i++) {
		void *p = cachemap_get(cm, (uint64_t)i << pshift, nhid, 0);
		fill(want, ps, (unsigned)i);
		if (!p || memcmp(p, want, ps)) return 21;
		free(p);
	}
	cachemap_free(cm);
	printf("drop_in_test ok: requests=%lu hits=%lu\n", (unsigned long)rq, (unsigned long)ht);
	free(page); free(want);
	return 0;
}
