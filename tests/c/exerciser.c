/*
 * exerciser.c — the workload shape of the reference's only exerciser (cachemap/cachemap_test.c:
 * capacity == object count, asynchronous inserts, two read sweeps, half of the objects put again
 * under new generation ids — which at capacity evicts one record per put, cachemap.c:17-48,186-197
 * — and two more read sweeps), written against the public API only, with per-phase hit counts
 * printed in a parseable form.  The same binary source is linked once against this repository's
 * libcachemap.so.0.0 and once against the compiled reference (oracle/_ref/libcachemap_ref.so):
 * eviction is random and wall-clock driven in both, so what must agree is the hit ratio of every
 * phase, not individual victims.
 *
 * usage: exerciser <dir> <objects> <pshift> <seed>
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "cachemap.h"
#include "filemap.h"

static uint64_t mix(uint64_t z) {
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

struct obj { uint64_t offset, nhid; uint32_t genid; };

static int sweep(struct cachemap *cm, struct obj *o, int n, int first, int count) {
	int hits = 0;
	for (int i = first; i < first + count; i++) {
		void *p = cachemap_get(cm, o[i].offset, o[i].nhid, o[i].genid);
		if (p) {
			if (((unsigned char *)p)[0] != (unsigned char)i || ((unsigned char *)p)[1] != (unsigned char)(i >> 8)) {
				printf("WRONG PAGE for object %d\n", i);
				_exit(3);
			}
			hits++;
			free(p);
		}
	}
	(void)n;
	return hits;
}

int main(int argc, char **argv) {
	if (argc != 5) { fprintf(stderr, "usage: %s dir objects pshift seed\n", argv[0]); return 2; }
	const int n = atoi(argv[2]), pshift = atoi(argv[3]);
	const uint64_t seed = strtoull(argv[4], NULL, 0);
	const size_t ps = (size_t)1 << pshift;
	struct cachemap *cm = cachemap_create(argv[1], (uint64_t)n, 12, pshift);
	if (!cm) { printf("cachemap_create failed\n"); return 1; }
	struct obj *o = calloc((size_t)n, sizeof(*o));
	unsigned char *page = calloc(1, ps);
	for (int i = 0; i < n; i++) {
		o[i].offset = (uint64_t)i * 4096u;            /* several objects per page number, told apart by nhid / genid */
		o[i].genid = (uint32_t)i;
		o[i].nhid = (uint64_t)i * (mix(seed + (uint64_t)i) >> 33);
	}
	for (int i = 0; i < n; i++) {                         /* asynchronous inserts; the page buffer is reused at once */
		page[0] = (unsigned char)i; page[1] = (unsigned char)(i >> 8);
		cachemap_put_async(cm, o[i].offset, o[i].nhid, o[i].genid, page);
	}
	/* the reference sleeps a second here; wait until the asynchronous puts have actually landed */
	struct filemap *fm = *(struct filemap **)cm;          /* cachemap.h: `pages` is the first member */
	for (int t = 0; t < 600 && filemap_entries(fm) < (uint64_t)n; t++) usleep(50000);
	printf("entries_after_insert %lu\n", (unsigned long)filemap_entries(fm));
	printf("phase read1 hits %d of %d\n", sweep(cm, o, n, 0, n), n);
	printf("phase read2 hits %d of %d\n", sweep(cm, o, n, 0, n), n);
	cachemap_print_stats(cm);
	for (int i = 0; i < n / 2; i++) {                     /* new generation ids = new keys: every put evicts */
		o[i].genid = (uint32_t)(mix(seed ^ 0x5151 ^ (uint64_t)i) & 0xfffff) | 0x80000u;
		page[0] = (unsigned char)i; page[1] = (unsigned char)(i >> 8);
		cachemap_put_async(cm, o[i].offset, o[i].nhid, o[i].genid, page);
	}
	uint64_t last = 0;
	for (int t = 0, same = 0; t < 600 && same < 6; t++) {  /* until the entry count has settled */
		usleep(50000);
		uint64_t e = filemap_entries(fm);
		same = (e == last) ? same + 1 : 0;
		last = e;
	}
	printf("entries_after_reput %lu\n", (unsigned long)filemap_entries(fm));
	printf("phase reput_new hits %d of %d\n", sweep(cm, o, n, 0, n / 2), n / 2);
	printf("phase reput_old hits %d of %d\n", sweep(cm, o, n, n / 2, n - n / 2), n - n / 2);
	printf("phase read4 hits %d of %d\n", sweep(cm, o, n, 0, n), n);
	cachemap_print_stats(cm);
	fflush(stdout);
	_exit(0);                                             /* no cachemap_free: it can hang in the reference */
}
