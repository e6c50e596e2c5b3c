/*
 * cachemap_api.c — the reference's C API (include/cachemap.h, include/filemap.h) over the B200
 * engine.  Host code stays C; everything heavy happens in the engine's kernels.
 *
 * What each reference function became:
 *   filemap_create/free        cachemap/filemap.c:35-110   -> config only; engine built lazily
 *   filemap_set/get/unset      cachemap/filemap.c:112-262  -> one request in a combining queue;
 *                              whichever caller finds no batch in flight becomes the leader, takes
 *                              every queued request (each from a different thread, all
 *                              outstanding at once, so any order is a valid linearisation) and
 *                              runs them as one GPU batch: unsets, then sets, then gets
 *   filemap_get_rand/entries   cachemap/filemap.c:264-330  -> table sample kernel / device counter
 *   cachemap_*                 cachemap/cachemap.c:107-239 -> same logic: address composition,
 *                              timestamps, evict-min-of-3 when entries >= capacity, counters,
 *                              async queue (one flusher thread that batches instead of 4 workers)
 * There is no CPU fallback: if the engine cannot be created the process stops with a message
 * (set CMB200_SOFT_FAIL=1 to degrade to "every put dropped, every get a miss" instead).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

#include "../../include/cachemap.h"
#include "../../include/cachemap_b200.h"

#define COMBINE_MAX 256         /* requests one leader takes per GPU batch */
#define PNUM_SHIFT 44           /* cachemap.c:155 */

enum req_kind { REQ_SET, REQ_GET, REQ_UNSET };

struct fm_req {
	enum req_kind kind;
	cmb200_addr addr;
	const void *page;       /* REQ_SET */
	uint64_t ts;
	void *out;              /* REQ_GET: malloc()ed page or NULL */
	int bad_entry;
	int done;
	struct fm_req *next;
};

struct filemap {
	uint64_t n;
	int compress;
	int bsize;
	int pshift;
	char destdir[2048];
	/* engine, built on first use (fork safety) */
	pthread_mutex_t init_mu;
	int init_state;         /* 0 = not yet, 1 = ready, -1 = failed */
	cmb200_engine *eng;
	uint8_t *h_stage;       /* page-locked, COMBINE_MAX pages */
	/* combining queue */
	pthread_mutex_t q_mu;
	pthread_cond_t q_cv;
	struct fm_req *q_head, *q_tail;
	int leader_active;
};

static long
env_long(const char *name, long dflt)
{
	const char *v = getenv(name);
	return (v && *v) ? strtol(v, NULL, 0) : dflt;
}

static int
filemap_engine_ready(struct filemap *m)
{
	if (__atomic_load_n(&m->init_state, __ATOMIC_ACQUIRE) == 1)
		return 1;
	pthread_mutex_lock(&m->init_mu);
	if (m->init_state == 0) {
		cmb200_config cfg;
		memset(&cfg, 0, sizeof(cfg));
		cfg.device = (int)env_long("CMB200_DEVICE", -1);
		cfg.pshift = m->pshift;
		cfg.accel = m->compress;
		cfg.capacity = m->n;
		cfg.arena_bytes = (uint64_t)env_long("CMB200_ARENA_MB", 0) << 20;
		cfg.table_slots = (uint64_t)env_long("CMB200_TABLE_SLOTS", 0);
		cfg.max_batch = (uint32_t)env_long("CMB200_MAX_BATCH", 0);
		cfg.flags = env_long("CMB200_FINGERPRINT", 0) ? CMB200_FINGERPRINT : 0;
		m->eng = cmb200_engine_create(&cfg);
		if (m->eng)
			m->h_stage = cmb200_host_alloc((size_t)COMBINE_MAX * m->bsize);
		if (!m->eng || !m->h_stage) {
			fprintf(stderr, "cachemap_b200: cannot start the GPU engine: %s\n", cmb200_last_error());
			if (!env_long("CMB200_SOFT_FAIL", 0)) {
				fprintf(stderr, "cachemap_b200: no CPU fallback exists; aborting "
				    "(CMB200_SOFT_FAIL=1 turns this into dropped puts / misses)\n");
				abort();
			}
			__atomic_store_n(&m->init_state, -1, __ATOMIC_RELEASE);
		} else {
			__atomic_store_n(&m->init_state, 1, __ATOMIC_RELEASE);
		}
	}
	pthread_mutex_unlock(&m->init_mu);
	return m->init_state == 1;
}

struct filemap *
filemap_create(char *destdir, uint64_t n, int compress_accel, int pshift)
{
	if (!destdir || strlen(destdir) >= sizeof(((struct filemap *)0)->destdir))
		return NULL;
	if (n < FILEMAP_SHARD_FACTOR)           /* filemap.c:51 */
		return NULL;
	if (pshift < 6 || pshift > 20)
		return NULL;
	struct filemap *m = calloc(1, sizeof(*m));
	if (!m)
		return NULL;
	m->n = n;
	m->compress = compress_accel;
	m->bsize = 1 << pshift;
	m->pshift = pshift;
	strcpy(m->destdir, destdir);
	pthread_mutex_init(&m->init_mu, NULL);
	pthread_mutex_init(&m->q_mu, NULL);
	pthread_cond_init(&m->q_cv, NULL);
	return m;
}

void
filemap_free(struct filemap *m)
{
	if (!m)
		return;
	if (m->h_stage)
		cmb200_host_free(m->h_stage);
	if (m->eng)
		cmb200_engine_destroy(m->eng);
	pthread_mutex_destroy(&m->init_mu);
	pthread_mutex_destroy(&m->q_mu);
	pthread_cond_destroy(&m->q_cv);
	free(m);
}

/* Runs one combined batch.  Called by the leader without q_mu held. */
static void
filemap_run_batch(struct filemap *m, struct fm_req **reqs, int count)
{
	cmb200_addr addr[COMBINE_MAX];
	uint64_t ts[COMBINE_MAX];
	int32_t status[COMBINE_MAX];
	int idx[COMBINE_MAX];
	int k;

	k = 0;
	for (int i = 0; i < count; i++)
		if (reqs[i]->kind == REQ_UNSET)
			addr[k++] = reqs[i]->addr;
	if (k)
		cmb200_unset_batch(m->eng, (size_t)k, addr);

	k = 0;
	for (int i = 0; i < count; i++) {
		if (reqs[i]->kind != REQ_SET)
			continue;
		addr[k] = reqs[i]->addr;
		ts[k] = reqs[i]->ts;
		memcpy(m->h_stage + (size_t)k * m->bsize, reqs[i]->page, (size_t)m->bsize);
		k++;
	}
	if (k)
		cmb200_put_batch(m->eng, (size_t)k, addr, NULL, m->h_stage, ts, NULL);

	k = 0;
	for (int i = 0; i < count; i++) {
		if (reqs[i]->kind != REQ_GET)
			continue;
		addr[k] = reqs[i]->addr;
		idx[k] = i;
		k++;
	}
	if (k && cmb200_get_batch(m->eng, (size_t)k, addr, NULL, m->h_stage, status) == 0) {
		for (int j = 0; j < k; j++) {
			struct fm_req *r = reqs[idx[j]];
			if (status[j] == CMB200_HIT) {
				r->out = malloc((size_t)m->bsize);      /* filemap.c:242 */
				if (r->out)
					memcpy(r->out, m->h_stage + (size_t)j * m->bsize, (size_t)m->bsize);
			} else if (status[j] == CMB200_BAD_ENTRY) {
				r->bad_entry = 1;
			}
		}
	}
}

static void
filemap_submit(struct filemap *m, struct fm_req *req)
{
	req->done = 0;
	req->next = NULL;
	pthread_mutex_lock(&m->q_mu);
	if (m->q_tail)
		m->q_tail->next = req;
	else
		m->q_head = req;
	m->q_tail = req;
	while (!req->done) {
		if (m->leader_active) {
			pthread_cond_wait(&m->q_cv, &m->q_mu);
			continue;
		}
		struct fm_req *batch[COMBINE_MAX];
		int count = 0;
		m->leader_active = 1;
		while (m->q_head && count < COMBINE_MAX) {
			batch[count++] = m->q_head;
			m->q_head = m->q_head->next;
		}
		if (!m->q_head)
			m->q_tail = NULL;
		pthread_mutex_unlock(&m->q_mu);
		filemap_run_batch(m, batch, count);
		pthread_mutex_lock(&m->q_mu);
		for (int i = 0; i < count; i++)
			batch[i]->done = 1;
		m->leader_active = 0;
		pthread_cond_broadcast(&m->q_cv);
	}
	pthread_mutex_unlock(&m->q_mu);
}

void
filemap_set(struct filemap *m, uint128_t *key, void *value, uint64_t attr)
{
	if (!filemap_engine_ready(m))
		return;
	struct fm_req r;
	memset(&r, 0, sizeof(r));
	r.kind = REQ_SET;
	r.addr.u = key->u;
	r.addr.l = key->l;
	r.page = value;
	r.ts = attr;
	filemap_submit(m, &r);
}

void
filemap_unset(struct filemap *m, uint128_t *key)
{
	if (!filemap_engine_ready(m))
		return;
	struct fm_req r;
	memset(&r, 0, sizeof(r));
	r.kind = REQ_UNSET;
	r.addr.u = key->u;
	r.addr.l = key->l;
	filemap_submit(m, &r);
}

void *
filemap_get(struct filemap *m, uint128_t *key)
{
	if (!filemap_engine_ready(m))
		return NULL;
	struct fm_req r;
	memset(&r, 0, sizeof(r));
	r.kind = REQ_GET;
	r.addr.u = key->u;
	r.addr.l = key->l;
	filemap_submit(m, &r);
	if (r.bad_entry)
		printf("bad entry\n");          /* filemap.c:237 */
	return r.out;
}

int
filemap_get_rand(struct filemap *m, uint128_t *key, uint64_t *ts)
{
	if (!filemap_engine_ready(m))
		return 0;
	/* filemap.c:271-274: a 64-bit draw built from rand() */
	uint64_t r = 0;
	for (int i = 0; i < 64; i += 30)
		r = r * ((uint64_t)RAND_MAX + 1) + (uint64_t)rand();
	cmb200_addr a;
	int32_t ok = 0;
	if (cmb200_sample(m->eng, 1, &r, &a, ts, &ok) != 0 || !ok)
		return 0;
	key->u = a.u;
	key->l = a.l;
	return 1;
}

uint64_t
filemap_entries(struct filemap *m)
{
	if (!filemap_engine_ready(m))
		return 0;
	return cmb200_entries(m->eng);
}

/* ------------------------------------------------------------------------------------------ */

struct async_node {
	struct async_node *next;
	cmb200_addr addr;
	uint64_t ts;
	void *page;
};

struct cachemap {
	struct filemap *pages;
	uint64_t capacity;
	uint64_t requests;
	uint64_t hits;
	/* async put queue (cachemap.c:50-105,199-216) */
	pthread_mutex_t a_mu;
	pthread_cond_t a_cv;
	struct async_node *a_head, *a_tail;
	pthread_t flusher;
	int flusher_started;
	int stop;
};

static uint64_t
now_ns(void)
{
	struct timespec tp;
	(void)clock_gettime(CLOCK_REALTIME_COARSE, &tp);        /* cachemap.c:10-15 */
	return (uint64_t)tp.tv_sec * 1000000000ULL + (uint64_t)tp.tv_nsec;
}

/* cachemap.c:151-166 */
static int
compose_addr(struct cachemap *cm, uint64_t offset, uint64_t nhid_small, uint32_t genid, cmb200_addr *out)
{
	uint64_t page = offset >> cm->pages->pshift;
	if (page >> PNUM_SHIFT)
		return -1;
	out->l = page | ((uint64_t)genid << PNUM_SHIFT);
	out->u = nhid_small;
	return 0;
}

struct cachemap *
cachemap_create(char *destdir, uint64_t capacity, int comp_accel, int pshift)
{
	struct stat sb;
	if (!destdir || stat(destdir, &sb) != 0 || !S_ISDIR(sb.st_mode))   /* cachemap.c:113-114 */
		return NULL;
	struct cachemap *cm = calloc(1, sizeof(*cm));
	if (!cm)
		return NULL;
	cm->pages = filemap_create(destdir, capacity, comp_accel, pshift);
	if (!cm->pages) {
		free(cm);
		return NULL;
	}
	/* mutex and condvar exist before any thread that uses them (the reference starts its
	 * workers first, cachemap.c:123-139, and can hang in cachemap_free because of it) */
	pthread_mutex_init(&cm->a_mu, NULL);
	pthread_cond_init(&cm->a_cv, NULL);
	cm->capacity = capacity;
	return cm;
}

/* Makes room for `incoming` puts: while entries + incoming > capacity, retire the oldest of three
 * random live records (cachemap.c:17-45).  Statistically the reference's policy; bitwise parity
 * is undefined there (wall-clock timestamps, rand()). */
static void
cachemap_make_room(struct cachemap *cm, uint64_t incoming)
{
	struct filemap *m = cm->pages;
	for (int round = 0; round < 8; round++) {
		uint64_t entries = filemap_entries(m);
		if (entries + incoming <= cm->capacity || entries == 0)
			return;
		uint64_t need = entries + incoming - cm->capacity;
		if (need > entries)
			need = entries;
		if (need > 1024)
			need = 1024;    /* per round; the loop continues */
		uint64_t *draws = malloc(3 * need * sizeof(uint64_t));
		uint64_t *ts = malloc(3 * need * sizeof(uint64_t));
		int32_t *ok = malloc(3 * need * sizeof(int32_t));
		cmb200_addr *cand = malloc(3 * need * sizeof(cmb200_addr));
		cmb200_addr *victim = malloc(need * sizeof(cmb200_addr));
		uint64_t nv = 0;
		if (draws && ts && ok && cand && victim) {
			for (uint64_t i = 0; i < 3 * need; i++) {
				uint64_t r = 0;
				for (int b = 0; b < 64; b += 30)
					r = r * ((uint64_t)RAND_MAX + 1) + (uint64_t)rand();
				draws[i] = r;
			}
			if (cmb200_sample(m->eng, (size_t)(3 * need), draws, cand, ts, ok) == 0) {
				for (uint64_t i = 0; i < need; i++) {
					uint64_t a = ts[3 * i], b = ts[3 * i + 1], c = ts[3 * i + 2];
					int pick;
					if (a < b)
						pick = (a > c) ? 2 : 0;         /* cachemap.c:29-41 */
					else
						pick = (b > c) ? 2 : 1;
					if (ok[3 * i + pick])
						victim[nv++] = cand[3 * i + pick];
				}
				if (nv)
					cmb200_unset_batch(m->eng, (size_t)nv, victim);
			}
		}
		free(draws); free(ts); free(ok); free(cand); free(victim);
		if (nv == 0)
			return;
		if (incoming == 1)
			return;         /* the reference evicts exactly one per put */
	}
}

void *
cachemap_get(struct cachemap *cm, uint64_t offset, uint64_t nhid_small, uint32_t genid)
{
	cmb200_addr a;
	if (compose_addr(cm, offset, nhid_small, genid, &a) != 0)
		return NULL;
	__atomic_fetch_add(&cm->requests, 1, __ATOMIC_RELAXED);         /* cachemap.c:176 */
	uint128_t key = { a.u, a.l };
	void *page = filemap_get(cm->pages, &key);
	if (page)
		__atomic_fetch_add(&cm->hits, 1, __ATOMIC_RELAXED);     /* cachemap.c:181 */
	return page;
}

void
cachemap_put(struct cachemap *cm, uint64_t offset, uint64_t nhid_small, uint32_t genid, const void *page)
{
	cmb200_addr a;
	if (compose_addr(cm, offset, nhid_small, genid, &a) != 0)
		return;
	uint64_t ts = now_ns();
	if (!filemap_engine_ready(cm->pages))
		return;
	cachemap_make_room(cm, 1);
	uint128_t key = { a.u, a.l };
	filemap_set(cm->pages, &key, (void *)page, ts);
}

static void
cachemap_flush_async(struct cachemap *cm, struct async_node *list)
{
	struct filemap *m = cm->pages;
	cmb200_addr addr[COMBINE_MAX];
	uint64_t ts[COMBINE_MAX];
	/* private page-locked gather buffer: the combiner's belongs to its leader */
	uint8_t *buf = cmb200_host_alloc((size_t)COMBINE_MAX * m->bsize);

	while (list) {
		struct async_node *first = list;
		int k = 0;
		while (list && k < COMBINE_MAX) {
			addr[k] = list->addr;
			ts[k] = list->ts;
			if (buf)
				memcpy(buf + (size_t)k * m->bsize, list->page, (size_t)m->bsize);
			list = list->next;
			k++;
		}
		if (buf) {
			cachemap_make_room(cm, (uint64_t)k);
			cmb200_put_batch(m->eng, (size_t)k, addr, NULL, buf, ts, NULL);
		}
		while (first != list) {
			struct async_node *d = first;
			first = first->next;
			free(d->page);
			free(d);
		}
	}
	cmb200_host_free(buf);
}

static void *
cachemap_flusher(void *arg)
{
	struct cachemap *cm = arg;
	pthread_mutex_lock(&cm->a_mu);
	while (cm->a_head || !cm->stop) {
		if (!cm->a_head) {
			pthread_cond_wait(&cm->a_cv, &cm->a_mu);
			continue;
		}
		struct async_node *list = cm->a_head;
		cm->a_head = cm->a_tail = NULL;
		pthread_mutex_unlock(&cm->a_mu);
		if (filemap_engine_ready(cm->pages)) {
			cachemap_flush_async(cm, list);
		} else {
			while (list) {
				struct async_node *d = list;
				list = list->next;
				free(d->page);
				free(d);
			}
		}
		pthread_mutex_lock(&cm->a_mu);
	}
	pthread_mutex_unlock(&cm->a_mu);
	return NULL;
}

void
cachemap_put_async(struct cachemap *cm, uint64_t offset, uint64_t nhid_small, uint32_t genid, const void *page)
{
	cmb200_addr a;
	if (compose_addr(cm, offset, nhid_small, genid, &a) != 0)
		return;
	struct async_node *n = malloc(sizeof(*n));
	if (!n)
		return;
	n->page = malloc((size_t)cm->pages->bsize);             /* cachemap.c:207-208 */
	if (!n->page) {
		free(n);
		return;
	}
	memcpy(n->page, page, (size_t)cm->pages->bsize);
	n->addr = a;
	n->ts = now_ns();
	n->next = NULL;
	pthread_mutex_lock(&cm->a_mu);
	if (!cm->flusher_started) {
		/* started on first use, i.e. in the process that actually caches (after any fork) */
		if (pthread_create(&cm->flusher, NULL, cachemap_flusher, cm) != 0) {
			pthread_mutex_unlock(&cm->a_mu);
			free(n->page);
			free(n);
			return;
		}
		cm->flusher_started = 1;
	}
	if (cm->a_tail)
		cm->a_tail->next = n;
	else
		cm->a_head = n;
	cm->a_tail = n;
	pthread_cond_signal(&cm->a_cv);
	pthread_mutex_unlock(&cm->a_mu);
}

void
cachemap_free(struct cachemap *cm)
{
	if (!cm)
		return;
	pthread_mutex_lock(&cm->a_mu);
	cm->stop = 1;
	pthread_cond_broadcast(&cm->a_cv);
	pthread_mutex_unlock(&cm->a_mu);
	if (cm->flusher_started)
		pthread_join(cm->flusher, NULL);
	pthread_mutex_destroy(&cm->a_mu);
	pthread_cond_destroy(&cm->a_cv);
	filemap_free(cm->pages);
	free(cm);
}

void
cachemap_print_stats(struct cachemap *cm)
{
	uint64_t rq = __atomic_load_n(&cm->requests, __ATOMIC_RELAXED);
	uint64_t ht = __atomic_load_n(&cm->hits, __ATOMIC_RELAXED);
	printf("requests: %lu, hits: %lu, ratio: %5.2f\n",              /* cachemap.c:237-238 */
	    (unsigned long)rq, (unsigned long)ht, ht * 100 / (float)rq);
}

/* ---- batch extension ---------------------------------------------------------------------- */

struct batch_keys {
	cmb200_addr *addr;
	uint8_t *valid;
	uint64_t *ts;
};

static int
batch_keys_build(struct cachemap *cm, uint64_t n, const uint64_t *offset, const uint64_t *nhid,
    const uint32_t *genid, int want_ts, struct batch_keys *bk)
{
	bk->addr = malloc((size_t)n * sizeof(cmb200_addr));
	bk->valid = malloc((size_t)n);
	bk->ts = want_ts ? malloc((size_t)n * 8) : NULL;
	if (!bk->addr || !bk->valid || (want_ts && !bk->ts)) {
		free(bk->addr); free(bk->valid); free(bk->ts);
		return -1;
	}
	uint64_t ts = want_ts ? now_ns() : 0;
	for (uint64_t i = 0; i < n; i++) {
		bk->valid[i] = compose_addr(cm, offset[i], nhid[i], genid ? genid[i] : 0, &bk->addr[i]) == 0;
		if (!bk->valid[i])
			memset(&bk->addr[i], 0, sizeof(cmb200_addr));
		if (want_ts)
			bk->ts[i] = ts;
	}
	return 0;
}

static void
batch_keys_free(struct batch_keys *bk)
{
	free(bk->addr); free(bk->valid); free(bk->ts);
}

static void
put_batch_common(struct cachemap *cm, uint64_t n, const uint64_t *offset, const uint64_t *nhid,
    const uint32_t *genid, const void *pages, int on_dev)
{
	struct batch_keys bk;
	if (n == 0 || !filemap_engine_ready(cm->pages))
		return;
	if (batch_keys_build(cm, n, offset, nhid, genid, 1, &bk) != 0)
		return;
	cachemap_make_room(cm, n);
	if (on_dev)
		cmb200_put_batch_dev(cm->pages->eng, (size_t)n, bk.addr, bk.valid, pages, bk.ts, NULL);
	else
		cmb200_put_batch(cm->pages->eng, (size_t)n, bk.addr, bk.valid, pages, bk.ts, NULL);
	batch_keys_free(&bk);
}

static void
get_batch_common(struct cachemap *cm, uint64_t n, const uint64_t *offset, const uint64_t *nhid,
    const uint32_t *genid, void *pages_out, uint8_t *hit_out, int on_dev)
{
	struct batch_keys bk;
	memset(hit_out, 0, (size_t)n);
	if (n == 0 || !filemap_engine_ready(cm->pages))
		return;
	if (batch_keys_build(cm, n, offset, nhid, genid, 0, &bk) != 0)
		return;
	int32_t *status = malloc((size_t)n * 4);
	int rc = -1;
	if (status)
		rc = on_dev ? cmb200_get_batch_dev(cm->pages->eng, (size_t)n, bk.addr, bk.valid, pages_out, status)
			    : cmb200_get_batch(cm->pages->eng, (size_t)n, bk.addr, bk.valid, pages_out, status);
	uint64_t rq = 0, ht = 0;
	for (uint64_t i = 0; i < n; i++) {
		if (!bk.valid[i])
			continue;                       /* cachemap.c:173-174: not a request */
		rq++;
		if (rc == 0 && status[i] == CMB200_HIT) {
			hit_out[i] = 1;
			ht++;
		} else if (rc == 0 && status[i] == CMB200_BAD_ENTRY) {
			printf("bad entry\n");
		}
	}
	__atomic_fetch_add(&cm->requests, rq, __ATOMIC_RELAXED);
	__atomic_fetch_add(&cm->hits, ht, __ATOMIC_RELAXED);
	free(status);
	batch_keys_free(&bk);
}

void
cachemap_put_batch(struct cachemap *cm, uint64_t n, const uint64_t *offset, const uint64_t *nhid_small,
    const uint32_t *genid, const void *pages)
{
	put_batch_common(cm, n, offset, nhid_small, genid, pages, 0);
}

void
cachemap_put_batch_dev(struct cachemap *cm, uint64_t n, const uint64_t *offset, const uint64_t *nhid_small,
    const uint32_t *genid, const void *pages_dev)
{
	put_batch_common(cm, n, offset, nhid_small, genid, pages_dev, 1);
}

void
cachemap_get_batch(struct cachemap *cm, uint64_t n, const uint64_t *offset, const uint64_t *nhid_small,
    const uint32_t *genid, void *pages_out, uint8_t *hit_out)
{
	get_batch_common(cm, n, offset, nhid_small, genid, pages_out, hit_out, 0);
}

void
cachemap_get_batch_dev(struct cachemap *cm, uint64_t n, const uint64_t *offset, const uint64_t *nhid_small,
    const uint32_t *genid, void *pages_out_dev, uint8_t *hit_out)
{
	get_batch_common(cm, n, offset, nhid_small, genid, pages_out_dev, hit_out, 1);
}

void
cachemap_get_counters(struct cachemap *cm, uint64_t *requests, uint64_t *hits)
{
	*requests = __atomic_load_n(&cm->requests, __ATOMIC_RELAXED);
	*hits = __atomic_load_n(&cm->hits, __ATOMIC_RELAXED);
}

struct cmb200_engine *
cachemap_engine(struct cachemap *cm)
{
	return filemap_engine_ready(cm->pages) ? cm->pages->eng : NULL;
}
