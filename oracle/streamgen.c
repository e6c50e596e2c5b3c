/*
 * streamgen.c — CPU statement of the synthetic chunk streams of the benchmark and
 * the multi-threaded parity gate of bench.py.  TEST / BASELINE INFRASTRUCTURE ONLY (ef_oracle.h).
 *
 * The stream is a definition of this repository, not of the reference (which has no benchmark
 * input): chunk `cid` of a stream seeded `seed` has content class (cid + (cid >> 3)) & 3
 *   0 R  incompressible  8-byte words  word(w) = mix(base + (w + 1) * G)
 *   1 T  text-like       byte j = 16 bits r of word(j / 4), field j % 4:
 *                        r & 3 != 0 -> 'a' + ((r >> 2) & 3), else r >> 8
 *   2 Z  zero page, bytes 0..1 = cid & 0xffff little-endian
 *   3 M  first half as T, second half repeats the first half
 * with base = mix(seed ^ cid * 0xD1B54A32D192ED03), G = 0x9E3779B97F4A7C15, mix = splitmix64's
 * output function; address: object = cid >> 14, page = cid & 16383,
 * nhid_small = mix((seed ^ object) + G), offset = page << pshift, genid = 0.
 * Written from that text, independently of the device generator (edge_fuse_b200/csrc/streamgen.cuh),
 * so that `bench.py --impl reference` needs nothing of the product library and the two
 * generators check each other (tests/test_oracle_pin.py).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GOLD 0x9E3779B97F4A7C15ULL

static uint64_t
mix64(uint64_t z)
{
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

static void
text_fill(uint64_t base, uint8_t *out, uint32_t nbytes)
{
	for (uint32_t j = 0; j < nbytes; j += 4) {
		uint64_t w = mix64(base + ((uint64_t)(j >> 2) + 1) * GOLD);
		for (uint32_t f = 0; f < 4 && j + f < nbytes; f++) {
			uint32_t r = (uint32_t)(w >> (16 * f)) & 0xFFFFu;
			out[j + f] = (r & 3u) ? (uint8_t)('a' + ((r >> 2) & 3u)) : (uint8_t)(r >> 8);
		}
	}
}

void
ef_gen_chunk(uint64_t seed, uint64_t cid, uint32_t bsize, uint8_t *out)
{
	const uint64_t base = mix64(seed ^ (cid * 0xD1B54A32D192ED03ULL));
	switch ((cid + (cid >> 3)) & 3) {
	case 0:
		for (uint32_t w = 0; w < bsize / 8; w++) {
			uint64_t v = mix64(base + ((uint64_t)w + 1) * GOLD);
			memcpy(out + 8 * (size_t)w, &v, 8);
		}
		break;
	case 1:
		text_fill(base, out, bsize);
		break;
	case 2:
		memset(out, 0, bsize);
		out[0] = (uint8_t)cid;
		out[1] = (uint8_t)(cid >> 8);
		break;
	default:
		text_fill(base, out, bsize / 2);
		memcpy(out + bsize / 2, out, bsize / 2);
		break;
	}
}

void
ef_gen_addr(uint64_t seed, const uint64_t *cids, size_t n, int pshift, uint64_t *offset_out, uint64_t *nhid_out)
{
	for (size_t i = 0; i < n; i++) {
		offset_out[i] = (cids[i] & 16383ULL) << pshift;
		nhid_out[i] = mix64((seed ^ (cids[i] >> 14)) + GOLD);
	}
}

/* Stream ids with a fraction `dup` of same-address repeats: position k repeats a
 * uniformly chosen earlier distinct chunk with probability dup (PRNG stream seed2), else it is the
 * next new chunk.  Returns the number of distinct chunks. */
uint64_t
ef_gen_stream_ids(uint64_t seed2, size_t n, double dup, uint64_t first_cid, uint64_t *cid_out)
{
	uint64_t state = seed2, distinct = 0;
	for (size_t k = 0; k < n; k++) {
		state += GOLD;
		uint64_t r = mix64(state);
		int repeat = distinct > 0 && (double)(r >> 11) * (1.0 / 9007199254740992.0) < dup;
		if (repeat) {
			state += GOLD;
			cid_out[k] = first_cid + mix64(state) % distinct;
		} else {
			cid_out[k] = first_cid + distinct++;
		}
	}
	return distinct;
}

struct gen_job { uint64_t seed; const uint64_t *cids; size_t first, count; uint32_t bsize; uint8_t *out; };

static void *
gen_worker(void *arg)
{
	struct gen_job *j = arg;
	for (size_t i = j->first; i < j->first + j->count; i++)
		ef_gen_chunk(j->seed, j->cids[i], j->bsize, j->out + i * (size_t)j->bsize);
	return NULL;
}

void
ef_gen_chunks(uint64_t seed, const uint64_t *cids, size_t n, uint32_t bsize, uint8_t *out, int threads)
{
	pthread_t tid[256];
	struct gen_job jobs[256];
	if (threads < 1) threads = 1;
	if (threads > 256) threads = 256;
	if ((size_t)threads > n) threads = n ? (int)n : 1;
	size_t per = n / threads, extra = n % threads, at = 0;
	for (int t = 0; t < threads; t++) {
		jobs[t] = (struct gen_job){ seed, cids, at, per + ((size_t)t < extra), bsize, out };
		at += jobs[t].count;
		pthread_create(&tid[t], NULL, gen_worker, &jobs[t]);
	}
	for (int t = 0; t < threads; t++)
		pthread_join(tid[t], NULL);
}

/* ---- parity gate of a measured run ------------------------------------------------------------
 * For each of n pages: block = enc(page) (the reference's LZ4_compress_fast when enc points into
 * oracle/_ref, else the port), expected record = ef_record_prefix(addr, len) + block; compared with
 * the record the GPU store holds (recs + i * rec_stride, rec_lens[i] bytes) and with the stored
 * length the put reported (put_lens[i], ignored when NULL).  out[0] = mismatching chunks,
 * out[1] = index of the first one (or -1), out[2] = total block bytes. */
typedef int (*enc_fn)(const char *src, char *dst, int n, int cap, int accel);
void ef_record_prefix(const void *a, int32_t compressed_length, uint8_t out[24]);

struct par_job {
	enc_fn enc; const uint8_t *pages; size_t first, count; int bsize, accel;
	const uint64_t *addr; const uint8_t *recs; size_t rec_stride; const int32_t *rec_lens, *put_lens;
	uint64_t bad; int64_t first_bad; uint64_t bytes;
};

static void *
par_worker(void *arg)
{
	struct par_job *j = arg;
	uint8_t *blk = malloc((size_t)j->bsize + 1024 + 64);
	uint8_t pre[24];
	j->bad = 0; j->first_bad = -1; j->bytes = 0;
	for (size_t i = j->first; i < j->first + j->count; i++) {
		int len = j->enc((const char *)(j->pages + i * (size_t)j->bsize), (char *)blk, j->bsize, j->bsize + 1024, j->accel);
		ef_record_prefix(j->addr + 2 * i, len, pre);
		const uint8_t *rec = j->recs + i * j->rec_stride;
		int ok = len > 0 && j->rec_lens[i] == 24 + len && (!j->put_lens || j->put_lens[i] == len) &&
		    memcmp(rec, pre, 24) == 0 && memcmp(rec + 24, blk, (size_t)len) == 0;
		j->bytes += len > 0 ? (uint64_t)len : 0;
		if (!ok) {
			if (j->first_bad < 0) j->first_bad = (int64_t)i;
			j->bad++;
		}
	}
	free(blk);
	return NULL;
}

void
ef_parity_records(enc_fn enc, const uint8_t *pages, size_t n, int bsize, int accel, const uint64_t *addr,
    const uint8_t *recs, size_t rec_stride, const int32_t *rec_lens, const int32_t *put_lens, int threads, double out[3])
{
	pthread_t tid[256];
	struct par_job jobs[256];
	if (threads < 1) threads = 1;
	if (threads > 256) threads = 256;
	if ((size_t)threads > n) threads = n ? (int)n : 1;
	size_t per = n / threads, extra = n % threads, at = 0;
	for (int t = 0; t < threads; t++) {
		jobs[t] = (struct par_job){ enc, pages, at, per + ((size_t)t < extra), bsize, accel, addr, recs, rec_stride,
		    rec_lens, put_lens, 0, -1, 0 };
		at += jobs[t].count;
		pthread_create(&tid[t], NULL, par_worker, &jobs[t]);
	}
	out[0] = 0; out[1] = -1; out[2] = 0;
	for (int t = 0; t < threads; t++) {
		pthread_join(tid[t], NULL);
		out[0] += (double)jobs[t].bad;
		if (jobs[t].first_bad >= 0 && out[1] < 0) out[1] = (double)jobs[t].first_bad;
		out[2] += (double)jobs[t].bytes;
	}
}
