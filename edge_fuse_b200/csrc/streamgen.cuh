// streamgen.cuh — deterministic synthetic chunk streams, one definition shared by
// the host generator (CPU baseline input, small tests) and the device generator (HBM-resident
// benchmark input).  Bench / test utility; not part of the reference's API.
//
// Chunk `cid` of a stream seeded `seed` (default 42) has content class (cid + (cid >> 3)) & 3 (the
// four classes take turns, and the turn order drifts every 8 chunks so that sharding chunk k to
// GPU k mod G never pins a class to a GPU):
//   0 R  incompressible: 8-byte words word(w) = mix(base + (w+1)*G)
//   1 T  low-entropy text: byte j takes 16 bits r of word(j/4) (field j%4);
//        r&3 != 0 -> 'a' + ((r>>2)&3)  (probability 3/4), else the random byte r>>8
//   2 Z  zero page stamped with cid (bytes 0,1 = cid little-endian)
//   3 M  first half as T, second half repeats the first half
// with base = mix(seed ^ cid*0xD1B54A32D192ED03), G = 0x9E3779B97F4A7C15 and mix = the splitmix64
// output function.  Addresses: object = cid >> 14, page = cid & 16383,
// nhid_small = mix((seed ^ object) + G), offset = page << pshift, genid = 0.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SG_HD __host__ __device__ __forceinline__
#else
#define SG_HD static inline
#endif

#define SG_GOLDEN 0x9E3779B97F4A7C15ULL

SG_HD uint64_t sg_mix(uint64_t z) {
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}
SG_HD uint32_t sg_class(uint64_t cid) { return (uint32_t)((cid + (cid >> 3)) & 3); }
SG_HD uint64_t sg_base(uint64_t seed, uint64_t cid) { return sg_mix(seed ^ (cid * 0xD1B54A32D192ED03ULL)); }
SG_HD uint64_t sg_word(uint64_t base, uint64_t w) { return sg_mix(base + (w + 1) * SG_GOLDEN); }
SG_HD uint8_t sg_text_byte(uint64_t base, uint32_t j) {
	uint32_t r = (uint32_t)(sg_word(base, j >> 2) >> (16 * (j & 3))) & 0xFFFFu;
	return (r & 3u) ? (uint8_t)('a' + ((r >> 2) & 3u)) : (uint8_t)(r >> 8);
}
// The 8 bytes [8*w8, 8*w8+8) of chunk cid, little-endian packed.
SG_HD uint64_t sg_chunk_word(uint64_t seed, uint64_t cid, uint32_t bsize, uint32_t w8) {
	uint64_t base = sg_base(seed, cid);
	uint32_t cls = sg_class(cid);
	if (cls == 0) return sg_word(base, w8);
	if (cls == 2) return w8 == 0 ? (cid & 0xFFFFu) : 0;
	uint32_t j0 = w8 * 8;
	if (cls == 3 && j0 >= bsize / 2) j0 -= bsize / 2;
	uint64_t v = 0;
	for (uint32_t k = 0; k < 8; k++) v |= (uint64_t)sg_text_byte(base, j0 + k) << (8 * k);
	return v;
}
SG_HD uint64_t sg_nhid(uint64_t seed, uint64_t cid) { return sg_mix((seed ^ (cid >> 14)) + SG_GOLDEN); }
SG_HD uint64_t sg_offset(uint64_t cid, int pshift) { return (cid & 16383ULL) << pshift; }
