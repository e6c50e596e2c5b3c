/*
 * ref_kat.c — prints known answers from the reference's OWN header (uint128.h, included from
 * where it lies under $(REF)/cachemap) for the golden fixture tests/golden/keys.json.
 * TEST INFRASTRUCTURE ONLY.  Built and run by tools/gen_golden.py in the authoring container.
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <uint128.h>

static uint64_t sm(uint64_t *s) {
	uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

int main(void) {
	const char *strs[] = { "", "a", "/bk1", "object-name.bin", "/cl/tn/bk/some%20dir/file" };
	uint64_t h, s = 2024;
	printf("{\"strings\": [");
	for (unsigned i = 0; i < sizeof(strs) / sizeof(strs[0]); i++) {
		FNV_hash(strs[i], (int)strlen(strs[i]), &h);
		printf("%s[\"%s\", \"%016lx\"]", i ? ", " : "", strs[i], (unsigned long)h);
	}
	printf("], \"addrs\": [");
	for (int i = 0; i < 64; i++) {
		uint128_t a;
		a.u = sm(&s);
		a.l = (sm(&s) & ((1ULL << 44) - 1)) | ((sm(&s) & 0xFFFFF) << 44);
		if (i == 0) { a.u = 0x1122334455667788ULL; a.l = (7ULL << 44) | 3; }
		FNV_hash(&a, sizeof(a), &h);
		printf("%s[\"%016lx\", \"%016lx\", \"%016lx\"]", i ? ", " : "", (unsigned long)a.u,
		    (unsigned long)a.l, (unsigned long)h);
	}
	printf("], \"sizeof_uint128\": %zu}\n", sizeof(uint128_t));
	return 0;
}
