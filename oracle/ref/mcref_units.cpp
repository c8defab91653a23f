// TEST INFRASTRUCTURE (oracle/_ref build only; never linked into the product).
// C entry points around individual functions of the UNMODIFIED reference, compiled from the sources where they lie
// (/root/reference/src/sketch.c, misc.c), so unit tests can call the reference itself through ctypes:
//   hash64 (sketch.c:27, static inline -> reachable by including the translation unit), mm_sketch_two (sketch.c:238),
//   mm_sketch_lh_ori (sketch.c:116), radix_sort_128x (misc.c:21-22 / ksort.h:108-157).
#include <stdint.h>
#include <string.h>
#include "sketch.c"      // found via -I/root/reference/src (brings minicom.h: mm128_t, mm128_v)

extern "C" {

uint64_t ref_hash64(uint64_t key, uint64_t mask) { return hash64(key, mask); }

void ref_sketch_two(const char *str, int len, int k, uint32_t rid, uint64_t *xy)
{
	mm128_t m;
	mm_sketch_two(str, len, k, rid, &m);
	xy[0] = m.x; xy[1] = m.y;
}

int64_t ref_sketch_lh_ori(const char *str, int len, int w, int k, uint32_t rid, uint64_t *xy, int64_t cap)
{
	mm128_v v = {0, 0, 0};
	mm_sketch_lh_ori(str, len, w, k, rid, &v);
	for (size_t i = 0; i < v.n && (int64_t)i < cap; ++i) { xy[2 * i] = v.a[i].x; xy[2 * i + 1] = v.a[i].y; }
	int64_t n = (int64_t)v.n;
	free(v.a);
	return n;
}

unsigned char ref_nt4(unsigned char c) { return seq_nt4_table[c]; }

}
