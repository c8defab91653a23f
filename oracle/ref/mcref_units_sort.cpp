// TEST INFRASTRUCTURE (oracle/_ref build only).  radix_sort_128x of the UNMODIFIED reference (misc.c:21-22, ksort.h:108-157)
// behind a C entry point.  Separate translation unit because breads.h and minicom.h (pulled in by sketch.c) cannot be
// included together.
#include <stdint.h>
#include "breads.h"

extern "C" void ref_radix_sort_128x(uint64_t *xy, int64_t n)
{
	radix_sort_128x((mm128_t*)xy, (mm128_t*)xy + n);
}
