// mcb_sketch_lh.cuh — windowed (w,k)-minimizers of one string, host + device.
//
// Restates mm_sketch_lh_ori (sketch.c:116-165), i.e. minimap-0.2 style winnowing with its exact tie rules:
//  * a k-mer equal to its reverse complement is skipped WITHOUT advancing the ring buffer (sketch.c:135);
//  * an ambiguous base resets the run length l but still occupies a ring slot (:139-140);
//  * when the first full window completes, ring entries equal to the current minimum are emitted first (:141-146);
//  * a new element <= the minimum replaces it and the old minimum is emitted once l >= w+k (:147-149);
//  * when the minimum leaves the window it is emitted, the ring is rescanned oldest-to-newest taking the LAST
//    smallest (>=), and other ring entries with the same hash are emitted (:150-161);
//  * the final minimum is emitted at the end (:163-164).
// Emission order is position order, which is what kthread_bucket.c:463 ("first m") depends on.
#pragma once
#include "mcb_common.cuh"

#define MCB_LH_WMAX 256

struct McbLhEmitArray {
	mcb_tuple *out; int64_t cap; int64_t n;
	MCB_HD void operator()(const mcb_tuple &t) { if (n < cap) out[n] = t; ++n; }
};

// `stop_after`: stop as soon as that many tuples were emitted (the index only keeps the first m); <0 = never.
template <class Emit>
MCB_HD void mcb_sketch_lh_core(const char *str, int len, int w, int k, uint32_t rid, mcb_tuple *ring, Emit &emit, int64_t stop_after)
{
	const uint64_t mask = (1ull << (2 * k)) - 1, shift1 = 2 * (k - 1);
	uint64_t fw = 0, rv = 0;
	mcb_tuple mn; mn.x = ~0ull; mn.y = ~0ull;
	int l = 0, bp = 0, mp = 0;
	for (int j = 0; j < w; ++j) { ring[j].x = ~0ull; ring[j].y = ~0ull; }
	for (int i = 0; i < len; ++i) {
		if (stop_after >= 0 && emit.n >= stop_after) return;
		unsigned c = mcb_code_of((unsigned char)str[i]);
		if (c == 5) {   // lower-case input is accepted by seq_nt4_table; map it the same way
			unsigned char u = (unsigned char)str[i];
			c = u == 'a' ? 0u : u == 'c' ? 1u : u == 'g' ? 2u : u == 't' ? 3u : 4u;
		}
		mcb_tuple info; info.x = ~0ull; info.y = ~0ull;
		if (c < 4) {
			fw = (fw << 2 | c) & mask;
			rv = (rv >> 2) | ((3ull ^ c) << shift1);
			if (fw == rv) continue;
			int z = fw < rv ? 0 : 1;
			if (++l >= k) { info.x = mcb_hash64_hd(z ? rv : fw, mask); info.y = (uint64_t)rid << 32 | (uint64_t)(uint32_t)i << 1 | (uint64_t)z; }
		} else l = 0;
		ring[bp] = info;
		if (l == w + k - 1) {
			for (int j = bp + 1; j < w; ++j) if (mn.x == ring[j].x && ring[j].y != mn.y) emit(ring[j]);
			for (int j = 0; j < bp; ++j) if (mn.x == ring[j].x && ring[j].y != mn.y) emit(ring[j]);
		}
		if (info.x <= mn.x) {
			if (l >= w + k) emit(mn);
			mn = info; mp = bp;
		} else if (bp == mp) {
			if (l >= w + k - 1) emit(mn);
			mn.x = ~0ull;
			for (int j = bp + 1; j < w; ++j) if (mn.x >= ring[j].x) { mn = ring[j]; mp = j; }
			for (int j = 0; j <= bp; ++j) if (mn.x >= ring[j].x) { mn = ring[j]; mp = j; }
			if (l >= w + k - 1) {
				for (int j = bp + 1; j < w; ++j) if (mn.x == ring[j].x && mn.y != ring[j].y) emit(ring[j]);
				for (int j = 0; j <= bp; ++j) if (mn.x == ring[j].x && mn.y != ring[j].y) emit(ring[j]);
			}
		}
		if (++bp == w) bp = 0;
	}
	if (stop_after >= 0 && emit.n >= stop_after) return;
	if (mn.x != ~0ull) emit(mn);
}
