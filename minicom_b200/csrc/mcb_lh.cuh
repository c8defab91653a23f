// mcb_lh.cuh — mm_sketch_lh_ori (sketch.c:116-165) on the device: one thread walks one contig.
// Used by kt_for_bucket (first m minimizers of every new seed contig, kthread_bucket.c:458-474) and by the contig merge
// (all minimizers of every contig, kthread_cb.c:234; first m of every new contig, :365,:418).
#pragma once
#include "mcb_common.cuh"

#define LH_THREADS 64
struct LhRingSmem {
	uint64_t *x; uint32_t *ps; int tid;
	__device__ __forceinline__ void set(int j, uint64_t hx, uint32_t p) { x[j * LH_THREADS + tid] = hx; ps[j * LH_THREADS + tid] = p; }
	__device__ __forceinline__ uint64_t hx(int j) const { return x[j * LH_THREADS + tid]; }
	__device__ __forceinline__ uint32_t p(int j) const { return ps[j * LH_THREADS + tid]; }
};

// mm_sketch_lh_ori (sketch.c:116-165) is a sequential scan with data-dependent tie rules, so one thread walks one contig; its
// characters arrive through aligned 64-bit loads, and the w-slot ring buffer (hash + position/strand per slot) lives in shared
// memory, slot-major so that the lanes of a warp hit different banks.  The walk stops after m outputs (kthread_bucket.c:463).
// The window minimum is found in constant time.  The reference rescans the whole ring whenever the minimum leaves the window
// (sketch.c:150-158); with 32 contigs per warp some lane needs that at almost every step, so every warp paid the w-slot scan
// (and the tie scan after it) all the time.  Here the ring is cut into blocks of w steps (one turn of bp): when bp wraps, one
// backward pass stores for every slot j the slot of the rightmost minimum of the previous block's slots [j, w) (strict <
// from the right = the reference's ">=" from the left); a running rightmost minimum P covers the slots [0, bp] written in the
// current block.  The minimum of the window is then min(S[bp+1], P), ties going to P (newer).  Both carry a "some other slot
// of the range holds the same hash" bit, so the reference's scan for identical k-mers (:159-161) runs only when there is one.
// bp is the same in all lanes of a warp unless a lane meets a symmetric k-mer (even k only), so the passes do not diverge.
#define LH2_SLOT_MASK 0x7Fu
#define LH2_WMAX MCB_LH_WMAX
template <bool WIDE>
__global__ void __launch_bounds__(LH_THREADS)
k_sketch_lh2(const char *__restrict__ cl_ref, const uint64_t *__restrict__ cl_ref_off, uint64_t cl_first, uint64_t cl_count, uint64_t cid_first,
             int w, int k, int m, mcb_tuple *__restrict__ mi, uint8_t *__restrict__ mi_cnt, const uint64_t *__restrict__ out_off, uint32_t *__restrict__ cnt32,
             const uint32_t *__restrict__ ch_contig = nullptr, const uint32_t *__restrict__ ch_start = nullptr, int ch_len = 0, int slot_cap = 0)
{
	extern __shared__ __align__(16) unsigned char lh_smem[];
	uint64_t *rx = (uint64_t*)lh_smem;                                        // [w][LH_THREADS]
	uint32_t *rp = (uint32_t*)(rx + (size_t)w * LH_THREADS);                  // [w][LH_THREADS]
	uint8_t *ss = (uint8_t*)(rp + (size_t)w * LH_THREADS);                    // [w][LH_THREADS] suffix-minimum slot | tie << 7
	const uint64_t ci = (uint64_t)blockIdx.x * LH_THREADS + threadIdx.x;
	if (ci >= cl_count) return;
	// Chunk mode (ALL mode, odd k, no ambiguous bases: the contig merge): work item ci is the stretch [start, start + ch_len) of
	// contig ch_contig[ci].  With odd k no k-mer is its own reverse complement, so every base occupies a ring slot, the current
	// minimum is the rightmost minimum of the last w k-mers, and what the walk emits at a step depends on the last w + 1 k-mers
	// and the position only: a walk that starts w + k - 1 bases early, counts positions from there and stays silent until
	// `start` emits exactly what the walk from the beginning of the string emits inside the stretch.
	const uint64_t crel = ch_contig ? ch_contig[ci] : ci;
	const uint64_t c = cl_first + crel;
	const uint64_t b = cl_ref_off[c], e = cl_ref_off[c + 1];
	const int len = (int)(e - b);
	const int start = ch_contig ? (int)ch_start[ci] : 0;
	const int end = ch_contig ? min(len, start + ch_len) : len;
	const int b0 = max(0, start - (w + k - 1));
	const uint32_t rid = (uint32_t)((cid_first + crel) << 8);        // ((clusters.n-1)<<8)+tid, tid 0 (kthread_bucket.c:458)
	const uint64_t *str8 = (const uint64_t*)(cl_ref + (b & ~(uint64_t)7));
	const int skew = (int)(b & 7);
	uint64_t chunk = 0;
	LhRingSmem ring; ring.x = rx; ring.ps = rp; ring.tid = threadIdx.x;
	uint8_t *sst = ss + threadIdx.x;
	// two uses: the first m tuples of contig c at mi[c*m ..] (kthread_bucket.c:463; mi_cnt given), or ALL tuples (m = INT_MAX, cnt32
	// given) at mi[out_off[ci] ..] — a pass with mi == null only counts them; with slot_cap > 0 the pass counts AND parks the first
	// slot_cap tuples of the item at mi[ci * slot_cap ..] (the caller compacts the slots once the counts are scanned)
	mcb_tuple *out = cnt32 ? (mi ? (slot_cap ? mi + ci * (uint64_t)slot_cap : mi + out_off[ci]) : nullptr) : mi + c * m;
	const int wlim = slot_cap ? slot_cap : m;
	int n_out = 0;
	const uint64_t mask = (1ull << (2 * k)) - 1, shift1 = 2 * (k - 1);
	const uint32_t mh = (uint32_t)(mask >> 32);
	const int s1 = WIDE ? 2 * (k - 1) - 32 : 0;
	const uint32_t t3 = 3u << s1;
	uint64_t fw = 0, rv = 0;
	uint32_t flo = 0, fhi = 0, rlo = 0, rhi = 0;
	uint64_t mn_x = ~0ull; uint32_t mn_p = ~0u;
	uint64_t px = ~0ull; int pslot = 0; bool ptie = false;             // rightmost minimum of the slots written in this block
	int l = b0, seen = 0, bp = 0, mp = 0;                               // l: bases since the last ambiguous one (= position here); seen: bases rolled into fw / rv
	bool on = b0 >= start;
#define LH_EMIT(hx_, p_) do { if (on) { if (out && n_out < wlim) { mcb_tuple t_; t_.x = (hx_); t_.y = (uint64_t)rid << 32 | (uint64_t)(p_); out[n_out] = t_; } ++n_out; } } while (0)
	for (int j = 0; j < w; ++j) { ring.set(j, ~0ull, ~0u); sst[j * LH_THREADS] = (uint8_t)(w - 1); }
	for (int i = b0; i < end && n_out < m; ++i) {
		const int ai = i + skew;
		if (i == b0 || (ai & 7) == 0) chunk = str8[ai >> 3];
		on = i >= start;
		const unsigned cc = mcb_code_of((unsigned char)(chunk >> (8 * (ai & 7))));   // consensus strings are upper-case ACGT (invert_code_rule)
		uint64_t ix = ~0ull; uint32_t ip = ~0u;
		if (cc < 4) {
			if (WIDE) {
				fhi = __funnelshift_l(flo, fhi, 2) & mh; flo = flo * 4u + cc;
				rlo = __funnelshift_r(rlo, rhi, 2); rhi = (rhi >> 2) | ((cc << s1) ^ t3);
				fw = (uint64_t)fhi << 32 | flo; rv = (uint64_t)rhi << 32 | rlo;
			} else {
				fw = (fw << 2 | cc) & mask;
				rv = (rv >> 2) | ((3ull ^ cc) << shift1);
			}
			if (fw == rv) continue;
			const int z = fw < rv ? 0 : 1;
			++l;
			if (++seen >= k) { ix = WIDE ? mcb_hash64_wide(z ? rv : fw, mh) : mcb_hash64_hd(z ? rv : fw, mask); ip = (uint32_t)i << 1 | (uint32_t)z; }
		} else { l = 0; seen = 0; }
		ring.set(bp, ix, ip);
		{
			const bool ple = ix <= px;
			ptie = ple ? (ix == px) : ptie; px = ple ? ix : px; pslot = ple ? bp : pslot;
		}
		if (l == w + k - 1) {
			for (int j = bp + 1; j < w; ++j) if (mn_x == ring.hx(j) && ring.p(j) != mn_p) LH_EMIT(ring.hx(j), ring.p(j));
			for (int j = 0; j < bp; ++j) if (mn_x == ring.hx(j) && ring.p(j) != mn_p) LH_EMIT(ring.hx(j), ring.p(j));
		}
		if (ix <= mn_x) {
			if (l >= w + k) LH_EMIT(mn_x, mn_p);
			mn_x = ix; mn_p = ip; mp = bp;
		} else if (bp == mp) {
			if (l >= w + k - 1) LH_EMIT(mn_x, mn_p);
			uint64_t nx = px; int ns = pslot; bool tie = ptie;
			if (bp + 1 < w) {
				const unsigned sv = sst[(bp + 1) * LH_THREADS];
				const int sslot = (int)(sv & LH2_SLOT_MASK);
				const uint64_t sx = ring.hx(sslot);
				if (sx < px) { nx = sx; ns = sslot; tie = (sv >> 7) != 0; }
				else if (sx == px) tie = true;
			}
			mn_x = nx; mp = ns; mn_p = ring.p(ns);
			if (tie && l >= w + k - 1) {
				for (int j = bp + 1; j < w; ++j) if (mn_x == ring.hx(j) && mn_p != ring.p(j)) LH_EMIT(ring.hx(j), ring.p(j));
				for (int j = 0; j <= bp; ++j) if (mn_x == ring.hx(j) && mn_p != ring.p(j)) LH_EMIT(ring.hx(j), ring.p(j));
			}
		}
		if (++bp == w) {
			bp = 0;
			uint64_t sx = ring.hx(w - 1); unsigned sv = (unsigned)(w - 1);
			sst[(w - 1) * LH_THREADS] = (uint8_t)sv;
			for (int j = w - 2; j >= 1; --j) {
				const uint64_t v = ring.hx(j);
				if (v < sx) { sx = v; sv = (unsigned)j; }
				else if (v == sx) sv |= 0x80u;
				sst[j * LH_THREADS] = (uint8_t)sv;
			}
			px = ~0ull; pslot = 0; ptie = false;
		}
	}
	on = end == len;                                                    // the closing minimum (sketch.c:163-164) belongs to the last stretch
	if (n_out < m && mn_x != ~0ull) LH_EMIT(mn_x, mn_p);
#undef LH_EMIT
	if (mi_cnt) mi_cnt[c] = (uint8_t)(n_out < m ? n_out : m);
	if (cnt32) cnt32[ci] = (uint32_t)n_out;
}
