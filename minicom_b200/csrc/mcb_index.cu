// mcb_index.cu — mm_idx_generation / mm_idx_get (kthread_idx.c:116-173, :84-101) on the device.
//
// The reference sorts every bucket with radix_sort_128x (misc.c:21-22, ksort.h:108-157): insertion sort up to 64
// elements (stable), above that an in-place most-significant-digit "American flag" sort whose cycle-leader permutation
// is NOT stable.  The order in which equal minimizers end up is observable: worker_post copies them into the posting
// list in that order (kthread_idx.c:153-157) and the contig merger takes the first acceptable hit
// (kthread_cb.c:267-343).  So the device index reproduces that permutation exactly: one warp per bucket, the
// data-dependent cycle walk on one lane, histograms / small-segment sorts / CSR construction on all lanes.
// The hash table of the reference (khash) is replaced by sorted distinct keys per bucket + CSR postings.
#include "mcb_common.cuh"
#include <chrono>
#include <stdlib.h>
#include <thread>
#include <algorithm>

#define IX_SMALL 64       // RS_MIN_SIZE (ksort.h:106)

struct IxSeg { uint32_t b, e; int s; };

// stable sort of a short segment by x: rank = #smaller + #equal-before.  Same result as the reference's insertion sort.
__device__ void ix_small_sort_warp(mcb_tuple *a, uint32_t n, int lane)
{
	// n <= 64: each lane owns elements lane and lane+32
	mcb_tuple e0, e1; bool h0 = lane < (int)n, h1 = lane + 32 < (int)n;
	if (h0) e0 = a[lane];
	if (h1) e1 = a[lane + 32];
	uint32_t r0 = 0, r1 = 0;
	for (uint32_t j = 0; j < n; ++j) {
		uint64_t xj = a[j].x;
		if (h0) r0 += (xj < e0.x) || (xj == e0.x && j < (uint32_t)lane);
		if (h1) r1 += (xj < e1.x) || (xj == e1.x && j < (uint32_t)lane + 32);
	}
	__syncwarp();
	if (h0) a[r0] = e0;
	if (h1) a[r1] = e1;
	__syncwarp();
}
#define IX_SMEM_STK 32

// ---------------------------------------------------------------- index-space formulation of the same sort
// The cycle-leader walk of ksort.h:131-145 only looks at digits: a slot of a digit region is visited once, left to right, and
// until then holds its ORIGINAL element.  So the walk can run on a byte array of digits and produce, for every original
// position, its destination (`dest`), touching 1-2 bytes per step instead of moving 16-byte tuples; the tuples are then
// moved by all lanes at once.  Per warp this needs 3 bytes per tuple of shared memory instead of 16, so 4-5x more buckets
// are in flight per SM — and the walk, a chain of dependent shared-memory loads, is bound by how many run concurrently.
// Tuples go through a scratch copy in global memory (L2-resident per bucket).
#define IX2_WARPS 4
template <class DT> struct Ix2Mem { uint32_t *cur, *end; DT *dest; uint8_t *dg; IxSeg *stk; };

// Only the digits and destinations live in shared memory (3 bytes per tuple); the keys needed to rank the small regions are
// read back from global memory after the walk's permutation has been applied, where a region is a contiguous, cache-resident run.
template <class DT>
__device__ void ix3_flag_sort(mcb_tuple *a, mcb_tuple *tmp, uint32_t n, const Ix2Mem<DT> &m, IxSeg *gstk, int lane)
{
	IxSeg *stk = m.stk;
	int top = 1;
	if (lane == 0) { stk[0].b = 0; stk[0].e = n; stk[0].s = 56; }
	__syncwarp();
	while (top > 0) {
		--top;
		const IxSeg sg = top < IX_SMEM_STK ? stk[top] : gstk[top - IX_SMEM_STK];
		const uint32_t sb = sg.b, se = sg.e, cnt = se - sb; const int s = sg.s;
		__syncwarp();
		for (int d = lane; d < 256; d += 32) m.cur[d] = 0;
		__syncwarp();
		for (uint32_t i = lane; i < cnt; i += 32) { const unsigned d = (unsigned)(a[sb + i].x >> s) & 255u; m.dg[i] = (uint8_t)d; atomicAdd(&m.cur[d], 1u); }
		__syncwarp();
		{
			uint32_t v[8], sum = 0;
#pragma unroll
			for (int q = 0; q < 8; ++q) { v[q] = m.cur[lane * 8 + q]; sum += v[q]; }
			uint32_t inc = sum;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { uint32_t x = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += x; }
			uint32_t run = inc - sum;
			__syncwarp();
#pragma unroll
			for (int q = 0; q < 8; ++q) { m.cur[lane * 8 + q] = run; run += v[q]; m.end[lane * 8 + q] = run; }
		}
		__syncwarp();
		if (lane == 0) {                          // the walk of ksort.h:131-145, on digits only
			for (int k = 0; k < 256;) {
				const uint32_t ck = m.cur[k];
				if (ck != m.end[k]) {
					uint32_t t = ck;
					int l = m.dg[t];
					if (l != k) {
						do {
							const uint32_t p = m.cur[l]; m.cur[l] = p + 1;
							m.dest[t] = (DT)p;
							t = p;
							l = m.dg[t];
						} while (l != k);
						m.dest[t] = (DT)m.cur[k]; m.cur[k] = m.cur[k] + 1;
					} else { m.dest[t] = (DT)ck; m.cur[k] = ck + 1; }
				} else ++k;
			}
		}
		__syncwarp();
		for (uint32_t i = lane; i < cnt; i += 32) tmp[sb + m.dest[i]] = a[sb + i];
		__syncwarp();
		for (uint32_t i = lane; i < cnt; i += 32) a[sb + i] = tmp[sb + i];
		__syncwarp();
		if (s > 0) {
			// regions of at most 64 tuples: stable rank by the whole key (= the reference's insertion sort), keys from global memory
			for (uint32_t p = lane; p < cnt; p += 32) {
				const uint64_t xi = a[sb + p].x;
				const unsigned d = (unsigned)(xi >> s) & 255u;
				const uint32_t rs = d ? m.end[d - 1] : 0u, re = m.end[d];
				uint32_t at = p;
				if (re - rs <= IX_SMALL && re - rs > 1) {
					uint32_t rank = 0;
					for (uint32_t q = rs; q < re; ++q) { const uint64_t xq = a[sb + q].x; rank += (xq < xi) || (xq == xi && q < p); }
					at = rs + rank;
				}
				m.dest[p] = (DT)at;
			}
			__syncwarp();
			for (uint32_t p = lane; p < cnt; p += 32) tmp[sb + m.dest[p]] = a[sb + p];
			__syncwarp();
			for (uint32_t p = lane; p < cnt; p += 32) a[sb + p] = tmp[sb + p];
			__syncwarp();
			const int s2 = s > 8 ? s - 8 : 0;
			for (int d0 = 0; d0 < 256; d0 += 32) {
				const int d = d0 + lane;
				const uint32_t rb = sb + (d == 0 ? 0u : m.end[d - 1]), re = sb + m.end[d];
				const bool big = re - rb > IX_SMALL;
				const unsigned bm = __ballot_sync(0xFFFFFFFFu, big);
				if (big) {
					const int slot = top + __popc(bm & ((1u << lane) - 1u));
					IxSeg ns; ns.b = rb; ns.e = re; ns.s = s2;
					if (slot < IX_SMEM_STK) stk[slot] = ns; else gstk[slot - IX_SMEM_STK] = ns;
				}
				top += __popc(bm);
			}
		}
		__syncwarp();
	}
}

// BIG = false: IX2_WARPS buckets per CTA, digits and 16-bit destinations in shared memory, buckets above `cap` tuples skipped.
// BIG = true: the skipped buckets, one warp per CTA, digits and 32-bit destinations in global scratch (dg_g / dest_g, one slot
// per tuple) — rare (a bucket above ~17 000 tuples), and correct for any size.
template <bool BIG>
__global__ void __launch_bounds__(IX2_WARPS * 32)
k_index_sort3(mcb_tuple *__restrict__ t, mcb_tuple *__restrict__ tmp, const uint64_t *__restrict__ boff, int nb, uint32_t cap, IxSeg *__restrict__ stacks,
              uint8_t *__restrict__ dg_g, uint32_t *__restrict__ dest_g)
{
	extern __shared__ __align__(16) unsigned char ix2_smem[];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const size_t per_warp = (2048 + IX_SMEM_STK * sizeof(IxSeg) + (BIG ? 0 : (size_t)cap * 3) + 15) & ~(size_t)15;
	unsigned char *base = ix2_smem + (size_t)wib * per_warp;
	const int nw = BIG ? 1 : IX2_WARPS;
	if (BIG && wib) return;
	for (int bk = blockIdx.x * nw + wib; bk < nb; bk += gridDim.x * nw) {
		const uint64_t B0 = boff[bk], B1 = boff[bk + 1];
		const uint32_t n = (uint32_t)(B1 - B0);
		if (BIG) {
			if (n <= cap) continue;
			Ix2Mem<uint32_t> m;
			m.cur = (uint32_t*)base; m.end = m.cur + 256; m.stk = (IxSeg*)(base + 2048);
			m.dest = dest_g + B0; m.dg = dg_g + B0;
			ix3_flag_sort<uint32_t>(t + B0, tmp + B0, n, m, stacks + (B0 / 65 + 8ull * bk), lane);
		} else {
			if (n <= 1 || n > cap) continue;
			if (n <= IX_SMALL) { ix_small_sort_warp(t + B0, n, lane); continue; }
			Ix2Mem<uint16_t> m;
			m.cur = (uint32_t*)base; m.end = m.cur + 256; m.stk = (IxSeg*)(base + 2048);
			m.dest = (uint16_t*)(base + 2048 + IX_SMEM_STK * sizeof(IxSeg)); m.dg = (uint8_t*)(m.dest + cap);
			ix3_flag_sort<uint16_t>(t + B0, tmp + B0, n, m, stacks + (B0 / 65 + 8ull * bk), lane);
		}
	}
}

struct mcb_index {
	int b = 14;
	uint64_t n_keys = 0, n_post = 0;
	uint64_t *keys = nullptr;      // [n_keys]   distinct minimizers, ascending inside each bucket
	uint32_t *kstart = nullptr;    // [n_keys+1] posting offsets
	uint64_t *post = nullptr;      // [n_post]   y values in the reference's order
	uint32_t *ub = nullptr;        // [2^b+1]    key range of each bucket
	HBuf slab;                     // one pinned allocation behind the four arrays (recycled through the context's pool)
	McbPinnedPool *pool = nullptr;
	std::vector<mcb_index*> parts; // several GPUs (mcb_group.cu): part r holds the buckets rank r owns (bucket * G >> b); this node holds nothing itself
};

mcb_index *mcb_index_make_group(std::vector<mcb_index*> &parts, int b)
{
	mcb_index *ix = new mcb_index();
	ix->b = b; ix->parts = parts;
	for (auto p : parts) { ix->n_keys += p->n_keys; ix->n_post += p->n_post; }
	return ix;
}

extern "C" void mcb_idx_destroy(mcb_index *ix)
{
	if (!ix) return;
	for (auto p : ix->parts) mcb_idx_destroy(p);
	if (ix->pool) { if (ix->pool->give(ix->slab)) delete ix->pool; }
	else ix->slab.release();
	delete ix;
}

extern "C" void mcb_idx_stats(const mcb_index *ix, uint64_t *n_keys, uint64_t *n_post)
{
	if (n_keys) *n_keys = ix ? ix->n_keys : 0;
	if (n_post) *n_post = ix ? ix->n_post : 0;
}

extern "C" void mcb_idx_arrays(const mcb_index *ix, const uint64_t **keys, const uint32_t **kstart, const uint64_t **post, const uint32_t **bucket_keys)
{
	if (ix && !ix->parts.empty()) ix = nullptr;          // a group index has no flat arrays of its own
	if (keys) *keys = ix ? ix->keys : nullptr;
	if (kstart) *kstart = ix ? ix->kstart : nullptr;
	if (post) *post = ix ? ix->post : nullptr;
	if (bucket_keys) *bucket_keys = ix ? ix->ub : nullptr;
}

extern "C" const uint64_t *mcb_idx_get(const mcb_index *ix, uint64_t minier, int *n)
{
	*n = 0;
	if (!ix || !ix->n_keys) return nullptr;
	const uint32_t bk = (uint32_t)(minier & ((1ull << ix->b) - 1));
	if (!ix->parts.empty()) return mcb_idx_get(ix->parts[((uint64_t)bk * ix->parts.size()) >> ix->b], minier, n);
	uint32_t lo = ix->ub[bk], hi = ix->ub[bk + 1];
	while (lo < hi) {
		uint32_t mid = lo + ((hi - lo) >> 1);
		uint64_t kx = ix->keys[mid];
		if (kx < minier) lo = mid + 1; else if (kx > minier) hi = mid; else {
			*n = (int)(ix->kstart[mid + 1] - ix->kstart[mid]);
			return ix->post + ix->kstart[mid];
		}
	}
	return nullptr;
}

// carve the four host arrays out of one pinned slab
static int idx_alloc_host(mcb_ctx *ctx, mcb_index *ix, uint64_t U, uint64_t n, int nb)
{
	const size_t o_keys = 0, o_post = o_keys + (U + 1) * 8, o_ks = o_post + (n + 1) * 8, o_ub = (o_ks + (U + 2) * 4 + 7) & ~(size_t)7, tot = o_ub + ((size_t)nb + 2) * 4;
	ix->pool = ctx->pool;
	ix->slab = ctx->pool->take(tot);
	MCB_TRY(ix->slab.ensure(tot));
	char *base = ix->slab.as<char>();
	ix->keys = (uint64_t*)(base + o_keys); ix->post = (uint64_t*)(base + o_post); ix->kstart = (uint32_t*)(base + o_ks); ix->ub = (uint32_t*)(base + o_ub);
	return MCB_OK;
}

// ---------------------------------------------------------------- build
// The tuples cross PCIe in (16 B each) and the postings and keys cross it out (8 + 12 B), and both dwarf the sort itself.  The
// build can therefore be cut into runs of buckets (MCB_IX_RUNS, at least a million tuples each): run c is uploaded on one copy
// stream, sorted and turned into keys/postings on the compute stream, and downloaded on a second copy stream while run c+1 is
// on its way in.  Key numbers are global: a device-side chain carries the distinct-key count from run to run (chain[c] = first
// key of run c).  Measured on C2 (seven builds of 1-7 M tuples per step, r2c): six runs save 1.8 ms of the 17.9 ms host time of
// the builds but triple their device time (a third of the buckets no longer fills the SMs with one warp per bucket), so the
// default is ONE run; the knob is for hosts whose builds are tens of millions of tuples.
// distinct keys of every (sorted) bucket: one warp per bucket
#define IXF_WARPS 8
__global__ void __launch_bounds__(IXF_WARPS * 32)
k_ix_count(const mcb_tuple *__restrict__ t, const uint64_t *__restrict__ boff, int nb, uint32_t *__restrict__ cntk)
{
	const int lane = threadIdx.x & 31, bk = blockIdx.x * IXF_WARPS + (threadIdx.x >> 5);
	if (bk >= nb) return;
	const uint64_t B0 = boff[bk], n = boff[bk + 1] - B0;
	uint32_t tot = 0;
	for (uint64_t i0 = 0; i0 < n; i0 += 32) {
		const uint64_t i = i0 + lane;
		const bool head = i < n && (i == 0 || t[B0 + i].x != t[B0 + i - 1].x);
		tot += __popc(__ballot_sync(0xFFFFFFFFu, head));
	}
	if (lane == 0) cntk[bk] = tot;
}
// first key number of every bucket of a run: exclusive scan of the counts (one CTA; a run has at most 2^b buckets) on top of
// the keys of the earlier runs (chain[c]); chain[c+1] = first key of the next run
__global__ void __launch_bounds__(1024)
k_ix_scan_run(const uint32_t *__restrict__ cntk, int b0, int b1, unsigned long long *__restrict__ chain, int c, uint32_t *__restrict__ ub, int nb_total_if_last)
{
	__shared__ uint32_t wtot[32], wexc[32];
	__shared__ uint32_t carry_s, tile_total;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	if (threadIdx.x == 0) carry_s = (uint32_t)chain[c];
	__syncthreads();
	for (int base = b0; base < b1; base += 1024) {
		const int bidx = base + threadIdx.x;
		const uint32_t v = bidx < b1 ? cntk[bidx] : 0u;
		uint32_t inc = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += x; }
		if (lane == 31) wtot[wid] = inc;
		__syncthreads();
		if (wid == 0) {
			const uint32_t w = wtot[lane];
			uint32_t winc = w;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, winc, o); if (lane >= o) winc += x; }
			wexc[lane] = winc - w;
			if (lane == 31) tile_total = winc;
		}
		__syncthreads();
		const uint32_t carry = carry_s;
		if (bidx < b1) ub[bidx] = carry + wexc[wid] + inc - v;
		__syncthreads();
		if (threadIdx.x == 0) carry_s = carry + tile_total;
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		chain[c + 1] = carry_s;
		if (nb_total_if_last >= 0) ub[nb_total_if_last] = carry_s;
	}
}
// keys, posting offsets and postings of every bucket: one warp per bucket
__global__ void __launch_bounds__(IXF_WARPS * 32)
k_ix_finalize(const mcb_tuple *__restrict__ t, const uint64_t *__restrict__ boff, int nb, const uint32_t *__restrict__ ub,
              uint64_t *__restrict__ keys, uint32_t *__restrict__ kstart, uint64_t *__restrict__ post)
{
	const int lane = threadIdx.x & 31, bk = blockIdx.x * IXF_WARPS + (threadIdx.x >> 5);
	if (bk >= nb) return;
	const uint64_t B0 = boff[bk], n = boff[bk + 1] - B0;
	uint32_t at = ub[bk];
	for (uint64_t i0 = 0; i0 < n; i0 += 32) {
		const uint64_t i = i0 + lane;
		mcb_tuple e; e.x = 0; e.y = 0;
		bool head = false;
		if (i < n) { e = t[B0 + i]; head = i == 0 || e.x != t[B0 + i - 1].x; post[B0 + i] = e.y; }
		const unsigned m = __ballot_sync(0xFFFFFFFFu, head);
		if (head) { const uint32_t u = at + __popc(m & ((1u << lane) - 1u)); keys[u] = e.x; kstart[u] = (uint32_t)(B0 + i); }
		at += __popc(m);
	}
}

// posting offsets end at the number of tuples: kstart[number of keys] = n
__global__ void k_ix_close(const uint32_t *__restrict__ ub, int nb, uint32_t *__restrict__ kstart, uint32_t n) { kstart[ub[nb]] = n; }

struct IxSource {            // where the tuples of a bucket range come from
	const mcb_tuple *flat = nullptr;                                    // bucket-major array, or
	const mcb_tuple *const *ptrs = nullptr; const uint64_t *cnt = nullptr; // one array per bucket (mm_idx_t::B[i].a), gathered through `stage`
	mcb_tuple *stage = nullptr; int n_threads = 1;
};

struct IxEvPair { cudaEvent_t a, b; };
static void ix_timer_add(mcb_ctx *ctx, const char *name, std::vector<IxEvPair> &v)
{
	float tot = 0;
	for (auto &p : v) { float ms = 0; if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) tot += ms; else cudaGetLastError(); cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
	if (ctx->tm.enabled && !v.empty()) { int id = ctx->tm.id(name); ctx->tm.ms[id] += tot; ctx->tm.cnt[id] += 1; }
	v.clear();
}

static int idx_build_pipelined(mcb_ctx *ctx, uint64_t n, const uint64_t *h_boff, const IxSource &src, mcb_index **out)
{
	const int b = ctx->prm.b, nb = 1 << b;
	mcb_index *ix = new mcb_index(); ix->b = b; ix->n_post = n;
	*out = ix;
	if (n == 0) {
		MCB_TRY(idx_alloc_host(ctx, ix, 0, 0, nb));
		memset(ix->ub, 0, ((size_t)nb + 1) * 4); ix->kstart[0] = 0;
		return MCB_OK;
	}
	if (n >= 0xFFFFFFFFull) { mcb_set_error("index too large"); return MCB_EINVAL; }
	MCB_TRY(mcb_copy_streams(ctx));
	// ---- runs of buckets
	static const int max_runs = getenv("MCB_IX_RUNS") ? std::max(1, std::min(16, atoi(getenv("MCB_IX_RUNS")))) : 1;   // tuning knob, see above
	const int C = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)max_runs, n >> 20));
	std::vector<int> cb((size_t)C + 1);
	cb[0] = 0; cb[C] = nb;
	for (int c = 1; c < C; ++c) {
		const uint64_t target = n / C * c;
		int q = (int)(std::lower_bound(h_boff, h_boff + nb + 1, target) - h_boff);
		cb[c] = std::max(cb[c - 1], std::min(q, nb));
	}
	MCB_TRY(ctx->d_scr[0].ensure(n * 16 + 16));
	MCB_TRY(ctx->d_scr[1].ensure(((size_t)nb + 1) * 8));
	MCB_TRY(ctx->d_scr[2].ensure((n / 65 + 8ull * nb + 8) * sizeof(IxSeg)));
	MCB_TRY(ctx->d_scr[3].ensure(((size_t)nb + 2) * 4)); // distinct keys per bucket
	MCB_TRY(ctx->d_scr[4].ensure(n * 8 + 16));           // keys (<= n)
	MCB_TRY(ctx->d_scr[5].ensure((n + 2) * 4));          // kstart
	MCB_TRY(ctx->d_scr[6].ensure(n * 8 + 16));           // postings
	MCB_TRY(ctx->d_scr[7].ensure(((size_t)nb + 2) * 4)); // ub
	MCB_TRY(ctx->d_scr[9].ensure((2 * (size_t)C + 2) * 8));
	MCB_TRY(ctx->h_in2.ensure((2 * (size_t)C + 2) * 8));
	mcb_tuple *dt = ctx->d_scr[0].as<mcb_tuple>();
	uint64_t *d_boff = ctx->d_scr[1].as<uint64_t>();
	uint32_t *d_flag = ctx->d_scr[3].as<uint32_t>();
	unsigned long long *d_chain = ctx->d_scr[9].as<unsigned long long>();
	volatile unsigned long long *h_chain = ctx->h_in2.as<unsigned long long>();
	MCB_TRY(mcb_h2d(ctx, d_boff, h_boff, ((size_t)nb + 1) * 8, 1));
	MCB_CUDA(cudaMemsetAsync(d_chain, 0, (2 * (size_t)C + 2) * 8, ctx->stream));
	MCB_TRY(idx_alloc_host(ctx, ix, n, n, nb));           // the number of distinct keys is not known yet: room for n
	// ---- sort kernel configuration (the same for all runs)
	uint64_t maxb = 0;
	for (int i = 0; i < nb; ++i) maxb = std::max<uint64_t>(maxb, h_boff[i + 1] - h_boff[i]);
	if (getenv("MCB_IX_DEBUG")) fprintf(stderr, "[mcb] idx build: n=%llu max bucket=%llu runs=%d\n", (unsigned long long)n, (unsigned long long)maxb, C);
	// index-space walk: 3 bytes of shared memory per tuple of the largest bucket, up to IX_CAP_MAX; larger buckets take the
	// global-scratch variant of the same kernel
	const uint32_t IX_CAP_MAX = 16384;
	const uint32_t cap = (uint32_t)std::min<uint64_t>(IX_CAP_MAX, std::max<uint64_t>(64, (maxb + 63) & ~63ull));
	const size_t smem3 = ((2048 + IX_SMEM_STK * sizeof(IxSeg) + (size_t)cap * 3 + 15) & ~(size_t)15) * IX2_WARPS;
	const size_t smem3_big = (2048 + IX_SMEM_STK * sizeof(IxSeg) + 15) & ~(size_t)15;
	const bool have_big = maxb > cap;
	MCB_TRY(ctx->d_scr[8].ensure(n * 16 + 16));
	if (smem3 > 48 * 1024) MCB_CUDA(cudaFuncSetAttribute(k_index_sort3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
	if (have_big) { MCB_TRY(ctx->d_scr[10].ensure(n + 16)); MCB_TRY(ctx->d_scr[11].ensure(n * 4 + 16)); }
	std::vector<cudaEvent_t> evH((size_t)C), evC((size_t)C);
	for (int c = 0; c < C; ++c) { cudaEventCreateWithFlags(&evH[c], cudaEventDisableTiming); cudaEventCreateWithFlags(&evC[c], cudaEventDisableTiming); }
	std::vector<IxEvPair> t_h2d, t_d2h;
	auto timed = [&](std::vector<IxEvPair> &v, cudaStream_t st, bool begin) {
		if (!ctx->tm.enabled) return;
		if (begin) { IxEvPair p; cudaEventCreate(&p.a); cudaEventCreate(&p.b); cudaEventRecord(p.a, st); v.push_back(p); }
		else cudaEventRecord(v.back().b, st);
	};
	int rc = MCB_OK;
	static const bool dbg = getenv("MCB_IX_DEBUG") != nullptr;
	auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	const double T0 = dbg ? now_ms() : 0;
	auto flush_keys = [&](int c) -> int {      // keys of run c to the host (their position needs chain[c], chain[c+1])
		MCB_CUDA(cudaEventSynchronize(evC[c]));
		const uint64_t k0 = h_chain[c], k1 = h_chain[c + 1];
		if (k1 > k0) {
			timed(t_d2h, ctx->copy_stream2, true);
			MCB_CUDA(cudaMemcpyAsync(ix->keys + k0, ctx->d_scr[4].as<uint64_t>() + k0, (k1 - k0) * 8, cudaMemcpyDeviceToHost, ctx->copy_stream2));
			MCB_CUDA(cudaMemcpyAsync(ix->kstart + k0, ctx->d_scr[5].as<uint32_t>() + k0, (k1 - k0) * 4, cudaMemcpyDeviceToHost, ctx->copy_stream2));
			timed(t_d2h, ctx->copy_stream2, false);
		}
		return MCB_OK;
	};
	auto run = [&](int c) -> int {
		const int b0 = cb[c], b1 = cb[c + 1];
		const uint64_t t0 = h_boff[b0], t1 = h_boff[b1], nc = t1 - t0;
		if (nc) {
			if (src.flat) {
				timed(t_h2d, ctx->copy_stream, true);
				MCB_TRY(mcb_h2d_on(ctx, ctx->copy_stream, dt + t0, src.flat + t0, nc * 16, 4));
				timed(t_h2d, ctx->copy_stream, false);
			} else {
				const int T = std::max(1, std::min(src.n_threads, (int)(nc / 65536) + 1));
				auto work = [&](int q0, int q1) { for (int i = q0; i < q1; ++i) if (src.cnt[i]) memcpy(src.stage + h_boff[i], src.ptrs[i], src.cnt[i] * 16); };
				if (T == 1) work(b0, b1);
				else {
					std::vector<std::thread> th;
					for (int t = 0; t < T; ++t) th.emplace_back(work, b0 + (int)((int64_t)(b1 - b0) * t / T), b0 + (int)((int64_t)(b1 - b0) * (t + 1) / T));
					for (auto &x : th) x.join();
				}
				timed(t_h2d, ctx->copy_stream, true);
				MCB_CUDA(cudaMemcpyAsync(dt + t0, src.stage + t0, nc * 16, cudaMemcpyHostToDevice, ctx->copy_stream));
				timed(t_h2d, ctx->copy_stream, false);
			}
		}
		MCB_CUDA(cudaEventRecord(evH[c], ctx->copy_stream));
		MCB_CUDA(cudaStreamWaitEvent(ctx->stream, evH[c], 0));
		{
			McbSpan sp(ctx->tm, "idx_build");
			if (nc) {
				MCB_LAUNCH(ctx, "index_sort", k_index_sort3<false>, mcb_grid_for(b1 - b0, IX2_WARPS), IX2_WARPS * 32, smem3, dt, ctx->d_scr[8].as<mcb_tuple>(), d_boff + b0, b1 - b0, cap,
				           ctx->d_scr[2].as<IxSeg>(), (uint8_t*)nullptr, (uint32_t*)nullptr);
				if (have_big) MCB_LAUNCH(ctx, "index_sort_big", k_index_sort3<true>, (unsigned)std::min(b1 - b0, 4 * ctx->sm_count), IX2_WARPS * 32, smem3_big, dt, ctx->d_scr[8].as<mcb_tuple>(),
				                         d_boff + b0, b1 - b0, cap, ctx->d_scr[2].as<IxSeg>(), ctx->d_scr[10].as<uint8_t>(), ctx->d_scr[11].as<uint32_t>());
				MCB_LAUNCH(ctx, "ix_count", k_ix_count, mcb_grid_for(b1 - b0, IXF_WARPS), IXF_WARPS * 32, 0, dt, d_boff + b0, b1 - b0, d_flag + b0);
			} else MCB_CUDA(cudaMemsetAsync(d_flag + b0, 0, (size_t)(b1 - b0) * 4, ctx->stream));
			MCB_LAUNCH(ctx, "ix_scan", k_ix_scan_run, 1, 1024, 0, d_flag, b0, b1, d_chain, c, ctx->d_scr[7].as<uint32_t>(), c == C - 1 ? nb : -1);
			if (nc) MCB_LAUNCH(ctx, "ix_finalize", k_ix_finalize, mcb_grid_for(b1 - b0, IXF_WARPS), IXF_WARPS * 32, 0, dt, d_boff + b0, b1 - b0, ctx->d_scr[7].as<uint32_t>() + b0,
			                   ctx->d_scr[4].as<uint64_t>(), ctx->d_scr[5].as<uint32_t>(), ctx->d_scr[6].as<uint64_t>());
		}
		MCB_CUDA(cudaMemcpyAsync((void*)h_chain, d_chain, (2 * (size_t)C + 2) * 8, cudaMemcpyDeviceToHost, ctx->stream));
		MCB_CUDA(cudaEventRecord(evC[c], ctx->stream));
		MCB_CUDA(cudaStreamWaitEvent(ctx->copy_stream2, evC[c], 0));
		if (nc) {
			timed(t_d2h, ctx->copy_stream2, true);
			MCB_CUDA(cudaMemcpyAsync(ix->post + t0, ctx->d_scr[6].as<uint64_t>() + t0, nc * 8, cudaMemcpyDeviceToHost, ctx->copy_stream2));
			timed(t_d2h, ctx->copy_stream2, false);
		}
		if (c >= 1) MCB_TRY(flush_keys(c - 1));
		return MCB_OK;
	};
	for (int c = 0; c < C && rc == MCB_OK; ++c) { rc = run(c); if (dbg) fprintf(stderr, "[mcb] idx run %d queued at +%.3f ms\n", c, now_ms() - T0); }
	if (rc == MCB_OK) rc = flush_keys(C - 1);
	if (rc == MCB_OK) {
		timed(t_d2h, ctx->copy_stream2, true);
		if (cudaMemcpyAsync(ix->ub, ctx->d_scr[7].p, ((size_t)nb + 1) * 4, cudaMemcpyDeviceToHost, ctx->copy_stream2) != cudaSuccess) { mcb_set_error("index d2h failed: %s", cudaGetErrorString(cudaGetLastError())); rc = MCB_ECUDA; }
		timed(t_d2h, ctx->copy_stream2, false);
	}
	if (dbg) fprintf(stderr, "[mcb] idx all queued at +%.3f ms\n", now_ms() - T0);
	cudaStreamSynchronize(ctx->copy_stream); if (dbg) fprintf(stderr, "[mcb] idx h2d done +%.3f ms\n", now_ms() - T0);
	cudaStreamSynchronize(ctx->stream); if (dbg) fprintf(stderr, "[mcb] idx compute done +%.3f ms\n", now_ms() - T0);
	cudaStreamSynchronize(ctx->copy_stream2); if (dbg) fprintf(stderr, "[mcb] idx d2h done +%.3f ms\n", now_ms() - T0);
	for (int c = 0; c < C; ++c) { cudaEventDestroy(evH[c]); cudaEventDestroy(evC[c]); }
	ix_timer_add(ctx, "h2d", t_h2d); ix_timer_add(ctx, "d2h", t_d2h);
	ctx->tm.collect();
	if (rc != MCB_OK) return rc;
	{ cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { mcb_set_error("index build failed: %s", cudaGetErrorString(e)); return MCB_ECUDA; } }
	const uint64_t U = h_chain[C];
	ix->n_keys = U;
	ix->kstart[U] = (uint32_t)n;
	return MCB_OK;
}

extern "C" int mcb_idx_build(mcb_ctx *ctx, const mcb_tuple *tuples, const uint64_t *bucket_off, mcb_index **out)
{
	if (!ctx || !out || !bucket_off) { mcb_set_error("mcb_idx_build: null argument"); return MCB_EINVAL; }
	MCB_CUDA(cudaSetDevice(ctx->prm.device));
	*out = nullptr;
	const int nb = 1 << ctx->prm.b;
	const uint64_t n = bucket_off[nb];
	if (n && !tuples) { mcb_set_error("mcb_idx_build: null tuples"); return MCB_EINVAL; }
	MCB_TRY(ctx->d_counters.ensure(64 * 8));
	IxSource src; src.flat = tuples;
	const int r = idx_build_pipelined(ctx, n, bucket_off, src, out);
	if (r != MCB_OK) { mcb_idx_destroy(*out); *out = nullptr; }
	return r;
}

extern "C" int mcb_idx_build_scattered(mcb_ctx *ctx, const mcb_tuple *const *ptrs, const uint64_t *cnt, int n_threads, mcb_index **out)
{
	if (!ctx || !out || !ptrs || !cnt) { mcb_set_error("mcb_idx_build_scattered: null argument"); return MCB_EINVAL; }
	MCB_CUDA(cudaSetDevice(ctx->prm.device));
	*out = nullptr;
	const int nb = 1 << ctx->prm.b;
	MCB_TRY(ctx->h_in1.ensure(((size_t)nb + 1) * 8));
	uint64_t *boff = ctx->h_in1.as<uint64_t>();
	uint64_t n = 0;
	for (int i = 0; i < nb; ++i) { boff[i] = n; n += cnt[i]; }
	boff[nb] = n;
	MCB_TRY(ctx->h_in0.ensure(n * 16 + 16));
	mcb_tuple *flat = ctx->h_in0.as<mcb_tuple>();
	MCB_TRY(ctx->d_counters.ensure(64 * 8));
	IxSource src; src.ptrs = ptrs; src.cnt = cnt; src.stage = flat; src.n_threads = n_threads;
	const int r = idx_build_pipelined(ctx, n, boff, src, out);
	if (r != MCB_OK) { mcb_idx_destroy(*out); *out = nullptr; }
	return r;
}

// ---------------------------------------------------------------- device-resident index (the contig merge, mcb_combine.cu)
// Same sort and CSR build as above on tuples that are already on the device, bucket-major (d_tuples, sorted in place); the key,
// offset and posting arrays stay on the device (d_scr[4..7]) for the lookups of the merge kernels.
int mcb_index_build_device(mcb_ctx *ctx, mcb_tuple *d_tuples, uint64_t n, const uint64_t *d_boff, const uint64_t *h_boff, McbDeviceIndex *out)
{
	McbMergeScope metric_scope(ctx->tm, 0);                    // mm_idx_generation is part of the metric even when the merge calls it
	const int b = ctx->prm.b, nb = 1 << b;
	memset(out, 0, sizeof *out);
	out->b = b; out->n_post = n;
	if (n >= 0xFFFFFFFFull) { mcb_set_error("index too large"); return MCB_EINVAL; }
	MCB_TRY(ctx->d_scr[2].ensure((n / 65 + 8ull * nb + 8) * sizeof(IxSeg)));
	MCB_TRY(ctx->d_scr[3].ensure(((size_t)nb + 2) * 4));
	MCB_TRY(ctx->d_scr[4].ensure(n * 8 + 16)); MCB_TRY(ctx->d_scr[5].ensure((n + 2) * 4)); MCB_TRY(ctx->d_scr[6].ensure(n * 8 + 16));
	MCB_TRY(ctx->d_scr[7].ensure(((size_t)nb + 2) * 4)); MCB_TRY(ctx->d_scr[9].ensure(4 * 8));
	uint64_t maxb = 0;
	for (int i = 0; i < nb; ++i) maxb = std::max<uint64_t>(maxb, h_boff[i + 1] - h_boff[i]);
	const uint32_t IX_CAP_MAX = 16384;
	const uint32_t cap = (uint32_t)std::min<uint64_t>(IX_CAP_MAX, std::max<uint64_t>(64, (maxb + 63) & ~63ull));
	const size_t smem3 = ((2048 + IX_SMEM_STK * sizeof(IxSeg) + (size_t)cap * 3 + 15) & ~(size_t)15) * IX2_WARPS;
	const size_t smem3_big = (2048 + IX_SMEM_STK * sizeof(IxSeg) + 15) & ~(size_t)15;
	const bool have_big = maxb > cap;
	unsigned long long *d_chain = ctx->d_scr[9].as<unsigned long long>();
	uint32_t *d_cnt = ctx->d_scr[3].as<uint32_t>(), *d_ub = ctx->d_scr[7].as<uint32_t>();
	McbSpan sp(ctx->tm, "idx_build");
	MCB_CUDA(cudaMemsetAsync(d_chain, 0, 4 * 8, ctx->stream));
	if (n) {
		MCB_TRY(ctx->d_scr[8].ensure(n * 16 + 16));
		if (smem3 > 48 * 1024) MCB_CUDA(cudaFuncSetAttribute(k_index_sort3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
		if (have_big) { MCB_TRY(ctx->d_scr[10].ensure(n + 16)); MCB_TRY(ctx->d_scr[11].ensure(n * 4 + 16)); }
		MCB_LAUNCH(ctx, "index_sort", k_index_sort3<false>, mcb_grid_for(nb, IX2_WARPS), IX2_WARPS * 32, smem3, d_tuples, ctx->d_scr[8].as<mcb_tuple>(), d_boff, nb, cap,
		           ctx->d_scr[2].as<IxSeg>(), (uint8_t*)nullptr, (uint32_t*)nullptr);
		if (have_big) MCB_LAUNCH(ctx, "index_sort_big", k_index_sort3<true>, (unsigned)std::min(nb, 4 * ctx->sm_count), IX2_WARPS * 32, smem3_big, d_tuples, ctx->d_scr[8].as<mcb_tuple>(),
		                         d_boff, nb, cap, ctx->d_scr[2].as<IxSeg>(), ctx->d_scr[10].as<uint8_t>(), ctx->d_scr[11].as<uint32_t>());
		MCB_LAUNCH(ctx, "ix_count", k_ix_count, mcb_grid_for(nb, IXF_WARPS), IXF_WARPS * 32, 0, d_tuples, d_boff, nb, d_cnt);
	} else MCB_CUDA(cudaMemsetAsync(d_cnt, 0, (size_t)nb * 4, ctx->stream));
	MCB_LAUNCH(ctx, "ix_scan", k_ix_scan_run, 1, 1024, 0, d_cnt, 0, nb, d_chain, 0, d_ub, nb);
	if (n) MCB_LAUNCH(ctx, "ix_finalize", k_ix_finalize, mcb_grid_for(nb, IXF_WARPS), IXF_WARPS * 32, 0, d_tuples, d_boff, nb, d_ub,
	                  ctx->d_scr[4].as<uint64_t>(), ctx->d_scr[5].as<uint32_t>(), ctx->d_scr[6].as<uint64_t>());
	MCB_LAUNCH(ctx, "ix_close", k_ix_close, 1, 1, 0, d_ub, nb, ctx->d_scr[5].as<uint32_t>(), (uint32_t)n);
	out->keys = ctx->d_scr[4].as<uint64_t>(); out->kstart = ctx->d_scr[5].as<uint32_t>(); out->post = ctx->d_scr[6].as<uint64_t>(); out->ub = d_ub;
	return MCB_OK;
}
