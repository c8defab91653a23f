// mcb_sort.cu — device primitives: stable LSD radix sort of 16-byte elements and exclusive scans.
//
// The reference sorts each of its 16384 buckets separately on the CPU (radix_sort_128x, ksort.h:108-157, called from
// kthread_bucket.c:391) and then qsorts each group (cmpcluster, kthread_bucket.c:442).  Here ONE stable sort over all
// tuples with the composite key (bucket, minimizer, adjusted position desc, rid) yields every bucket, every group and
// the member order at once; CSR group boundaries fall out of a head-flag scan.
//
// Sort: 8-bit digits, three kernels per pass (per-tile histogram, scan of the digit-major histogram table, stable
// scatter).  HBM-bound: each pass reads the elements twice (histogram + scatter) and writes them once, 48 B/element.
// Tiles are 2048 elements (8 warps x 8 steps x 32 lanes); a warp owns a contiguous run of its tile so that ranks
// computed with __match_any_sync are stable.
#include "mcb_common.cuh"

#define SORT_THREADS 256
#define SORT_WARPS 8
#define SORT_STEPS 8
#define SORT_TILE (SORT_THREADS * SORT_STEPS)

// digit extractors: (a) a bit field of one word of a 16-byte element, (b) a bit field of the bucket hash of a Stage-2
// contig-table entry (low 32 bits of the lt-mer << 32 | position)
struct DigitPair {
	int word, shift; unsigned mask;
	__device__ __forceinline__ unsigned operator()(const ulonglong2 &e) const { unsigned long long v = word ? e.y : e.x; return (unsigned)(v >> shift) & mask; }
};
struct DigitKmer {
	int pbits, shift; unsigned mask, b_lo;
	__device__ __forceinline__ unsigned operator()(const unsigned long long &e) const { return ((mcb_kmer_bucket(e >> MCB_S2_POS_BITS, pbits) - b_lo) >> shift) & mask; }
};

template <class E, class D>
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_hist(const E *__restrict__ in, uint64_t n, D digit, uint32_t *__restrict__ hist, unsigned nblocks)
{
	__shared__ unsigned h[256];
	h[threadIdx.x] = 0;
	__syncthreads();
	uint64_t base = (uint64_t)blockIdx.x * SORT_TILE;
#pragma unroll
	for (int s = 0; s < SORT_STEPS; ++s) {
		uint64_t i = base + (uint64_t)s * SORT_THREADS + threadIdx.x;
		if (i < n) atomicAdd(&h[digit(in[i])], 1u);
	}
	__syncthreads();
	hist[(uint64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// Exclusive scan of every digit's row of the digit-major tile histogram (row d = counts of digit d per tile), one CTA per
// digit; the row totals go to rowsum[256].  Together with a 256-value scan inside the scatter kernel this replaces a
// general multi-level scan (3-5 launches per pass) by one launch.
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_rowscan(uint32_t *__restrict__ hist, unsigned nblocks, uint32_t *__restrict__ rowsum)
{
	__shared__ unsigned wsum[SORT_WARPS];
	__shared__ unsigned carry_s;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t *row = hist + (uint64_t)blockIdx.x * nblocks;
	if (threadIdx.x == 0) carry_s = 0;
	__syncthreads();
	for (unsigned base = 0; base < nblocks; base += SORT_THREADS * 4) {
		unsigned v[4], tot = 0;
#pragma unroll
		for (int j = 0; j < 4; ++j) { unsigned i = base + threadIdx.x * 4 + j; v[j] = i < nblocks ? row[i] : 0u; tot += v[j]; }
		unsigned inc = tot;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
		if (lane == 31) wsum[w] = inc;
		__syncthreads();
		unsigned wb = 0, all = 0;
#pragma unroll
		for (int ww = 0; ww < SORT_WARPS; ++ww) { unsigned t = wsum[ww]; if (ww < w) wb += t; all += t; }
		unsigned run = carry_s + wb + inc - tot;
#pragma unroll
		for (int j = 0; j < 4; ++j) { unsigned i = base + threadIdx.x * 4 + j; if (i < nblocks) row[i] = run; run += v[j]; }
		__syncthreads();
		if (threadIdx.x == 0) carry_s += all;
		__syncthreads();
	}
	if (threadIdx.x == 0) rowsum[blockIdx.x] = carry_s;
}

// Stable scatter of one tile.  Elements are first ranked inside the tile (warp-private digit counters +
// __match_any_sync), parked in shared memory in digit order, and then written out so that consecutive threads write
// consecutive addresses of a digit's run: global stores are whole sectors instead of 32 scattered 8/16-byte pieces.
template <class E, class D>
__global__ void __launch_bounds__(SORT_THREADS, sizeof(E) == 16 ? 4 : 6)      // registers capped so that 4 (16-byte elements) / 6 (8-byte) CTAs fit an SM
k_sort_scatter(const E *__restrict__ in, E *__restrict__ out, uint64_t n, D digit, const uint32_t *__restrict__ hist_scanned, unsigned nblocks,
               const uint32_t *__restrict__ rowsum)
{
	__shared__ unsigned wcnt[SORT_WARPS][256];
	__shared__ unsigned gbase[256];
	__shared__ unsigned wsum[SORT_WARPS], rsum[SORT_WARPS];
	__shared__ E stage[SORT_TILE];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	for (int i = threadIdx.x; i < SORT_WARPS * 256; i += SORT_THREADS) (&wcnt[0][0])[i] = 0;
	__syncthreads();
	const uint64_t tbase = (uint64_t)blockIdx.x * SORT_TILE;
	const uint64_t wbase = tbase + (uint64_t)w * (32 * SORT_STEPS);
	E e[SORT_STEPS];
	unsigned dg[SORT_STEPS];
#pragma unroll
	for (int s = 0; s < SORT_STEPS; ++s) {
		uint64_t i = wbase + s * 32 + lane;
		if (i < n) {
			e[s] = in[i];
			dg[s] = digit(e[s]);
			atomicAdd(&wcnt[w][dg[s]], 1u);
		} else dg[s] = 256u;
	}
	__syncthreads();
	{   // thread d owns digit d: tile-local start of the digit (exclusive scan over the 256 digit totals), per-warp starts, global base
		const unsigned d = threadIdx.x;
		unsigned c[SORT_WARPS], tot = 0;
#pragma unroll
		for (int ww = 0; ww < SORT_WARPS; ++ww) { c[ww] = wcnt[ww][d]; tot += c[ww]; }
		const unsigned rs = rowsum[d];                           // elements with digit d in the whole input
		unsigned inc = tot, rinc = rs;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			unsigned t = __shfl_up_sync(0xFFFFFFFFu, inc, o), u = __shfl_up_sync(0xFFFFFFFFu, rinc, o);
			if (lane >= o) { inc += t; rinc += u; }
		}
		if (lane == 31) { wsum[w] = inc; rsum[w] = rinc; }
		__syncthreads();
		unsigned wb = 0, rb = 0;
#pragma unroll
		for (int ww = 0; ww < SORT_WARPS; ++ww) if (ww < w) { wb += wsum[ww]; rb += rsum[ww]; }
		unsigned run = wb + inc - tot;                           // tile-local start of digit d
		const unsigned dstart = rb + rinc - rs;                  // global start of digit d
		gbase[d] = dstart + hist_scanned[(uint64_t)d * nblocks + blockIdx.x] - run;
#pragma unroll
		for (int ww = 0; ww < SORT_WARPS; ++ww) { wcnt[ww][d] = run; run += c[ww]; }
	}
	__syncthreads();
	const unsigned lt = (1u << lane) - 1u;
#pragma unroll
	for (int s = 0; s < SORT_STEPS; ++s) {
		unsigned peers = __match_any_sync(0xFFFFFFFFu, dg[s]);
		bool valid = dg[s] < 256u;
		unsigned base = 0;
		if (valid) base = wcnt[w][dg[s]];
		__syncwarp();
		if (valid && lane == (__ffs(peers) - 1)) wcnt[w][dg[s]] = base + __popc(peers);
		__syncwarp();
		if (valid) stage[base + __popc(peers & lt)] = e[s];
	}
	__syncthreads();
	const unsigned cnt = (unsigned)min((uint64_t)SORT_TILE, n - tbase);
	for (unsigned i = threadIdx.x; i < cnt; i += SORT_THREADS) {
		const E v = stage[i];
		out[(uint32_t)(gbase[digit(v)] + i)] = v;          // 32-bit wrap-around arithmetic: gbase may be "negative"
	}
}

// ---------------------------------------------------------------- scans
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

template <class T>
__device__ __forceinline__ T block_exclusive_scan(T v, T *total, T *smem /* [SCAN_THREADS/32] */)
{
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	T inc = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { T t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
	if (lane == 31) smem[w] = inc;
	__syncthreads();
	if (w == 0) {
		T s = lane < SCAN_THREADS / 32 ? smem[lane] : (T)0, si = s;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { T t = __shfl_up_sync(0xFFFFFFFFu, si, o); if (lane >= o) si += t; }
		if (lane < SCAN_THREADS / 32) smem[lane] = si - s;       // exclusive warp bases
		if (lane == SCAN_THREADS / 32 - 1) *total = si;
	}
	__syncthreads();
	T r = smem[w] + inc - v;
	return r;
}

template <class T>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const T *__restrict__ d, uint64_t n, T *__restrict__ sums)
{
	__shared__ T sm[SCAN_THREADS / 32];
	uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
	T s = 0;
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) if (base + i < n) s += d[base + i];
#pragma unroll
	for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
	if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
	__syncthreads();
	if (threadIdx.x == 0) { T t = 0; for (int i = 0; i < SCAN_THREADS / 32; ++i) t += sm[i]; sums[blockIdx.x] = t; }
}

// in-place exclusive scan of one tile per block, plus the block's base offset (sums may be null for the single-tile case)
template <class T>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(T *__restrict__ d, uint64_t n, const T *__restrict__ sums, uint64_t *total_out)
{
	__shared__ T sm[SCAN_THREADS / 32];
	__shared__ T tot;
	uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
	T v[SCAN_ITEMS], s = 0;
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = base + i < n ? d[base + i] : (T)0; s += v[i]; }
	T ex = block_exclusive_scan<T>(s, &tot, sm);
	T off = sums ? sums[blockIdx.x] : (T)0;
	T run = ex + off;
#pragma unroll
	for (int i = 0; i < SCAN_ITEMS; ++i) { if (base + i < n) d[base + i] = run; run += v[i]; }
	if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0 && !sums) *total_out = (uint64_t)tot;
}

template <class T>
static int scan_rec(mcb_ctx *ctx, T *d, uint64_t n, uint64_t *d_total, int level)
{
	if (n == 0) {
		if (d_total) MCB_CUDA(cudaMemsetAsync(d_total, 0, 8, ctx->stream));
		return MCB_OK;
	}
	uint64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
	if (nb == 1) {
		MCB_LAUNCH(ctx, "scan_apply", k_scan_apply<T>, 1, SCAN_THREADS, 0, d, n, (const T*)nullptr, d_total);
		return MCB_OK;
	}
	if (level >= 4) { mcb_set_error("scan too deep"); return MCB_EINVAL; }
	MCB_TRY(ctx->d_scan_tmp[level].ensure(nb * sizeof(T)));
	T *sums = ctx->d_scan_tmp[level].as<T>();
	MCB_LAUNCH(ctx, "scan_reduce", k_scan_reduce<T>, (unsigned)nb, SCAN_THREADS, 0, d, n, sums);
	MCB_TRY(scan_rec<T>(ctx, sums, nb, d_total, level + 1));
	MCB_LAUNCH(ctx, "scan_apply", k_scan_apply<T>, (unsigned)nb, SCAN_THREADS, 0, d, n, (const T*)sums, (uint64_t*)nullptr);
	return MCB_OK;
}

int mcb_exclusive_scan_u32(mcb_ctx *ctx, uint32_t *d, uint64_t n, uint64_t *d_total) { return scan_rec<uint32_t>(ctx, d, n, d_total, 0); }
int mcb_exclusive_scan_u64(mcb_ctx *ctx, uint64_t *d, uint64_t n, uint64_t *d_total) { return scan_rec<unsigned long long>(ctx, (unsigned long long*)d, n, d_total, 0); }

// ---------------------------------------------------------------- sort driver
int mcb_add_bit_passes(std::vector<McbSortPass> &v, int word, int lo, int hi)
{
	while (lo < hi) {
		int bits = hi - lo > 8 ? 8 : hi - lo;
		McbSortPass p = { word, lo, bits };
		v.push_back(p);
		lo += bits;
	}
	return (int)v.size();
}

int mcb_radix_sort(mcb_ctx *ctx, ulonglong2 *a, ulonglong2 *b, uint64_t n, const McbSortPass *passes, int n_passes, ulonglong2 **sorted_out)
{
	*sorted_out = a;
	if (n <= 1 || n_passes == 0) return MCB_OK;
	uint64_t nb = (n + SORT_TILE - 1) / SORT_TILE;
	if (nb > 0x7FFFFFFFull) { mcb_set_error("sort input too large"); return MCB_EINVAL; }
	MCB_TRY(ctx->d_sort_hist.ensure((nb + 1) * 256 * sizeof(uint32_t)));
	uint32_t *hist = ctx->d_sort_hist.as<uint32_t>(), *rowsum = hist + nb * 256;
	ulonglong2 *src = a, *dst = b;
	for (int p = 0; p < n_passes; ++p) {
		DigitPair dg = { passes[p].word, passes[p].shift, (1u << passes[p].bits) - 1u };
		MCB_LAUNCH(ctx, "radix_hist", (k_sort_hist<ulonglong2, DigitPair>), (unsigned)nb, SORT_THREADS, 0, src, n, dg, hist, (unsigned)nb);
		MCB_LAUNCH(ctx, "sort_rowscan", k_sort_rowscan, 256, SORT_THREADS, 0, hist, (unsigned)nb, rowsum);
		MCB_LAUNCH(ctx, "radix_scatter", (k_sort_scatter<ulonglong2, DigitPair>), (unsigned)nb, SORT_THREADS, 0, src, dst, n, dg, hist, (unsigned)nb, rowsum);
		ulonglong2 *t = src; src = dst; dst = t;
	}
	*sorted_out = src;
	return MCB_OK;
}

// Stage-2 contig table: order the 8-byte entries by the `pbits`-bit bucket hash of their lt-mer (order inside a bucket is free)
int mcb_radix_sort_kmers(mcb_ctx *ctx, unsigned long long *a, unsigned long long *b, uint64_t n, int pbits, uint32_t b_lo, uint32_t b_hi, unsigned long long **sorted_out)
{
	*sorted_out = a;
	if (n <= 1) return MCB_OK;
	uint64_t nb = (n + SORT_TILE - 1) / SORT_TILE;
	if (nb > 0x7FFFFFFFull) { mcb_set_error("sort input too large"); return MCB_EINVAL; }
	MCB_TRY(ctx->d_sort_hist.ensure((nb + 1) * 256 * sizeof(uint32_t)));
	uint32_t *hist = ctx->d_sort_hist.as<uint32_t>(), *rowsum = hist + nb * 256;
	unsigned long long *src = a, *dst = b;
	const int kbits = mcb_bits_for(b_hi - b_lo > 1 ? b_hi - b_lo - 1 : 1);            // bits of (bucket - b_lo)
	const int n_passes = (kbits + 7) / 8;
	for (int p = 0, lo = 0; p < n_passes; ++p) {
		const int bits = (kbits - lo + (n_passes - p) - 1) / (n_passes - p);      // spread the bits evenly over the passes
		DigitKmer dg = { pbits, lo, (1u << bits) - 1u, b_lo };
		MCB_LAUNCH(ctx, "s2_kmer_sort_hist", (k_sort_hist<unsigned long long, DigitKmer>), (unsigned)nb, SORT_THREADS, 0, src, n, dg, hist, (unsigned)nb);
		MCB_LAUNCH(ctx, "sort_rowscan", k_sort_rowscan, 256, SORT_THREADS, 0, hist, (unsigned)nb, rowsum);
		MCB_LAUNCH(ctx, "s2_kmer_sort_scatter", (k_sort_scatter<unsigned long long, DigitKmer>), (unsigned)nb, SORT_THREADS, 0, src, dst, n, dg, hist, (unsigned)nb, rowsum);
		unsigned long long *t = src; src = dst; dst = t;
		lo += bits;
	}
	*sorted_out = src;
	return MCB_OK;
}

// ---------------------------------------------------------------- sharding: group tuples by owning rank (stable)
// owner of a tuple = owner of its bucket: contiguous bucket ranges, rank = bucket * n_ranks >> 14;
// elements that are not tuples (reads that were not sketched) go last and are not counted
struct DigitOwner {
	int n_ranks;
	__device__ __forceinline__ unsigned operator()(const ulonglong2 &e) const
	{ return e.x == MCB_K1_INVALID ? (unsigned)n_ranks : (unsigned)(((e.x >> 50) * (unsigned long long)n_ranks) >> 14); }
};
int mcb_partition_by_owner(mcb_ctx *ctx, ulonglong2 *a, ulonglong2 *b, uint64_t n, int n_ranks, uint64_t *counts)
{
	for (int i = 0; i <= n_ranks; ++i) counts[i] = 0;
	if (n == 0) return MCB_OK;
	if (n == 1) {    // nothing to move; classify the single element on the host side of a tiny copy
		ulonglong2 e;
		MCB_CUDA(cudaMemcpyAsync(&e, a, 16, cudaMemcpyDeviceToHost, ctx->stream));
		MCB_CUDA(cudaStreamSynchronize(ctx->stream));
		counts[e.x == MCB_K1_INVALID ? n_ranks : (int)(((e.x >> 50) * (unsigned long long)n_ranks) >> 14)] = 1;
		return MCB_OK;
	}
	uint64_t nb = (n + SORT_TILE - 1) / SORT_TILE;
	if (nb > 0x7FFFFFFFull) { mcb_set_error("sort input too large"); return MCB_EINVAL; }
	MCB_TRY(ctx->d_sort_hist.ensure((nb + 1) * 256 * sizeof(uint32_t)));
	uint32_t *hist = ctx->d_sort_hist.as<uint32_t>(), *rowsum = hist + nb * 256;
	DigitOwner dg = { n_ranks };
	MCB_LAUNCH(ctx, "shard_hist", (k_sort_hist<ulonglong2, DigitOwner>), (unsigned)nb, SORT_THREADS, 0, a, n, dg, hist, (unsigned)nb);
	MCB_LAUNCH(ctx, "sort_rowscan", k_sort_rowscan, 256, SORT_THREADS, 0, hist, (unsigned)nb, rowsum);
	MCB_LAUNCH(ctx, "shard_scatter", (k_sort_scatter<ulonglong2, DigitOwner>), (unsigned)nb, SORT_THREADS, 0, a, b, n, dg, hist, (unsigned)nb, rowsum);
	uint32_t sums[256];          // elements per digit = per owner
	MCB_CUDA(cudaMemcpyAsync(sums, rowsum, sizeof sums, cudaMemcpyDeviceToHost, ctx->stream));
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	for (int i = 0; i <= n_ranks; ++i) counts[i] = sums[i];
	return MCB_OK;
}

// ---------------------------------------------------------------- bucket-local tuple sort
// The tuple sort of kt_for_bucket orders by (bucket, minimizer, adjusted position desc, rid): 70+ key bits, nine 8-bit LSD passes
// over all tuples.  Radix passes on the leading key bits alone — the 14 bucket bits plus, for large inputs, the top `e` bits of
// the minimizer, chosen so that a sub-bucket holds a few hundred tuples — bring every sub-bucket together; the rest of the key
// is then sorted inside the sub-bucket, in shared memory, by one CTA (bitonic network on the 128-bit key — keys are unique, rid
// is part of them).  Global traffic drops from 9 x 48 to 2 x 48 + 32 bytes per tuple (3 x 48 above 2^16 sub-buckets).
// Elements that are not tuples (reads that were not sketched: N-rich, poly-A/T) get a digit of their own past the last
// sub-bucket, so they never inflate one.  Sub-buckets above BL_CAP tuples (one minimizer shared by thousands of reads) are
// not sorted here: they raise *overflow and the caller falls back to the full LSD sort.
#define BL_CAP 2048
#define BL_THREADS 256

// sub-bucket of a tuple: (bucket - b_lo) << e | which of 2^e value ranges of x >> 14 the minimizer falls into (K1 = bucket << 50 |
// x >> 14).  A minimizer is the MINIMUM of the read's k-mer hashes, so its leading bits are almost always zero: equal-width
// ranges would put every tuple into the first one.  The ranges are therefore the 2^e quantiles of the minimum of n_k uniform
// hashes (thr[j], ascending, computed on the host); the rank of a value among them is monotone in the value, so sub-bucket
// order is sort order.  A skewed input only unbalances the sub-buckets (-> the LSD fallback), it cannot misorder anything.
struct SubBucket {
	int e;
	unsigned b_lo;           // first bucket this context holds (sharded: a contiguous range of the 16384)
	const unsigned long long *thr;      // [2^e - 1]
	__device__ __forceinline__ unsigned operator()(unsigned long long k1) const
	{
		const unsigned long long xv = k1 & 0x3FFFFFFFFFFFFull;
		unsigned lo = 0, hi = (1u << e) - 1u;                  // number of thresholds <= xv
		while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (thr[mid] <= xv) lo = mid + 1; else hi = mid; }
		return (((unsigned)(k1 >> 50) - b_lo) << e) | lo;
	}
};
struct DigitSub {
	SubBucket sb; int shift; unsigned mask, invalid_digit;     // invalid_digit: digit of the non-tuples (top pass: one past the largest), 0 elsewhere
	__device__ __forceinline__ unsigned operator()(const ulonglong2 &el) const
	{ return el.x == MCB_K1_INVALID ? invalid_digit : (sb(el.x) >> shift) & mask; }
};

__global__ void k_bucket_bounds(const ulonglong2 *__restrict__ e, uint64_t n, SubBucket sb, unsigned nsub, uint32_t *__restrict__ boff)
{
	const unsigned b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b > nsub) return;
	uint64_t lo = 0, hi = n;                  // first element whose sub-bucket is >= b
	while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (sb(e[mid].x) < b) lo = mid + 1; else hi = mid; }
	boff[b] = (uint32_t)lo;
}

__device__ __forceinline__ bool bl_less(const ulonglong2 &a, const ulonglong2 &b) { return a.x < b.x || (a.x == b.x && a.y < b.y); }

__global__ void __launch_bounds__(BL_THREADS)
k_bucket_local_sort(ulonglong2 *__restrict__ e, const uint32_t *__restrict__ boff, unsigned long long *__restrict__ overflow)
{
	__shared__ ulonglong2 s[BL_CAP];
	const uint32_t b0 = boff[blockIdx.x], n = boff[blockIdx.x + 1] - b0;
	if (n <= 1) return;
	if (n > BL_CAP) { if (threadIdx.x == 0) atomicAdd(overflow, 1ull); return; }
	uint32_t P = 2; while (P < n) P <<= 1;
	const ulonglong2 pad = make_ulonglong2(~0ull, ~0ull);
	for (uint32_t i = threadIdx.x; i < P; i += BL_THREADS) s[i] = i < n ? e[b0 + i] : pad;
	__syncthreads();
	for (uint32_t k = 2; k <= P; k <<= 1)
		for (uint32_t j = k >> 1; j > 0; j >>= 1) {
			for (uint32_t t = threadIdx.x; t < (P >> 1); t += BL_THREADS) {
				const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), p = i | j;      // pair (i, i+j)
				const bool up = (i & k) == 0;
				const ulonglong2 x = s[i], y = s[p];
				if (bl_less(y, x) == up) { s[i] = y; s[p] = x; }
			}
			__syncthreads();
		}
	for (uint32_t i = threadIdx.x; i < n; i += BL_THREADS) e[b0 + i] = s[i];
}

// a/b: double buffer; n elements of which n_valid are tuples whose minimizers have `kbits` significant bits; boff: device scratch
// u32[*n_sub_out + 2] (at most n_valid/256 + 16386 entries); overflow: device counter (must be zero on entry).  The sorted tuples
// occupy [0, n_valid) of *sorted_out, the other elements follow.
#include <math.h>
int mcb_bucket_sort(mcb_ctx *ctx, ulonglong2 *a, ulonglong2 *b, uint64_t n, uint64_t n_valid, int kbits, int n_kmers, uint32_t *boff, unsigned long long *overflow, ulonglong2 **sorted_out)
{
	*sorted_out = a;
	if (n <= 1) return MCB_OK;
	const uint64_t nb = (n + SORT_TILE - 1) / SORT_TILE;
	if (nb > 0x7FFFFFFFull) { mcb_set_error("sort input too large"); return MCB_EINVAL; }
	MCB_TRY(ctx->d_sort_hist.ensure((nb + 1) * 256 * sizeof(uint32_t)));
	uint32_t *hist = ctx->d_sort_hist.as<uint32_t>(), *rowsum = hist + nb * 256;
	const int kx = kbits > 14 ? kbits - 14 : 0;                       // bits of x >> 14
	int e = 0;
	// about 256..512 tuples per sub-bucket; a sharded context holds 1/shard_n of the buckets, so its buckets are shard_n times fuller
	while (e < 9 && e < kx && ((n_valid * (uint64_t)ctx->shard_n) >> (14 + e)) > 512) ++e;
	// the buckets of this context: all of them, or the contiguous range a sharded context owns (owner = bucket * G >> 14)
	const unsigned G = (unsigned)ctx->shard_n, rk = (unsigned)ctx->shard_rank;
	const unsigned b_lo = (rk * 16384u + G - 1) / G, b_hi = ((rk + 1) * 16384u + G - 1) / G;
	// quantiles of the minimum of n_kmers uniform kbits-bit hashes, on the scale of x >> 14: F(x) = 1 - (1 - x / 2^kbits)^n_kmers
	MCB_TRY(ctx->d_subthr.ensure(((size_t)1 << e) * 8 + 16));
	if (e > 0) {
		unsigned long long thr[512];
		const long double scale = ldexpl(1.0L, kx), nk = (long double)(n_kmers > 0 ? n_kmers : 1);
		for (int j = 1; j < (1 << e); ++j) {
			const long double q = (long double)j / (long double)(1 << e);
			long double t = scale * (1.0L - powl(1.0L - q, 1.0L / nk));
			if (t < 0) t = 0;
			thr[j - 1] = t >= scale ? (unsigned long long)scale - 1 : (unsigned long long)t;
			if (j > 1 && thr[j - 1] < thr[j - 2]) thr[j - 1] = thr[j - 2];
		}
		MCB_CUDA(cudaMemcpyAsync(ctx->d_subthr.p, thr, ((size_t)(1 << e) - 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
		MCB_CUDA(cudaStreamSynchronize(ctx->stream));             // thr lives on this stack frame
	}
	const SubBucket sb = { e, b_lo, ctx->d_subthr.as<unsigned long long>() };
	const int P = mcb_bits_for(b_hi - b_lo > 1 ? b_hi - b_lo - 1 : 1) + e;
	// digits, least significant first; the top one has at most 7 bits so that the non-tuples fit behind it as digit 128
	int bits[4], np = 0;
	{
		const int top = P < 7 ? P : 7, rest = P - top, nlow = (rest + 7) / 8;
		for (int i = 0, left = rest; i < nlow; ++i) { const int bq = (left + (nlow - i) - 1) / (nlow - i); bits[np++] = bq; left -= bq; }
		bits[np++] = top;
	}
	ulonglong2 *src = a, *dst = b;
	for (int p = 0, lo = 0; p < np; ++p) {
		const DigitSub dg = { sb, lo, (1u << bits[p]) - 1u, p == np - 1 ? 128u : 0u };
		MCB_LAUNCH(ctx, "sort_hist", (k_sort_hist<ulonglong2, DigitSub>), (unsigned)nb, SORT_THREADS, 0, src, n, dg, hist, (unsigned)nb);
		MCB_LAUNCH(ctx, "sort_rowscan", k_sort_rowscan, 256, SORT_THREADS, 0, hist, (unsigned)nb, rowsum);
		MCB_LAUNCH(ctx, "sort_scatter", (k_sort_scatter<ulonglong2, DigitSub>), (unsigned)nb, SORT_THREADS, 0, src, dst, n, dg, hist, (unsigned)nb, rowsum);
		ulonglong2 *t = src; src = dst; dst = t;
		lo += bits[p];
	}
	*sorted_out = src;
	if (n_valid <= 1) return MCB_OK;
	const unsigned nsub = (b_hi - b_lo) << e;
	MCB_LAUNCH(ctx, "bucket_bounds", k_bucket_bounds, (nsub + 1 + 255) / 256, 256, 0, src, n_valid, sb, nsub, boff);
	MCB_LAUNCH(ctx, "bucket_local_sort", k_bucket_local_sort, nsub, BL_THREADS, 0, src, boff, overflow);
	return MCB_OK;
}
