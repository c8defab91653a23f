// mcb_common.cuh — context, buffers, timers and the small integer helpers shared by all kernels.
//
// Data layout in HBM:
//   packed reads   u64[N][WS]   2-bit codes A0 C1 G2 T3 (seq_nt4_table, sketch.c:8-25), base i in word i/32 at bits
//                               2*(i%32); WS = round_up(ceil(L/32),2) so a row is a whole number of 16-byte vectors.
//   sort elements  ulonglong2   .x = K1 = bucket<<50 | x>>14   (bucket = x & 0x3FFF, kthread_reads.c:213)
//                               .y = K2 = posinv<<33 | rid<<1 | strand  (cmpcluster order, kthread_bucket.c:44-62)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <map>
#include "../../include/minicom_b200.h"

#define MCB_HD __host__ __device__ __forceinline__
#define MCB_LH_WMAX 128     // widest contig minimizer window (-w): k_sketch_lh2 keeps a 7-bit slot number per ring entry

// ---------------------------------------------------------------- errors
void mcb_set_error(const char *fmt, ...);
#define MCB_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	mcb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
	return MCB_ECUDA; } } while (0)
#define MCB_TRY(call) do { int r_ = (call); if (r_ != MCB_OK) return r_; } while (0)

// ---------------------------------------------------------------- growable buffers
struct DBuf {            // device
	void *p = nullptr; size_t cap = 0;
	int ensure(size_t bytes) {
		if (bytes <= cap) return MCB_OK;
		if (p) { cudaFree(p); p = nullptr; cap = 0; }
		size_t want = bytes + (bytes >> 3) + 256;
		MCB_CUDA(cudaMalloc(&p, want));
		cap = want; return MCB_OK;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
	template <class T> T *as() const { return (T*)p; }
};
struct HBuf {            // pinned host
	void *p = nullptr; size_t cap = 0;
	int ensure(size_t bytes) {
		if (bytes <= cap) return MCB_OK;
		if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
		size_t want = bytes + (bytes >> 3) + 256;
		MCB_CUDA(cudaMallocHost(&p, want));
		cap = want; return MCB_OK;
	}
	void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
	template <class T> T *as() const { return (T*)p; }
};

// Pool of pinned host slabs that outlive single calls (the arrays of an mcb_index are handed to the caller and read by the
// host for as long as the index lives).  cudaMallocHost costs ~0.2 ms per MiB, so slabs are recycled; the pool is
// reference-counted because an index may be destroyed after its context.
#include <mutex>
struct McbPinnedPool {
	std::mutex mu;
	std::vector<HBuf> free_slabs;
	int refs = 1;               // the context + every live index
	bool ctx_alive = true;
	HBuf take(size_t bytes) {
		std::lock_guard<std::mutex> g(mu);
		int best = -1;
		for (size_t i = 0; i < free_slabs.size(); ++i)
			if (free_slabs[i].cap >= bytes && (best < 0 || free_slabs[i].cap < free_slabs[best].cap)) best = (int)i;
		HBuf b;
		if (best >= 0) { b = free_slabs[best]; free_slabs.erase(free_slabs.begin() + best); }
		++refs;
		return b;                 // caller ensure()s the size (no-op when recycled)
	}
	// returns true when the pool itself must be deleted by the caller
	bool give(HBuf b) {
		std::lock_guard<std::mutex> g(mu);
		if (ctx_alive && free_slabs.size() < 8) free_slabs.push_back(b); else b.release();
		return --refs == 0;
	}
	bool close() {                // context is going away
		std::lock_guard<std::mutex> g(mu);
		ctx_alive = false;
		for (auto &b : free_slabs) b.release();
		free_slabs.clear();
		return --refs == 0;
	}
};

// ---------------------------------------------------------------- timers (CUDA events on ctx->stream)
struct McbTimers {
	struct Rec { int name; cudaEvent_t a, b; };
	bool enabled = false;
	std::vector<std::string> names;
	std::map<std::string, int> idx;
	std::vector<double> ms;
	std::vector<uint64_t> cnt;
	std::vector<Rec> pending;
	std::vector<cudaEvent_t> pool;
	cudaStream_t stream = 0;
	uint64_t launches = 0;
	int id(const char *n) {
		auto it = idx.find(n);
		if (it != idx.end()) return it->second;
		int i = (int)names.size(); names.push_back(n); idx[n] = i; ms.push_back(0); cnt.push_back(0); return i;
	}
	cudaEvent_t ev() {
		if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
		cudaEvent_t e; cudaEventCreate(&e); return e;
	}
	int merge_scope = 0;          // inside mcb_combine: kernel timers are booked as "k:cb_<name>" (the merge is outside the metric)
	int begin(const char *n) {
		if (!enabled) return -1;
		Rec r;
		if (merge_scope && n[0] == 'k' && n[1] == ':' && strncmp(n + 2, "cb_", 3) != 0) { char buf[96]; snprintf(buf, sizeof buf, "k:cb_%s", n + 2); r.name = id(buf); }
		else r.name = id(n);
		r.a = ev(); r.b = ev();
		cudaEventRecord(r.a, stream); pending.push_back(r); return (int)pending.size() - 1;
	}
	void end(int h) { if (h >= 0) cudaEventRecord(pending[h].b, stream); }
	void collect() {  // caller has synchronized the stream
		for (auto &r : pending) {
			float t = 0;
			if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) { cudaGetLastError(); t = 0; }   // span abandoned by an early return
			ms[r.name] += t; cnt[r.name] += 1; pool.push_back(r.a); pool.push_back(r.b);
		}
		pending.clear();
	}
	void reset() { collect(); for (auto &v : ms) v = 0; for (auto &v : cnt) v = 0; launches = 0; }
};

struct McbMergeScope {          // RAII: on / off for the duration of a block
	McbTimers &t; int saved;
	McbMergeScope(McbTimers &t_, int on) : t(t_), saved(t_.merge_scope) { t_.merge_scope = on; }
	~McbMergeScope() { t.merge_scope = saved; }
};
// RAII span for whole-entry-point / copy timers
struct McbSpan {
	McbTimers &t; int h;
	McbSpan(McbTimers &t_, const char *n) : t(t_), h(t_.begin(n)) {}
	~McbSpan() { t.end(h); }
};

#define MCB_LAUNCH(ctx, kname, kernel, grid, block, smem, ...) do { \
	int h_ = (ctx)->tm.begin("k:" kname); \
	kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__); \
	(ctx)->tm.end(h_); (ctx)->tm.launches++; \
	cudaError_t le_ = cudaGetLastError(); if (le_ != cudaSuccess) { \
		mcb_set_error("kernel launch %s failed at %s:%d: %s", kname, __FILE__, __LINE__, cudaGetErrorString(le_)); return MCB_ECUDA; } \
	} while (0)

// slots of the device scalar block ctx->d_counters (u64[64])
enum { CT_SKETCHED = 0, CT_BADCHAR = 1, CT_DEGENERATE = 2, CT_NREADS = 3, CT_REFCURSOR = 4, CT_G = 5, CT_TOT_CL = 6, CT_TOT_MEM = 7,
       CT_TOT_REF = 8, CT_TOT_SG = 9, CT_TOT_RESK = 10, CT_ERR = 11, CT_SCRATCH_IDX = 12,
       CT_S2_U = 16, CT_S2_PROBES = 17, CT_S2_CAND = 18, CT_S2_NEEDEXACT = 19, CT_S2_ERR = 20 /* adjacent to NEEDEXACT: summed over the ranks together */, CT_S2_FPA = 21, CT_S2_FPT = 22, CT_S2_CLAIMS = 23,
       CT_S2_MAXBIN = 24, CT_S2_DIFF = 25, CT_SORT_OVERFLOW = 13, CT_S2_NCAND = 26, CT_S2_NBIGMEM = 27, CT_S2_NEVENTS = 28, CT_WORK0 = 32 /* ..35: consensus work-list sizes */ };

// ---------------------------------------------------------------- context
struct mcb_index;   // defined in mcb_index.cu

// Stage 2: the contigs of the last mcb_realign call and the k-mer index built over them (mcb_stage2.cu).  The contigs do
// not change between the rounds of the -e/-S/-E schedule (preprocess.c:197-232), so the index is built once per run.
struct McbContigIndex {
	bool valid = false;
	uint64_t n_contigs = 0, ref_bytes = 0, total_words = 0, n_windows = 0, n_entries = 0;
	int L = 0, lt = 0, pbits = 0;
	uint32_t b_lo = 0, b_hi = 0;         // table buckets [b_lo, b_hi) of the 2^pbits (the whole table)
	DBuf refs;      // ASCII consensus strings, concatenated (kept to recognise an identical contig set)
	DBuf roff;      // u64[n_contigs+1] offsets into refs
	DBuf cwo;       // u64[n_contigs+1] word offsets of the packed contigs
	DBuf wo;        // u64[n_contigs+1] window offsets (prefix sums of len-L+1)
	DBuf cw;        // packed contigs, 2 bits per base, each contig padded to whole words + 1
	DBuf pblk;      // u32[ref_bytes/512+1] contig holding base 512*i
	DBuf ptab;      // u32[2^pbits+1] bucket ends of the k-mer table
	DBuf ents, ents2; // u64[n_entries] key<<30 | global base position, ordered by a hash of the key (double buffer of the sort)
	unsigned long long *ents_sorted = nullptr;
	DBuf meta;      // S2ContigMeta[n_contigs+1]: {ref_off, cw_off, woff, len} per contig, 32 bytes
	DBuf eoff;      // u64[n_contigs+1] entry offsets (prefix sums of len-lt+1 over contigs with windows)
	// the table holds only the lt-mers that some single of the building call could ask for (Bloom filter over their keys)
	bool table_valid = false;  // contigs packed AND table built
	bool filtered = false;
	int nd = 0, dstart0 = 0;   // dictionary geometry the filter was built for
	DBuf flt;       // u32[flt_words] Bloom filter over the singles' dictionary keys and their reverse complements
	uint64_t flt_words = 0;
	uint64_t n_table = 0;      // entries kept in this context's table
	DBuf sgmap;     // u32[n_reads/32+1] bitmap of the read ids whose keys are in the filter
};

// kt_for_bucket as a resumable round loop (the sharded driver exchanges tuples between the rounds)
struct McbBucketState {
	bool active = false;
	int r = 0, is_last = 0;
	ulonglong2 *cur = nullptr, *alt = nullptr;
	uint64_t n_in = 0, n_valid = 0;                       // elements handed to the sort / valid tuples among them
	uint64_t tot_cl = 0, tot_mem = 0, tot_ref = 0, tot_sg = 0, tot_sk = 0;
	uint64_t n = 0, G = 0, n_cl_new = 0, n_mem_new = 0, n_sg_new = 0, n_ref_new = 0, n_resk = 0;   // current round
	bool half = false;                                    // round_a done, round_b pending
	std::vector<uint64_t> round_cl, round_mem, round_ref, round_sg;   // what each round appended (sharded merge order)
};

// mcb_realign between its two halves (search / claim emission)
struct McbRealignState {
	bool pending = false;
	uint64_t S = 0;
	int nd = 0;
	bool have_index = false;                  // sharded: d_x[1] holds the position of every local single in the job's sg list
};

struct mcb_ctx {
	mcb_params prm;
	cudaStream_t stream = 0, copy_stream = 0, copy_stream2 = 0;   // compute; host->device pipeline; device->host pipeline
	int sm_count = 148;
	McbTimers tm;
	// geometry
	int L = 0, Wd = 0, WS = 0;          // words per read (used / row stride)
	// ---- persistent read state (mcb_for_reads)
	uint64_t n_reads = 0;                // size of the read-id space (all reads of the job; == n_local unless sharded)
	uint64_t n_local = 0, rid_base = 0;  // this context's slice [rid_base, rid_base + n_local)
	int shard_rank = 0, shard_n = 1;
	void *comm = nullptr;                // ncclComm_t of this rank (mcb_shard.cu); null on a single GPU
	bool own_comm = false;
	DBuf d_subthr;                        // sub-bucket thresholds of the tuple sort (mcb_sort.cu)
	DBuf d_rows_send, d_rows_recv, d_coll;   // sharded Stage 1: packed rows travelling with their tuples; small collectives
	HBuf h_coll;
	uint64_t elem_cap = 0;               // capacity (elements) of d_elemA / d_elemB
	bool reads_loaded = false, bucket_done = false;
	bool seed_results_on_device_only = false;   // mcb_for_bucket_keep
	McbBucketState bs;
	McbRealignState rs;
	DBuf d_ascii;                        // N*L bytes (kept until the N masks are extracted)
	DBuf d_packed;                       // u64[N][WS]
	DBuf d_cls;                          // u8[N]
	DBuf d_elemA, d_elemB;               // ulonglong2[N] sort double buffer
	DBuf d_counters;                     // u64[64] device scalars
	uint64_t n_valid_round1 = 0;
	// reads containing N: sorted rid list + N masks (1 bit per base, same 2-bit-field layout: bit 2*(i%32) of word i/32)
	uint64_t n_nreads = 0;
	DBuf d_nread_rid;                    // u32[n_nreads]
	DBuf d_nread_mask;                   // u64[n_nreads][WS]
	// scratch shared by the phases
	DBuf d_scr[12];
	DBuf d_sort_hist, d_scan_tmp[4];
	DBuf d_x[4];                         // stage-2 extras
	DBuf d_out[8];                       // kt_for_bucket outputs accumulated over the rounds
	McbContigIndex cix;                  // stage-2 contig k-mer index (cached across threshold rounds)
	McbPinnedPool *pool = nullptr;       // pinned slabs of the minimizer indexes
	struct McbCombineState *cb = nullptr; // contig merge (mcb_combine.cu), created on first use
	// host result buffers
	HBuf h_cls, h_nrid, h_nrepl, h_noff, h_npos, h_nmask, h_counters, h_stage;
	HBuf h_cl_n, h_cl_a_off, h_cl_a, h_cl_ref_off, h_cl_ref, h_sg, h_mi_cnt, h_mi;
	HBuf h_claim_c, h_claim_s, h_claim_y, h_claim_p, h_fpA, h_fpT, h_in0, h_in1, h_in2;
	std::vector<uint64_t> tmp_off;
};

// ---------------------------------------------------------------- integer helpers (host + device)
// Invertible integer mix restricted to 2k bits (sketch.c:27-37).
// The shift-add forms of the reference are multiplications modulo 2^64 (the mask is applied afterwards in both):
//   ~key + (key<<21) = key*(2^21-1) - 1,  key + (key<<3) + (key<<8) = key*265,  key + (key<<2) + (key<<4) = key*21,
//   key + (key<<31) = key*(2^31+1).  A 64x32-bit multiply is 2-3 instructions on the device, each shift-add chain 6-8.
#ifdef __CUDACC__
// device form for 2k > 32 (mask covers the whole low word): the key lives in two 32-bit halves, every stage is one wide
// multiply-add, one multiply-add, one AND on the high half and a funnel-shift XOR: 24 instructions instead of 34.
__device__ __forceinline__ uint64_t mcb_hash64_wide(uint64_t key, uint32_t mh)
{
	uint32_t lo = (uint32_t)key, hi = (uint32_t)(key >> 32);
	uint64_t t = (uint64_t)lo * 0x1FFFFFu + 0xFFFFFFFFFFFFFFFFull;
	hi = (hi * 0x1FFFFFu + (uint32_t)(t >> 32)) & mh; lo = (uint32_t)t;
	lo ^= __funnelshift_r(lo, hi, 24); hi ^= hi >> 24;
	t = (uint64_t)lo * 265u;
	hi = (hi * 265u + (uint32_t)(t >> 32)) & mh; lo = (uint32_t)t;
	lo ^= __funnelshift_r(lo, hi, 14); hi ^= hi >> 14;
	t = (uint64_t)lo * 21u;
	hi = (hi * 21u + (uint32_t)(t >> 32)) & mh; lo = (uint32_t)t;
	lo ^= __funnelshift_r(lo, hi, 28); hi ^= hi >> 28;
	t = (uint64_t)lo * 0x80000001u;
	hi = (hi * 0x80000001u + (uint32_t)(t >> 32)) & mh; lo = (uint32_t)t;
	return (uint64_t)hi << 32 | lo;
}
#endif
MCB_HD uint64_t mcb_hash64_hd(uint64_t key, uint64_t mask)
{
#ifdef __CUDA_ARCH__
	if ((uint32_t)mask == 0xFFFFFFFFu) return mcb_hash64_wide(key, (uint32_t)(mask >> 32));
#endif
	key = (key * 0x1FFFFFull - 1ull) & mask;
	key ^= key >> 24;
	key = (key * 265ull) & mask;
	key ^= key >> 14;
	key = (key * 21ull) & mask;
	key ^= key >> 28;
	key = (key * 0x80000001ull) & mask;
	return key;
}

// base i (0..L-1) of a packed row
MCB_HD unsigned mcb_base_at(const uint64_t *row, int i) { return (unsigned)(row[i >> 5] >> (2 * (i & 31))) & 3u; }

// ASCII -> code: A0 C1 G2 T3 N4, anything else 5 (rejected; the reference's behaviour on such input is undefined:
// process_reads counts only upper-case ACGTN, kthread_reads.c:56-73, and construct_ref indexes count_table[4] for 'N').
MCB_HD unsigned mcb_code_of(unsigned char c)
{
	// branch-free: (c>>1)&3 maps A,C,T,G to 0,1,2,3 and x^(x>>1) turns that into A0 C1 G2 T3; the guess is accepted iff "ACGT"[code] == c.
	// (A chain of comparisons compiles to a four-way branch that serialises whatever follows in a warp.)
	const unsigned x = (c >> 1) & 3u, code = x ^ (x >> 1);
	const unsigned ok = ((0x54474341u >> (8u * code)) & 0xFFu) == c;
	return ok ? code : (c == 'N' ? 4u : 5u);
}

// reverse the order of the 32 two-bit fields of a word and complement them (A<->T, C<->G == 3-code)
MCB_HD uint64_t mcb_rc_word(uint64_t w)
{
#ifdef __CUDA_ARCH__
	uint64_t r = __brevll(w);
#else
	uint64_t r = w;
	r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
	r = ((r >> 2) & 0x3333333333333333ull) | ((r & 0x3333333333333333ull) << 2);
	r = ((r >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((r & 0x0F0F0F0F0F0F0F0Full) << 4);
	r = ((r >> 8) & 0x00FF00FF00FF00FFull) | ((r & 0x00FF00FF00FF00FFull) << 8);
	r = ((r >> 16) & 0x0000FFFF0000FFFFull) | ((r & 0x0000FFFF0000FFFFull) << 16);
	r = (r >> 32) | (r << 32);
#endif
	r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);  // undo the swap inside each field
	return ~r;
}

// Sort keys.  K1 orders by (bucket, minimizer); K2 by cmpcluster: strand-adjusted position descending, rid ascending.
#define MCB_K1_INVALID 0xFFFFFFFFFFFFFFFFull
MCB_HD uint64_t mcb_make_k1(uint64_t x) { return ((x & 0x3FFFull) << 50) | (x >> 14); }
MCB_HD uint64_t mcb_k1_to_x(uint64_t k1) { return ((k1 & 0x3FFFFFFFFFFFFull) << 14) | (k1 >> 50); }
#define MCB_POSINV_SHIFT 33
#define MCB_POSINV_FIELD 0x1FFull   // 9 bits: posinv = pbase - pos', pbase = L + max_rounds <= 256 + 35; all ones = "not a tuple"
// pos' (kthread_bucket.c:51-56): pos for strand 0, L - pos + k - 2 for strand 1 (k = reads->k, NOT the round's kmer)
MCB_HD int mcb_adj_pos(int pos, int strand, int L, int k) { return strand ? L - pos + k - 2 : pos; }
MCB_HD uint64_t mcb_make_k2(uint32_t rid, int pos, int strand, int L, int k, int pbase)
{
	int pa = mcb_adj_pos(pos, strand, L, k);
	return ((uint64_t)(pbase - pa) << MCB_POSINV_SHIFT) | ((uint64_t)rid << 1) | (uint64_t)strand;
}
MCB_HD uint64_t mcb_make_k2_invalid(uint32_t rid) { return (MCB_POSINV_FIELD << MCB_POSINV_SHIFT) | ((uint64_t)rid << 1); }
MCB_HD int mcb_k2_adjpos(uint64_t k2, int pbase) { return pbase - (int)(k2 >> MCB_POSINV_SHIFT); }
MCB_HD uint32_t mcb_k2_rid(uint64_t k2) { return (uint32_t)(k2 >> 1); }
MCB_HD int mcb_k2_strand(uint64_t k2) { return (int)(k2 & 1); }

// mm_sketch_two (sketch.c:238-289) over a packed, N-free read.  Returns x (UINT64_MAX if no valid k-mer) and
// the last position / strand of the chosen k-mer.
// While the first k-1 bases are shifted in, f == r is impossible (the top field of r is the complement of the newest base and
// the top field of f is still empty; the bottom fields likewise), so the reference's counter l (sketch.c:262-270) is simply
// "bases seen" there; from the first full window on a symmetric k-mer is skipped and every other one is hashed.  With odd k a
// full window is never symmetric (its middle base would be its own complement) and the test disappears.
template <bool WIDE, bool ODD>
MCB_HD uint64_t mcb_sketch_two_packed_t(const uint64_t *row, int L, int k, int *pos_out, int *strand_out)
{
	const uint64_t mask = (1ull << (2 * k)) - 1, shift1 = 2 * (k - 1);
	uint64_t f = 0, r = 0, best = ~0ull;
	int bp = 0, bz = 0;
	for (int w = 0; w * 32 < L; ++w) {
		uint64_t word = row[w];
		int lim = L - w * 32; if (lim > 32) lim = 32;
		for (int j = 0; j < lim; ++j) {
			uint64_t c = word & 3; word >>= 2;
			f = (f << 2 | c) & mask;
			r = (r >> 2) | ((3ull ^ c) << shift1);
			if (w * 32 + j < k - 1) continue;
			if (!ODD && f == r) continue;          // symmetric k-mer: skipped (sketch.c:265)
			const int z = f < r ? 0 : 1;
#ifdef __CUDA_ARCH__
			const uint64_t h = WIDE ? mcb_hash64_wide(z ? r : f, (uint32_t)(mask >> 32)) : mcb_hash64_hd(z ? r : f, mask);
#else
			const uint64_t h = mcb_hash64_hd(z ? r : f, mask);
#endif
			if (h < best) { best = h; bp = w * 32 + j; bz = z; }   // strict <: leftmost on ties (sketch.c:275)
		}
	}
	*pos_out = bp; *strand_out = bz;
	return best;
}
#ifdef __CUDACC__
// Device form for 16 < k < 32: k-mers in 32-bit halves, sixteen bases per 32-bit piece of the packed row, the partial-window
// bases rolled in a loop of their own, and the strand of the winner recomputed once at the end instead of carried along.
template <bool ODD>
__device__ __forceinline__ uint64_t mcb_sketch_two_wide(const uint64_t *row, int L, int k, int *pos_out, int *strand_out)
{
	const uint32_t mh = (1u << (2 * k - 32)) - 1u;
	const int s1 = 2 * (k - 1) - 32;
	const uint32_t t3 = 3u << s1;
	uint32_t flo = 0, fhi = 0, rlo = 0, rhi = 0, blo = ~0u, bhi = ~0u;
	int bp = -1;
#define MCB_S2W_ROLL() { const uint32_t c = x & 3u; x >>= 2; fhi = __funnelshift_l(flo, fhi, 2) & mh; flo = flo * 4u + c; \
                         rlo = __funnelshift_r(rlo, rhi, 2); rhi = (rhi >> 2) | ((c << s1) ^ t3); }
	for (int h = 0; h * 16 < L; ++h) {
		uint32_t x = (uint32_t)(row[h >> 1] >> (32 * (h & 1)));
		const int base = h * 16;
		const int lim = min(16, L - base);
		const int split = max(0, min(lim, k - 1 - base));
		for (int j = 0; j < split; ++j) MCB_S2W_ROLL();
		for (int j = split; j < lim; ++j) {
			MCB_S2W_ROLL();
			if (!ODD && flo == rlo && fhi == rhi) continue;          // symmetric k-mer: skipped (sketch.c:265)
			const bool fwd = ((uint64_t)fhi << 32 | flo) < ((uint64_t)rhi << 32 | rlo);
			const uint64_t hv = mcb_hash64_wide((uint64_t)(fwd ? fhi : rhi) << 32 | (fwd ? flo : rlo), mh);
			const uint32_t hlo = (uint32_t)hv, hhi = (uint32_t)(hv >> 32);
			const bool better = hv < ((uint64_t)bhi << 32 | blo);           // strict <: leftmost on ties (sketch.c:275)
			blo = better ? hlo : blo; bhi = better ? hhi : bhi; bp = better ? base + j : bp;
		}
	}
#undef MCB_S2W_ROLL
	if (bp < 0) { *pos_out = 0; *strand_out = 0; return ~0ull; }
	// strand of the winning window [bp-k+1, bp]: r is the complemented window as packed, f its field reversal
	const int o = 2 * (bp - k + 1), wi = o >> 6, sh = o & 63;
	uint64_t v = row[wi] >> sh;
	if (sh + 2 * k > 64) v |= row[wi + 1] << (64 - sh);
	const uint64_t mask = ((uint64_t)mh << 32) | 0xFFFFFFFFull;
	v &= mask;
	const uint64_t r = ~v & mask;
	uint64_t f = __brevll(v) >> (64 - 2 * k);
	f = ((f >> 1) & 0x5555555555555555ull) | ((f & 0x5555555555555555ull) << 1);
	*pos_out = bp; *strand_out = f < r ? 0 : 1;
	return (uint64_t)bhi << 32 | blo;
}
#endif
// ---- "top-aligned" arithmetic for 16 < k < 32 (device): a 2k-bit k-mer or hash lives in the TOP 2k bits of a 64-bit pair of
// 32-bit halves (value << (64 - 2k)).  Multiplication modulo 2^(2k) is then plain wrap-around modulo 2^64, so the four "& mask"
// after hash64's multiplications (sketch.c:28-36) disappear; x ^= x >> s needs the shifted-in low bits cleared, which folds into
// the XOR (one LOP3: a ^ (b & c)); order is preserved, so minima and the f < r test are those of the plain values.
// Pipes: the loop is bound by the ALU pipe (LOP3 / SHF / ISETP / SEL; one warp instruction per two cycles per SM sub-partition,
// like the FMA pipe), so the high half's x >> s, the reverse k-mer's high half and one funnel shift run as IMAD / IMAD.HI with
// powers of two that arrive as kernel arguments (ptxas cannot turn those back into shifts).  Measured on B200 with
// scripts/microbench/sketch_variants.cu: 1.77 -> 1.43 ms per 10 M reads of 100 bases (profiles/r2_sketch_variants.txt).
struct McbTaMul { uint32_t lm, csh, m24, m14, m28, p30, neg30, sub1lo, sub1hi; int32_t sh; };
static inline McbTaMul mcb_ta_mul(int k)
{
	McbTaMul M;
	memset(&M, 0, sizeof M);
	if (k <= 16 || k >= 32) return M;
	M.sh = 64 - 2 * k; M.lm = ~((1u << M.sh) - 1u); M.csh = 1u << M.sh; M.m24 = 1u << 8; M.m14 = 1u << 18; M.m28 = 1u << 4;
	M.p30 = 1u << 30; M.neg30 = 0u - (1u << 30);
	const uint64_t sub = 0ull - (1ull << M.sh);
	M.sub1lo = (uint32_t)sub; M.sub1hi = (uint32_t)(sub >> 32);
	return M;
}
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t mcb_madhi(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
// hash64 of a top-aligned key, top-aligned result
__device__ __forceinline__ uint64_t mcb_hash64_ta(uint32_t lo, uint32_t hi, const McbTaMul &M)
{
	uint64_t t = (uint64_t)lo * 0x1FFFFFu + ((uint64_t)M.sub1hi << 32 | M.sub1lo);
	hi = hi * 0x1FFFFFu + (uint32_t)(t >> 32); lo = (uint32_t)t;
	{ const uint32_t fl = mcb_madhi(lo, M.m24, hi * M.m24), u = __umulhi(hi, M.m24); lo ^= fl & M.lm; hi ^= u; }
	t = (uint64_t)lo * 265u; hi = hi * 265u + (uint32_t)(t >> 32); lo = (uint32_t)t;
	{ const uint32_t fl = __funnelshift_r(lo, hi, 14), u = __umulhi(hi, M.m14); lo ^= fl & M.lm; hi ^= u; }
	t = (uint64_t)lo * 21u; hi = hi * 21u + (uint32_t)(t >> 32); lo = (uint32_t)t;
	{ const uint32_t fl = __funnelshift_r(lo, hi, 28), u = __umulhi(hi, M.m28); lo ^= fl & M.lm; hi ^= u; }
	t = (uint64_t)lo * 0x80000001u; hi = hi * 0x80000001u + (uint32_t)(t >> 32); lo = (uint32_t)t;
	return (uint64_t)hi << 32 | lo;
}
// one base into the top-aligned forward / reverse k-mer (sketch.c:257-259)
#define MCB_TA_ROLL(c, flo, fhi, rlo, rhi, M) { \
	fhi = __funnelshift_l(flo, fhi, 2); flo = flo * 4u + (c) * (M).csh; \
	rlo = __funnelshift_r(rlo, rhi, 2) & (M).lm; \
	rhi = mcb_madhi(rhi, (M).p30, (c) * (M).neg30 + 0xC0000000u); }
// mm_sketch_two (sketch.c:238-289) over a packed, N-free read, 16 < k < 32: same walk as mcb_sketch_two_wide on top-aligned values
template <bool ODD>
__device__ __forceinline__ uint64_t mcb_sketch_two_ta(const uint64_t *row, int L, int k, const McbTaMul &M, int *pos_out, int *strand_out)
{
	uint32_t flo = 0, fhi = 0, rlo = 0, rhi = 0, blo = ~0u, bhi = ~0u;
	int bp = -1;
	for (int h = 0; h * 16 < L; ++h) {
		uint32_t x = (uint32_t)(row[h >> 1] >> (32 * (h & 1)));
		const int base = h * 16;
		const int lim = min(16, L - base);
		const int split = max(0, min(lim, k - 1 - base));
		for (int j = 0; j < split; ++j) { const uint32_t c = x & 3u; x >>= 2; MCB_TA_ROLL(c, flo, fhi, rlo, rhi, M); }
		for (int j = split; j < lim; ++j) {
			const uint32_t c = x & 3u; x >>= 2;
			MCB_TA_ROLL(c, flo, fhi, rlo, rhi, M);
			if (!ODD && flo == rlo && fhi == rhi) continue;          // symmetric k-mer: skipped (sketch.c:265)
			const bool fwd = ((uint64_t)fhi << 32 | flo) < ((uint64_t)rhi << 32 | rlo);
			const uint64_t hv = mcb_hash64_ta(fwd ? flo : rlo, fwd ? fhi : rhi, M);
			const bool better = hv < ((uint64_t)bhi << 32 | blo);           // strict <: leftmost on ties (sketch.c:275)
			blo = better ? (uint32_t)hv : blo; bhi = better ? (uint32_t)(hv >> 32) : bhi; bp = better ? base + j : bp;
		}
	}
	if (bp < 0) { *pos_out = 0; *strand_out = 0; return ~0ull; }
	// strand of the winning window [bp-k+1, bp]: r is the complemented window as packed, f its field reversal
	const int o = 2 * (bp - k + 1), wi = o >> 6, sh = o & 63;
	uint64_t v = row[wi] >> sh;
	if (sh + 2 * k > 64) v |= row[wi + 1] << (64 - sh);
	const uint64_t mask = (1ull << (2 * k)) - 1;
	v &= mask;
	const uint64_t r = ~v & mask;
	uint64_t f = __brevll(v) >> (64 - 2 * k);
	f = ((f >> 1) & 0x5555555555555555ull) | ((f & 0x5555555555555555ull) << 1);
	*pos_out = bp; *strand_out = f < r ? 0 : 1;
	return ((uint64_t)bhi << 32 | blo) >> M.sh;
}
// the device kernels' entry: top-aligned walk for 16 < k < 32, the generic one otherwise
__device__ __forceinline__ uint64_t mcb_sketch_two_dev(const uint64_t *row, int L, int k, const McbTaMul &M, int *pos_out, int *strand_out)
{
	if (k > 16 && k < 32) return (k & 1) ? mcb_sketch_two_ta<true>(row, L, k, M, pos_out, strand_out) : mcb_sketch_two_ta<false>(row, L, k, M, pos_out, strand_out);
	return mcb_sketch_two_packed_t<false, false>(row, L, k, pos_out, strand_out);
}
#endif
MCB_HD uint64_t mcb_sketch_two_packed(const uint64_t *row, int L, int k, int *pos_out, int *strand_out)
{
#ifdef __CUDA_ARCH__
	if (k > 16 && k < 32) return (k & 1) ? mcb_sketch_two_wide<true>(row, L, k, pos_out, strand_out) : mcb_sketch_two_wide<false>(row, L, k, pos_out, strand_out);
#endif
	return mcb_sketch_two_packed_t<false, false>(row, L, k, pos_out, strand_out);
}

// Stage-2 contig table entries: low 32 bits of the lt-mer << 32 | global base position; buckets by a multiplicative hash of
// those 32 bits.  (lt = 17 gives 34-bit lt-mers: the 17th base is confirmed by k_s2_verify, which has both sequences at hand.)
#define MCB_S2_POS_BITS 32
MCB_HD uint32_t mcb_kmer_bucket(uint64_t key, int pbits) { return (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> (64 - pbits)); }

// ---------------------------------------------------------------- primitives (mcb_sort.cu)
struct McbSortPass { int word; int shift; int bits; };  // word 0 = .x, 1 = .y
int mcb_radix_sort(mcb_ctx *ctx, ulonglong2 *a, ulonglong2 *b, uint64_t n, const McbSortPass *passes, int n_passes,
                   ulonglong2 **sorted_out);
int mcb_radix_sort_kmers(mcb_ctx *ctx, unsigned long long *a, unsigned long long *b, uint64_t n, int pbits, uint32_t b_lo, uint32_t b_hi, unsigned long long **sorted_out);
int mcb_bucket_sort(mcb_ctx *ctx, ulonglong2 *a, ulonglong2 *b, uint64_t n, uint64_t n_valid, int kbits, int n_kmers, uint32_t *boff, unsigned long long *overflow, ulonglong2 **sorted_out);
int mcb_add_bit_passes(std::vector<McbSortPass> &v, int word, int lo, int hi);  // digits covering bits [lo,hi)
int mcb_exclusive_scan_u32(mcb_ctx *ctx, uint32_t *d_data, uint64_t n, uint64_t *d_total /* device u64, may be null */);
int mcb_exclusive_scan_u64(mcb_ctx *ctx, uint64_t *d_data, uint64_t n, uint64_t *d_total);

// the minimizer index as device arrays (mcb_index.cu): keys[U] distinct minimizers, bucket-major ascending; kstart[U+1] posting
// offsets; post[n] y values in the reference's order; ub[2^b+1] first key of every bucket
struct McbDeviceIndex { int b; uint64_t n_post; const uint64_t *keys; const uint32_t *kstart; const uint64_t *post; const uint32_t *ub; };
int mcb_index_build_device(mcb_ctx *ctx, mcb_tuple *d_tuples, uint64_t n, const uint64_t *d_boff, const uint64_t *h_boff, McbDeviceIndex *out);

// host -> device copy that is a true DMA whatever the source: page-locked sources go in one piece, pageable ones are staged
// through pinned chunks filled by n_threads host threads while the previous chunk is in flight (mcb_api.cu)
int mcb_h2d(mcb_ctx *ctx, void *dst, const void *src, size_t bytes, int n_threads);
int mcb_h2d_on(mcb_ctx *ctx, cudaStream_t stream, void *dst, const void *src, size_t bytes, int n_threads);
int mcb_copy_streams(mcb_ctx *ctx);   // creates copy_stream / copy_stream2 on first use

// collectives over the context's communicator, enqueued on ctx->stream (mcb_shard.cu)
int mcb_coll_allreduce_sum_u64(mcb_ctx *ctx, unsigned long long *d_inout, size_t n);
int mcb_coll_allreduce_sum_u32(mcb_ctx *ctx, uint32_t *d_inout, size_t n);
int mcb_coll_allgather_inplace_u64(mcb_ctx *ctx, uint64_t *d_buf, size_t chunk_words);      // rank r contributes d_buf[r*chunk, (r+1)*chunk)
int mcb_coll_allgatherv(mcb_ctx *ctx, const void *d_send, uint64_t bytes, std::vector<unsigned long long> &host_out);   // variable-size all-gather of 8-byte words to the host, rank order

static inline unsigned mcb_grid_for(uint64_t n, unsigned block, unsigned cap = 0x7FFFFFFFu)
{
	uint64_t g = (n + block - 1) / block; if (g < 1) g = 1; if (g > cap) g = cap; return (unsigned)g;
}
static inline int mcb_bits_for(uint64_t n) { int b = 0; while (b < 64 && (n >> b)) ++b; return b ? b : 1; }  // bits to hold values < n .. roughly
