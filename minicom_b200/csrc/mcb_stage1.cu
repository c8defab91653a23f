// mcb_stage1.cu — kt_for_reads + kt_for_bucket on the device.
//
//   K01 k_pack_classify_sketch   process_reads (kthread_reads.c:40-230) + mm_sketch_two (sketch.c:238-289)
//   K2  sort + head flags        radix_sort_128x + grouping + qsort(cmpcluster) (kthread_bucket.c:391-442)
//   K3  k_consensus              construct_ref (kthread_bucket.c:69-377)
//   K4  k_sketch_lh              mm_sketch_lh_ori (sketch.c:116-165), first `first_mininum` tuples per seed contig
//   round loop                   kt_for_bucket (kthread_bucket.c:562-629)
#include "mcb_common.cuh"
#include "mcb_lh.cuh"
#include <algorithm>
#include <thread>

// ---------------------------------------------------------------- classification (kthread_reads.c:84-226)
struct McbCounts { int a, c, g, t, n; };
MCB_HD int mcb_classify(const McbCounts &q, int L, int e, int *repl_code)
{
	*repl_code = -1;
	if (q.a == L) return MCB_CLS_ALLA;
	if (q.t == L) return MCB_CLS_ALLT;
	if (q.n == L) return MCB_CLS_ALLN;
	if (q.t + q.g + q.c + q.n <= e) return MCB_CLS_FPA;
	if (q.a + q.g + q.c + q.n <= e) return MCB_CLS_FPT;
	if (q.a + q.t + q.g + q.c <= e) return MCB_CLS_FPN;
	if (5 * q.n > 2 * L) return MCB_CLS_NFILE;                 // cntN > 0.4*L  (kthread_reads.c:182)
	if (q.n > 0) {                                             // most frequent base, ties in the order A,T,G,C (:185-201)
		int mx = q.a; if (q.t > mx) mx = q.t; if (q.g > mx) mx = q.g; if (q.c > mx) mx = q.c;
		*repl_code = mx == q.a ? 0 : mx == q.t ? 3 : mx == q.g ? 2 : 1;
	}
	return MCB_CLS_SKETCHED;
}


#define RD_THREADS 128
#define HASN_BIT 0x80

__global__ void __launch_bounds__(RD_THREADS)
k_pack_classify_sketch(const uint8_t *__restrict__ ascii, uint64_t n, uint64_t rid_base, int L, int Wd, int WS, int k, McbTaMul TM, int e, int pbase,
                       uint64_t *__restrict__ packed, uint8_t *__restrict__ cls, ulonglong2 *__restrict__ elem,
                       unsigned long long *__restrict__ counters)
{
	extern __shared__ __align__(16) unsigned char smem[];
	uint64_t (*sp)[9] = (uint64_t (*)[9])smem;                                   // packed words per thread
	unsigned char *rows = smem + RD_THREADS * 9 * sizeof(uint64_t);              // ASCII tile
	const uint64_t first = (uint64_t)blockIdx.x * RD_THREADS;
	const int nrow = (int)min((uint64_t)RD_THREADS, n - first);
	const size_t nbytes = (size_t)nrow * L;
	const uint8_t *src = ascii + first * L;
	{   // coalesced 128-bit tile load (tile base is 16-byte aligned: 128*L is a multiple of 16)
		const size_t nvec = nbytes >> 4;
		const uint4 *s4 = (const uint4*)src; uint4 *d4 = (uint4*)rows;
		for (size_t i = threadIdx.x; i < nvec; i += RD_THREADS) d4[i] = s4[i];
		for (size_t i = (nvec << 4) + threadIdx.x; i < nbytes; i += RD_THREADS) rows[i] = src[i];
	}
	__syncthreads();
	const int t = threadIdx.x;
	const uint64_t lid = first + t, rid = rid_base + lid;       // index inside this slice / global read id
	bool sketched = false, bad = false, degenerate = false, hasn = false;
	if (t < nrow) {
		// Four characters per step, SIMD inside a 32-bit register.  (c>>1)&3 maps A,C,T,G (0x41,0x43,0x54,0x47) to 0,1,2,3;
		// x ^ (x>>1) turns that into the reference's codes A0 C1 G2 T3 (seq_nt4_table).  A character is accepted iff it
		// is 'N' or equals "ACGT"[code]; everything else is rejected (see mcb_code_of).
		const uint32_t *rows32 = (const uint32_t*)rows;
		const unsigned rowb = (unsigned)t * (unsigned)L;
		McbCounts q = {0, 0, 0, 0, 0};
		uint64_t nmw[8];
		unsigned badacc = 0;
#pragma unroll
		for (int w = 0; w < 8; ++w) {
			uint64_t pw = 0, nw = 0;
			if (w < Wd) {
				const int lim = min(32, L - w * 32);
#pragma unroll
				for (int j = 0; j < 8; ++j) {
					if (j * 4 < lim) {
						const unsigned a = rowb + (unsigned)(w * 32 + j * 4);
						const unsigned ai = a >> 2, sh = (a & 3u) * 8u;
						unsigned v = __funnelshift_r(rows32[ai], rows32[ai + 1], sh);          // staging buffer is padded by 8 bytes
						const int nb = lim - j * 4;                                            // bytes of this word that belong to the read
						if (nb < 4) v = (v & (0xFFFFFFFFu >> (8 * (4 - nb)))) | (0x41414141u << (8 * nb));   // pad with 'A': code 0, valid
						const unsigned x = (v >> 1) & 0x03030303u;
						unsigned code = x ^ ((x >> 1) & 0x01010101u);
						const unsigned tn = v ^ 0x4E4E4E4Eu;
						const unsigned isn = ~(((tn & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | tn | 0x7F7F7F7Fu);            // 0x80 in every byte that is 'N'
						const unsigned sel = (code & 0x3u) | ((code >> 4) & 0x30u) | ((code >> 8) & 0x300u) | ((code >> 12) & 0x3000u);
						const unsigned df = __byte_perm(0x54474341u, 0u, sel) ^ v;
						badacc |= (((df & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | df) & 0x80808080u & ~isn;
						const unsigned n1 = isn >> 7;
						code &= ~(n1 * 3u);
						unsigned pc = code | (code >> 6); pc = (pc | (pc >> 12)) & 0xFFu;
						unsigned pn = n1 | (n1 >> 6); pn = (pn | (pn >> 12)) & 0xFFu;
						if (nb < 4) pn &= (1u << (2 * nb)) - 1u;
						pw |= (uint64_t)pc << (8 * j); nw |= (uint64_t)pn << (8 * j);
					}
				}
				const uint64_t valid = lim == 32 ? 0x5555555555555555ull : (0x5555555555555555ull & ((1ull << (2 * lim)) - 1ull));
				const uint64_t lo = pw & 0x5555555555555555ull, hi = (pw >> 1) & 0x5555555555555555ull;
				const int nn = __popcll(nw), nt = __popcll(lo & hi), ng = __popcll(hi & ~lo), nc = __popcll(lo & ~hi);
				q.n += nn; q.t += nt; q.g += ng; q.c += nc; q.a += lim - nn - nt - ng - nc;
				(void)valid;
				sp[t][w] = pw;
			}
			nmw[w] = nw;
		}
		bad = badacc != 0;
		int repl;
		int c = mcb_classify(q, L, e, &repl);
		hasn = q.n > 0;
		if (c == MCB_CLS_SKETCHED && repl > 0) {
#pragma unroll
			for (int w = 0; w < 8; ++w) if (w < Wd) sp[t][w] |= nmw[w] * (uint64_t)repl;
		}
		cls[lid] = (uint8_t)(c | (hasn ? HASN_BIT : 0));
		ulonglong2 el; el.x = MCB_K1_INVALID; el.y = mcb_make_k2_invalid((uint32_t)rid);
		if (c == MCB_CLS_SKETCHED && !bad) {
			int pos, z;
			uint64_t x = mcb_sketch_two_dev(sp[t], L, k, TM, &pos, &z);
			if (x == ~0ull) degenerate = true;
			else { el.x = mcb_make_k1(x); el.y = mcb_make_k2((uint32_t)rid, pos, z, L, k, pbase); sketched = true; }
		}
		elem[lid] = el;
	}
	__syncthreads();
	{   // coalesced store of the packed rows of this tile
		uint64_t *dst = packed + (rid_base + first) * WS;
		const int total = nrow * WS;
		for (int i = threadIdx.x; i < total; i += RD_THREADS) { int r = i / WS, w = i - r * WS; dst[i] = w < Wd ? sp[r][w] : 0ull; }
	}
	unsigned m;
	m = __ballot_sync(0xFFFFFFFFu, sketched);   if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counters[CT_SKETCHED], (unsigned long long)__popc(m));
	m = __ballot_sync(0xFFFFFFFFu, bad);        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counters[CT_BADCHAR], (unsigned long long)__popc(m));
	m = __ballot_sync(0xFFFFFFFFu, degenerate); if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counters[CT_DEGENERATE], (unsigned long long)__popc(m));
	m = __ballot_sync(0xFFFFFFFFu, hasn);       if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counters[CT_NREADS], (unsigned long long)__popc(m));
}

// reads with N: flag -> scan -> (rid, mask, replacement) side table
__global__ void k_hasn_flags(const uint8_t *__restrict__ cls, uint64_t n, uint32_t *__restrict__ flag)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) flag[i] = (cls[i] & HASN_BIT) ? 1u : 0u;
}
__global__ void k_hasn_extract(const uint8_t *__restrict__ ascii, uint8_t *__restrict__ cls, uint64_t n, uint64_t rid_base, int L, int Wd, int WS, int e,
                               const uint32_t *__restrict__ pos, uint32_t *__restrict__ nrid, uint64_t *__restrict__ nmask, uint8_t *__restrict__ nrepl)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint8_t c = cls[i];
	if (!(c & HASN_BIT)) return;
	cls[i] = c & ~HASN_BIT;
	uint32_t o = pos[i];
	nrid[o] = (uint32_t)(rid_base + i);
	const uint8_t *row = ascii + i * L;
	McbCounts q = {0, 0, 0, 0, 0};
	for (int w = 0; w < WS; ++w) {
		uint64_t nw = 0;
		if (w < Wd) {
			int lim = L - w * 32; if (lim > 32) lim = 32;
			for (int j = 0; j < lim; ++j) {
				unsigned cc = mcb_code_of(row[w * 32 + j]);
				q.a += cc == 0; q.c += cc == 1; q.g += cc == 2; q.t += cc == 3; q.n += cc == 4;
				if (cc == 4) nw |= 1ull << (2 * j);
			}
		}
		nmask[(uint64_t)o * WS + w] = nw;
	}
	int repl; mcb_classify(q, L, e, &repl);
	nrepl[o] = repl < 0 ? 0 : (uint8_t)"ACGT"[repl];
}

// ---------------------------------------------------------------- kt_for_reads on packed rows (N3: the host parser of
// mcb_fastq.cu already turned the characters into 2-bit codes, N as code 0 + a side table of N masks)
__global__ void k_mark_nreads(const uint32_t *__restrict__ nrid_local, uint64_t nn, uint64_t n, uint32_t *__restrict__ bits, unsigned long long *__restrict__ counters)
{
	uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= nn) return;
	const uint32_t r = nrid_local[j];
	if (r >= n || (j && nrid_local[j - 1] >= r)) { atomicAdd(&counters[CT_BADCHAR], 1ull); return; }      // out of range / not ascending
	atomicOr(&bits[r >> 5], 1u << (r & 31));
}

// A, C, G, T counts of a packed row whose unused bits are zero (N positions, code 0, are taken out of `a` by the caller)
__device__ __forceinline__ bool mcb_count_packed(const uint64_t *row, int L, int Wd, int WS, McbCounts *q)
{
	int nt = 0, ng = 0, nc = 0;
	bool bad = false;
#pragma unroll
	for (int w = 0; w < 8; ++w) {
		if (w < WS) {
			const uint64_t v = row[w];
			if (w < Wd) {
				const uint64_t lo = v & 0x5555555555555555ull, hi = (v >> 1) & 0x5555555555555555ull;
				nt += __popcll(lo & hi); ng += __popcll(hi & ~lo); nc += __popcll(lo & ~hi);
				const int lim = L - w * 32;
				if (lim < 32 && (v >> (2 * lim))) bad = true;
			} else if (v) bad = true;
		}
	}
	q->t = nt; q->g = ng; q->c = nc; q->n = 0; q->a = L - nt - ng - nc;
	return bad;
}

// thread per read without N: classification from popcounts, sketch, sort element.  src == dst when the rows were uploaded in
// place; otherwise (rows resident in the caller's device buffer) the row is also copied into the context's read table
__global__ void __launch_bounds__(RD_THREADS)
k_classify_sketch_packed(const uint64_t *__restrict__ src, uint64_t *__restrict__ dst, uint64_t n, uint64_t lid0, uint64_t rid_base, int L, int Wd, int WS, int k, McbTaMul TM, int e, int pbase,
                         const uint32_t *__restrict__ hasn_bits, uint8_t *__restrict__ cls, ulonglong2 *__restrict__ elem, unsigned long long *__restrict__ counters)
{
	__shared__ uint64_t sp[RD_THREADS][9];
	const int t = threadIdx.x;
	const uint64_t i = (uint64_t)blockIdx.x * RD_THREADS + t;
	bool sketched = false, bad = false, degenerate = false;
	if (i < n) {
		const uint64_t lid = lid0 + i, rid = rid_base + lid;
		const bool hasn = hasn_bits && ((hasn_bits[lid >> 5] >> (lid & 31)) & 1u);
		if (!hasn) {                                              // reads with N: k_nreads_packed
			const uint4 *s4 = (const uint4*)(src + i * WS);
#pragma unroll
			for (int v = 0; v < 4; ++v)
				if (2 * v < WS) {
					const uint4 q4 = s4[v];
					sp[t][2 * v] = (uint64_t)q4.y << 32 | q4.x; sp[t][2 * v + 1] = (uint64_t)q4.w << 32 | q4.z;
					if (dst != src) ((uint4*)(dst + i * WS))[v] = q4;
				}
			McbCounts q;
			bad = mcb_count_packed(sp[t], L, Wd, WS, &q);
			int repl;
			const int c = mcb_classify(q, L, e, &repl);
			cls[lid] = (uint8_t)c;
			ulonglong2 el; el.x = MCB_K1_INVALID; el.y = mcb_make_k2_invalid((uint32_t)rid);
			if (c == MCB_CLS_SKETCHED && !bad) {
				int pos, z;
				const uint64_t x = mcb_sketch_two_dev(sp[t], L, k, TM, &pos, &z);
				if (x == ~0ull) degenerate = true;
				else { el.x = mcb_make_k1(x); el.y = mcb_make_k2((uint32_t)rid, pos, z, L, k, pbase); sketched = true; }
			}
			elem[lid] = el;
		}
	}
	unsigned m;
	m = __ballot_sync(0xFFFFFFFFu, sketched);   if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counters[CT_SKETCHED], (unsigned long long)__popc(m));
	m = __ballot_sync(0xFFFFFFFFu, bad);        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counters[CT_BADCHAR], (unsigned long long)__popc(m));
	m = __ballot_sync(0xFFFFFFFFu, degenerate); if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counters[CT_DEGENERATE], (unsigned long long)__popc(m));
}

// thread per read with N: counts with the mask, classification, N replacement in the read table (kthread_reads.c:185-204), sketch,
// and the context's side table (global read id, mask, replacement character)
__global__ void __launch_bounds__(RD_THREADS)
k_nreads_packed(const uint64_t *__restrict__ src, uint64_t *__restrict__ dst, const uint32_t *__restrict__ nrid_local, const uint64_t *__restrict__ nmask, uint64_t nn, uint64_t n,
                uint64_t rid_base, int L, int Wd, int WS, int k, McbTaMul TM, int e, int pbase, uint32_t *__restrict__ nrid_out, uint8_t *__restrict__ nrepl,
                uint8_t *__restrict__ cls, ulonglong2 *__restrict__ elem, unsigned long long *__restrict__ counters)
{
	__shared__ uint64_t sp[RD_THREADS][9];
	const int t = threadIdx.x;
	const uint64_t j = (uint64_t)blockIdx.x * RD_THREADS + t;
	if (j >= nn) return;
	const uint64_t lid = nrid_local[j];
	if (lid >= n) return;                                        // counted by k_mark_nreads
	const uint64_t rid = rid_base + lid;
	nrid_out[j] = (uint32_t)rid;
	for (int w = 0; w < WS; ++w) sp[t][w] = src[lid * WS + w];
	McbCounts q;
	bool bad = mcb_count_packed(sp[t], L, Wd, WS, &q);
	int nn_ = 0;
	for (int w = 0; w < WS; ++w) {
		const uint64_t m = nmask[j * WS + w];
		const int lim = L - w * 32;
		if ((m & 0xAAAAAAAAAAAAAAAAull) || (lim <= 0 && m) || (lim > 0 && lim < 32 && (m >> (2 * lim))) || (sp[t][w] & (m * 3ull))) bad = true;
		nn_ += __popcll(m);
	}
	if (nn_ == 0) bad = true;                                    // listed as a read with N, but has none
	q.n = nn_; q.a -= nn_;
	int repl;
	const int c = mcb_classify(q, L, e, &repl);
	nrepl[j] = repl < 0 ? 0 : (uint8_t)"ACGT"[repl];
	if (c == MCB_CLS_SKETCHED && repl > 0)
		for (int w = 0; w < Wd; ++w) sp[t][w] |= nmask[j * WS + w] * (uint64_t)repl;
	for (int w = 0; w < WS; ++w) dst[lid * WS + w] = sp[t][w];
	cls[lid] = (uint8_t)c;
	ulonglong2 el; el.x = MCB_K1_INVALID; el.y = mcb_make_k2_invalid((uint32_t)rid);
	if (bad) atomicAdd(&counters[CT_BADCHAR], 1ull);
	else if (c == MCB_CLS_SKETCHED) {
		int pos, z;
		const uint64_t x = mcb_sketch_two_dev(sp[t], L, k, TM, &pos, &z);
		if (x == ~0ull) atomicAdd(&counters[CT_DEGENERATE], 1ull);
		else { el.x = mcb_make_k1(x); el.y = mcb_make_k2((uint32_t)rid, pos, z, L, k, pbase); atomicAdd(&counters[CT_SKETCHED], 1ull); }
	}
	elem[lid] = el;
	atomicAdd(&counters[CT_NREADS], 1ull);
}

// batched mm_sketch_two over already packed reads (rounds >= 2: kthread_bucket.c:205,489)
__global__ void k_resketch(const uint64_t *__restrict__ packed, int WS, int L, int k_orig, int kmer, McbTaMul TM, int pbase,
                           const uint32_t *__restrict__ rids, uint64_t n, ulonglong2 *__restrict__ elem, unsigned long long *__restrict__ counters)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t rid = rids[i];
	uint64_t row[8];
#pragma unroll
	for (int w = 0; w < 8; ++w) row[w] = w < WS ? packed[(uint64_t)rid * WS + w] : 0ull;
	int pos, z;
	uint64_t x = mcb_sketch_two_dev(row, L, kmer, TM, &pos, &z);
	ulonglong2 el;
	if (x == ~0ull) { atomicAdd(&counters[CT_DEGENERATE], 1ull); el.x = MCB_K1_INVALID; el.y = mcb_make_k2_invalid(rid); }
	else { el.x = mcb_make_k1(x); el.y = mcb_make_k2(rid, pos, z, L, k_orig, pbase); }
	elem[i] = el;
}

// raw tuples for tests: elem (rid order) -> (x,y)
__global__ void k_elem_to_tuple(const ulonglong2 *__restrict__ elem, uint64_t n, int L, int k, int pbase, mcb_tuple *__restrict__ out, int by_rid, uint64_t rid_base)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	ulonglong2 e = elem[i];
	mcb_tuple t;
	if (e.x == MCB_K1_INVALID) { t.x = ~0ull; t.y = ~0ull; }
	else {
		int z = mcb_k2_strand(e.y), pa = mcb_k2_adjpos(e.y, pbase);
		int pos = z ? L + k - 2 - pa : pa;
		t.x = mcb_k1_to_x(e.x); t.y = (uint64_t)mcb_k2_rid(e.y) << 32 | (uint64_t)pos << 1 | (uint64_t)z;
	}
	out[by_rid ? (uint64_t)mcb_k2_rid(e.y) - rid_base : i] = t;
}

__global__ void k_unpack(const uint64_t *__restrict__ packed, uint64_t n, int L, int WS, char *__restrict__ out)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n * (uint64_t)L) return;
	uint64_t r = i / L; int p = (int)(i - r * L);
	out[i] = "ACGT"[mcb_base_at(packed + r * WS, p)];
}

// ---------------------------------------------------------------- grouping
__global__ void k_heads(const ulonglong2 *__restrict__ e, uint64_t n, uint32_t *__restrict__ flag)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) flag[i] = (i == 0 || e[i].x != e[i - 1].x) ? 1u : 0u;
}
__global__ void k_gstart(const ulonglong2 *__restrict__ e, uint64_t n, const uint32_t *__restrict__ hscan, const unsigned long long *__restrict__ G,
                         uint32_t *__restrict__ gstart)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if (i == 0 || e[i].x != e[i - 1].x) gstart[hscan[i]] = (uint32_t)i;
	if (i == n - 1) gstart[*G] = (uint32_t)n;
}

// ---------------------------------------------------------------- K3 consensus (construct_ref, kthread_bucket.c:69-377)
// One warp per group of >= 2 tuples.  Lane owns consensus columns lane, lane+32, ...; members are streamed in batches.
#define CS_WARPS 4
#define CS_MB 4          // members fetched per batch (CS_MB * 8 word slots = 32 lanes)

struct ConsOut {
	uint8_t *status;      // per element: 0 singleton, 1 kept member, 2 rejected, 3 lone survivor
	uint32_t *erank;      // per element: rank inside its output list
	uint64_t *newrec;     // per element: rid<<32 | offset<<1 | dir (kept members)
	uint32_t *g_iscl, *g_kept, *g_sg, *g_resk;   // per group
	unsigned long long *g_reflen;                // per group
	unsigned long long *g_refoff;                // per group: offset into reftmp
	char *reftmp;
};

__device__ __forceinline__ unsigned member_base(const uint64_t *mw, int dir, int p, int L)
{
	return dir ? 3u - mcb_base_at(mw, L - 1 - p) : mcb_base_at(mw, p);
}

__global__ void __launch_bounds__(CS_WARPS * 32)
k_consensus(const ulonglong2 *__restrict__ el, const uint32_t *__restrict__ gstart, uint64_t G, const uint64_t *__restrict__ packed,
            int WS, int L, int e_thr, int pbase, int NC, int last_round, ConsOut o, unsigned long long *__restrict__ counters, uint64_t reftmp_cap,
            const uint32_t *__restrict__ worklist, const unsigned long long *__restrict__ n_work_ptr)
{
	extern __shared__ __align__(16) unsigned char smem[];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const size_t per_warp = (size_t)NC * 16 + NC + CS_MB * 8 * 8 + CS_MB * 16;
	unsigned char *wsm = smem + (size_t)wib * ((per_warp + 15) & ~(size_t)15);
	uint32_t *cnt = (uint32_t*)wsm;                                   // [4][NC]
	uint64_t *mw = (uint64_t*)(wsm + (size_t)NC * 16);                // [CS_MB][8]
	int *moff = (int*)(mw + CS_MB * 8);                               // [CS_MB]
	int *mdir = moff + CS_MB;                                         // [CS_MB]
	uint32_t *mrid = (uint32_t*)(mdir + CS_MB);                       // [CS_MB]
	uint8_t *msel = (uint8_t*)(mrid + CS_MB);                         // [CS_MB] (padding inside the 16 B/member budget)
	unsigned char *cons = wsm + (size_t)NC * 16 + CS_MB * 8 * 8 + CS_MB * 16;   // [NC]
	const uint64_t nwarps = (uint64_t)gridDim.x * CS_WARPS;
	const uint64_t n_work = *n_work_ptr;                  // only the groups too large for k_consensus_bs come here
	for (uint64_t wi = (uint64_t)blockIdx.x * CS_WARPS + wib; wi < n_work; wi += nwarps) {
		const uint64_t g = worklist[wi];
		const uint32_t s = gstart[g], cntm = gstart[g + 1] - s;
		if (cntm < 2) {
			if (lane == 0) {
				o.status[s] = 0; o.erank[s] = 0;
				o.g_iscl[g] = 0; o.g_kept[g] = 0; o.g_sg[g] = 1; o.g_resk[g] = 0; o.g_reflen[g] = 0; o.g_refoff[g] = 0;
			}
			continue;
		}
		const int posinv0 = (int)(el[s].y >> MCB_POSINV_SHIFT);
		const int off_last = (int)(el[s + cntm - 1].y >> MCB_POSINV_SHIFT) - posinv0;
		int ncol = off_last + L;
		if (ncol > NC) { if (lane == 0) atomicAdd(&counters[CT_ERR], 1ull); ncol = NC; }
		// ---- pass 1: column counts over all members, consensus (ties -> lowest code), stop at first empty column
		for (int c = lane; c < 4 * NC; c += 32) cnt[c] = 0;
		__syncwarp();
		int kept = 0;
		for (int pass = 0; pass < 3; ++pass) {
			// pass 0: count all; pass 1: per-member mismatches vs consensus; pass 2: recount kept members
			if (pass == 2) {
				if (kept < 2) break;
				for (int c = lane; c < 4 * NC; c += 32) cnt[c] = 0;
				__syncwarp();
			}
			for (uint32_t b0 = 0; b0 < cntm; b0 += CS_MB) {
				const int mb = min((uint32_t)CS_MB, cntm - b0);
				{   // fetch a batch: lane = member*8 + word
					int m = lane >> 3, w = lane & 7;
					if (m < mb) {
						unsigned long long k2 = el[s + b0 + m].y;
						uint32_t rid = mcb_k2_rid(k2);
						mw[m * 8 + w] = w < WS ? packed[(uint64_t)rid * WS + w] : 0ull;
						if (w == 0) {
							moff[m] = (int)(k2 >> MCB_POSINV_SHIFT) - posinv0; mdir[m] = (int)(k2 & 1); mrid[m] = rid;
							msel[m] = pass == 2 ? (o.status[s + b0 + m] == 1) : 1;
						}
					}
				}
				__syncwarp();
				for (int m = 0; m < mb; ++m) {
					const int off = moff[m], dir = mdir[m];
					const uint64_t *w64 = mw + m * 8;
					if (pass != 1) {
						if (msel[m]) {
							for (int c = lane; c < ncol; c += 32) {
								int p = c - off;
								if ((unsigned)p < (unsigned)L) cnt[member_base(w64, dir, p, L) * NC + c]++;
							}
						}
					} else {
						int mism = 0;
						for (int p = lane; p < L; p += 32) {
							int c = off + p;
							unsigned char rc = c < NC ? cons[c] : 0xFF;
							mism += rc != member_base(w64, dir, p, L);
						}
						mism = __reduce_add_sync(0xFFFFFFFFu, mism);
						const bool keep = mism <= e_thr;
						kept += keep;
						if (lane == 0) o.status[s + b0 + m] = keep ? 1 : 2;
					}
				}
				__syncwarp();
			}
			if (pass == 0) {
				// consensus over [0, ncol); a column with no coverage would end the string (kthread_bucket.c:117-120)
				int reflen1 = ncol;
				for (int c0 = 0; c0 < ncol; c0 += 32) {
					int c = c0 + lane; bool empty = false;
					if (c < ncol) {
						uint32_t m0 = cnt[c], best = 0;
						for (int b = 1; b < 4; ++b) { uint32_t v = cnt[b * NC + c]; if (v > m0) { m0 = v; best = b; } }
						cons[c] = (unsigned char)best; empty = m0 == 0;
					}
					unsigned em = __ballot_sync(0xFFFFFFFFu, empty);
					if (em) { reflen1 = c0 + __ffs(em) - 1; break; }
				}
				for (int c = reflen1 + lane; c < NC; c += 32) cons[c] = 0xFF;   // beyond the end nothing matches
				__syncwarp();
			}
		}
		// ---- finalize
		const bool iscl = kept >= 2;
		int sv = 0, rend = 0;
		unsigned long long refoff = 0;
		if (iscl) {
			// sv = first covered column, rend = end of the right-most kept member (kthread_bucket.c:285-317)
			int first_cov = ncol;
			for (int c0 = 0; c0 < ncol; c0 += 32) {
				int c = c0 + lane; bool cov = false;
				if (c < ncol) cov = (cnt[c] | cnt[NC + c] | cnt[2 * NC + c] | cnt[3 * NC + c]) != 0;
				unsigned cm = __ballot_sync(0xFFFFFFFFu, cov);
				if (cm) { first_cov = c0 + __ffs(cm) - 1; break; }
			}
			sv = first_cov;
			int last_cov = 0;
			for (int c0 = ((ncol - 1) / 32) * 32; c0 >= 0; c0 -= 32) {
				int c = c0 + lane; bool cov = false;
				if (c < ncol) cov = (cnt[c] | cnt[NC + c] | cnt[2 * NC + c] | cnt[3 * NC + c]) != 0;
				unsigned cm = __ballot_sync(0xFFFFFFFFu, cov);
				if (cm) { last_cov = c0 + 31 - __clz(cm); break; }
			}
			rend = last_cov + 1;    // == max(off)+L over kept members: a member covers every column of its span
			const int reflen2 = rend - sv;
			if (lane == 0) refoff = atomicAdd(&counters[CT_REFCURSOR], (unsigned long long)reflen2);
			refoff = __shfl_sync(0xFFFFFFFFu, refoff, 0);
			if (refoff + reflen2 > reftmp_cap) { if (lane == 0) atomicAdd(&counters[CT_ERR], 1ull); }
			else for (int c = sv + lane; c < rend; c += 32) {
				uint32_t m0 = cnt[c], best = 0;
				for (int b = 1; b < 4; ++b) { uint32_t v = cnt[b * NC + c]; if (v > m0) { m0 = v; best = b; } }
				o.reftmp[refoff + (c - sv)] = "ACGT"[best];
			}
			if (lane == 0) { o.g_reflen[g] = (unsigned long long)reflen2; o.g_refoff[g] = refoff; }
		} else if (lane == 0) { o.g_reflen[g] = 0; o.g_refoff[g] = 0; }
		// per-member ranks and records
		int rk_keep = 0, rk_rej = 0;
		const int nrej = (int)cntm - kept;
		for (uint32_t j0 = 0; j0 < cntm; j0 += 32) {
			uint32_t j = j0 + lane; bool valid = j < cntm; bool keep = false;
			unsigned long long k2 = 0;
			if (valid) { keep = o.status[s + j] == 1; k2 = el[s + j].y; }
			unsigned km = __ballot_sync(0xFFFFFFFFu, valid && keep), rm = __ballot_sync(0xFFFFFFFFu, valid && !keep);
			unsigned lt = (1u << lane) - 1u;
			if (valid) {
				if (keep) {
					if (iscl) {
						int off = (int)(k2 >> MCB_POSINV_SHIFT) - posinv0;
						o.erank[s + j] = rk_keep + __popc(km & lt);
						o.newrec[s + j] = ((unsigned long long)mcb_k2_rid(k2) << 32) | ((unsigned long long)(off - sv) << 1) | (k2 & 1);
					} else { o.status[s + j] = 3; o.erank[s + j] = nrej; }     // lone survivor: after the rejects (:476-498)
				} else o.erank[s + j] = rk_rej + __popc(rm & lt);
			}
			rk_keep += __popc(km); rk_rej += __popc(rm);
		}
		if (lane == 0) {
			const uint32_t nout = iscl ? nrej : cntm;
			o.g_iscl[g] = iscl; o.g_kept[g] = iscl ? kept : 0;
			o.g_sg[g] = last_round ? nout : 0; o.g_resk[g] = last_round ? 0 : nout;
		}
		__syncwarp();
	}
}

// ---------------------------------------------------------------- K3, bit-sliced formulation
// The same construct_ref, reorganised around 32-column words.  A group is handled by GW lanes (GW = 8, 16 or 32, the
// smallest that covers the widest possible pile-up of the round); lane j owns columns [32j, 32j+32).  Per column and
// base, the coverage counters are bit-sliced: plane q of a counter word holds bit q of the 32 columns' counts, so adding
// one member is a ripple-carry over K planes (K = bits needed for the group size) on whole words.  A member's bases are
// split into a low-bit and a high-bit plane (32 bases per word), reverse-complemented at plane level when on the
// reverse strand, and shifted to the member's offset with two shuffles per plane.  Majority, first empty column,
// mismatch counts (XOR against the consensus planes + popcount) and the final trim all work on these words; rejected
// members are subtracted from the counters as soon as they are found, which is pass 3 of the reference (recount over
// the kept members) without touching the kept ones again.  Several groups share a warp (32/GW), walking their members
// in lockstep.
__device__ __forceinline__ uint32_t compress_even(uint64_t w)
{
	uint64_t x = w & 0x5555555555555555ull;
	x = (x | (x >> 1)) & 0x3333333333333333ull;
	x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0Full;
	x = (x | (x >> 4)) & 0x00FF00FF00FF00FFull;
	x = (x | (x >> 8)) & 0x0000FFFF0000FFFFull;
	x = x | (x >> 16);
	return (uint32_t)x;
}
template <int K> __device__ __forceinline__ void bs_add(uint32_t (&c)[K], uint32_t m)
{
	uint32_t carry = m;
#pragma unroll
	for (int q = 0; q < K; ++q) { const uint32_t t = c[q] & carry; c[q] ^= carry; carry = t; }
}
template <int K> __device__ __forceinline__ void bs_sub(uint32_t (&c)[K], uint32_t m)
{
	uint32_t borrow = m;
#pragma unroll
	for (int q = 0; q < K; ++q) { const uint32_t t = ~c[q] & borrow; c[q] ^= borrow; borrow = t; }
}
template <int K> __device__ __forceinline__ uint32_t bs_gt(const uint32_t (&x)[K], const uint32_t (&y)[K])
{
	uint32_t gt = 0, eq = ~0u;
#pragma unroll
	for (int q = K - 1; q >= 0; --q) { gt |= eq & x[q] & ~y[q]; eq &= ~(x[q] ^ y[q]); }
	return gt;
}
// per column: the most frequent base, ties to the lowest code (strict > at kthread_bucket.c:111); cov = some base seen
template <int K> __device__ __forceinline__ void bs_majority(const uint32_t (&cA)[K], const uint32_t (&cC)[K], const uint32_t (&cG)[K], const uint32_t (&cT)[K],
                                                            uint32_t &lo, uint32_t &hi, uint32_t &cov)
{
	uint32_t best[K];
#pragma unroll
	for (int q = 0; q < K; ++q) best[q] = cA[q];
	lo = 0; hi = 0;
	uint32_t g = bs_gt<K>(cC, best);
	lo = g;
#pragma unroll
	for (int q = 0; q < K; ++q) best[q] = (cC[q] & g) | (best[q] & ~g);
	g = bs_gt<K>(cG, best);
	lo &= ~g; hi = g;
#pragma unroll
	for (int q = 0; q < K; ++q) best[q] = (cG[q] & g) | (best[q] & ~g);
	g = bs_gt<K>(cT, best);
	lo |= g; hi |= g;
	cov = 0;
#pragma unroll
	for (int q = 0; q < K; ++q) cov |= (cT[q] & g) | (best[q] & ~g);
}

struct BsMember { uint32_t lo, hi, valid; };
// 32-column chunk (word wq) of member `k2` of the sub-group, oriented and shifted to its offset
template <int GW>
__device__ __forceinline__ BsMember bs_member_chunk(const uint64_t *__restrict__ packed, int WS, int Wd, int L, unsigned long long k2, int posinv0, int wq, unsigned sgmask, bool active)
{
	const int dir = (int)(k2 & 1), off = (int)(k2 >> MCB_POSINV_SHIFT) - posinv0;
	uint32_t plo = 0, phi = 0;
	if (active && wq < Wd) {       // this lane holds plane word wq of the oriented read
		const uint64_t *row = packed + (uint64_t)mcb_k2_rid(k2) * WS;
		uint64_t w;
		if (!dir) w = row[wq];
		else {
			const int pad = Wd * 32 - L;
			const uint64_t a = mcb_rc_word(row[Wd - 1 - wq]);
			const uint64_t b = wq + 1 < Wd ? mcb_rc_word(row[Wd - 2 - wq]) : 0ull;
			w = pad ? (a >> (2 * pad)) | (b << (64 - 2 * pad)) : a;
		}
		if (wq == Wd - 1 && (L & 31)) w &= (1ull << (2 * (L & 31))) - 1;
		plo = compress_even(w); phi = compress_even(w >> 1);
	}
	// columns [32wq, 32wq+32) are read positions [32wq-off, ...): plane words i0, i0+1 shifted by sh
	const int b0 = 32 * wq - off, i0 = b0 >> 5, sh = b0 & 31;
	const int base = (threadIdx.x & 31) & ~(GW - 1);
	const int s0 = (i0 >= 0 && i0 < GW) ? i0 : 0, s1 = (i0 + 1 >= 0 && i0 + 1 < GW) ? i0 + 1 : 0;
	uint32_t l0 = __shfl_sync(sgmask, plo, base + s0), l1 = __shfl_sync(sgmask, plo, base + s1);
	uint32_t h0 = __shfl_sync(sgmask, phi, base + s0), h1 = __shfl_sync(sgmask, phi, base + s1);
	if (i0 < 0 || i0 >= GW) { l0 = 0; h0 = 0; }
	if (i0 + 1 < 0 || i0 + 1 >= GW) { l1 = 0; h1 = 0; }
	BsMember m;
	m.lo = sh ? (l0 >> sh) | (l1 << (32 - sh)) : l0;
	m.hi = sh ? (h0 >> sh) | (h1 << (32 - sh)) : h0;
	const int vb0 = max(0, off - 32 * wq), vb1 = min(32, off + L - 32 * wq);     // valid bits: columns off .. off+L-1
	m.valid = (active && vb1 > vb0) ? ((vb1 >= 32 ? ~0u : (1u << vb1) - 1u) & ~((1u << vb0) - 1u)) : 0u;
	m.lo &= m.valid; m.hi &= m.valid;
	return m;
}

template <int GW> __device__ __forceinline__ int sg_min(int v, unsigned m) { for (int o = GW / 2; o; o >>= 1) v = min(v, __shfl_xor_sync(m, v, o)); return v; }
template <int GW> __device__ __forceinline__ int sg_max(int v, unsigned m) { for (int o = GW / 2; o; o >>= 1) v = max(v, __shfl_xor_sync(m, v, o)); return v; }
template <int GW> __device__ __forceinline__ int sg_sum(int v, unsigned m) { for (int o = GW / 2; o; o >>= 1) v += __shfl_xor_sync(m, v, o); return v; }

// one batch of 32/GW groups (one per sub-group) whose sizes all fit K counter planes
template <int GW, int K>
__device__ void cons_groups_bs(const ulonglong2 *__restrict__ el, uint32_t s, uint32_t cntm, uint64_t g, bool have, const uint64_t *__restrict__ packed,
                               int WS, int Wd, int L, int e_thr, int last_round, const ConsOut &o, unsigned long long *__restrict__ counters, uint64_t reftmp_cap)
{
	const unsigned FULL = 0xFFFFFFFFu;
	const int lane = threadIdx.x & 31, wq = lane & (GW - 1);
	const bool lead = wq == 0;
	uint32_t cA[K], cC[K], cG[K], cT[K];
#pragma unroll
	for (int q = 0; q < K; ++q) { cA[q] = 0; cC[q] = 0; cG[q] = 0; cT[q] = 0; }
	const int posinv0 = have ? (int)(el[s].y >> MCB_POSINV_SHIFT) : 0;
	const int ncol = have ? (int)(el[s + cntm - 1].y >> MCB_POSINV_SHIFT) - posinv0 + L : 0;
	if (have && ncol > 32 * GW && lead) atomicAdd(&counters[CT_ERR], 1ull);
	const uint32_t colmask = ncol >= 32 * (wq + 1) ? ~0u : (ncol > 32 * wq ? (1u << (ncol - 32 * wq)) - 1u : 0u);
	uint32_t maxm = cntm;
	for (int ofs = 16; ofs; ofs >>= 1) maxm = max(maxm, __shfl_xor_sync(FULL, maxm, ofs));
	// ---- pass 1: pile-up of all members
	for (uint32_t m = 0; m < maxm; ++m) {
		const bool act = have && m < cntm;
		const unsigned long long k2 = act ? el[s + m].y : 0ull;
		const BsMember mb = bs_member_chunk<GW>(packed, WS, Wd, L, k2, posinv0, wq, FULL, act);
		bs_add<K>(cA, ~mb.lo & ~mb.hi & mb.valid); bs_add<K>(cC, mb.lo & ~mb.hi); bs_add<K>(cG, ~mb.lo & mb.hi); bs_add<K>(cT, mb.lo & mb.hi);
	}
	uint32_t conlo, conhi, cov;
	bs_majority<K>(cA, cC, cG, cT, conlo, conhi, cov);
	// the consensus string ends at the first column nobody covers (kthread_bucket.c:117-120); beyond it nothing matches
	const uint32_t empty = ~cov & colmask;
	const int reflen1 = sg_min<GW>(empty ? 32 * wq + __ffs(empty) - 1 : ncol, FULL);
	const uint32_t inref = reflen1 >= 32 * (wq + 1) ? ~0u : (reflen1 > 32 * wq ? (1u << (reflen1 - 32 * wq)) - 1u : 0u);
	// ---- pass 2: mismatches of every member against the consensus; rejected members leave the counters (= pass 3's recount)
	int kept = 0, nrej = 0;
	for (uint32_t m = 0; m < maxm; ++m) {
		const bool act = have && m < cntm;
		const unsigned long long k2 = act ? el[s + m].y : 0ull;
		const BsMember mb = bs_member_chunk<GW>(packed, WS, Wd, L, k2, posinv0, wq, FULL, act);
		const uint32_t diff = (((mb.lo ^ conlo) | (mb.hi ^ conhi)) | ~inref) & mb.valid;
		const int mism = sg_sum<GW>(__popc(diff), FULL);
		const bool keep = mism <= e_thr;
		if (act) {
			if (lead) { o.status[s + m] = keep ? 1 : 2; o.erank[s + m] = keep ? kept : nrej; }
			if (keep) ++kept;
			else {
				++nrej;
				bs_sub<K>(cA, ~mb.lo & ~mb.hi & mb.valid); bs_sub<K>(cC, mb.lo & ~mb.hi); bs_sub<K>(cG, ~mb.lo & mb.hi); bs_sub<K>(cT, mb.lo & mb.hi);
			}
		}
	}
	// ---- finalize: trim to the covered span of the kept members and write the consensus
	const bool iscl = have && kept >= 2;
	int sv = 0;
	if (have) {
		bs_majority<K>(cA, cC, cG, cT, conlo, conhi, cov);
		cov &= colmask;
	} else cov = 0;
	{
		const int first_cov = sg_min<GW>(cov ? 32 * wq + __ffs(cov) - 1 : 1 << 30, FULL);
		const int last_cov = sg_max<GW>(cov ? 32 * wq + 31 - __clz(cov) : -1, FULL);
		unsigned long long refoff = 0;
		int reflen2 = 0;
		if (iscl) { sv = first_cov; reflen2 = last_cov + 1 - sv; }
		// The consensus goes to the scratch string buffer as whole 32-byte pieces: lane wq writes the characters of its 32 columns
		// with two 16-byte stores into a 32-byte aligned slot of ceil(ncol/32) pieces, and the string proper (the covered span) is
		// recorded as slot + sv.  (Byte stores, one column per lane and iteration, were 27 M store sectors for 1.1 M groups.)
		const int npiece = (ncol + 31) >> 5;
		if (iscl && lead) refoff = (atomicAdd(&counters[CT_REFCURSOR], (unsigned long long)(32 * npiece + 32)) + 31ull) & ~31ull;
		refoff = __shfl_sync(FULL, refoff, lane & ~(GW - 1));
		if (iscl) {
			if (refoff + 32ull * npiece > reftmp_cap) { if (lead) atomicAdd(&counters[CT_ERR], 1ull); }
			else if (wq < npiece) {
				uint32_t wd[8];
#pragma unroll
				for (int j = 0; j < 8; ++j) {
					uint32_t l = (conlo >> (4 * j)) & 0xFu, h = (conhi >> (4 * j)) & 0xFu;
					l = (l | (l << 6)) & 0x0303u; l = (l | (l << 3)) & 0x1111u;                             // bit i -> bit 4i (no carries)
					h = (h | (h << 6)) & 0x0303u; h = (h | (h << 3)) & 0x1111u;
					const uint32_t sel = l | (h << 1);                                                     // code of column 4j+i in nibble i
					wd[j] = __byte_perm(0x54474341u, 0u, sel);                                             // "ACGT"[code]
				}
				uint4 *dst = (uint4*)(o.reftmp + refoff + 32ull * wq);
				dst[0] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
				dst[1] = make_uint4(wd[4], wd[5], wd[6], wd[7]);
			}
			refoff += (unsigned long long)sv;
		}
		if (have && lead) {
			const uint32_t nout = iscl ? (uint32_t)nrej : cntm;
			o.g_iscl[g] = iscl; o.g_kept[g] = iscl ? kept : 0;
			o.g_sg[g] = last_round ? nout : 0; o.g_resk[g] = last_round ? 0 : nout;
			o.g_reflen[g] = (unsigned long long)reflen2; o.g_refoff[g] = iscl ? refoff : 0ull;
		}
	}
	__syncwarp();
	// ---- member records (need sv) / lone survivor placement
	for (uint32_t j0 = 0; j0 < maxm; j0 += GW) {
		const uint32_t j = j0 + wq;
		if (have && j < cntm) {
			const unsigned long long k2 = el[s + j].y;
			if (o.status[s + j] == 1) {
				if (iscl) o.newrec[s + j] = ((unsigned long long)mcb_k2_rid(k2) << 32) | ((unsigned long long)((int)(k2 >> MCB_POSINV_SHIFT) - posinv0 - sv) << 1) | (k2 & 1);
				else { o.status[s + j] = 3; o.erank[s + j] = nrej; }            // lone survivor: after the rejects (:476-498)
			}
		}
	}
	__syncwarp();
}

// groups taken from a work list of one size class (K counter planes hold sizes below 2^K); singletons never get here
template <int GW, int K>
__global__ void __launch_bounds__(128)
k_consensus_bs(const ulonglong2 *__restrict__ el, const uint32_t *__restrict__ gstart, const uint32_t *__restrict__ worklist, const unsigned long long *__restrict__ n_work_ptr,
               const uint64_t *__restrict__ packed, int WS, int Wd, int L, int e_thr, int last_round, ConsOut o, unsigned long long *__restrict__ counters, uint64_t reftmp_cap)
{
	constexpr int SGW = 32 / GW;                                   // sub-groups per warp
	const uint64_t n_work = *n_work_ptr;
	const int lane = threadIdx.x & 31;
	const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	for (uint64_t w0 = warp * SGW; w0 < n_work; w0 += nwarps * SGW) {
		const uint64_t wi = w0 + lane / GW;
		const bool have = wi < n_work;
		uint64_t g = 0; uint32_t s = 0, cntm = 0;
		if (have) { g = worklist[wi]; s = gstart[g]; cntm = gstart[g + 1] - s; }
		cons_groups_bs<GW, K>(el, s, cntm, g, have, packed, WS, Wd, L, e_thr, last_round, o, counters, reftmp_cap);
	}
}

#define CONS_BS_MAX_MEMBERS 60000u
// singletons are settled here; groups of 2..15 / 16..255 / 256..CONS_BS_MAX_MEMBERS members go to the work lists of
// k_consensus_bs (4, 8, 16 counter planes), larger ones to the last list, for k_consensus.  The sub-groups of a warp walk their
// members in lock step up to the largest of them, so the 2..15 range (nearly all groups at 20x coverage) is split into seven
// lists of nearly equal sizes: the warps of one launch then carry groups of the same length.
#define CONS_NCLS 10
__global__ void k_cons_worklist(const uint32_t *__restrict__ gstart, uint64_t G, ConsOut o, uint32_t *__restrict__ worklist, unsigned long long *__restrict__ counters)
{
	const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	int cls = -1;
	if (g < G) {
		const uint32_t s = gstart[g], cntm = gstart[g + 1] - s;
		if (cntm < 2) {
			o.status[s] = 0; o.erank[s] = 0;
			o.g_iscl[g] = 0; o.g_kept[g] = 0; o.g_sg[g] = 1; o.g_resk[g] = 0; o.g_reflen[g] = 0; o.g_refoff[g] = 0;
		} else cls = cntm == 2 ? 0 : cntm == 3 ? 1 : cntm == 4 ? 2 : cntm <= 6 ? 3 : cntm <= 8 ? 4 : cntm <= 11 ? 5 : cntm < 16 ? 6
		           : cntm < 256 ? 7 : cntm <= CONS_BS_MAX_MEMBERS ? 8 : 9;
	}
	// one atomic per list and CTA: warp ballots -> per-warp counts in shared memory -> offsets inside the CTA
	__shared__ unsigned long long cta_base[CONS_NCLS];
	__shared__ uint32_t wcnt[8][CONS_NCLS];                  // launched with 256 threads
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	unsigned mybal = 0;
#pragma unroll
	for (int c = 0; c < CONS_NCLS; ++c) {
		const unsigned bal = __ballot_sync(0xFFFFFFFFu, cls == c);
		if (lane == 0) wcnt[wid][c] = __popc(bal);
		if (cls == c) mybal = bal;
	}
	__syncthreads();
	if (threadIdx.x < CONS_NCLS) {
		const int c = threadIdx.x;
		uint32_t tot = 0;
		for (int w = 0; w < 8; ++w) { const uint32_t v = wcnt[w][c]; wcnt[w][c] = tot; tot += v; }
		cta_base[c] = tot ? atomicAdd(&counters[CT_WORK0 + c], (unsigned long long)tot) : 0ull;
	}
	__syncthreads();
	if (cls >= 0) worklist[(uint64_t)cls * G + cta_base[cls] + wcnt[wid][cls] + __popc(mybal & ((1u << lane) - 1u))] = (uint32_t)g;
}

// scatter per element into the compact outputs (bases come from exclusive scans over the group arrays)
struct ScatIn {
	const uint8_t *status; const uint32_t *erank; const uint64_t *newrec; const uint32_t *hscan;
	const uint32_t *g_iscl, *g_kept, *g_sg, *g_resk;     // exclusive-scanned
};
__global__ void k_scatter_members(const ulonglong2 *__restrict__ el, uint64_t n, ScatIn in, int last_round,
                                  uint64_t mem_base, uint64_t sg_base, uint64_t *__restrict__ cl_a, uint32_t *__restrict__ sg, uint32_t *__restrict__ resk)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	bool head = i == 0 || el[i].x != el[i - 1].x;
	uint32_t g = in.hscan[i] - (head ? 0u : 1u);
	uint8_t st = in.status[i];
	uint32_t rid = mcb_k2_rid(el[i].y);
	if (st == 1) cl_a[mem_base + in.g_kept[g] + in.erank[i]] = in.newrec[i];
	else if (st == 0 || last_round) sg[sg_base + in.g_sg[g] + in.erank[i]] = rid;
	else resk[in.g_resk[g] + in.erank[i]] = rid;
}
// one warp per group: cluster table entries + consensus string copy
__global__ void k_scatter_clusters(const uint32_t *__restrict__ gstart, uint64_t G, const uint32_t *__restrict__ g_iscl_scan, const uint32_t *__restrict__ g_kept_scan,
                                   const unsigned long long *__restrict__ g_reflen_scan, const unsigned long long *__restrict__ g_refoff,
                                   const unsigned long long *__restrict__ totals /* [CT_*] */, const char *__restrict__ reftmp,
                                   uint64_t cl_base, uint64_t mem_base, uint64_t ref_base,
                                   uint32_t *__restrict__ cl_n, uint64_t *__restrict__ cl_a_off, uint64_t *__restrict__ cl_ref_off, char *__restrict__ cl_ref)
{
	const int lane = threadIdx.x & 31;
	uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (g >= G) return;
	const uint32_t ci = g_iscl_scan[g];
	const uint32_t ci_next = g + 1 < G ? g_iscl_scan[g + 1] : (uint32_t)totals[CT_TOT_CL];
	if (ci_next == ci) return;                                    // not a cluster
	const uint32_t kb = g_kept_scan[g];
	const uint32_t kb_next = g + 1 < G ? g_kept_scan[g + 1] : (uint32_t)totals[CT_TOT_MEM];
	const unsigned long long rb = g_reflen_scan[g];
	const unsigned long long rb_next = g + 1 < G ? g_reflen_scan[g + 1] : totals[CT_TOT_REF];
	if (lane == 0) {
		cl_n[cl_base + ci] = kb_next - kb;
		cl_a_off[cl_base + ci] = mem_base + kb;
		cl_ref_off[cl_base + ci] = ref_base + rb;
	}
	const unsigned long long len = rb_next - rb, so = g_refoff[g];
	for (unsigned long long c = lane; c < len; c += 32) cl_ref[ref_base + rb + c] = reftmp[so + c];
}

// K4 (first m windowed minimizers of each new seed contig): k_sketch_lh2 in mcb_lh.cuh

// ================================================================= host side
static int sync_counters(mcb_ctx *ctx, unsigned long long **hc)
{
	MCB_TRY(ctx->h_counters.ensure(64 * 8));
	MCB_CUDA(cudaMemcpyAsync(ctx->h_counters.p, ctx->d_counters.p, 64 * 8, cudaMemcpyDeviceToHost, ctx->stream));
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	*hc = ctx->h_counters.as<unsigned long long>();
	return MCB_OK;
}

static int ensure_elems(mcb_ctx *ctx, uint64_t n, bool keep_cur)
{
	if (n <= ctx->elem_cap) return MCB_OK;
	(void)keep_cur;
	MCB_TRY(ctx->d_elemA.ensure((size_t)n * 16 + 16));
	MCB_TRY(ctx->d_elemB.ensure((size_t)n * 16 + 16));
	ctx->elem_cap = std::min(ctx->d_elemA.cap, ctx->d_elemB.cap) / 16;
	return MCB_OK;
}

// h_rows != nullptr: page-locked host rows that still have to be copied to d_rows; the copy is cut into chunks on a second
// stream and the pack/sketch kernel of a chunk runs while the next chunk is on the bus
static int for_reads_impl(mcb_ctx *ctx, const uint8_t *d_rows, uint64_t n, mcb_reads_result *res, const char *h_rows = nullptr)
{
	const int L = ctx->L, Wd = ctx->Wd, WS = ctx->WS;
	const int pbase = L + ctx->prm.max_rounds;
	if (ctx->shard_n > 1) {
		if (ctx->rid_base + n > ctx->n_reads) { mcb_set_error("mcb_for_reads: slice [%llu,+%llu) exceeds the %llu reads declared by mcb_shard_begin", (unsigned long long)ctx->rid_base, (unsigned long long)n, (unsigned long long)ctx->n_reads); return MCB_EINVAL; }
	} else { ctx->n_reads = n; ctx->rid_base = 0; }
	ctx->n_local = n; ctx->reads_loaded = false; ctx->bucket_done = false; ctx->bs.active = false;
	ctx->cix.valid = false;              // a new read set starts a new run: contigs of the previous one are never reused
	const uint64_t rid_base = ctx->rid_base;
	MCB_TRY(ctx->d_packed.ensure((size_t)(ctx->n_reads + ctx->shard_n) * WS * 8 + 16));   // slack: equal-sized all-gather chunks may overhang by < n_ranks rows
	MCB_TRY(ctx->d_cls.ensure(n + 16));
	ctx->elem_cap = std::min(ctx->d_elemA.cap, ctx->d_elemB.cap) / 16;
	MCB_TRY(ensure_elems(ctx, n + 1, false));
	MCB_TRY(ctx->d_counters.ensure(64 * 8));
	MCB_CUDA(cudaMemsetAsync(ctx->d_counters.p, 0, 64 * 8, ctx->stream));
	unsigned long long *dc = ctx->d_counters.as<unsigned long long>();
	unsigned long long *hc = nullptr;
	const size_t smem = RD_THREADS * 9 * sizeof(uint64_t) + (((size_t)RD_THREADS * L + 15) & ~(size_t)15) + 16;
	if (n && !h_rows) {
		McbSpan sp(ctx->tm, "for_reads");
		MCB_LAUNCH(ctx, "pack_classify_sketch", k_pack_classify_sketch, mcb_grid_for(n, RD_THREADS), RD_THREADS, smem,
		           d_rows, n, rid_base, L, Wd, WS, ctx->prm.k, mcb_ta_mul(ctx->prm.k), ctx->prm.diff_threshold, pbase,
		           ctx->d_packed.as<uint64_t>(), ctx->d_cls.as<uint8_t>(), ctx->d_elemA.as<ulonglong2>(), dc);
	} else if (n) {
		if (!ctx->copy_stream) MCB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
		const uint64_t CH = 1u << 20;                                   // reads per chunk (a multiple of 128: tiles and 16-byte alignment hold)
		const int nch = (int)((n + CH - 1) / CH);
		std::vector<cudaEvent_t> ev((size_t)nch);
		cudaEvent_t t0, t1;
		cudaEventCreate(&t0); cudaEventCreate(&t1);
		cudaEventRecord(t0, ctx->copy_stream);
		for (int c = 0; c < nch; ++c) {
			const uint64_t off = (uint64_t)c * CH, cnt = std::min(CH, n - off);
			cudaEventCreateWithFlags(&ev[c], cudaEventDisableTiming);
			MCB_CUDA(cudaMemcpyAsync((char*)d_rows + off * L, h_rows + off * L, cnt * L, cudaMemcpyHostToDevice, ctx->copy_stream));
			cudaEventRecord(ev[c], ctx->copy_stream);
		}
		cudaEventRecord(t1, ctx->copy_stream);
		for (int c = 0; c < nch; ++c) {
			const uint64_t off = (uint64_t)c * CH, cnt = std::min(CH, n - off);
			MCB_CUDA(cudaStreamWaitEvent(ctx->stream, ev[c], 0));
			McbSpan sp(ctx->tm, "for_reads");
			MCB_LAUNCH(ctx, "pack_classify_sketch", k_pack_classify_sketch, mcb_grid_for(cnt, RD_THREADS), RD_THREADS, smem,
			           d_rows + off * L, cnt, rid_base + off, L, Wd, WS, ctx->prm.k, mcb_ta_mul(ctx->prm.k), ctx->prm.diff_threshold, pbase,
			           ctx->d_packed.as<uint64_t>(), ctx->d_cls.as<uint8_t>() + off, ctx->d_elemA.as<ulonglong2>() + off, dc);
		}
		MCB_CUDA(cudaStreamSynchronize(ctx->copy_stream));
		MCB_CUDA(cudaStreamSynchronize(ctx->stream));
		if (ctx->tm.enabled) { float ms = 0; if (cudaEventElapsedTime(&ms, t0, t1) == cudaSuccess) { int id = ctx->tm.id("h2d"); ctx->tm.ms[id] += ms; ctx->tm.cnt[id] += 1; } }
		for (auto e : ev) cudaEventDestroy(e);
		cudaEventDestroy(t0); cudaEventDestroy(t1);
	}
	MCB_TRY(sync_counters(ctx, &hc));
	if (hc[CT_BADCHAR]) { mcb_set_error("%llu reads contain characters other than A,C,G,T,N (unsupported; the reference's behaviour on them is undefined)", hc[CT_BADCHAR]); return MCB_EINPUT; }
	if (hc[CT_DEGENERATE]) { mcb_set_error("%llu reads have no valid k-mer (reference would index read 0xFFFFFFFF)", hc[CT_DEGENERATE]); return MCB_EINPUT; }
	const uint64_t n_sk = hc[CT_SKETCHED], nn = hc[CT_NREADS];
	ctx->n_valid_round1 = n_sk; ctx->n_nreads = nn;
	// ---- side table of reads with N
	MCB_TRY(ctx->d_nread_rid.ensure(nn * 4 + 16));
	MCB_TRY(ctx->d_nread_mask.ensure(nn * WS * 8 + 16));
	MCB_TRY(ctx->d_scr[0].ensure(n * 4 + 16));
	MCB_TRY(ctx->d_scr[1].ensure(nn + 16));
	{
		McbSpan sp(ctx->tm, "for_reads");
		if (n && nn) {
			MCB_LAUNCH(ctx, "hasn_flags", k_hasn_flags, mcb_grid_for(n, 256), 256, 0, ctx->d_cls.as<uint8_t>(), n, ctx->d_scr[0].as<uint32_t>());
			MCB_TRY(mcb_exclusive_scan_u32(ctx, ctx->d_scr[0].as<uint32_t>(), n, nullptr));
			MCB_LAUNCH(ctx, "hasn_extract", k_hasn_extract, mcb_grid_for(n, 256), 256, 0, d_rows, ctx->d_cls.as<uint8_t>(), n, rid_base, L, Wd, WS,
			           ctx->prm.diff_threshold, ctx->d_scr[0].as<uint32_t>(), ctx->d_nread_rid.as<uint32_t>(), ctx->d_nread_mask.as<uint64_t>(), ctx->d_scr[1].as<uint8_t>());
		}
	}
	// ---- results to the host
	MCB_TRY(ctx->h_cls.ensure(n + 16));
	MCB_TRY(ctx->h_nrid.ensure(nn * 4 + 16));
	MCB_TRY(ctx->h_nrepl.ensure(nn + 16));
	MCB_TRY(ctx->h_nmask.ensure(nn * WS * 8 + 16));
	{
		McbSpan sp(ctx->tm, "d2h");
		if (n) MCB_CUDA(cudaMemcpyAsync(ctx->h_cls.p, ctx->d_cls.p, n, cudaMemcpyDeviceToHost, ctx->stream));
		if (nn) {
			MCB_CUDA(cudaMemcpyAsync(ctx->h_nrid.p, ctx->d_nread_rid.p, nn * 4, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(ctx->h_nrepl.p, ctx->d_scr[1].p, nn, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(ctx->h_nmask.p, ctx->d_nread_mask.p, nn * WS * 8, cudaMemcpyDeviceToHost, ctx->stream));
		}
	}
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->tm.collect();
	// N positions from the masks
	MCB_TRY(ctx->h_noff.ensure((nn + 1) * 8));
	uint64_t *noff = ctx->h_noff.as<uint64_t>();
	const uint64_t *nm = ctx->h_nmask.as<uint64_t>();
	uint64_t tot = 0;
	for (uint64_t i = 0; i < nn; ++i) { noff[i] = tot; for (int w = 0; w < Wd; ++w) tot += __builtin_popcountll(nm[i * WS + w]); }
	noff[nn] = tot;
	MCB_TRY(ctx->h_npos.ensure(tot * 4 + 16));
	uint32_t *np = ctx->h_npos.as<uint32_t>();
	for (uint64_t i = 0, o = 0; i < nn; ++i)
		for (int w = 0; w < Wd; ++w) { uint64_t m = nm[i * WS + w]; while (m) { int bit = __builtin_ctzll(m); np[o++] = (uint32_t)(w * 32 + bit / 2); m &= m - 1; } }
	res->n_reads = n; res->cls = ctx->h_cls.as<uint8_t>();
	res->n_nreads = nn; res->nread_rid = ctx->h_nrid.as<uint32_t>(); res->nread_repl = ctx->h_nrepl.as<uint8_t>();
	res->nread_off = noff; res->npos = np; res->n_sketched = n_sk;
	ctx->reads_loaded = true;
	return MCB_OK;
}

static int check_ctx(mcb_ctx *ctx) { if (!ctx) { mcb_set_error("null context"); return MCB_EINVAL; } MCB_CUDA(cudaSetDevice(ctx->prm.device)); return MCB_OK; }

extern "C" int mcb_for_reads_device(mcb_ctx *ctx, const char *d_rows, uint64_t n, mcb_reads_result *res)
{
	MCB_TRY(check_ctx(ctx));
	if (!res || (n && !d_rows)) { mcb_set_error("mcb_for_reads_device: null argument"); return MCB_EINVAL; }
	if (((uintptr_t)d_rows & 15) != 0) { mcb_set_error("mcb_for_reads_device: rows must be 16-byte aligned"); return MCB_EINVAL; }
	if (n >= (1ull << 31)) { mcb_set_error("too many reads (rid is a signed 32-bit int in the reference, kthread_bucket.c:48)"); return MCB_EINVAL; }
	return for_reads_impl(ctx, (const uint8_t*)d_rows, n, res);
}

extern "C" int mcb_for_reads(mcb_ctx *ctx, const char *rows, uint64_t n, mcb_reads_result *res)
{
	MCB_TRY(check_ctx(ctx));
	if (!res || (n && !rows)) { mcb_set_error("mcb_for_reads: null argument"); return MCB_EINVAL; }
	const size_t bytes = (size_t)n * ctx->L;
	MCB_TRY(ctx->d_ascii.ensure(bytes + 16));
	if (n >= (1ull << 31)) { mcb_set_error("too many reads (rid is a signed 32-bit int in the reference, kthread_bucket.c:48)"); return MCB_EINVAL; }
	cudaPointerAttributes pa;
	const bool pinned = cudaPointerGetAttributes(&pa, rows) == cudaSuccess && pa.type == cudaMemoryTypeHost;
	cudaGetLastError();
	if (pinned && n) return for_reads_impl(ctx, ctx->d_ascii.as<uint8_t>(), n, res, rows);     // copy and compute overlap chunk by chunk
	{
		McbSpan sp(ctx->tm, "h2d");
		MCB_TRY(mcb_h2d(ctx, ctx->d_ascii.p, rows, bytes, 4));
	}
	return mcb_for_reads_device(ctx, ctx->d_ascii.as<char>(), n, res);
}

extern "C" int mcb_for_reads_ptrs(mcb_ctx *ctx, const void *first_seq_ptr, size_t stride, uint64_t n, int n_threads, mcb_reads_result *res)
{
	MCB_TRY(check_ctx(ctx));
	if (!res || (n && !first_seq_ptr)) { mcb_set_error("mcb_for_reads_ptrs: null argument"); return MCB_EINVAL; }
	const int L = ctx->L;
	const size_t bytes = (size_t)n * L;
	MCB_TRY(ctx->d_ascii.ensure(bytes + 16));
	if (n_threads < 1) n_threads = 1;
	{
		McbSpan sp(ctx->tm, "h2d");
		const uint64_t RCH = std::max<uint64_t>(1, (32u << 20) / L);          // reads per chunk
		const size_t CH = (size_t)RCH * L;
		MCB_TRY(ctx->h_stage.ensure(2 * CH));
		cudaEvent_t ev[2]; cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming); cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
		int slot = 0; bool short_read = false;
		for (uint64_t r0 = 0; r0 < n; r0 += RCH, slot ^= 1) {
			uint64_t cnt = std::min(RCH, n - r0);
			char *st = ctx->h_stage.as<char>() + (size_t)slot * CH;
			cudaEventSynchronize(ev[slot]);
			auto work = [&](uint64_t a, uint64_t b) {
				for (uint64_t i = a; i < b; ++i) {
					const char *s = *(const char *const *)((const char*)first_seq_ptr + (r0 + i) * stride);
					if (memchr(s, 0, L)) short_read = true; else memcpy(st + i * L, s, L);
				}
			};
			int T = (int)std::min<uint64_t>(n_threads, (cnt + 4095) / 4096);
			if (T <= 1) work(0, cnt);
			else {
				std::vector<std::thread> th;
				for (int t = 0; t < T; ++t) th.emplace_back(work, cnt * t / T, cnt * (t + 1) / T);
				for (auto &x : th) x.join();
			}
			if (short_read) { cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]); mcb_set_error("a read is shorter than readlen=%d", L); return MCB_EINPUT; }
			MCB_CUDA(cudaMemcpyAsync(ctx->d_ascii.as<char>() + r0 * L, st, cnt * L, cudaMemcpyHostToDevice, ctx->stream));
			cudaEventRecord(ev[slot], ctx->stream);
		}
		cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
	}
	return mcb_for_reads_device(ctx, ctx->d_ascii.as<char>(), n, res);
}

// ---------------------------------------------------------------- kt_for_reads on packed rows
// rows / nrid / nmask: host memory (on_device = false: page-locked rows are uploaded chunk by chunk on the copy stream, straight
// into the read table, and the kernel of a chunk runs while the next chunk is on the bus) or device memory (on_device = true)
static int for_reads_packed_impl(mcb_ctx *ctx, const uint64_t *rows, uint64_t n, const uint32_t *nrid, const uint64_t *nmask, uint64_t nn, bool on_device, mcb_reads_result *res)
{
	const int L = ctx->L, Wd = ctx->Wd, WS = ctx->WS;
	const int pbase = L + ctx->prm.max_rounds;
	if (nn > n) { mcb_set_error("mcb_for_reads_packed: more reads with N (%llu) than reads (%llu)", (unsigned long long)nn, (unsigned long long)n); return MCB_EINVAL; }
	if (ctx->shard_n > 1) {
		if (ctx->rid_base + n > ctx->n_reads) { mcb_set_error("mcb_for_reads_packed: slice [%llu,+%llu) exceeds the %llu reads declared by mcb_shard_begin", (unsigned long long)ctx->rid_base, (unsigned long long)n, (unsigned long long)ctx->n_reads); return MCB_EINVAL; }
	} else { ctx->n_reads = n; ctx->rid_base = 0; }
	ctx->n_local = n; ctx->reads_loaded = false; ctx->bucket_done = false; ctx->bs.active = false;
	ctx->cix.valid = false;
	const uint64_t rid_base = ctx->rid_base;
	MCB_TRY(ctx->d_packed.ensure((size_t)(ctx->n_reads + ctx->shard_n) * WS * 8 + 16));
	MCB_TRY(ctx->d_cls.ensure(n + 16));
	ctx->elem_cap = std::min(ctx->d_elemA.cap, ctx->d_elemB.cap) / 16;
	MCB_TRY(ensure_elems(ctx, n + 1, false));
	MCB_TRY(ctx->d_counters.ensure(64 * 8));
	MCB_CUDA(cudaMemsetAsync(ctx->d_counters.p, 0, 64 * 8, ctx->stream));
	unsigned long long *dc = ctx->d_counters.as<unsigned long long>();
	unsigned long long *hc = nullptr;
	uint64_t *slice = ctx->d_packed.as<uint64_t>() + rid_base * WS;
	// ---- side table of the reads with N, and the bitmap that keeps the main kernel off them
	MCB_TRY(ctx->d_nread_rid.ensure(nn * 4 + 16));
	MCB_TRY(ctx->d_nread_mask.ensure(nn * WS * 8 + 16));
	MCB_TRY(ctx->d_scr[0].ensure((n / 32 + 1) * 4 + 16));
	MCB_TRY(ctx->d_scr[1].ensure(nn + 16));
	MCB_TRY(ctx->d_scr[2].ensure(nn * 4 + 16));
	const uint32_t *d_nrid_local = nullptr;
	const uint64_t *d_nmask = nullptr;
	uint32_t *bits = nullptr;
	if (nn) {
		if (on_device) { d_nrid_local = nrid; d_nmask = nmask; }
		else {
			McbSpan sp(ctx->tm, "h2d");
			MCB_TRY(mcb_h2d(ctx, ctx->d_scr[2].p, nrid, nn * 4, 1));
			MCB_TRY(mcb_h2d(ctx, ctx->d_nread_mask.p, nmask, nn * WS * 8, 1));
			d_nrid_local = ctx->d_scr[2].as<uint32_t>(); d_nmask = ctx->d_nread_mask.as<uint64_t>();
		}
		McbSpan sp(ctx->tm, "for_reads");
		bits = ctx->d_scr[0].as<uint32_t>();
		MCB_CUDA(cudaMemsetAsync(bits, 0, (n / 32 + 1) * 4, ctx->stream));
		MCB_LAUNCH(ctx, "mark_nreads", k_mark_nreads, mcb_grid_for(nn, 256), 256, 0, d_nrid_local, nn, n, bits, dc);
	}
	// ---- the rows
	bool pinned = false;
	if (!on_device && n) {
		cudaPointerAttributes pa;
		pinned = cudaPointerGetAttributes(&pa, rows) == cudaSuccess && pa.type == cudaMemoryTypeHost;
		cudaGetLastError();
	}
	if (n && (on_device || !pinned)) {
		const uint64_t *src = rows;
		if (!on_device) {
			McbSpan sp(ctx->tm, "h2d");
			MCB_TRY(mcb_h2d(ctx, slice, rows, (size_t)n * WS * 8, 4));
			src = slice;
		}
		McbSpan sp(ctx->tm, "for_reads");
		MCB_LAUNCH(ctx, "classify_sketch_packed", k_classify_sketch_packed, mcb_grid_for(n, RD_THREADS), RD_THREADS, 0, src, slice, n, (uint64_t)0, rid_base, L, Wd, WS,
		           ctx->prm.k, mcb_ta_mul(ctx->prm.k), ctx->prm.diff_threshold, pbase, bits, ctx->d_cls.as<uint8_t>(), ctx->d_elemA.as<ulonglong2>(), dc);
		if (nn) MCB_LAUNCH(ctx, "nreads_packed", k_nreads_packed, mcb_grid_for(nn, RD_THREADS), RD_THREADS, 0, src, slice, d_nrid_local, d_nmask, nn, n, rid_base, L, Wd, WS,
		                   ctx->prm.k, mcb_ta_mul(ctx->prm.k), ctx->prm.diff_threshold, pbase, ctx->d_nread_rid.as<uint32_t>(), ctx->d_scr[1].as<uint8_t>(), ctx->d_cls.as<uint8_t>(), ctx->d_elemA.as<ulonglong2>(), dc);
	} else if (n) {
		if (!ctx->copy_stream) MCB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
		const uint64_t CH = 1u << 21;                                   // reads per chunk: 64 MB at L = 100
		const int nch = (int)((n + CH - 1) / CH);
		std::vector<cudaEvent_t> ev((size_t)nch);
		cudaEvent_t t0, t1;
		cudaEventCreate(&t0); cudaEventCreate(&t1);
		cudaEventRecord(t0, ctx->copy_stream);
		for (int c = 0; c < nch; ++c) {
			const uint64_t off = (uint64_t)c * CH, cnt = std::min(CH, n - off);
			cudaEventCreateWithFlags(&ev[c], cudaEventDisableTiming);
			MCB_CUDA(cudaMemcpyAsync(slice + off * WS, rows + off * WS, cnt * WS * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
			cudaEventRecord(ev[c], ctx->copy_stream);
		}
		cudaEventRecord(t1, ctx->copy_stream);
		for (int c = 0; c < nch; ++c) {
			const uint64_t off = (uint64_t)c * CH, cnt = std::min(CH, n - off);
			MCB_CUDA(cudaStreamWaitEvent(ctx->stream, ev[c], 0));
			McbSpan sp(ctx->tm, "for_reads");
			MCB_LAUNCH(ctx, "classify_sketch_packed", k_classify_sketch_packed, mcb_grid_for(cnt, RD_THREADS), RD_THREADS, 0, slice + off * WS, slice + off * WS, cnt, off, rid_base, L, Wd, WS,
			           ctx->prm.k, mcb_ta_mul(ctx->prm.k), ctx->prm.diff_threshold, pbase, bits, ctx->d_cls.as<uint8_t>(), ctx->d_elemA.as<ulonglong2>(), dc);
		}
		if (nn) {
			McbSpan sp(ctx->tm, "for_reads");
			MCB_LAUNCH(ctx, "nreads_packed", k_nreads_packed, mcb_grid_for(nn, RD_THREADS), RD_THREADS, 0, slice, slice, d_nrid_local, d_nmask, nn, n, rid_base, L, Wd, WS,
			           ctx->prm.k, mcb_ta_mul(ctx->prm.k), ctx->prm.diff_threshold, pbase, ctx->d_nread_rid.as<uint32_t>(), ctx->d_scr[1].as<uint8_t>(), ctx->d_cls.as<uint8_t>(), ctx->d_elemA.as<ulonglong2>(), dc);
		}
		MCB_CUDA(cudaStreamSynchronize(ctx->copy_stream));
		MCB_CUDA(cudaStreamSynchronize(ctx->stream));
		if (ctx->tm.enabled) { float ms = 0; if (cudaEventElapsedTime(&ms, t0, t1) == cudaSuccess) { int id = ctx->tm.id("h2d"); ctx->tm.ms[id] += ms; ctx->tm.cnt[id] += 1; } }
		for (auto e : ev) cudaEventDestroy(e);
		cudaEventDestroy(t0); cudaEventDestroy(t1);
	}
	if (on_device && nn && d_nmask != ctx->d_nread_mask.as<uint64_t>())
		MCB_CUDA(cudaMemcpyAsync(ctx->d_nread_mask.p, d_nmask, nn * WS * 8, cudaMemcpyDeviceToDevice, ctx->stream));
	MCB_TRY(sync_counters(ctx, &hc));
	if (hc[CT_BADCHAR]) { mcb_set_error("%llu packed reads are malformed (non-zero unused bits, or an N table that is unsorted, out of range or inconsistent with the rows)", hc[CT_BADCHAR]); return MCB_EINPUT; }
	if (hc[CT_DEGENERATE]) { mcb_set_error("%llu reads have no valid k-mer (reference would index read 0xFFFFFFFF)", hc[CT_DEGENERATE]); return MCB_EINPUT; }
	if (hc[CT_NREADS] != nn) { mcb_set_error("internal: %llu of %llu reads with N processed", hc[CT_NREADS], (unsigned long long)nn); return MCB_ECUDA; }
	ctx->n_valid_round1 = hc[CT_SKETCHED]; ctx->n_nreads = nn;
	// ---- results to the host
	MCB_TRY(ctx->h_cls.ensure(n + 16));
	MCB_TRY(ctx->h_nrid.ensure(nn * 4 + 16));
	MCB_TRY(ctx->h_nrepl.ensure(nn + 16));
	MCB_TRY(ctx->h_nmask.ensure(nn * WS * 8 + 16));
	{
		McbSpan sp(ctx->tm, "d2h");
		if (n) MCB_CUDA(cudaMemcpyAsync(ctx->h_cls.p, ctx->d_cls.p, n, cudaMemcpyDeviceToHost, ctx->stream));
		if (nn) {
			MCB_CUDA(cudaMemcpyAsync(ctx->h_nrid.p, ctx->d_nread_rid.p, nn * 4, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(ctx->h_nrepl.p, ctx->d_scr[1].p, nn, cudaMemcpyDeviceToHost, ctx->stream));
			if (on_device) MCB_CUDA(cudaMemcpyAsync(ctx->h_nmask.p, ctx->d_nread_mask.p, nn * WS * 8, cudaMemcpyDeviceToHost, ctx->stream));
		}
	}
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->tm.collect();
	const uint64_t *nm = on_device ? ctx->h_nmask.as<uint64_t>() : nmask;       // the caller's masks are the N positions
	MCB_TRY(ctx->h_noff.ensure((nn + 1) * 8));
	uint64_t *noff = ctx->h_noff.as<uint64_t>();
	uint64_t tot = 0;
	for (uint64_t i = 0; i < nn; ++i) { noff[i] = tot; for (int w = 0; w < Wd; ++w) tot += __builtin_popcountll(nm[i * WS + w]); }
	noff[nn] = tot;
	MCB_TRY(ctx->h_npos.ensure(tot * 4 + 16));
	uint32_t *np = ctx->h_npos.as<uint32_t>();
	for (uint64_t i = 0, o = 0; i < nn; ++i)
		for (int w = 0; w < Wd; ++w) { uint64_t m = nm[i * WS + w]; while (m) { int bit = __builtin_ctzll(m); np[o++] = (uint32_t)(w * 32 + bit / 2); m &= m - 1; } }
	res->n_reads = n; res->cls = ctx->h_cls.as<uint8_t>();
	res->n_nreads = nn; res->nread_rid = ctx->h_nrid.as<uint32_t>(); res->nread_repl = ctx->h_nrepl.as<uint8_t>();
	res->nread_off = noff; res->npos = np; res->n_sketched = hc[CT_SKETCHED];
	ctx->reads_loaded = true;
	return MCB_OK;
}

extern "C" int mcb_for_reads_packed(mcb_ctx *ctx, const uint64_t *packed, uint64_t n, const uint32_t *nread_rid, const uint64_t *nmask, uint64_t n_nreads, mcb_reads_result *res)
{
	MCB_TRY(check_ctx(ctx));
	if (!res || (n && !packed) || (n_nreads && (!nread_rid || !nmask))) { mcb_set_error("mcb_for_reads_packed: null argument"); return MCB_EINVAL; }
	if (n >= (1ull << 31)) { mcb_set_error("too many reads (rid is a signed 32-bit int in the reference, kthread_bucket.c:48)"); return MCB_EINVAL; }
	return for_reads_packed_impl(ctx, packed, n, nread_rid, nmask, n_nreads, false, res);
}

extern "C" int mcb_for_reads_packed_device(mcb_ctx *ctx, const uint64_t *d_packed, uint64_t n, const uint32_t *d_nread_rid, const uint64_t *d_nmask, uint64_t n_nreads, mcb_reads_result *res)
{
	MCB_TRY(check_ctx(ctx));
	if (!res || (n && !d_packed) || (n_nreads && (!d_nread_rid || !d_nmask))) { mcb_set_error("mcb_for_reads_packed_device: null argument"); return MCB_EINVAL; }
	if (((uintptr_t)d_packed & 15) != 0) { mcb_set_error("mcb_for_reads_packed_device: rows must be 16-byte aligned"); return MCB_EINVAL; }
	if (n >= (1ull << 31)) { mcb_set_error("too many reads (rid is a signed 32-bit int in the reference, kthread_bucket.c:48)"); return MCB_EINVAL; }
	return for_reads_packed_impl(ctx, d_packed, n, d_nread_rid, d_nmask, n_nreads, true, res);
}

extern "C" int mcb_debug_read_tuples(mcb_ctx *ctx, mcb_tuple *out)
{
	MCB_TRY(check_ctx(ctx));
	if (!ctx->reads_loaded || ctx->bucket_done || ctx->bs.active) { mcb_set_error("mcb_debug_read_tuples: call right after mcb_for_reads"); return MCB_ESTATE; }
	const uint64_t n = ctx->n_local;
	if (!n) return MCB_OK;
	MCB_TRY(ctx->d_scr[0].ensure(n * 16));
	MCB_LAUNCH(ctx, "elem_to_tuple", k_elem_to_tuple, mcb_grid_for(n, 256), 256, 0, ctx->d_elemA.as<ulonglong2>(), n, ctx->L, ctx->prm.k,
	           ctx->L + ctx->prm.max_rounds, ctx->d_scr[0].as<mcb_tuple>(), 1, ctx->rid_base);
	MCB_CUDA(cudaMemcpyAsync(out, ctx->d_scr[0].p, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	return MCB_OK;
}

extern "C" int mcb_debug_sketch_two(mcb_ctx *ctx, const uint32_t *rids, uint64_t n, int k, mcb_tuple *out)
{
	MCB_TRY(check_ctx(ctx));
	if (!ctx->reads_loaded) { mcb_set_error("mcb_debug_sketch_two: no reads loaded"); return MCB_ESTATE; }
	if (k < 1 || k > 31) { mcb_set_error("k out of range"); return MCB_EINVAL; }
	if (!n) return MCB_OK;
	for (uint64_t i = 0; i < n; ++i) if (rids[i] >= ctx->n_reads) { mcb_set_error("rid out of range"); return MCB_EINVAL; }
	MCB_TRY(ctx->d_scr[0].ensure(n * 4)); MCB_TRY(ctx->d_scr[1].ensure(n * 16)); MCB_TRY(ctx->d_scr[2].ensure(n * 16));
	MCB_CUDA(cudaMemcpyAsync(ctx->d_scr[0].p, rids, n * 4, cudaMemcpyHostToDevice, ctx->stream));
	// note: positions are encoded against prm.k, decode with the same
	MCB_LAUNCH(ctx, "resketch", k_resketch, mcb_grid_for(n, 128), 128, 0, ctx->d_packed.as<uint64_t>(), ctx->WS, ctx->L, ctx->prm.k, k, mcb_ta_mul(k),
	           ctx->L + ctx->prm.max_rounds, ctx->d_scr[0].as<uint32_t>(), n, ctx->d_scr[1].as<ulonglong2>(), ctx->d_counters.as<unsigned long long>());
	MCB_LAUNCH(ctx, "elem_to_tuple", k_elem_to_tuple, mcb_grid_for(n, 256), 256, 0, ctx->d_scr[1].as<ulonglong2>(), n, ctx->L, ctx->prm.k,
	           ctx->L + ctx->prm.max_rounds, ctx->d_scr[2].as<mcb_tuple>(), 0, (uint64_t)0);
	MCB_CUDA(cudaMemcpyAsync(out, ctx->d_scr[2].p, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	return MCB_OK;
}

extern "C" int mcb_debug_unpack_reads(mcb_ctx *ctx, char *rows_out)
{
	MCB_TRY(check_ctx(ctx));
	if (!ctx->reads_loaded) { mcb_set_error("mcb_debug_unpack_reads: no reads loaded"); return MCB_ESTATE; }
	const uint64_t tot = ctx->n_reads * (uint64_t)ctx->L;
	if (!tot) return MCB_OK;
	MCB_TRY(ctx->d_scr[0].ensure(tot));
	MCB_LAUNCH(ctx, "unpack", k_unpack, mcb_grid_for(tot, 256), 256, 0, ctx->d_packed.as<uint64_t>(), ctx->n_reads, ctx->L, ctx->WS, ctx->d_scr[0].as<char>());
	MCB_CUDA(cudaMemcpyAsync(rows_out, ctx->d_scr[0].p, tot, cudaMemcpyDeviceToHost, ctx->stream));
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	return MCB_OK;
}

// ---------------------------------------------------------------- kt_for_bucket
// device accumulation buffers that must survive growth
static int grow_preserve(mcb_ctx *ctx, DBuf &b, size_t used, size_t need)
{
	if (need <= b.cap) return MCB_OK;
	DBuf nb; MCB_TRY(nb.ensure(need + need / 2));
	if (used) MCB_CUDA(cudaMemcpyAsync(nb.p, b.p, used, cudaMemcpyDeviceToDevice, ctx->stream));
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	b.release(); b = nb;
	return MCB_OK;
}

// ---- the round loop of kt_for_bucket (kthread_bucket.c:562-629) as begin / round_a / round_b / finish, so that the sharded
// driver can exchange tuples and cluster counts between the halves; mcb_for_bucket below is the single-GPU loop over them.
struct BucketBufs {
	DBuf &b_hs, &b_gs, &b_st, &b_er, &b_nr, &b_gc, &b_gk, &b_gsg, &b_grk, &b_grl, &b_gro, &b_rk;
	DBuf &d_cl_n, &d_cl_aoff, &d_cl_a, &d_cl_roff, &d_cl_ref, &d_sg, &d_mi, &d_micnt;
	// d_scr roles: 0 head scan, 1 gstart, 2 status, 3 erank, 4 newrec, 5..9 group arrays, 10 refoff, 11 resk list
	// d_out: accumulated outputs (kept in the context: allocating them per call costs more than all the kernels of this entry point together)
	explicit BucketBufs(mcb_ctx *c) : b_hs(c->d_scr[0]), b_gs(c->d_scr[1]), b_st(c->d_scr[2]), b_er(c->d_scr[3]), b_nr(c->d_scr[4]), b_gc(c->d_scr[5]), b_gk(c->d_scr[6]),
		b_gsg(c->d_scr[7]), b_grk(c->d_scr[8]), b_grl(c->d_scr[9]), b_gro(c->d_scr[10]), b_rk(c->d_scr[11]),
		d_cl_n(c->d_out[0]), d_cl_aoff(c->d_out[1]), d_cl_a(c->d_out[2]), d_cl_roff(c->d_out[3]), d_cl_ref(c->d_out[4]), d_sg(c->d_out[5]), d_mi(c->d_out[6]), d_micnt(c->d_out[7]) {}
};
static inline int consensus_cols(const mcb_ctx *ctx) { return ((2 * ctx->L + 2 * ctx->prm.max_rounds + 8 + 31) / 32) * 32; }

int mcb_bucket_begin(mcb_ctx *ctx)
{
	if (!ctx->reads_loaded || ctx->bucket_done) { mcb_set_error("kt_for_bucket: needs a fresh mcb_for_reads"); return MCB_ESTATE; }
	McbBucketState &bs = ctx->bs;
	bs = McbBucketState();
	bs.active = true;
	bs.cur = ctx->d_elemA.as<ulonglong2>(); bs.alt = ctx->d_elemB.as<ulonglong2>();
	bs.n_in = ctx->n_local;            // round 1: every read of the slice, invalid ones sort last
	bs.n_valid = ctx->n_valid_round1;
	bs.tot_sk = ctx->n_valid_round1;
	return MCB_OK;
}

// sort + group + consensus of the current tuples; leaves the per-round totals in bs (n_cl_new, n_mem_new, n_sg_new, n_resk)
int mcb_bucket_round_a_impl(mcb_ctx *ctx, int r, int is_last)
{
	McbBucketState &bs = ctx->bs;
	if (!bs.active || bs.half) { mcb_set_error("bucket round: call order violated"); return MCB_ESTATE; }
	BucketBufs B(ctx);
	const int L = ctx->L, WS = ctx->WS, k = ctx->prm.k, max_rounds = ctx->prm.max_rounds;
	const int pbase = L + max_rounds, kmer = k - r;
	const uint64_t N = ctx->n_reads;
	unsigned long long *dc = ctx->d_counters.as<unsigned long long>();
	unsigned long long *hc = nullptr;
	bs.r = r; bs.is_last = is_last;
	bs.n = bs.G = bs.n_cl_new = bs.n_mem_new = bs.n_sg_new = bs.n_ref_new = bs.n_resk = 0;
	bs.half = true;
	if (bs.n_valid == 0) return MCB_OK;
	const int NC = consensus_cols(ctx);
	const size_t cs_per_warp = (((size_t)NC * 16 + NC + CS_MB * 8 * 8 + CS_MB * 16) + 15) & ~(size_t)15;
	const size_t cs_smem = cs_per_warp * CS_WARPS;
	if (cs_smem > 48 * 1024) MCB_CUDA(cudaFuncSetAttribute(k_consensus, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cs_smem));
	// the ASCII buffer is dead after mcb_for_reads: reuse it for consensus strings before compaction
	// (room for the 32-byte aligned slots of k_consensus_bs: a group of g reads takes at most g*L + 95 bytes)
	MCB_TRY(ctx->d_ascii.ensure((size_t)std::max<uint64_t>(bs.n_valid, ctx->n_local) * (L + 48) + 256));
	const uint64_t reftmp_cap = ctx->d_ascii.cap;
	McbSpan sp(ctx->tm, "for_bucket");
	// ---- K2: one stable sort; key = (bucket, minimizer, adjusted pos desc, rid asc)
	std::vector<McbSortPass> passes;
	if (r > 1 || ctx->shard_n > 1) mcb_add_bit_passes(passes, 1, 0, 1 + mcb_bits_for(N));      // strand + rid (a single GPU's round 1 arrives in rid order)
	mcb_add_bit_passes(passes, 1, MCB_POSINV_SHIFT, MCB_POSINV_SHIFT + mcb_bits_for(pbase));
	const int kbits = r == 1 ? 2 * k : 2 * (kmer + 1);                      // tuples of round r were hashed with k-(r-1) (r>=2) or k
	mcb_add_bit_passes(passes, 0, 0, std::max(1, kbits - 14));
	mcb_add_bit_passes(passes, 0, 50, 64);
	// Buckets are brought together by two radix passes on the bucket bits and sorted by the rest of the key inside shared
	// memory (mcb_bucket_sort); a bucket too large for that makes the whole round fall back to LSD passes over the full key.
	static const bool lsd_only = getenv("MCB_SORT_LSD") != nullptr;
	ulonglong2 *sorted = nullptr;
	const uint64_t n = bs.n_valid;
	bs.n = n;
	MCB_TRY(B.b_hs.ensure(std::max<uint64_t>(n, n / 256 + 16400) * 4 + 16)); MCB_TRY(B.b_gs.ensure((n + 2) * 4));
	unsigned long long *hc0 = nullptr;
	for (int attempt = lsd_only ? 1 : 0;; ++attempt) {
		if (attempt == 0) {
			MCB_CUDA(cudaMemsetAsync(&dc[CT_SORT_OVERFLOW], 0, 8, ctx->stream));
			MCB_TRY(mcb_bucket_sort(ctx, bs.cur, bs.alt, bs.n_in, bs.n_valid, kbits, L - (r == 1 ? k : kmer + 1) + 1, B.b_hs.as<uint32_t>(), &dc[CT_SORT_OVERFLOW], &sorted));     // b_hs doubles as the bucket-offset scratch
		} else {
			if (ctx->tm.enabled) { const int id = ctx->tm.id("sort_lsd_fallbacks"); ctx->tm.ms[id] += 1; ctx->tm.cnt[id] += 1; }      // (a count, not milliseconds)
			if (getenv("MCB_SORT_DEBUG")) {          // which sub-buckets were too large?
				std::vector<uint32_t> hb(70000);
				cudaMemcpy(hb.data(), B.b_hs.p, hb.size() * 4, cudaMemcpyDeviceToHost);
				uint32_t mx = 0, at = 0, bad = 0;
				for (size_t i = 0; i + 1 < hb.size(); ++i) { const uint32_t d = hb[i + 1] - hb[i]; if (d > 2048 && d < 0x80000000u) { if (d > mx) { mx = d; at = (uint32_t)i; } ++bad; } }
				fprintf(stderr, "[mcb rank %d] round %d: bucket sort overflowed (%llu sub-buckets reported): n_in=%llu n_valid=%llu largest %u at sub-bucket %u (%u above 2048 among the first 70000); boff[0..3]=%u %u %u %u\n",
				        ctx->shard_rank, r, hc0[CT_SORT_OVERFLOW], (unsigned long long)bs.n_in, (unsigned long long)bs.n_valid, mx, at, bad, hb[0], hb[1], hb[2], hb[3]);
			}
			MCB_TRY(mcb_radix_sort(ctx, bs.cur, bs.alt, bs.n_in, passes.data(), (int)passes.size(), &sorted));
		}
		if (sorted != bs.cur) { bs.alt = bs.cur; bs.cur = sorted; }
		// ---- groups
		MCB_LAUNCH(ctx, "heads", k_heads, mcb_grid_for(n, 256), 256, 0, bs.cur, n, B.b_hs.as<uint32_t>());
		MCB_TRY(mcb_exclusive_scan_u32(ctx, B.b_hs.as<uint32_t>(), n, (uint64_t*)&dc[CT_G]));
		MCB_LAUNCH(ctx, "gstart", k_gstart, mcb_grid_for(n, 256), 256, 0, bs.cur, n, B.b_hs.as<uint32_t>(), &dc[CT_G], B.b_gs.as<uint32_t>());
		MCB_TRY(sync_counters(ctx, &hc0));
		if (attempt > 0 || hc0[CT_SORT_OVERFLOW] == 0) break;
	}
	hc = hc0;
	ulonglong2 *cur = bs.cur;
	const uint64_t G = hc[CT_G];
	bs.G = G;
	// ---- K3 consensus
	MCB_TRY(B.b_st.ensure(n + 16)); MCB_TRY(B.b_er.ensure(n * 4 + 16)); MCB_TRY(B.b_nr.ensure(n * 8 + 16));
	MCB_TRY(B.b_gc.ensure(G * 4 + 16)); MCB_TRY(B.b_gk.ensure(G * 4 + 16)); MCB_TRY(B.b_gsg.ensure(G * 4 + 16)); MCB_TRY(B.b_grk.ensure(G * 4 + 16));
	MCB_TRY(B.b_grl.ensure(G * 8 + 16)); MCB_TRY(B.b_gro.ensure(G * 8 + 16));
	MCB_CUDA(cudaMemsetAsync(&dc[CT_REFCURSOR], 0, 8, ctx->stream));
	ConsOut co; co.status = B.b_st.as<uint8_t>(); co.erank = B.b_er.as<uint32_t>(); co.newrec = B.b_nr.as<uint64_t>();
	co.g_iscl = B.b_gc.as<uint32_t>(); co.g_kept = B.b_gk.as<uint32_t>(); co.g_sg = B.b_gsg.as<uint32_t>(); co.g_resk = B.b_grk.as<uint32_t>();
	co.g_reflen = B.b_grl.as<unsigned long long>(); co.g_refoff = B.b_gro.as<unsigned long long>(); co.reftmp = ctx->d_ascii.as<char>();
	{
		// widest pile-up a group of this round can have: offsets are differences of strand-adjusted positions in [kmer', L+r-2]
		// (kmer' = length of the k-mers these tuples were sketched with), so columns <= 2L + 2r - 2 - k  (+ margin)
		const int max_cols = 2 * L + 2 * r - k + 2;
		const int ncw = (max_cols + 31) / 32;
		MCB_TRY(ctx->d_x[2].ensure((size_t)CONS_NCLS * G * 4 + 16));
		uint32_t *worklist = ctx->d_x[2].as<uint32_t>();
		MCB_CUDA(cudaMemsetAsync(&dc[CT_WORK0], 0, CONS_NCLS * 8, ctx->stream));
		MCB_LAUNCH(ctx, "cons_worklist", k_cons_worklist, mcb_grid_for(G, 256), 256, 0, B.b_gs.as<uint32_t>(), G, co, worklist, dc);
		const unsigned bgrid = (unsigned)ctx->sm_count * 16;
#define CONS_BS_LAUNCH(GWv, Kv, cls) MCB_LAUNCH(ctx, "consensus", (k_consensus_bs<GWv, Kv>), bgrid, 128, 0, cur, B.b_gs.as<uint32_t>(), worklist + (uint64_t)(cls) * G, \
		&dc[CT_WORK0 + (cls)], ctx->d_packed.as<uint64_t>(), WS, ctx->Wd, L, ctx->prm.diff_threshold, is_last, co, dc, reftmp_cap)
		for (int cls = 0; cls < 7; ++cls) {
			if (ncw <= 8) CONS_BS_LAUNCH(8, 4, cls); else if (ncw <= 16) CONS_BS_LAUNCH(16, 4, cls); else CONS_BS_LAUNCH(32, 4, cls);
		}
		if (ncw <= 8) { CONS_BS_LAUNCH(8, 8, 7); CONS_BS_LAUNCH(8, 16, 8); }
		else if (ncw <= 16) { CONS_BS_LAUNCH(16, 8, 7); CONS_BS_LAUNCH(16, 16, 8); }
		else { CONS_BS_LAUNCH(32, 8, 7); CONS_BS_LAUNCH(32, 16, 8); }
#undef CONS_BS_LAUNCH
		// groups too large for the bit-sliced counters (more than CONS_BS_MAX_MEMBERS members): the column-count kernel
		MCB_LAUNCH(ctx, "consensus_huge", k_consensus, (unsigned)ctx->sm_count, CS_WARPS * 32, cs_smem, cur, B.b_gs.as<uint32_t>(), G, ctx->d_packed.as<uint64_t>(), WS, L,
		           ctx->prm.diff_threshold, pbase, NC, is_last, co, dc, reftmp_cap, worklist + 9ull * G, &dc[CT_WORK0 + 9]);
	}
	// ---- bases
	MCB_TRY(mcb_exclusive_scan_u32(ctx, B.b_gc.as<uint32_t>(), G, (uint64_t*)&dc[CT_TOT_CL]));
	MCB_TRY(mcb_exclusive_scan_u32(ctx, B.b_gk.as<uint32_t>(), G, (uint64_t*)&dc[CT_TOT_MEM]));
	MCB_TRY(mcb_exclusive_scan_u32(ctx, B.b_gsg.as<uint32_t>(), G, (uint64_t*)&dc[CT_TOT_SG]));
	MCB_TRY(mcb_exclusive_scan_u32(ctx, B.b_grk.as<uint32_t>(), G, (uint64_t*)&dc[CT_TOT_RESK]));
	MCB_TRY(mcb_exclusive_scan_u64(ctx, B.b_grl.as<uint64_t>(), G, (uint64_t*)&dc[CT_TOT_REF]));
	MCB_TRY(sync_counters(ctx, &hc));
	if (hc[CT_ERR]) { mcb_set_error("internal: consensus table overflow (%llu groups)", hc[CT_ERR]); return MCB_EINVAL; }
	bs.n_cl_new = hc[CT_TOT_CL]; bs.n_mem_new = hc[CT_TOT_MEM]; bs.n_resk = hc[CT_TOT_RESK];
	bs.n_sg_new = hc[CT_TOT_SG]; bs.n_ref_new = hc[CT_TOT_REF];
	return MCB_OK;
}

// scatter into the accumulated outputs, index tuples of the new seed contigs (ids cid_first, cid_first+1, ...), re-sketch of the rejects
int mcb_bucket_round_b_impl(mcb_ctx *ctx, uint64_t cid_first)
{
	McbBucketState &bs = ctx->bs;
	if (!bs.active || !bs.half) { mcb_set_error("bucket round: call order violated"); return MCB_ESTATE; }
	bs.half = false;
	BucketBufs B(ctx);
	const int L = ctx->L, WS = ctx->WS, k = ctx->prm.k, m = ctx->prm.first_mininum, max_rounds = ctx->prm.max_rounds;
	const int pbase = L + max_rounds, kmer = k - bs.r, is_last = bs.is_last;
	unsigned long long *dc = ctx->d_counters.as<unsigned long long>();
	const uint64_t n = bs.n, G = bs.G, n_cl_new = bs.n_cl_new, n_mem_new = bs.n_mem_new, n_resk = bs.n_resk, n_sg_new = bs.n_sg_new, n_ref_new = bs.n_ref_new;
	const uint64_t tot_cl = bs.tot_cl, tot_mem = bs.tot_mem, tot_ref = bs.tot_ref, tot_sg = bs.tot_sg;
	bs.round_cl.push_back(n_cl_new); bs.round_mem.push_back(n_mem_new); bs.round_ref.push_back(n_ref_new); bs.round_sg.push_back(n_sg_new);
	if (n == 0) { bs.n_in = bs.n_valid = 0; return MCB_OK; }
	ulonglong2 *cur = bs.cur;
	{
		McbSpan sp(ctx->tm, "for_bucket");
		MCB_TRY(grow_preserve(ctx, B.d_sg, tot_sg * 4, (tot_sg + n_sg_new) * 4 + 16));
		MCB_TRY(grow_preserve(ctx, B.d_cl_n, tot_cl * 4, (tot_cl + n_cl_new) * 4 + 16));
		MCB_TRY(grow_preserve(ctx, B.d_cl_aoff, tot_cl * 8, (tot_cl + n_cl_new + 1) * 8 + 16));
		MCB_TRY(grow_preserve(ctx, B.d_cl_roff, tot_cl * 8, (tot_cl + n_cl_new + 1) * 8 + 16));
		MCB_TRY(grow_preserve(ctx, B.d_cl_a, tot_mem * 8, (tot_mem + n_mem_new) * 8 + 16));
		MCB_TRY(grow_preserve(ctx, B.d_cl_ref, tot_ref, tot_ref + n_ref_new + 16));
		MCB_TRY(B.b_rk.ensure(n_resk * 4 + 16));
		ScatIn si; si.status = B.b_st.as<uint8_t>(); si.erank = B.b_er.as<uint32_t>(); si.newrec = B.b_nr.as<uint64_t>(); si.hscan = B.b_hs.as<uint32_t>();
		si.g_iscl = B.b_gc.as<uint32_t>(); si.g_kept = B.b_gk.as<uint32_t>(); si.g_sg = B.b_gsg.as<uint32_t>(); si.g_resk = B.b_grk.as<uint32_t>();
		MCB_LAUNCH(ctx, "scatter_members", k_scatter_members, mcb_grid_for(n, 256), 256, 0, cur, n, si, is_last, tot_mem, tot_sg,
		           B.d_cl_a.as<uint64_t>(), B.d_sg.as<uint32_t>(), B.b_rk.as<uint32_t>());
		MCB_LAUNCH(ctx, "scatter_clusters", k_scatter_clusters, mcb_grid_for(G * 32, 256), 256, 0, B.b_gs.as<uint32_t>(), G, B.b_gc.as<uint32_t>(), B.b_gk.as<uint32_t>(),
		           B.b_grl.as<unsigned long long>(), B.b_gro.as<unsigned long long>(), dc, ctx->d_ascii.as<char>(), tot_cl, tot_mem, tot_ref,
		           B.d_cl_n.as<uint32_t>(), B.d_cl_aoff.as<uint64_t>(), B.d_cl_roff.as<uint64_t>(), B.d_cl_ref.as<char>());
		// close the offset tables
		{
			uint64_t endv[2] = { tot_mem + n_mem_new, tot_ref + n_ref_new };
			MCB_CUDA(cudaMemcpyAsync(B.d_cl_aoff.as<uint64_t>() + tot_cl + n_cl_new, &endv[0], 8, cudaMemcpyHostToDevice, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(B.d_cl_roff.as<uint64_t>() + tot_cl + n_cl_new, &endv[1], 8, cudaMemcpyHostToDevice, ctx->stream));
			MCB_CUDA(cudaStreamSynchronize(ctx->stream));
		}
		// ---- K4: first m minimizers of each new seed contig (window rw, k = reads->k; kthread_bucket.c:458)
		MCB_TRY(grow_preserve(ctx, B.d_mi, tot_cl * m * 16, (tot_cl + n_cl_new) * m * 16 + 16));
		MCB_TRY(grow_preserve(ctx, B.d_micnt, tot_cl, tot_cl + n_cl_new + 16));
		if (n_cl_new) {
			const int rw = ctx->prm.rw;                                   // mcb_create admits rw <= MCB_LH_WMAX = LH2_WMAX
			const size_t lh2_smem = (size_t)rw * LH_THREADS * 13;
			auto kern = (k > 16 && k < 32) ? k_sketch_lh2<true> : k_sketch_lh2<false>;
			if (lh2_smem > 48 * 1024) MCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lh2_smem));
			MCB_LAUNCH(ctx, "sketch_lh", kern, mcb_grid_for(n_cl_new, LH_THREADS), LH_THREADS, lh2_smem, B.d_cl_ref.as<char>(), B.d_cl_roff.as<uint64_t>(), tot_cl, n_cl_new, cid_first,
			           rw, k, m, B.d_mi.as<mcb_tuple>(), B.d_micnt.as<uint8_t>(), (const uint64_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (const uint32_t*)nullptr, 0, 0);
		}
		bs.tot_cl += n_cl_new; bs.tot_mem += n_mem_new; bs.tot_ref += n_ref_new; bs.tot_sg += n_sg_new;
		// ---- rejected reads go to the next round with a shorter k-mer (kthread_bucket.c:205-212,488-496)
		if (!is_last && n_resk) {
			MCB_LAUNCH(ctx, "resketch", k_resketch, mcb_grid_for(n_resk, 128), 128, 0, ctx->d_packed.as<uint64_t>(), WS, L, k, kmer, mcb_ta_mul(kmer), pbase,
			           B.b_rk.as<uint32_t>(), n_resk, bs.alt, dc);
			ulonglong2 *t = bs.cur; bs.cur = bs.alt; bs.alt = t;
			bs.tot_sk += n_resk;
		}
	}
	bs.n_in = bs.n_valid = (is_last ? 0 : n_resk);
	return MCB_OK;
}

int mcb_bucket_finish_impl(mcb_ctx *ctx, mcb_bucket_result *res, uint64_t *round_counts, int cap_rounds)
{
	{
		const McbBucketState &b0 = ctx->bs;
		const int nr = (int)b0.round_cl.size();
		if (round_counts) {
			if (cap_rounds < nr) { mcb_set_error("kt_for_bucket: %d rounds, room for %d", nr, cap_rounds); return MCB_EINVAL; }
			for (int i = 0; i < nr; ++i) { round_counts[4 * i] = b0.round_cl[i]; round_counts[4 * i + 1] = b0.round_mem[i]; round_counts[4 * i + 2] = b0.round_ref[i]; round_counts[4 * i + 3] = b0.round_sg[i]; }
		}
	}
	McbBucketState &bs = ctx->bs;
	if (!bs.active || bs.half) { mcb_set_error("bucket finish: call order violated"); return MCB_ESTATE; }
	BucketBufs B(ctx);
	const int m = ctx->prm.first_mininum;
	unsigned long long *hc = nullptr;
	MCB_TRY(sync_counters(ctx, &hc));
	if (hc[CT_DEGENERATE]) { mcb_set_error("%llu re-sketched reads have no valid k-mer", hc[CT_DEGENERATE]); return MCB_EINPUT; }
	const uint64_t tot_cl = bs.tot_cl, tot_mem = bs.tot_mem, tot_ref = bs.tot_ref, tot_sg = bs.tot_sg;
	// ---- results to the host
	MCB_TRY(ctx->h_sg.ensure(tot_sg * 4 + 16)); MCB_TRY(ctx->h_cl_a_off.ensure(16)); MCB_TRY(ctx->h_cl_ref_off.ensure(16));
	if (!ctx->seed_results_on_device_only) {
		MCB_TRY(ctx->h_cl_n.ensure(tot_cl * 4 + 16)); MCB_TRY(ctx->h_cl_a_off.ensure((tot_cl + 1) * 8)); MCB_TRY(ctx->h_cl_ref_off.ensure((tot_cl + 1) * 8));
		MCB_TRY(ctx->h_cl_a.ensure(tot_mem * 8 + 16)); MCB_TRY(ctx->h_cl_ref.ensure(tot_ref + 16));
		MCB_TRY(ctx->h_mi.ensure(tot_cl * m * 16 + 16)); MCB_TRY(ctx->h_mi_cnt.ensure(tot_cl + 16));
	}
	{
		McbSpan sp(ctx->tm, "d2h");
		if (tot_cl && !ctx->seed_results_on_device_only) {
			MCB_CUDA(cudaMemcpyAsync(ctx->h_cl_n.p, B.d_cl_n.p, tot_cl * 4, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(ctx->h_cl_a_off.p, B.d_cl_aoff.p, (tot_cl + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(ctx->h_cl_ref_off.p, B.d_cl_roff.p, (tot_cl + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(ctx->h_cl_a.p, B.d_cl_a.p, tot_mem * 8, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(ctx->h_cl_ref.p, B.d_cl_ref.p, tot_ref, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(ctx->h_mi.p, B.d_mi.p, tot_cl * m * 16, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(ctx->h_mi_cnt.p, B.d_micnt.p, tot_cl, cudaMemcpyDeviceToHost, ctx->stream));
		} else { ctx->h_cl_a_off.as<uint64_t>()[0] = 0; ctx->h_cl_ref_off.as<uint64_t>()[0] = 0; }
		if (tot_sg) MCB_CUDA(cudaMemcpyAsync(ctx->h_sg.p, B.d_sg.p, tot_sg * 4, cudaMemcpyDeviceToHost, ctx->stream));
	}
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->tm.collect();
	res->n_clusters = tot_cl; res->cl_n = ctx->h_cl_n.as<uint32_t>(); res->cl_a_off = ctx->h_cl_a_off.as<uint64_t>(); res->cl_a = ctx->h_cl_a.as<uint64_t>();
	res->cl_ref_off = ctx->h_cl_ref_off.as<uint64_t>(); res->cl_ref = ctx->h_cl_ref.as<char>();
	if (ctx->seed_results_on_device_only) { res->cl_n = nullptr; res->cl_a_off = nullptr; res->cl_a = nullptr; res->cl_ref_off = nullptr; res->cl_ref = nullptr; }
	res->n_sg = tot_sg; res->sg = ctx->h_sg.as<uint32_t>();
	res->mi_cnt = ctx->seed_results_on_device_only ? nullptr : ctx->h_mi_cnt.as<uint8_t>(); res->mi = ctx->seed_results_on_device_only ? nullptr : ctx->h_mi.as<mcb_tuple>();
	res->rounds = bs.r; res->n_sketched_total = bs.tot_sk; res->n_grouped = tot_mem;
	bs.active = false;
	ctx->bucket_done = true;
	return MCB_OK;
}

// loop control of kt_for_bucket (kthread_bucket.c:584-585,594,607-622): feed it the members gained in the round
extern "C" void mcb_round_control_init(mcb_round_control *rc) { memset(rc, 0, sizeof *rc); }
extern "C" int mcb_round_control_begin(mcb_round_control *rc, int k, int max_rounds)
{
	rc->round += 1;
	if (k - rc->round <= 9) ++rc->last_rounds;
	if (rc->round == max_rounds - 1) ++rc->last_rounds;
	return rc->last_rounds ? 1 : 0;                    // is_last for this round
}
extern "C" int mcb_round_control_end(mcb_round_control *rc, uint64_t members_total)
{
	if (rc->last_rounds) ++rc->last_rounds;
	if ((long long)members_total - rc->pre_members < 100) ++rc->last_rounds;
	rc->pre_members = (long long)members_total;
	return rc->last_rounds > 1 ? 1 : 0;                // stop
}

// kt_for_bucket for a caller that goes on with mcb_combine: the seed contigs and their index tuples stay on the device, only the
// singles (and the counts) come back; the array pointers of *res other than sg are NULL
extern "C" int mcb_for_bucket_keep(mcb_ctx *ctx, mcb_bucket_result *res)
{
	if (!ctx) { mcb_set_error("null context"); return MCB_EINVAL; }
	ctx->seed_results_on_device_only = true;
	const int rc = mcb_for_bucket(ctx, res);
	ctx->seed_results_on_device_only = false;
	return rc;
}

extern "C" int mcb_for_bucket(mcb_ctx *ctx, mcb_bucket_result *res)
{
	MCB_TRY(check_ctx(ctx));
	if (!res) { mcb_set_error("mcb_for_bucket: null result"); return MCB_EINVAL; }
	if (ctx->shard_n > 1) { mcb_set_error("mcb_for_bucket: context is sharded, use mcb_shard_for_bucket"); return MCB_ESTATE; }
	MCB_TRY(mcb_bucket_begin(ctx));
	mcb_round_control rc; mcb_round_control_init(&rc);
	for (;;) {
		const int is_last = mcb_round_control_begin(&rc, ctx->prm.k, ctx->prm.max_rounds);
		MCB_TRY(mcb_bucket_round_a_impl(ctx, rc.round, is_last));
		MCB_TRY(mcb_bucket_round_b_impl(ctx, ctx->bs.tot_cl));
		if (mcb_round_control_end(&rc, ctx->bs.tot_mem)) break;
	}
	return mcb_bucket_finish_impl(ctx, res, nullptr, 0);
}

// room for n_tuples (+1) in both halves of the sort double buffer; the current half (bs.cur, bs.n_in elements) survives
int mcb_elems_reserve(mcb_ctx *ctx, uint64_t n_tuples)
{
	McbBucketState &bs = ctx->bs;
	if (n_tuples + 1 <= ctx->elem_cap) return MCB_OK;
	const bool cur_is_A = bs.cur == ctx->d_elemA.as<ulonglong2>();
	DBuf &keep = cur_is_A ? ctx->d_elemA : ctx->d_elemB, &other = cur_is_A ? ctx->d_elemB : ctx->d_elemA;
	MCB_TRY(other.ensure((n_tuples + 1) * 16 + 16));
	MCB_TRY(grow_preserve(ctx, keep, bs.n_in * 16, (n_tuples + 1) * 16 + 16));
	ctx->elem_cap = std::min(ctx->d_elemA.cap, ctx->d_elemB.cap) / 16;
	bs.cur = keep.as<ulonglong2>(); bs.alt = other.as<ulonglong2>();
	return MCB_OK;
}
