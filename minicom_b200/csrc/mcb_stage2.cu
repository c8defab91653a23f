// mcb_stage2.cu — realign_hash (kthread_hash_realign.c:569-594) on the device.
//
// The reference builds, per threshold round, `numdict` dictionaries over substrings of the leftover reads ("singles",
// constructdictionary_realign :3-140) and slides every contig window over them (realign_hash_search :316-508):
// W windows x (2*numdict-1) random dictionary probes, of which ~98 % miss.  The join is symmetric, and the contigs do not
// change between rounds (preprocess.c:197-232), so this implementation turns it around:
//
//   once per contig set   K5a k_s2_pack_refs    2-bit packed contigs (+ one 32-byte record per contig, k_s2_contig_meta)
//                         K5b k_s2_kmer_emit    every lt-mer of every contig (lt = 17, or 11 for L <= 80) as an 8-byte entry
//                                               lt-mer (low 32 bits)<<32 | position, radix-sorted by a hash of those bits into 2^p buckets
//                                               (mcb_radix_sort_kmers) + k_s2_bucket_ends
//   every round           K6  k_s2_singles      singleRead2bitset (bbhashdict.c:127-227): 2-bit singles, near-poly-A/T
//                                               diversion, and a count-min sketch of the dictionary bins (bin sizes matter:
//                                               the reference scans only the last `maxsearch` entries of a bin, :388)
//                         K7a k_s2_probe_table  one thread per (single, dictionary): its substring key, and the reverse
//                                               complement of the key, are looked up in the contig table; every entry with
//                                               an equal lt-mer is a (window, single) candidate = exactly the pairs the
//                                               reference's probes meet
//                         K7b k_s2_verify       one thread per candidate: XOR/popcount verification and the encode_byte
//                                               gate (:283-314); the single-threaded reference lets the FIRST probe step that
//                                               matches a read claim it, so every verified pair proposes priority
//                                               P=(window, phase, l) and atomicMin keeps the first
//                         K8  claims            compacted and sorted by (P, sg index descending) = the reference's append order
//   rare                  realign_exact_bins    bins larger than maxsearch: their singles are replayed in order on the host
//
// Probes per round drop from 9 per window (4.7e8 at 10 M reads) to 2 per (single, dictionary) (3e7 in the first round,
// 3e6 later), and only true key matches touch the contigs.  Sharded over G GPUs (mcb_shard_realign, mcb_shard.cu) every rank
// runs this same path on the singles IT produced in Stage 1 (their packed rows are local) against all contigs — packed contigs
// are a few bytes per read — so no claim ever crosses ranks; only the dictionary bin-size guard is a collective.
//
// Distance is the popcount of the XOR of 2-bit codes (bbhashdict.c:247-254), not a base count: A<->T and C<->G cost 2.
// The reference's code A=00,G=01,C=10,T=11 (kthread_hash_realign.c:251-258) and ours (A0 C1 G2 T3) differ only by
// swapping the two bits of a field, so popcounts, key equality and bin contents are identical.
#include "mcb_common.cuh"
#include <algorithm>
#include <stdlib.h>

#define S2_MAXD 16
#define S2_POS_BITS MCB_S2_POS_BITS
#define S2_POS_MASK 0xFFFFFFFFull
#define S2_KEY32(k) ((uint64_t)(k) & 0xFFFFFFFFull)      // what a table entry keeps of an lt-mer
#define S2_BLK_SHIFT 9
struct S2Geom {
	int L, Wd, WS, nd, lt;
	int dstart[S2_MAXD];
	int enc_limit;           // floor(0.4*L)  (readlen*0.4, kthread_hash_realign.c:313, bbhashdict.c:177)
	int thr, maxsearch;
};

__device__ __forceinline__ int ndigits(int v) { return v >= 100 ? 3 : v >= 10 ? 2 : 1; }

// bits [2*base, 2*(base+nbases)) of a packed row
__device__ __forceinline__ uint64_t extract_bases(const uint64_t *w, int base, int nbases)
{
	int bit = 2 * base, wi = bit >> 6, sh = bit & 63;
	uint64_t v = w[wi] >> sh;
	if (sh + 2 * nbases > 64) v |= w[wi + 1] << (64 - sh);
	return v & ((1ull << (2 * nbases)) - 1);
}

__device__ __forceinline__ uint64_t mix64(uint64_t k)
{
	k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
	return k;
}
__device__ __forceinline__ uint32_t kmer_bucket(uint64_t key, int pbits) { return mcb_kmer_bucket(key, pbits); }

// encode_byte (kthread_hash_realign.c:283-314) on a mismatch pattern given as XOR words, positions ascending.
// Reproduces the missing `eq_char_num = 0` of the literal branch (:301-305).
__device__ bool enc_ok(const uint64_t *x, int L, int limit)
{
	int len = 0, eq = 0;
	for (int i = 0; i < L; ++i) {
		bool mism = ((x[i >> 5] >> (2 * (i & 31))) & 3) != 0;
		if (mism) {
			if (eq > 1) { len += ndigits(eq); eq = 0; }
			else len += eq;
			++len;
		} else ++eq;
	}
	if (len == 0) len = 1;
	return len <= limit;
}

// ---------------------------------------------------------------- K5a contig packing
// contig c occupies words [cw_off[c], cw_off[c+1]) (ceil(len/32)+1 words, zero padded)
__global__ void k_s2_pack_refs(const char *__restrict__ refs, const uint64_t *__restrict__ ref_off, const uint64_t *__restrict__ cw_off, uint64_t n_contigs,
                               uint64_t w_begin, uint64_t w_end, uint64_t *__restrict__ cw, unsigned long long *__restrict__ counters)
{
	const uint64_t q = w_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;      // words [w_begin, w_end): all of them, or this rank's share
	if (q >= w_end) return;
	uint64_t lo = 0, hi = n_contigs;          // last c with cw_off[c] <= q
	while (hi - lo > 1) { uint64_t mid = (lo + hi) >> 1; if (cw_off[mid] <= q) lo = mid; else hi = mid; }
	const uint64_t c = lo, wq = q - cw_off[c];
	const uint64_t b = ref_off[c], len = ref_off[c + 1] - b;
	uint64_t v = 0; bool bad = false;
	for (int j = 0; j < 32; ++j) {
		uint64_t p = wq * 32 + j;
		if (p < len) { unsigned code = mcb_code_of((unsigned char)refs[b + p]); bad |= code > 3; v |= (uint64_t)(code & 3) << (2 * j); }
	}
	if (bad) atomicAdd(&counters[CT_S2_ERR], 1ull);
	cw[q] = v;
}

// contig that holds base 512*i of the concatenated consensus strings
__global__ void k_s2_pos_blocks(const uint64_t *__restrict__ ref_off, uint64_t n_contigs, uint64_t n_blocks, uint32_t *__restrict__ pblk)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_blocks) return;
	const uint64_t P = i << S2_BLK_SHIFT;
	uint64_t lo = 0, hi = n_contigs;          // last c with ref_off[c] <= P
	while (hi - lo > 1) { uint64_t mid = (lo + hi) >> 1; if (ref_off[mid] <= P) lo = mid; else hi = mid; }
	pblk[i] = (uint32_t)lo;
}

// one 32-byte record per contig, so that a candidate costs one sector for all its contig coordinates
struct __align__(32) S2ContigMeta { uint64_t ref_off, cw_off, woff, len; };
__global__ void k_s2_contig_meta(const uint64_t *__restrict__ ref_off, const uint64_t *__restrict__ cw_off, const uint64_t *__restrict__ woff, uint64_t n_contigs,
                                 S2ContigMeta *__restrict__ meta)
{
	uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c > n_contigs) return;
	S2ContigMeta m;
	m.ref_off = ref_off[c]; m.cw_off = cw_off[c]; m.woff = woff[c]; m.len = c < n_contigs ? ref_off[c + 1] - ref_off[c] : ~0ull >> 1;   // sentinel ends every walk
	meta[c] = m;
}

// were the contigs of this call the ones the cached index was built from?
__global__ void k_s2_same(const uint64_t *__restrict__ a, const uint64_t *__restrict__ b, uint64_t nwords, unsigned long long *__restrict__ counters)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	bool diff = i < nwords && a[i] != b[i];
	if (__any_sync(0xFFFFFFFFu, diff) && (threadIdx.x & 31) == 0) atomicAdd(&counters[CT_S2_DIFF], 1ull);
}

// ---------------------------------------------------------------- key filter
// Only contig lt-mers equal to a dictionary key of some single (or to the reverse complement of one) can ever be hit.  A Bloom
// filter over those keys (one 32-bit word, three bits per key) lets the table builder drop the rest: at 20x coverage about one
// contig lt-mer in eight survives, so the table, its sort and the probes' DRAM footprint shrink accordingly.  The filter belongs
// to the singles of the call that built the table; later rounds may only use the table if their singles are a subset (sgmap).
__device__ __forceinline__ void flt_slot(uint64_t key, uint64_t wmask, uint64_t *word, uint32_t *bits)
{
	const uint64_t h = mix64(key);
	*word = (h >> 20) & wmask;
	*bits = (1u << (h & 31)) | (1u << ((h >> 5) & 31)) | (1u << ((h >> 10) & 31));
}
__device__ __forceinline__ bool flt_has(const uint32_t *__restrict__ flt, uint64_t wmask, uint64_t key)
{
	uint64_t w; uint32_t b;
	flt_slot(key, wmask, &w, &b);
	return (flt[w] & b) == b;
}
__device__ __forceinline__ uint64_t rev_fields(uint64_t v, int nbases)
{
	// reverse the order of the low `nbases` 2-bit fields
	uint64_t r = __brevll(v) >> (64 - 2 * nbases);
	return ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
}
__global__ void k_s2_filter_insert(const uint64_t *__restrict__ rd, const uint8_t *__restrict__ flagged, const uint32_t *__restrict__ sg, uint64_t S, S2Geom gm,
                                   uint32_t *__restrict__ flt, uint64_t wmask, uint32_t *__restrict__ sgmap)
{
	const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= S * (uint64_t)gm.nd) return;
	const int l = (int)(idx / S); const uint64_t s = idx - (uint64_t)l * S;
	if (l == 0) { const uint32_t rid = sg[s]; atomicOr(&sgmap[rid >> 5], 1u << (rid & 31)); }
	(void)flagged;          // diverted singles are inserted too: whether a single is diverted depends on the round's threshold
	const uint64_t *row = rd + s * gm.WS;
	const int lt = gm.lt, ds = gm.dstart[l];
	const uint64_t kmask = (1ull << (2 * lt)) - 1;
	const int bit = 2 * ds, wi = bit >> 6, sh = bit & 63;
	uint64_t v = row[wi] >> sh;
	if (sh + 2 * lt > 64) v |= row[wi + 1] << (64 - sh);
	const uint64_t key_f = v & kmask;
	uint64_t w; uint32_t b;
	flt_slot(key_f, wmask, &w, &b); atomicOr(&flt[w], b);
	if (ds > 0) {
		const uint64_t key_r = rev_fields(~key_f & kmask, lt);
		flt_slot(key_r, wmask, &w, &b); atomicOr(&flt[w], b);
	}
}
// are all these singles among the ones the filter was built from?
__global__ void k_s2_sg_subset(const uint32_t *__restrict__ sg, uint64_t S, const uint32_t *__restrict__ sgmap, unsigned long long *__restrict__ counters)
{
	const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	bool miss = false;
	if (s < S) { const uint32_t rid = sg[s]; miss = !((sgmap[rid >> 5] >> (rid & 31)) & 1u); }
	if (__any_sync(0xFFFFFFFFu, miss) && (threadIdx.x & 31) == 0) atomicAdd(&counters[CT_S2_DIFF], 1ull);
}

// ---------------------------------------------------------------- K5b contig lt-mer table
// One thread per packed word finds its contig; the warp then walks its 32 words together, lane j writing the lt-mer that
// starts at base j of the word, so every store instruction covers 32 consecutive entries.
__global__ void __launch_bounds__(256)
k_s2_kmer_emit(const uint64_t *__restrict__ cw, const uint64_t *__restrict__ cw_off, const uint64_t *__restrict__ ref_off, const uint64_t *__restrict__ ent_off,
               uint64_t n_contigs, uint64_t total_words, int L, int lt, unsigned long long *__restrict__ ents,
               int filter, unsigned long long *__restrict__ counter, unsigned long long ents_cap,
               const uint32_t *__restrict__ flt, uint64_t flt_mask)
{
	const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int lane = threadIdx.x & 31;
	uint64_t w0 = 0, w1 = 0, dst = 0, pos0 = 0; int nstart = 0;
	if (q < total_words) {
		uint64_t lo = 0, hi = n_contigs;
		while (hi - lo > 1) { uint64_t mid = (lo + hi) >> 1; if (cw_off[mid] <= q) lo = mid; else hi = mid; }
		const uint64_t c = lo, wq = q - cw_off[c];
		const uint64_t base = ref_off[c], len = ref_off[c + 1] - base;
		if (len >= (uint64_t)L && wq * 32 + lt <= len) {                 // contigs shorter than a read have no windows
			w0 = cw[q]; w1 = cw[q + 1];
			nstart = (int)min((uint64_t)32, len - lt + 1 - wq * 32);
			dst = ent_off[c] + wq * 32; pos0 = base + wq * 32;
		}
	}
	const uint64_t kmask = (1ull << (2 * lt)) - 1;
	if (!filter) {
		for (int t = 0; t < 32; ++t) {
			const int cnt = __shfl_sync(0xFFFFFFFFu, nstart, t);
			if (cnt == 0) continue;
			const uint64_t a = __shfl_sync(0xFFFFFFFFu, w0, t), b = __shfl_sync(0xFFFFFFFFu, w1, t);
			const uint64_t d = __shfl_sync(0xFFFFFFFFu, dst, t), p = __shfl_sync(0xFFFFFFFFu, pos0, t);
			if (lane < cnt) {
				const uint64_t key = ((a >> (2 * lane)) | (lane ? b << (64 - 2 * lane) : 0ull)) & kmask;
				ents[d + lane] = (key << S2_POS_BITS) | (p + lane);
			}
		}
		return;
	}
	// filtered table: keep the lt-mers that pass the key filter.  Each thread walks
	// the 32 start positions of ITS word, the warp reserves room with one atomic, and the kept entries are written packed (their
	// order is irrelevant: they are sorted next).  The keep decisions are remembered in a bit mask for the second sweep.
	unsigned keepmask = 0;
#pragma unroll 4
	for (int j = 0; j < nstart; ++j) {
		const uint64_t key = ((w0 >> (2 * j)) | (j ? w1 << (64 - 2 * j) : 0ull)) & kmask;
		const bool keep = flt_has(flt, flt_mask, key);
		keepmask |= (unsigned)keep << j;
	}
	const unsigned mine = __popc(keepmask);
	unsigned inc = mine;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
	const unsigned total = __shfl_sync(0xFFFFFFFFu, inc, 31);
	if (total == 0) return;
	unsigned long long base = 0;
	if (lane == 0) base = atomicAdd(counter, (unsigned long long)total);
	base = __shfl_sync(0xFFFFFFFFu, base, 0);
	unsigned long long at = base + (inc - mine);
	for (unsigned km = keepmask; km; km &= km - 1) {
		const int j = __ffs(km) - 1;
		const uint64_t key = ((w0 >> (2 * j)) | (j ? w1 << (64 - 2 * j) : 0ull)) & kmask;
		if (at < ents_cap) ents[at] = (key << S2_POS_BITS) | (pos0 + j);
		++at;
	}
}
// entries are ordered by bucket: ptab[b] = end of bucket b (= start of bucket b+1)
__global__ void k_s2_bucket_ends(const unsigned long long *__restrict__ ents, uint64_t n, int pbits, uint32_t b_lo, uint32_t b_hi, uint32_t *__restrict__ ptab)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t b = kmer_bucket(ents[i] >> S2_POS_BITS, pbits) - b_lo;
	const uint32_t bn = i + 1 < n ? kmer_bucket(ents[i + 1] >> S2_POS_BITS, pbits) - b_lo : b_hi - b_lo;
	if (i == 0) for (uint32_t q = 0; q < b; ++q) ptab[q] = 0;
	for (uint32_t q = b; q < bn; ++q) ptab[q] = (uint32_t)(i + 1);
}

// ---------------------------------------------------------------- K6
__global__ void k_s2_singles(const uint32_t *__restrict__ sg, uint64_t S, const uint64_t *__restrict__ packed, S2Geom gm,
                             const uint32_t *__restrict__ nread_rid, const uint64_t *__restrict__ nread_mask, uint64_t n_nreads, uint64_t n_reads,
                             uint64_t *__restrict__ rd, uint8_t *__restrict__ flagged, uint32_t *__restrict__ cm, uint64_t cm_mask,
                             unsigned long long *__restrict__ counters)
{
	uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	unsigned maxbin = 0;
	if (s < S) {
		const uint32_t rid = sg[s];
		if (rid >= n_reads) atomicAdd(&counters[CT_S2_ERR], 1ull);
		else {
			const int L = gm.L, Wd = gm.Wd, WS = gm.WS;
			uint64_t w[9];
#pragma unroll
			for (int i = 0; i < 9; ++i) w[i] = i < Wd ? packed[(uint64_t)rid * WS + i] : 0ull;
			int pcA = 0, pcT = 0;
			for (int i = 0; i < Wd; ++i) {
				uint64_t valid = (i == Wd - 1 && (L & 31)) ? (1ull << (2 * (L & 31))) - 1 : ~0ull;
				pcA += __popcll(w[i]); pcT += __popcll(~w[i] & valid);
			}
			for (int i = 0; i < WS; ++i) rd[s * WS + i] = w[i];
			// near-poly-A / near-poly-T (bbhashdict.c:157-216); the run-length code is measured on the ORIGINAL characters (N restored)
			uint8_t fl = 0;
			const bool nearA = pcA <= gm.thr, nearT = !nearA && pcT <= gm.thr;
			if (nearA || nearT) {
				const uint64_t *nm = nullptr;
				{   // binary search the side table of reads that contained N
					uint64_t lo = 0, hi = n_nreads;
					while (lo < hi) { uint64_t mid = (lo + hi) >> 1; uint32_t v = nread_rid[mid]; if (v < rid) lo = mid + 1; else hi = mid; }
					if (lo < n_nreads && nread_rid[lo] == rid) nm = nread_mask + lo * WS;
				}
				const unsigned want = nearA ? 0u : 3u;
				int len = 0, eq = 0;
				for (int i = 0; i < L; ++i) {
					bool isn = nm ? ((nm[i >> 5] >> (2 * (i & 31))) & 1) != 0 : false;
					bool same = !isn && (((w[i >> 5] >> (2 * (i & 31))) & 3) == want);
					if (!same) { if (eq > 0) { len += ndigits(eq); eq = 0; } ++len; } else ++eq;
				}
				if (len == 0) len = 1;
				if (len <= gm.enc_limit) fl = nearA ? 1 : 2;
			}
			flagged[s] = fl;
			// count-min sketch of the dictionary bins: an upper bound of every bin size (a bin = singles sharing a key in dictionary l)
			for (int l = 0; l < gm.nd; ++l) {
				const uint64_t k0 = extract_bases(w, gm.dstart[l], gm.lt);
				unsigned long long key = ((unsigned long long)l << 34) | k0;
				unsigned v = atomicAdd(&cm[mix64(key) & cm_mask], 1u) + 1u;
				maxbin = max(maxbin, v);
			}
		}
	}
	maxbin = __reduce_max_sync(0xFFFFFFFFu, maxbin);
	if ((threadIdx.x & 31) == 0 && maxbin) atomicMax(&counters[CT_S2_MAXBIN], (unsigned long long)maxbin);
}

// exact bin sizes (only when the sketch reports a possible bin above maxsearch): open addressing (l,key) -> count
#define S2_EMPTY 0xFFFFFFFFFFFFFFFFull
__global__ void k_s2_bins_exact(const uint64_t *__restrict__ rd, uint64_t S, S2Geom gm, unsigned long long *__restrict__ tkey, uint32_t *__restrict__ tcnt, uint64_t hmask)
{
	uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= S * (uint64_t)gm.nd) return;
	const int l = (int)(idx / S); const uint64_t s = idx - (uint64_t)l * S;
	uint64_t w[9];
#pragma unroll
	for (int i = 0; i < 9; ++i) w[i] = i < gm.Wd ? rd[s * gm.WS + i] : 0ull;
	const unsigned long long key = ((unsigned long long)l << 34) | extract_bases(w, gm.dstart[l], gm.lt);
	uint64_t h = mix64(key) & hmask;
	for (;;) {
		unsigned long long old = atomicCAS(&tkey[h], S2_EMPTY, key);
		if (old == S2_EMPTY || old == key) { atomicAdd(&tcnt[h], 1u); return; }
		h = (h + 1) & hmask;
	}
}
__device__ __forceinline__ uint32_t bins_exact_count(const unsigned long long *__restrict__ tkey, const uint32_t *__restrict__ tcnt, uint64_t hmask, unsigned long long key)
{
	uint64_t h = mix64(key) & hmask;
	for (;;) {
		unsigned long long k = tkey[h];
		if (k == key) return tcnt[h];
		if (k == S2_EMPTY) return 0;
		h = (h + 1) & hmask;
	}
}

// exact mode: every (single, dictionary) whose bin is larger than maxsearch -> inbig[single] = 1 and a (bin key, single) record
// (bins are judged by the exact table on one GPU and by the job-wide count-min sketch when sharded; records carry the single's
// position in the job's sg list and its diversion flag, so that the replay can run on any rank)
__global__ void k_s2_mark_big(const uint64_t *__restrict__ rd, const uint8_t *__restrict__ flagged, const uint32_t *__restrict__ sidx, uint64_t S, S2Geom gm,
                              const unsigned long long *__restrict__ tkey, const uint32_t *__restrict__ tcnt, uint64_t hmask, const uint32_t *__restrict__ cmg, uint64_t cmg_mask,
                              uint8_t *__restrict__ inbig, unsigned long long *__restrict__ members, unsigned long long cap, unsigned long long *__restrict__ counters)
{
	uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= S * (uint64_t)gm.nd) return;
	const int l = (int)(idx / S); const uint64_t s = idx - (uint64_t)l * S;
	const uint64_t *row = rd + s * gm.WS;
	const int bit = 2 * gm.dstart[l], wi = bit >> 6, sh = bit & 63;
	uint64_t v = row[wi] >> sh;
	if (sh + 2 * gm.lt > 64) v |= row[wi + 1] << (64 - sh);
	const unsigned long long key = ((unsigned long long)l << 34) | (v & ((1ull << (2 * gm.lt)) - 1));
	if ((tkey ? bins_exact_count(tkey, tcnt, hmask, key) : cmg[mix64(key) & cmg_mask]) <= (uint32_t)gm.maxsearch) return;
	inbig[s] = 1;
	const unsigned long long at = atomicAdd(&counters[CT_S2_NBIGMEM], 1ull);
	if (at < cap) { members[2 * at] = key; members[2 * at + 1] = (unsigned long long)(sidx ? sidx[s] : (uint32_t)s) | ((unsigned long long)flagged[s] << 32); }
}
__global__ void k_s2_apply_claims(const unsigned long long *__restrict__ pairs, uint64_t n, unsigned long long *__restrict__ claim)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) claim[pairs[2 * i]] = pairs[2 * i + 1];
}

// ---------------------------------------------------------------- K7
struct S2Join {
	uint64_t S;
	const uint64_t *rd; const uint8_t *flagged;
	const uint32_t *ptab; const unsigned long long *ents; int pbits;
	const uint32_t *pblk; const S2ContigMeta *meta; const uint64_t *cw;
	unsigned long long *claim;            // [S] min priority
	unsigned long long *counters;
	const unsigned long long *xkey; const uint32_t *xcnt; uint64_t xmask;   // exact bin sizes (null: trust the sketch)
	const uint32_t *cmg; uint64_t cmg_mask;                                 // sharded: count-min sketch summed over the ranks (bounds every bin of the JOB)
	// exact mode (some dictionary bin exceeds maxsearch): singles that sit in such a bin do not claim on the device; their
	// verified matches are listed as events {priority, bin key, single} and replayed in order on the host
	const uint8_t *inbig; unsigned long long *events; unsigned long long events_cap;
	const uint32_t *sidx;                 // position of every single in the job's sg list (null = identity)
};


// K7a: probe.  One thread per (single, dictionary): the key and its reverse complement are looked up in the contig table;
// every entry with an equal lt-mer becomes a candidate record (pair index | phase | contig position), appended with one
// atomic per warp.  No verification here: key matches are sparse (a fraction of a match per thread), and
// verifying them inside this loop left 31 lanes idle around each one.
#define S2_CAND_PHASE_BIT 32                 // candidate record: (single, dictionary) index << 33 | phase << 32 | contig position
__global__ void __launch_bounds__(256) k_s2_probe_table(S2Join p, S2Geom gm, unsigned long long *__restrict__ cand, unsigned long long cand_cap)
{
	const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int lane = threadIdx.x & 31;
	uint64_t key2[2] = { 0, 0 };
	uint32_t lo2[2] = { 0, 0 }, hi2[2] = { 0, 0 };
	if (idx < p.S * (uint64_t)gm.nd) {
		const int l = (int)(idx / p.S); const uint64_t s = idx - (uint64_t)l * p.S;
		if (!p.flagged[s]) {                  // sg_flag already set by the poly-A/T diversion: can never be claimed
			const int lt = gm.lt, ds = gm.dstart[l];
			const uint64_t *__restrict__ row = p.rd + s * gm.WS;
			const uint64_t kmask = (1ull << (2 * lt)) - 1;
			const int bit = 2 * ds, wi = bit >> 6, sh = bit & 63;
			uint64_t v = row[wi] >> sh;
			if (sh + 2 * lt > 64) v |= row[wi + 1] << (64 - sh);
			// forward: the window holds the key at [ds, ds+lt).  reverse: the reverse-complemented window holds it there,
			// i.e. the window itself holds the key's reverse complement at [L-ds-lt, L-ds).
			key2[0] = v & kmask; key2[1] = rev_fields(~key2[0] & kmask, lt);
			const int nphase = ds > 0 ? 2 : 1;                      // kthread_hash_realign.c:440 (j = 0): no reverse probe for a dictionary at 0
#pragma unroll
			for (int phase = 0; phase < 2; ++phase) {
				if (phase < nphase) {
					const uint32_t b = kmer_bucket(S2_KEY32(key2[phase]), p.pbits);
					lo2[phase] = b ? p.ptab[b - 1] : 0u;
					hi2[phase] = p.ptab[b];
				}
			}
		}
	}
	// count this lane's key matches, reserve room for the whole warp with ONE atomic, then write them (the second sweep over
	// the few table entries hits L1)
	unsigned mine = 0;
#pragma unroll
	for (int phase = 0; phase < 2; ++phase)
		for (uint32_t i = lo2[phase]; i < hi2[phase]; ++i) mine += (p.ents[i] >> S2_POS_BITS) == S2_KEY32(key2[phase]);
	unsigned inc = mine;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
	const unsigned total = __shfl_sync(0xFFFFFFFFu, inc, 31);
	if (total == 0) return;
	unsigned long long base = 0;
	if (lane == 0) base = atomicAdd(&p.counters[CT_S2_NCAND], (unsigned long long)total);
	base = __shfl_sync(0xFFFFFFFFu, base, 0);
	unsigned long long at = base + (inc - mine);
	if (mine) {
#pragma unroll
		for (int phase = 0; phase < 2; ++phase)
			for (uint32_t i = lo2[phase]; i < hi2[phase]; ++i) {
				const unsigned long long e = p.ents[i];
				if ((e >> S2_POS_BITS) == S2_KEY32(key2[phase])) {
					if (at < cand_cap) cand[at] = (idx << 33) | ((unsigned long long)phase << S2_CAND_PHASE_BIT) | (e & S2_POS_MASK);
					++at;
				}
			}
	}
}

// K7b: verify.  One thread per candidate: contig coordinates, window validity, XOR/popcount against the single (or its
// reverse complement), the encode_byte gate, and the claim.
__global__ void __launch_bounds__(128) k_s2_verify(S2Join p, S2Geom gm, const unsigned long long *__restrict__ cand, uint64_t n_cand_listed)
{
	const uint64_t ci = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long n_valid = 0;
	if (ci < n_cand_listed) {
		const unsigned long long cr = cand[ci];
		const uint64_t idx = cr >> 33, P = cr & S2_POS_MASK;
		const int phase = (int)((cr >> S2_CAND_PHASE_BIT) & 1);
		const int l = (int)(idx / p.S); const uint64_t s = idx - (uint64_t)l * p.S;
		const int L = gm.L, Wd = gm.Wd, lt = gm.lt, ds = gm.dstart[l];
		const int koff = phase ? L - ds - lt : ds;
		uint32_t c = p.pblk[P >> S2_BLK_SHIFT];
		S2ContigMeta cm = p.meta[c];
		while (cm.ref_off + cm.len <= P) cm = p.meta[++c];
		const long long jj = (long long)(P - cm.ref_off) - koff;
		if (jj >= 0 && (uint64_t)jj + L <= cm.len) {
			n_valid = 1;
			const uint64_t *__restrict__ row = p.rd + s * gm.WS;
			const uint64_t tailmask = (L & 31) ? (1ull << (2 * (L & 31))) - 1 : ~0ull;
			const int pad = Wd * 32 - L;
			const uint64_t *__restrict__ src = p.cw + cm.cw_off + ((uint64_t)jj >> 5);
			const int sh = 2 * (int)(jj & 31);
			// word q of (window XOR single); for the reverse phase the single is reverse-complemented on the fly:
			// reverse the fields over Wd words, then drop the pad fields that moved to the bottom
			auto xword = [&](int q) -> uint64_t {
				const uint64_t a = src[q];
				uint64_t w = sh ? (a >> sh) | (src[q + 1] << (64 - sh)) : a;
				uint64_t r;
				if (!phase) r = row[q];
				else {
					const uint64_t hi = mcb_rc_word(row[Wd - 1 - q]);
					const uint64_t lo = q + 1 < Wd ? mcb_rc_word(row[Wd - 2 - q]) : 0ull;
					r = pad ? (hi >> (2 * pad)) | (lo << (64 - 2 * pad)) : hi;
				}
				if (q == Wd - 1) { w &= tailmask; r &= tailmask; }
				return w ^ r;
			};
			// the table compared only the first 16 bases of the lt-mer: a candidate whose remaining bases differ was never a
			// dictionary hit (the reference would not have looked at this pair through this dictionary)
			// (lt is 11 or 17: at most ONE base, field koff+16 of the XOR, is left to confirm; it is picked up on the way)
			const int kq = lt > 16 ? (koff + 16) >> 5 : -1, ksh = 2 * ((koff + 16) & 31);
			bool keyeq = true;
			int pc = 0;
#pragma unroll
			for (int q = 0; q < 8; ++q) if (q < Wd) {
				const uint64_t x = xword(q);
				pc += __popcll(x);
				if (q == kq) keyeq = ((x >> ksh) & 3ull) == 0;
			}
			bool ok = keyeq && pc <= gm.thr;
			if (!keyeq) n_valid = 0;
			if (ok && (phase == 0 || gm.thr > 24)) {                                             // encode_byte gate, :393 / :461
				int len_e = 0, eq = 0;
#pragma unroll 1
				for (int q = 0; q < Wd; ++q) {
					uint64_t x = xword(q);
					const int lim = min(32, L - q * 32);
					for (int j = 0; j < lim; ++j, x >>= 2) {
						if (x & 3) {
							if (eq > 1) { len_e += ndigits(eq); eq = 0; }
							else len_e += eq;                                                   // stale-counter quirk of :301-305
							++len_e;
						} else ++eq;
					}
				}
				if (len_e == 0) len_e = 1;
				ok = len_e <= gm.enc_limit;
			}
			if (ok && p.inbig) {                 // exact mode
				const unsigned long long g = cm.woff + (unsigned long long)jj;
				const unsigned long long prio = (g << 5) | ((unsigned long long)phase << 4) | (unsigned long long)l;
				if (p.inbig[s]) {
					const int bit = 2 * ds, wi = bit >> 6, shk = bit & 63;
					uint64_t v = row[wi] >> shk;
					if (shk + 2 * lt > 64) v |= row[wi + 1] << (64 - shk);
					const unsigned long long at = atomicAdd(&p.counters[CT_S2_NEVENTS], 1ull);
					if (at < p.events_cap) { p.events[3 * at] = prio; p.events[3 * at + 1] = ((unsigned long long)l << 34) | (v & ((1ull << (2 * lt)) - 1)); p.events[3 * at + 2] = p.sidx ? p.sidx[s] : (uint32_t)s; }
				} else atomicMin(&p.claim[s], prio);
				ok = false;
			}
			if (ok) {
				if (p.counters[CT_S2_MAXBIN] > (unsigned long long)gm.maxsearch) {
					// The reference scans only the last `maxsearch` live entries of a bin (:388); with every bin at most that
					// large the scan sees everything and "all matches, first one wins" is exact.
					bool big = true;
					if (p.xkey || p.cmg) {
						const int bit = 2 * ds, wi = bit >> 6, shk = bit & 63;
						uint64_t v = row[wi] >> shk;
						if (shk + 2 * lt > 64) v |= row[wi + 1] << (64 - shk);
						const unsigned long long bkey = ((unsigned long long)l << 34) | (v & ((1ull << (2 * lt)) - 1));
						big = p.xkey ? bins_exact_count(p.xkey, p.xcnt, p.xmask, bkey) > (uint32_t)gm.maxsearch : p.cmg[mix64(bkey) & p.cmg_mask] > (uint32_t)gm.maxsearch;
					}
					if (big) atomicAdd(&p.counters[CT_S2_NEEDEXACT], 1ull);
				}
				const unsigned long long g = cm.woff + (unsigned long long)jj;
				atomicMin(&p.claim[s], (g << 5) | ((unsigned long long)phase << 4) | (unsigned long long)l);
			}
		}
	}
	for (int o = 16; o; o >>= 1) n_valid += __shfl_xor_sync(0xFFFFFFFFu, n_valid, o);
	if ((threadIdx.x & 31) == 0 && n_valid) atomicAdd(&p.counters[CT_S2_CAND], n_valid);
}

// ---------------------------------------------------------------- K8
// sidx: position of every single in the job's sg list (sharded: the local singles are a subsequence of it); null = identity
#define S2_NOCLAIM ((unsigned long long)MCB_CLAIM_NONE)
__global__ void k_s2_claim_flags(const unsigned long long *__restrict__ claim, const uint8_t *__restrict__ flagged, uint64_t S,
                                 uint32_t *__restrict__ f_claim, uint32_t *__restrict__ f_a, uint32_t *__restrict__ f_t)
{
	uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= S) return;
	f_claim[s] = claim[s] != S2_NOCLAIM; f_a[s] = flagged[s] == 1; f_t[s] = flagged[s] == 2;
}
__global__ void k_s2_claim_compact(const unsigned long long *__restrict__ claim, const uint8_t *__restrict__ flagged, uint64_t S, const uint32_t *__restrict__ sidx,
                                   const uint32_t *__restrict__ p_claim, const uint32_t *__restrict__ p_a, const uint32_t *__restrict__ p_t,
                                   ulonglong2 *__restrict__ el, uint32_t *__restrict__ fpa, uint32_t *__restrict__ fpt)
{
	uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= S) return;
	const unsigned long long c = claim[s];
	if (c != S2_NOCLAIM) { ulonglong2 e; e.x = c; e.y = 0xFFFFFFFFull - s; el[p_claim[s]] = e; }   // ascending y == descending sg index (local order == job order)
	const uint32_t gs = sidx ? sidx[s] : (uint32_t)s;
	if (flagged[s] == 1) fpa[p_a[s]] = gs;
	if (flagged[s] == 2) fpt[p_t[s]] = gs;
}
__global__ void k_s2_claim_emit(const ulonglong2 *__restrict__ el, uint64_t n, const uint64_t *__restrict__ woff, uint64_t n_contigs, const uint32_t *__restrict__ sg,
                                const uint32_t *__restrict__ sidx, uint32_t *__restrict__ out_c, uint32_t *__restrict__ out_s, uint64_t *__restrict__ out_y, uint64_t *__restrict__ out_p)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	ulonglong2 e = el[i];
	const uint64_t g = e.x >> 5; const unsigned dir = (unsigned)(e.x >> 4) & 1u;
	const uint32_t s = (uint32_t)(0xFFFFFFFFull - e.y);
	uint64_t lo = 0, hi = n_contigs;
	while (hi - lo > 1) { uint64_t mid = (lo + hi) >> 1; if (woff[mid] <= g) lo = mid; else hi = mid; }
	out_c[i] = (uint32_t)lo; out_s[i] = sidx ? sidx[s] : s;
	out_y[i] = ((uint64_t)sg[s] << 32) | ((g - woff[lo]) << 1) | dir;
	out_p[i] = e.x;
}

// ================================================================= host
// Upload the contigs and (re)build the lt-mer table unless the cached one was built from identical contigs.
// d_refs != null: the strings are already on the device (the contig merge just made them): copied device to device, never compared.
static int contig_index_update(mcb_ctx *ctx, const char *refs, const uint64_t *ref_off, uint64_t n_contigs, int lt, const char *d_refs = nullptr)
{
	McbContigIndex &cx = ctx->cix;
	const int L = ctx->L;
	unsigned long long *dc = ctx->d_counters.as<unsigned long long>();
	const uint64_t ref_bytes = n_contigs ? ref_off[n_contigs] : 0;
	if (ref_bytes >= 0xFFFFFFFFull) { mcb_set_error("mcb_realign: %llu contig bases exceed the 2^32 positions of one call", (unsigned long long)ref_bytes); return MCB_EINVAL; }
	// ---- host-side offsets
	uint64_t total_words = 0, n_windows = 0, n_entries = 0;
	MCB_TRY(ctx->h_in0.ensure((n_contigs + 1) * 8)); MCB_TRY(ctx->h_in1.ensure((n_contigs + 1) * 8));
	MCB_TRY(ctx->h_in2.ensure((n_contigs + 1) * 8));
	uint64_t *cwo = ctx->h_in0.as<uint64_t>(), *wo = ctx->h_in1.as<uint64_t>(), *eo = ctx->h_in2.as<uint64_t>();
	for (uint64_t c = 0; c < n_contigs; ++c) {
		if (ref_off[c + 1] < ref_off[c]) { mcb_set_error("mcb_realign: ref_off is not monotonic"); return MCB_EINVAL; }
		const uint64_t len = ref_off[c + 1] - ref_off[c];
		cwo[c] = total_words; wo[c] = n_windows; eo[c] = n_entries;
		total_words += (len + 31) / 32 + 1;
		if (len >= (uint64_t)L) { n_windows += len - L + 1; n_entries += len - lt + 1; }
	}
	cwo[n_contigs] = total_words + 1; wo[n_contigs] = n_windows; eo[n_contigs] = n_entries;   // +1 guard word: loaders read one word ahead
	if (n_windows >= (1ull << 58)) { mcb_set_error("too many windows"); return MCB_EINVAL; }
	// ---- same contigs as last time?  (compare on the device: the strings have to be uploaded to find out)
	// Sharded: every rank needs all packed contigs, but each uploads and packs only its share of the words (an equal slice of the
	// padded word array) and one in-place all-gather over NVLink completes the array; nothing is compared with the cache then
	// (the caller says "same contigs" by passing NULL).
	const bool sharded = ctx->shard_n > 1 && ctx->comm;
	const uint64_t chunk = sharded ? (total_words + 2 + ctx->shard_n - 1) / ctx->shard_n : total_words + 2;
	const uint64_t w_begin = sharded ? std::min<uint64_t>((uint64_t)ctx->shard_rank * chunk, total_words) : 0;
	const uint64_t w_end = sharded ? std::min<uint64_t>(w_begin + chunk, total_words) : total_words;
	uint64_t b_lo = 0, b_hi = ref_bytes;                  // characters the words [w_begin, w_end) are made of
	if (sharded && n_contigs) {
		if (w_end > w_begin) {
			const uint64_t c0 = (uint64_t)(std::upper_bound(cwo, cwo + n_contigs, w_begin) - cwo) - 1, c1 = (uint64_t)(std::upper_bound(cwo, cwo + n_contigs, w_end - 1) - cwo) - 1;
			b_lo = std::min(ref_off[c0] + (w_begin - cwo[c0]) * 32, ref_off[c0 + 1]);
			b_hi = std::min(ref_off[c1] + (w_end - cwo[c1]) * 32, ref_off[c1 + 1]);
		} else b_lo = b_hi = 0;
	}
	const bool maybe_same = !sharded && !d_refs && cx.valid && cx.n_contigs == n_contigs && cx.ref_bytes == ref_bytes && cx.lt == lt && cx.L == L;
	DBuf &stage_refs = maybe_same ? ctx->d_scr[1] : cx.refs, &stage_off = maybe_same ? ctx->d_scr[2] : cx.roff;
	const size_t refs_pad = (ref_bytes + 7) & ~(size_t)7;
	MCB_TRY(stage_refs.ensure(refs_pad + 16)); MCB_TRY(stage_off.ensure((n_contigs + 1) * 8));
	cx.valid = cx.valid && maybe_same;
	if (d_refs) {
		if (ref_bytes) MCB_CUDA(cudaMemcpyAsync(stage_refs.p, d_refs, ref_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
		if (refs_pad > ref_bytes) MCB_CUDA(cudaMemsetAsync(stage_refs.as<char>() + ref_bytes, 0, refs_pad - ref_bytes, ctx->stream));
		MCB_CUDA(cudaMemcpyAsync(stage_off.p, ref_off, (n_contigs + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
	} else {   // the contig strings go up on the copy stream: the singles kernels already queued on the compute stream run meanwhile
		MCB_TRY(mcb_copy_streams(ctx));
		struct EvPair {                   // destroyed on every exit path
			cudaEvent_t a = nullptr, b = nullptr;
			~EvPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
		} ev;
		MCB_CUDA(cudaEventCreate(&ev.a)); MCB_CUDA(cudaEventCreate(&ev.b));
		MCB_CUDA(cudaEventRecord(ev.a, ctx->copy_stream));             // (every earlier call has synchronized: nothing else touches these buffers)
		if (refs_pad > ref_bytes) MCB_CUDA(cudaMemsetAsync(stage_refs.as<char>() + (refs_pad - 8), 0, 8, ctx->copy_stream));
		if (b_hi > b_lo) MCB_TRY(mcb_h2d_on(ctx, ctx->copy_stream, stage_refs.as<char>() + b_lo, refs + b_lo, b_hi - b_lo, 4));
		MCB_TRY(mcb_h2d_on(ctx, ctx->copy_stream, stage_off.p, ref_off, (n_contigs + 1) * 8, 1));
		MCB_CUDA(cudaEventRecord(ev.b, ctx->copy_stream));
		MCB_CUDA(cudaStreamWaitEvent(ctx->stream, ev.b, 0));
		if (ctx->tm.enabled) {
			cudaEventSynchronize(ev.b);
			float ms = 0; if (cudaEventElapsedTime(&ms, ev.a, ev.b) == cudaSuccess) { int id = ctx->tm.id("h2d"); ctx->tm.ms[id] += ms; ctx->tm.cnt[id] += 1; }
		}
	}
	if (maybe_same) {
		McbSpan sp(ctx->tm, "realign");
		MCB_CUDA(cudaMemsetAsync(&dc[CT_S2_DIFF], 0, 8, ctx->stream));
		if (refs_pad) MCB_LAUNCH(ctx, "s2_same", k_s2_same, mcb_grid_for(refs_pad / 8, 256), 256, 0, stage_refs.as<uint64_t>(), cx.refs.as<uint64_t>(), refs_pad / 8, dc);
		MCB_LAUNCH(ctx, "s2_same", k_s2_same, mcb_grid_for(n_contigs + 1, 256), 256, 0, stage_off.as<uint64_t>(), cx.roff.as<uint64_t>(), n_contigs + 1, dc);
		MCB_CUDA(cudaMemcpyAsync(ctx->h_counters.p, dc, 64 * 8, cudaMemcpyDeviceToHost, ctx->stream));
		MCB_CUDA(cudaStreamSynchronize(ctx->stream));
		if (ctx->h_counters.as<unsigned long long>()[CT_S2_DIFF] == 0) return MCB_OK;      // identical: keep the table
		cx.valid = false;
		std::swap(cx.refs, ctx->d_scr[1]); std::swap(cx.roff, ctx->d_scr[2]);
	}
	// ---- pack
	cx.table_valid = false;
	cx.n_contigs = n_contigs; cx.ref_bytes = ref_bytes; cx.total_words = total_words; cx.n_windows = n_windows; cx.n_entries = n_entries; cx.L = L; cx.lt = lt;
	const uint64_t n_blocks = (ref_bytes >> S2_BLK_SHIFT) + 1;
	MCB_TRY(cx.cwo.ensure((n_contigs + 1) * 8)); MCB_TRY(cx.wo.ensure((n_contigs + 1) * 8)); MCB_TRY(cx.cw.ensure((chunk * (sharded ? ctx->shard_n : 1) + 2) * 8));
	MCB_TRY(cx.pblk.ensure(n_blocks * 4 + 16));
	MCB_TRY(cx.eoff.ensure((n_contigs + 1) * 8)); MCB_TRY(cx.meta.ensure((n_contigs + 2) * 32));
	{
		McbSpan sp(ctx->tm, "h2d");
		MCB_CUDA(cudaMemcpyAsync(cx.cwo.p, cwo, (n_contigs + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
		MCB_CUDA(cudaMemcpyAsync(cx.wo.p, wo, (n_contigs + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
		MCB_CUDA(cudaMemcpyAsync(cx.eoff.p, eo, (n_contigs + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
	}
	if (n_contigs == 0 || n_windows == 0) { cx.valid = true; return MCB_OK; }
	McbSpan sp(ctx->tm, "realign");
	MCB_LAUNCH(ctx, "s2_contig_meta", k_s2_contig_meta, mcb_grid_for(n_contigs + 1, 256), 256, 0, cx.roff.as<uint64_t>(), cx.cwo.as<uint64_t>(), cx.wo.as<uint64_t>(), n_contigs,
	           cx.meta.as<S2ContigMeta>());
	if (w_end > w_begin) MCB_LAUNCH(ctx, "s2_pack_refs", k_s2_pack_refs, mcb_grid_for(w_end - w_begin, 256), 256, 0, cx.refs.as<char>(), cx.roff.as<uint64_t>(), cx.cwo.as<uint64_t>(),
	                                n_contigs, w_begin, w_end, cx.cw.as<uint64_t>(), dc);
	if (sharded) {                 // ("nccl_in:" = a collective inside the entry point's own span)
		McbSpan sp2(ctx->tm, "nccl_in:contigs");
		MCB_TRY(mcb_coll_allgather_inplace_u64(ctx, cx.cw.as<uint64_t>(), chunk));
	}
	MCB_CUDA(cudaMemsetAsync(cx.cw.as<uint64_t>() + total_words, 0, 16, ctx->stream));              // guard words: loaders read one word ahead
	MCB_LAUNCH(ctx, "s2_pos_blocks", k_s2_pos_blocks, mcb_grid_for(n_blocks, 256), 256, 0, cx.roff.as<uint64_t>(), n_contigs, n_blocks, cx.pblk.as<uint32_t>());
	cx.valid = true;
	return MCB_OK;
}

// The lt-mer table of the packed contigs: (lt-mer, position) entries sorted by hashed bucket, with bucket end offsets.  Built
// after the singles of the call are known, so that only the lt-mers some single can ask for are kept (key filter above); the
// table is rebuilt when a later call brings singles that were not in the filter, or another dictionary geometry.
static int contig_table_update(mcb_ctx *ctx, const S2Geom &gm, const uint32_t *d_sg, uint64_t S, const uint64_t *d_rd, const uint8_t *d_fl)
{
	McbContigIndex &cx = ctx->cix;
	unsigned long long *dc = ctx->d_counters.as<unsigned long long>();
	unsigned long long *hc = ctx->h_counters.as<unsigned long long>();
	static const bool use_filter = !(getenv("MCB_S2_NOFILTER") && atoi(getenv("MCB_S2_NOFILTER")));
	if (cx.n_windows == 0 || cx.n_contigs == 0) { cx.table_valid = true; cx.filtered = false; return MCB_OK; }
	if (cx.table_valid) {
		if (!cx.filtered) return MCB_OK;
		if (cx.nd == gm.nd && cx.dstart0 == gm.dstart[0]) {
			MCB_CUDA(cudaMemsetAsync(&dc[CT_S2_DIFF], 0, 8, ctx->stream));
			if (S) MCB_LAUNCH(ctx, "s2_sg_subset", k_s2_sg_subset, mcb_grid_for(S, 256), 256, 0, d_sg, S, cx.sgmap.as<uint32_t>(), dc);
			MCB_CUDA(cudaMemcpyAsync(ctx->h_counters.p, dc, 64 * 8, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaStreamSynchronize(ctx->stream));
			if (hc[CT_S2_DIFF] == 0) return MCB_OK;
		}
		cx.table_valid = false;
	}
	const uint64_t n_all = cx.n_entries;          // all lt-mer start positions of the contigs that have a window
	const uint64_t nkv = S * (uint64_t)gm.nd;
	const bool filtered = use_filter;
	uint64_t wmask = 0;
	if (filtered) {
		static const int flt_shift = getenv("MCB_S2_FLT") ? atoi(getenv("MCB_S2_FLT")) : -1;             // tuning knob: filter words per (single, dictionary), log2
		const uint64_t want = flt_shift >= 0 ? (nkv + 1) << flt_shift : (nkv + 1) >> -flt_shift;
		uint64_t W = 1024; while (W < want && W < (1ull << 28)) W <<= 1;
		MCB_TRY(cx.flt.ensure(W * 4)); MCB_TRY(cx.sgmap.ensure((ctx->n_reads / 32 + 2) * 4));
		MCB_CUDA(cudaMemsetAsync(cx.flt.p, 0, W * 4, ctx->stream));
		MCB_CUDA(cudaMemsetAsync(cx.sgmap.p, 0, (ctx->n_reads / 32 + 2) * 4, ctx->stream));
		wmask = W - 1; cx.flt_words = W;
		if (nkv) MCB_LAUNCH(ctx, "s2_filter_insert", k_s2_filter_insert, mcb_grid_for(nkv, 256), 256, 0, d_rd, d_fl, d_sg, S, gm, cx.flt.as<uint32_t>(), wmask, cx.sgmap.as<uint32_t>());
	}
	const bool compacting = filtered;
	uint64_t ents_cap = !compacting ? n_all : n_all / 4 + (1u << 20);
	if (ents_cap > n_all) ents_cap = n_all;
	uint64_t n_own = n_all;
	for (int tries = 0;; ++tries) {
		MCB_TRY(cx.ents.ensure(ents_cap * 8 + 16));
		if (compacting) MCB_CUDA(cudaMemsetAsync(&dc[CT_S2_NCAND], 0, 8, ctx->stream));
		MCB_LAUNCH(ctx, "s2_kmer_emit", k_s2_kmer_emit, mcb_grid_for(cx.total_words, 256), 256, 0, cx.cw.as<uint64_t>(), cx.cwo.as<uint64_t>(), cx.roff.as<uint64_t>(),
		           cx.eoff.as<uint64_t>(), cx.n_contigs, cx.total_words, cx.L, cx.lt, cx.ents.as<unsigned long long>(), compacting ? 1 : 0, &dc[CT_S2_NCAND],
		           (unsigned long long)ents_cap, filtered ? cx.flt.as<uint32_t>() : (const uint32_t*)nullptr, wmask);
		if (!compacting) break;
		MCB_CUDA(cudaMemcpyAsync(ctx->h_counters.p, dc, 64 * 8, cudaMemcpyDeviceToHost, ctx->stream));
		MCB_CUDA(cudaStreamSynchronize(ctx->stream));
		n_own = hc[CT_S2_NCAND];
		if (n_own <= ents_cap) break;
		if (tries) { mcb_set_error("mcb_realign: lt-mer table overflow (%llu entries, room for %llu)", (unsigned long long)n_own, (unsigned long long)ents_cap); return MCB_EINVAL; }
		ents_cap = n_own;                                                               // the count is exact: emit again into enough room
	}
	cx.n_table = n_own;
	static const int load_shift = getenv("MCB_S2_LOAD") ? atoi(getenv("MCB_S2_LOAD")) : 2;                   // tuning knob: 2^load_shift .. 2^(load_shift+1) entries per bucket
	int pbits = 10;
	while (pbits < 30 && pbits < 2 * cx.lt && ((1ull << load_shift) << pbits) < n_own) ++pbits;
	cx.pbits = pbits;
	cx.b_lo = 0; cx.b_hi = 1u << pbits;
	const uint64_t nbk = cx.b_hi - cx.b_lo;
	MCB_TRY(cx.ents2.ensure(n_own * 8 + 16)); MCB_TRY(cx.ptab.ensure((nbk + 1) * 4));
	MCB_TRY(mcb_radix_sort_kmers(ctx, cx.ents.as<unsigned long long>(), cx.ents2.as<unsigned long long>(), n_own, pbits, cx.b_lo, cx.b_hi, &cx.ents_sorted));
	MCB_CUDA(cudaMemsetAsync(cx.ptab.p, 0, (nbk + 1) * 4, ctx->stream));
	if (n_own) MCB_LAUNCH(ctx, "s2_bucket_ends", k_s2_bucket_ends, mcb_grid_for(n_own, 256), 256, 0, cx.ents_sorted, n_own, pbits, cx.b_lo, cx.b_hi, cx.ptab.as<uint32_t>());
	cx.table_valid = true; cx.filtered = filtered; cx.nd = gm.nd; cx.dstart0 = gm.dstart[0];
	return MCB_OK;
}

// d_scr roles of one mcb_realign: 0 sg, 3 flags (K8), 4 fpA|fpT, 5 count-min sketch / exact bins, 6 rd, 8 flagged, 9/10 claim elements, 11 claim outputs
static int realign_geometry(mcb_ctx *ctx, int threshold, int maxsearch, int ininumdict, S2Geom *out)
{
	const int L = ctx->L;
	S2Geom gm; memset(&gm, 0, sizeof gm);
	gm.L = L; gm.Wd = ctx->Wd; gm.WS = ctx->WS; gm.thr = threshold; gm.maxsearch = maxsearch;
	gm.lt = L <= 80 ? 11 : 17;
	gm.nd = L / gm.lt;
	if (ininumdict > 1 && ininumdict < gm.nd) gm.nd = ininumdict;
	if (gm.nd > S2_MAXD) { mcb_set_error("numdict %d unsupported", gm.nd); return MCB_EINVAL; }
	int st0 = (ininumdict > 0 && ininumdict < gm.nd) ? L / 2 - (gm.lt * gm.nd) / 2 : 0;
	for (int i = 0; i < gm.nd; ++i) gm.dstart[i] = st0 + i * gm.lt;
	gm.enc_limit = (int)((double)L * 0.4);
	*out = gm;
	return MCB_OK;
}

// Sequential bin-window emulation for the singles that sit in a dictionary bin larger than maxsearch.
// The device lists (a) the members of every such bin and (b) every verified match of every such single ("events").  The
// host replays the events in the reference's probe order (priority ascending; inside one probe the bin is scanned from the
// highest sg index down, kthread_hash_realign.c:388) with the reference's rules: a match counts only if the single is among
// the last `maxsearch` LIVE entries of the probed bin; a claimed single leaves all its bins after the probe (:409-424),
// except that the last entry of a bin is never removed and the bin is closed instead (bbhashdict.c:51-55); singles set
// aside by the poly-A/T diversion stay in the bins for good.  Singles outside big bins keep the parallel first-wins rule.
#include <unordered_map>
struct BigBin { std::vector<uint32_t> mem; std::vector<int> fen; uint32_t live = 0; bool closed = false; };
static void fen_add(std::vector<int> &f, size_t i, int d) { for (++i; i < f.size(); i += i & (~i + 1)) f[i] += d; }
static int fen_sum(const std::vector<int> &f, size_t i) { int r = 0; for (; i > 0; i -= i & (~i + 1)) r += f[i]; return r; }   // live among the first i members

static int realign_exact_bins(mcb_ctx *ctx, S2Join jn, const S2Geom &gm, uint64_t S, uint64_t nkv, uint64_t n_listed, const unsigned long long *d_cand, const uint32_t *h_sidx)
{
	unsigned long long *dc = ctx->d_counters.as<unsigned long long>();
	unsigned long long *hc = ctx->h_counters.as<unsigned long long>();
	DBuf b_inbig, b_mem, b_ev, b_pairs;
	struct Rel { DBuf *b[4]; ~Rel() { for (auto x : b) x->release(); } } rel = {{&b_inbig, &b_mem, &b_ev, &b_pairs}};
	MCB_TRY(b_inbig.ensure(S + 16)); MCB_TRY(b_mem.ensure(nkv * 16 + 16)); MCB_TRY(b_ev.ensure(n_listed * 24 + 24));
	MCB_CUDA(cudaMemsetAsync(b_inbig.p, 0, S + 16, ctx->stream));
	MCB_CUDA(cudaMemsetAsync(&dc[CT_S2_NBIGMEM], 0, 16, ctx->stream));      // NBIGMEM, NEVENTS
	MCB_CUDA(cudaMemsetAsync(&dc[CT_S2_CAND], 0, 8, ctx->stream)); MCB_CUDA(cudaMemsetAsync(&dc[CT_S2_NEEDEXACT], 0, 8, ctx->stream));
	if (S) MCB_CUDA(cudaMemsetAsync(jn.claim, 0x7F, S * 8, ctx->stream));
	if (nkv) MCB_LAUNCH(ctx, "s2_mark_big", k_s2_mark_big, mcb_grid_for(nkv, 256), 256, 0, jn.rd, jn.flagged, jn.sidx, S, gm, jn.xkey, jn.xcnt, jn.xmask, jn.cmg, jn.cmg_mask,
	                    b_inbig.as<uint8_t>(), b_mem.as<unsigned long long>(), (unsigned long long)nkv, dc);
	jn.inbig = b_inbig.as<uint8_t>(); jn.events = b_ev.as<unsigned long long>(); jn.events_cap = n_listed;
	if (n_listed) MCB_LAUNCH(ctx, "s2_verify", k_s2_verify, mcb_grid_for(n_listed, 128), 128, 0, jn, gm, d_cand, n_listed);
	MCB_CUDA(cudaMemcpyAsync(ctx->h_counters.p, dc, 64 * 8, cudaMemcpyDeviceToHost, ctx->stream));
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	uint64_t n_mem = hc[CT_S2_NBIGMEM], n_ev = hc[CT_S2_NEVENTS];
	std::vector<unsigned long long> mem, ev;
	if (ctx->shard_n > 1) {
		// sharded: the bins and the events of the whole job on every rank (they are few); each rank replays all of them and keeps
		// the winners among its own singles
		MCB_TRY(mcb_coll_allgatherv(ctx, b_mem.p, n_mem * 16, mem));
		MCB_TRY(mcb_coll_allgatherv(ctx, b_ev.p, n_ev * 24, ev));
		n_mem = mem.size() / 2; n_ev = ev.size() / 3;
	} else {
		mem.resize(2 * n_mem); ev.resize(3 * n_ev);
		if (n_mem) MCB_CUDA(cudaMemcpyAsync(mem.data(), b_mem.p, n_mem * 16, cudaMemcpyDeviceToHost, ctx->stream));
		if (n_ev) MCB_CUDA(cudaMemcpyAsync(ev.data(), b_ev.p, n_ev * 24, cudaMemcpyDeviceToHost, ctx->stream));
		MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	}
	// bins
	std::unordered_map<unsigned long long, BigBin> bins;
	std::unordered_map<uint32_t, std::vector<unsigned long long>> bins_of;      // single -> keys of its big bins
	std::unordered_map<uint32_t, uint8_t> flagged;                                // singles of the big bins set aside by the poly-A/T diversion
	for (uint64_t i = 0; i < n_mem; ++i) {
		const uint32_t sgi = (uint32_t)mem[2 * i + 1];
		bins[mem[2 * i]].mem.push_back(sgi); bins_of[sgi].push_back(mem[2 * i]);
		if (mem[2 * i + 1] >> 32) flagged[sgi] = 1;
	}
	for (auto &kv : bins) {
		BigBin &b = kv.second;
		std::sort(b.mem.begin(), b.mem.end());
		b.live = (uint32_t)b.mem.size();
		b.fen.assign(b.mem.size() + 1, 0);
		for (size_t i = 0; i < b.mem.size(); ++i) fen_add(b.fen, i, 1);
	}
	// events in probe order; inside a probe from the highest sg index down
	std::vector<uint64_t> order(n_ev);
	for (uint64_t i = 0; i < n_ev; ++i) order[i] = i;
	std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) {
		if (ev[3 * a] != ev[3 * b]) return ev[3 * a] < ev[3 * b];
		return ev[3 * a + 2] > ev[3 * b + 2];
	});
	std::unordered_map<uint32_t, unsigned long long> claimed;                  // single -> winning priority
	std::vector<uint32_t> removed;
	for (uint64_t i = 0; i < n_ev;) {
		uint64_t j = i;
		while (j < n_ev && ev[3 * order[j]] == ev[3 * order[i]]) ++j;            // one probe = one (window, phase, l) = one bin
		const unsigned long long prio = ev[3 * order[i]], key = ev[3 * order[i] + 1];
		auto bit = bins.find(key);
		BigBin *bb = bit == bins.end() ? nullptr : &bit->second;
		removed.clear();
		if (!bb || !bb->closed) {
			for (uint64_t q = i; q < j; ++q) {
				const uint32_t sgi = (uint32_t)ev[3 * order[q] + 2];
				if (flagged.count(sgi) || claimed.count(sgi)) continue;
				if (bb) {                                                              // among the last maxsearch live entries?
					const size_t pos = (size_t)(std::lower_bound(bb->mem.begin(), bb->mem.end(), sgi) - bb->mem.begin());
					const int live_above = (int)bb->live - fen_sum(bb->fen, pos + 1);
					if (live_above >= gm.maxsearch) continue;
				}
				claimed[sgi] = prio;
				removed.push_back(sgi);
			}
		}
		for (uint32_t sgi : removed)                                                  // :409-424: out of every dictionary
			for (unsigned long long k2 : bins_of[sgi]) {
				BigBin &b = bins[k2];
				if (b.live == 1) { b.closed = true; continue; }                       // last entry stays, bin is closed
				const size_t pos = (size_t)(std::lower_bound(b.mem.begin(), b.mem.end(), sgi) - b.mem.begin());
				fen_add(b.fen, pos, -1); --b.live;
			}
		i = j;
	}
	// hand the winners among this context's singles back to the device claim array (job position -> local position)
	std::vector<unsigned long long> pairs; pairs.reserve(2 * claimed.size());
	for (auto &kv : claimed) {
		uint64_t local = kv.first;
		if (h_sidx) {
			const uint32_t *q = std::lower_bound(h_sidx, h_sidx + S, kv.first);
			if (q == h_sidx + S || *q != kv.first) continue;                            // another rank's single
			local = (uint64_t)(q - h_sidx);
		}
		pairs.push_back(local); pairs.push_back(kv.second);
	}
	if (!pairs.empty()) {
		MCB_TRY(b_pairs.ensure(pairs.size() * 8));
		MCB_CUDA(cudaMemcpyAsync(b_pairs.p, pairs.data(), pairs.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
		MCB_LAUNCH(ctx, "s2_apply_claims", k_s2_apply_claims, mcb_grid_for(pairs.size() / 2, 256), 256, 0, b_pairs.as<unsigned long long>(), (uint64_t)pairs.size() / 2, jn.claim);
		MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	}
	return MCB_OK;
}

// first half: contigs, singles, join.  Leaves the claim priorities in ctx->d_x[0] (u64[S]).
// sg_index (sharded): position of every local single in the job's sg list, ascending; n_sg_total = length of that list.  The
// dictionary bins of the reference span the singles of the whole job, so the bin-size guard is summed over the ranks.
static int realign_search(mcb_ctx *ctx, const uint32_t *sg, const uint32_t *sg_index, uint64_t S, uint64_t n_sg_total, const char *refs, const uint64_t *ref_off, uint64_t n_contigs,
                          int threshold, int maxsearch, int ininumdict, mcb_realign_result *res)
{
	if (!ctx->reads_loaded) { mcb_set_error("mcb_realign: no reads loaded"); return MCB_ESTATE; }
	if (S && !sg) { mcb_set_error("mcb_realign: null input"); return MCB_EINVAL; }
	if (n_sg_total >= 0xFFFFFFFFull || S > n_sg_total) { mcb_set_error("mcb_realign: too many singles"); return MCB_EINVAL; }
	const bool sharded = ctx->shard_n > 1;
	if (sharded && !ctx->comm) { mcb_set_error("mcb_realign: sharded context without a communicator (mcb_shard_init)"); return MCB_ESTATE; }
	if (sharded && S && !sg_index) { mcb_set_error("mcb_shard_realign: sg_index is required"); return MCB_EINVAL; }
	memset(res, 0, sizeof(*res));
	const int L = ctx->L, WS = ctx->WS;
	S2Geom gm;
	MCB_TRY(realign_geometry(ctx, threshold, maxsearch, ininumdict, &gm));       // setglobalarrays_realign, kthread_hash_realign.c:150-207
	res->numdict = gm.nd;
	MCB_TRY(ctx->d_counters.ensure(64 * 8)); MCB_TRY(ctx->h_counters.ensure(64 * 8));
	unsigned long long *dc = ctx->d_counters.as<unsigned long long>();
	unsigned long long *hc = ctx->h_counters.as<unsigned long long>();
	MCB_CUDA(cudaMemsetAsync(dc + 16, 0, 16 * 8, ctx->stream));
	// ---- singles first: they do not depend on the contigs, and the contig upload (copy stream) overlaps them
	const uint64_t nkv = S * (uint64_t)gm.nd, nkv_job = n_sg_total * (uint64_t)gm.nd;
	if (nkv >= 0x7FFFFFFFull) { mcb_set_error("dictionary too large (singles x dictionaries must stay below 2^31)"); return MCB_EINVAL; }
	DBuf &b_sg = ctx->d_scr[0], &b_cm = ctx->d_scr[5], &b_rd = ctx->d_scr[6], &b_fl = ctx->d_scr[8], &b_sidx = ctx->d_x[1];
	MCB_TRY(b_sg.ensure(S * 4 + 16)); MCB_TRY(b_rd.ensure(S * WS * 8 + 16)); MCB_TRY(b_fl.ensure(S + 16));
	ctx->rs.have_index = sg_index != nullptr;
	if (sg_index) MCB_TRY(b_sidx.ensure(S * 4 + 16));
	// count-min sketch of the bin sizes: it only has to bound the largest bin from above (a bound above maxsearch sends the call
	// through the exact count), so about one counter per two (single, dictionary) pairs is enough, and at that size (32 MB
	// for 15 M pairs) the atomics resolve in L2 instead of DRAM.  Sized from the JOB's singles: every rank uses the same array.
	static const int cm_shift = getenv("MCB_S2_CM") ? atoi(getenv("MCB_S2_CM")) : -1;
	const uint64_t cm_want = cm_shift >= 0 ? nkv_job << cm_shift : nkv_job >> -cm_shift;
	uint64_t CM = 1024; while (CM < cm_want && CM < (1ull << 23)) CM <<= 1;                          // at most 32 MB: it has to stay in L2
	if (nkv_job) MCB_TRY(b_cm.ensure(CM * 4));
	if (nkv) {
		{
			McbSpan sp(ctx->tm, "h2d");
			MCB_CUDA(cudaMemcpyAsync(b_sg.p, sg, S * 4, cudaMemcpyHostToDevice, ctx->stream));
			if (sg_index) MCB_CUDA(cudaMemcpyAsync(b_sidx.p, sg_index, S * 4, cudaMemcpyHostToDevice, ctx->stream));
		}
		McbSpan span(ctx->tm, "realign");
		MCB_CUDA(cudaMemsetAsync(b_cm.p, 0, CM * 4, ctx->stream));
		MCB_LAUNCH(ctx, "s2_singles", k_s2_singles, mcb_grid_for(S, 128), 128, 0, b_sg.as<uint32_t>(), S, ctx->d_packed.as<uint64_t>(), gm,
		           ctx->d_nread_rid.as<uint32_t>(), ctx->d_nread_mask.as<uint64_t>(), ctx->n_nreads, ctx->n_reads,
		           b_rd.as<uint64_t>(), b_fl.as<uint8_t>(), b_cm.as<uint32_t>(), CM - 1, dc);
	} else if (nkv_job && sharded) MCB_CUDA(cudaMemsetAsync(b_cm.p, 0, CM * 4, ctx->stream));
	// sharded: a bin of the job is at most the sum of the ranks' largest bins — one scalar all-reduce, and usually the end of it
	if (sharded && nkv_job) { McbSpan span(ctx->tm, "nccl:guard"); MCB_TRY(mcb_coll_allreduce_sum_u64(ctx, &dc[CT_S2_MAXBIN], 1)); }
	// ---- contigs: refs == NULL reuses the contigs (and their table) of the previous call
	McbContigIndex &cx = ctx->cix;
	if (refs || ref_off) {
		if (n_contigs && (!refs || !ref_off)) { mcb_set_error("mcb_realign: refs and ref_off must be given together"); return MCB_EINVAL; }
		MCB_TRY(contig_index_update(ctx, refs, ref_off, n_contigs, gm.lt));
	} else {
		if (!cx.valid || cx.L != L || cx.lt != gm.lt) { mcb_set_error("mcb_realign: refs == NULL but no contigs from a previous call are cached"); return MCB_ESTATE; }
		if (n_contigs && n_contigs != cx.n_contigs) { mcb_set_error("mcb_realign: refs == NULL with a different contig count (%llu, cached %llu)", (unsigned long long)n_contigs, (unsigned long long)cx.n_contigs); return MCB_EINVAL; }
	}
	const uint64_t n_windows = cx.n_windows;
	if (n_windows >= (1ull << 57)) { mcb_set_error("too many windows"); return MCB_EINVAL; }
	res->n_windows = n_windows;
	{
		int nrev = 0; for (int l = 0; l < gm.nd; ++l) nrev += gm.dstart[l] > 0;
		res->n_probes = n_windows * (uint64_t)(gm.nd + nrev);                 // what the reference's window loop would issue (:355-504)
	}
	ctx->rs.pending = true; ctx->rs.S = S; ctx->rs.nd = gm.nd;
	MCB_TRY(ctx->d_x[0].ensure(S * 8 + 16));                   // claim priorities
	unsigned long long *claim = ctx->d_x[0].as<unsigned long long>();
	if (S) MCB_CUDA(cudaMemsetAsync(claim, 0x7F, S * 8, ctx->stream));          // MCB_CLAIM_NONE
	if (n_sg_total == 0 || gm.nd == 0) return MCB_OK;
	McbSpan span(ctx->tm, "realign");
	if (S) MCB_TRY(contig_table_update(ctx, gm, b_sg.as<uint32_t>(), S, b_rd.as<uint64_t>(), b_fl.as<uint8_t>()));
	S2Join jn; memset(&jn, 0, sizeof jn);
	jn.S = S; jn.rd = b_rd.as<uint64_t>(); jn.flagged = b_fl.as<uint8_t>(); jn.ptab = cx.ptab.as<uint32_t>(); jn.ents = cx.ents_sorted; jn.pbits = cx.pbits;
	jn.pblk = cx.pblk.as<uint32_t>(); jn.meta = cx.meta.as<S2ContigMeta>(); jn.cw = cx.cw.as<uint64_t>();
	jn.claim = claim; jn.counters = dc; jn.sidx = sg_index ? b_sidx.as<uint32_t>() : nullptr;
	// candidate list: sized from the previous call's count, grown (and the probe repeated) when it overflows
	DBuf &b_cand = ctx->d_x[3];
	uint64_t n_listed = 0;
	if (S && n_windows) {
		if (b_cand.cap < (nkv / 2 + 1024) * 8) MCB_TRY(b_cand.ensure((nkv / 2 + 1024) * 8));
		for (int tries = 0;; ++tries) {
			const uint64_t cap = b_cand.cap / 8;
			MCB_CUDA(cudaMemsetAsync(&dc[CT_S2_NCAND], 0, 8, ctx->stream));
			MCB_LAUNCH(ctx, "s2_probe_table", k_s2_probe_table, mcb_grid_for(nkv, 256), 256, 0, jn, gm, b_cand.as<unsigned long long>(), (unsigned long long)cap);
			MCB_CUDA(cudaMemcpyAsync(ctx->h_counters.p, dc, 64 * 8, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaStreamSynchronize(ctx->stream));
			n_listed = hc[CT_S2_NCAND];
			if (n_listed <= cap) break;
			if (tries) { mcb_set_error("mcb_realign: candidate list overflow"); return MCB_EINVAL; }
			MCB_TRY(b_cand.ensure(n_listed * 8 + 1024));
		}
	}
	for (int attempt = 0;; ++attempt) {
		if (n_listed) MCB_LAUNCH(ctx, "s2_verify", k_s2_verify, mcb_grid_for(n_listed, 128), 128, 0, jn, gm, b_cand.as<unsigned long long>(), n_listed);
		// every rank must take the same way through the guard: the decision counters are summed over the ranks first
		if (sharded) MCB_TRY(mcb_coll_allreduce_sum_u64(ctx, &dc[CT_S2_NEEDEXACT], 2));       // NEEDEXACT and ERR (inside the "realign" span)
		MCB_CUDA(cudaMemcpyAsync(ctx->h_counters.p, dc, 64 * 8, cudaMemcpyDeviceToHost, ctx->stream));
		MCB_CUDA(cudaStreamSynchronize(ctx->stream));
		if (hc[CT_S2_ERR]) { mcb_set_error("mcb_realign: %llu invalid inputs (sg id out of range or non-ACGT contig character)", hc[CT_S2_ERR]); return MCB_EINPUT; }
		if (!hc[CT_S2_NEEDEXACT]) break;
		if (attempt == 1) {
			// Some verified match lies in a bin larger than maxsearch: the reference's scan of "the last maxsearch live entries"
			// (kthread_hash_realign.c:388) depends on which reads were already claimed.  Replay exactly those singles in order.
			MCB_TRY(realign_exact_bins(ctx, jn, gm, S, nkv, n_listed, b_cand.as<unsigned long long>(), sg_index));
			MCB_CUDA(cudaMemcpyAsync(ctx->h_counters.p, dc, 64 * 8, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaStreamSynchronize(ctx->stream));
			break;
		}
		// the bound was too coarse: judge every bin by itself and join again
		if (!sharded) {               // exact bin sizes
			uint64_t H = 1024; while (H < 2 * nkv) H <<= 1;
			MCB_TRY(b_cm.ensure(H * 12 + 16));
			unsigned long long *tkey = b_cm.as<unsigned long long>(); uint32_t *tcnt = (uint32_t*)(b_cm.as<char>() + H * 8);
			MCB_CUDA(cudaMemsetAsync(tkey, 0xFF, H * 8, ctx->stream)); MCB_CUDA(cudaMemsetAsync(tcnt, 0, H * 4, ctx->stream));
			MCB_LAUNCH(ctx, "s2_bins_exact", k_s2_bins_exact, mcb_grid_for(nkv, 256), 256, 0, b_rd.as<uint64_t>(), S, gm, tkey, tcnt, H - 1);
			jn.xkey = tkey; jn.xcnt = tcnt; jn.xmask = H - 1;
		} else {                      // the count-min sketch summed over the ranks bounds every bin of the job
			MCB_TRY(mcb_coll_allreduce_sum_u32(ctx, b_cm.as<uint32_t>(), CM));
			jn.cmg = b_cm.as<uint32_t>(); jn.cmg_mask = CM - 1;
		}
		MCB_CUDA(cudaMemsetAsync(&dc[CT_S2_CAND], 0, 8, ctx->stream)); MCB_CUDA(cudaMemsetAsync(&dc[CT_S2_NEEDEXACT], 0, 8, ctx->stream));
		if (S) MCB_CUDA(cudaMemsetAsync(claim, 0x7F, S * 8, ctx->stream));
	}
	res->n_candidates = hc[CT_S2_CAND];
	return MCB_OK;
}

// second half (K8): compact the claims, order them, copy the lists to the host
static int realign_claims(mcb_ctx *ctx, mcb_realign_result *res)
{
	if (!ctx->rs.pending) { mcb_set_error("mcb_realign: no search pending"); return MCB_ESTATE; }
	ctx->rs.pending = false;
	const uint64_t S = ctx->rs.S;
	McbContigIndex &cx = ctx->cix;
	const uint64_t n_windows = cx.n_windows, n_contigs = cx.n_contigs;
	if (S == 0 || ctx->rs.nd == 0) return MCB_OK;
	unsigned long long *dc = ctx->d_counters.as<unsigned long long>();
	unsigned long long *hc = ctx->h_counters.as<unsigned long long>();
	unsigned long long *claim = ctx->d_x[0].as<unsigned long long>();
	const uint32_t *sidx = ctx->rs.have_index ? ctx->d_x[1].as<uint32_t>() : nullptr;
	DBuf &b_sg = ctx->d_scr[0], &b_fl3 = ctx->d_scr[3], &b_fp = ctx->d_scr[4], &b_fl = ctx->d_scr[8], &b_elA = ctx->d_scr[9], &b_elB = ctx->d_scr[10], &b_out = ctx->d_scr[11];
	MCB_TRY(b_fl3.ensure(S * 12 + 64));
	uint32_t *f_c = b_fl3.as<uint32_t>(), *f_a = f_c + S, *f_t = f_a + S;
	MCB_TRY(b_elA.ensure(S * 16 + 16)); MCB_TRY(b_elB.ensure(S * 16 + 16)); MCB_TRY(b_fp.ensure(S * 8 + 64));
	ulonglong2 *elA = b_elA.as<ulonglong2>(), *elB = b_elB.as<ulonglong2>();
	uint32_t *d_fpa = b_fp.as<uint32_t>(), *d_fpt = d_fpa + S;
	const int span_h = ctx->tm.begin("realign");
	MCB_LAUNCH(ctx, "s2_claim_flags", k_s2_claim_flags, mcb_grid_for(S, 256), 256, 0, claim, b_fl.as<uint8_t>(), S, f_c, f_a, f_t);
	MCB_TRY(mcb_exclusive_scan_u32(ctx, f_c, S, (uint64_t*)&dc[CT_S2_CLAIMS]));
	MCB_TRY(mcb_exclusive_scan_u32(ctx, f_a, S, (uint64_t*)&dc[CT_S2_FPA]));
	MCB_TRY(mcb_exclusive_scan_u32(ctx, f_t, S, (uint64_t*)&dc[CT_S2_FPT]));
	MCB_LAUNCH(ctx, "s2_claim_compact", k_s2_claim_compact, mcb_grid_for(S, 256), 256, 0, claim, b_fl.as<uint8_t>(), S, sidx, f_c, f_a, f_t, elA, d_fpa, d_fpt);
	MCB_CUDA(cudaMemcpyAsync(ctx->h_counters.p, dc, 64 * 8, cudaMemcpyDeviceToHost, ctx->stream));
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	const uint64_t ncl = hc[CT_S2_CLAIMS], nfa = hc[CT_S2_FPA], nft = hc[CT_S2_FPT];
	std::vector<McbSortPass> passes;
	mcb_add_bit_passes(passes, 1, 0, mcb_bits_for(S));                                // 0xFFFFFFFF-s: only the low bits vary
	mcb_add_bit_passes(passes, 0, 0, 5 + mcb_bits_for(n_windows));
	ulonglong2 *els = nullptr;
	MCB_TRY(mcb_radix_sort(ctx, elA, elB, ncl, passes.data(), (int)passes.size(), &els));
	MCB_TRY(b_out.ensure(ncl * 24 + 64));
	uint64_t *o_y = b_out.as<uint64_t>(), *o_p = o_y + ncl; uint32_t *o_c = (uint32_t*)(o_p + ncl), *o_s = o_c + ncl;
	if (ncl) MCB_LAUNCH(ctx, "s2_claim_emit", k_s2_claim_emit, mcb_grid_for(ncl, 256), 256, 0, els, ncl, cx.wo.as<uint64_t>(), n_contigs, b_sg.as<uint32_t>(), sidx, o_c, o_s, o_y, o_p);
	ctx->tm.end(span_h);
	MCB_TRY(ctx->h_claim_c.ensure(ncl * 4 + 16)); MCB_TRY(ctx->h_claim_s.ensure(ncl * 4 + 16)); MCB_TRY(ctx->h_claim_y.ensure(ncl * 8 + 16)); MCB_TRY(ctx->h_claim_p.ensure(ncl * 8 + 16));
	MCB_TRY(ctx->h_fpA.ensure(nfa * 4 + 16)); MCB_TRY(ctx->h_fpT.ensure(nft * 4 + 16));
	{
		McbSpan sp(ctx->tm, "d2h");
		if (ncl) {
			MCB_CUDA(cudaMemcpyAsync(ctx->h_claim_c.p, o_c, ncl * 4, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(ctx->h_claim_s.p, o_s, ncl * 4, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(ctx->h_claim_y.p, o_y, ncl * 8, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(ctx->h_claim_p.p, o_p, ncl * 8, cudaMemcpyDeviceToHost, ctx->stream));
		}
		if (nfa) MCB_CUDA(cudaMemcpyAsync(ctx->h_fpA.p, d_fpa, nfa * 4, cudaMemcpyDeviceToHost, ctx->stream));
		if (nft) MCB_CUDA(cudaMemcpyAsync(ctx->h_fpT.p, d_fpt, nft * 4, cudaMemcpyDeviceToHost, ctx->stream));
	}
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->tm.collect();
	res->n_claims = ncl; res->claim_contig = ctx->h_claim_c.as<uint32_t>(); res->claim_sg = ctx->h_claim_s.as<uint32_t>(); res->claim_y = ctx->h_claim_y.as<uint64_t>();
	res->claim_prio = ctx->h_claim_p.as<uint64_t>();
	res->n_fpA = nfa; res->n_fpT = nft; res->fpA_sg = ctx->h_fpA.as<uint32_t>(); res->fpT_sg = ctx->h_fpT.as<uint32_t>();
	return MCB_OK;
}

extern "C" int mcb_realign(mcb_ctx *ctx, const uint32_t *sg, uint64_t S, const char *refs, const uint64_t *ref_off, uint64_t n_contigs,
                           int threshold, int maxsearch, int ininumdict, mcb_realign_result *res)
{
	if (!ctx || !res) { mcb_set_error("mcb_realign: null argument"); return MCB_EINVAL; }
	MCB_CUDA(cudaSetDevice(ctx->prm.device));
	if (ctx->shard_n > 1) { mcb_set_error("mcb_realign: context is sharded, use mcb_shard_realign"); return MCB_ESTATE; }
	MCB_TRY(realign_search(ctx, sg, nullptr, S, S, refs, ref_off, n_contigs, threshold, maxsearch, ininumdict, res));
	return realign_claims(ctx, res);
}

// Sharded Stage 2: this rank's singles — a subsequence of the job's sg list, given with their positions in it —
// against ALL contigs.  Claims come back in the reference's append order restricted to these singles, with claim_sg / fpA_sg /
// fpT_sg as positions in the job's list and claim_prio as the merge key: the job's list is the ranks' lists merged by
// (claim_prio ascending, claim_sg descending).  Collective: every rank of the communicator must call it.
extern "C" int mcb_shard_realign(mcb_ctx *ctx, const uint32_t *sg, const uint32_t *sg_index, uint64_t n_sg_local, uint64_t n_sg_total,
                                 const char *refs, const uint64_t *ref_off, uint64_t n_contigs, int threshold, int maxsearch, int ininumdict, mcb_realign_result *res)
{
	if (!ctx || !res) { mcb_set_error("mcb_shard_realign: null argument"); return MCB_EINVAL; }
	MCB_CUDA(cudaSetDevice(ctx->prm.device));
	MCB_TRY(realign_search(ctx, sg, ctx->shard_n > 1 ? sg_index : (sg_index ? sg_index : nullptr), n_sg_local, n_sg_total, refs, ref_off, n_contigs, threshold, maxsearch, ininumdict, res));
	return realign_claims(ctx, res);
}

// The contig merge hands its result straight to Stage 2: the next mcb_realign calls with refs == NULL use these contigs.
int mcb_realign_prime_contigs(mcb_ctx *ctx, const char *d_refs, const uint64_t *h_ref_off, uint64_t n_contigs)
{
	const int lt = ctx->L <= 80 ? 11 : 17;               // realign_geometry
	if (ctx->shard_n > 1) return MCB_OK;
	return contig_index_update(ctx, nullptr, h_ref_off, n_contigs, lt, d_refs);
}
