// mcb_combine.cu — combine_cluster (kthread_cb.c:570-630) on the device: the contig merge that sits between kt_for_bucket and
// realign_hash.  SURVEY.md 8(f) N1.
//
// The reference walks the contigs in order; contig i sketches its consensus (all (w,k)-minimizers, kthread_cb.c:234), looks every
// minimizer up in the index of the first-m minimizers of all contigs (:269), and merges with the FIRST posting whose contig is
// still unmerged, lies on the same strand and whose consensus differs from i's in at most cbthreshold characters over their
// overlap (match_pro, :36-53, :289-290).  Both contigs are then flagged; merged contigs come first in the new set, the untouched
// ones are copied behind them (:474-494); the loop ends when the number of contigs changes by less than 100 (:625).
//
// Everything but the flags is independent of the order of processing: which postings contig i meets, in which order, and which
// of them pass the strand and match_pro tests.  So one iteration is
//   1  index of the current first-m tuples, built on the device (mcb_index_build_device: the reference's posting order)
//   2  all minimizers of every contig                   k_sketch_lh2 in its ALL mode, count + emit
//   3  (minimizer, posting) pairs in the reference's visiting order, strand / self test, match_pro on 2-bit packed contigs
//      -> per contig the ordered list of acceptable partners                      k_cb_lookup, k_cb_pairs, k_cb_match
//   4  the sequential part, which is tiny: walk the contigs in order, take the first partner that is still free
//      (= greedy matching by edge priority (i, position in i's list))             host loop over the compacted lists
//   5  the new contig set: member lists concatenated with the offset shift (:300-317) and stably sorted by (position, strand)
//      (construct_ref2's qsort, :107), consensus by column majority over the oriented reads (:120-150), untouched contigs
//      copied, first-m minimizers of every new contig                             k_cb_members, radix sort, k_cb_consensus, k_sketch_lh2
// The contig set, the packed reads and the index never leave the device; the host sees two small lists per iteration.
#include "mcb_common.cuh"
#include "mcb_lh.cuh"
#include <algorithm>
#include <limits.h>

struct CbSet {                       // one contig set on the device
	DBuf n, aoff, a, roff, ref;
	DBuf coff;                         // u64[ncl]: where the per-column base counts of the contig start in the arena McbCombineState::cols
	DBuf mins, moff;                   // ALL (w,k)-minimizers of every contig, contig-major in position order; moff u64[ncl+1].  The first
	                                   // first_mininum of a contig's list are its index tuples (kthread_cb.c:359-368: the same sketch, cut short)
	uint64_t ncl = 0, nmem = 0, nref = 0, nmin = 0;
	void release() { n.release(); aoff.release(); a.release(); roff.release(); ref.release(); coff.release(); mins.release(); moff.release(); }
};
struct McbCombineState {
	CbSet set[2];
	DBuf alive, best, pick, tup, tup2, boff, chn, chc, chs, cho, cho64, pcnt, cand, pass, plist, loff, cw, cwo, flag, pairs, src, keyA, keyB, len2, tmp32, slots;
	DBuf cols;                         // uint4 per contig column: how many member reads put A, C, G, T there.  Append-only over the iterations:
	uint64_t cols_used = 0;            // a merged contig's counts are the sum of its two parents' (construct_ref2 counts every member again,
	                                   // kthread_cb.c:120-137, and counting is additive), untouched contigs keep pointing at theirs
	HBuf h_boff, h_list, h_loff, h_pairs, h_flag, h_small;
	HBuf h_cl_n, h_cl_a_off, h_cl_a, h_cl_ref_off, h_cl_ref;
	void release()
	{
		set[0].release(); set[1].release(); cols.release(); cols_used = 0;
		DBuf *d[] = { &alive, &best, &pick, &tup, &tup2, &boff, &chn, &chc, &chs, &cho, &cho64, &pcnt, &cand, &pass, &plist, &loff, &cw, &cwo, &flag, &pairs, &src, &keyA, &keyB, &len2, &tmp32, &slots };
		for (auto b : d) b->release();
		HBuf *h[] = { &h_boff, &h_list, &h_loff, &h_pairs, &h_flag, &h_small, &h_cl_n, &h_cl_a_off, &h_cl_a, &h_cl_ref_off, &h_cl_ref };
		for (auto b : h) b->release();
	}
};
int mcb_realign_prime_contigs(mcb_ctx *ctx, const char *d_refs, const uint64_t *h_ref_off, uint64_t n_contigs);     // mcb_stage2.cu
void mcb_combine_release(mcb_ctx *ctx) { if (ctx->cb) { ctx->cb->release(); delete ctx->cb; ctx->cb = nullptr; } }

// ---------------------------------------------------------------- 1: tuples of the current set, in push order
__global__ void k_cb_first_m_counts(const uint64_t *__restrict__ moff, uint64_t ncl, int m, uint32_t *__restrict__ out)
{
	const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c < ncl) out[c] = (uint32_t)min((uint64_t)m, moff[c + 1] - moff[c]);
}
__global__ void k_cb_gather_tuples(const mcb_tuple *__restrict__ mins, const uint64_t *__restrict__ moff, const uint32_t *__restrict__ off, uint64_t ncl, int m, ulonglong2 *__restrict__ out)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= ncl * (uint64_t)m) return;
	const uint64_t c = i / m, j = i - c * m, b = moff[c];
	if (b + j < moff[c + 1]) { const mcb_tuple t = mins[b + j]; out[off[c] + j] = make_ulonglong2(t.x, t.y); }
}
// long contigs are sketched in stretches of CB_CHUNK bases, one thread each (k_sketch_lh2's chunk mode)
#define CB_CHUNK 192
__global__ void k_cb_chunk_counts(const uint64_t *__restrict__ roff, uint64_t count, int chunk, uint32_t *__restrict__ nch)
{
	const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c < count) { const uint64_t len = roff[c + 1] - roff[c]; nch[c] = chunk ? (uint32_t)max((uint64_t)1, (len + chunk - 1) / chunk) : 1u; }
}
__global__ void k_cb_chunk_table(const uint32_t *__restrict__ first, uint64_t count, uint64_t nch, int chunk, uint32_t *__restrict__ ch_contig, uint32_t *__restrict__ ch_start)
{
	const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nch) return;
	uint64_t lo = 0, hi = count;              // last contig whose first chunk is <= t
	while (hi - lo > 1) { const uint64_t mid = (lo + hi) >> 1; if (first[mid] <= t) lo = mid; else hi = mid; }
	ch_contig[t] = (uint32_t)lo; ch_start[t] = (uint32_t)(t - first[lo]) * (uint32_t)chunk;
}
// list length of a contig = the tuples of its stretches
__global__ void k_cb_contig_counts(const uint32_t *__restrict__ first, uint64_t count, uint64_t nch, const uint32_t *__restrict__ choff, const unsigned long long *__restrict__ total, uint32_t *__restrict__ cnt)
{
	const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= count) return;
	const uint32_t a = choff[first[c]], b = c + 1 < count ? choff[first[c + 1]] : (uint32_t)*total;
	cnt[c] = b - a;
}
__global__ void k_cb_widen(const uint32_t *__restrict__ in, uint64_t n, uint64_t *__restrict__ out, const unsigned long long *__restrict__ total, uint32_t cap, unsigned long long *__restrict__ overflow)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	out[i] = in[i];
	const uint32_t cnt = (i + 1 < n ? in[i + 1] : (uint32_t)*total) - in[i];           // `in` is the exclusive scan of the per-stretch counts
	if (cap && cnt > cap) atomicAdd(overflow, 1ull);
}
// the tuples a counting pass parked in fixed-size slots (k_sketch_lh2 with slot_cap) move to their place in the contig-major list
__global__ void k_cb_unslot(const mcb_tuple *__restrict__ slots, uint32_t cap, const uint32_t *__restrict__ off, uint64_t nch, const unsigned long long *__restrict__ total, mcb_tuple *__restrict__ mins)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t t = i / cap; const uint32_t j = (uint32_t)(i - t * cap);
	if (t >= nch) return;
	const uint32_t o = off[t], cnt = (t + 1 < nch ? off[t + 1] : (uint32_t)*total) - o;
	if (j < cnt) mins[(uint64_t)o + j] = slots[t * cap + j];
}
// minimizer lists of the untouched contigs move to the new set with their new contig id
__global__ void k_cb_copy_min_counts(uint64_t nm, const uint32_t *__restrict__ src, uint64_t n2, const uint64_t *__restrict__ moff, uint32_t *__restrict__ cnt2)
{
	const uint64_t q = nm + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q < n2) { const uint32_t s = src[q - nm]; cnt2[q] = (uint32_t)(moff[s + 1] - moff[s]); }
}
__global__ void k_cb_copy_mins(uint64_t nm, const uint32_t *__restrict__ src, uint64_t n2, const uint64_t *__restrict__ moff, const mcb_tuple *__restrict__ mins,
                               const uint64_t *__restrict__ moff2, mcb_tuple *__restrict__ mins2)
{
	const uint64_t q = nm + (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3);           // eight lanes per contig
	const int lane = threadIdx.x & 7;
	if (q >= n2) return;
	const uint32_t s = src[q - nm];
	const uint64_t b = moff[s], cnt = moff[s + 1] - b, o = moff2[q];
	for (uint64_t u = lane; u < cnt; u += 8) { mcb_tuple t = mins[b + u]; t.y = ((q << 8) << 32) | (uint64_t)(uint32_t)t.y; mins2[o + u] = t; }
}

__global__ void k_cb_bucket_bounds(const ulonglong2 *__restrict__ t, uint64_t n, int nb, uint64_t *__restrict__ boff)
{
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b > nb) return;
	uint64_t lo = 0, hi = n;                  // first tuple whose bucket is >= b
	while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if ((int)(t[mid].x & (uint64_t)(nb - 1)) < b) lo = mid + 1; else hi = mid; }
	boff[b] = lo;
}

// ---------------------------------------------------------------- 3: contigs packed 2 bits per base (as in Stage 2), lookups, match_pro
__global__ void k_cb_word_offsets(const uint64_t *__restrict__ roff, uint64_t ncl, uint64_t *__restrict__ wcnt)
{
	const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c < ncl) wcnt[c] = (roff[c + 1] - roff[c] + 31) / 32 + 1;
}
__global__ void k_cb_pack(const char *__restrict__ ref, const uint64_t *__restrict__ roff, const uint64_t *__restrict__ cwo, uint64_t ncl, uint64_t total_words, uint64_t *__restrict__ cw)
{
	const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= total_words) return;
	uint64_t lo = 0, hi = ncl;                // last c with cwo[c] <= q
	while (hi - lo > 1) { const uint64_t mid = (lo + hi) >> 1; if (cwo[mid] <= q) lo = mid; else hi = mid; }
	const uint64_t c = lo, wq = q - cwo[c], b = roff[c], len = roff[c + 1] - b;
	uint64_t v = 0;
	for (int j = 0; j < 32; ++j) {
		const uint64_t p = wq * 32 + j;
		if (p < len) v |= (uint64_t)(mcb_code_of((unsigned char)ref[b + p]) & 3u) << (2 * j);
	}
	cw[q] = v;
}
// postings of one minimizer in the device index (mm_idx_get, kthread_idx.c:84-101)
__device__ __forceinline__ uint32_t cb_lookup(const McbDeviceIndex &ix, uint64_t x, uint32_t *first)
{
	const uint32_t bk = (uint32_t)(x & ((1ull << ix.b) - 1));
	uint32_t lo = ix.ub[bk], hi = ix.ub[bk + 1];
	while (lo < hi) {
		const uint32_t mid = lo + ((hi - lo) >> 1);
		const uint64_t kx = ix.keys[mid];
		if (kx < x) lo = mid + 1; else if (kx > x) hi = mid; else { *first = ix.kstart[mid]; return ix.kstart[mid + 1] - ix.kstart[mid]; }
	}
	*first = 0;
	return 0;
}
__global__ void k_cb_lookup(const mcb_tuple *__restrict__ mins, uint64_t M, McbDeviceIndex ix, uint32_t *__restrict__ pcnt, uint32_t *__restrict__ pfirst)
{
	const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= M) return;
	uint32_t first;
	pcnt[t] = cb_lookup(ix, mins[t].x, &first);
	pfirst[t] = first;
}
// one record per (minimizer of contig i, posting): the partner contig and the two anchor positions; partner = ~0 when the posting
// is contig i itself or lies on the other strand (kthread_cb.c:276-289)
struct CbCand { uint32_t i, c, pos_ori, pos; };
__global__ void k_cb_pairs(const mcb_tuple *__restrict__ mins, uint64_t M, McbDeviceIndex ix, const uint32_t *__restrict__ poff, const uint32_t *__restrict__ pfirst,
                           const uint32_t *__restrict__ pcnt, CbCand *__restrict__ cand)
{
	const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= M) return;
	const mcb_tuple mn = mins[t];
	const uint32_t rid_ori = (uint32_t)(mn.y >> 32), pos_ori = (uint32_t)mn.y >> 1, dir_ori = (uint32_t)mn.y & 1u;
	const uint32_t n = pcnt[t], f = pfirst[t], o = poff[t];
	for (uint32_t q = 0; q < n; ++q) {
		const uint64_t y = ix.post[f + q];
		const uint32_t rid = (uint32_t)(y >> 32);
		CbCand r; r.i = rid_ori >> 8; r.pos_ori = pos_ori; r.pos = (uint32_t)y >> 1;
		r.c = (rid != rid_ori && ((uint32_t)y & 1u) == dir_ori) ? rid >> 8 : 0xFFFFFFFFu;
		cand[o + q] = r;
	}
}
// match_pro (kthread_cb.c:36-53): mismatching characters over the overlap of the two consensus strings aligned at (pos_ori, pos)
__device__ __forceinline__ uint64_t cb_bases32(const uint64_t *__restrict__ w, int64_t base)       // 32 bases starting at `base` >= 0 (guard word behind every contig)
{
	const int64_t wi = base >> 5; const int sh = 2 * (int)(base & 31);
	const uint64_t a = w[wi];
	return sh ? (a >> sh) | (w[wi + 1] << (64 - sh)) : a;
}
__global__ void k_cb_match(const CbCand *__restrict__ cand, uint64_t n, const uint64_t *__restrict__ cw, const uint64_t *__restrict__ cwo, const uint64_t *__restrict__ roff,
                           int cbthr, uint32_t *__restrict__ pass)
{
	const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n) return;
	const CbCand r = cand[e];
	uint32_t ok = 0;
	if (r.c != 0xFFFFFFFFu) {
		const int64_t li = (int64_t)(roff[r.i + 1] - roff[r.i]), lc = (int64_t)(roff[r.c + 1] - roff[r.c]);
		const int64_t d = (int64_t)r.pos_ori - (int64_t)r.pos;                 // column x of contig i meets column x - d of contig c
		const int64_t lo = d > 0 ? d : 0, hi = li < lc + d ? li : lc + d;
		const uint64_t *wi = cw + cwo[r.i], *wc = cw + cwo[r.c];
		int mism = 0;
		for (int64_t x = lo; x < hi && mism <= cbthr; x += 32) {
			uint64_t v = cb_bases32(wi, x) ^ cb_bases32(wc, x - d);
			v = (v | (v >> 1)) & 0x5555555555555555ull;
			if (hi - x < 32) v &= (1ull << (2 * (hi - x))) - 1;
			mism += __popcll(v);
		}
		ok = mism <= cbthr;
	}
	pass[e] = ok;
}
__global__ void k_cb_compact(const CbCand *__restrict__ cand, const uint32_t *__restrict__ pass_scan, const uint32_t *__restrict__ passed, uint64_t n, CbCand *__restrict__ out)
{
	const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e < n && passed[e]) out[pass_scan[e]] = cand[e];
}
__global__ void k_cb_list_bounds(const CbCand *__restrict__ list, uint64_t P, uint64_t ncl, uint32_t *__restrict__ loff)
{
	const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c > ncl) return;
	uint64_t lo = 0, hi = P;                  // first entry whose contig is >= c
	while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (list[mid].i < c) lo = mid + 1; else hi = mid; }
	loff[c] = (uint32_t)lo;
}

// a partner that already appears earlier in the same contig's list can never be the pick (if it is free the earlier entry takes it)
__global__ void k_cb_dedupe(const CbCand *__restrict__ list, uint64_t P, uint32_t *__restrict__ keep)
{
	const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= P) return;
	const CbCand r = list[e];
	uint32_t k = 1;
	for (uint64_t f = e; f-- > 0 && list[f].i == r.i;) if (list[f].c == r.c) { k = 0; break; }
	keep[e] = k;
}

// ---------------------------------------------------------------- 4: first-come resolution
// The reference takes the contigs in order and gives contig i the first entry of its list whose partner is still free; both
// are then taken.  That is greedy matching over the entries in list order (entry e of the concatenated lists has priority e):
// an entry is accepted iff no earlier accepted entry touches either of its contigs.  Equivalent rounds: every live entry
// bids for both of its contigs with its index (atomicMin); an entry that holds the minimum at BOTH contigs cannot be
// pre-empted by anything earlier and is accepted; entries that touch a taken contig die.  The earliest live entry always
// wins its round, so the loop ends, and contig ids are in hash order, so dependency chains stay short (tens of rounds).
#define CB_NONE 0xFFFFFFFFu
__global__ void k_cb_bid(const CbCand *__restrict__ list, uint64_t P, const uint8_t *__restrict__ taken, uint8_t *__restrict__ alive, uint32_t *__restrict__ best,
                         unsigned long long *__restrict__ n_live)
{
	const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	bool live = false;
	if (e < P && alive[e]) {
		const CbCand r = list[e];
		if (taken[r.i] || taken[r.c]) alive[e] = 0;
		else { atomicMin(&best[r.i], (uint32_t)e); atomicMin(&best[r.c], (uint32_t)e); live = true; }
	}
	const unsigned m = __ballot_sync(0xFFFFFFFFu, live);
	if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_live, (unsigned long long)__popc(m));
}
__global__ void k_cb_accept(const CbCand *__restrict__ list, uint64_t P, uint8_t *__restrict__ taken, uint8_t *__restrict__ alive, const uint32_t *__restrict__ best, uint32_t *__restrict__ pick)
{
	const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= P || !alive[e]) return;
	const CbCand r = list[e];
	if (best[r.i] == (uint32_t)e && best[r.c] == (uint32_t)e) { taken[r.i] = 1; taken[r.c] = 1; pick[r.i] = (uint32_t)e; alive[e] = 0; }
}
__global__ void k_cb_round_reset(const CbCand *__restrict__ list, uint64_t P, const uint8_t *__restrict__ alive, uint32_t *__restrict__ best)
{
	const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (e < P && alive[e]) { best[list[e].i] = CB_NONE; best[list[e].c] = CB_NONE; }
}
// the outcome as two ordered lists: the accepted entries in the order of their first contig, the untouched contigs in order
__global__ void k_cb_outcome_flags(const uint32_t *__restrict__ pick, const uint8_t *__restrict__ taken, uint64_t ncl, uint32_t *__restrict__ f_pick, uint32_t *__restrict__ f_free)
{
	const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c < ncl) { f_pick[c] = pick[c] != CB_NONE; f_free[c] = !taken[c]; }
}
__global__ void k_cb_outcome(const CbCand *__restrict__ list, const uint32_t *__restrict__ pick, const uint8_t *__restrict__ taken, uint64_t ncl,
                             const uint32_t *__restrict__ s_pick, const uint32_t *__restrict__ s_free, CbCand *__restrict__ pairs, uint32_t *__restrict__ src)
{
	const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= ncl) return;
	if (pick[c] != CB_NONE) pairs[s_pick[c]] = list[pick[c]];
	if (!taken[c]) src[s_free[c]] = (uint32_t)c;
}

// ---------------------------------------------------------------- 5: the new contig set
// new contig q < nm is the merge pairs[q] = (i, c, pos_ori, pos); q >= nm is the copy of old contig src[q - nm]
__global__ void k_cb_new_sizes(const CbCand *__restrict__ pairs, uint64_t nm, const uint32_t *__restrict__ src, uint64_t n2, const uint32_t *__restrict__ cl_n, uint32_t *__restrict__ n_new)
{
	const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= n2) return;
	n_new[q] = q < nm ? cl_n[pairs[q].i] + cl_n[pairs[q].c] : cl_n[src[q - nm]];
}
__global__ void k_cb_prefix64(const uint32_t *__restrict__ scan32, uint64_t n, uint64_t *__restrict__ out, const unsigned long long *__restrict__ total)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = scan32[i];
	if (i == n) out[n] = *total;
}
// one warp per new contig: members of the first contig as they are, those of the second shifted by the anchor difference
// (kthread_cb.c:300-317); merged contigs also get the sort key (new contig, position, strand) of every member
__global__ void k_cb_members(const CbCand *__restrict__ pairs, uint64_t nm, const uint32_t *__restrict__ src, uint64_t n2, const uint64_t *__restrict__ aoff, const uint64_t *__restrict__ a,
                             const uint64_t *__restrict__ aoff2, uint64_t *__restrict__ a2, ulonglong2 *__restrict__ keyed)
{
	const uint64_t q = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (q >= n2) return;
	const uint64_t o = aoff2[q];
	if (q >= nm) {
		const uint32_t s = src[q - nm];
		const uint64_t b = aoff[s], cnt = aoff[s + 1] - b;
		for (uint64_t u = lane; u < cnt; u += 32) a2[o + u] = a[b + u];
		return;
	}
	const CbCand p = pairs[q];
	const bool i_first = p.pos_ori >= p.pos;
	const uint32_t first = i_first ? p.i : p.c, second = i_first ? p.c : p.i;
	const uint64_t shift = i_first ? p.pos_ori - p.pos : p.pos - p.pos_ori;
	const uint64_t b1 = aoff[first], n1 = aoff[first + 1] - b1, b2 = aoff[second], n2m = aoff[second + 1] - b2;
	for (uint64_t u = lane; u < n1 + n2m; u += 32) {
		uint64_t y;
		if (u < n1) y = a[b1 + u];
		else { const uint64_t v = a[b2 + u - n1]; y = (v >> 32 << 32) | ((((uint64_t)((uint32_t)v >> 1)) + shift) << 1) | (v & 1); }
		keyed[o + u] = make_ulonglong2((q << 32) | (uint64_t)(uint32_t)y, y);          // (new contig, position << 1 | strand)
	}
}
__global__ void k_cb_unkey(const ulonglong2 *__restrict__ keyed, uint64_t n, uint64_t *__restrict__ a2)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) a2[i] = keyed[i].y;
}
// consensus lengths: a merged contig ends with its right-most member (members sorted by position), a copy keeps its length
__global__ void k_cb_new_lengths(uint64_t nm, const uint32_t *__restrict__ src, uint64_t n2, const uint64_t *__restrict__ aoff2, const uint64_t *__restrict__ a2,
                                 const uint64_t *__restrict__ roff, int L, uint64_t *__restrict__ len2, unsigned long long *__restrict__ max_len)
{
	const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long len = 0;
	if (q < n2) {
		if (q < nm) len = (uint64_t)((uint32_t)a2[aoff2[q + 1] - 1] >> 1) + (uint64_t)L;
		else { const uint32_t s = src[q - nm]; len = roff[s + 1] - roff[s]; }
		len2[q] = len;
	}
	for (int o = 16; o; o >>= 1) len = max(len, __shfl_xor_sync(0xFFFFFFFFu, len, o));
	if ((threadIdx.x & 31) == 0 && len) atomicMax(max_len, len);
}
// construct_ref2 (kthread_cb.c:105-150) recounts, for every column of a merged contig, the bases its oriented members put there;
// the consensus is 'A' unless a base has strictly more votes, in the order A, C, G, T.  Counting is additive, so the counts of a
// merged contig are the sums of its parents' counts, the second parent shifted by the anchor difference (:300-317): one thread per
// column adds two uint4 and takes the majority, whatever the coverage.  The counts of the seed contigs are made once (below).
__device__ __forceinline__ char cb_majority(uint4 v)
{
	unsigned best = 0, mx = v.x;
	if (v.y > mx) { mx = v.y; best = 1; }
	if (v.z > mx) { mx = v.z; best = 2; }
	if (v.w > mx) { mx = v.w; best = 3; }
	return "ACGT"[best];
}
// one warp per merged contig: lanes stride over its columns, so the parents' counts arrive as coalesced 16-byte loads
__global__ void __launch_bounds__(256)
k_cb_merge_counts(const CbCand *__restrict__ pairs, uint64_t nm, const uint64_t *__restrict__ roff2, const uint64_t *__restrict__ roff, const uint64_t *__restrict__ coff,
                  uint4 *__restrict__ cols, uint64_t new_base, char *__restrict__ ref2)
{
	const uint64_t q = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (q >= nm) return;
	const CbCand p = pairs[q];
	const bool i_first = p.pos_ori >= p.pos;
	const uint32_t first = i_first ? p.i : p.c, second = i_first ? p.c : p.i;
	const uint64_t shift = i_first ? p.pos_ori - p.pos : p.pos - p.pos_ori;
	const uint64_t l1 = roff[first + 1] - roff[first], l2 = roff[second + 1] - roff[second];
	const uint64_t o2 = roff2[q], len = roff2[q + 1] - o2;
	const uint4 *c1 = cols + coff[first], *c2 = cols + coff[second];
	for (uint64_t col = lane; col < len; col += 32) {
		uint4 v = make_uint4(0, 0, 0, 0);
		if (col < l1) v = c1[col];
		if (col >= shift && col - shift < l2) { const uint4 u = c2[col - shift]; v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
		cols[new_base + o2 + col] = v;
		ref2[o2 + col] = cb_majority(v);
	}
}
__global__ void k_cb_new_coff(uint64_t nm, const uint32_t *__restrict__ src, uint64_t n2, const uint64_t *__restrict__ roff2, const uint64_t *__restrict__ coff, uint64_t new_base, uint64_t *__restrict__ coff2)
{
	const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q < n2) coff2[q] = q < nm ? new_base + roff2[q] : coff[src[q - nm]];
}
// the counts of the seed contigs: one warp per contig takes the members in turn, its lanes the bases of the member, and votes
// into a table in shared memory (a member's bases fall into distinct columns, so the lanes never meet; members are serialised
// by __syncwarp).  Column g of the set's concatenated consensus strings = arena entry g (coff = roff).
#define CB_SEED_WARPS 4
__global__ void __launch_bounds__(CB_SEED_WARPS * 32)
k_cb_seed_counts(uint64_t ncl, const uint64_t *__restrict__ roff, const uint64_t *__restrict__ aoff, const uint64_t *__restrict__ a,
                 const uint64_t *__restrict__ packed, int WS, int L, int maxcols, uint4 *__restrict__ cols, unsigned long long *__restrict__ err)
{
	extern __shared__ uint32_t cb_votes[];                       // [CB_SEED_WARPS][4][maxcols]
	const int wrp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint64_t q = (uint64_t)blockIdx.x * CB_SEED_WARPS + wrp;
	if (q >= ncl) return;
	uint32_t *v = cb_votes + (size_t)wrp * 4 * maxcols;
	const uint64_t o = roff[q];
	const int len = (int)(roff[q + 1] - o);
	if (len > maxcols) { if (lane == 0) atomicAdd(err, 1ull); return; }
	for (int i = lane; i < 4 * maxcols; i += 32) v[i] = 0;
	__syncwarp();
	const uint64_t mb = aoff[q], me = aoff[q + 1];
	for (uint64_t u = mb; u < me; ++u) {
		const uint64_t y = a[u];
		const int pos = (int)((uint32_t)y >> 1);
		const uint64_t *row = packed + (y >> 32) * (uint64_t)WS;
		if (pos + L <= len)
			for (int b = lane; b < L; b += 32) {
				const unsigned bse = (y & 1) ? 3u - mcb_base_at(row, L - 1 - b) : mcb_base_at(row, b);
				v[bse * maxcols + pos + b] += 1;
			}
		else if (lane == 0) atomicAdd(err, 1ull);
		__syncwarp();
	}
	for (int c = lane; c < len; c += 32) cols[o + c] = make_uint4(v[c], v[maxcols + c], v[2 * maxcols + c], v[3 * maxcols + c]);
}
__global__ void k_cb_copy_refs(uint64_t nm, const uint32_t *__restrict__ src, uint64_t n2, const uint64_t *__restrict__ roff, const char *__restrict__ ref,
                               const uint64_t *__restrict__ roff2, char *__restrict__ ref2)
{
	const uint64_t q = nm + (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
	const int lane = threadIdx.x & 31;
	if (q >= n2) return;
	const uint32_t s = src[q - nm];
	const uint64_t b = roff[s], len = roff[s + 1] - b, o = roff2[q];
	for (uint64_t u = lane; u < len; u += 32) ref2[o + u] = ref[b + u];
}

// ================================================================= host
// the count arena grows without losing what it holds
static int cb_cols_reserve(mcb_ctx *ctx, McbCombineState &cb, uint64_t columns)
{
	const size_t need = (size_t)columns * 16 + 16;
	if (need <= cb.cols.cap) return MCB_OK;
	DBuf bigger;
	MCB_TRY(bigger.ensure(need + need / 2));
	if (cb.cols_used) MCB_CUDA(cudaMemcpyAsync(bigger.p, cb.cols.p, (size_t)cb.cols_used * 16, cudaMemcpyDeviceToDevice, ctx->stream));
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	cb.cols.release();
	cb.cols = bigger;
	return MCB_OK;
}

static int cb_counters(mcb_ctx *ctx)          // the device scalars to the host (synchronizes)
{
	MCB_CUDA(cudaMemcpyAsync(ctx->h_counters.p, ctx->d_counters.p, 64 * 8, cudaMemcpyDeviceToHost, ctx->stream));
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	return MCB_OK;
}
enum { CT_CB_A = 48, CT_CB_B = 49, CT_CB_C = 50, CT_CB_D = 51, CT_CB_E = 52, CT_CB_F = 53 };          // scratch slots of ctx->d_counters (32..41 are the consensus work lists)
#define CB_HC(ctx, slot) ((ctx)->h_counters.as<unsigned long long>()[slot])

// all (w,k)-minimizers (mm_sketch_lh_ori, sketch.c:116-165; window rw, kthread_cb.c:234 / :359 with win_step = 0) of the work items
// [0, n_items) of a set: whole contigs, or stretches of contigs (ch_* given).  cnt32 receives the number of tuples per item; with
// `emit` they are written to S.mins at off[item]
static int cb_sketch(mcb_ctx *ctx, CbSet &S, uint64_t n_items, uint32_t *cnt32, const uint64_t *off, const uint32_t *ch_contig, const uint32_t *ch_start, int ch_len,
                     mcb_tuple *slots = nullptr, int slot_cap = 0)
{
	if (!n_items) return MCB_OK;
	const int rw = ctx->prm.rw, k = ctx->prm.k;
	const size_t smem = (size_t)rw * LH_THREADS * 13;
	auto kern = (k > 16 && k < 32) ? k_sketch_lh2<true> : k_sketch_lh2<false>;
	if (smem > 48 * 1024) MCB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	MCB_LAUNCH(ctx, "cb_sketch", kern, mcb_grid_for(n_items, LH_THREADS), LH_THREADS, smem, S.ref.as<char>(), S.roff.as<uint64_t>(), (uint64_t)0, n_items, (uint64_t)0,
	           rw, k, INT_MAX, slots ? slots : off ? S.mins.as<mcb_tuple>() : (mcb_tuple*)nullptr, (uint8_t*)nullptr, off, cnt32, ch_contig, ch_start, ch_len, slots ? slot_cap : 0);
	return MCB_OK;
}

// minimizer lists of a new set: the merged contigs [0, nm) are sketched — in stretches, so that a contig of thousands of bases is
// not one thread's sequential walk — and the untouched ones take their lists along (d_src = their old ids; the seed set has
// nm == ncl and sketches everything)
static int cb_min_lists(mcb_ctx *ctx, McbCombineState &cb, CbSet &S, uint64_t nm, const CbSet *old, const uint32_t *d_src)
{
	unsigned long long *dc = ctx->d_counters.as<unsigned long long>();
	const uint64_t n2 = S.ncl;
	const int chunk = (ctx->prm.k & 1) ? CB_CHUNK : 0;          // stretches need odd k (no k-mer equal to its reverse complement); else whole contigs
	MCB_TRY(cb.tmp32.ensure((n2 + 2) * 4)); MCB_TRY(S.moff.ensure((n2 + 2) * 8)); MCB_TRY(cb.chn.ensure((nm + 2) * 4));
	uint32_t *cnt = cb.tmp32.as<uint32_t>(), *first = cb.chn.as<uint32_t>();
	uint64_t nch = 0;
	// one sketch pass instead of two: the counting pass parks the tuples of every stretch in a slot of `cap` tuples (three times what
	// the window density 2/(w+1) predicts); they are moved once the counts are scanned.  A stretch that needs more than a slot (long
	// runs of equal hashes) makes the call fall back to the second, emitting pass.
	uint32_t cap = (uint32_t)std::min(CB_CHUNK, std::max(16, 3 * CB_CHUNK / (ctx->prm.rw + 1) + 8));
	if (const char *e = getenv("MCB_CB_SLOTCAP")) cap = (uint32_t)std::max(1, atoi(e));          // tests: a tiny slot forces the fallback
	const bool use_slots = chunk != 0 && !getenv("MCB_CB_TWOPASS");
	if (nm) {
		MCB_LAUNCH(ctx, "cb_chunk_counts", k_cb_chunk_counts, mcb_grid_for(nm, 256), 256, 0, S.roff.as<uint64_t>(), nm, chunk, first);
		MCB_TRY(mcb_exclusive_scan_u32(ctx, first, nm, (uint64_t*)&dc[CT_CB_A]));
		MCB_TRY(cb_counters(ctx));
		nch = CB_HC(ctx, CT_CB_A);
		MCB_TRY(cb.chc.ensure(nch * 4 + 16)); MCB_TRY(cb.chs.ensure(nch * 4 + 16)); MCB_TRY(cb.cho.ensure((nch + 2) * 4)); MCB_TRY(cb.cho64.ensure((nch + 2) * 8));
		MCB_LAUNCH(ctx, "cb_chunk_table", k_cb_chunk_table, mcb_grid_for(nch, 256), 256, 0, first, nm, nch, chunk ? chunk : INT_MAX, cb.chc.as<uint32_t>(), cb.chs.as<uint32_t>());
		if (use_slots) MCB_TRY(cb.slots.ensure(nch * cap * 16 + 16));
		MCB_TRY(cb_sketch(ctx, S, nch, cb.cho.as<uint32_t>(), nullptr, cb.chc.as<uint32_t>(), cb.chs.as<uint32_t>(), chunk ? chunk : INT_MAX, use_slots ? cb.slots.as<mcb_tuple>() : nullptr, (int)cap));
		MCB_TRY(mcb_exclusive_scan_u32(ctx, cb.cho.as<uint32_t>(), nch, (uint64_t*)&dc[CT_CB_B]));
		MCB_LAUNCH(ctx, "cb_contig_counts", k_cb_contig_counts, mcb_grid_for(nm, 256), 256, 0, first, nm, nch, cb.cho.as<uint32_t>(), &dc[CT_CB_B], cnt);
		MCB_CUDA(cudaMemsetAsync(&dc[CT_CB_C], 0, 8, ctx->stream));
		MCB_LAUNCH(ctx, "cb_widen", k_cb_widen, mcb_grid_for(nch, 256), 256, 0, cb.cho.as<uint32_t>(), nch, cb.cho64.as<uint64_t>(), &dc[CT_CB_B], use_slots ? cap : 0u, &dc[CT_CB_C]);
	}
	if (n2 > nm) MCB_LAUNCH(ctx, "cb_copy_min_counts", k_cb_copy_min_counts, mcb_grid_for(n2 - nm, 256), 256, 0, nm, d_src, n2, old->moff.as<uint64_t>(), cnt);
	MCB_TRY(mcb_exclusive_scan_u32(ctx, cnt, n2, (uint64_t*)&dc[CT_CB_A]));
	MCB_LAUNCH(ctx, "cb_prefix64", k_cb_prefix64, mcb_grid_for(n2 + 1, 256), 256, 0, cnt, n2, S.moff.as<uint64_t>(), &dc[CT_CB_A]);
	MCB_TRY(cb_counters(ctx));
	S.nmin = CB_HC(ctx, CT_CB_A);
	if (S.nmin >= 0xFFFFFFFFull) { mcb_set_error("mcb_combine: too many contig minimizers"); return MCB_EINVAL; }
	MCB_TRY(S.mins.ensure(S.nmin * 16 + 16));
	// the merged contigs come first in the set, so the tuples of stretch t start at cho[t] in the set's list as well
	if (nm && use_slots && CB_HC(ctx, CT_CB_C) == 0)
		MCB_LAUNCH(ctx, "cb_unslot", k_cb_unslot, mcb_grid_for(nch * cap, 256), 256, 0, cb.slots.as<mcb_tuple>(), cap, cb.cho.as<uint32_t>(), nch, &dc[CT_CB_B], S.mins.as<mcb_tuple>());
	else if (nm) MCB_TRY(cb_sketch(ctx, S, nch, cb.cho.as<uint32_t>() /* free again: rewritten with the lengths */, cb.cho64.as<uint64_t>(), cb.chc.as<uint32_t>(), cb.chs.as<uint32_t>(), chunk ? chunk : INT_MAX));
	if (n2 > nm) MCB_LAUNCH(ctx, "cb_copy_mins", k_cb_copy_mins, mcb_grid_for((n2 - nm) * 8, 256), 256, 0, nm, d_src, n2, old->moff.as<uint64_t>(), old->mins.as<mcb_tuple>(),
	                        S.moff.as<uint64_t>(), S.mins.as<mcb_tuple>());
	return MCB_OK;
}

static int combine_iteration(mcb_ctx *ctx, McbCombineState &cb, CbSet &cur, CbSet &nxt, int cbthr, uint64_t *n_merged, uint64_t *n_tuples, uint64_t *max_len)
{
	const int L = ctx->L, WS = ctx->WS, m = ctx->prm.first_mininum, nb = 1 << ctx->prm.b;
	unsigned long long *dc = ctx->d_counters.as<unsigned long long>();
	const uint64_t ncl = cur.ncl, M = cur.nmin;
	*n_merged = 0;
	// ---- 1: index of the first-m tuples (mm_idx_generation, kthread_cb.c:574), and meanwhile-independent preparations
	uint64_t T = 0, total_words = 0, C = 0;
	McbDeviceIndex ix;
	MCB_TRY(cb.pcnt.ensure((M + 2) * 12));
	uint32_t *pcnt = cb.pcnt.as<uint32_t>(), *pfirst = pcnt + (M + 1), *poff = pfirst + (M + 1);
	{
		McbSpan sp(ctx->tm, "combine");
		MCB_TRY(cb.tmp32.ensure((ncl + 2) * 4)); MCB_TRY(cb.cwo.ensure((ncl + 2) * 8));
		MCB_LAUNCH(ctx, "cb_first_m_counts", k_cb_first_m_counts, mcb_grid_for(ncl, 256), 256, 0, cur.moff.as<uint64_t>(), ncl, m, cb.tmp32.as<uint32_t>());
		MCB_TRY(mcb_exclusive_scan_u32(ctx, cb.tmp32.as<uint32_t>(), ncl, (uint64_t*)&dc[CT_CB_A]));
		MCB_LAUNCH(ctx, "cb_word_offsets", k_cb_word_offsets, mcb_grid_for(ncl, 256), 256, 0, cur.roff.as<uint64_t>(), ncl, cb.cwo.as<uint64_t>());
		MCB_TRY(mcb_exclusive_scan_u64(ctx, cb.cwo.as<uint64_t>(), ncl, (uint64_t*)&dc[CT_CB_B]));
		MCB_TRY(cb_counters(ctx));
		T = CB_HC(ctx, CT_CB_A); total_words = CB_HC(ctx, CT_CB_B);
		*n_tuples = T;
		MCB_TRY(cb.tup.ensure(T * 16 + 16)); MCB_TRY(cb.tup2.ensure(T * 16 + 16)); MCB_TRY(cb.boff.ensure(((size_t)nb + 2) * 8)); MCB_TRY(cb.h_boff.ensure(((size_t)nb + 2) * 8));
		MCB_LAUNCH(ctx, "cb_gather_tuples", k_cb_gather_tuples, mcb_grid_for(ncl * m, 256), 256, 0, cur.mins.as<mcb_tuple>(), cur.moff.as<uint64_t>(), cb.tmp32.as<uint32_t>(), ncl, m, cb.tup.as<ulonglong2>());
		McbSortPass bp[2] = { {0, 0, 7}, {0, 7, 7} };        // stable by bucket: inside a bucket the tuples keep their push order
		ulonglong2 *sorted = nullptr;
		MCB_TRY(mcb_radix_sort(ctx, cb.tup.as<ulonglong2>(), cb.tup2.as<ulonglong2>(), T, bp, 2, &sorted));
		if (sorted != cb.tup.as<ulonglong2>()) std::swap(cb.tup, cb.tup2);
		MCB_LAUNCH(ctx, "cb_bucket_bounds", k_cb_bucket_bounds, (nb + 1 + 255) / 256, 256, 0, cb.tup.as<ulonglong2>(), T, nb, cb.boff.as<uint64_t>());
		MCB_CUDA(cudaMemcpyAsync(cb.h_boff.p, cb.boff.p, ((size_t)nb + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
		// the packed contigs do not depend on the index: queued before the host waits for the bucket bounds
		MCB_TRY(cb.cw.ensure((total_words + 2) * 8));
		MCB_CUDA(cudaMemsetAsync(cb.cw.as<uint64_t>() + total_words, 0, 16, ctx->stream));
		if (total_words) MCB_LAUNCH(ctx, "cb_pack", k_cb_pack, mcb_grid_for(total_words, 256), 256, 0, cur.ref.as<char>(), cur.roff.as<uint64_t>(), cb.cwo.as<uint64_t>(), ncl, total_words, cb.cw.as<uint64_t>());
		MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	}
	MCB_TRY(mcb_index_build_device(ctx, cb.tup.as<mcb_tuple>(), T, cb.boff.as<uint64_t>(), cb.h_boff.as<uint64_t>(), &ix));
	McbSpan sp(ctx->tm, "combine");
	// ---- 3: (minimizer, posting) pairs in visiting order, match_pro, ordered partner lists without repeats
	uint64_t P = 0, P2 = 0;
	{
		if (M) MCB_LAUNCH(ctx, "cb_lookup", k_cb_lookup, mcb_grid_for(M, 256), 256, 0, cur.mins.as<mcb_tuple>(), M, ix, pcnt, pfirst);
		MCB_CUDA(cudaMemcpyAsync(poff, pcnt, M * 4, cudaMemcpyDeviceToDevice, ctx->stream));
		MCB_TRY(mcb_exclusive_scan_u32(ctx, poff, M, (uint64_t*)&dc[CT_CB_C]));
		MCB_TRY(cb_counters(ctx));
		C = CB_HC(ctx, CT_CB_C);
		if (C >= 0xFFFFFFFFull) { mcb_set_error("mcb_combine: too many (minimizer, posting) pairs"); return MCB_EINVAL; }
		MCB_TRY(cb.cand.ensure(C * 16 + 16)); MCB_TRY(cb.pass.ensure((C + 2) * 8));
		uint32_t *passed = cb.pass.as<uint32_t>(), *pscan = passed + (C + 1);
		if (M) MCB_LAUNCH(ctx, "cb_pairs", k_cb_pairs, mcb_grid_for(M, 256), 256, 0, cur.mins.as<mcb_tuple>(), M, ix, poff, pfirst, pcnt, cb.cand.as<CbCand>());
		if (C) MCB_LAUNCH(ctx, "cb_match", k_cb_match, mcb_grid_for(C, 256), 256, 0, cb.cand.as<CbCand>(), C, cb.cw.as<uint64_t>(), cb.cwo.as<uint64_t>(), cur.roff.as<uint64_t>(), cbthr, passed);
		MCB_CUDA(cudaMemcpyAsync(pscan, passed, C * 4, cudaMemcpyDeviceToDevice, ctx->stream));
		MCB_TRY(mcb_exclusive_scan_u32(ctx, pscan, C, (uint64_t*)&dc[CT_CB_D]));
		MCB_TRY(cb_counters(ctx));
		P = CB_HC(ctx, CT_CB_D);
		MCB_TRY(cb.plist.ensure(P * 16 + 16)); MCB_TRY(cb.loff.ensure((ncl + 2) * 4));
		if (C) MCB_LAUNCH(ctx, "cb_compact", k_cb_compact, mcb_grid_for(C, 256), 256, 0, cb.cand.as<CbCand>(), pscan, passed, C, cb.plist.as<CbCand>());
		// second compaction: the first occurrence of every partner only (cand is free again: it takes the short lists)
		uint32_t *keep = cb.pass.as<uint32_t>(), *kscan = keep + (C + 1);
		if (P) MCB_LAUNCH(ctx, "cb_dedupe", k_cb_dedupe, mcb_grid_for(P, 256), 256, 0, cb.plist.as<CbCand>(), P, keep);
		MCB_CUDA(cudaMemcpyAsync(kscan, keep, P * 4, cudaMemcpyDeviceToDevice, ctx->stream));
		MCB_TRY(mcb_exclusive_scan_u32(ctx, kscan, P, (uint64_t*)&dc[CT_CB_E]));
		if (P) MCB_LAUNCH(ctx, "cb_compact", k_cb_compact, mcb_grid_for(P, 256), 256, 0, cb.plist.as<CbCand>(), kscan, keep, P, cb.cand.as<CbCand>());
		MCB_TRY(cb_counters(ctx));
		P2 = CB_HC(ctx, CT_CB_E);
		MCB_LAUNCH(ctx, "cb_list_bounds", k_cb_list_bounds, mcb_grid_for(ncl + 1, 256), 256, 0, cb.cand.as<CbCand>(), P2, ncl, cb.loff.as<uint32_t>());
	}
	// ---- 4: first-come resolution in contig order (kthread_cb.c:460-466 with one thread), as bidding rounds on the device
	uint64_t nm = 0, n_copy = 0;
	{
		MCB_TRY(cb.alive.ensure(P2 + 16)); MCB_TRY(cb.best.ensure((ncl + 2) * 4)); MCB_TRY(cb.pick.ensure((ncl + 2) * 4)); MCB_TRY(cb.flag.ensure(ncl + 16));
		MCB_TRY(cb.pairs.ensure((ncl / 2 + 2) * 16)); MCB_TRY(cb.src.ensure((ncl + 2) * 4)); MCB_TRY(cb.tmp32.ensure((2 * ncl + 4) * 4));
		uint8_t *alive = cb.alive.as<uint8_t>(), *taken = cb.flag.as<uint8_t>();
		uint32_t *best = cb.best.as<uint32_t>(), *pick = cb.pick.as<uint32_t>();
		const CbCand *list = cb.cand.as<CbCand>();
		MCB_CUDA(cudaMemsetAsync(alive, 1, P2, ctx->stream)); MCB_CUDA(cudaMemsetAsync(taken, 0, ncl, ctx->stream));
		MCB_CUDA(cudaMemsetAsync(best, 0xFF, ncl * 4, ctx->stream)); MCB_CUDA(cudaMemsetAsync(pick, 0xFF, ncl * 4, ctx->stream));
		for (int round = 0; P2; ++round) {
			if (round > 100000) { mcb_set_error("mcb_combine: resolution does not converge"); return MCB_EINVAL; }
			const bool check = (round & 3) == 3;          // every fourth round the host looks at the number of live entries
			if (check) MCB_CUDA(cudaMemsetAsync(&dc[CT_CB_A], 0, 8, ctx->stream));
			MCB_LAUNCH(ctx, "cb_bid", k_cb_bid, mcb_grid_for(P2, 256), 256, 0, list, P2, taken, alive, best, &dc[check ? CT_CB_A : CT_CB_E]);
			MCB_LAUNCH(ctx, "cb_accept", k_cb_accept, mcb_grid_for(P2, 256), 256, 0, list, P2, taken, alive, best, pick);
			MCB_LAUNCH(ctx, "cb_round_reset", k_cb_round_reset, mcb_grid_for(P2, 256), 256, 0, list, P2, alive, best);
			if (check) {
				MCB_TRY(cb_counters(ctx));
				if (CB_HC(ctx, CT_CB_A) == 0) break;
			}
		}
		uint32_t *f_pick = cb.tmp32.as<uint32_t>(), *f_free = f_pick + (ncl + 1);
		MCB_LAUNCH(ctx, "cb_outcome_flags", k_cb_outcome_flags, mcb_grid_for(ncl, 256), 256, 0, pick, taken, ncl, f_pick, f_free);
		MCB_TRY(mcb_exclusive_scan_u32(ctx, f_pick, ncl, (uint64_t*)&dc[CT_CB_A]));
		MCB_TRY(mcb_exclusive_scan_u32(ctx, f_free, ncl, (uint64_t*)&dc[CT_CB_B]));
		MCB_LAUNCH(ctx, "cb_outcome", k_cb_outcome, mcb_grid_for(ncl, 256), 256, 0, list, pick, taken, ncl, f_pick, f_free, cb.pairs.as<CbCand>(), cb.src.as<uint32_t>());
		MCB_TRY(cb_counters(ctx));
		nm = CB_HC(ctx, CT_CB_A); n_copy = CB_HC(ctx, CT_CB_B);
	}
	*n_merged = nm;
	const uint64_t n2 = nm + n_copy;
	// ---- 5: the new set
	nxt.ncl = n2; nxt.nmem = cur.nmem;
	MCB_TRY(nxt.n.ensure((n2 + 2) * 4)); MCB_TRY(nxt.aoff.ensure((n2 + 2) * 8)); MCB_TRY(nxt.a.ensure(cur.nmem * 8 + 16)); MCB_TRY(nxt.roff.ensure((n2 + 2) * 8));
	MCB_TRY(cb.tmp32.ensure((n2 + 2) * 4)); MCB_TRY(cb.len2.ensure((n2 + 2) * 8));
	if (!n2) { nxt.nref = 0; nxt.nmin = 0; MCB_TRY(nxt.moff.ensure(16)); MCB_CUDA(cudaMemsetAsync(nxt.moff.p, 0, 8, ctx->stream)); return MCB_OK; }
	MCB_LAUNCH(ctx, "cb_new_sizes", k_cb_new_sizes, mcb_grid_for(n2, 256), 256, 0, cb.pairs.as<CbCand>(), nm, cb.src.as<uint32_t>(), n2, cur.n.as<uint32_t>(), nxt.n.as<uint32_t>());
	MCB_CUDA(cudaMemcpyAsync(cb.tmp32.p, nxt.n.p, n2 * 4, cudaMemcpyDeviceToDevice, ctx->stream));
	MCB_TRY(mcb_exclusive_scan_u32(ctx, cb.tmp32.as<uint32_t>(), n2, (uint64_t*)&dc[CT_CB_A]));
	MCB_LAUNCH(ctx, "cb_prefix64", k_cb_prefix64, mcb_grid_for(n2 + 1, 256), 256, 0, cb.tmp32.as<uint32_t>(), n2, nxt.aoff.as<uint64_t>(), &dc[CT_CB_A]);
	// members of the merged contigs go through a stable sort on (new contig, position, strand); the copies are final at once
	uint64_t n_mm = 0;                                           // members of merged contigs = aoff2[nm]
	if (nm) {
		MCB_CUDA(cudaMemcpyAsync(ctx->h_counters.p, nxt.aoff.as<uint64_t>() + nm, 8, cudaMemcpyDeviceToHost, ctx->stream));
		MCB_CUDA(cudaStreamSynchronize(ctx->stream));
		n_mm = ctx->h_counters.as<unsigned long long>()[0];
	}
	MCB_TRY(cb.keyA.ensure(n_mm * 16 + 16)); MCB_TRY(cb.keyB.ensure(n_mm * 16 + 16));
	MCB_LAUNCH(ctx, "cb_members", k_cb_members, mcb_grid_for(n2 * 32, 256), 256, 0, cb.pairs.as<CbCand>(), nm, cb.src.as<uint32_t>(), n2, cur.aoff.as<uint64_t>(), cur.a.as<uint64_t>(),
	           nxt.aoff.as<uint64_t>(), nxt.a.as<uint64_t>(), cb.keyA.as<ulonglong2>());
	if (n_mm) {
		std::vector<McbSortPass> passes;
		mcb_add_bit_passes(passes, 0, 0, 1 + mcb_bits_for(2 * *max_len));     // position << 1 | strand; a shifted position stays below the sum of two old lengths
		mcb_add_bit_passes(passes, 0, 32, 32 + mcb_bits_for(nm));
		ulonglong2 *sorted = nullptr;
		MCB_TRY(mcb_radix_sort(ctx, cb.keyA.as<ulonglong2>(), cb.keyB.as<ulonglong2>(), n_mm, passes.data(), (int)passes.size(), &sorted));
		MCB_LAUNCH(ctx, "cb_unkey", k_cb_unkey, mcb_grid_for(n_mm, 256), 256, 0, sorted, n_mm, nxt.a.as<uint64_t>());
	}
	// consensus strings
	MCB_CUDA(cudaMemsetAsync(&dc[CT_CB_C], 0, 8, ctx->stream));
	MCB_LAUNCH(ctx, "cb_new_lengths", k_cb_new_lengths, mcb_grid_for(n2, 256), 256, 0, nm, cb.src.as<uint32_t>(), n2, nxt.aoff.as<uint64_t>(), nxt.a.as<uint64_t>(), cur.roff.as<uint64_t>(), L,
	           cb.len2.as<uint64_t>(), &dc[CT_CB_C]);
	MCB_CUDA(cudaMemcpyAsync(nxt.roff.p, cb.len2.p, n2 * 8, cudaMemcpyDeviceToDevice, ctx->stream));
	MCB_TRY(mcb_exclusive_scan_u64(ctx, nxt.roff.as<uint64_t>(), n2, (uint64_t*)&dc[CT_CB_B]));
	MCB_CUDA(cudaMemcpyAsync(nxt.roff.as<uint64_t>() + n2, &dc[CT_CB_B], 8, cudaMemcpyDeviceToDevice, ctx->stream));
	if (nm) MCB_CUDA(cudaMemcpyAsync(&dc[CT_CB_D], nxt.roff.as<uint64_t>() + nm, 8, cudaMemcpyDeviceToDevice, ctx->stream));
	MCB_TRY(cb_counters(ctx));
	const uint64_t nref2 = CB_HC(ctx, CT_CB_B), merged_cols = nm ? CB_HC(ctx, CT_CB_D) : 0;
	*max_len = std::max<uint64_t>(1, CB_HC(ctx, CT_CB_C));
	nxt.nref = nref2;
	MCB_TRY(nxt.ref.ensure(nref2 + 32));
	MCB_CUDA(cudaMemsetAsync(nxt.ref.as<char>() + (nref2 & ~(uint64_t)7), 0, 24, ctx->stream));      // the sketch kernel reads whole 8-byte words
	MCB_TRY(cb_cols_reserve(ctx, cb, cb.cols_used + merged_cols));
	MCB_TRY(nxt.coff.ensure((n2 + 2) * 8));
	if (merged_cols) MCB_LAUNCH(ctx, "cb_consensus", k_cb_merge_counts, mcb_grid_for(nm * 32, 256), 256, 0, cb.pairs.as<CbCand>(), nm, nxt.roff.as<uint64_t>(), cur.roff.as<uint64_t>(),
	                            cur.coff.as<uint64_t>(), cb.cols.as<uint4>(), cb.cols_used, nxt.ref.as<char>());
	MCB_LAUNCH(ctx, "cb_new_coff", k_cb_new_coff, mcb_grid_for(n2, 256), 256, 0, nm, cb.src.as<uint32_t>(), n2, nxt.roff.as<uint64_t>(), cur.coff.as<uint64_t>(), cb.cols_used, nxt.coff.as<uint64_t>());
	cb.cols_used += merged_cols;
	if (n_copy) MCB_LAUNCH(ctx, "cb_copy_refs", k_cb_copy_refs, mcb_grid_for(n_copy * 32, 256), 256, 0, nm, cb.src.as<uint32_t>(), n2, cur.roff.as<uint64_t>(), cur.ref.as<char>(),
	                       nxt.roff.as<uint64_t>(), nxt.ref.as<char>());
	// minimizers of the new set: sketches of the merged contigs, lists of the others taken along
	MCB_TRY(cb_min_lists(ctx, cb, nxt, nm, &cur, cb.src.as<uint32_t>()));
	return MCB_OK;
}

extern "C" int mcb_combine(mcb_ctx *ctx, int cbthreshold, mcb_combine_result *res)
{
	if (!ctx || !res) { mcb_set_error("mcb_combine: null argument"); return MCB_EINVAL; }
	MCB_CUDA(cudaSetDevice(ctx->prm.device));
	if (!ctx->bucket_done) { mcb_set_error("mcb_combine: needs mcb_for_bucket first"); return MCB_ESTATE; }
	if (ctx->shard_n > 1) { mcb_set_error("mcb_combine: the contig merge runs on one GPU (gather the seed contigs first)"); return MCB_ESTATE; }
	if (cbthreshold < 0) { mcb_set_error("mcb_combine: bad cbthreshold"); return MCB_EINVAL; }
	memset(res, 0, sizeof *res);
	McbMergeScope merge_scope(ctx->tm, 1);                     // kernel timers of the merge are booked apart from the metric path's
	if (!ctx->cb) ctx->cb = new McbCombineState();
	McbCombineState &cb = *ctx->cb;
	MCB_TRY(ctx->h_counters.ensure(64 * 8));
	const McbBucketState &bs = ctx->bs;
	// ---- the seed contigs of kt_for_bucket are still on the device (ctx->d_out): they are set 0
	CbSet &s0 = cb.set[0];
	s0.ncl = bs.tot_cl; s0.nmem = bs.tot_mem; s0.nref = bs.tot_ref;
	MCB_TRY(s0.n.ensure((s0.ncl + 2) * 4)); MCB_TRY(s0.aoff.ensure((s0.ncl + 2) * 8)); MCB_TRY(s0.a.ensure(s0.nmem * 8 + 16)); MCB_TRY(s0.roff.ensure((s0.ncl + 2) * 8));
	MCB_TRY(s0.ref.ensure(s0.nref + 32));
	{
		McbSpan sp(ctx->tm, "combine");
		const cudaMemcpyKind DD = cudaMemcpyDeviceToDevice;
		MCB_CUDA(cudaMemsetAsync(s0.ref.as<char>() + (s0.nref & ~(uint64_t)7), 0, 24, ctx->stream));     // zero tail first: the sketch kernel reads whole 8-byte words
		MCB_CUDA(cudaMemsetAsync(s0.roff.p, 0, 8, ctx->stream)); MCB_CUDA(cudaMemsetAsync(s0.aoff.p, 0, 8, ctx->stream));
		if (s0.ncl) {
			MCB_CUDA(cudaMemcpyAsync(s0.n.p, ctx->d_out[0].p, s0.ncl * 4, DD, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(s0.aoff.p, ctx->d_out[1].p, (s0.ncl + 1) * 8, DD, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(s0.a.p, ctx->d_out[2].p, s0.nmem * 8, DD, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(s0.roff.p, ctx->d_out[3].p, (s0.ncl + 1) * 8, DD, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(s0.ref.p, ctx->d_out[4].p, s0.nref, DD, ctx->stream));
		}
		// per-column base counts of the seed contigs: the start of the arena, column g of the concatenated strings at entry g
		cb.cols_used = 0;
		MCB_TRY(cb_cols_reserve(ctx, cb, 2 * s0.nref));
		MCB_TRY(s0.coff.ensure((s0.ncl + 2) * 8));
		if (s0.ncl) MCB_CUDA(cudaMemcpyAsync(s0.coff.p, s0.roff.p, s0.ncl * 8, DD, ctx->stream));
		if (s0.nref) {
			// a seed contig spans fewer than 2L + 2 max_rounds columns (construct_ref: every member overlaps the first one's minimizer)
			const int maxcols = ((2 * ctx->L + 2 * ctx->prm.max_rounds + 8 + 31) / 32) * 32;
			const size_t smem = (size_t)CB_SEED_WARPS * 4 * maxcols * 4;
			if (smem > 48 * 1024) MCB_CUDA(cudaFuncSetAttribute(k_cb_seed_counts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
			unsigned long long *dc = ctx->d_counters.as<unsigned long long>();
			MCB_CUDA(cudaMemsetAsync(&dc[CT_CB_F], 0, 8, ctx->stream));
			MCB_LAUNCH(ctx, "cb_seed_counts", k_cb_seed_counts, mcb_grid_for(s0.ncl, CB_SEED_WARPS), CB_SEED_WARPS * 32, smem, s0.ncl, s0.roff.as<uint64_t>(), s0.aoff.as<uint64_t>(), s0.a.as<uint64_t>(),
			           ctx->d_packed.as<uint64_t>(), ctx->WS, ctx->L, maxcols, cb.cols.as<uint4>(), &dc[CT_CB_F]);
		}
		cb.cols_used = s0.nref;
		MCB_TRY(cb_min_lists(ctx, cb, s0, s0.ncl, nullptr, nullptr));          // every seed contig is sketched once; later only what a merge created
		if (s0.ncl && CB_HC(ctx, CT_CB_F)) { mcb_set_error("mcb_combine: internal: %llu seed contigs / members outside the expected column range", CB_HC(ctx, CT_CB_F)); return MCB_ECUDA; }
	}
	int cur = 0, iterations = 0;
	long pre_tot = 0;
	uint64_t merges_total = 0, tuples_total = 0, max_len = (uint64_t)4 * ctx->L + 4 * ctx->prm.max_rounds;      // seed contigs span fewer than 2L + 2 max_rounds columns
	for (;;) {
		uint64_t nm = 0, nt = 0;
		MCB_TRY(combine_iteration(ctx, cb, cb.set[cur], cb.set[cur ^ 1], cbthreshold, &nm, &nt, &max_len));
		cur ^= 1; ++iterations; merges_total += nm; tuples_total += nt;
		const long tot = (long)cb.set[cur].ncl;
		if (labs(pre_tot - tot) < 100) break;                      // kthread_cb.c:625
		pre_tot = tot;
	}
	// ---- the final contigs to the host
	CbSet &F = cb.set[cur];
	MCB_TRY(cb.h_cl_n.ensure(F.ncl * 4 + 16)); MCB_TRY(cb.h_cl_a_off.ensure((F.ncl + 1) * 8)); MCB_TRY(cb.h_cl_a.ensure(F.nmem * 8 + 16));
	MCB_TRY(cb.h_cl_ref_off.ensure((F.ncl + 1) * 8)); MCB_TRY(cb.h_cl_ref.ensure(F.nref + 16));
	{
		McbSpan sp(ctx->tm, "d2h");
		if (F.ncl) {
			MCB_CUDA(cudaMemcpyAsync(cb.h_cl_n.p, F.n.p, F.ncl * 4, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(cb.h_cl_a_off.p, F.aoff.p, (F.ncl + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(cb.h_cl_a.p, F.a.p, F.nmem * 8, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(cb.h_cl_ref_off.p, F.roff.p, (F.ncl + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(cb.h_cl_ref.p, F.ref.p, F.nref, cudaMemcpyDeviceToHost, ctx->stream));
		} else { cb.h_cl_a_off.as<uint64_t>()[0] = 0; cb.h_cl_ref_off.as<uint64_t>()[0] = 0; }
	}
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	// the merged contigs are what realign_hash runs against (preprocess.c:197-232): Stage 2 takes them from the device
	MCB_TRY(mcb_realign_prime_contigs(ctx, F.ref.as<char>(), cb.h_cl_ref_off.as<uint64_t>(), F.ncl));
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->tm.collect();
	res->n_clusters = F.ncl; res->cl_n = cb.h_cl_n.as<uint32_t>(); res->cl_a_off = cb.h_cl_a_off.as<uint64_t>(); res->cl_a = cb.h_cl_a.as<uint64_t>();
	res->cl_ref_off = cb.h_cl_ref_off.as<uint64_t>(); res->cl_ref = cb.h_cl_ref.as<char>();
	res->iterations = iterations; res->n_merges = merges_total; res->n_index_tuples = tuples_total;
	return MCB_OK;
}
