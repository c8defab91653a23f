// mcb_stage2.cu — realign_hash (kthread_hash_realign.c:569-594) on the device.
//
//   K5 k_s2_singles   singleRead2bitset (bbhashdict.c:127-227): gather the 2-bit reads of the singletons, their reverse
//                     complements, the near-poly-A / near-poly-T diversion, and the per-dictionary substring keys
//   K6 sort + table   constructdictionary_realign (kthread_hash_realign.c:3-140): one stable radix sort of (dict, key) ->
//                     CSR bins with ascending sg index; an open-addressing table replaces the BooPHF minimal perfect hash
//                     (its value only selects a bin; the `ull == ull1` re-check at :385 makes the lookup an exact match)
//   K7 k_s2_probe     realign_hash_search (:316-508): one thread per contig window; forward probes l=0..nd-1, reverse
//                     probes for dictionaries with dict_start>0; XOR/popcount verification, encode_byte gate (:283-314)
//   K8 claims         the single-threaded reference lets the FIRST probe step that matches a read claim it.  Here every
//                     matching step proposes priority P=(window, phase, l); atomicMin keeps the first; claims are then
//                     sorted by (P, sg index descending) = the reference's append order.
//
// Distance is the popcount of the XOR of 2-bit codes (bbhashdict.c:247-254), not a base count: A<->T and C<->G cost 2.
// The reference's code A=00,G=01,C=10,T=11 (kthread_hash_realign.c:251-258) and ours (A0 C1 G2 T3) differ only by
// swapping the two bits of a field, so popcounts, key equality and bin contents are identical.
#include "mcb_common.cuh"
#include <algorithm>

#define S2_MAXD 16
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

// ---------------------------------------------------------------- K5
__global__ void k_s2_singles(const uint32_t *__restrict__ sg, uint64_t S, const uint64_t *__restrict__ packed, S2Geom gm,
                             const uint32_t *__restrict__ nread_rid, const uint64_t *__restrict__ nread_mask, uint64_t n_nreads, uint64_t n_reads,
                             uint64_t *__restrict__ rd, uint64_t *__restrict__ rdrc, uint8_t *__restrict__ flagged, ulonglong2 *__restrict__ kv,
                             unsigned long long *__restrict__ counters)
{
	uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= S) return;
	const uint32_t rid = sg[s];
	if (rid >= n_reads) { atomicAdd(&counters[CT_S2_ERR], 1ull); return; }
	const int L = gm.L, Wd = gm.Wd, WS = gm.WS;
	uint64_t w[9], r[9];
#pragma unroll
	for (int i = 0; i < 9; ++i) { w[i] = i < Wd ? packed[(uint64_t)rid * WS + i] : 0ull; }
	// reverse complement: reverse fields over Wd words, then drop the pad fields that moved to the bottom
#pragma unroll
	for (int i = 0; i < 9; ++i) r[i] = 0;
	const int pad = Wd * 32 - L;
	for (int i = 0; i < Wd; ++i) {
		uint64_t a = mcb_rc_word(w[Wd - 1 - i]);
		uint64_t b = i + 1 < Wd ? mcb_rc_word(w[Wd - 2 - i]) : 0ull;
		r[i] = pad ? (a >> (2 * pad)) | (b << (64 - 2 * pad)) : a;
	}
	if (L & 31) r[Wd - 1] &= (1ull << (2 * (L & 31))) - 1;
	int pcA = 0, pcT = 0;
	for (int i = 0; i < Wd; ++i) {
		uint64_t valid = (i == Wd - 1 && (L & 31)) ? (1ull << (2 * (L & 31))) - 1 : ~0ull;
		pcA += __popcll(w[i]); pcT += __popcll(~w[i] & valid);
	}
	for (int i = 0; i < WS; ++i) { rd[s * WS + i] = w[i]; rdrc[s * WS + i] = r[i]; }
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
	for (int l = 0; l < gm.nd; ++l) {
		ulonglong2 e; e.x = ((unsigned long long)l << 34) | extract_bases(w, gm.dstart[l], gm.lt); e.y = s;
		kv[(uint64_t)l * S + s] = e;
	}
}

__global__ void k_s2_heads(const ulonglong2 *__restrict__ e, uint64_t n, uint32_t *__restrict__ flag)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) flag[i] = (i == 0 || e[i].x != e[i - 1].x) ? 1u : 0u;
}
// distinct keys -> bin start; also flatten the sorted values into the bin array
__global__ void k_s2_bins(const ulonglong2 *__restrict__ e, uint64_t n, const uint32_t *__restrict__ hscan, const unsigned long long *__restrict__ U,
                          uint64_t *__restrict__ ukey, uint32_t *__restrict__ bstart, uint32_t *__restrict__ bin_sg)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	ulonglong2 v = e[i];
	bin_sg[i] = (uint32_t)v.y;
	if (i == 0 || v.x != e[i - 1].x) { uint32_t u = hscan[i]; ukey[u] = v.x; bstart[u] = (uint32_t)i; }
	if (i == n - 1) bstart[*U] = (uint32_t)n;
}
#define S2_EMPTY 0xFFFFFFFFFFFFFFFFull
__global__ void k_s2_table_insert(const uint64_t *__restrict__ ukey, uint64_t U, unsigned long long *__restrict__ tkey, uint32_t *__restrict__ tval, uint64_t hmask)
{
	uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (u >= U) return;
	const unsigned long long key = ukey[u];
	uint64_t h = mix64(key) & hmask;
	for (;;) {
		unsigned long long old = atomicCAS(&tkey[h], S2_EMPTY, key);
		if (old == S2_EMPTY || old == key) { tval[h] = (uint32_t)u; return; }
		h = (h + 1) & hmask;
	}
}
__device__ __forceinline__ bool table_find(const unsigned long long *__restrict__ tkey, const uint32_t *__restrict__ tval, uint64_t hmask, unsigned long long key, uint32_t *u)
{
	uint64_t h = mix64(key) & hmask;
	for (;;) {
		unsigned long long k = tkey[h];
		if (k == key) { *u = tval[h]; return true; }
		if (k == S2_EMPTY) return false;
		h = (h + 1) & hmask;
	}
}

// ---------------------------------------------------------------- contig packing
// contig c occupies words [cw_off[c], cw_off[c+1]) (ceil(len/32)+1 words, zero padded)
__global__ void k_s2_pack_refs(const char *__restrict__ refs, const uint64_t *__restrict__ ref_off, const uint64_t *__restrict__ cw_off, uint64_t n_contigs,
                               uint64_t total_words, uint64_t *__restrict__ cw, unsigned long long *__restrict__ counters)
{
	uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= total_words) return;
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

// ---------------------------------------------------------------- K7
struct S2Probe {
	const uint64_t *cw, *cw_off, *woff;   // packed contigs, word offsets, window offsets (prefix sums)
	uint64_t n_contigs, n_windows;
	const unsigned long long *tkey; const uint32_t *tval; uint64_t hmask;
	const uint32_t *bstart, *bin_sg;
	const uint64_t *rd, *rdrc; const uint8_t *flagged;
	unsigned long long *claim;            // [S] min priority
	unsigned long long *counters;
};

__device__ __forceinline__ uint64_t rev_fields(uint64_t v, int nbases)
{
	// reverse the order of the low `nbases` 2-bit fields
	uint64_t r = __brevll(v) >> (64 - 2 * nbases);
	return ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
}

__global__ void __launch_bounds__(128) k_s2_probe(S2Probe p, S2Geom gm)
{
	const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long n_probe = 0, n_cand = 0;
	if (g < p.n_windows) {
	uint64_t lo = 0, hi = p.n_contigs;        // last c with woff[c] <= g (contigs without windows have equal offsets: take the last)
	while (hi - lo > 1) { uint64_t mid = (lo + hi) >> 1; if (p.woff[mid] <= g) lo = mid; else hi = mid; }
	const uint64_t c = lo;
	const uint64_t jj = g - p.woff[c];
	const int L = gm.L, Wd = gm.Wd, WS = gm.WS;
	// window bits
	uint64_t W[9];
	{
		const uint64_t *src = p.cw + p.cw_off[c] + (jj >> 5);
		const int sh = 2 * (int)(jj & 31);
#pragma unroll
		for (int i = 0; i < 9; ++i) {
			if (i < Wd) { uint64_t a = src[i]; W[i] = sh ? (a >> sh) | (src[i + 1] << (64 - sh)) : a; } else W[i] = 0;
		}
		if (L & 31) W[Wd - 1] &= (1ull << (2 * (L & 31))) - 1;
	}
	for (int phase = 0; phase < 2; ++phase) {
		for (int l = 0; l < gm.nd; ++l) {
			unsigned long long key;
			if (phase == 0) key = extract_bases(W, gm.dstart[l], gm.lt);
			else {
				if (gm.dstart[l] <= 0) continue;                                  // kthread_hash_realign.c:440 (j = 0)
				// key of the reverse-complemented window: complement and reverse the bases [L-dstart-lt, L-dstart)
				uint64_t v = extract_bases(W, L - gm.dstart[l] - gm.lt, gm.lt);
				key = rev_fields(~v & ((1ull << (2 * gm.lt)) - 1), gm.lt);
			}
			key |= (unsigned long long)l << 34;
			++n_probe;
			uint32_t u;
			if (!table_find(p.tkey, p.tval, p.hmask, key, &u)) continue;
			const uint32_t bs = p.bstart[u], be = p.bstart[u + 1];
			const unsigned long long prio = (g << 5) | ((unsigned long long)phase << 4) | (unsigned long long)l;
			for (uint32_t i = be; i-- > bs;) {
				const uint32_t s = p.bin_sg[i];
				const uint64_t *r = (phase ? p.rdrc : p.rd) + (uint64_t)s * WS;
				uint64_t X[9]; int pc = 0;
#pragma unroll
				for (int q = 0; q < 9; ++q) { X[q] = q < Wd ? (W[q] ^ r[q]) : 0ull; pc += __popcll(X[q]); }
				++n_cand;
				if (pc > gm.thr) continue;
				if ((phase == 0 || gm.thr > 24) && !enc_ok(X, L, gm.enc_limit)) continue;   // :393 / :461
				if (p.flagged[s]) continue;                                                // sg_flag already set by the poly-A/T diversion
				if (be - 1 - i >= (uint32_t)gm.maxsearch) atomicAdd(&p.counters[CT_S2_NEEDEXACT], 1ull);   // beyond the static scan window (:388)
				atomicMin(&p.claim[s], prio);
			}
		}
	}
	}
	// warp-aggregated statistics
	for (int o = 16; o; o >>= 1) { n_probe += __shfl_xor_sync(0xFFFFFFFFu, n_probe, o); n_cand += __shfl_xor_sync(0xFFFFFFFFu, n_cand, o); }
	if ((threadIdx.x & 31) == 0) { atomicAdd(&p.counters[CT_S2_PROBES], n_probe); atomicAdd(&p.counters[CT_S2_CAND], n_cand); }
}

// ---------------------------------------------------------------- K8
__global__ void k_s2_claim_flags(const unsigned long long *__restrict__ claim, const uint8_t *__restrict__ flagged, uint64_t S,
                                 uint32_t *__restrict__ f_claim, uint32_t *__restrict__ f_a, uint32_t *__restrict__ f_t)
{
	uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= S) return;
	f_claim[s] = claim[s] != S2_EMPTY; f_a[s] = flagged[s] == 1; f_t[s] = flagged[s] == 2;
}
__global__ void k_s2_claim_compact(const unsigned long long *__restrict__ claim, const uint8_t *__restrict__ flagged, uint64_t S,
                                   const uint32_t *__restrict__ p_claim, const uint32_t *__restrict__ p_a, const uint32_t *__restrict__ p_t,
                                   ulonglong2 *__restrict__ el, uint32_t *__restrict__ fpa, uint32_t *__restrict__ fpt)
{
	uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= S) return;
	unsigned long long c = claim[s];
	if (c != S2_EMPTY) { ulonglong2 e; e.x = c; e.y = 0xFFFFFFFFull - s; el[p_claim[s]] = e; }   // ascending y == descending sg index
	if (flagged[s] == 1) fpa[p_a[s]] = (uint32_t)s;
	if (flagged[s] == 2) fpt[p_t[s]] = (uint32_t)s;
}
__global__ void k_s2_claim_emit(const ulonglong2 *__restrict__ el, uint64_t n, const uint64_t *__restrict__ woff, uint64_t n_contigs, const uint32_t *__restrict__ sg,
                                uint32_t *__restrict__ out_c, uint32_t *__restrict__ out_s, uint64_t *__restrict__ out_y)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	ulonglong2 e = el[i];
	const uint64_t g = e.x >> 5; const unsigned dir = (unsigned)(e.x >> 4) & 1u;
	const uint32_t s = (uint32_t)(0xFFFFFFFFull - e.y);
	uint64_t lo = 0, hi = n_contigs;
	while (hi - lo > 1) { uint64_t mid = (lo + hi) >> 1; if (woff[mid] <= g) lo = mid; else hi = mid; }
	out_c[i] = (uint32_t)lo; out_s[i] = s;
	out_y[i] = ((uint64_t)sg[s] << 32) | ((g - woff[lo]) << 1) | dir;
}

// ================================================================= host
extern "C" int mcb_realign(mcb_ctx *ctx, const uint32_t *sg, uint64_t S, const char *refs, const uint64_t *ref_off, uint64_t n_contigs,
                           int threshold, int maxsearch, int ininumdict, mcb_realign_result *res)
{
	if (!ctx || !res) { mcb_set_error("mcb_realign: null argument"); return MCB_EINVAL; }
	MCB_CUDA(cudaSetDevice(ctx->prm.device));
	if (!ctx->reads_loaded) { mcb_set_error("mcb_realign: no reads loaded"); return MCB_ESTATE; }
	if ((S && !sg) || (n_contigs && (!refs || !ref_off))) { mcb_set_error("mcb_realign: null input"); return MCB_EINVAL; }
	if (S >= 0xFFFFFFFFull) { mcb_set_error("mcb_realign: too many singles"); return MCB_EINVAL; }
	memset(res, 0, sizeof(*res));
	const int L = ctx->L, Wd = ctx->Wd, WS = ctx->WS;
	// ---- geometry (setglobalarrays_realign, kthread_hash_realign.c:150-207)
	S2Geom gm; memset(&gm, 0, sizeof gm);
	gm.L = L; gm.Wd = Wd; gm.WS = WS; gm.thr = threshold; gm.maxsearch = maxsearch;
	gm.lt = L <= 80 ? 11 : 17;
	gm.nd = L / gm.lt;
	if (ininumdict > 1 && ininumdict < gm.nd) gm.nd = ininumdict;
	if (gm.nd > S2_MAXD) { mcb_set_error("numdict %d unsupported", gm.nd); return MCB_EINVAL; }
	int st0 = (ininumdict > 0 && ininumdict < gm.nd) ? L / 2 - (gm.lt * gm.nd) / 2 : 0;
	for (int i = 0; i < gm.nd; ++i) gm.dstart[i] = st0 + i * gm.lt;
	gm.enc_limit = (int)((double)L * 0.4);
	res->numdict = gm.nd;
	MCB_TRY(ctx->d_counters.ensure(64 * 8));
	unsigned long long *dc = ctx->d_counters.as<unsigned long long>();
	MCB_CUDA(cudaMemsetAsync(dc + 16, 0, 16 * 8, ctx->stream));
	// ---- host-side offsets
	uint64_t total_words = 0, n_windows = 0, ref_bytes = n_contigs ? ref_off[n_contigs] : 0;
	MCB_TRY(ctx->h_in0.ensure((n_contigs + 1) * 8)); MCB_TRY(ctx->h_in1.ensure((n_contigs + 1) * 8));
	uint64_t *cwo = ctx->h_in0.as<uint64_t>(), *wo = ctx->h_in1.as<uint64_t>();
	for (uint64_t c = 0; c < n_contigs; ++c) {
		uint64_t len = ref_off[c + 1] - ref_off[c];
		cwo[c] = total_words; wo[c] = n_windows;
		total_words += (len + 31) / 32 + 1;
		if (len >= (uint64_t)L) n_windows += len - L + 1;
	}
	cwo[n_contigs] = total_words + 1; wo[n_contigs] = n_windows;   // +1 guard word: the window loader reads one word ahead
	if (n_windows >= (1ull << 58)) { mcb_set_error("too many windows"); return MCB_EINVAL; }
	res->n_windows = n_windows;
	const uint64_t nkv = S * (uint64_t)gm.nd;
	if (S == 0 || n_windows == 0 || gm.nd == 0) return MCB_OK;
	if (nkv >= 0xFFFFFFFFull) { mcb_set_error("dictionary too large"); return MCB_EINVAL; }
	// d_scr roles: 0 sg, 1 refs ascii, 2 ref_off, 3 cw_off, 4 woff, 5 cw, 6 rd, 7 rdrc, 8 flagged, 9 kv A, 10 kv B, 11 misc
	DBuf &b_sg = ctx->d_scr[0], &b_refs = ctx->d_scr[1], &b_roff = ctx->d_scr[2], &b_cwo = ctx->d_scr[3], &b_wo = ctx->d_scr[4], &b_cw = ctx->d_scr[5];
	DBuf &b_rd = ctx->d_scr[6], &b_rc = ctx->d_scr[7], &b_fl = ctx->d_scr[8], &b_kva = ctx->d_scr[9], &b_kvb = ctx->d_scr[10], &b_misc = ctx->d_scr[11];
	MCB_TRY(b_sg.ensure(S * 4 + 16)); MCB_TRY(b_refs.ensure(ref_bytes + 16)); MCB_TRY(b_roff.ensure((n_contigs + 1) * 8));
	MCB_TRY(b_cwo.ensure((n_contigs + 1) * 8)); MCB_TRY(b_wo.ensure((n_contigs + 1) * 8)); MCB_TRY(b_cw.ensure((total_words + 2) * 8));
	MCB_TRY(b_rd.ensure(S * WS * 8 + 16)); MCB_TRY(b_rc.ensure(S * WS * 8 + 16)); MCB_TRY(b_fl.ensure(S + 16));
	MCB_TRY(b_kva.ensure(nkv * 16 + 16)); MCB_TRY(b_kvb.ensure(nkv * 16 + 16));
	{
		McbSpan sp(ctx->tm, "h2d");
		MCB_CUDA(cudaMemcpyAsync(b_sg.p, sg, S * 4, cudaMemcpyHostToDevice, ctx->stream));
		MCB_CUDA(cudaMemcpyAsync(b_refs.p, refs, ref_bytes, cudaMemcpyHostToDevice, ctx->stream));
		MCB_CUDA(cudaMemcpyAsync(b_roff.p, ref_off, (n_contigs + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
		MCB_CUDA(cudaMemcpyAsync(b_cwo.p, cwo, (n_contigs + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
		MCB_CUDA(cudaMemcpyAsync(b_wo.p, wo, (n_contigs + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
	}
	const int span_h = ctx->tm.begin("realign");
	MCB_CUDA(cudaMemsetAsync(b_cw.as<uint64_t>() + total_words, 0, 16, ctx->stream));
	MCB_LAUNCH(ctx, "s2_pack_refs", k_s2_pack_refs, mcb_grid_for(total_words, 256), 256, 0, b_refs.as<char>(), b_roff.as<uint64_t>(), b_cwo.as<uint64_t>(),
	           n_contigs, total_words, b_cw.as<uint64_t>(), dc);
	MCB_LAUNCH(ctx, "s2_singles", k_s2_singles, mcb_grid_for(S, 128), 128, 0, b_sg.as<uint32_t>(), S, ctx->d_packed.as<uint64_t>(), gm,
	           ctx->d_nread_rid.as<uint32_t>(), ctx->d_nread_mask.as<uint64_t>(), ctx->n_nreads, ctx->n_reads,
	           b_rd.as<uint64_t>(), b_rc.as<uint64_t>(), b_fl.as<uint8_t>(), b_kva.as<ulonglong2>(), dc);
	// ---- K6: dictionaries
	std::vector<McbSortPass> passes;
	mcb_add_bit_passes(passes, 0, 0, 2 * gm.lt);
	if (gm.nd > 1) mcb_add_bit_passes(passes, 0, 34, 34 + mcb_bits_for(gm.nd - 1));
	ulonglong2 *kvs = nullptr;
	MCB_TRY(mcb_radix_sort(ctx, b_kva.as<ulonglong2>(), b_kvb.as<ulonglong2>(), nkv, passes.data(), (int)passes.size(), &kvs));
	ulonglong2 *kvfree = kvs == b_kva.as<ulonglong2>() ? b_kvb.as<ulonglong2>() : b_kva.as<ulonglong2>();
	// misc layout: hscan u32[nkv] | ukey u64[nkv] | bstart u32[nkv+1] | bin_sg u32[nkv]
	size_t o_h = 0, o_uk = (nkv * 4 + 15) & ~(size_t)15, o_bs = o_uk + nkv * 8, o_bn = (o_bs + (nkv + 1) * 4 + 15) & ~(size_t)15, o_end = o_bn + nkv * 4;
	MCB_TRY(b_misc.ensure(o_end + 16));
	uint32_t *hscan = (uint32_t*)(b_misc.as<char>() + o_h); uint64_t *ukey = (uint64_t*)(b_misc.as<char>() + o_uk);
	uint32_t *bstart = (uint32_t*)(b_misc.as<char>() + o_bs), *bin_sg = (uint32_t*)(b_misc.as<char>() + o_bn);
	MCB_LAUNCH(ctx, "s2_heads", k_s2_heads, mcb_grid_for(nkv, 256), 256, 0, kvs, nkv, hscan);
	MCB_TRY(mcb_exclusive_scan_u32(ctx, hscan, nkv, (uint64_t*)&dc[CT_S2_U]));
	MCB_LAUNCH(ctx, "s2_bins", k_s2_bins, mcb_grid_for(nkv, 256), 256, 0, kvs, nkv, hscan, &dc[CT_S2_U], ukey, bstart, bin_sg);
	MCB_TRY(ctx->h_counters.ensure(64 * 8));
	MCB_CUDA(cudaMemcpyAsync(ctx->h_counters.p, dc, 64 * 8, cudaMemcpyDeviceToHost, ctx->stream));
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	unsigned long long *hc = ctx->h_counters.as<unsigned long long>();
	if (hc[CT_S2_ERR]) { mcb_set_error("mcb_realign: %llu invalid inputs (sg id out of range or non-ACGT contig character)", hc[CT_S2_ERR]); return MCB_EINPUT; }
	const uint64_t U = hc[CT_S2_U];
	res->n_dict_keys = U;
	uint64_t H = 1024; while (H < 2 * U) H <<= 1;
	// the free half of the kv double buffer holds the table (keys u64[H] | vals u32[H]); grow it if needed
	DBuf &b_tab = (kvfree == b_kva.as<ulonglong2>()) ? b_kva : b_kvb;
	MCB_TRY(b_tab.ensure(H * 12 + 16));
	unsigned long long *tkey = b_tab.as<unsigned long long>(); uint32_t *tval = (uint32_t*)(b_tab.as<char>() + H * 8);
	MCB_CUDA(cudaMemsetAsync(tkey, 0xFF, H * 8, ctx->stream));
	MCB_LAUNCH(ctx, "s2_table_insert", k_s2_table_insert, mcb_grid_for(U, 256), 256, 0, ukey, U, tkey, tval, H - 1);
	// ---- K7
	MCB_TRY(ctx->d_x[0].ensure(S * 8 + 16));                   // claim priorities
	unsigned long long *claim = ctx->d_x[0].as<unsigned long long>();
	MCB_CUDA(cudaMemsetAsync(claim, 0xFF, S * 8, ctx->stream));
	S2Probe pr; pr.cw = b_cw.as<uint64_t>(); pr.cw_off = b_cwo.as<uint64_t>(); pr.woff = b_wo.as<uint64_t>(); pr.n_contigs = n_contigs; pr.n_windows = n_windows;
	pr.tkey = tkey; pr.tval = tval; pr.hmask = H - 1; pr.bstart = bstart; pr.bin_sg = bin_sg; pr.rd = b_rd.as<uint64_t>(); pr.rdrc = b_rc.as<uint64_t>();
	pr.flagged = b_fl.as<uint8_t>(); pr.claim = claim; pr.counters = dc;
	MCB_LAUNCH(ctx, "s2_probe", k_s2_probe, mcb_grid_for(n_windows, 128), 128, 0, pr, gm);
	// ---- K8
	// reuse: refs ascii buffer is dead -> flags; sized S*12
	MCB_TRY(b_refs.ensure(S * 12 + 64));
	uint32_t *f_c = b_refs.as<uint32_t>(), *f_a = f_c + S, *f_t = f_a + S;
	MCB_LAUNCH(ctx, "s2_claim_flags", k_s2_claim_flags, mcb_grid_for(S, 256), 256, 0, claim, b_fl.as<uint8_t>(), S, f_c, f_a, f_t);
	MCB_TRY(mcb_exclusive_scan_u32(ctx, f_c, S, (uint64_t*)&dc[CT_S2_CLAIMS]));
	MCB_TRY(mcb_exclusive_scan_u32(ctx, f_a, S, (uint64_t*)&dc[CT_S2_FPA]));
	MCB_TRY(mcb_exclusive_scan_u32(ctx, f_t, S, (uint64_t*)&dc[CT_S2_FPT]));
	// claim elements go to the kv buffer that held the sorted pairs (dead now: bins were flattened)
	DBuf &b_el = (kvs == b_kva.as<ulonglong2>()) ? b_kva : b_kvb;
	MCB_TRY(ctx->d_x[1].ensure(S * 16 + 16));
	ulonglong2 *elA = b_el.as<ulonglong2>(), *elB = ctx->d_x[1].as<ulonglong2>();
	MCB_TRY(b_cw.ensure(S * 8 + 64));                           // packed contigs are dead after the probe: fpA | fpT lists
	uint32_t *d_fpa = b_cw.as<uint32_t>(), *d_fpt = d_fpa + S;
	MCB_LAUNCH(ctx, "s2_claim_compact", k_s2_claim_compact, mcb_grid_for(S, 256), 256, 0, claim, b_fl.as<uint8_t>(), S, f_c, f_a, f_t, elA, d_fpa, d_fpt);
	MCB_CUDA(cudaMemcpyAsync(ctx->h_counters.p, dc, 64 * 8, cudaMemcpyDeviceToHost, ctx->stream));
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	const uint64_t ncl = hc[CT_S2_CLAIMS], nfa = hc[CT_S2_FPA], nft = hc[CT_S2_FPT];
	if (hc[CT_S2_NEEDEXACT]) {
		mcb_set_error("mcb_realign: %llu matches lie beyond the maxsearch=%d scan window of their bin; the sequential bin-window emulation is not implemented",
		              hc[CT_S2_NEEDEXACT], maxsearch);
		return MCB_EINPUT;
	}
	passes.clear();
	mcb_add_bit_passes(passes, 1, 0, mcb_bits_for(S));                                // 0xFFFFFFFF-s: only the low bits vary
	mcb_add_bit_passes(passes, 0, 0, 5 + mcb_bits_for(n_windows));
	ulonglong2 *els = nullptr;
	MCB_TRY(mcb_radix_sort(ctx, elA, elB, ncl, passes.data(), (int)passes.size(), &els));
	MCB_TRY(b_rd.ensure(ncl * 16 + 64));                                                // rd is dead: claim outputs c | s | y
	uint64_t *o_y = b_rd.as<uint64_t>(); uint32_t *o_c = (uint32_t*)(o_y + ncl), *o_s = o_c + ncl;
	if (ncl) MCB_LAUNCH(ctx, "s2_claim_emit", k_s2_claim_emit, mcb_grid_for(ncl, 256), 256, 0, els, ncl, b_wo.as<uint64_t>(), n_contigs, b_sg.as<uint32_t>(), o_c, o_s, o_y);
	ctx->tm.end(span_h);
	MCB_TRY(ctx->h_claim_c.ensure(ncl * 4 + 16)); MCB_TRY(ctx->h_claim_s.ensure(ncl * 4 + 16)); MCB_TRY(ctx->h_claim_y.ensure(ncl * 8 + 16));
	MCB_TRY(ctx->h_fpA.ensure(nfa * 4 + 16)); MCB_TRY(ctx->h_fpT.ensure(nft * 4 + 16));
	{
		McbSpan sp(ctx->tm, "d2h");
		if (ncl) {
			MCB_CUDA(cudaMemcpyAsync(ctx->h_claim_c.p, o_c, ncl * 4, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(ctx->h_claim_s.p, o_s, ncl * 4, cudaMemcpyDeviceToHost, ctx->stream));
			MCB_CUDA(cudaMemcpyAsync(ctx->h_claim_y.p, o_y, ncl * 8, cudaMemcpyDeviceToHost, ctx->stream));
		}
		if (nfa) MCB_CUDA(cudaMemcpyAsync(ctx->h_fpA.p, d_fpa, nfa * 4, cudaMemcpyDeviceToHost, ctx->stream));
		if (nft) MCB_CUDA(cudaMemcpyAsync(ctx->h_fpT.p, d_fpt, nft * 4, cudaMemcpyDeviceToHost, ctx->stream));
		MCB_CUDA(cudaMemcpyAsync(ctx->h_counters.p, dc, 64 * 8, cudaMemcpyDeviceToHost, ctx->stream));
	}
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->tm.collect();
	res->n_claims = ncl; res->claim_contig = ctx->h_claim_c.as<uint32_t>(); res->claim_sg = ctx->h_claim_s.as<uint32_t>(); res->claim_y = ctx->h_claim_y.as<uint64_t>();
	res->n_fpA = nfa; res->n_fpT = nft; res->fpA_sg = ctx->h_fpA.as<uint32_t>(); res->fpT_sg = ctx->h_fpT.as<uint32_t>();
	res->n_probes = hc[CT_S2_PROBES]; res->n_candidates = hc[CT_S2_CAND];
	return MCB_OK;
}
