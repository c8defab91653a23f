/* mc_oracle.c — TEST INFRASTRUCTURE.  Single-threaded CPU restatement, in plain C, of the front end of
 * yuansliu/minicom (the path libminicom_b200.so replaces).  Only tests/, __graft_entry__.smoke() and the cpu_baseline
 * leg of bench.py may load this; the product never does.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors, so this restatement is pinned
 * against the reference itself: tests/test_oracle.py compares every function below with state dumps of the unmodified
 * reference built from /root/reference/src with num_thr=1 (oracle/ref/build_ref.sh -> oracle/_ref/), on the committed
 * fixtures under tests/golden/ and on larger seeded sets, plus the known-answer vectors of SURVEY.md §8c.
 *
 * Every function names the reference lines it restates (paths relative to /root/reference/src).  It follows the
 * num_thr=1 schedule, the only one in which the reference is deterministic.  readlen is a run-time value here (the
 * reference compiles it in, bbhashdict.h:57-63), so one library serves every read length.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static char *dup_str(const char *s) { size_t n = strlen(s) + 1; char *d = (char*)malloc(n); memcpy(d, s, n); return d; }

typedef struct { uint64_t x, y; } mco_tuple;

typedef struct {
	int32_t readlen, k, b, rw, first_mininum, diff_threshold, max_rounds;
} mco_params;

/* ------------------------------------------------------------------ growable arrays */
#define VEC(T) struct { T *a; size_t n, m; }
#define vpush(T, v, val) do { if ((v).n == (v).m) { (v).m = (v).m ? (v).m * 2 : 16; (v).a = (T*)realloc((v).a, (v).m * sizeof(T)); } (v).a[(v).n++] = (val); } while (0)
typedef VEC(uint8_t) vec_u8;
typedef VEC(uint32_t) vec_u32;
typedef VEC(uint64_t) vec_u64;
typedef VEC(mco_tuple) vec_tup;
typedef VEC(char) vec_chr;

/* ------------------------------------------------------------------ a2/a3: base codes and the k-mer hash */
/* seq_nt4_table (sketch.c:8-25): A/a 0, C/c 1, G/g 2, T/t 3, everything else 4 */
static int nt4(unsigned char c)
{
	switch (c) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': return 3;
	default: return 4;
	}
}
int mco_nt4(int c) { return nt4((unsigned char)c); }

/* hash64 (sketch.c:27-37): invertible integer mix, truncated to 2k bits after every addition */
uint64_t mco_hash64(uint64_t key, uint64_t mask)
{
	key = (~key + (key << 21)) & mask;
	key ^= key >> 24;
	key = (key + (key << 3) + (key << 8)) & mask;
	key ^= key >> 14;
	key = (key + (key << 2) + (key << 4)) & mask;
	key ^= key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}

/* ------------------------------------------------------------------ a4: mm_sketch_two (sketch.c:238-289) */
/* ONE minimizer per read: the smallest hash over all canonical k-mers, leftmost on ties.  No ambiguity handling: a
 * code-4 character is shifted in as is (:257-258).  k-mers equal to their reverse complement are skipped WITHOUT
 * advancing the valid-length counter (:265-272).  No valid k-mer -> {~0,~0}. */
void mco_sketch_two(const char *s, int len, int k, uint32_t rid, mco_tuple *out)
{
	const uint64_t mask = (1ULL << 2 * k) - 1, top = 2 * (k - 1);
	uint64_t fw = 0, rv = 0;
	mco_tuple best = { UINT64_MAX, UINT64_MAX };
	int have = 0;
	for (int i = 0; i < len; ++i) {
		uint64_t c = (uint64_t)nt4((unsigned char)s[i]);
		fw = (fw << 2 | c) & mask;
		rv = (rv >> 2) | (3ULL ^ c) << top;
		if (fw == rv) continue;
		int strand = fw < rv ? 0 : 1;
		if (++have < k) continue;
		uint64_t h = mco_hash64(strand ? rv : fw, mask);
		if (h < best.x) { best.x = h; best.y = (uint64_t)rid << 32 | (uint64_t)(uint32_t)i << 1 | (uint64_t)strand; }
	}
	*out = best;
}

/* ------------------------------------------------------------------ a11: mm_sketch_lh_ori (sketch.c:116-165) */
/* minimap-style (w,k) window minimizers over a contig string, robust-winnowing tie rules, output in position order.
 * Returns the number of tuples the reference would push; writes the first `cap` of them. */
int64_t mco_sketch_lh(const char *s, int len, int w, int k, uint32_t rid, mco_tuple *out, int64_t cap)
{
	if (len <= 0 || w <= 0 || k <= 0) return 0;
	const uint64_t mask = (1ULL << 2 * k) - 1, top = 2 * (k - 1);
	const mco_tuple none = { UINT64_MAX, UINT64_MAX };
	mco_tuple *ring = (mco_tuple*)malloc((size_t)w * sizeof(mco_tuple));
	for (int j = 0; j < w; ++j) ring[j] = none;
	uint64_t fw = 0, rv = 0;
	mco_tuple cur_min = none;
	int64_t n = 0;
	int have = 0, slot = 0, min_slot = 0;
#define EMIT(t) do { if (n < cap) out[n] = (t); ++n; } while (0)
	for (int i = 0; i < len; ++i) {
		int c = nt4((unsigned char)s[i]);
		mco_tuple here = none;
		if (c < 4) {
			fw = (fw << 2 | (uint64_t)c) & mask;
			rv = (rv >> 2) | (3ULL ^ (uint64_t)c) << top;
			if (fw == rv) continue;                              /* :136, skips the ring update too */
			int strand = fw < rv ? 0 : 1;
			if (++have >= k) { here.x = mco_hash64(strand ? rv : fw, mask); here.y = (uint64_t)rid << 32 | (uint64_t)(uint32_t)i << 1 | (uint64_t)strand; }
		} else have = 0;                                         /* :140 */
		ring[slot] = here;
		if (have == w + k - 1) {                                 /* first full window: duplicates of the minimum (:142-147) */
			for (int j = slot + 1; j < w; ++j) if (ring[j].x == cur_min.x && ring[j].y != cur_min.y) EMIT(ring[j]);
			for (int j = 0; j < slot; ++j) if (ring[j].x == cur_min.x && ring[j].y != cur_min.y) EMIT(ring[j]);
		}
		if (here.x <= cur_min.x) {                               /* new minimum: the old one is final (:148-150) */
			if (have >= w + k) EMIT(cur_min);
			cur_min = here; min_slot = slot;
		} else if (slot == min_slot) {                           /* minimum slid out of the window (:151-163) */
			if (have >= w + k - 1) EMIT(cur_min);
			cur_min.x = UINT64_MAX;
			for (int j = slot + 1; j < w; ++j) if (cur_min.x >= ring[j].x) { cur_min = ring[j]; min_slot = j; }
			for (int j = 0; j <= slot; ++j) if (cur_min.x >= ring[j].x) { cur_min = ring[j]; min_slot = j; }
			if (have >= w + k - 1) {
				for (int j = slot + 1; j < w; ++j) if (ring[j].x == cur_min.x && ring[j].y != cur_min.y) EMIT(ring[j]);
				for (int j = 0; j <= slot; ++j) if (ring[j].x == cur_min.x && ring[j].y != cur_min.y) EMIT(ring[j]);
			}
		}
		if (++slot == w) slot = 0;
	}
	if (cur_min.x != UINT64_MAX) EMIT(cur_min);                  /* :163-164 */
#undef EMIT
	free(ring);
	return n;
}

/* ------------------------------------------------------------------ a6: radix_sort_128x (misc.c:21-22, ksort.h:108-157) */
/* Sort by x only.  <=64 elements: insertion sort (stable).  Otherwise an in-place most-significant-digit sort whose
 * cycle-leader permutation is NOT stable; the exact order it leaves equal keys in is observable. */
static void insertion_by_x(mco_tuple *a, size_t n)
{
	for (size_t i = 1; i < n; ++i) {
		if (!(a[i].x < a[i - 1].x)) continue;
		mco_tuple t = a[i];
		size_t j = i;
		while (j > 0 && t.x < a[j - 1].x) { a[j] = a[j - 1]; --j; }
		a[j] = t;
	}
}
static void flag_sort_by_x(mco_tuple *a, size_t n, int shift)
{
	size_t cnt[256], head[256], tail[256], run = 0;
	memset(cnt, 0, sizeof cnt);
	for (size_t i = 0; i < n; ++i) ++cnt[(a[i].x >> shift) & 255];
	for (int d = 0; d < 256; ++d) { head[d] = run; run += cnt[d]; tail[d] = run; }
	for (int d = 0; d < 256;) {                                  /* ksort.h:131-145 */
		if (head[d] == tail[d]) { ++d; continue; }
		mco_tuple t = a[head[d]];
		int g = (int)((t.x >> shift) & 255);
		if (g == d) { ++head[d]; continue; }
		do {
			mco_tuple u = a[head[g]];
			a[head[g]++] = t;
			t = u;
			g = (int)((t.x >> shift) & 255);
		} while (g != d);
		a[head[d]++] = t;
	}
	if (shift == 0) return;
	int next = shift > 8 ? shift - 8 : 0;
	for (int d = 0; d < 256; ++d) {
		mco_tuple *seg = a + (tail[d] - cnt[d]);
		if (cnt[d] > 64) flag_sort_by_x(seg, cnt[d], next);
		else if (cnt[d] > 1) insertion_by_x(seg, cnt[d]);
	}
}
void mco_radix_sort_x(mco_tuple *a, int64_t n)
{
	if (n <= 64) insertion_by_x(a, (size_t)n);
	else flag_sort_by_x(a, (size_t)n, 56);
}

/* ------------------------------------------------------------------ a5: process_reads (kthread_reads.c:40-230) */
enum { MCO_SKETCHED = 0, MCO_ALLA, MCO_ALLT, MCO_ALLN, MCO_FPA, MCO_FPT, MCO_FPN, MCO_NFILE };

/* class of one read, tests in the reference's order (:84-226); *repl = replacement base for N (0 if none) */
int mco_classify(const char *s, int L, int e, char *repl)
{
	int a = 0, c = 0, g = 0, t = 0, n = 0;
	for (int i = 0; i < L; ++i) {
		switch (s[i]) { case 'A': ++a; break; case 'T': ++t; break; case 'G': ++g; break; case 'C': ++c; break; case 'N': ++n; break; default: break; }
	}
	*repl = 0;
	if (a == L) return MCO_ALLA;
	if (t == L) return MCO_ALLT;
	if (n == L) return MCO_ALLN;
	if (t + g + c + n <= e) return MCO_FPA;
	if (a + g + c + n <= e) return MCO_FPT;
	if (a + t + g + c <= e) return MCO_FPN;
	if (!(n <= 0.4 * L)) return MCO_NFILE;                       /* :182 (double comparison, as written) */
	if (n > 0) {                                                 /* most frequent base, ties A,T,G,C (:185-201) */
		int mx = a; if (t > mx) mx = t; if (g > mx) mx = g; if (c > mx) mx = c;
		*repl = mx == a ? 'A' : mx == t ? 'T' : mx == g ? 'G' : 'C';
	}
	return MCO_SKETCHED;
}

/* ------------------------------------------------------------------ stage 1 state */
typedef struct mco_stage1 {
	mco_params p;
	uint64_t n;
	char *rows;                 /* n*L, N-replaced in place like reads->seq[i].seq */
	char *orig;                 /* n*L, as given */
	vec_u8 cls;                 /* class per read */
	vec_tup tuples;             /* per read: tuple after kt_for_reads ({~0,~0} if not sketched) */
	/* kt_for_bucket results */
	vec_u32 cl_n; vec_u64 cl_a_off, cl_a, cl_ref_off; vec_chr cl_ref;
	vec_u32 sg;
	vec_tup mi;                 /* tuples pushed to mi[0], in push order */
	int rounds;
	uint64_t n_sketched_total;
} mco_stage1;

static mco_stage1 *g_sort_ctx;   /* cmpcluster reads two globals of the reference (reads->seq_len, reads->k) */

/* a8: cmpcluster (kthread_bucket.c:44-62): strand-adjusted position descending, then rid ascending */
static int adj_pos(uint64_t y, int L, int k) { int pos = (int)((uint32_t)y >> 1); return (y & 1) ? L - pos + k - 2 : pos; }
static int cmp_members(const void *pa, const void *pb)
{
	uint64_t a = *(const uint64_t*)pa, b = *(const uint64_t*)pb;
	int L = g_sort_ctx->p.readlen, k = g_sort_ctx->p.k;
	int qa = adj_pos(a, L, k), qb = adj_pos(b, L, k);
	if (qa == qb) return (int)(a >> 32) - (int)(b >> 32);
	return qb - qa;
}

/* reverse_complement (preprocess.c:22-37) of read rid into buf */
static void oriented(const mco_stage1 *S, uint32_t rid, int dir, char *buf)
{
	const int L = S->p.readlen;
	const char *r = S->rows + (size_t)rid * L;
	if (!dir) { memcpy(buf, r, L); return; }
	for (int i = 0; i < L; ++i) {
		char c = r[L - 1 - i];
		buf[i] = c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'G' ? 'C' : c == 'C' ? 'G' : c;
	}
}

static char majority(uint32_t *const cnt[4], int col, uint32_t *max_out)
{
	uint32_t mx = cnt[0][col]; char ch = 'A';
	for (int b = 1; b < 4; ++b) if (cnt[b][col] > mx) { mx = cnt[b][col]; ch = "ACGT"[b]; }   /* strict >: ties go to the lowest code (:111) */
	*max_out = mx;
	return ch;
}

typedef struct { vec_tup next; vec_u32 sg_tail; } round_out;

/* a9: construct_ref (kthread_bucket.c:69-377).  members: y values sorted by cmp_members, n >= 2.
 * Returns the number of kept members (written back to members[0..kept)) and the consensus in *ref_out (malloc'd).
 * Rejected members go to S->sg (last round) or are re-sketched with `kmer` into `next`, in member order. */
static int construct_ref(mco_stage1 *S, uint64_t *members, int n, int kmer, int last_round, vec_tup *next, char **ref_out)
{
	const int L = S->p.readlen, k = S->p.k, e = S->p.diff_threshold;
	const int tlen = L << 2;                                     /* `readlen<<1 + 1` parses as readlen<<2 (:72) */
	uint32_t *cnt[4];
	for (int b = 0; b < 4; ++b) cnt[b] = (uint32_t*)calloc((size_t)tlen, sizeof(uint32_t));
	char *buf = (char*)malloc((size_t)L + 1);
	char *cons = (char*)calloc((size_t)tlen + 1, 1);
	int pos0 = L;
	/* pass 1 (:78-121): pile-up of all members at offsets pos0 - pos', consensus up to the first empty column */
	for (int i = 0; i < n; ++i) {
		uint64_t y = members[i];
		uint32_t rid = (uint32_t)(y >> 32); int dir = (int)(y & 1);
		int pos = adj_pos(y, L, k);
		oriented(S, rid, dir, buf);
		if (i == 0) pos0 = pos;
		for (int s = 0; s < L; ++s) ++cnt[nt4((unsigned char)buf[s])][pos0 - pos + s];
		members[i] = (uint64_t)rid << 32 | (uint64_t)(pos0 - pos) << 1 | (uint64_t)dir;
	}
	for (int s = 0; s < tlen; ++s) {
		uint32_t mx; char ch = majority(cnt, s, &mx);
		if (mx == 0) break;
		cons[s] = ch;
	}
	const int ref_len = (int)strlen(cons);
	/* pass 2 (:165-216): character mismatches of every member against the consensus */
	int kept = 0;
	for (int i = 0; i < n; ++i) {
		uint64_t y = members[i];
		uint32_t rid = (uint32_t)(y >> 32); int off = (int)((uint32_t)y >> 1), dir = (int)(y & 1);
		oriented(S, rid, dir, buf);
		int dif = 0;
		for (int t = 0; t < L; ++t) if (cons[off + t] != buf[t]) ++dif;
		if (dif <= e) members[kept++] = y;
		else if (last_round) vpush(uint32_t, S->sg, rid);
		else {
			mco_tuple t;
			mco_sketch_two(S->rows + (size_t)rid * L, L, kmer, rid, &t);
			vpush(mco_tuple, *next, t);
			++S->n_sketched_total;
		}
	}
	/* pass 3 (:244-350): recount the kept members, trim leading empty columns, rebuild the consensus */
	char *ref = (char*)calloc((size_t)tlen + 1, 1);
	memcpy(ref, cons, (size_t)ref_len);
	if (kept > 0) {
		for (int b = 0; b < 4; ++b) memset(cnt[b], 0, (size_t)tlen * sizeof(uint32_t));
		int rend = 0;
		for (int i = 0; i < kept; ++i) {
			uint64_t y = members[i];
			uint32_t rid = (uint32_t)(y >> 32); int off = (int)((uint32_t)y >> 1), dir = (int)(y & 1);
			oriented(S, rid, dir, buf);
			for (int s = 0; s < L; ++s) ++cnt[nt4((unsigned char)buf[s])][off + s];
			if (off + L > rend) rend = off + L;
		}
		int s = 0;
		for (; s < ref_len; ++s) { uint32_t mx; majority(cnt, s, &mx); if (mx != 0) break; }
		const int sv = s;
		int r = 0;
		for (; s < rend; ++s, ++r) { uint32_t mx; ref[r] = majority(cnt, s, &mx); }
		ref[r] = 0;
		for (int i = 0; i < kept; ++i) {
			uint64_t y = members[i];
			members[i] = (y >> 32 << 32) | (uint64_t)((int)((uint32_t)y >> 1) - sv) << 1 | (y & 1);
		}
	}
	for (int b = 0; b < 4; ++b) free(cnt[b]);
	free(buf); free(cons);
	*ref_out = ref;
	return kept;
}

/* a7: process_bucket (kthread_bucket.c:381-509) for one bucket of the current round */
static void process_bucket(mco_stage1 *S, mco_tuple *a, size_t n, int kmer, int last_round, vec_tup *next)
{
	const int L = S->p.readlen;
	if (n == 0) return;
	mco_radix_sort_x(a, (int64_t)n);
	for (size_t g0 = 0; g0 < n;) {
		size_t g1 = g0 + 1;
		while (g1 < n && a[g1].x == a[g0].x) ++g1;
		const int cnt = (int)(g1 - g0);
		if (cnt < 2) vpush(uint32_t, S->sg, (uint32_t)(a[g0].y >> 32));      /* :402-413 */
		else {
			uint64_t *mem = (uint64_t*)malloc((size_t)cnt * 8);
			for (int i = 0; i < cnt; ++i) mem[i] = a[g0 + i].y;
			g_sort_ctx = S;
			qsort(mem, (size_t)cnt, 8, cmp_members);                           /* :442 */
			char *ref = NULL;
			int kept = construct_ref(S, mem, cnt, kmer, last_round, next, &ref);
			if (kept > 1) {                                                    /* a seed contig (:451-475) */
				const uint32_t cid = (uint32_t)S->cl_n.n << 8;                 /* ((clusters.n-1)<<8)+tid with tid=0 (:458) */
				vpush(uint32_t, S->cl_n, (uint32_t)kept);
				for (int i = 0; i < kept; ++i) vpush(uint64_t, S->cl_a, mem[i]);
				vpush(uint64_t, S->cl_a_off, (uint64_t)S->cl_a.n);
				const int len = (int)strlen(ref);
				for (int i = 0; i < len; ++i) vpush(char, S->cl_ref, ref[i]);
				vpush(uint64_t, S->cl_ref_off, (uint64_t)S->cl_ref.n);
				const int m = S->p.first_mininum;
				mco_tuple *tmp = (mco_tuple*)malloc((size_t)m * sizeof(mco_tuple));
				int64_t got = mco_sketch_lh(ref, len, S->p.rw, S->p.k, cid, tmp, m);
				for (int i = 0; i < m && i < got; ++i) vpush(mco_tuple, S->mi, tmp[i]);
				free(tmp);
			} else if (kept == 1) {                                            /* lone survivor (:476-498) */
				uint32_t rid = (uint32_t)(mem[0] >> 32);
				if (last_round) vpush(uint32_t, S->sg, rid);
				else {
					mco_tuple t;
					mco_sketch_two(S->rows + (size_t)rid * L, L, kmer, rid, &t);
					vpush(mco_tuple, *next, t);
					++S->n_sketched_total;
				}
			}
			free(ref); free(mem);
		}
		g0 = g1;
	}
}

static void bucket_partition(const mco_tuple *t, size_t n, int b, mco_tuple *out, size_t *off /* [2^b+1] */)
{
	const size_t nb = (size_t)1 << b, mask = nb - 1;
	memset(off, 0, (nb + 1) * sizeof(size_t));
	for (size_t i = 0; i < n; ++i) ++off[(t[i].x & mask) + 1];
	for (size_t i = 0; i < nb; ++i) off[i + 1] += off[i];
	size_t *cur = (size_t*)malloc(nb * sizeof(size_t));
	memcpy(cur, off, nb * sizeof(size_t));
	for (size_t i = 0; i < n; ++i) out[cur[t[i].x & mask]++] = t[i];
	free(cur);
}

/* kt_for_reads (kthread_reads.c:247) + kt_for_bucket (kthread_bucket.c:562-629).  rows: n*L characters, not modified. */
mco_stage1 *mco_stage1_run(const mco_params *p, const char *rows, uint64_t n)
{
	mco_stage1 *S = (mco_stage1*)calloc(1, sizeof *S);
	const int L = p->readlen, k = p->k;
	S->p = *p; S->n = n;
	S->rows = (char*)malloc((size_t)n * L + 1); memcpy(S->rows, rows, (size_t)n * L);
	S->orig = (char*)malloc((size_t)n * L + 1); memcpy(S->orig, rows, (size_t)n * L);
	vec_tup cur = {0, 0, 0};
	for (uint64_t rid = 0; rid < n; ++rid) {
		char *r = S->rows + (size_t)rid * L, repl;
		int c = mco_classify(r, L, p->diff_threshold, &repl);
		vpush(uint8_t, S->cls, (uint8_t)c);
		mco_tuple t = { UINT64_MAX, UINT64_MAX };
		if (c == MCO_SKETCHED) {
			if (repl) for (int i = 0; i < L; ++i) if (r[i] == 'N') r[i] = repl;
			mco_sketch_two(r, L, k, (uint32_t)rid, &t);
			vpush(mco_tuple, cur, t);
			++S->n_sketched_total;
		}
		vpush(mco_tuple, S->tuples, t);
	}
	vpush(uint64_t, S->cl_a_off, 0); vpush(uint64_t, S->cl_ref_off, 0);
	const size_t nb = (size_t)1 << p->b;
	size_t *off = (size_t*)malloc((nb + 1) * sizeof(size_t));
	int last_rounds = 0;
	long long prev_members = 0;
	for (int r = 1;; ++r) {
		if (k - r <= 9) ++last_rounds;                                   /* kthread_bucket.c:584 */
		if (r == p->max_rounds - 1) ++last_rounds;                       /* :585 */
		vec_tup next = {0, 0, 0};
		mco_tuple *part = (mco_tuple*)malloc((cur.n + 1) * sizeof(mco_tuple));
		bucket_partition(cur.a, cur.n, p->b, part, off);
		for (size_t bk = 0; bk < nb; ++bk) process_bucket(S, part + off[bk], off[bk + 1] - off[bk], k - r, last_rounds != 0, &next);
		free(part); free(cur.a);
		cur = next;
		S->rounds = r;
		if (last_rounds) ++last_rounds;                                  /* :594 */
		long long members = (long long)S->cl_a.n;
		if (members - prev_members < 100) ++last_rounds;                 /* :614-617 */
		prev_members = members;
		if (last_rounds > 1) break;                                      /* :622 */
	}
	free(cur.a); free(off);
	return S;
}

void mco_stage1_free(mco_stage1 *S)
{
	if (!S) return;
	free(S->rows); free(S->orig); free(S->cls.a); free(S->tuples.a); free(S->cl_n.a); free(S->cl_a_off.a); free(S->cl_a.a);
	free(S->cl_ref_off.a); free(S->cl_ref.a); free(S->sg.a); free(S->mi.a); free(S);
}

enum { MCO_F_CLS = 0, MCO_F_TUPLES, MCO_F_ROWS, MCO_F_CL_N, MCO_F_CL_A_OFF, MCO_F_CL_A, MCO_F_CL_REF_OFF, MCO_F_CL_REF, MCO_F_SG, MCO_F_MI };
const void *mco_stage1_field(const mco_stage1 *S, int what, uint64_t *count)
{
	switch (what) {
	case MCO_F_CLS: *count = S->cls.n; return S->cls.a;
	case MCO_F_TUPLES: *count = S->tuples.n; return S->tuples.a;
	case MCO_F_ROWS: *count = S->n * (uint64_t)S->p.readlen; return S->rows;
	case MCO_F_CL_N: *count = S->cl_n.n; return S->cl_n.a;
	case MCO_F_CL_A_OFF: *count = S->cl_a_off.n; return S->cl_a_off.a;
	case MCO_F_CL_A: *count = S->cl_a.n; return S->cl_a.a;
	case MCO_F_CL_REF_OFF: *count = S->cl_ref_off.n; return S->cl_ref_off.a;
	case MCO_F_CL_REF: *count = S->cl_ref.n; return S->cl_ref.a;
	case MCO_F_SG: *count = S->sg.n; return S->sg.a;
	case MCO_F_MI: *count = S->mi.n; return S->mi.a;
	default: *count = 0; return NULL;
	}
}
int mco_stage1_rounds(const mco_stage1 *S) { return S->rounds; }
uint64_t mco_stage1_sketched(const mco_stage1 *S) { return S->n_sketched_total; }

/* ------------------------------------------------------------------ a12/a13: the minimizer index (kthread_idx.c:84-173) */
/* worker_post sorts each bucket with radix_sort_128x and stores, per distinct minimizer, its y values in the order
 * the sort left them.  The khash of the reference is replaced by the sorted key array (its layout is not observable). */
typedef struct mco_index {
	int b;
	size_t nb;
	size_t *koff;               /* [nb+1] key range of each bucket */
	uint64_t *keys;             /* distinct minimizers, ascending inside a bucket */
	size_t *pstart;             /* [n_keys+1] */
	uint64_t *post;             /* y values */
	size_t n_keys, n_post;
} mco_index;

mco_index *mco_idx_build(const mco_tuple *tuples, const uint64_t *bucket_off, int b)
{
	mco_index *ix = (mco_index*)calloc(1, sizeof *ix);
	ix->b = b; ix->nb = (size_t)1 << b;
	const size_t n = (size_t)bucket_off[ix->nb];
	mco_tuple *t = (mco_tuple*)malloc((n + 1) * sizeof(mco_tuple));
	memcpy(t, tuples, n * sizeof(mco_tuple));
	ix->koff = (size_t*)calloc(ix->nb + 1, sizeof(size_t));
	ix->keys = (uint64_t*)malloc((n + 1) * 8); ix->pstart = (size_t*)malloc((n + 2) * sizeof(size_t)); ix->post = (uint64_t*)malloc((n + 1) * 8);
	size_t nk = 0;
	for (size_t bk = 0; bk < ix->nb; ++bk) {
		size_t s = (size_t)bucket_off[bk], e = (size_t)bucket_off[bk + 1];
		ix->koff[bk] = nk;
		mco_radix_sort_x(t + s, (int64_t)(e - s));                       /* kthread_idx.c:126 */
		for (size_t i = s; i < e; ++i) {
			if (i == s || t[i].x != t[i - 1].x) { ix->keys[nk] = t[i].x; ix->pstart[nk] = i; ++nk; }
			ix->post[i] = t[i].y;                                        /* :153-157 */
		}
	}
	ix->koff[ix->nb] = nk; ix->pstart[nk] = n; ix->n_keys = nk; ix->n_post = n;
	free(t);
	return ix;
}
/* mm_idx_get (kthread_idx.c:84-101) */
const uint64_t *mco_idx_get(const mco_index *ix, uint64_t x, int *n)
{
	*n = 0;
	size_t bk = (size_t)(x & (ix->nb - 1)), lo = ix->koff[bk], hi = ix->koff[bk + 1];
	while (lo < hi) {
		size_t mid = lo + (hi - lo) / 2;
		if (ix->keys[mid] < x) lo = mid + 1; else if (ix->keys[mid] > x) hi = mid;
		else { *n = (int)(ix->pstart[mid + 1] - ix->pstart[mid]); return ix->post + ix->pstart[mid]; }
	}
	return NULL;
}
void mco_idx_stats(const mco_index *ix, uint64_t *n_keys, uint64_t *n_post) { *n_keys = ix->n_keys; *n_post = ix->n_post; }
/* all keys in (bucket, key) order with their posting ranges, for whole-index comparisons */
const uint64_t *mco_idx_keys(const mco_index *ix) { return ix->keys; }
const uint64_t *mco_idx_postings(const mco_index *ix) { return ix->post; }
void mco_idx_key_starts(const mco_index *ix, uint64_t *out) { for (size_t i = 0; i <= ix->n_keys; ++i) out[i] = ix->pstart[i]; }
void mco_idx_free(mco_index *ix) { if (!ix) return; free(ix->koff); free(ix->keys); free(ix->pstart); free(ix->post); free(ix); }

/* ------------------------------------------------------------------ N1: combine_cluster (kthread_cb.c:570-630), the contig merge */
/* Restated for the GPU merge (mcb_combine); follows the num_thr=1 schedule: clusters in array order, flags set by the first
 * acceptable hit (kthread_cb.c:267-343), merged clusters first, untouched ones copied behind them (:474-494), until the
 * number of clusters changes by less than 100 (:625). */
typedef struct { vec_u64 a; char *ref; } mco_cl;
typedef struct mco_combine_out {
	vec_u32 cl_n; vec_u64 cl_a_off, cl_a, cl_ref_off; vec_chr cl_ref;
	int iterations;
	vec_u64 iter_tuples_off;      /* [iterations+1] into iter_tuples: the tuples handed to mm_idx_generation number j (push order) */
	vec_tup iter_tuples;
	vec_u32 iter_merges;          /* merges per iteration */
} mco_combine_out;

/* match_pro (kthread_cb.c:36-53): character mismatches over the overlap of two strings aligned at (i_, j_) */
static int match_pro(const char *s0, int n0, const char *s1, int n1, int i_, int j_)
{
	int tot = 0, match = 0;
	for (int i = i_, j = j_; i < n0 && j < n1; ++i, ++j) { ++tot; match += s0[i] == s1[j]; }
	for (int i = i_ - 1, j = j_ - 1; i >= 0 && j >= 0; --i, --j) { ++tot; match += s0[i] == s1[j]; }
	return tot - match;
}
/* cmpcluster2 (kthread_cb.c:55-70): position ascending, then direction */
static int cmp_pos_dir(const void *a_, const void *b_)
{
	uint64_t a = *(const uint64_t*)a_, b = *(const uint64_t*)b_;
	int pa = (int)((uint32_t)a >> 1), pb = (int)((uint32_t)b >> 1);
	if (pa == pb) return (int)(a & 1) - (int)(b & 1);
	return pa - pb;
}
/* construct_ref2 (kthread_cb.c:105-218): sort the members, pile the oriented reads up, majority per column ('A' unless a
 * base has strictly more votes, in the order A, C, G, T), length = end of the right-most member */
static char *construct_ref2(const mco_stage1 *S, vec_u64 *m)
{
	const int L = S->p.readlen;
	qsort(m->a, m->n, sizeof(uint64_t), cmp_pos_dir);              /* :107, the same libc sort as the reference */
	const int tot_len = (int)((uint32_t)m->a[m->n - 1] >> 1) + 2 * L;
	uint32_t *cnt = (uint32_t*)calloc((size_t)4 * (tot_len + 1), sizeof(uint32_t));
	char *buf = (char*)malloc(L + 1);
	int rend = 0;
	for (size_t q = 0; q < m->n; ++q) {
		const uint64_t y = m->a[q];
		const int pos = (int)((uint32_t)y >> 1);
		oriented(S, (uint32_t)(y >> 32), (int)(y & 1), buf);
		for (int t = 0; t < L; ++t) ++cnt[(size_t)nt4((unsigned char)buf[t]) * (tot_len + 1) + pos + t];
		if (pos + L > rend) rend = pos + L;
	}
	char *ref = (char*)calloc((size_t)rend + 1, 1);
	for (int c = 0; c < rend; ++c) {
		uint32_t mx = cnt[c]; char ch = 'A';
		for (int b = 1; b < 4; ++b) if (cnt[(size_t)b * (tot_len + 1) + c] > mx) { mx = cnt[(size_t)b * (tot_len + 1) + c]; ch = "ACGT"[b]; }
		ref[c] = ch;
	}
	free(cnt); free(buf);
	return ref;
}

static void push_first_m(const mco_stage1 *S, const char *ref, uint32_t cid, vec_tup *tuples, mco_tuple *scratch, int64_t cap)
{
	const int64_t n = mco_sketch_lh(ref, (int)strlen(ref), S->p.rw, S->p.k, cid << 8, scratch, cap);      /* win + win_step, win_step = 0 (:572) */
	for (int64_t j = 0; j < n && j < S->p.first_mininum && j < cap; ++j) vpush(mco_tuple, *tuples, scratch[j]);
}

mco_combine_out *mco_combine(const mco_stage1 *S, int cbthreshold)
{
	mco_combine_out *o = (mco_combine_out*)calloc(1, sizeof *o);
	const int b = S->p.b;
	const size_t nb = (size_t)1 << b;
	size_t n = S->cl_n.n;
	mco_cl *cur = (mco_cl*)calloc(n + 1, sizeof(mco_cl));
	for (size_t c = 0; c < n; ++c) {
		for (uint64_t q = S->cl_a_off.a[c]; q < S->cl_a_off.a[c + 1]; ++q) vpush(uint64_t, cur[c].a, S->cl_a.a[q]);
		const size_t len = (size_t)(S->cl_ref_off.a[c + 1] - S->cl_ref_off.a[c]);
		cur[c].ref = (char*)calloc(len + 1, 1);
		memcpy(cur[c].ref, S->cl_ref.a + S->cl_ref_off.a[c], len);
	}
	vec_tup tuples = {0, 0, 0};
	for (size_t i = 0; i < S->mi.n; ++i) vpush(mco_tuple, tuples, S->mi.a[i]);
	const int64_t cap = 1 << 20;
	mco_tuple *mini = (mco_tuple*)malloc((size_t)cap * sizeof(mco_tuple)), *scratch = (mco_tuple*)malloc((size_t)cap * sizeof(mco_tuple));
	long pre_tot = 0;
	vpush(uint64_t, o->iter_tuples_off, 0);
	for (;;) {
		/* mm_idx_generation over the tuples pushed for this cluster set (:574) */
		for (size_t i = 0; i < tuples.n; ++i) vpush(mco_tuple, o->iter_tuples, tuples.a[i]);
		vpush(uint64_t, o->iter_tuples_off, o->iter_tuples.n);
		mco_tuple *bm = (mco_tuple*)malloc((tuples.n + 1) * sizeof(mco_tuple));
		size_t *off = (size_t*)calloc(nb + 1, sizeof(size_t));
		bucket_partition(tuples.a, tuples.n, b, bm, off);
		uint64_t *off64 = (uint64_t*)malloc((nb + 1) * 8);
		for (size_t i = 0; i <= nb; ++i) off64[i] = off[i];
		mco_index *ix = mco_idx_build(bm, off64, b);
		free(bm); free(off); free(off64);
		tuples.n = 0;
		/* kt_find_next_for (:502-568) with one thread */
		uint8_t *flag = (uint8_t*)calloc(n + 1, 1);
		mco_cl *nxt = (mco_cl*)calloc(n + 1, sizeof(mco_cl));
		size_t nn = 0; uint32_t merges = 0;
		for (size_t i = 0; i < n; ++i) {
			if (flag[i]) continue;
			/* find_next (:220-372) */
			const char *ref = cur[i].ref;
			const int len = (int)strlen(ref);
			const uint32_t rid_ori = (uint32_t)i << 8;
			const int64_t nm = mco_sketch_lh(ref, len, S->p.rw, S->p.k, rid_ori, mini, cap);
			int done = 0;
			for (int64_t j = 0; j < nm && j < cap && !done; ++j) {
				int np = 0;
				const uint64_t *r = mco_idx_get(ix, mini[j].x, &np);
				const uint32_t pos_ori = (uint32_t)mini[j].y >> 1, dir_ori = (uint32_t)(mini[j].y & 1);
				for (int q = 0; q < np && !done; ++q) {
					const uint32_t rid = (uint32_t)(r[q] >> 32);
					if (rid == rid_ori) continue;
					const size_t cid = rid >> 8;
					const uint32_t pos = (uint32_t)r[q] >> 1, dir = (uint32_t)(r[q] & 1);
					if (dir != dir_ori || flag[i] || flag[cid]) continue;
					if (match_pro(ref, len, cur[cid].ref, (int)strlen(cur[cid].ref), (int)pos_ori, (int)pos) > cbthreshold) continue;
					mco_cl t; memset(&t, 0, sizeof t);
					const mco_cl *first = pos_ori >= pos ? &cur[i] : &cur[cid], *second = pos_ori >= pos ? &cur[cid] : &cur[i];
					const uint64_t shift = pos_ori >= pos ? pos_ori - pos : pos - pos_ori;
					for (size_t u = 0; u < first->a.n; ++u) vpush(uint64_t, t.a, first->a.a[u]);
					for (size_t u = 0; u < second->a.n; ++u) {
						const uint64_t y = second->a.a[u];
						vpush(uint64_t, t.a, y >> 32 << 32 | (((uint64_t)((uint32_t)y >> 1) + shift) << 1) | (y & 1));      /* :300-317 */
					}
					t.ref = construct_ref2(S, &t.a);
					flag[i] = flag[cid] = 1; done = 1; ++merges;
					nxt[nn] = t;
					push_first_m(S, t.ref, (uint32_t)nn, &tuples, scratch, cap);                                       /* :359-368 */
					++nn;
				}
			}
		}
		for (size_t i = 0; i < n; ++i) {           /* cp_cluster (:374-403) for everything left */
			if (flag[i]) continue;
			mco_cl t; memset(&t, 0, sizeof t);
			for (size_t u = 0; u < cur[i].a.n; ++u) vpush(uint64_t, t.a, cur[i].a.a[u]);
			t.ref = dup_str(cur[i].ref);
			nxt[nn] = t;
			push_first_m(S, t.ref, (uint32_t)nn, &tuples, scratch, cap);
			++nn;
		}
		vpush(uint32_t, o->iter_merges, merges);
		for (size_t c = 0; c < n; ++c) { free(cur[c].a.a); free(cur[c].ref); }
		free(cur); free(flag); mco_idx_free(ix);
		cur = nxt; n = nn;
		++o->iterations;
		if (labs(pre_tot - (long)n) < 100) break;          /* :625 */
		pre_tot = (long)n;
	}
	vpush(uint64_t, o->cl_a_off, 0); vpush(uint64_t, o->cl_ref_off, 0);
	for (size_t c = 0; c < n; ++c) {
		vpush(uint32_t, o->cl_n, (uint32_t)cur[c].a.n);
		for (size_t u = 0; u < cur[c].a.n; ++u) vpush(uint64_t, o->cl_a, cur[c].a.a[u]);
		vpush(uint64_t, o->cl_a_off, o->cl_a.n);
		for (const char *q = cur[c].ref; *q; ++q) vpush(char, o->cl_ref, *q);
		vpush(uint64_t, o->cl_ref_off, o->cl_ref.n);
		free(cur[c].a.a); free(cur[c].ref);
	}
	free(cur); free(tuples.a); free(mini); free(scratch);
	return o;
}
enum { MCO_CB_CL_N = 0, MCO_CB_CL_A_OFF, MCO_CB_CL_A, MCO_CB_CL_REF_OFF, MCO_CB_CL_REF, MCO_CB_ITER_OFF, MCO_CB_ITER_TUPLES, MCO_CB_ITER_MERGES };
const void *mco_combine_field(const mco_combine_out *o, int what, uint64_t *count)
{
	switch (what) {
	case MCO_CB_CL_N: *count = o->cl_n.n; return o->cl_n.a;
	case MCO_CB_CL_A_OFF: *count = o->cl_a_off.n; return o->cl_a_off.a;
	case MCO_CB_CL_A: *count = o->cl_a.n; return o->cl_a.a;
	case MCO_CB_CL_REF_OFF: *count = o->cl_ref_off.n; return o->cl_ref_off.a;
	case MCO_CB_CL_REF: *count = o->cl_ref.n; return o->cl_ref.a;
	case MCO_CB_ITER_OFF: *count = o->iter_tuples_off.n; return o->iter_tuples_off.a;
	case MCO_CB_ITER_TUPLES: *count = o->iter_tuples.n; return o->iter_tuples.a;
	case MCO_CB_ITER_MERGES: *count = o->iter_merges.n; return o->iter_merges.a;
	default: *count = 0; return NULL;
	}
}
int mco_combine_iterations(const mco_combine_out *o) { return o->iterations; }
void mco_combine_free(mco_combine_out *o)
{
	if (!o) return;
	free(o->cl_n.a); free(o->cl_a_off.a); free(o->cl_a.a); free(o->cl_ref_off.a); free(o->cl_ref.a); free(o->iter_tuples_off.a); free(o->iter_tuples.a); free(o->iter_merges.a);
	free(o);
}

/* ------------------------------------------------------------------ stage 2: realign_hash (kthread_hash_realign.c:569-594) */
/* 2-bit read images.  The reference's std::bitset<2L> puts base i at bits 2i,2i+1 with A=00 C=10 G=01 T=11 (bit 2i first;
 * :251-258).  Here a read is an array of 64-bit words with the same bit positions, so popcounts of XORs (basediff,
 * bbhashdict.c:247-254) and substring keys (`(b & mask1[l]) >> 2*dict_start[l]`, :28-29) are bit-for-bit the reference's. */
static void to_bits(const char *s, int L, uint64_t *w, int nw)
{
	memset(w, 0, (size_t)nw * 8);
	for (int i = 0; i < L; ++i) {
		uint64_t v = s[i] == 'C' ? 2 : s[i] == 'G' ? 1 : s[i] == 'T' ? 3 : 0;    /* bit 2i = G|T, bit 2i+1 = C|T */
		w[(2 * i) >> 6] |= v << ((2 * i) & 63);
	}
}
static uint64_t bits_range(const uint64_t *w, int lo_bit, int nbits)
{
	int wi = lo_bit >> 6, sh = lo_bit & 63;
	uint64_t v = w[wi] >> sh;
	if (sh && sh + nbits > 64) v |= w[wi + 1] << (64 - sh);
	return nbits >= 64 ? v : v & ((1ULL << nbits) - 1);
}
static int popdiff(const uint64_t *a, const uint64_t *b, int nw) { int c = 0; for (int i = 0; i < nw; ++i) c += __builtin_popcountll(a[i] ^ b[i]); return c; }

static int ndigits(int v) { int d = 1; while (v >= 10) { v /= 10; ++d; } return d; }

/* a19: encode_byte (kthread_hash_realign.c:283-314), including the stale-counter quirk of its literal branch (:301-305) */
int mco_encode_byte_ok(const char *read_oriented, const char *ref_window, int L)
{
	int len = 0, eq = 0;
	for (int t = 0; t < L; ++t) {
		if (ref_window[t] != read_oriented[t]) {
			if (eq > 1) { len += ndigits(eq); eq = 0; }
			else len += eq;                                            /* copies `eq` literal characters, eq is NOT reset */
			++len;
		} else ++eq;
	}
	if (len == 0) len = 1;
	return len <= L * 0.4;
}

/* run-length code length of singleRead2bitset's near-poly-A/T test (bbhashdict.c:157-176,190-209) */
static int polyrun_len(const char *s, int L, char base)
{
	int len = 0, eq = 0;
	for (int t = 0; t < L; ++t) {
		if (s[t] != base) { if (eq > 0) { len += ndigits(eq); eq = 0; } ++len; }
		else ++eq;
	}
	return len ? len : 1;
}

typedef struct {
	uint64_t n_claims;
	vec_u32 claim_contig, claim_sg; vec_u64 claim_y;
	vec_u32 fpA, fpT;           /* sg indices */
	uint8_t *flag;              /* sg_flag[n_sg] */
	uint64_t n_windows, n_probes, n_candidates, n_beyond_static_window;
	int numdict;
} mco_realign_out;

typedef struct { uint64_t key; uint32_t s; } keyed;
static int cmp_keyed(const void *a, const void *b)
{
	const keyed *p = (const keyed*)a, *q = (const keyed*)b;
	if (p->key != q->key) return p->key < q->key ? -1 : 1;
	return p->s < q->s ? -1 : p->s > q->s;
}

typedef struct {              /* one dictionary (bbhashdict.h:21-43) without the perfect hash */
	size_t n_keys;
	uint64_t *keys;           /* sorted distinct keys */
	uint32_t *start;          /* [n_keys+1] */
	uint32_t *ids;            /* bins, ascending sg index (:105-127); removal shifts the tail left (bbhashdict.c:49-67) */
	uint32_t *live;           /* live entries per bin (the reference keeps this in the bin's last slots, bbhashdict.c:33-43) */
	uint8_t *empty;           /* empty_bin: the last read of a bin is never physically removed (:51-55) */
} dict_t;

static long dict_find(const dict_t *d, uint64_t key)
{
	size_t lo = 0, hi = d->n_keys;
	while (lo < hi) { size_t mid = lo + (hi - lo) / 2; if (d->keys[mid] < key) lo = mid + 1; else hi = mid; }
	return (lo < d->n_keys && d->keys[lo] == key) ? (long)lo : -1;
}
/* bbhashdict::remove (bbhashdict.c:45-67) */
static void dict_remove(dict_t *d, long bin, uint32_t s)
{
	uint32_t *ids = d->ids + d->start[bin];
	uint32_t live = d->live[bin];
	if (live == 1) { d->empty[bin] = 1; return; }
	uint32_t lo = 0, hi = live;
	while (lo < hi) { uint32_t mid = (lo + hi) / 2; if (ids[mid] < s) lo = mid + 1; else hi = mid; }
	memmove(ids + lo, ids + lo + 1, (size_t)(live - lo - 1) * 4);
	d->live[bin] = live - 1;
}

/* realign_hash for one threshold: singleRead2bitset (bbhashdict.c:127-227), constructdictionary_realign
 * (kthread_hash_realign.c:3-140), realign_hash_search over every contig in order (:316-508), num_thr=1 schedule.
 * sg: read ids of the singles; refs/ref_off: contig consensus strings. */
mco_realign_out *mco_realign(const mco_stage1 *S, const uint32_t *sg, uint64_t n_sg, const char *refs, const uint64_t *ref_off, uint64_t n_contigs,
                             int threshold, int maxsearch, int ininumdict)
{
	const int L = S->p.readlen, nw = (2 * L + 63) / 64 + 1;
	mco_realign_out *o = (mco_realign_out*)calloc(1, sizeof *o);
	o->flag = (uint8_t*)calloc(n_sg + 1, 1);
	/* setglobalarrays_realign (:150-207) */
	const int lt = L <= 80 ? 11 : 17;
	int nd = L / lt;
	if (ininumdict > 1 && ininumdict < nd) nd = ininumdict;
	int *dstart = (int*)malloc((size_t)(nd + 1) * sizeof(int));
	dstart[0] = (ininumdict > 0 && ininumdict < nd) ? L / 2 - (lt * nd) / 2 : 0;      /* never true after the line above */
	for (int i = 1; i < nd; ++i) dstart[i] = dstart[i - 1] + lt;
	o->numdict = nd;
	/* singleRead2bitset */
	uint64_t *bits = (uint64_t*)calloc((size_t)(n_sg + 1) * nw, 8);
	uint64_t *allA = (uint64_t*)calloc((size_t)nw, 8), *allT = (uint64_t*)calloc((size_t)nw, 8);
	char *tmp = (char*)malloc((size_t)L + 1), *tmp2 = (char*)malloc((size_t)L + 1);
	memset(tmp, 'A', L); to_bits(tmp, L, allA, nw);
	memset(tmp, 'T', L); to_bits(tmp, L, allT, nw);
	for (uint64_t s = 0; s < n_sg; ++s) {
		const char *r = S->rows + (size_t)sg[s] * L, *orig = S->orig + (size_t)sg[s] * L;
		to_bits(r, L, bits + s * nw, nw);
		for (int i = 0; i < L; ++i) tmp[i] = orig[i] == 'N' ? 'N' : r[i];               /* N restored (bbhashdict.c:148-154) */
		if (popdiff(bits + s * nw, allA, nw) <= threshold) {
			if (polyrun_len(tmp, L, 'A') <= L * 0.4) { o->flag[s] = 1; vpush(uint32_t, o->fpA, (uint32_t)s); }
		} else if (popdiff(bits + s * nw, allT, nw) <= threshold) {
			if (polyrun_len(tmp, L, 'T') <= L * 0.4) { o->flag[s] = 1; vpush(uint32_t, o->fpT, (uint32_t)s); }
		}
	}
	/* dictionaries */
	dict_t *D = (dict_t*)calloc((size_t)nd + 1, sizeof(dict_t));
	keyed *kv = (keyed*)malloc((n_sg + 1) * sizeof(keyed));
	for (int l = 0; l < nd; ++l) {
		dict_t *d = &D[l];
		for (uint64_t s = 0; s < n_sg; ++s) { kv[s].key = bits_range(bits + s * nw, 2 * dstart[l], 2 * lt); kv[s].s = (uint32_t)s; }
		qsort(kv, n_sg, sizeof(keyed), cmp_keyed);
		d->keys = (uint64_t*)malloc((n_sg + 1) * 8); d->start = (uint32_t*)malloc((n_sg + 2) * 4); d->ids = (uint32_t*)malloc((n_sg + 1) * 4);
		d->live = (uint32_t*)malloc((n_sg + 1) * 4); d->empty = (uint8_t*)calloc(n_sg + 1, 1);
		size_t nk = 0;
		for (uint64_t i = 0; i < n_sg; ++i) {
			if (i == 0 || kv[i].key != kv[i - 1].key) { d->keys[nk] = kv[i].key; d->start[nk] = (uint32_t)i; ++nk; }
			d->ids[i] = kv[i].s;
		}
		d->start[nk] = (uint32_t)n_sg; d->n_keys = nk;
		for (size_t q = 0; q < nk; ++q) d->live[q] = d->start[q + 1] - d->start[q];
	}
	free(kv);
	/* realign_hash_search, contigs in order */
	uint64_t *win = (uint64_t*)calloc((size_t)nw, 8), *rcw = (uint64_t*)calloc((size_t)nw, 8);
	for (uint64_t c = 0; c < n_contigs; ++c) {
		const char *ref = refs + ref_off[c];
		const long reflen = (long)(ref_off[c + 1] - ref_off[c]);
		for (long jj = 0; jj + L <= reflen; ++jj) {
			++o->n_windows;
			to_bits(ref + jj, L, win, nw);
			for (int i = 0; i < L; ++i) { char ch = ref[jj + L - 1 - i]; tmp[i] = ch == 'A' ? 'T' : ch == 'C' ? 'G' : ch == 'G' ? 'C' : 'A'; }   /* reverse_complement_ */
			to_bits(tmp, L, rcw, nw);
			for (int phase = 0; phase < 2; ++phase) {
				const uint64_t *probe = phase ? rcw : win;
				for (int l = 0; l < nd; ++l) {
					if (phase == 0 ? (dstart[l] + lt - 1 >= L) : (dstart[l] <= 0)) continue;       /* :363 / :440 with j = 0 */
					++o->n_probes;
					dict_t *d = &D[l];
					long bin = dict_find(d, bits_range(probe, 2 * dstart[l], 2 * lt));            /* MPHF lookup + key re-check (:368-386) */
					if (bin < 0 || d->empty[bin]) continue;
					const uint32_t *ids = d->ids + d->start[bin];
					const long live = (long)d->live[bin];
					vec_u32 removed = {0, 0, 0};
					for (long i = live - 1; i >= 0 && i >= live - maxsearch; --i) {                 /* :388 / :458 */
						const uint32_t s = ids[i];
						++o->n_candidates;
						if (popdiff(probe, bits + (size_t)s * nw, nw) > threshold) continue;
						if (phase == 0 || threshold > 24) {                                          /* :393 / :461 */
							const char *r = S->rows + (size_t)sg[s] * L;
							if (phase) { for (int q = 0; q < L; ++q) { char ch = r[L - 1 - q]; tmp2[q] = ch == 'A' ? 'T' : ch == 'T' ? 'A' : ch == 'G' ? 'C' : ch == 'C' ? 'G' : ch; } }
							if (!mco_encode_byte_ok(phase ? tmp2 : r, ref + jj, L)) continue;
						}
						if (o->flag[s]) continue;
						o->flag[s] = 1;
						vpush(uint32_t, o->claim_contig, (uint32_t)c); vpush(uint32_t, o->claim_sg, s);
						vpush(uint64_t, o->claim_y, (uint64_t)sg[s] << 32 | (uint64_t)jj << 1 | (uint64_t)phase);
						vpush(uint32_t, removed, s);
					}
					for (size_t q = 0; q < removed.n; ++q)                                           /* :409-424: out of every dictionary */
						for (int l1 = 0; l1 < nd; ++l1) {
							long b1 = dict_find(&D[l1], bits_range(bits + (size_t)removed.a[q] * nw, 2 * dstart[l1], 2 * lt));
							dict_remove(&D[l1], b1, removed.a[q]);
						}
					free(removed.a);
				}
			}
		}
	}
	o->n_claims = o->claim_y.n;
	for (int l = 0; l < nd; ++l) { free(D[l].keys); free(D[l].start); free(D[l].ids); free(D[l].live); free(D[l].empty); }
	free(D); free(bits); free(allA); free(allT); free(tmp); free(tmp2); free(win); free(rcw); free(dstart);
	return o;
}

enum { MCO_R_CONTIG = 0, MCO_R_SG, MCO_R_Y, MCO_R_FPA, MCO_R_FPT, MCO_R_FLAG };
const void *mco_realign_field(const mco_realign_out *o, int what, uint64_t *count, uint64_t n_sg)
{
	switch (what) {
	case MCO_R_CONTIG: *count = o->claim_contig.n; return o->claim_contig.a;
	case MCO_R_SG: *count = o->claim_sg.n; return o->claim_sg.a;
	case MCO_R_Y: *count = o->claim_y.n; return o->claim_y.a;
	case MCO_R_FPA: *count = o->fpA.n; return o->fpA.a;
	case MCO_R_FPT: *count = o->fpT.n; return o->fpT.a;
	case MCO_R_FLAG: *count = n_sg; return o->flag;
	default: *count = 0; return NULL;
	}
}
void mco_realign_counters(const mco_realign_out *o, uint64_t *out4) { out4[0] = o->n_windows; out4[1] = o->n_probes; out4[2] = o->n_candidates; out4[3] = (uint64_t)o->numdict; }
void mco_realign_free(mco_realign_out *o)
{
	if (!o) return;
	free(o->claim_contig.a); free(o->claim_sg.a); free(o->claim_y.a); free(o->fpA.a); free(o->fpT.a); free(o->flag); free(o);
}

/* ------------------------------------------------------------------ N3: FASTQ records (bseq.c:38-66, kseq.h:185-224) */
/* kseq's buffered stream over a file, here over memory: ks_getc (kseq.h:78-91) and ks_getuntil2 (kseq.h:94-141) */
typedef struct { const unsigned char *p; size_t n, at; } mco_stream;
static int st_getc(mco_stream *s) { return s->at < s->n ? s->p[s->at++] : -1; }
static int st_isspace(int c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
/* delimiter 0 = white space, 2 = newline (KS_SEP_SPACE / KS_SEP_LINE); returns the string length or -1 at end of input */
static long st_getuntil(mco_stream *s, int delimiter, vec_chr *str, int *dret, int append)
{
	if (dret) *dret = 0;
	if (!append) str->n = 0;
	if (s->at >= s->n) return -1;
	while (s->at < s->n) {
		int c = s->p[s->at++];
		if (delimiter == 2 ? c == '\n' : st_isspace(c)) { if (dret) *dret = c; break; }
		vpush(char, *str, (char)c);
	}
	if (delimiter == 2 && str->n > 1 && str->a[str->n - 1] == '\r') --str->n;     /* kseq.h:138 */
	return (long)str->n;
}

/* kseq_read + the loop of bseq_read: every sequence of the input, concatenated into seqs (no terminators), its length in
 * lens[i]; stops like the reference at end of input or at the first record whose quality string is missing or of another
 * length (kseq_read < 0, bseq.c:44).  Returns the number of sequences. */
int64_t mco_kseq_all(const char *buf, uint64_t len, char *seqs, uint64_t seqs_cap, uint32_t *lens, int64_t lens_cap)
{
	mco_stream s = { (const unsigned char*)buf, (size_t)len, 0 };
	vec_chr name = {0, 0, 0}, comment = {0, 0, 0}, seq = {0, 0, 0}, qual = {0, 0, 0};
	int last_char = 0, c;
	int64_t n = 0;
	uint64_t used = 0;
	for (;;) {
		if (last_char == 0) {                                    /* kseq.h:189-193 */
			while ((c = st_getc(&s)) != -1 && c != '>' && c != '@');
			if (c == -1) break;
			last_char = c;
		}
		seq.n = qual.n = comment.n = 0;
		if (st_getuntil(&s, 0, &name, &c, 0) < 0) break;          /* :195 */
		if (c != '\n') st_getuntil(&s, 2, &comment, 0, 0);        /* :196 */
		while ((c = st_getc(&s)) != -1 && c != '>' && c != '+' && c != '@') {   /* :201-205 */
			if (c == '\n') continue;
			vpush(char, seq, (char)c);
			st_getuntil(&s, 2, &seq, 0, 1);
		}
		if (c == '>' || c == '@') last_char = c;                  /* :206 */
		if (c == '+') {                                           /* :213-222 */
			while ((c = st_getc(&s)) != -1 && c != '\n');
			if (c == -1) break;                                   /* -2: no quality string */
			while (st_getuntil(&s, 2, &qual, 0, 1) >= 0 && qual.n < seq.n);
			last_char = 0;
			if (seq.n != qual.n) break;                           /* -2 */
		}
		if (n < lens_cap) lens[n] = (uint32_t)seq.n;
		if (used + seq.n <= seqs_cap) memcpy(seqs + used, seq.a, seq.n);
		used += seq.n; ++n;
	}
	free(name.a); free(comment.a); free(seq.a); free(qual.a);
	return n;
}

/* one read in the device layout: base j -> word j/32, bits 2*(j%32), A0 C1 G2 T3 (seq_nt4_table, upper case
 * only: process_reads counts nothing else, kthread_reads.c:56-73), N as code 0 with its position in mask (seq->n_pos, :69-80).
 * Returns 0, 1 (has N) or -1 (a character outside ACGTN). */
int mco_pack_row(const char *s, int L, int WS, uint64_t *row, uint64_t *mask)
{
	int hasn = 0;
	for (int w = 0; w < WS; ++w) row[w] = mask[w] = 0;
	for (int j = 0; j < L; ++j) {
		uint64_t code;
		switch (s[j]) {
		case 'A': code = 0; break;
		case 'C': code = 1; break;
		case 'G': code = 2; break;
		case 'T': code = 3; break;
		case 'N': code = 0; hasn = 1; mask[j / 32] |= 1ull << (2 * (j % 32)); break;
		default: return -1;
		}
		row[j / 32] |= code << (2 * (j % 32));
	}
	return hasn;
}

/* ------------------------------------------------------------------ N2: the per-read diff encoding of the dump stage */
/* print_encode's inner loop (kthread_dump.c:66-118, identical in the ORDER, default and _PE variants): the read as it was before
 * N replacement (N put back, :70-75), reverse-complemented when dir (preprocess.c:22-36: N stays N), compared with the consensus
 * window; a run of >= 2 equal characters becomes its decimal length, a run of 1 is copied, every mismatching character is copied;
 * the trailing run is dropped; no mismatch at all -> "0".  read: L characters (may hold N); ref_window: L consensus characters;
 * out: at least L+1 bytes.  Returns the length (no terminator). */
int mco_print_encode(const char *read, int dir, const char *ref_window, int L, char *out)
{
	char *tmp = (char*)malloc((size_t)L + 1);
	if (dir) for (int i = L - 1, j = 0; i >= 0; --i, ++j) { char c = read[i]; tmp[j] = c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'G' ? 'C' : c == 'C' ? 'G' : 'N'; }
	else memcpy(tmp, read, (size_t)L);
	int n = 0, eq = 0;
	for (int tj = 0; tj < L; ++tj) {
		if (ref_window[tj] != tmp[tj]) {
			if (eq > 1) {
				char digits[12]; int nd = 0, v = eq;
				while (v) { digits[nd++] = (char)('0' + v % 10); v /= 10; }
				while (nd) out[n++] = digits[--nd];
			} else for (int i = tj - eq; i < tj; ++i) out[n++] = tmp[i];
			eq = 0;
			out[n++] = tmp[tj];
		} else ++eq;
	}
	if (n == 0) out[n++] = '0';
	free(tmp);
	return n;
}
