// TEST INFRASTRUCTURE (oracle/_ref build only; never linked into the product).
//
// Link-time wrappers (GNU ld --wrap) around the hot-path entry points that the
// reference's kept objects call: kt_for_reads, kt_for_bucket,
// mm_idx_generation, combine_cluster, realign_hash.  Each wrapper times the
// call and, when MC_DUMP=<dir> is set, writes the host-visible state before /
// after it as flat little-endian arrays that tests read with numpy.  The same
// wrappers are linked around the reference objects (oracle arm) and around the
// B200 drop-in (product arm), so the two arms can be compared state by state.
//
// Compiled against the reference's own headers where they lie
// (/root/reference/src/breads.h, kvec.h); nothing is copied.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include "breads.h"
#include "kvec.h"

#define DECL(ret, mangled, ...) \
	extern "C" ret __real_##mangled(__VA_ARGS__); \
	extern "C" ret __wrap_##mangled(__VA_ARGS__)

static const char *dump_dir() { const char *s = getenv("MC_DUMP"); return (s && *s) ? s : 0; }
static double g_t_reads, g_t_bucket, g_t_idx, g_t_combine, g_t_realign;
static int g_n_idx, g_n_realign;
static std::string g_realign_detail;

static FILE *dopen(const char *name)
{
	std::string p = std::string(dump_dir()) + "/" + name;
	FILE *f = fopen(p.c_str(), "wb");
	if (!f) { fprintf(stderr, "mcref: cannot write %s\n", p.c_str()); exit(1); }
	return f;
}
template <class T> static void dump_arr(const char *name, const T *a, size_t n)
{
	FILE *f = dopen(name);
	if (n) fwrite(a, sizeof(T), n, f);
	fclose(f);
}
static void dump_buckets(const char *prefix, mm128_v *B, int nb)
{
	std::vector<uint64_t> cnt(nb), xy;
	for (int i = 0; i < nb; ++i) {
		cnt[i] = B[i].n;
		for (size_t j = 0; j < B[i].n; ++j) { xy.push_back(B[i].a[j].x); xy.push_back(B[i].a[j].y); }
	}
	dump_arr((std::string(prefix) + "_counts.u64").c_str(), cnt.data(), cnt.size());
	dump_arr((std::string(prefix) + "_xy.u64").c_str(), xy.data(), xy.size());
}
static void dump_clusters(const char *prefix, int idx)
{
	std::vector<uint64_t> pert, cn, ca, rl;
	std::string refs;
	for (int t = 0; t < n_threads; ++t) {
		cluster_v *v = &reads->clusters[idx][t];
		pert.push_back(v->n);
		for (size_t i = 0; i < v->n; ++i) {
			cluster_t *p = &v->a[i];
			cn.push_back(p->n);
			for (size_t j = 0; j < p->n; ++j) ca.push_back(p->a[j]);
			size_t l = p->ref ? strlen(p->ref) : 0;
			rl.push_back(l);
			if (l) refs.append(p->ref, l);
		}
	}
	dump_arr((std::string(prefix) + "_pert.u64").c_str(), pert.data(), pert.size());
	dump_arr((std::string(prefix) + "_n.u64").c_str(), cn.data(), cn.size());
	dump_arr((std::string(prefix) + "_a.u64").c_str(), ca.data(), ca.size());
	dump_arr((std::string(prefix) + "_reflen.u64").c_str(), rl.data(), rl.size());
	dump_arr((std::string(prefix) + "_ref.u8").c_str(), refs.data(), refs.size());
}

// ---- kt_for_reads(int, reads_t*, long) : kthread_reads.c:247 ----
DECL(void, _Z12kt_for_readsiP7reads_tl, int nt, reads_t *r, long n)
{
	double t0 = realtime();
	__real__Z12kt_for_readsiP7reads_tl(nt, r, n);
	g_t_reads += realtime() - t0;
	if (!dump_dir()) return;
	dump_buckets("r_B0", r->B[0], 1 << r->b);
	dump_arr("r_allA.u32", r->sp->allA_id.a, r->sp->allA_id.n);
	dump_arr("r_allT.u32", r->sp->allT_id.a, r->sp->allT_id.n);
	dump_arr("r_allN.u32", r->sp->allN_id.a, r->sp->allN_id.n);
	dump_arr("r_fpA.u32", r->fpA_id.a, r->fpA_id.n);
	dump_arr("r_fpT.u32", r->fpT_id.a, r->fpT_id.n);
	dump_arr("r_fpN.u32", r->fpN_id.a, r->fpN_id.n);
	dump_arr("r_Nfile.u32", r->Nfile_id.a, r->Nfile_id.n);
	if (getenv("MC_DUMP_SEQS")) {
		FILE *f = dopen("r_seqs.txt");
		for (long i = 0; i < n; ++i) { fputs(r->seq[i].seq, f); fputc('\n', f); }
		fclose(f);
		std::vector<uint32_t> np;  // per read: count, positions...
		for (long i = 0; i < n; ++i) {
			uint32_v *v = (uint32_v*)r->seq[i].n_pos;
			np.push_back(v ? (uint32_t)v->n : 0);
			if (v) for (size_t j = 0; j < v->n; ++j) np.push_back(v->a[j]);
		}
		dump_arr("r_npos.u32", np.data(), np.size());
	}
}

// ---- kt_for_bucket(int, reads_t*, long) : kthread_bucket.c:562 ----
DECL(void, _Z13kt_for_bucketiP7reads_tl, int nt, reads_t *r, long n)
{
	double t0 = realtime();
	__real__Z13kt_for_bucketiP7reads_tl(nt, r, n);
	g_t_bucket += realtime() - t0;
	if (!dump_dir()) return;
	dump_clusters("b_cl", 0);
	dump_arr("b_sg.u32", r->sg.a, r->sg.n);
	std::vector<mm128_v> B(1 << r->b);
	for (int i = 0; i < (1 << r->b); ++i) B[i] = r->mi[0]->B[i].a;
	dump_buckets("b_mi", B.data(), 1 << r->b);
}

// ---- mm_idx_generation(int, mm_idx_t*) : kthread_idx.c:170 ----
DECL(void, _Z17mm_idx_generationiP8mm_idx_t, int nt, mm_idx_t *mi)
{
	int nb = 1 << reads->b;
	std::vector<std::vector<uint64_t> > keys;
	char name[64];
	if (dump_dir()) {
		std::vector<mm128_v> B(nb);
		keys.resize(nb);
		for (int i = 0; i < nb; ++i) {
			B[i] = mi->B[i].a;
			for (size_t j = 0; j < B[i].n; ++j) keys[i].push_back(B[i].a[j].x);
			std::sort(keys[i].begin(), keys[i].end());
			keys[i].erase(std::unique(keys[i].begin(), keys[i].end()), keys[i].end());
		}
		snprintf(name, sizeof name, "i%d_in", g_n_idx);
		dump_buckets(name, B.data(), nb);
	}
	double t0 = realtime();
	__real__Z17mm_idx_generationiP8mm_idx_t(nt, mi);
	g_t_idx += realtime() - t0;
	if (dump_dir()) {
		std::vector<uint64_t> post; // records: x, n, y[0..n)
		for (int i = 0; i < nb; ++i)
			for (size_t j = 0; j < keys[i].size(); ++j) {
				int n = 0;
				const uint64_t *y = mm_idx_get(mi, keys[i][j], &n);
				post.push_back(keys[i][j]); post.push_back((uint64_t)n);
				for (int k = 0; k < n; ++k) post.push_back(y[k]);
			}
		snprintf(name, sizeof name, "i%d_post.u64", g_n_idx);
		dump_arr(name, post.data(), post.size());
	}
	++g_n_idx;
}

// ---- combine_cluster(int, reads_t*, int*) : kthread_cb.c:570 (host, kept) ----
DECL(void, _Z15combine_clusteriP7reads_tPi, int nt, reads_t *r, int *index)
{
	double t0 = realtime(), i0 = g_t_idx;
	__real__Z15combine_clusteriP7reads_tPi(nt, r, index);
	g_t_combine += (realtime() - t0) - (g_t_idx - i0); // host merge only
	if (dump_dir()) {
		dump_clusters("c_cl", *index);
		uint64_t v = (uint64_t)*index;
		dump_arr("c_idxv.u64", &v, 1);
	}
}

// ---- realign_hash(int, reads_t*, int, int) : kthread_hash_realign.c:569 ----
DECL(void, _Z12realign_hashiP7reads_tii, int nt, reads_t *r, int index, int thr)
{
	char name[64];
	std::vector<uint64_t> pre;
	size_t preA = r->fpA_id.n, preT = r->fpT_id.n;
	if (dump_dir()) {
		for (int t = 0; t < nt; ++t)
			for (size_t i = 0; i < r->clusters[index][t].n; ++i) pre.push_back(r->clusters[index][t].a[i].n);
		snprintf(name, sizeof name, "h%d_sg.u32", g_n_realign);
		dump_arr(name, r->sg.a, r->sg.n);
		uint64_t v = (uint64_t)thr;
		snprintf(name, sizeof name, "h%d_thr.u64", g_n_realign);
		dump_arr(name, &v, 1);
	}
	double t0 = realtime();
	__real__Z12realign_hashiP7reads_tii(nt, r, index, thr);
	double dt = realtime() - t0;
	g_t_realign += dt;
	char buf[128];
	snprintf(buf, sizeof buf, "%s{\"thr\": %d, \"singles\": %zu, \"sec\": %.6f}", g_n_realign ? ", " : "", thr, (size_t)r->sg.n, dt);
	g_realign_detail += buf;
	if (dump_dir()) {
		std::vector<uint64_t> cnt, app;
		size_t c = 0;
		for (int t = 0; t < nt; ++t)
			for (size_t i = 0; i < r->clusters[index][t].n; ++i, ++c) {
				cluster_t *p = &r->clusters[index][t].a[i];
				cnt.push_back(p->n - pre[c]);
				for (size_t j = pre[c]; j < p->n; ++j) app.push_back(p->a[j]);
			}
		snprintf(name, sizeof name, "h%d_app_cnt.u64", g_n_realign);
		dump_arr(name, cnt.data(), cnt.size());
		snprintf(name, sizeof name, "h%d_app_y.u64", g_n_realign);
		dump_arr(name, app.data(), app.size());
		snprintf(name, sizeof name, "h%d_flag.u8", g_n_realign);
		dump_arr(name, (const uint8_t*)r->sg_flag, r->sg.n);
		snprintf(name, sizeof name, "h%d_fpA.u32", g_n_realign);
		dump_arr(name, r->fpA_id.a + preA, r->fpA_id.n - preA);
		snprintf(name, sizeof name, "h%d_fpT.u32", g_n_realign);
		dump_arr(name, r->fpT_id.a + preT, r->fpT_id.n - preT);
	}
	++g_n_realign;
}

// Timing summary, printed when the process exits (after cluster_dump).
struct McrefAtExit {
	~McrefAtExit()
	{
		const char *p = getenv("MC_TIMING");
		if (!p || !*p) return;
		FILE *f = fopen(p, "w");
		if (!f) return;
		fprintf(f, "{\"n_reads\": %d, \"readlen\": %d, \"threads\": %d, \"kt_for_reads\": %.6f, \"kt_for_bucket\": %.6f, "
			"\"mm_idx_generation\": %.6f, \"n_idx\": %d, \"host_combine\": %.6f, \"realign_hash\": %.6f, \"n_realign\": %d, "
			"\"realign_rounds\": [%s]}\n",
			reads ? reads->n_seq : 0, reads ? reads->seq_len : 0, n_threads, g_t_reads, g_t_bucket, g_t_idx, g_n_idx,
			g_t_combine, g_t_realign, g_n_realign, g_realign_detail.c_str());
		fclose(f);
	}
};
static McrefAtExit g_mcref_at_exit;
