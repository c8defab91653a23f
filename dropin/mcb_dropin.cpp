// mcb_dropin.cpp — the reference-side binding of libminicom_b200.so.
//
// Defines, with the reference's own (C++-mangled) signatures, exactly the ten symbols that the KEPT objects of
// yuansliu/minicom (minicommain.o preprocess.o kthread_cb.o kthread_dump[_pe].o misc.o) import from the
// objects this project replaces (sketch.o kthread_reads.o kthread_bucket.o kthread_idx.o kthread_hash_realign.o
// bbhashdict.o) — SURVEY.md §8b:
//     kt_for_reads  kt_for_bucket  mm_idx_init  mm_idx_generation  mm_idx_get  mm_idx_destroy  realign_hash
//     mm_sketch_lh_ori  seq_nt4_table  invert_code_rule
// plus combine_cluster (N1: the contig merge on the device, kthread_cb.o's own is kept under another name), the five
// functions of bseq.o (N3: bseq_open bseq_read bseq_read_second bseq_close bseq_eof — the FASTQ reader, which here packs the
// reads for the device while it parses; build with MCB_KEEP_BSEQ=1 to link the reference's bseq.o instead) and kt_dump_for /
// kt_dump_pe_for (N2: the per-contig half of the dump stage — member sort, diff encoding of every read on the device, and the
// per-thread-slot files; kthread_dump[_pe].o stays linked for cluster_dump[_pe], its own worker is weakened by build_dropin.sh).
// It is compiled inside the reference tree against the reference's headers (breads.h, kvec.h and the generated
// config.h), the way a maintainer would add it; it contains marshalling only — every computation is
// a call into the C-ABI (include/minicom_b200.h).  Cluster placement equals the num_thr=1 layout of the reference
// (all contigs in clusters[idx][0], ids idx<<8), which is the only deterministic configuration of the reference.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "minicom_b200.h"
// config.h defines `readlen` and `ininumdict` as macros: everything that spells those identifiers comes before it
static inline void set_readlen(mcb_params *p, int v) { p->readlen = v; }
#include "breads.h"
#include "kvec.h"
#include "config.h"
#include "mcb_dump_writer.h"

// ---- data symbols (sketch.c:8-25, kthread_bucket.c:64)
unsigned char seq_nt4_table[256] = {
#define R4 4, 4, 4, 4
#define R16 R4, R4, R4, R4
	0, 1, 2, 3, R4, R4, R4, R16, R16, R16,
	4, 0, 4, 1, 4, 4, 4, 2, R4, R4, 4, 4, 4, 4, 3, 4, 4, 4, R4, R4,
	4, 0, 4, 1, 4, 4, 4, 2, R4, R4, 4, 4, 4, 4, 3, 4, 4, 4, R4, R4,
	R16, R16, R16, R16, R16, R16, R16, R16
};
const char invert_code_rule[4] = {'A', 'C', 'G', 'T'};

static mcb_ctx *g_ctx = 0;
static mcb_group *g_grp = 0;    // MCB_DEVICES=0,1,...: the same calls over several GPUs (one context per device inside the library)
static double g_wall[4];      // for_reads, for_bucket, idx_build, realign (host wall seconds inside the entry points)
static int g_calls[4];
static double g_wall_combine = 0;   // combine_cluster (the contig merge)
static double g_wall_dump = 0;      // kt_dump_for (the per-contig half of the dump stage)
static bool g_contigs_on_device = false;   // the device merge handed its contigs to Stage 2: realign_hash need not ship them
static std::string g_realign_detail;

// MCB_RECORD=<dir>: write the host-side inputs of every mm_idx_generation / realign_hash call as flat little-endian
// files, so bench.py can replay the exact sequence of C-ABI calls of this run without the host stages in between.
static const char *record_dir() { const char *s = getenv("MCB_RECORD"); return (s && *s) ? s : 0; }
static void record(const char *name, int idx, const char *suffix, const void *data, size_t bytes)
{
	char path[4096];
	snprintf(path, sizeof path, "%s/%s_%d.%s", record_dir(), name, idx, suffix);
	FILE *f = fopen(path, "wb");
	if (!f || (bytes && fwrite(data, 1, bytes, f) != bytes)) { fprintf(stderr, "minicom_b200: cannot write %s\n", path); exit(1); }
	fclose(f);
}

static void die(const char *where, int rc)
{
	fprintf(stderr, "minicom_b200: %s failed (%d): %s\n", where, rc, mcb_last_error());
	exit(1);
}

// ---- bseq_open / bseq_read / bseq_read_second / bseq_close / bseq_eof (bseq.c:19-101), N3.
// The parser of the library reads the file (zlib, kseq record grammar), hands back the NUL-terminated strings the kept host
// stages go on using (reads->seq[i].seq: dump, N replacement) and keeps every read 2-bit packed in page-locked memory;
// kt_for_reads then uploads those rows (a quarter of the characters) instead of gathering and shipping the strings.
static mcb_readset *g_rs = 0;
#ifndef MCB_KEEP_BSEQ
#include "bseq.h"
struct bseq_file_s { std::string path; int is_eof; };

bseq_file_t *bseq_open(const char *fn)
{
	if (!fn || !strcmp(fn, "-")) { fprintf(stderr, "minicom_b200: reading from stdin is not supported by the packed reader\n"); return 0; }
	FILE *f = fopen(fn, "rb");
	if (!f) return 0;                                        // preprocess.c:46-49 reports it
	fclose(f);
	bseq_file_t *fp = new bseq_file_s();
	fp->path = fn; fp->is_eof = 0;
	return fp;
}

void bseq_close(bseq_file_t *fp) { delete fp; }
int bseq_eof(bseq_file_t *fp) { return fp->is_eof; }

static void read_into(bseq1_t *&seqs, bseq_file_t *fp, int *n_, int seq_len, int n_before)
{
	if (!g_rs) { int rc = mcb_readset_create(seq_len, &g_rs); if (rc) die("bseq_read", rc); }
	char *ascii = 0;
	uint64_t added = 0;
	int rc = mcb_readset_add_fastq(g_rs, fp->path.c_str(), n_threads > 0 ? n_threads : 1, &ascii, &added);
	if (rc) {
		if (strstr(mcb_last_error(), "Length of reads are different")) fprintf(stderr, "Length of reads are different. The program can not compress it.\n");   // bseq.c:55
		die("bseq_read", rc);
	}
	seqs = (bseq1_t*)realloc(seqs, ((size_t)n_before + added + 1) * sizeof(bseq1_t));
	for (uint64_t i = 0; i < added; ++i) {
		bseq1_t *s = &seqs[n_before + i];
		s->rid = 0; s->n_pos = 0;
		s->seq = ascii + i * (size_t)(seq_len + 1);          // one block for the whole file; nothing in the reference frees the strings
	}
	if (added == 0) fp->is_eof = 1;
	*n_ = n_before + (int)added;
}

bseq1_t *bseq_read(bseq_file_t *fp, int *n_, int seq_len)
{
	bseq1_t *seqs = 0;
	if (g_rs) { mcb_readset_destroy(g_rs); g_rs = 0; }
	read_into(seqs, fp, n_, seq_len, 0);
	if (*n_ == 0) { free(seqs); seqs = 0; }
	return seqs;
}

void bseq_read_second(bseq1_t *&seqs, bseq_file_t *fp, int *n_, int seq_len)
{
	read_into(seqs, fp, n_, seq_len, *n_);
}
#endif

struct McbAtExit {
	~McbAtExit()
	{
		const char *p = getenv("MCB_TIMING");
		if (p && *p) {
			FILE *f = fopen(p, "w");
			if (f) {
				static char buf[1 << 16];
				buf[0] = 0;
				if (g_grp) mcb_group_timers_dump(g_grp, buf, sizeof buf);
				else if (g_ctx) mcb_timers_dump(g_ctx, buf, sizeof buf);
				std::string dev = "{";
				for (char *line = strtok(buf, "\n"); line; line = strtok(0, "\n")) {
					char name[128]; double ms; unsigned long long cnt;
					if (sscanf(line, "%127s %lf %llu", name, &ms, &cnt) == 3) {
						char e[256]; snprintf(e, sizeof e, "%s\"%s\": [%.6f, %llu]", dev.size() > 1 ? ", " : "", name, ms, cnt);
						dev += e;
					}
				}
				dev += "}";
				fprintf(f, "{\"n_reads\": %d, \"readlen\": %d, \"threads\": %d, \"kt_for_reads\": %.6f, \"kt_for_bucket\": %.6f, "
				        "\"mm_idx_generation\": %.6f, \"n_idx\": %d, \"realign_hash\": %.6f, \"n_realign\": %d, \"realign_rounds\": [%s], "
				        "\"kernel_launches\": %llu, \"combine_cluster\": %.6f, \"kt_dump_for\": %.6f, \"device_ms\": %s}\n",
				        reads ? reads->n_seq : 0, reads ? reads->seq_len : 0, n_threads, g_wall[0], g_wall[1], g_wall[2], g_calls[2], g_wall[3], g_calls[3],
				        g_realign_detail.c_str(), g_ctx ? (unsigned long long)mcb_kernel_launches(g_ctx) : 0ull, g_wall_combine, g_wall_dump, dev.c_str());     // (group: rank 0's launches, per-timer maximum over the ranks)
				fclose(f);
			}
		}
		if (g_grp) { mcb_group_destroy(g_grp); g_grp = 0; g_ctx = 0; }
		if (g_ctx) { mcb_destroy(g_ctx); g_ctx = 0; }
	}
};
static McbAtExit g_at_exit;

static mcb_ctx *ctx_for(reads_t *r)
{
	if (g_ctx) return g_ctx;
	mcb_params p;
	memset(&p, 0, sizeof p);
	set_readlen(&p, r->seq_len); p.k = r->k; p.b = r->b; p.rw = r->rw;            // resolved by main()/pre_process()
	p.first_mininum = first_mininum; p.diff_threshold = diff_threshold; p.max_rounds = max_rounds;
	const char *dev = getenv("MCB_DEVICE");
	p.device = dev ? atoi(dev) : 0;
	std::vector<int> devs;                                    // MCB_DEVICES=0,1,...  (more than one device: the sharded path)
	if (const char *list = getenv("MCB_DEVICES"))
		for (const char *q = list; *q;) { char *e; long v = strtol(q, &e, 10); if (e == q) break; devs.push_back((int)v); q = *e ? e + 1 : e; }
	if (devs.size() > 1) {
		int rc = mcb_group_create(&p, devs.data(), (int)devs.size(), &g_grp);
		if (rc) die("mcb_group_create", rc);
		g_ctx = mcb_group_context(g_grp, 0);
		if (getenv("MCB_TIMING")) for (int i = 0; i < mcb_group_size(g_grp); ++i) mcb_timers_enable(mcb_group_context(g_grp, i), 1);
		return g_ctx;
	}
	if (devs.size() == 1) p.device = devs[0];
	int rc = mcb_create(&p, &g_ctx);
	if (rc) die("mcb_create", rc);
	if (getenv("MCB_TIMING")) mcb_timers_enable(g_ctx, 1);
	return g_ctx;
}

// ---- kt_for_reads (kthread_reads.c:247)
void kt_for_reads(int n_threads_, reads_t *r, long n)
{
	double t0 = realtime();
	mcb_ctx *ctx = ctx_for(r);
	mcb_reads_result res;
	mcb_readset_view v;
	memset(&v, 0, sizeof v);
	if (g_rs) mcb_readset_get(g_rs, &v);
	int rc;
	if (g_grp) rc = mcb_group_for_reads_ptrs(g_grp, &r->seq[0].seq, sizeof(bseq1_t), (uint64_t)n, n_threads_, &res);
	else if (g_rs && v.n_reads == (uint64_t)n && !getenv("MCB_ASCII_READS"))       // the reads as the shim's own bseq_read packed them
		rc = mcb_for_reads_packed(ctx, v.packed, v.n_reads, v.nread_rid, v.nmask, v.n_nreads, &res);
	else rc = mcb_for_reads_ptrs(ctx, &r->seq[0].seq, sizeof(bseq1_t), (uint64_t)n, n_threads_, &res);
	if (rc) die("kt_for_reads", rc);
	if (g_rs && !g_grp) { mcb_readset_destroy(g_rs); g_rs = 0; }                   // the packed rows are on the device now
	sp_reads_t *s = r->sp;
	for (long i = 0; i < n; ++i) {
		r->seq[i].n_pos = NULL;
		switch (res.cls[i]) {
		case MCB_CLS_ALLA: s->allA++; kv_push(uint32_t, s->allA_id, (uint32_t)i); break;
		case MCB_CLS_ALLT: s->allT++; kv_push(uint32_t, s->allT_id, (uint32_t)i); break;
		case MCB_CLS_ALLN: s->allN++; kv_push(uint32_t, s->allN_id, (uint32_t)i); break;
		case MCB_CLS_FPA: kv_push(uint32_t, r->fpA_id, (uint32_t)i); break;
		case MCB_CLS_FPT: kv_push(uint32_t, r->fpT_id, (uint32_t)i); break;
		case MCB_CLS_FPN: kv_push(uint32_t, r->fpN_id, (uint32_t)i); break;
		case MCB_CLS_NFILE: kv_push(uint32_t, r->Nfile_id, (uint32_t)i); break;
		default: break;
		}
	}
	for (uint64_t j = 0; j < res.n_nreads; ++j) {
		uint32_t rid = res.nread_rid[j];
		uint32_v *np = (uint32_v*)calloc(1, sizeof(uint32_v));
		for (uint64_t q = res.nread_off[j]; q < res.nread_off[j + 1]; ++q) {
			kv_push(uint32_t, *np, res.npos[q]);
			if (res.nread_repl[j]) r->seq[rid].seq[res.npos[q]] = (char)res.nread_repl[j];   // kthread_reads.c:202-204
		}
		r->seq[rid].n_pos = np;
	}
	g_wall[0] += realtime() - t0; g_calls[0]++;
}

// ---- kt_for_bucket (kthread_bucket.c:562)
void kt_for_bucket(int n_threads_, reads_t *r, long n)
{
	double t0 = realtime();
	mcb_ctx *ctx = ctx_for(r);
	mcb_bucket_result res;
	// with the contig merge on the device (the default on one GPU) the host never looks at the seed contigs: only the singles come back
	const bool keep = !g_grp && !getenv("MCB_HOST_MERGE");
	int rc = g_grp ? mcb_group_for_bucket(g_grp, &res) : keep ? mcb_for_bucket_keep(ctx, &res) : mcb_for_bucket(ctx, &res);
	if (rc) die("kt_for_bucket", rc);
	cluster_v *cv = &r->clusters[0][0];
	for (uint64_t c = 0; c < (keep ? 0 : res.n_clusters); ++c) {
		cluster_t *p;
		kv_pushp(cluster_t, *cv, &p);
		kv_init(*p);
		size_t nm = res.cl_n[c], len = (size_t)(res.cl_ref_off[c + 1] - res.cl_ref_off[c]);
		kv_resize(uint64_t, *p, nm);
		memcpy(p->a, res.cl_a + res.cl_a_off[c], nm * sizeof(uint64_t));
		p->n = nm;
		p->ref = (char*)calloc(len + 1, 1);
		memcpy(p->ref, res.cl_ref + res.cl_ref_off[c], len);
		p->ennum = 0;
	}
	for (uint64_t i = 0; i < res.n_sg; ++i) kv_push(uint32_t, r->sg, res.sg[i]);
	const int mask = (1 << r->b) - 1, m = first_mininum;
	for (uint64_t c = 0; c < (keep ? 0 : res.n_clusters); ++c)
		for (int j = 0; j < res.mi_cnt[c]; ++j) {
			const mcb_tuple *t = &res.mi[c * m + j];
			mm128_t v; v.x = t->x; v.y = t->y;
			mm128_v *pp = &r->mi[0]->B[v.x & mask].a;
			kv_push(mm128_t, *pp, v);
		}
	g_wall[1] += realtime() - t0; g_calls[1]++;
}

// ---- combine_cluster (kthread_cb.c:570): the contig merge, on the seed contigs kt_for_bucket left on the device.
// The reference's own host merge is linked under another name (build_dropin.sh) and used when MCB_HOST_MERGE is set or when the
// job runs on several GPUs (the merge needs the contigs and the packed reads of the whole job on one device).
void mcb_ref_combine_cluster(int n_threads, reads_t *reads, int *index_);
void combine_cluster(int n_threads_, reads_t *r, int *index_)
{
	if (getenv("MCB_HOST_MERGE") || g_grp) { mcb_ref_combine_cluster(n_threads_, r, index_); return; }
	double t0 = realtime();
	mcb_ctx *ctx = ctx_for(r);
	mcb_combine_result res;
	int rc = mcb_combine(ctx, cbthreshold, &res);
	if (rc) die("combine_cluster", rc);
	const int index = *index_;
	// what the reference's loop leaves behind (:573-630): the seed contigs destroyed, the final set in clusters[idxv][0] with
	// idxv flipped once per iteration, every mm_idx_t but the caller's first one gone
	for (int t = 0; t < n_threads_; ++t) {
		cluster_v *old = &r->clusters[index][t];
		for (size_t i = 0; i < old->n; ++i) cluster_destroy(old->a[i]);
		old->n = 0;
	}
	const int idxv = index ^ (res.iterations & 1);
	for (int t = 0; t < n_threads_; ++t) { kv_init(r->clusters[idxv][t]); kv_resize(cluster_t, r->clusters[idxv][t], 1 << 10); }
	if (idxv != index) for (int t = 0; t < n_threads_; ++t) { free(r->clusters[index][t].a); kv_init(r->clusters[index][t]); kv_resize(cluster_t, r->clusters[index][t], 1 << 10); }
	cluster_v *cv = &r->clusters[idxv][0];
	for (uint64_t c = 0; c < res.n_clusters; ++c) {
		cluster_t *p;
		kv_pushp(cluster_t, *cv, &p);
		kv_init(*p);
		size_t nm = res.cl_n[c], len = (size_t)(res.cl_ref_off[c + 1] - res.cl_ref_off[c]);
		kv_resize(uint64_t, *p, nm);
		memcpy(p->a, res.cl_a + res.cl_a_off[c], nm * sizeof(uint64_t));
		p->n = nm;
		p->ref = (char*)calloc(len + 1, 1);
		memcpy(p->ref, res.cl_ref + res.cl_ref_off[c], len);
		p->ennum = 0;                                           // written by construct_ref2, never read (kthread_cb.c:329 is commented out)
	}
	mm_idx_destroy(r->mi[index]);                              // the index objects the host loop would have gone through (:585,:628)
	r->mi[index] = 0;
	*index_ = idxv;
	g_contigs_on_device = true;
	g_wall_combine += realtime() - t0;
}

// ---- minimizer index (kthread_idx.c:77-173).  mm_idx_t keeps the reference's layout; the device-built index hangs off B[0].h.
mm_idx_t *mm_idx_init(int b)
{
	mm_idx_t *mi = (mm_idx_t*)calloc(1, sizeof(mm_idx_t));
	mi->B = (mm_idx_bucket_t*)calloc((size_t)1 << b, sizeof(mm_idx_bucket_t));
	return mi;
}

void mm_idx_generation(int n_threads_, mm_idx_t *mi)
{
	double t0 = realtime();
	mcb_ctx *ctx = ctx_for(reads);
	const int nb = 1 << reads->b;
	std::vector<const mcb_tuple*> ptrs(nb);
	std::vector<uint64_t> cnt(nb);
	for (int i = 0; i < nb; ++i) { ptrs[i] = (const mcb_tuple*)mi->B[i].a.a; cnt[i] = mi->B[i].a.n; }
	if (record_dir()) {
		std::vector<uint64_t> off(nb + 1, 0);
		for (int i = 0; i < nb; ++i) off[i + 1] = off[i] + cnt[i];
		std::vector<mcb_tuple> flat(off[nb]);
		for (int i = 0; i < nb; ++i) if (cnt[i]) memcpy(&flat[off[i]], ptrs[i], cnt[i] * sizeof(mcb_tuple));
		record("idx", g_calls[2], "off.u64", off.data(), off.size() * 8);
		record("idx", g_calls[2], "xy.u64", flat.data(), flat.size() * sizeof(mcb_tuple));
	}
	mcb_index *ix = 0;
	int rc = g_grp ? mcb_group_idx_build_scattered(g_grp, ptrs.data(), cnt.data(), n_threads_, &ix) : mcb_idx_build_scattered(ctx, ptrs.data(), cnt.data(), n_threads_, &ix);
	if (rc) die("mm_idx_generation", rc);
	for (int i = 0; i < nb; ++i) { free(mi->B[i].a.a); mi->B[i].a.a = 0; mi->B[i].a.n = mi->B[i].a.m = 0; }   // kthread_idx.c:166-167
	if (mi->B[0].h) mcb_idx_destroy((mcb_index*)mi->B[0].h);
	mi->B[0].h = ix;
	g_wall[2] += realtime() - t0; g_calls[2]++;
}

const uint64_t *mm_idx_get(const mm_idx_t *mi, uint64_t minier, int *n)
{
	return mcb_idx_get((const mcb_index*)mi->B[0].h, minier, n);
}

void mm_idx_destroy(mm_idx_t *mi)
{
	if (mi == 0) return;
	mcb_idx_destroy((mcb_index*)mi->B[0].h);
	for (int i = 0; i < 1 << reads->b; ++i) free(mi->B[i].a.a);
	free(mi->B);
	free(mi);
}

// ---- mm_sketch_lh_ori (sketch.c:116): per-contig call from the host merger (kthread_cb.c:234,365,418)
void mm_sketch_lh_ori(const char *str, int len, int w, int k, uint32_t rid, mm128_v *p)
{
	mcb_tuple stackbuf[512];
	int64_t n = mcb_sketch_lh_host(str, len, w, k, rid, stackbuf, 512);
	const mcb_tuple *src = stackbuf;
	std::vector<mcb_tuple> big;
	if (n > 512) { big.resize((size_t)n); mcb_sketch_lh_host(str, len, w, k, rid, big.data(), n); src = big.data(); }
	for (int64_t i = 0; i < n; ++i) { mm128_t v; v.x = src[i].x; v.y = src[i].y; kv_push(mm128_t, *p, v); }
}

// ---- realign_hash (kthread_hash_realign.c:569)
void realign_hash(int n_threads_, reads_t *r, int index, int max_threshold)
{
	double t0 = realtime();
	mcb_ctx *ctx = ctx_for(r);
	std::vector<cluster_t*> contigs;
	std::vector<uint64_t> off;
	std::string refs;
	for (int t = 0; t < n_threads_; ++t)
		for (size_t i = 0; i < r->clusters[index][t].n; ++i) {
			cluster_t *p = &r->clusters[index][t].a[i];
			qsort(p->a, p->n, sizeof(uint64_t), cmpcluster2);          // kthread_hash_realign.c:318
			contigs.push_back(p);
			off.push_back(refs.size());
			refs.append(p->ref);
		}
	off.push_back(refs.size());
	if (record_dir()) {
		uint64_t meta[3] = { (uint64_t)max_threshold, (uint64_t)maxsearch, (uint64_t)ininumdict };
		record("realign", g_calls[3], "meta.u64", meta, sizeof meta);
		record("realign", g_calls[3], "sg.u32", r->sg.a, r->sg.n * 4);
		record("realign", g_calls[3], "off.u64", off.data(), off.size() * 8);
		record("realign", g_calls[3], "refs.u8", refs.data(), refs.size());
	}
	// Between two realign_hash calls of one run only updateSingle() executes (preprocess.c:197-232): the consensus strings are
	// the ones of the previous round, so only the first round ships them; later rounds reuse the device-side contig table.
	// "Same contigs" is decided on content (count, total length, a 64-bit hash of the strings), never on pointers.
	static uint64_t sent_sig[3] = { 0, 0, 0 };
	uint64_t sig[3] = { (uint64_t)contigs.size(), (uint64_t)refs.size(), 0x9E3779B97F4A7C15ull };
	{
		const size_t nw = refs.size() / 8;
		uint64_t h[4] = { 1, 2, 3, 4 };                      // four independent lanes: the multiply chain is latency bound
		const char *d = refs.data();
		size_t i = 0;
		for (; i + 4 <= nw; i += 4)
			for (int q = 0; q < 4; ++q) { uint64_t w; memcpy(&w, d + (i + q) * 8, 8); h[q] = (h[q] ^ w) * 0x9E3779B97F4A7C15ull; }
		for (; i < nw; ++i) { uint64_t w; memcpy(&w, d + i * 8, 8); h[0] = (h[0] ^ w) * 0x9E3779B97F4A7C15ull; }
		uint64_t tail = 0; memcpy(&tail, d + nw * 8, refs.size() - nw * 8);
		sig[2] = ((h[0] ^ tail) * 0x9E3779B97F4A7C15ull) ^ (h[1] * 3) ^ (h[2] * 5) ^ (h[3] * 7);
		for (size_t c = 0; c < off.size(); ++c) sig[2] = (sig[2] ^ off[c]) * 0x9E3779B97F4A7C15ull;
	}
	const bool same = ((g_calls[3] > 0 && !memcmp(sig, sent_sig, sizeof sig)) || (g_calls[3] == 0 && g_contigs_on_device)) && !getenv("MCB_RESEND_CONTIGS");
	mcb_realign_result res;
	const char *rp = same ? NULL : refs.data();
	const uint64_t *op = same ? NULL : off.data();
	int rc = g_grp ? mcb_group_realign(g_grp, r->sg.a, r->sg.n, rp, op, contigs.size(), max_threshold, maxsearch, ininumdict, &res)
	               : mcb_realign(ctx, r->sg.a, r->sg.n, rp, op, contigs.size(), max_threshold, maxsearch, ininumdict, &res);
	memcpy(sent_sig, sig, sizeof sig);
	if (rc) die("realign_hash", rc);
	for (uint64_t i = 0; i < res.n_fpA; ++i) { r->sg_flag[res.fpA_sg[i]] = true; kv_push(uint32_t, r->fpA_id, r->sg.a[res.fpA_sg[i]]); }
	for (uint64_t i = 0; i < res.n_fpT; ++i) { r->sg_flag[res.fpT_sg[i]] = true; kv_push(uint32_t, r->fpT_id, r->sg.a[res.fpT_sg[i]]); }
	for (uint64_t i = 0; i < res.n_claims; ++i) {
		cluster_t *p = contigs[res.claim_contig[i]];
		kv_push(uint64_t, *p, res.claim_y[i]);
		r->sg_flag[res.claim_sg[i]] = true;
	}
	double dt = realtime() - t0;
	char buf[160];
	snprintf(buf, sizeof buf, "%s{\"thr\": %d, \"singles\": %zu, \"sec\": %.6f, \"claims\": %llu}", g_calls[3] ? ", " : "", max_threshold, (size_t)r->sg.n, dt,
	         (unsigned long long)res.n_claims);
	g_realign_detail += buf;
	g_wall[3] += dt; g_calls[3]++;
}

// ---- kt_dump_for (kthread_dump.c:333) / kt_dump_pe_for (kthread_dump_pe.c:180): N2.
// The diff encoding of every member read (print_encode's loop) runs on the device, on the reads kt_for_reads left there; the
// writer (mcb_dump_writer.h) lays the bytes out in the reference's files.  On several GPUs the reads are spread over the ranks,
// so the encoding uses a context of its own on the first device, loaded from the packed read set.  MCB_HOST_DUMP=1 runs the
// reference's own workers instead (linked as mcb_ref_kt_dump_[pe_]for by build_dropin.sh).
#ifndef MCB_KEEP_DUMP
struct DeviceEncoder : McbDumpEncoder {
	mcb_ctx *ctx;
	std::vector<uint64_t> off;
	explicit DeviceEncoder(mcb_ctx *c) : ctx(c) {}
	int encode(const uint64_t *members, const uint64_t *moff, const char *refs, const uint64_t *roff, uint64_t nc, const uint64_t **enc_off, const char **enc)
	{
		mcb_encode_result res;
		int rc = mcb_dump_encode(ctx, members, moff, refs, roff, nc, &res);
		if (rc) die("cluster_dump (mcb_dump_encode)", rc);
		*enc_off = res.enc_off; *enc = res.enc;
		return 0;
	}
};
#ifdef _PE
void mcb_ref_kt_dump_pe_for(int n_threads, reads_t *reads, int index);
void kt_dump_pe_for(int n_threads_, reads_t *r, int index)
#else
void mcb_ref_kt_dump_for(int n_threads, reads_t *reads, int index);
void kt_dump_for(int n_threads_, reads_t *r, int index)
#endif
{
	if (getenv("MCB_HOST_DUMP")) {
#ifdef _PE
		mcb_ref_kt_dump_pe_for(n_threads_, r, index);
#else
		mcb_ref_kt_dump_for(n_threads_, r, index);
#endif
		return;
	}
	double t0 = realtime();
	mcb_ctx *ctx = ctx_for(r);
	mcb_ctx *own = 0;
	if (g_grp) {                                                 // reads are spread over the ranks: a context of its own on the first device
		mcb_params p;
		memset(&p, 0, sizeof p);
		set_readlen(&p, r->seq_len); p.k = r->k; p.b = r->b; p.rw = r->rw;
		p.first_mininum = first_mininum; p.diff_threshold = diff_threshold; p.max_rounds = max_rounds;
		p.device = 0;
		if (const char *list = getenv("MCB_DEVICES")) p.device = atoi(list);
		int rc = mcb_create(&p, &own);
		if (rc) die("cluster_dump (mcb_create)", rc);
		mcb_reads_result rr;
		mcb_readset_view v;
		memset(&v, 0, sizeof v);
		if (g_rs) mcb_readset_get(g_rs, &v);
		// the host strings carry the N replacement by now, the packed read set does not need it: both give the same table + N side table
		rc = (g_rs && v.n_reads == (uint64_t)r->n_seq) ? mcb_for_reads_packed(own, v.packed, v.n_reads, v.nread_rid, v.nmask, v.n_nreads, &rr)
		                                                : MCB_ESTATE;
		if (rc == MCB_ESTATE) { fprintf(stderr, "minicom_b200: the dump stage on several GPUs needs the packed read set of the shim's FASTQ reader (or MCB_HOST_DUMP=1)\n"); exit(1); }
		if (rc) die("cluster_dump (mcb_for_reads_packed)", rc);
		ctx = own;
	}
	DeviceEncoder E(ctx);
	mcb_dump_workers(n_threads_, r, index, E);
	if (own) mcb_destroy(own);
	g_wall_dump += realtime() - t0;
}
#endif
