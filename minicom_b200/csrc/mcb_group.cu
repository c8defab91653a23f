// mcb_group.cu — several GPUs behind the single-GPU call shapes, for a host program that is ONE process (the reference's
// minicommain.c behind dropin/mcb_dropin.cpp): one context per device, one host thread per context for the duration of a call,
// an NCCL communicator made with ncclCommInitAll, and the merge of the ranks' results into exactly what one GPU returns
// (round-major / rank-minor for kt_for_bucket, by claim priority for realign_hash — see mcb_shard.cu).
#include "mcb_common.cuh"
#include <nccl.h>
#include <algorithm>
#include <functional>
#include <thread>

struct mcb_index;
mcb_index *mcb_index_make_group(std::vector<mcb_index*> &parts, int b);          // mcb_index.cu

struct mcb_group {
	int G = 0;
	mcb_params prm;
	std::vector<mcb_ctx*> ctx;
	std::vector<int> devices;
	std::vector<ncclComm_t> comms;
	uint64_t n_reads = 0;
	std::vector<uint8_t> sg_owner;              // per read id: rank that produced it as a single in Stage 1 (0xFF: none)
	// merged results (the arrays the result structs point into)
	std::vector<uint8_t> cls, nrepl, mi_cnt;
	std::vector<uint32_t> nrid, npos, cl_n, sg, claim_c, claim_s, fpA, fpT;
	std::vector<uint64_t> noff, cl_a_off, cl_a, cl_ref_off, claim_y, claim_p;
	std::vector<char> cl_ref;
	std::vector<mcb_tuple> mi;
};

// run fn(rank) on one thread per rank; the first failure's message becomes the caller's mcb_last_error()
static int group_run(mcb_group *g, const std::function<int(int)> &fn)
{
	std::vector<int> rc((size_t)g->G, MCB_OK);
	std::vector<std::string> msg((size_t)g->G);
	std::vector<std::thread> th;
	for (int r = 0; r < g->G; ++r)
		th.emplace_back([&, r] { rc[r] = fn(r); if (rc[r] != MCB_OK) msg[r] = mcb_last_error(); });
	for (auto &t : th) t.join();
	for (int r = 0; r < g->G; ++r)
		if (rc[r] != MCB_OK) { mcb_set_error("device %d (rank %d): %s", g->ctx[r]->prm.device, r, msg[r].c_str()); return rc[r]; }
	return MCB_OK;
}

extern "C" void mcb_group_destroy(mcb_group *g)
{
	if (!g) return;
	for (auto c : g->ctx) if (c) mcb_destroy(c);                 // (the contexts do not own the communicators)
	for (size_t r = 0; r < g->comms.size(); ++r) if (g->comms[r]) { cudaSetDevice(g->devices[r]); ncclCommDestroy(g->comms[r]); }
	delete g;
}

extern "C" int mcb_group_create(const mcb_params *p, const int *devices, int n_devices, mcb_group **out)
{
	if (!p || !devices || !out || n_devices < 1 || n_devices > 64) { mcb_set_error("mcb_group_create: bad arguments"); return MCB_EINVAL; }
	*out = nullptr;
	mcb_group *g = new mcb_group();
	g->G = n_devices; g->prm = *p; g->devices.assign(devices, devices + n_devices);
	g->ctx.assign((size_t)n_devices, nullptr); g->comms.assign((size_t)n_devices, nullptr);
	for (int r = 0; r < n_devices; ++r) {
		mcb_params q = *p; q.device = devices[r];
		const int rc = mcb_create(&q, &g->ctx[r]);
		if (rc != MCB_OK) { mcb_group_destroy(g); return rc; }
	}
	if (n_devices > 1) {
		ncclResult_t nr = ncclCommInitAll(g->comms.data(), n_devices, devices);
		if (nr != ncclSuccess) { mcb_set_error("ncclCommInitAll failed: %s", ncclGetErrorString(nr)); mcb_group_destroy(g); return MCB_ECUDA; }
		for (int r = 0; r < n_devices; ++r) {
			const int rc = mcb_shard_attach(g->ctx[r], g->comms[r], r, n_devices);
			if (rc != MCB_OK) { mcb_group_destroy(g); return rc; }
		}
	}
	*out = g;
	return MCB_OK;
}

extern "C" int mcb_group_size(const mcb_group *g) { return g ? g->G : 0; }
extern "C" mcb_ctx *mcb_group_context(mcb_group *g, int rank) { return (g && rank >= 0 && rank < g->G) ? g->ctx[rank] : nullptr; }

static inline void rid_range(uint64_t n, int r, int G, uint64_t *lo, uint64_t *hi)
{
	const uint64_t cap = (n + G - 1) / G;
	*lo = std::min<uint64_t>((uint64_t)r * cap, n); *hi = std::min<uint64_t>(*lo + cap, n);
}

// ---- kt_for_reads: rank r loads the read-id range [lo_r, hi_r)
extern "C" int mcb_group_for_reads_ptrs(mcb_group *g, const void *first_seq_ptr, size_t stride, uint64_t n, int n_threads, mcb_reads_result *res)
{
	if (!g || !res || (n && !first_seq_ptr)) { mcb_set_error("mcb_group_for_reads_ptrs: null argument"); return MCB_EINVAL; }
	const int G = g->G;
	std::vector<mcb_reads_result> rr((size_t)G);
	g->n_reads = n;
	MCB_TRY(group_run(g, [&](int r) -> int {
		uint64_t lo, hi; rid_range(n, r, G, &lo, &hi);
		MCB_TRY(mcb_shard_begin(g->ctx[r], n, lo));
		return mcb_for_reads_ptrs(g->ctx[r], (const char*)first_seq_ptr + lo * stride, stride, hi - lo, std::max(1, n_threads / G), &rr[r]);
	}));
	g->cls.clear(); g->nrid.clear(); g->nrepl.clear(); g->noff.assign(1, 0); g->npos.clear();
	uint64_t n_sk = 0;
	for (int r = 0; r < G; ++r) {
		const mcb_reads_result &q = rr[r];
		g->cls.insert(g->cls.end(), q.cls, q.cls + q.n_reads);
		g->nrid.insert(g->nrid.end(), q.nread_rid, q.nread_rid + q.n_nreads);
		g->nrepl.insert(g->nrepl.end(), q.nread_repl, q.nread_repl + q.n_nreads);
		const uint64_t base = g->npos.size();
		for (uint64_t i = 0; i < q.n_nreads; ++i) g->noff.push_back(base + q.nread_off[i + 1]);
		g->npos.insert(g->npos.end(), q.npos, q.npos + q.nread_off[q.n_nreads]);
		n_sk += q.n_sketched;
	}
	res->n_reads = n; res->cls = g->cls.data();
	res->n_nreads = g->nrid.size(); res->nread_rid = g->nrid.data(); res->nread_repl = g->nrepl.data(); res->nread_off = g->noff.data(); res->npos = g->npos.data();
	res->n_sketched = n_sk;
	return MCB_OK;
}

// ---- kt_for_bucket: the job's lists = the ranks' contributions in rank order, round by round
extern "C" int mcb_group_for_bucket(mcb_group *g, mcb_bucket_result *res)
{
	if (!g || !res) { mcb_set_error("mcb_group_for_bucket: null argument"); return MCB_EINVAL; }
	const int G = g->G, m = g->prm.first_mininum, CAP = 64;
	std::vector<mcb_bucket_result> br((size_t)G);
	std::vector<uint64_t> rc((size_t)G * 4 * CAP, 0);
	MCB_TRY(group_run(g, [&](int r) -> int { return mcb_shard_for_bucket(g->ctx[r], &br[r], &rc[(size_t)r * 4 * CAP], CAP); }));
	int rounds = 0;
	for (int r = 0; r < G; ++r) rounds = std::max(rounds, (int)br[r].rounds);
	g->cl_n.clear(); g->cl_a_off.assign(1, 0); g->cl_a.clear(); g->cl_ref_off.assign(1, 0); g->cl_ref.clear(); g->sg.clear(); g->mi_cnt.clear(); g->mi.clear();
	g->sg_owner.assign(g->n_reads, 0xFF);
	std::vector<uint64_t> c_cl((size_t)G, 0), c_sg((size_t)G, 0);
	uint64_t n_sk = 0, n_grp = 0;
	for (int rd = 0; rd < rounds; ++rd)
		for (int r = 0; r < G; ++r) {
			if (rd >= br[r].rounds) continue;
			const mcb_bucket_result &q = br[r];
			const uint64_t ncl = rc[((size_t)r * CAP + rd) * 4], nsg = rc[((size_t)r * CAP + rd) * 4 + 3];
			for (uint64_t c = c_cl[r]; c < c_cl[r] + ncl; ++c) {
				g->cl_n.push_back(q.cl_n[c]);
				g->cl_a.insert(g->cl_a.end(), q.cl_a + q.cl_a_off[c], q.cl_a + q.cl_a_off[c + 1]);
				g->cl_a_off.push_back(g->cl_a.size());
				g->cl_ref.insert(g->cl_ref.end(), q.cl_ref + q.cl_ref_off[c], q.cl_ref + q.cl_ref_off[c + 1]);
				g->cl_ref_off.push_back(g->cl_ref.size());
				g->mi_cnt.push_back(q.mi_cnt[c]);
				g->mi.insert(g->mi.end(), q.mi + c * m, q.mi + (c + 1) * m);
			}
			for (uint64_t i = c_sg[r]; i < c_sg[r] + nsg; ++i) { g->sg.push_back(q.sg[i]); if (q.sg[i] < g->n_reads) g->sg_owner[q.sg[i]] = (uint8_t)r; }
			c_cl[r] += ncl; c_sg[r] += nsg;
		}
	for (int r = 0; r < G; ++r) { n_sk += br[r].n_sketched_total; n_grp += br[r].n_grouped; }
	res->n_clusters = g->cl_n.size(); res->cl_n = g->cl_n.data(); res->cl_a_off = g->cl_a_off.data(); res->cl_a = g->cl_a.data();
	res->cl_ref_off = g->cl_ref_off.data(); res->cl_ref = g->cl_ref.data(); res->n_sg = g->sg.size(); res->sg = g->sg.data();
	res->mi_cnt = g->mi_cnt.data(); res->mi = g->mi.data(); res->rounds = rounds; res->n_sketched_total = n_sk; res->n_grouped = n_grp;
	return MCB_OK;
}

// ---- mm_idx_generation: every rank builds the buckets it owns; lookups go to the owner's part
extern "C" int mcb_group_idx_build_scattered(mcb_group *g, const mcb_tuple *const *ptrs, const uint64_t *cnt, int n_threads, mcb_index **out)
{
	if (!g || !ptrs || !cnt || !out) { mcb_set_error("mcb_group_idx_build_scattered: null argument"); return MCB_EINVAL; }
	const int G = g->G, nb = 1 << g->prm.b;
	*out = nullptr;
	std::vector<mcb_index*> parts((size_t)G, nullptr);
	std::vector<std::vector<uint64_t>> mine((size_t)G);
	const int rcode = group_run(g, [&](int r) -> int {
		const int b0 = (int)(((int64_t)r * nb + G - 1) / G), b1 = (int)(((int64_t)(r + 1) * nb + G - 1) / G);      // owner = bucket * G >> b
		mine[r].assign((size_t)nb, 0);
		for (int i = b0; i < b1; ++i) mine[r][i] = cnt[i];
		return mcb_idx_build_scattered(g->ctx[r], ptrs, mine[r].data(), std::max(1, n_threads / G), &parts[r]);
	});
	if (rcode != MCB_OK) { for (auto p : parts) if (p) mcb_idx_destroy(p); return rcode; }
	*out = mcb_index_make_group(parts, g->prm.b);
	return MCB_OK;
}

// ---- realign_hash: a single is realigned by the rank that produced it in Stage 1
extern "C" int mcb_group_realign(mcb_group *g, const uint32_t *sg, uint64_t n_sg, const char *refs, const uint64_t *ref_off, uint64_t n_contigs,
                                 int threshold, int maxsearch, int ininumdict, mcb_realign_result *res)
{
	if (!g || !res || (n_sg && !sg)) { mcb_set_error("mcb_group_realign: null argument"); return MCB_EINVAL; }
	const int G = g->G;
	std::vector<std::vector<uint32_t>> loc((size_t)G), pos((size_t)G);
	for (uint64_t i = 0; i < n_sg; ++i) {
		const uint32_t rid = sg[i];
		const uint8_t o = rid < g->sg_owner.size() ? g->sg_owner[rid] : 0xFF;
		if (o == 0xFF) { mcb_set_error("mcb_group_realign: read %u is not a single of this job's kt_for_bucket", rid); return MCB_EINVAL; }
		loc[o].push_back(rid); pos[o].push_back((uint32_t)i);
	}
	std::vector<mcb_realign_result> rr((size_t)G);
	MCB_TRY(group_run(g, [&](int r) -> int {
		return mcb_shard_realign(g->ctx[r], loc[r].data(), pos[r].data(), loc[r].size(), n_sg, refs, ref_off, n_contigs, threshold, maxsearch, ininumdict, &rr[r]);
	}));
	// claims: merge the ranks' lists by (priority ascending, position in sg descending); diversions: ascending position
	uint64_t tot = 0;
	for (int r = 0; r < G; ++r) tot += rr[r].n_claims;
	g->claim_c.resize(tot); g->claim_s.resize(tot); g->claim_y.resize(tot); g->claim_p.resize(tot);
	std::vector<uint64_t> at((size_t)G, 0);
	for (uint64_t o = 0; o < tot; ++o) {
		int best = -1;
		for (int r = 0; r < G; ++r) {
			if (at[r] >= rr[r].n_claims) continue;
			if (best < 0) { best = r; continue; }
			const uint64_t pa = rr[r].claim_prio[at[r]], pb = rr[best].claim_prio[at[best]];
			if (pa < pb || (pa == pb && rr[r].claim_sg[at[r]] > rr[best].claim_sg[at[best]])) best = r;
		}
		const uint64_t i = at[best]++;
		g->claim_c[o] = rr[best].claim_contig[i]; g->claim_s[o] = rr[best].claim_sg[i]; g->claim_y[o] = rr[best].claim_y[i]; g->claim_p[o] = rr[best].claim_prio[i];
	}
	g->fpA.clear(); g->fpT.clear();
	for (int r = 0; r < G; ++r) { g->fpA.insert(g->fpA.end(), rr[r].fpA_sg, rr[r].fpA_sg + rr[r].n_fpA); g->fpT.insert(g->fpT.end(), rr[r].fpT_sg, rr[r].fpT_sg + rr[r].n_fpT); }
	std::sort(g->fpA.begin(), g->fpA.end()); std::sort(g->fpT.begin(), g->fpT.end());
	*res = rr[0];
	res->n_claims = tot; res->claim_contig = g->claim_c.data(); res->claim_sg = g->claim_s.data(); res->claim_y = g->claim_y.data(); res->claim_prio = g->claim_p.data();
	res->n_fpA = g->fpA.size(); res->n_fpT = g->fpT.size(); res->fpA_sg = g->fpA.data(); res->fpT_sg = g->fpT.data();
	res->n_candidates = 0;
	for (int r = 0; r < G; ++r) res->n_candidates += rr[r].n_candidates;
	return MCB_OK;
}

// device-time accounting of the whole group: per timer the maximum over the ranks (the ranks run side by side)
extern "C" size_t mcb_group_timers_dump(mcb_group *g, char *buf, size_t cap)
{
	std::map<std::string, std::pair<double, uint64_t>> mx;
	if (g) for (auto c : g->ctx)
		for (size_t i = 0; i < c->tm.names.size(); ++i) {
			auto &e = mx[c->tm.names[i]];
			e.first = std::max(e.first, c->tm.ms[i]); e.second = std::max<uint64_t>(e.second, c->tm.cnt[i]);
		}
	std::string s;
	for (auto &kv : mx) { char line[256]; snprintf(line, sizeof line, "%s %.6f %llu\n", kv.first.c_str(), kv.second.first, (unsigned long long)kv.second.second); s += line; }
	if (buf && cap) { size_t n = s.size() < cap - 1 ? s.size() : cap - 1; memcpy(buf, s.data(), n); buf[n] = 0; }
	return s.size() + 1;
}
