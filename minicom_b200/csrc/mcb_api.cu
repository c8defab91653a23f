// mcb_api.cu — context lifetime, error reporting, timers and the host-side boundary helpers of the C-ABI.
#include "mcb_common.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <algorithm>

static thread_local char g_err[1024] = "";

void mcb_set_error(const char *fmt, ...)
{
	va_list ap; va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
}

extern "C" const char *mcb_last_error(void) { return g_err; }
extern "C" const char *mcb_version(void) { return "minicom_b200 0.1 (sm_100a)"; }

// minicommain.c:92-127 and preprocess.c:89-107
extern "C" void mcb_resolve_params(mcb_params *p, int readlen, int inik, int inithr, int iniw, int inim, int inimaxrounds)
{
	memset(p, 0, sizeof *p);
	p->readlen = readlen;
	p->k = readlen < 80 ? 17 : 31;
	if (inik > 0) p->k = inik;
	p->diff_threshold = inithr > 0 ? inithr : 4;
	p->first_mininum = inim > 0 ? inim : 6;
	p->max_rounds = 35;
	if (inimaxrounds > 0 && inimaxrounds < p->max_rounds) p->max_rounds = inimaxrounds;
	p->b = 14;
	p->rw = readlen >= 70 ? readlen / 2 - p->k : 3;
	if (iniw > 0) p->rw = iniw;
	p->device = 0;
}

extern "C" int mcb_create(const mcb_params *p, mcb_ctx **out)
{
	if (!p || !out) { mcb_set_error("mcb_create: null argument"); return MCB_EINVAL; }
	*out = nullptr;
	if (p->readlen < 12 || p->readlen > 256) { mcb_set_error("readlen %d out of range [12,256] (minicom:51-54)", p->readlen); return MCB_EINVAL; }
	if (p->k < 10 || p->k > 31 || p->k > p->readlen) { mcb_set_error("k=%d out of range [10,31]", p->k); return MCB_EINVAL; }
	if (p->b != 14) { mcb_set_error("bucket bits b=%d unsupported (the reference fixes b=14, minicommain.c:175)", p->b); return MCB_EINVAL; }
	if (p->rw < 1 || p->rw > MCB_LH_WMAX) { mcb_set_error("contig window rw=%d out of range [1,%d]", p->rw, MCB_LH_WMAX); return MCB_EINVAL; }
	if (p->first_mininum < 1 || p->first_mininum > 255) { mcb_set_error("first_mininum=%d out of range [1,255]", p->first_mininum); return MCB_EINVAL; }
	if (p->diff_threshold < 0 || p->max_rounds < 2 || p->max_rounds > 35) { mcb_set_error("bad diff_threshold/max_rounds"); return MCB_EINVAL; }
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev <= 0) {
		mcb_set_error("no CUDA device available (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
		cudaGetLastError();
		return MCB_ECUDA;
	}
	if (p->device < 0 || p->device >= ndev) { mcb_set_error("device %d out of range (have %d)", p->device, ndev); return MCB_EINVAL; }
	MCB_CUDA(cudaSetDevice(p->device));
	mcb_ctx *ctx = new mcb_ctx();
	ctx->prm = *p;
	ctx->L = p->readlen; ctx->Wd = (p->readlen + 31) / 32; ctx->WS = (ctx->Wd + 1) & ~1;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, p->device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
	if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; mcb_set_error("cudaStreamCreate failed"); return MCB_ECUDA; }
	ctx->tm.stream = ctx->stream;
	ctx->pool = new McbPinnedPool();
	if (ctx->d_counters.ensure(64 * 8) != MCB_OK) { cudaStreamDestroy(ctx->stream); delete ctx; return MCB_ECUDA; }
	cudaMemsetAsync(ctx->d_counters.p, 0, 64 * 8, ctx->stream);
	*out = ctx;
	return MCB_OK;
}

void mcb_shard_release(mcb_ctx *ctx);      // mcb_shard.cu
void mcb_combine_release(mcb_ctx *ctx);    // mcb_combine.cu

extern "C" void mcb_destroy(mcb_ctx *ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->prm.device);
	cudaStreamSynchronize(ctx->stream);
	mcb_shard_release(ctx);
	mcb_combine_release(ctx);
	ctx->tm.collect();
	for (auto e : ctx->tm.pool) cudaEventDestroy(e);
	DBuf *db[] = { &ctx->d_ascii, &ctx->d_packed, &ctx->d_cls, &ctx->d_elemA, &ctx->d_elemB, &ctx->d_counters, &ctx->d_nread_rid, &ctx->d_nread_mask, &ctx->d_sort_hist,
	               &ctx->d_rows_send, &ctx->d_rows_recv, &ctx->d_coll, &ctx->d_subthr };
	for (auto b : db) b->release();
	for (auto &b : ctx->d_scr) b.release();
	for (auto &b : ctx->d_scan_tmp) b.release();
	for (auto &b : ctx->d_x) b.release();
	for (auto &b : ctx->d_out) b.release();
	{ McbContigIndex &c = ctx->cix; DBuf *cb[] = { &c.refs, &c.roff, &c.cwo, &c.wo, &c.cw, &c.pblk, &c.ptab, &c.ents, &c.ents2, &c.eoff, &c.meta, &c.flt, &c.sgmap }; for (auto b : cb) b->release(); }
	HBuf *hb[] = { &ctx->h_cls, &ctx->h_nrid, &ctx->h_nrepl, &ctx->h_noff, &ctx->h_npos, &ctx->h_nmask, &ctx->h_counters, &ctx->h_stage,
	               &ctx->h_cl_n, &ctx->h_cl_a_off, &ctx->h_cl_a, &ctx->h_cl_ref_off, &ctx->h_cl_ref, &ctx->h_sg, &ctx->h_mi_cnt, &ctx->h_mi,
	               &ctx->h_claim_c, &ctx->h_claim_s, &ctx->h_claim_y, &ctx->h_claim_p, &ctx->h_fpA, &ctx->h_fpT, &ctx->h_in0, &ctx->h_in1, &ctx->h_in2, &ctx->h_coll };
	for (auto b : hb) b->release();
	cudaStreamDestroy(ctx->stream);
	if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
	if (ctx->copy_stream2) cudaStreamDestroy(ctx->copy_stream2);
	if (ctx->pool && ctx->pool->close()) delete ctx->pool;
	delete ctx;
}

// ---------------------------------------------------------------- host -> device staging
#include <thread>
static void par_memcpy(char *dst, const char *src, size_t bytes, int n_threads)
{
	int T = (int)std::min<size_t>((size_t)std::max(1, n_threads), (bytes + (4u << 20) - 1) / (4u << 20));
	if (T <= 1) { memcpy(dst, src, bytes); return; }
	std::vector<std::thread> th;
	for (int t = 0; t < T; ++t) {
		size_t a = bytes * t / T, b = bytes * (t + 1) / T;
		th.emplace_back([=] { memcpy(dst + a, src + a, b - a); });
	}
	for (auto &x : th) x.join();
}

int mcb_copy_streams(mcb_ctx *ctx)
{
	if (!ctx->copy_stream) MCB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
	if (!ctx->copy_stream2) MCB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream2, cudaStreamNonBlocking));
	return MCB_OK;
}
int mcb_h2d(mcb_ctx *ctx, void *dst, const void *src, size_t bytes, int n_threads) { return mcb_h2d_on(ctx, ctx->stream, dst, src, bytes, n_threads); }
int mcb_h2d_on(mcb_ctx *ctx, cudaStream_t stream, void *dst, const void *src, size_t bytes, int n_threads)
{
	if (!bytes) return MCB_OK;
	cudaPointerAttributes pa;
	const bool pinned = cudaPointerGetAttributes(&pa, src) == cudaSuccess && pa.type == cudaMemoryTypeHost;
	cudaGetLastError();
	if (pinned) { MCB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream)); return MCB_OK; }
	const size_t CH = 16u << 20;
	MCB_TRY(ctx->h_stage.ensure(2 * CH));
	cudaEvent_t ev[2];
	cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming); cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
	int slot = 0, rc = MCB_OK;
	for (size_t o = 0; o < bytes; o += CH, slot ^= 1) {
		const size_t len = std::min(CH, bytes - o);
		char *st = ctx->h_stage.as<char>() + (size_t)slot * CH;
		cudaEventSynchronize(ev[slot]);
		par_memcpy(st, (const char*)src + o, len, n_threads);
		if (cudaMemcpyAsync((char*)dst + o, st, len, cudaMemcpyHostToDevice, stream) != cudaSuccess) { mcb_set_error("h2d staging copy failed: %s", cudaGetErrorString(cudaGetLastError())); rc = MCB_ECUDA; break; }
		cudaEventRecord(ev[slot], stream);
	}
	cudaEventSynchronize(ev[0]); cudaEventSynchronize(ev[1]);
	cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
	return rc;
}

// ---------------------------------------------------------------- timers
extern "C" void mcb_timers_enable(mcb_ctx *ctx, int on) { if (ctx) ctx->tm.enabled = on != 0; }
extern "C" void mcb_timers_reset(mcb_ctx *ctx) { if (ctx) { cudaStreamSynchronize(ctx->stream); ctx->tm.reset(); } }
extern "C" double mcb_timer_get(mcb_ctx *ctx, const char *name, uint64_t *count)
{
	if (count) *count = 0;
	if (!ctx || !name) return 0;
	auto it = ctx->tm.idx.find(name);
	if (it == ctx->tm.idx.end()) return 0;
	if (count) *count = ctx->tm.cnt[it->second];
	return ctx->tm.ms[it->second];
}
extern "C" size_t mcb_timers_dump(mcb_ctx *ctx, char *buf, size_t cap)
{
	std::string s;
	if (ctx) for (size_t i = 0; i < ctx->tm.names.size(); ++i) {
		char line[256];
		snprintf(line, sizeof line, "%s %.6f %llu\n", ctx->tm.names[i].c_str(), ctx->tm.ms[i], (unsigned long long)ctx->tm.cnt[i]);
		s += line;
	}
	if (buf && cap) { size_t n = s.size() < cap - 1 ? s.size() : cap - 1; memcpy(buf, s.data(), n); buf[n] = 0; }
	return s.size() + 1;
}
extern "C" uint64_t mcb_kernel_launches(mcb_ctx *ctx) { return ctx ? ctx->tm.launches : 0; }

// ---------------------------------------------------------------- host boundary helpers
extern "C" uint64_t mcb_hash64(uint64_t key, uint64_t mask) { return mcb_hash64_hd(key, mask); }

// mm_sketch_two (sketch.c:238-289) straight from ASCII; characters go through the same table as the reference
// (seq_nt4_table: anything that is not ACGT/acgt has code 4 and is ORed in unmasked, sketch.c:257-261).
extern "C" void mcb_sketch_two_host(const char *str, int len, int k, uint32_t rid, mcb_tuple *out)
{
	const uint64_t mask = (1ull << (2 * k)) - 1, shift1 = 2 * (k - 1);
	uint64_t f = 0, r = 0;
	mcb_tuple best; best.x = ~0ull; best.y = ~0ull;
	int l = 0;
	for (int i = 0; i < len; ++i) {
		unsigned char ch = (unsigned char)str[i];
		uint64_t c = (ch == 'A' || ch == 'a') ? 0 : (ch == 'C' || ch == 'c') ? 1 : (ch == 'G' || ch == 'g') ? 2 : (ch == 'T' || ch == 't') ? 3 : 4;
		f = (f << 2 | c) & mask;
		r = (r >> 2) | ((3ull ^ c) << shift1);
		if (f == r) continue;
		int z = f < r ? 0 : 1;
		if (++l >= k) {
			uint64_t h = mcb_hash64_hd(z ? r : f, mask);
			if (h < best.x) { best.x = h; best.y = (uint64_t)rid << 32 | (uint64_t)(uint32_t)i << 1 | (uint64_t)z; }
		}
	}
	*out = best;
}

// mm_sketch_lh_ori (sketch.c:116-165) for one string, host side (the per-contig call of the host contig merger).  Same
// formulation as the device kernel k_sketch_lh2 (mcb_stage1.cu): the w-slot ring is cut into blocks of w steps; when the ring
// pointer wraps, one backward pass records for every slot the rightmost minimum of the block's suffix, a running rightmost
// minimum covers the slots written since, and the window minimum is the smaller of the two (ties to the newer one) instead of
// the reference's rescan.  Both carry a "same hash elsewhere" bit, so the scan for identical k-mers runs only when one exists.
// Emission order and tie rules are the reference's (property-tested against the oracle in tests/test_device_algorithms_cpu.py).
extern "C" int64_t mcb_sketch_lh_host(const char *str, int len, int w, int k, uint32_t rid, mcb_tuple *out, int64_t cap)
{
	if (len <= 0 || w <= 0 || k <= 0 || k > 31) return 0;
	const uint64_t mask = (1ull << (2 * k)) - 1, shift1 = 2 * (k - 1), NONE = ~0ull;
	std::vector<uint64_t> rx((size_t)w, NONE), rp((size_t)w, NONE);
	std::vector<int> suf((size_t)w, w - 1);           // slot of the rightmost minimum of the previous block's slots [j, w)
	std::vector<char> suf_tie((size_t)w, 0);
	int64_t n_out = 0;
	auto emit = [&](uint64_t x, uint64_t p) { if (out && n_out < cap) { out[n_out].x = x; out[n_out].y = (uint64_t)rid << 32 | p; } ++n_out; };
	uint64_t fw = 0, rv = 0, mn_x = NONE, mn_p = NONE, run_x = NONE;
	int run_slot = 0; bool run_tie = false;
	int l = 0, bp = 0, mp = 0;
	for (int i = 0; i < len; ++i) {
		const unsigned char ch = (unsigned char)str[i];
		const unsigned c = (ch == 'A' || ch == 'a') ? 0u : (ch == 'C' || ch == 'c') ? 1u : (ch == 'G' || ch == 'g') ? 2u : (ch == 'T' || ch == 't') ? 3u : 4u;
		uint64_t ix = NONE, ip = NONE;
		if (c < 4) {
			fw = (fw << 2 | c) & mask;
			rv = (rv >> 2) | ((3ull ^ c) << shift1);
			if (fw == rv) continue;                                  // symmetric k-mer: no ring slot (sketch.c:135)
			const int z = fw < rv ? 0 : 1;
			if (++l >= k) { ix = mcb_hash64_hd(z ? rv : fw, mask); ip = (uint64_t)(uint32_t)i << 1 | (uint64_t)z; }
		} else l = 0;                                                // ambiguous base: run restarts, slot is still used (:139-140)
		rx[bp] = ix; rp[bp] = ip;
		if (ix <= run_x) { run_tie = ix == run_x; run_x = ix; run_slot = bp; }
		auto emit_equal = [&](int lo, int hi) { for (int j = lo; j < hi; ++j) if (mn_x == rx[j] && rp[j] != mn_p) emit(rx[j], rp[j]); };
		if (l == w + k - 1) { emit_equal(bp + 1, w); emit_equal(0, bp); }      // first full window (:141-146)
		if (ix <= mn_x) {
			if (l >= w + k) emit(mn_x, mn_p);
			mn_x = ix; mn_p = ip; mp = bp;
		} else if (bp == mp) {                                       // the minimum leaves the window (:150-161)
			if (l >= w + k - 1) emit(mn_x, mn_p);
			uint64_t nx = run_x; int ns = run_slot; bool tie = run_tie;
			if (bp + 1 < w) {
				const int ss = suf[bp + 1];
				if (rx[ss] < run_x) { nx = rx[ss]; ns = ss; tie = suf_tie[bp + 1] != 0; }
				else if (rx[ss] == run_x) tie = true;
			}
			mn_x = nx; mp = ns; mn_p = rp[ns];
			if (tie && l >= w + k - 1) { emit_equal(bp + 1, w); emit_equal(0, bp + 1); }
		}
		if (++bp == w) {
			bp = 0;
			uint64_t sx = rx[w - 1]; int ss = w - 1; char st = 0;
			suf[w - 1] = ss; suf_tie[w - 1] = 0;
			for (int j = w - 2; j >= 1; --j) {
				if (rx[j] < sx) { sx = rx[j]; ss = j; st = 0; }
				else if (rx[j] == sx) st = 1;
				suf[j] = ss; suf_tie[j] = st;
			}
			run_x = NONE; run_slot = 0; run_tie = false;
		}
	}
	if (mn_x != NONE) emit(mn_x, mn_p);
	return n_out;
}
