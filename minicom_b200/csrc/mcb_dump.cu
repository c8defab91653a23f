// mcb_dump.cu — N2, first slice: the per-read diff encoding of the dump stage on the device.
//
// print_encode (kthread_dump.c:33-236; the ORDER, default and _PE variants share this loop, :66-118) takes every member of a
// contig — the read as it was before N replacement, reverse-complemented when dir — and writes one line of dif_char.txt: runs of
// >= 2 characters equal to the consensus as their decimal length, a run of 1 copied, every mismatching character copied, the
// trailing run dropped, "0" when nothing differs.  That is the only part of the dump stage that touches every base; it is
// independent per member, so: one thread per member, a counting pass, a scan, an emitting pass.  The reads are the 2-bit rows
// and the N side table mcb_for_reads* left on the device; the contigs are the ones Stage 2 holds (or are passed in).
// What the dump stage does around it (member sort, position deltas, direction bits, packed consensus, the ten output files)
// stays the reference's host code: this entry point is not yet bound by the shim.
#include "mcb_common.cuh"

struct DumpIn {
	const uint64_t *members, *moff;      // y = rid<<32 | pos<<1 | dir;  member offsets per contig [n_contigs+1]
	const char *refs; const uint64_t *roff;
	const uint64_t *packed; const uint32_t *nrid; const uint64_t *nmask;
	uint64_t n_members, n_contigs, n_nreads, n_reads;
	int L, WS;
};

template <bool EMIT>
__global__ void k_dump_encode(DumpIn in, uint32_t *__restrict__ cnt, const uint32_t *__restrict__ off, char *__restrict__ out, unsigned long long *__restrict__ err)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= in.n_members) return;
	uint64_t lo = 0, hi = in.n_contigs;       // last contig whose first member is <= i
	while (hi - lo > 1) { const uint64_t mid = (lo + hi) >> 1; if (in.moff[mid] <= i) lo = mid; else hi = mid; }
	const uint64_t y = in.members[i];
	const uint32_t rid = (uint32_t)(y >> 32);
	const int pos = (int)((uint32_t)y >> 1), dir = (int)(y & 1), L = in.L;
	const uint64_t rb = in.roff[lo], rlen = in.roff[lo + 1] - rb;
	if (rid >= in.n_reads || (uint64_t)pos + (uint64_t)L > rlen) { if (!EMIT) { cnt[i] = 0; atomicAdd(err, 1ull); } return; }
	const char *ref = in.refs + rb + pos;
	const uint64_t *row = in.packed + (uint64_t)rid * in.WS;
	const uint64_t *nm = nullptr;             // the read's N mask, if it has one (side table sorted by read id)
	{
		uint64_t a = 0, b = in.n_nreads;
		while (a < b) { const uint64_t mid = (a + b) >> 1; const uint32_t r = in.nrid[mid]; if (r < rid) a = mid + 1; else if (r > rid) b = mid; else { nm = in.nmask + mid * in.WS; break; } }
	}
	char *o = EMIT ? out + off[i] : nullptr;
	int n = 0, eq = 0;
	char prev = 0;
	for (int tj = 0; tj < L; ++tj) {
		const int p = dir ? L - 1 - tj : tj;
		const unsigned code = mcb_base_at(row, p);
		const bool isn = nm && ((nm[p >> 5] >> (2 * (p & 31))) & 1ull);
		const char ch = isn ? 'N' : "ACGT"[dir ? 3u - code : code];
		if (ref[tj] != ch) {
			if (eq > 1) {
				const int d2 = eq / 100, d1 = (eq / 10) % 10, d0 = eq % 10;
				if (eq >= 100) { if (EMIT) o[n] = (char)('0' + d2); ++n; }
				if (eq >= 10) { if (EMIT) o[n] = (char)('0' + d1); ++n; }
				if (EMIT) o[n] = (char)('0' + d0); ++n;
			} else if (eq == 1) { if (EMIT) o[n] = prev; ++n; }
			eq = 0;
			if (EMIT) o[n] = ch; ++n;
		} else ++eq;
		prev = ch;
	}
	if (n == 0) { if (EMIT) o[0] = '0'; n = 1; }
	if (!EMIT) cnt[i] = (uint32_t)n;
}

extern "C" int mcb_dump_encode(mcb_ctx *ctx, const uint64_t *members, const uint64_t *member_off, const char *refs, const uint64_t *ref_off, uint64_t n_contigs,
                               mcb_encode_result *res)
{
	if (!ctx || !res || !member_off || (n_contigs && !members && member_off[n_contigs])) { mcb_set_error("mcb_dump_encode: null argument"); return MCB_EINVAL; }
	MCB_CUDA(cudaSetDevice(ctx->prm.device));
	if (!ctx->reads_loaded) { mcb_set_error("mcb_dump_encode: no reads loaded (mcb_for_reads first)"); return MCB_ESTATE; }
	if ((refs == nullptr) != (ref_off == nullptr)) { mcb_set_error("mcb_dump_encode: refs and ref_off go together"); return MCB_EINVAL; }
	memset(res, 0, sizeof *res);
	const uint64_t nm = n_contigs ? member_off[n_contigs] : 0;
	McbContigIndex &cx = ctx->cix;
	DumpIn in;
	memset(&in, 0, sizeof in);
	in.n_members = nm; in.n_contigs = n_contigs; in.n_nreads = ctx->n_nreads; in.n_reads = ctx->n_reads; in.L = ctx->L; in.WS = ctx->WS;
	in.packed = ctx->d_packed.as<uint64_t>(); in.nrid = ctx->d_nread_rid.as<uint32_t>(); in.nmask = ctx->d_nread_mask.as<uint64_t>();
	MCB_TRY(ctx->d_scr[3].ensure(nm * 8 + 16)); MCB_TRY(ctx->d_scr[4].ensure((n_contigs + 2) * 8)); MCB_TRY(ctx->d_scr[5].ensure((nm + 2) * 4));
	MCB_TRY(ctx->h_in0.ensure((nm + 2) * 8));
	MCB_TRY(ctx->d_counters.ensure(64 * 8));
	unsigned long long *dc = ctx->d_counters.as<unsigned long long>();
	{
		McbSpan sp(ctx->tm, "h2d");
		if (nm) MCB_TRY(mcb_h2d(ctx, ctx->d_scr[3].p, members, nm * 8, 4));
		MCB_TRY(mcb_h2d(ctx, ctx->d_scr[4].p, member_off, (n_contigs + 1) * 8, 1));
		if (refs) {
			const uint64_t rbytes = ref_off[n_contigs];
			MCB_TRY(ctx->d_scr[6].ensure(rbytes + 16)); MCB_TRY(ctx->d_scr[7].ensure((n_contigs + 2) * 8));
			if (rbytes) MCB_TRY(mcb_h2d(ctx, ctx->d_scr[6].p, refs, rbytes, 4));
			MCB_TRY(mcb_h2d(ctx, ctx->d_scr[7].p, ref_off, (n_contigs + 1) * 8, 1));
			in.refs = ctx->d_scr[6].as<char>(); in.roff = ctx->d_scr[7].as<uint64_t>();
		} else {
			if (!cx.valid || cx.n_contigs != n_contigs) { mcb_set_error("mcb_dump_encode: no contigs on the device for this call (%llu given, %llu held)", (unsigned long long)n_contigs, (unsigned long long)(cx.valid ? cx.n_contigs : 0)); return MCB_ESTATE; }
			in.refs = cx.refs.as<char>(); in.roff = cx.roff.as<uint64_t>();
		}
	}
	in.members = ctx->d_scr[3].as<uint64_t>(); in.moff = ctx->d_scr[4].as<uint64_t>();
	uint32_t *cnt = ctx->d_scr[5].as<uint32_t>();
	uint64_t total = 0;
	if (nm) {
		McbSpan sp(ctx->tm, "dump_encode");
		MCB_CUDA(cudaMemsetAsync(&dc[CT_ERR], 0, 8, ctx->stream));
		MCB_LAUNCH(ctx, "dump_encode_count", k_dump_encode<false>, mcb_grid_for(nm, 128), 128, 0, in, cnt, (const uint32_t*)nullptr, (char*)nullptr, &dc[CT_ERR]);
		MCB_TRY(mcb_exclusive_scan_u32(ctx, cnt, nm, (uint64_t*)&dc[CT_SCRATCH_IDX]));
	}
	MCB_TRY(ctx->h_counters.ensure(64 * 8));
	MCB_CUDA(cudaMemcpyAsync(ctx->h_counters.p, ctx->d_counters.p, 64 * 8, cudaMemcpyDeviceToHost, ctx->stream));
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	if (nm) {
		const unsigned long long *hc = ctx->h_counters.as<unsigned long long>();
		if (hc[CT_ERR]) { mcb_set_error("mcb_dump_encode: %llu members with a read id or a position outside their contig", hc[CT_ERR]); return MCB_EINPUT; }
		total = hc[CT_SCRATCH_IDX];
		if (total >= 0xFFFFFFFFull) { mcb_set_error("mcb_dump_encode: more than 4 GB of encodings in one call"); return MCB_EINVAL; }
	}
	MCB_TRY(ctx->d_scr[8].ensure(total + 16)); MCB_TRY(ctx->h_in1.ensure(total + 16)); MCB_TRY(ctx->h_in2.ensure((nm + 2) * 4));
	if (nm) {
		{
			McbSpan sp(ctx->tm, "dump_encode");
			MCB_LAUNCH(ctx, "dump_encode_emit", k_dump_encode<true>, mcb_grid_for(nm, 128), 128, 0, in, (uint32_t*)nullptr, (const uint32_t*)cnt, ctx->d_scr[8].as<char>(), &dc[CT_ERR]);
		}
		McbSpan sp(ctx->tm, "d2h");
		MCB_CUDA(cudaMemcpyAsync(ctx->h_in2.p, cnt, nm * 4, cudaMemcpyDeviceToHost, ctx->stream));
		if (total) MCB_CUDA(cudaMemcpyAsync(ctx->h_in1.p, ctx->d_scr[8].p, total, cudaMemcpyDeviceToHost, ctx->stream));
	}
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->tm.collect();
	uint64_t *eo = ctx->h_in0.as<uint64_t>();
	const uint32_t *o32 = ctx->h_in2.as<uint32_t>();
	for (uint64_t i = 0; i < nm; ++i) eo[i] = o32[i];
	eo[nm] = total;
	res->n_members = nm; res->enc_off = eo; res->enc = ctx->h_in1.as<char>(); res->n_bytes = total;
	return MCB_OK;
}
