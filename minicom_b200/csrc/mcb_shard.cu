// mcb_shard.cu — the front end over the GPUs of one box: one context per GPU, NCCL (NVLink / NVSwitch) underneath the C-ABI.
//
// Partitioning:
//   reads     contiguous read-id ranges; a context sketches its own slice (mcb_for_reads* after mcb_shard_begin)
//   buckets   the 16384 minimizer buckets (x & 0x3FFF, kthread_reads.c:213) in contiguous ranges, owner = bucket * G >> 14
//   Stage 1   every round: tuples grouped by owner (one stable radix pass), then ONE grouped ncclSend/ncclRecv exchange that
//             carries each 16-byte tuple together with the 2-bit packed row of its read (WS words).  The owner of a bucket
//             therefore holds the rows of exactly the reads it builds consensus from, scattered into its packed table at their
//             global read ids, and every kernel of kt_for_bucket runs unchanged on it.  Nothing is all-gathered: a rank sends
//             (16 + 8 WS) bytes per local read per round instead of receiving 8 WS bytes for every read of the job.
//             Two 8-byte counters per rank and round (new seed contigs -> global contig ids, members -> the loop control of
//             kthread_bucket.c:607-622) travel in one tiny all-gather.
//   Stage 2   a rank realigns the singles IT produced in Stage 1 (their rows are local) against all contigs, see
//             mcb_shard_realign in mcb_stage2.cu; the only collectives are scalar guards of the dictionary bin sizes.
// Concatenating the ranks' Stage-1 results in rank order, round by round, and merging their claim lists by priority reproduces
// the single-GPU (= single-threaded reference) result bit for bit (tests/test_gpu_shard.py).
#include "mcb_common.cuh"
#include <nccl.h>
#include <algorithm>

#define MCB_NCCL(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { \
	mcb_set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, ncclGetErrorString(r_)); return MCB_ECUDA; } } while (0)

static inline ncclComm_t comm_of(mcb_ctx *ctx) { return (ncclComm_t)ctx->comm; }
static int check_ctx(mcb_ctx *ctx) { if (!ctx) { mcb_set_error("null context"); return MCB_EINVAL; } MCB_CUDA(cudaSetDevice(ctx->prm.device)); return MCB_OK; }

// implemented in mcb_stage1.cu / mcb_sort.cu
int mcb_bucket_begin(mcb_ctx *ctx);
int mcb_bucket_round_a_impl(mcb_ctx *ctx, int r, int is_last);
int mcb_bucket_round_b_impl(mcb_ctx *ctx, uint64_t cid_first);
int mcb_bucket_finish_impl(mcb_ctx *ctx, mcb_bucket_result *res, uint64_t *round_counts, int cap_rounds);
int mcb_elems_reserve(mcb_ctx *ctx, uint64_t n_tuples);      // room for n_tuples in both halves of the sort double buffer, current half preserved
int mcb_partition_by_owner(mcb_ctx *ctx, ulonglong2 *a, ulonglong2 *b, uint64_t n, int n_ranks, uint64_t *counts /* host [n_ranks+1] */);

// ---------------------------------------------------------------- communicator
extern "C" int mcb_shard_unique_id(void *id128)
{
	if (!id128) { mcb_set_error("mcb_shard_unique_id: null argument"); return MCB_EINVAL; }
	static_assert(sizeof(ncclUniqueId) == MCB_NCCL_ID_BYTES, "ncclUniqueId size");
	ncclUniqueId id;
	MCB_NCCL(ncclGetUniqueId(&id));
	memcpy(id128, &id, sizeof id);
	return MCB_OK;
}

static void shard_reset(mcb_ctx *ctx, int rank, int n_ranks)
{
	ctx->shard_rank = rank; ctx->shard_n = n_ranks;
	ctx->reads_loaded = false; ctx->bucket_done = false; ctx->bs.active = false; ctx->cix.valid = false;
}

extern "C" int mcb_shard_init(mcb_ctx *ctx, const void *id128, int rank, int n_ranks)
{
	MCB_TRY(check_ctx(ctx));
	if (!id128 || n_ranks < 1 || n_ranks > 255 || rank < 0 || rank >= n_ranks) { mcb_set_error("mcb_shard_init: bad arguments"); return MCB_EINVAL; }
	if (ctx->comm) { mcb_set_error("mcb_shard_init: context already has a communicator"); return MCB_ESTATE; }
	ncclUniqueId id; memcpy(&id, id128, sizeof id);
	ncclComm_t c;
	MCB_NCCL(ncclCommInitRank(&c, n_ranks, id, rank));
	ctx->comm = c; ctx->own_comm = true;
	shard_reset(ctx, rank, n_ranks);
	return MCB_OK;
}

extern "C" int mcb_shard_attach(mcb_ctx *ctx, void *nccl_comm, int rank, int n_ranks)
{
	MCB_TRY(check_ctx(ctx));
	if (!nccl_comm || n_ranks < 1 || n_ranks > 255 || rank < 0 || rank >= n_ranks) { mcb_set_error("mcb_shard_attach: bad arguments"); return MCB_EINVAL; }
	if (ctx->comm) { mcb_set_error("mcb_shard_attach: context already has a communicator"); return MCB_ESTATE; }
	ctx->comm = nccl_comm; ctx->own_comm = false;
	shard_reset(ctx, rank, n_ranks);
	return MCB_OK;
}

void mcb_shard_release(mcb_ctx *ctx)          // from mcb_destroy
{
	if (ctx->comm && ctx->own_comm) ncclCommDestroy(comm_of(ctx));
	ctx->comm = nullptr;
}

extern "C" int mcb_shard_begin(mcb_ctx *ctx, uint64_t n_total, uint64_t rid_base)
{
	MCB_TRY(check_ctx(ctx));
	if (ctx->shard_n > 1 && !ctx->comm) { mcb_set_error("mcb_shard_begin: no communicator (mcb_shard_init / mcb_shard_attach)"); return MCB_ESTATE; }
	if (rid_base > n_total) { mcb_set_error("mcb_shard_begin: bad arguments"); return MCB_EINVAL; }
	if (n_total >= (1ull << 31)) { mcb_set_error("too many reads (rid is a signed 32-bit int in the reference, kthread_bucket.c:48)"); return MCB_EINVAL; }
	ctx->n_reads = n_total; ctx->rid_base = rid_base;
	ctx->reads_loaded = false; ctx->bucket_done = false; ctx->bs.active = false;
	return MCB_OK;
}

// ---------------------------------------------------------------- small collectives
int mcb_coll_allreduce_sum_u64(mcb_ctx *ctx, unsigned long long *d, size_t n)
{
	if (ctx->shard_n <= 1) return MCB_OK;
	MCB_NCCL(ncclAllReduce(d, d, n, ncclUint64, ncclSum, comm_of(ctx), ctx->stream));
	return MCB_OK;
}
int mcb_coll_allreduce_sum_u32(mcb_ctx *ctx, uint32_t *d, size_t n)
{
	if (ctx->shard_n <= 1) return MCB_OK;
	MCB_NCCL(ncclAllReduce(d, d, n, ncclUint32, ncclSum, comm_of(ctx), ctx->stream));
	return MCB_OK;
}

int mcb_coll_allgather_inplace_u64(mcb_ctx *ctx, uint64_t *d_buf, size_t chunk_words)
{
	if (ctx->shard_n <= 1 || !chunk_words) return MCB_OK;
	MCB_NCCL(ncclAllGather(d_buf + (size_t)ctx->shard_rank * chunk_words, d_buf, chunk_words, ncclUint64, comm_of(ctx), ctx->stream));
	return MCB_OK;
}

// every rank contributes `n` 8-byte words, all ranks get the [n_ranks][n] table on the host
static int allgather_words(mcb_ctx *ctx, const unsigned long long *h_mine, int n, unsigned long long *h_all)
{
	const int G = ctx->shard_n;
	MCB_TRY(ctx->d_coll.ensure((size_t)(G + 1) * n * 8 + 64));
	unsigned long long *d_all = ctx->d_coll.as<unsigned long long>(), *d_mine = d_all + (size_t)G * n;
	MCB_CUDA(cudaMemcpyAsync(d_mine, h_mine, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
	{
		McbSpan sp(ctx->tm, "nccl:counts");
		MCB_NCCL(ncclAllGather(d_mine, d_all, (size_t)n, ncclUint64, comm_of(ctx), ctx->stream));
	}
	MCB_CUDA(cudaMemcpyAsync(h_all, d_all, (size_t)G * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	return MCB_OK;
}

int mcb_coll_allgatherv(mcb_ctx *ctx, const void *d_send, uint64_t bytes, std::vector<unsigned long long> &host_out)
{
	const int G = ctx->shard_n, me = ctx->shard_rank;
	std::vector<unsigned long long> sizes((size_t)G);
	unsigned long long mine = bytes;
	MCB_TRY(allgather_words(ctx, &mine, 1, sizes.data()));
	uint64_t total = 0; std::vector<uint64_t> off((size_t)G + 1, 0);
	for (int q = 0; q < G; ++q) { off[q] = total; total += sizes[q]; }
	off[G] = total;
	host_out.assign(total / 8, 0ull);
	if (!total) return MCB_OK;
	DBuf tmp;
	struct Rel { DBuf &b; ~Rel() { b.release(); } } rel{tmp};
	MCB_TRY(tmp.ensure(total + 16));
	if (bytes) MCB_CUDA(cudaMemcpyAsync(tmp.as<char>() + off[me], d_send, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
	MCB_NCCL(ncclGroupStart());
	for (int q = 0; q < G; ++q)
		if (sizes[q]) MCB_NCCL(ncclBroadcast(tmp.as<char>() + off[q], tmp.as<char>() + off[q], sizes[q], ncclUint8, q, comm_of(ctx), ctx->stream));
	MCB_NCCL(ncclGroupEnd());
	MCB_CUDA(cudaMemcpyAsync(host_out.data(), tmp.p, total, cudaMemcpyDeviceToHost, ctx->stream));
	MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	return MCB_OK;
}

// ---------------------------------------------------------------- Stage 1 exchange
// rows travel in the order of their tuples
__global__ void k_shard_gather_rows(const ulonglong2 *__restrict__ el, uint64_t n, const uint64_t *__restrict__ packed, int WS, uint64_t *__restrict__ rows)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n * (uint64_t)WS) return;
	const uint64_t t = i / WS; const int w = (int)(i - t * WS);
	rows[i] = packed[(uint64_t)mcb_k2_rid(el[t].y) * WS + w];
}
__global__ void k_shard_scatter_rows(const ulonglong2 *__restrict__ el, uint64_t n, uint64_t *__restrict__ packed, int WS, const uint64_t *__restrict__ rows)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n * (uint64_t)WS) return;
	const uint64_t t = i / WS; const int w = (int)(i - t * WS);
	packed[(uint64_t)mcb_k2_rid(el[t].y) * WS + w] = rows[i];
}

// one round's repartition: the tuples of this rank (sketched from its slice in round 1, re-sketched rejects later) go to the owners
// of their buckets, each with the packed row of its read
static int shard_exchange(mcb_ctx *ctx)
{
	McbBucketState &bs = ctx->bs;
	const int G = ctx->shard_n, me = ctx->shard_rank, WS = ctx->WS;
	std::vector<uint64_t> cnt((size_t)G + 1, 0);
	{
		McbSpan sp(ctx->tm, "for_bucket");
		MCB_TRY(mcb_partition_by_owner(ctx, bs.cur, bs.alt, bs.n_in, G, cnt.data()));
	}
	if (bs.n_in > 1) std::swap(bs.cur, bs.alt);
	std::vector<unsigned long long> mine((size_t)G), all((size_t)G * G);
	uint64_t n_send = 0;
	for (int q = 0; q < G; ++q) { mine[q] = cnt[q]; n_send += cnt[q]; }
	MCB_TRY(allgather_words(ctx, mine.data(), G, all.data()));
	std::vector<uint64_t> so((size_t)G + 1, 0), ro((size_t)G + 1, 0);
	for (int q = 0; q < G; ++q) { so[q + 1] = so[q] + all[(size_t)me * G + q]; ro[q + 1] = ro[q] + all[(size_t)q * G + me]; }
	const uint64_t n_recv = ro[G];
	MCB_TRY(mcb_elems_reserve(ctx, std::max(n_recv, bs.n_in)));
	MCB_TRY(ctx->d_rows_send.ensure(n_send * WS * 8 + 16)); MCB_TRY(ctx->d_rows_recv.ensure(n_recv * WS * 8 + 16));
	uint64_t *rs = ctx->d_rows_send.as<uint64_t>(), *rr = ctx->d_rows_recv.as<uint64_t>();
	{
		McbSpan sp(ctx->tm, "for_bucket");
		if (n_send) MCB_LAUNCH(ctx, "shard_gather_rows", k_shard_gather_rows, mcb_grid_for(n_send * WS, 256), 256, 0, bs.cur, n_send, ctx->d_packed.as<uint64_t>(), WS, rs);
	}
	{
		McbSpan sp(ctx->tm, "nccl:tuples+rows");
		MCB_NCCL(ncclGroupStart());
		for (int q = 0; q < G; ++q) {
			const uint64_t ns = so[q + 1] - so[q], nr = ro[q + 1] - ro[q];
			if (ns) {
				MCB_NCCL(ncclSend(bs.cur + so[q], ns * 16, ncclUint8, q, comm_of(ctx), ctx->stream));
				MCB_NCCL(ncclSend(rs + so[q] * WS, ns * WS * 8, ncclUint8, q, comm_of(ctx), ctx->stream));
			}
			if (nr) {
				MCB_NCCL(ncclRecv(bs.alt + ro[q], nr * 16, ncclUint8, q, comm_of(ctx), ctx->stream));
				MCB_NCCL(ncclRecv(rr + ro[q] * WS, nr * WS * 8, ncclUint8, q, comm_of(ctx), ctx->stream));
			}
		}
		MCB_NCCL(ncclGroupEnd());
	}
	{
		McbSpan sp(ctx->tm, "for_bucket");
		if (n_recv) MCB_LAUNCH(ctx, "shard_scatter_rows", k_shard_scatter_rows, mcb_grid_for(n_recv * WS, 256), 256, 0, bs.alt, n_recv, ctx->d_packed.as<uint64_t>(), WS, rr);
	}
	std::swap(bs.cur, bs.alt);                  // the receive buffer is the input of this round
	bs.n_in = bs.n_valid = n_recv;
	if (ctx->tm.enabled) { int id = ctx->tm.id("nccl_bytes_sent"); ctx->tm.ms[id] += (double)((n_send - (so[me + 1] - so[me])) * (16 + (uint64_t)WS * 8)); ctx->tm.cnt[id] += 1; }
	return MCB_OK;
}

// the side table of reads that contained N (rare): Stage 2 needs the original characters of ANY single for the near-poly-A/T test
// (bbhashdict.c:148-154), so every rank gets the whole table, ascending read id (= rank order)
static int shard_gather_nreads(mcb_ctx *ctx)
{
	const int WS = ctx->WS;
	std::vector<unsigned long long> rid_all, mask_all;
	// the local table is on the device (d_nread_rid u32[n], d_nread_mask u64[n][WS]); pad the read ids to 8-byte words for the gather
	const uint64_t n = ctx->n_nreads;
	DBuf wide;
	struct Rel { DBuf &b; ~Rel() { b.release(); } } rel{wide};
	std::vector<unsigned long long> h_wide(n);
	const uint32_t *hr = ctx->h_nrid.as<uint32_t>();
	for (uint64_t i = 0; i < n; ++i) h_wide[i] = hr[i];
	MCB_TRY(wide.ensure(n * 8 + 16));
	if (n) MCB_CUDA(cudaMemcpyAsync(wide.p, h_wide.data(), n * 8, cudaMemcpyHostToDevice, ctx->stream));
	MCB_TRY(mcb_coll_allgatherv(ctx, wide.p, n * 8, rid_all));
	MCB_TRY(mcb_coll_allgatherv(ctx, ctx->d_nread_mask.p, n * WS * 8, mask_all));
	const uint64_t nt = rid_all.size();
	if (mask_all.size() != nt * WS) { mcb_set_error("internal: N side table gather is inconsistent"); return MCB_EINVAL; }
	std::vector<uint32_t> rid32(nt);
	for (uint64_t i = 0; i < nt; ++i) rid32[i] = (uint32_t)rid_all[i];
	MCB_TRY(ctx->d_nread_rid.ensure(nt * 4 + 16)); MCB_TRY(ctx->d_nread_mask.ensure(nt * WS * 8 + 16));
	if (nt) {
		MCB_CUDA(cudaMemcpyAsync(ctx->d_nread_rid.p, rid32.data(), nt * 4, cudaMemcpyHostToDevice, ctx->stream));
		MCB_CUDA(cudaMemcpyAsync(ctx->d_nread_mask.p, mask_all.data(), nt * WS * 8, cudaMemcpyHostToDevice, ctx->stream));
		MCB_CUDA(cudaStreamSynchronize(ctx->stream));
	}
	ctx->n_nreads = nt;
	return MCB_OK;
}

// kt_for_bucket (kthread_bucket.c:562-629) over the whole job; this rank returns the seed contigs / singles of the buckets it owns
// in its own round order, and round_counts[4 r .. 4 r + 3] = contigs, members, consensus bytes, singles that round r contributed
// (the job's lists are the ranks' contributions concatenated in rank order, round by round).  Contig ids inside the index tuples
// (res->mi) are global.  Collective.
extern "C" int mcb_shard_for_bucket(mcb_ctx *ctx, mcb_bucket_result *res, uint64_t *round_counts, int cap_rounds)
{
	MCB_TRY(check_ctx(ctx));
	if (!res) { mcb_set_error("mcb_shard_for_bucket: null result"); return MCB_EINVAL; }
	if (ctx->shard_n <= 1) {       // a one-rank "job": plain kt_for_bucket, counts included
		MCB_TRY(mcb_bucket_begin(ctx));
		mcb_round_control rc; mcb_round_control_init(&rc);
		for (;;) {
			const int is_last = mcb_round_control_begin(&rc, ctx->prm.k, ctx->prm.max_rounds);
			MCB_TRY(mcb_bucket_round_a_impl(ctx, rc.round, is_last));
			MCB_TRY(mcb_bucket_round_b_impl(ctx, ctx->bs.tot_cl));
			if (mcb_round_control_end(&rc, ctx->bs.tot_mem)) break;
		}
		return mcb_bucket_finish_impl(ctx, res, round_counts, cap_rounds);
	}
	if (!ctx->comm) { mcb_set_error("mcb_shard_for_bucket: no communicator"); return MCB_ESTATE; }
	const int G = ctx->shard_n, me = ctx->shard_rank;
	MCB_TRY(mcb_bucket_begin(ctx));
	MCB_TRY(shard_gather_nreads(ctx));
	mcb_round_control rc; mcb_round_control_init(&rc);
	uint64_t tot_cl = 0, members = 0;
	std::vector<unsigned long long> all((size_t)2 * G);
	for (;;) {
		const int is_last = mcb_round_control_begin(&rc, ctx->prm.k, ctx->prm.max_rounds);
		MCB_TRY(shard_exchange(ctx));
		MCB_TRY(mcb_bucket_round_a_impl(ctx, rc.round, is_last));
		unsigned long long mine[2] = { ctx->bs.n_cl_new, ctx->bs.n_mem_new };
		MCB_TRY(allgather_words(ctx, mine, 2, all.data()));
		uint64_t before = 0, cl_round = 0, mem_round = 0;
		for (int q = 0; q < G; ++q) { if (q < me) before += all[2 * q]; cl_round += all[2 * q]; mem_round += all[2 * q + 1]; }
		MCB_TRY(mcb_bucket_round_b_impl(ctx, tot_cl + before));          // global id of this rank's first new contig in the single-GPU order
		tot_cl += cl_round; members += mem_round;
		if (mcb_round_control_end(&rc, members)) break;
	}
	return mcb_bucket_finish_impl(ctx, res, round_counts, cap_rounds);
}
