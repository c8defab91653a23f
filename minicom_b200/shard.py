"""Sharding of the front end across the GPUs of one box: one process per GPU, `torch.distributed` (NCCL over NVLink) for the
plumbing, libminicom_b200.so for the device work (include/minicom_b200.h, "sharding across the GPUs of one box").

Partitioning rules:
  reads      contiguous read-id ranges, rank r owns [r*cap, min((r+1)*cap, N)), cap = ceil(N / G)
  buckets    the 16384 minimizer buckets (x & 0x3FFF, kthread_reads.c:213) in contiguous ranges, owner = bucket*G >> 14
  contigs    contiguous ranges of the contig list, cut so that every rank gets about the same number of bases

Exchanges:
  Stage 1    all-gather of the 2-bit packed reads (any rank builds the consensus of its groups from any read);
             per round an all-to-all of the 16-byte (minimizer, read, position, strand) tuples by bucket owner;
             an all-gather of two counters per round (new seed contigs for the global contig ids, members for the loop control)
  Stage 2    contigs and singles on every rank; the contig lt-mer table partitioned by hash range (each rank probes only the
             lt-mers it owns, so probes per rank stay constant); one all-reduce(MIN) of the per-single claim priorities per
             threshold round; claims emitted by window range
Concatenating the ranks' results in rank order — round by round for Stage 1 — reproduces the single-GPU (= single-threaded
reference) order exactly; `merge_stage1` / `merge_claims` do that on one host for the callers that stay on one host (contig merge).

The pure partition/merge arithmetic has no CUDA dependency and is tested on CPU with the gloo backend (tests/test_shard_cpu.py).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

NB = 1 << 14          # minimizer buckets (reads->b = 14, minicommain.c:175)


# ---------------------------------------------------------------- partition arithmetic (pure)
def rid_range(n_total: int, rank: int, n_ranks: int):
    cap = (n_total + n_ranks - 1) // n_ranks
    lo = min(rank * cap, n_total)
    return lo, min(lo + cap, n_total)


def owner_of_bucket(bucket, n_ranks: int):
    return (np.asarray(bucket, dtype=np.int64) * n_ranks) >> 14


def bucket_range(rank: int, n_ranks: int):
    """[b0, b1): the buckets rank owns (owner_of_bucket is monotone)."""
    b0 = (rank * NB + n_ranks - 1) // n_ranks
    b1 = ((rank + 1) * NB + n_ranks - 1) // n_ranks
    return b0, b1


def contig_partition(ref_off: np.ndarray, n_ranks: int, readlen: int):
    """Cut the contig list into n_ranks contiguous ranges of about equal bases.
    Returns (cuts[n_ranks+1] contig indices, window_base[n_ranks]) with windows = max(0, len-L+1) per contig."""
    ref_off = np.asarray(ref_off, dtype=np.uint64).astype(np.int64)
    n = len(ref_off) - 1
    total = int(ref_off[-1]) if n > 0 else 0
    cuts = [0]
    for r in range(1, n_ranks):
        cuts.append(int(np.searchsorted(ref_off, total * r // n_ranks, side="left")) if n else 0)
    cuts.append(n)
    cuts = np.minimum.accumulate(np.array(cuts[::-1]))[::-1]          # monotone, ends at n
    cuts = np.maximum.accumulate(np.minimum(cuts, n))
    lens = np.diff(ref_off)
    win = np.where(lens >= readlen, lens - readlen + 1, 0)
    wcum = np.concatenate([[0], np.cumsum(win)])
    return cuts.astype(np.int64), wcum[cuts[:-1]].astype(np.int64)


def exchange_plan(send_counts: np.ndarray, recv_counts: np.ndarray):
    """Row offsets of an all-to-all with these split sizes."""
    so = np.concatenate([[0], np.cumsum(send_counts)]).astype(np.int64)
    ro = np.concatenate([[0], np.cumsum(recv_counts)]).astype(np.int64)
    return so, ro


@dataclass
class Stage1Part:
    """One rank's share of kt_for_bucket, in its own round order, plus what every round contributed."""
    cl_n: np.ndarray
    cl_a: np.ndarray
    cl_ref: np.ndarray
    cl_reflen: np.ndarray
    sg: np.ndarray
    mi: np.ndarray           # (n_clusters, m, 2)
    mi_cnt: np.ndarray
    rounds: np.ndarray       # (n_rounds, 4): contigs, members, consensus bytes, singles


def merge_stage1(parts: list[Stage1Part]) -> Stage1Part:
    """Single-GPU order = for every round, the ranks in order (bucket ranges ascend with the rank)."""
    n_rounds = max(len(p.rounds) for p in parts)
    cur = [dict(c=0, a=0, r=0, s=0) for _ in parts]
    out = dict(cl_n=[], cl_a=[], cl_ref=[], cl_reflen=[], sg=[], mi=[], mi_cnt=[])
    rounds = np.zeros((n_rounds, 4), dtype=np.uint64)
    for rd in range(n_rounds):
        for p, c in zip(parts, cur):
            if rd >= len(p.rounds):
                continue
            ncl, nmem, nref, nsg = (int(x) for x in p.rounds[rd])
            out["cl_n"].append(p.cl_n[c["c"]:c["c"] + ncl])
            out["cl_reflen"].append(p.cl_reflen[c["c"]:c["c"] + ncl])
            out["mi"].append(p.mi[c["c"]:c["c"] + ncl])
            out["mi_cnt"].append(p.mi_cnt[c["c"]:c["c"] + ncl])
            out["cl_a"].append(p.cl_a[c["a"]:c["a"] + nmem])
            out["cl_ref"].append(p.cl_ref[c["r"]:c["r"] + nref])
            out["sg"].append(p.sg[c["s"]:c["s"] + nsg])
            c["c"] += ncl; c["a"] += nmem; c["r"] += nref; c["s"] += nsg
            rounds[rd] += np.array([ncl, nmem, nref, nsg], dtype=np.uint64)
    cat = lambda k, like: np.concatenate(out[k]) if out[k] else like[:0]      # noqa: E731
    p0 = parts[0]
    return Stage1Part(cat("cl_n", p0.cl_n), cat("cl_a", p0.cl_a), cat("cl_ref", p0.cl_ref), cat("cl_reflen", p0.cl_reflen), cat("sg", p0.sg),
                      cat("mi", p0.mi), cat("mi_cnt", p0.mi_cnt), rounds)


def merge_claims(parts: list[tuple[np.ndarray, np.ndarray, np.ndarray]], contig_cuts: np.ndarray | None = None):
    """Per-rank (claim_contig, claim_sg, claim_y) -> global lists in the reference's append order (rank order = window order).
    contig_cuts: first contig of every rank when the contig indices are rank-local (contig-range sharding); None when they are
    global already (key sharding)."""
    cc = [p[0].astype(np.int64) + (int(contig_cuts[r]) if contig_cuts is not None else 0) for r, p in enumerate(parts)]
    return np.concatenate(cc).astype(np.uint32), np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts])


# ---------------------------------------------------------------- device plumbing
class _DevMem:
    """Zero-copy torch view of library-owned device memory."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def dev_view(ptr: int, nbytes: int, device):
    import torch
    if nbytes == 0:
        return torch.empty(0, dtype=torch.uint8, device=device)
    return torch.as_tensor(_DevMem(ptr, nbytes), device=device)


def all_to_all_rows(dist, send, send_counts, device):
    """All-to-all of row blocks: `send` is (n, w) with the rows for rank q at [so[q], so[q+1]).  Returns (recv_counts, recv).
    Works on CPU tensors (gloo) and CUDA tensors (NCCL)."""
    import torch
    world = dist.get_world_size()
    sc = torch.as_tensor(np.asarray(send_counts, dtype=np.int64), device=device)
    rc = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(rc, sc)
    recv_counts = rc.cpu().numpy()
    recv = torch.empty((int(recv_counts.sum()),) + tuple(send.shape[1:]), dtype=send.dtype, device=device)
    dist.all_to_all_single(recv, send, output_split_sizes=[int(x) for x in recv_counts], input_split_sizes=[int(x) for x in send_counts])
    return recv_counts, recv


class ShardedFrontEnd:
    """Drives one rank of the sharded front end.  `ctx` is an api.Context on this rank's GPU."""

    def __init__(self, ctx, dist, device):
        import torch
        self.ctx, self.dist, self.device, self.torch = ctx, dist, device, torch
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.bytes_exchanged = 0
        self.time_collectives = False
        self._events = []

    def _coll(self, fn, *a, **kw):
        """run a collective; when asked, bracket it with CUDA events on the stream it is enqueued on"""
        if not self.time_collectives:
            return fn(*a, **kw)
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **kw)
        e1.record()
        self._events.append((e0, e1))
        return r

    def collective_ms(self, reset=True):
        self.torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in self._events)
        if reset:
            self._events = []
        return ms

    # ---- Stage 1: kt_for_reads + kt_for_bucket over the whole read set
    def stage1(self, rows_local, n_total: int, device_resident: bool = False):
        """rows_local: this rank's reads, (n_local, L) uint8 (numpy, or a CUDA tensor when device_resident).
        Returns (ReadsResult of the slice, Stage1Part of this rank)."""
        from . import api
        torch, dist, ctx = self.torch, self.dist, self.ctx
        lo, hi = rid_range(n_total, self.rank, self.world)
        ctx.shard_begin(self.rank, self.world, n_total, lo)
        rr = ctx.for_reads_device(rows_local.data_ptr(), hi - lo) if device_resident else ctx.for_reads(rows_local)
        # all-gather of the packed reads (equal chunks of `cap` rows; the last one overhangs into the slack the library allocated)
        ptr, rb = ctx.shard_packed()
        cap = (n_total + self.world - 1) // self.world
        whole = dev_view(ptr, self.world * cap * rb, self.device)
        mine = whole[self.rank * cap * rb:(self.rank + 1) * cap * rb].clone()
        self._coll(dist.all_gather_into_tensor, whole, mine)
        self.bytes_exchanged += whole.numel()
        # the (rare) reads that contained N: every rank needs all of them for the near-poly-A/T test of Stage 2
        nrid, nmask = ctx.shard_get_nreads()
        ncnt = torch.tensor([len(nrid)], dtype=torch.int64, device=self.device)
        nall = torch.empty(self.world, dtype=torch.int64, device=self.device)
        self._coll(dist.all_gather_into_tensor, nall, ncnt)
        if int(nall.sum().item()):
            gathered = [None] * self.world
            dist.all_gather_object(gathered, (np.array(nrid), np.array(nmask)))
            ctx.shard_set_nreads(np.concatenate([g[0] for g in gathered]), np.concatenate([g[1] for g in gathered]))
        torch.cuda.synchronize()
        # rounds
        rc = api.RoundControl()
        ctx.lib.mcb_round_control_init(C.byref(rc))
        tot_cl = members = 0
        while True:
            is_last = ctx.lib.mcb_round_control_begin(C.byref(rc), ctx.params.k, ctx.params.max_rounds)
            counts, _ = ctx.shard_partition(self.world)
            sc = torch.as_tensor(counts.astype(np.int64), device=self.device)
            rcnt = torch.empty(self.world, dtype=torch.int64, device=self.device)
            self._coll(dist.all_to_all_single, rcnt, sc)
            recv_counts = rcnt.cpu().numpy()
            n_send, n_recv = int(counts.sum()), int(recv_counts.sum())
            rptr, sptr = ctx.shard_recv_buffer(n_recv)
            send = dev_view(sptr, n_send * 16, self.device).view(torch.int64).view(-1, 2)
            recv = dev_view(rptr, n_recv * 16, self.device).view(torch.int64).view(-1, 2)
            self._coll(dist.all_to_all_single, recv, send, output_split_sizes=[int(x) for x in recv_counts], input_split_sizes=[int(x) for x in counts])
            torch.cuda.synchronize()
            self.bytes_exchanged += n_send * 16
            ctx.shard_set_tuples(n_recv)
            n_cl_new, n_mem_new, _, _ = ctx.bucket_round_a(int(rc.round), int(is_last))
            mine2 = torch.tensor([n_cl_new, n_mem_new], dtype=torch.int64, device=self.device)
            allc = torch.empty((self.world, 2), dtype=torch.int64, device=self.device)
            self._coll(dist.all_gather_into_tensor, allc, mine2)
            allc = allc.cpu().numpy()
            ctx.bucket_round_b(tot_cl + int(allc[:self.rank, 0].sum()))
            tot_cl += int(allc[:, 0].sum())
            members += int(allc[:, 1].sum())
            if ctx.lib.mcb_round_control_end(C.byref(rc), members):
                break
        br, rounds = ctx.bucket_finish()
        return rr, Stage1Part(br.cl_n, br.cl_a, br.cl_ref, np.diff(br.cl_ref_off.astype(np.int64)).astype(np.uint64), br.sg, br.mi, br.mi_cnt, rounds)

    # ---- Stage 2: one threshold round of realign_hash
    def realign(self, sg, refs, ref_off, g_lo: int, g_hi: int, threshold: int, maxsearch: int, ininumdict: int = 0):
        """Key-sharded: every rank passes ALL contigs (refs/ref_off; None/None = same as the previous round) and all singles,
        holds its hash range of the contig lt-mer table and probes only the lt-mers it owns; the claim priorities are
        min-reduced; this rank then emits the claims of windows [g_lo, g_hi) (contig_partition gives the ranges) with global
        contig indices."""
        torch, dist, ctx = self.torch, self.dist, self.ctx
        err, ptr, maxbin = None, 0, 0
        try:
            ptr, maxbin = ctx.realign_begin_keyed(sg, refs, ref_off, self.rank, self.world, g_lo, g_hi, threshold, maxsearch, ininumdict)
        except Exception as e:      # every rank has to reach the collective
            err = e
        flag = torch.tensor([1 if err else 0, maxbin], dtype=torch.int64, device=self.device)
        self._coll(dist.all_reduce, flag, op=dist.ReduceOp.MAX)
        bad, gmax = (int(x) for x in flag.cpu())
        if bad:
            raise err or RuntimeError("mcb_realign_begin_keyed failed on another rank")
        if gmax > maxsearch:
            raise RuntimeError(f"a dictionary bin may hold {gmax} singles (> maxsearch={maxsearch}): the sequential bin-window replay is single-GPU only")
        claim = dev_view(ptr, len(sg) * 8, self.device).view(torch.int64)
        if len(sg):
            self._coll(dist.all_reduce, claim, op=dist.ReduceOp.MIN)
            self.bytes_exchanged += len(sg) * 8
        torch.cuda.synchronize()
        return ctx.realign_finish()

    # ---- index builds: every rank sorts the buckets it owns
    def idx_build(self, tuples, bucket_off):
        """tuples bucket-major (n,2) uint64 + bucket_off[16385] of the WHOLE index; builds this rank's bucket range."""
        b0, b1 = bucket_range(self.rank, self.world)
        off = np.asarray(bucket_off, dtype=np.uint64)
        lo, hi = int(off[b0]), int(off[b1])
        mine = np.full(NB + 1, hi - lo, dtype=np.uint64)
        mine[:b0] = 0
        mine[b0:b1 + 1] = off[b0:b1 + 1] - np.uint64(lo)
        return self.ctx.idx_build(np.asarray(tuples).reshape(-1, 2)[lo:hi], mine)
