"""The front end over the GPUs of one box, one process per GPU: thin driver over the sharded C-ABI (include/minicom_b200.h,
"the same path over the GPUs of one box").  All exchanges between GPUs — NCCL grouped send/recv over NVLink — happen inside
libminicom_b200.so (csrc/mcb_shard.cu); this module only hands every rank its inputs and knows how the ranks' results combine.

Partitioning rules:
  reads      contiguous read-id ranges, rank r owns [r*cap, min((r+1)*cap, N)), cap = ceil(N / G)
  buckets    the 16384 minimizer buckets (x & 0x3FFF, kthread_reads.c:213) in contiguous ranges, owner = bucket*G >> 14
  singles    a single belongs to the rank whose bucket it was left alone in (that rank holds its packed row)

Exchanges:
  Stage 1    per round ONE exchange of (16-byte tuple + packed row of its read) by bucket owner, plus two counters per rank
             (new seed contigs for the global contig ids, members for the loop control)
  Stage 2    none for the data: every rank realigns its own singles against all contigs; scalar guards of the bin sizes
Concatenating the ranks' Stage-1 results in rank order — round by round — and merging the claim lists by priority reproduces
the single-GPU (= single-threaded reference) order exactly; `merge_stage1` / `merge_claims` do that on one host.

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


def merge_claims(parts):
    """Per-rank (claim_contig, claim_sg, claim_y, claim_prio) -> the job's lists in the reference's append order: ascending
    priority (window, forward-then-reverse, dictionary), inside one step descending position in the job's sg list."""
    c = np.concatenate([p[0] for p in parts]); s = np.concatenate([p[1] for p in parts])
    y = np.concatenate([p[2] for p in parts]); pr = np.concatenate([p[3] for p in parts])
    order = np.lexsort((np.uint32(0xFFFFFFFF) - s, pr))
    return c[order], s[order], y[order], pr[order]


def owned_mask(n_reads: int, sg_of_rank: np.ndarray) -> np.ndarray:
    """bitmap over read ids: the singles this rank produced in Stage 1 (it holds their packed rows)"""
    m = np.zeros(n_reads, dtype=bool)
    m[np.asarray(sg_of_rank, dtype=np.int64)] = True
    return m


def local_singles(mask: np.ndarray, sg_job: np.ndarray):
    """the entries of the job's sg list that this rank owns, with their positions in the list"""
    pos = np.nonzero(mask[np.asarray(sg_job, dtype=np.int64)])[0]
    return np.ascontiguousarray(np.asarray(sg_job)[pos], dtype=np.uint32), pos.astype(np.uint32)


# ---------------------------------------------------------------- host model of the exchange (CPU tests)
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


def make_unique_id(dist) -> bytes:
    """rank 0 asks NCCL for a unique id (through the library) and the process group hands it to everyone"""
    from . import api
    box = [None]
    if dist.get_rank() == 0:
        buf = C.create_string_buffer(api.NCCL_ID_BYTES)
        lib = api.load_library()
        if lib.mcb_shard_unique_id(buf) != 0:
            raise api.McbError(lib.mcb_last_error().decode())
        box[0] = buf.raw
    dist.broadcast_object_list(box, src=0)
    return box[0]


class ShardedFrontEnd:
    """Drives one rank of the sharded front end.  `ctx` is an api.Context on this rank's GPU; the communicator is created
    inside the library from `unique_id` (make_unique_id)."""

    def __init__(self, ctx, rank: int, world: int, unique_id: bytes):
        self.ctx, self.rank, self.world = ctx, rank, world
        ctx.shard_init(unique_id, rank, world)
        self.mask = None

    # ---- Stage 1: kt_for_reads + kt_for_bucket over the whole read set
    def stage1(self, rows_local, n_total: int, device_resident: bool = False, keep_mask: bool = False):
        """rows_local: this rank's reads: (n_local, L) uint8 ASCII rows (numpy, or a CUDA tensor when device_resident), an
        api.ReadSet (packed on the host by the library's FASTQ reader), or a tuple (d_packed, d_nread_rid, d_nmask, n_nreads) of
        device pointers to the same packed arrays.
        Returns (ReadsResult of the slice, Stage1Part of this rank).  keep_mask: the ownership bitmap of an earlier call on the
        same reads is still right (repeated runs of one job), do not rebuild it."""
        ctx = self.ctx
        lo, hi = rid_range(n_total, self.rank, self.world)
        ctx.shard_begin(n_total, lo)
        if isinstance(rows_local, tuple):
            rr = ctx.for_reads_packed_device(rows_local[0], hi - lo, rows_local[1], rows_local[2], rows_local[3])
        elif hasattr(rows_local, "view") and hasattr(rows_local, "add_rows"):
            assert len(rows_local) == hi - lo
            rr = ctx.for_reads_packed(rows_local)
        else:
            rr = ctx.for_reads_device(rows_local.data_ptr(), hi - lo) if device_resident else ctx.for_reads(rows_local)
        br, rounds = ctx.shard_for_bucket()
        if not (keep_mask and self.mask is not None):
            self.mask = owned_mask(n_total, br.sg)
        return rr, Stage1Part(br.cl_n, br.cl_a, br.cl_ref, np.diff(br.cl_ref_off.astype(np.int64)).astype(np.uint64), br.sg, br.mi, br.mi_cnt, rounds)

    # ---- Stage 2: one threshold round of realign_hash
    def select(self, sg_job):
        """(this rank's singles of the job's list, their positions in it)"""
        return local_singles(self.mask, sg_job)

    def realign(self, sg_local, sg_index, n_sg_total: int, refs, ref_off, threshold: int, maxsearch: int, ininumdict: int = 0):
        """ALL contigs (refs/ref_off; None/None = the previous round's) against this rank's singles.  Collective."""
        return self.ctx.shard_realign(sg_local, sg_index, n_sg_total, refs, ref_off, threshold, maxsearch, ininumdict)

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
