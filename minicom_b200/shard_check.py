"""Parity check of the sharded front end against the single-GPU path, run under torchrun with one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 -m minicom_b200.shard_check

Every rank runs its share of kt_for_reads / kt_for_bucket (tuple all-to-all by bucket owner, packed-read all-gather), the
bucket-sharded index build and three threshold rounds of realign_hash (contigs partitioned, claim priorities min-reduced);
rank 0 also runs the whole job on one GPU and compares everything bit for bit.  Prints `SHARD CHECK OK`."""
import argparse
import os
import sys

import numpy as np


def main():
    import torch
    import torch.distributed as dist
    from . import api, shard, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=60000)
    ap.add_argument("--readlen", type=int, default=100)
    ap.add_argument("--genome", type=int, default=300000)
    ap.add_argument("--seed", type=int, default=17)
    ap.add_argument("--special", type=float, default=0.01)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    L, n_total = args.readlen, args.reads
    reads = synth.make_reads(n_total, L, args.genome, seed=args.seed, special=args.special)
    lo, hi = shard.rid_range(n_total, rank, world)
    ctx = api.Context(api.resolve_params(L, device=local))
    fe = shard.ShardedFrontEnd(ctx, dist, dev)
    rr, part = fe.stage1(np.ascontiguousarray(reads[lo:hi]), n_total)
    parts = [None] * world
    dist.all_gather_object(parts, part)
    cls_all = [None] * world
    dist.all_gather_object(cls_all, rr.cls)
    merged = shard.merge_stage1(parts)
    ok = True

    def check(name, a, b):
        nonlocal ok
        good = a.shape == b.shape and np.array_equal(a, b)
        if not good:
            ok = False
            print(f"[rank {rank}] MISMATCH {name}: {a.shape} vs {b.shape}", flush=True)

    single = None
    if rank == 0:
        single = api.Context(api.resolve_params(L, device=local))
        r1 = single.for_reads(reads)
        b1 = single.for_bucket()
        check("read classes", np.concatenate(cls_all), r1.cls)
        check("cl_n", merged.cl_n, b1.cl_n)
        check("cl_a", merged.cl_a, b1.cl_a)
        check("cl_ref", merged.cl_ref, b1.cl_ref)
        check("cl_reflen", merged.cl_reflen, np.diff(b1.cl_ref_off.astype(np.int64)).astype(np.uint64))
        check("sg", merged.sg, b1.sg)
        check("mi_cnt", merged.mi_cnt, b1.mi_cnt)
        m = ctx.params.first_mininum
        keep = np.arange(m)[None, :] < b1.mi_cnt[:, None]
        check("mi", merged.mi[keep], b1.mi[keep])
        print(f"[rank 0] stage 1: {len(b1.cl_n)} seed contigs, {len(b1.sg)} singles, {b1.rounds} rounds; sharded parts: {[len(p.cl_n) for p in parts]}", flush=True)
    # ---- index build over the seed-contig tuples, sharded by bucket range
    m = ctx.params.first_mininum
    keep = np.arange(m)[None, :] < merged.mi_cnt[:, None]
    flat = merged.mi[keep]
    bk = (flat[:, 0] & np.uint64(shard.NB - 1)).astype(np.int64)
    order = np.argsort(bk, kind="stable")
    off = np.zeros(shard.NB + 1, dtype=np.uint64)
    np.cumsum(np.bincount(bk, minlength=shard.NB), out=off[1:])
    xy = flat[order]
    ix = fe.idx_build(xy, off)
    b0, b1r = shard.bucket_range(rank, world)
    if rank == 0:
        ix1 = single.idx_build(xy, off)
    keys = np.unique(xy[:, 0])
    mine_keys = keys[(shard.owner_of_bucket((keys & np.uint64(shard.NB - 1)).astype(np.int64), world) == rank)]
    got = {int(x): ix.get(int(x)) for x in mine_keys[:4000]}
    allgot = [None] * world
    dist.all_gather_object(allgot, got)
    if rank == 0:
        n_chk = 0
        for g in allgot:
            for x, ys in g.items():
                n_chk += 1
                if not np.array_equal(ys, ix1.get(x)):
                    ok = False
                    print(f"[rank 0] MISMATCH index key {x:#x}", flush=True)
                    break
        print(f"[rank 0] index: {n_chk} keys checked over {world} bucket ranges", flush=True)
    # ---- Stage 2 over the seed contigs (no host merge in this check), three threshold rounds
    ref_off = np.concatenate([[0], np.cumsum(merged.cl_reflen.astype(np.int64))]).astype(np.uint64)
    cuts, wbase = shard.contig_partition(ref_off, world, L)
    lens = np.diff(ref_off.astype(np.int64))
    n_win = int(np.where(lens >= L, lens - L + 1, 0).sum())
    g_lo = int(wbase[rank]); g_hi = int(wbase[rank + 1]) if rank + 1 < world else n_win
    sg = merged.sg.copy()
    e = ctx.params.diff_threshold
    for rnd, thr in enumerate((e, 2 * e, 3 * e)):
        r = fe.realign(sg, merged.cl_ref if rnd == 0 else None, ref_off if rnd == 0 else None, g_lo, g_hi, thr, 2000)
        cl = [None] * world
        dist.all_gather_object(cl, (r.claim_contig, r.claim_sg, r.claim_y))
        gc, gs, gy = shard.merge_claims(cl)
        if rank == 0:
            w = single.realign(sg, merged.cl_ref if rnd == 0 else None, ref_off if rnd == 0 else None, thr, 2000)
            check(f"claims thr {thr}: y", gy, w.claim_y)
            check(f"claims thr {thr}: contig", gc, w.claim_contig)
            check(f"claims thr {thr}: sg", gs, w.claim_sg)
            check(f"fpA thr {thr}", r.fpA_sg, w.fpA_sg)
            print(f"[rank 0] realign thr {thr}: {len(sg)} singles, {len(w.claim_y)} claims; per rank {[len(c[2]) for c in cl]}", flush=True)
        flag = np.zeros(len(sg), dtype=bool)
        flag[gs] = True
        flag[r.fpA_sg] = True
        flag[r.fpT_sg] = True
        sg = sg[~flag]
    t = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(t)
    if rank == 0:
        print("SHARD CHECK OK" if int(t.item()) == 0 else "SHARD CHECK FAILED", f"({world} ranks, {n_total} reads, {fe.bytes_exchanged} bytes sent by rank 0)", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 0 else 1)


if __name__ == "__main__":
    main()
