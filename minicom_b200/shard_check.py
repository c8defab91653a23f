"""Parity check of the sharded front end against the single-GPU path, run under torchrun with one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 -m minicom_b200.shard_check

Every rank runs its share of kt_for_reads / kt_for_bucket (tuples + packed rows exchanged by bucket owner inside the library),
the bucket-sharded index build and three threshold rounds of realign_hash on the singles it owns; rank 0 also runs the whole job
on one GPU and compares everything bit for bit.  `--repeats` plants a repeat family (many near-identical singles sharing their
dictionary keys) so that dictionary bins exceed `--maxsearch` and the sequential bin-window replay runs sharded.
torch.distributed is only the launcher's plumbing here (unique id, gathering results for the comparison).  Prints `SHARD CHECK OK`."""
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
    ap.add_argument("--repeats", type=int, default=0, help="plant this many mutated copies of one read (big dictionary bins)")
    ap.add_argument("--maxsearch", type=int, default=2000)
    ap.add_argument("--bigbins", action="store_true", help="Stage 2 only, on a high-duplication set whose dictionary bins exceed a small maxsearch")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")                       # host-side plumbing only: the data path uses the library's own communicator
    L, n_total = args.readlen, args.reads
    if args.bigbins:
        sys.exit(big_bins(args, rank, world, local, dist))
    reads = synth.make_reads(n_total, L, args.genome, seed=args.seed, special=args.special)
    if args.repeats:
        # a family of reads that differ from one template in 6-9 scattered bases: they fail the Stage-1 consensus test (e = 4) or
        # stay alone, become singles, and share most of their dictionary keys -> bins far above a small maxsearch
        rng = np.random.default_rng(args.seed + 5)
        tpl = reads[7].copy()
        rows = rng.choice(n_total, size=args.repeats, replace=False)
        for r in rows:
            v = tpl.copy()
            pos = rng.choice(L, size=int(rng.integers(6, 10)), replace=False)
            v[pos] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=pos.size)]
            reads[r] = v
    lo, hi = shard.rid_range(n_total, rank, world)
    ctx = api.Context(api.resolve_params(L, device=local))
    ctx.timers_enable(True)
    fe = shard.ShardedFrontEnd(ctx, rank, world, shard.make_unique_id(dist))
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
    sg = merged.sg.copy()
    e = ctx.params.diff_threshold
    for rnd, thr in enumerate((e, 2 * e, 3 * e)):
        mine, idx = fe.select(sg)
        r = fe.realign(mine, idx, len(sg), merged.cl_ref if rnd == 0 else None, ref_off if rnd == 0 else None, thr, args.maxsearch)
        cl = [None] * world
        dist.all_gather_object(cl, (r.claim_contig, r.claim_sg, r.claim_y, r.claim_prio, r.fpA_sg, r.fpT_sg))
        gc, gs, gy, _ = shard.merge_claims(cl)
        fpa, fpt = np.sort(np.concatenate([c[4] for c in cl])), np.sort(np.concatenate([c[5] for c in cl]))
        if rank == 0:
            w = single.realign(sg, merged.cl_ref if rnd == 0 else None, ref_off if rnd == 0 else None, thr, args.maxsearch)
            check(f"claims thr {thr}: y", gy, w.claim_y)
            check(f"claims thr {thr}: contig", gc, w.claim_contig)
            check(f"claims thr {thr}: sg", gs, w.claim_sg)
            check(f"fpA thr {thr}", fpa, w.fpA_sg)
            check(f"fpT thr {thr}", fpt, w.fpT_sg)
            print(f"[rank 0] realign thr {thr}: {len(sg)} singles, {len(w.claim_y)} claims; per rank {[len(c[2]) for c in cl]}", flush=True)
        flag = np.zeros(len(sg), dtype=bool)
        flag[gs] = True
        flag[fpa] = True
        flag[fpt] = True
        sg = sg[~flag]
    tm = ctx.timers()
    sent = int(tm.get("nccl_bytes_sent", (0.0, 0))[0])
    t = torch.tensor([0 if ok else 1])
    dist.all_reduce(t)
    if rank == 0:
        print("SHARD CHECK OK" if int(t.item()) == 0 else "SHARD CHECK FAILED", f"({world} ranks, {n_total} reads, {sent} bytes sent by rank 0 in Stage 1, "
              f"collectives {sum(v[0] for k, v in tm.items() if k.startswith('nccl:')):.2f} ms)", flush=True)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 0 else 1)


def big_bins(args, rank, world, local, dist):
    """The shape of tests/test_gpu_parity.py::test_realign_exact_replay_of_big_bins, sharded: 1500 reads over a 130 bp genome
    (31 start positions, both strands) against the genome and pieces of it as contigs, small maxsearch.  Every rank realigns the
    sketched reads of its own read-id slice; the bins span the ranks, so the bound, the events and the replay are collective."""
    import torch
    from . import api, shard, synth
    L, n_total = 100, 1500
    genome = synth.make_genome(130, 77)
    reads = synth.make_reads(n_total, L, 130, seed=77, genome=genome, special=0.01)
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    contigs = [genome, comp[genome[::-1]], genome[:115], genome[12:]]
    refs = np.concatenate(contigs)
    off = np.concatenate([[0], np.cumsum([len(c) for c in contigs])]).astype(np.uint64)
    lo, hi = shard.rid_range(n_total, rank, world)
    ctx = api.Context(api.resolve_params(L, device=local))
    fe = shard.ShardedFrontEnd(ctx, rank, world, shard.make_unique_id(dist))
    ctx.shard_begin(n_total, lo)
    rr = ctx.for_reads(np.ascontiguousarray(reads[lo:hi]))
    cls_all = [None] * world
    dist.all_gather_object(cls_all, rr.cls)
    sg = np.nonzero(np.concatenate(cls_all) == 0)[0].astype(np.uint32)
    fe.mask = np.zeros(n_total, dtype=bool)
    fe.mask[lo:hi] = True
    ok = True
    single = None
    if rank == 0:
        single = api.Context(api.resolve_params(L, device=local))
        single.for_reads(reads)
    for maxsearch in (3, 11, 40):
        cur = sg.copy()
        for thr in (4, 8, 30):
            mine, idx = fe.select(cur)
            r = fe.realign(mine, idx, len(cur), refs, off, thr, maxsearch)
            cl = [None] * world
            dist.all_gather_object(cl, (r.claim_contig, r.claim_sg, r.claim_y, r.claim_prio, r.fpA_sg, r.fpT_sg))
            gc, gs, gy, _ = shard.merge_claims(cl)
            fpa, fpt = np.sort(np.concatenate([c[4] for c in cl])), np.sort(np.concatenate([c[5] for c in cl]))
            if rank == 0:
                w = single.realign(cur, refs, off, thr, maxsearch)
                same = np.array_equal(gy, w.claim_y) and np.array_equal(gc, w.claim_contig) and np.array_equal(gs, w.claim_sg) and np.array_equal(fpa, w.fpA_sg) and np.array_equal(fpt, w.fpT_sg)
                print(f"[rank 0] big bins, maxsearch {maxsearch}, thr {thr}: {len(cur)} singles, {len(w.claim_y)} claims (sharded {len(gy)}): {'equal' if same else 'MISMATCH'}", flush=True)
                ok = ok and same
            flag = np.zeros(len(cur), dtype=bool)
            flag[gs] = True
            flag[fpa] = True
            flag[fpt] = True
            cur = cur[~flag]
            if len(cur) == 0:
                break
    t = torch.tensor([0 if ok else 1])
    dist.all_reduce(t)
    if rank == 0:
        print("SHARD CHECK OK" if int(t.item()) == 0 else "SHARD CHECK FAILED", f"(big bins, {world} ranks)", flush=True)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    return 0 if int(t.item()) == 0 else 1


if __name__ == "__main__":
    main()
