"""CPU tests of the sharding host logic (minicom_b200/shard.py): partition arithmetic, the tuple all-to-all over the gloo
backend with world_size 2, and the merge orders.  No GPU, no CUDA library calls."""
import os
import sys
import tempfile

import numpy as np
import pytest

from minicom_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rid_ranges_cover_everything():
    for n in (0, 1, 7, 1000, 10_000_019):
        for g in (1, 2, 3, 8):
            r = [shard.rid_range(n, q, g) for q in range(g)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(g - 1))
            cap = (n + g - 1) // g
            assert all(hi - lo <= cap for lo, hi in r)


def test_bucket_ranges_match_owner_rule():
    b = np.arange(shard.NB)
    for g in (1, 2, 3, 4, 5, 8):
        own = shard.owner_of_bucket(b, g)
        assert own.min() == 0 and own.max() == g - 1 and np.all(np.diff(own) >= 0)
        for q in range(g):
            b0, b1 = shard.bucket_range(q, g)
            assert np.array_equal(np.nonzero(own == q)[0], np.arange(b0, b1))


def test_local_singles_pick_the_owned_entries_in_list_order():
    n = 50
    own0, own1 = np.array([3, 9, 40, 41]), np.array([5, 7, 30])
    m0, m1 = shard.owned_mask(n, own0), shard.owned_mask(n, own1)
    job = np.array([41, 5, 3, 30, 9, 7], dtype=np.uint32)                 # the job's sg list after some updateSingle()
    s0, i0 = shard.local_singles(m0, job)
    s1, i1 = shard.local_singles(m1, job)
    assert list(s0) == [41, 3, 9] and list(i0) == [0, 2, 4]
    assert list(s1) == [5, 30, 7] and list(i1) == [1, 3, 5]
    assert sorted(list(i0) + list(i1)) == list(range(len(job)))


def test_merge_stage1_is_round_major_then_rank():
    m = 2
    def part(tag, rounds):
        ncl = sum(r[0] for r in rounds); nmem = sum(r[1] for r in rounds); nref = sum(r[2] for r in rounds); nsg = sum(r[3] for r in rounds)
        return shard.Stage1Part(np.full(ncl, tag, np.uint32), np.arange(nmem, dtype=np.uint64) + np.uint64(1000 * tag), np.full(nref, 65 + tag, np.uint8),
                                np.full(ncl, 7, np.uint64), np.arange(nsg, dtype=np.uint32) + np.uint32(100 * tag), np.zeros((ncl, m, 2), np.uint64),
                                np.full(ncl, m, np.uint8), np.array(rounds, dtype=np.uint64).reshape(-1, 4))
    a = part(1, [(2, 5, 14, 3), (1, 2, 7, 0)])
    b = part(2, [(1, 3, 7, 2), (0, 0, 0, 4)])
    mg = shard.merge_stage1([a, b])
    assert list(mg.cl_n) == [1, 1, 2, 1]                       # round 1: rank 0 (2 contigs), rank 1 (1); round 2: rank 0 (1)
    assert list(mg.cl_a) == [1000, 1001, 1002, 1003, 1004, 2000, 2001, 2002, 1005, 1006]
    assert list(mg.sg) == [100, 101, 102, 200, 201, 202, 203, 204, 205]
    assert mg.cl_ref.tobytes() == b"B" * 14 + b"C" * 7 + b"B" * 7
    assert mg.rounds.tolist() == [[3, 8, 21, 5], [1, 2, 7, 4]]


def test_merge_claims_by_priority_then_descending_single():
    # (contig, position in the job's sg list, y, priority); one step of the window loop (priority 64) claims on both ranks
    p0 = (np.array([0, 3, 9], np.uint32), np.array([5, 4, 1], np.uint32), np.array([11, 12, 13], np.uint64), np.array([32, 64, 900], np.uint64))
    p1 = (np.array([1, 3, 14], np.uint32), np.array([7, 6, 2], np.uint32), np.array([21, 22, 23], np.uint64), np.array([40, 64, 70], np.uint64))
    c, s, y, pr = shard.merge_claims([p0, p1])
    assert list(pr) == [32, 40, 64, 64, 70, 900]
    assert list(s) == [5, 7, 6, 4, 2, 1]                                  # inside priority 64: higher sg position first (kthread_hash_realign.c:388)
    assert list(y) == [11, 21, 22, 12, 23, 13] and list(c) == [0, 1, 3, 3, 14, 9]


WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from minicom_b200 import shard
out = sys.argv[2]
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n_total = 5003
lo, hi = shard.rid_range(n_total, rank, world)
rng = np.random.default_rng(99)                              # same stream on every rank: the global tuple list
x = rng.integers(0, 1 << 62, size=n_total, dtype=np.uint64)
rid = np.arange(n_total, dtype=np.uint64)
mine = np.stack([x[lo:hi], rid[lo:hi] << np.uint64(32)], axis=1)
own = shard.owner_of_bucket((mine[:, 0] & np.uint64(shard.NB - 1)).astype(np.int64), world)
order = np.argsort(own, kind="stable")                       # what mcb_shard_partition does on the device
counts = np.bincount(own, minlength=world)
send = torch.from_numpy(mine[order].view(np.int64))
recv_counts, recv = shard.all_to_all_rows(dist, send, counts, "cpu")
np.save(os.path.join(out, f"recv_{rank}.npy"), recv.numpy().view(np.uint64))
np.save(os.path.join(out, f"counts_{rank}.npy"), recv_counts)
dist.barrier()
dist.destroy_process_group()
'''


def test_tuple_all_to_all_over_gloo_world2():
    world = 2
    with tempfile.TemporaryDirectory() as td:
        script = os.path.join(td, "worker.py")
        with open(script, "w") as f:
            f.write(WORKER)
        import subprocess
        p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                            "--master-port", "29623", script, ROOT, td], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
        assert p.returncode == 0, p.stdout.decode()[-3000:]
        n_total = 5003
        rng = np.random.default_rng(99)
        x = rng.integers(0, 1 << 62, size=n_total, dtype=np.uint64)
        own = shard.owner_of_bucket((x & np.uint64(shard.NB - 1)).astype(np.int64), world)
        for q in range(world):
            got = np.load(os.path.join(td, f"recv_{q}.npy"))
            want_rid = np.nonzero(own == q)[0]                 # arrival order = source rank order = read-id order
            assert np.array_equal(got[:, 1] >> np.uint64(32), want_rid.astype(np.uint64))
            assert np.array_equal(got[:, 0], x[want_rid])
            cnt = np.load(os.path.join(td, f"counts_{q}.npy"))
            assert cnt.sum() == len(want_rid)
