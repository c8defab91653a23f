"""GPU parity: every entry point of the C-ABI against the state dumps of the reference itself (oracle/_ref, num_thr=1).

Bit-exact comparison of: bucket tuples after kt_for_reads, read classes, N bookkeeping, N-replaced reads, seed
contigs / singles / index tuples after kt_for_bucket, every minimizer index the host merger builds (posting order
included), and per threshold round of realign_hash the reads claimed by each contig in append order plus the
near-poly-A/T diversions.
"""
import numpy as np
import pytest

import refdump
from minicom_b200 import api, synth

pytestmark = pytest.mark.gpu

# (name, n_reads, L, genome, seed, special fraction, reference env options)
CASES = [
    ("20k_special", 20000, 100, 100000, 3, 0.01, {}),
    ("60k_plain", 60000, 100, 300000, 5, 0.0, {}),
    ("30k_k25_e6", 30000, 100, 150000, 7, 0.005, {"MC_K": 25, "MC_E": 6, "MC_M": 4, "MC_W": 12, "MC_S": 3, "MC_STEP": 3, "MC_EMAX": 30}),
]


def params_for(L, env):
    return api.resolve_params(L, k=int(env.get("MC_K", 0)), e=int(env.get("MC_E", 0)), w=int(env.get("MC_W", 0)),
                              m=int(env.get("MC_M", 0)), max_rounds=int(env.get("MC_MAXROUNDS", 0)))


def bucket_major(tuples, b=14):
    """stable partition of (n,2) tuples by bucket = x & (2^b-1): the order per-bucket pushes would give."""
    bk = (tuples[:, 0] & np.uint64((1 << b) - 1)).astype(np.int64)
    order = np.argsort(bk, kind="stable")
    cnt = np.bincount(bk, minlength=1 << b).astype(np.uint64)
    off = np.zeros((1 << b) + 1, dtype=np.uint64)
    np.cumsum(cnt, out=off[1:])
    return off, tuples[order]


def check_reads(ctx, reads, dump, log):
    n, L = reads.shape
    rr = ctx.for_reads(reads)
    for code, name in ((1, "r_allA.u32"), (2, "r_allT.u32"), (3, "r_allN.u32"), (4, "r_fpA.u32"), (5, "r_fpT.u32"), (6, "r_fpN.u32"), (7, "r_Nfile.u32")):
        want = dump.arr(name)
        got = np.nonzero(rr.cls == code)[0].astype(np.uint32)
        assert np.array_equal(got, want), f"class list {name}: got {len(got)} want {len(want)}"
    t = ctx.debug_read_tuples(n)
    valid = t[:, 0] != np.uint64(0xFFFFFFFFFFFFFFFF)
    assert np.array_equal(valid, rr.cls == 0)
    assert rr.n_sketched == int(valid.sum())
    off, mine = bucket_major(t[valid])
    roff, rxy = dump.buckets("r_B0")
    assert np.array_equal(off, roff), "bucket sizes differ"
    assert np.array_equal(mine, rxy), "bucket tuples differ"
    # N bookkeeping
    want_np = dump.npos(n)
    has = np.array([len(x) > 0 for x in want_np])
    assert np.array_equal(np.nonzero(has)[0].astype(np.uint32), rr.nread_rid)
    for i, rid in enumerate(rr.nread_rid):
        assert np.array_equal(rr.npos[int(rr.nread_off[i]):int(rr.nread_off[i + 1])], want_np[int(rid)])
    # N-replaced sequences of the sketched reads
    seqs = dump.seqs()
    un = ctx.debug_unpack_reads(n)
    sk = np.nonzero(rr.cls == 0)[0]
    want = np.frombuffer(b"".join(seqs[i] for i in sk), dtype=np.uint8).reshape(len(sk), L)
    assert np.array_equal(un[sk], want), "N-replaced reads differ"
    for i, rid in enumerate(rr.nread_rid):
        if rr.cls[rid] == 0:
            assert rr.nread_repl[i] == seqs[int(rid)][int(rr.npos[int(rr.nread_off[i])])]
    log.append(f"reads: {n} reads, {rr.n_sketched} sketched, {len(rr.nread_rid)} with N: OK")
    return rr


def check_bucket(ctx, dump, log):
    br = ctx.for_bucket()
    cl = dump.clusters("b_cl")
    assert len(br.cl_n) == len(cl["n"]), f"cluster count {len(br.cl_n)} vs {len(cl['n'])}"
    assert np.array_equal(br.cl_n.astype(np.uint64), cl["n"])
    assert np.array_equal(br.cl_a, cl["a"]), "cluster members differ"
    assert np.array_equal(br.cl_ref_off, cl["ref_off"]), "consensus lengths differ"
    assert np.array_equal(br.cl_ref, cl["ref"]), "consensus strings differ"
    assert np.array_equal(br.sg, dump.arr("b_sg.u32")), "singles (order) differ"
    m = ctx.params.first_mininum
    keep = np.arange(m)[None, :] < br.mi_cnt[:, None]
    flat = br.mi[keep]
    off, mine = bucket_major(flat)
    roff, rxy = dump.buckets("b_mi")
    assert np.array_equal(off, roff) and np.array_equal(mine, rxy), "mi[0] tuples differ"
    log.append(f"bucket: {len(br.cl_n)} seed contigs, {len(br.sg)} singles, {len(flat)} index tuples, {br.rounds} rounds: OK")
    return br


def check_index(ctx, dump, log):
    for j in range(dump.n_idx()):
        off, xy = dump.buckets(f"i{j}_in")
        ix = ctx.idx_build(xy, off)
        post = dump.postings(j)
        nk, npost = ix.stats()
        assert nk == len(post) and npost == len(xy)
        for x, ys in post:
            got = ix.get(x)
            assert np.array_equal(got, ys), f"index {j} key {x:#x}: {got} vs {ys}"
        assert len(ix.get(0x123456789)) == 0 or any(x == 0x123456789 for x, _ in post)
        ix.close()
        log.append(f"index {j}: {len(xy)} tuples, {nk} keys: OK")


def check_realign(ctx, dump, n_sg_stage1, env, log):
    cl = dump.clusters("c_cl")
    maxsearch = 2000 if n_sg_stage1 <= 5000000 else 500
    for j in range(dump.n_realign()):
        sg = dump.arr(f"h{j}_sg.u32")
        thr = int(dump.arr(f"h{j}_thr.u64")[0])
        rr = ctx.realign(sg, cl["ref"], cl["ref_off"], thr, maxsearch, int(env.get("MC_S", 0)))
        want_cnt = dump.arr(f"h{j}_app_cnt.u64")
        want_y = dump.arr(f"h{j}_app_y.u64")
        got_cnt = np.bincount(rr.claim_contig, minlength=len(want_cnt)).astype(np.uint64)
        assert len(rr.claim_y) == len(want_y), f"round {j} thr {thr}: {len(rr.claim_y)} claims vs {len(want_y)}"
        assert np.array_equal(got_cnt, want_cnt), f"round {j}: per-contig claim counts differ"
        assert np.all(np.diff(rr.claim_contig.astype(np.int64)) >= 0)
        assert np.array_equal(rr.claim_y, want_y), f"round {j}: claims / append order differ"
        assert np.array_equal(sg[rr.claim_sg], (rr.claim_y >> np.uint64(32)).astype(np.uint32))
        assert np.array_equal(sg[rr.fpA_sg], dump.arr(f"h{j}_fpA.u32")), "near-poly-A list differs"
        assert np.array_equal(sg[rr.fpT_sg], dump.arr(f"h{j}_fpT.u32")), "near-poly-T list differs"
        flag = np.zeros(len(sg), dtype=np.uint8)
        flag[rr.claim_sg] = 1
        flag[rr.fpA_sg] = 1
        flag[rr.fpT_sg] = 1
        assert np.array_equal(flag, dump.arr(f"h{j}_flag.u8")), "sg_flag differs"
        log.append(f"realign {j}: thr {thr}, {len(sg)} singles, {rr.n_windows} windows, {rr.n_probes} probes, {rr.n_candidates} candidates, "
                   f"{len(rr.claim_y)} claims, {len(rr.fpA_sg)}+{len(rr.fpT_sg)} polyA/T: OK")


def run_case(case, log):
    name, n, L, G, seed, special, env = case
    reads, dump = refdump.cached_reference(n, L, G, seed, special, "sg", env)
    with api.Context(params_for(L, env)) as ctx:
        check_reads(ctx, reads, dump, log)
        br = check_bucket(ctx, dump, log)
        check_index(ctx, dump, log)
        check_realign(ctx, dump, len(br.sg), env, log)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_front_end_matches_reference(case):
    log = []
    try:
        run_case(case, log)
    finally:
        print("\n".join(log))


@pytest.mark.parametrize("name", refdump.golden_names())
def test_front_end_matches_golden_fixture(name):
    """Same state-by-state comparison against the committed fixtures (tests/golden/*.npz: dumps of the unmodified
    reference) — needs neither /root/reference nor a cache."""
    reads, meta, dump, _ = refdump.load_golden(name)
    env, log = meta["env"], []
    try:
        with api.Context(params_for(meta["L"], env)) as ctx:
            check_reads(ctx, reads, dump, log)
            br = check_bucket(ctx, dump, log)
            check_index(ctx, dump, log)
            check_realign(ctx, dump, len(br.sg), env, log)
    finally:
        print("\n".join(log))


# shapes without a reference build: the oracle restatement (oracle/mc_oracle.c, itself pinned against the reference) is the checker
ORACLE_CASES = [
    # (name, n_reads, L, genome, seed, special, options)
    ("L36_min_k", 6000, 36, 20000, 31, 0.01, {"k": 12}),
    ("L64", 6000, 64, 30000, 32, 0.01, {}),
    ("L125", 5000, 125, 30000, 33, 0.01, {}),
    ("L200_w30", 4000, 200, 40000, 34, 0.01, {"w": 30, "m": 3}),
    ("L256_max", 3000, 256, 40000, 35, 0.01, {"e": 8}),
    ("n1", 1, 100, 1000, 36, 0.0, {}),
    ("n127", 127, 100, 600, 37, 0.02, {}),
    ("n129", 129, 100, 600, 38, 0.02, {}),
    ("dup_heavy", 8000, 100, 4000, 39, 0.0, {}),          # 200x coverage: big groups, long contigs, buckets above 64 tuples
    ("rounds_many", 8000, 100, 40000, 40, 0.0, {"k": 14, "e": 2}),
    ("huge_groups", 150000, 100, 130, 42, 0.0, {}),        # 31 start positions: groups beyond the bit-sliced counters (> 60000 members) and of every size class
]


@pytest.mark.parametrize("case", ORACLE_CASES, ids=[c[0] for c in ORACLE_CASES])
def test_front_end_matches_oracle(case):
    import oracle_lib as O
    name, n, L, G, seed, special, opt = case
    reads = synth.make_reads(n, L, G, seed=seed, special=special)
    _check_front_end_against_oracle(reads, L, opt)


def _tandem_genome(seed, total):
    """Random stretches interleaved with tandem repeats (period 1..40): windows that hold the same k-mer several times,
    the case mm_sketch_lh_ori's identical-hash scans (sketch.c:141-146,159-161) exist for."""
    rng = np.random.default_rng([seed, 0x74616E])
    parts, n = [], 0
    while n < total:
        parts.append(synth.make_genome(int(rng.integers(40, 200)), int(rng.integers(1 << 30))))
        unit = synth.make_genome(int(rng.integers(1, 41)), int(rng.integers(1 << 30)))
        parts.append(np.tile(unit, int(rng.integers(80, 400)) // len(unit) + 1))
        n += len(parts[-1]) + len(parts[-2])
    return np.concatenate(parts)


@pytest.mark.parametrize("L,opt,seed", [(100, {}, 51), (100, {"k": 30}, 52), (64, {"k": 20, "w": 7}, 53), (150, {"w": 40, "m": 12}, 54)],
                         ids=["L100", "L100_even_k", "L64_w7", "L150_w40_m12"])
def test_front_end_matches_oracle_on_tandem_repeats(L, opt, seed):
    genome = _tandem_genome(seed, 30000)
    reads = synth.make_reads(8000, L, len(genome), seed=seed, sub_rate=0.003, genome=genome)
    _check_front_end_against_oracle(reads, L, opt)


def _check_front_end_against_oracle(reads, L, opt):
    import oracle_lib as O
    n = len(reads)
    S = O.Stage1(O.resolve_params(L, **opt), reads)
    with api.Context(api.resolve_params(L, **opt)) as ctx:
        rr = ctx.for_reads(reads)
        assert np.array_equal(rr.cls, S.cls)
        t = ctx.debug_read_tuples(n)
        assert np.array_equal(t, S.tuples)
        sk = np.nonzero(S.cls == 0)[0]
        assert np.array_equal(ctx.debug_unpack_reads(n)[sk], S.rows[sk])
        br = ctx.for_bucket()
        assert br.rounds == S.rounds
        assert np.array_equal(br.cl_n, S.cl_n) and np.array_equal(br.cl_a, S.cl_a)
        assert np.array_equal(br.cl_ref_off, S.cl_ref_off) and np.array_equal(br.cl_ref, S.cl_ref)
        assert np.array_equal(br.sg, S.sg)
        m = ctx.params.first_mininum
        flat = br.mi[np.arange(m)[None, :] < br.mi_cnt[:, None]]
        assert np.array_equal(flat, S.mi)
        assert br.n_sketched_total == S.n_sketched_total
        # index over the seed-contig tuples (the first mm_idx_generation of the host merge)
        off, xy = O.bucket_major(S.mi)
        ix, ox = ctx.idx_build(xy, off), O.Index(xy, off)
        keys, st, post = ox.flat()
        assert ix.stats() == (ox.n_keys, ox.n_post)
        for i, x in enumerate(keys):
            assert np.array_equal(ix.get(int(x)), post[int(st[i]):int(st[i + 1])])
        ix.close()
        # the threshold schedule over the seed contigs (no host merge in between: contigs = seed contigs)
        sg = S.sg.copy()
        e = ctx.params.diff_threshold
        for rnd, thr in enumerate((e, 2 * e, 3 * e, 28)):
            if len(sg) == 0 or len(S.cl_n) == 0:
                break
            want = S.realign(sg, S.cl_ref, S.cl_ref_off, thr, 2000)
            # later rounds reuse the contigs of the first call (refs=None), as the drop-in shim does
            got = ctx.realign(sg, S.cl_ref, S.cl_ref_off, thr, 2000) if rnd == 0 else ctx.realign(sg, None, None, thr, 2000)
            assert np.array_equal(got.claim_y, want["claim_y"]), f"thr {thr}: claims / append order differ"
            assert np.array_equal(got.claim_contig, want["claim_contig"]) and np.array_equal(got.claim_sg, want["claim_sg"])
            assert np.array_equal(got.fpA_sg, want["fpA_sg"]) and np.array_equal(got.fpT_sg, want["fpT_sg"])
            assert (got.n_windows, got.n_probes, got.numdict) == (want["n_windows"], want["n_probes"], want["numdict"])
            sg = sg[want["flag"] == 0]
    S.close()


def test_realign_contig_cache_is_invalidated_by_content():
    """Two contig sets of identical shape but different bases: the cached device table must not survive the second call."""
    import oracle_lib as O
    L = 100
    genome = synth.make_genome(6000, 41)
    reads = synth.make_reads(1500, L, 6000, seed=41, genome=genome)
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    contigs_a = genome.copy()
    contigs_b = comp[genome[::-1]].copy()                  # same length, reverse complement
    off = np.array([0, 2500, 2560, 6000], dtype=np.uint64)   # includes a contig shorter than a read
    sg = np.arange(len(reads), dtype=np.uint32)
    S = O.Stage1(O.resolve_params(L), reads)
    with api.Context(api.resolve_params(L)) as ctx:
        ctx.for_reads(reads)
        for refs in (contigs_a, contigs_b, contigs_b, contigs_a):
            want = S.realign(sg, refs, off, 8, 2000)
            got = ctx.realign(sg, refs, off, 8, 2000)
            assert len(want["claim_y"]) > 500
            assert np.array_equal(got.claim_y, want["claim_y"]) and np.array_equal(got.claim_contig, want["claim_contig"])
    S.close()


def test_realign_bins_above_maxsearch():
    """20 identical singles share every dictionary bin.  The reference scans only the last `maxsearch` live entries of a bin
    per probe (kthread_hash_realign.c:388) and removes claimed reads afterwards, so with maxsearch 4 window 0 claims them four
    at a time, dictionary after dictionary.  The library detects the oversized bins and replays those singles sequentially."""
    import oracle_lib as O
    rng = np.random.default_rng(9)
    L = 100
    read = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=L)]
    reads = np.tile(read, (20, 1))
    contig = np.concatenate([read, read])
    off = np.array([0, 2 * L], dtype=np.uint64)
    sg = np.arange(20, dtype=np.uint32)
    S = O.Stage1(O.resolve_params(L), reads)
    with api.Context(api.resolve_params(L)) as ctx:
        ctx.for_reads(reads)
        for maxsearch in (2000, 4, 1, 7):
            want = S.realign(sg, contig, off, 4, maxsearch)
            got = ctx.realign(sg, contig, off, 4, maxsearch)
            assert np.array_equal(got.claim_y, want["claim_y"]) and np.array_equal(got.claim_sg, want["claim_sg"]), f"maxsearch {maxsearch}"
    S.close()


@pytest.mark.parametrize("maxsearch", [3, 11, 40])
def test_realign_exact_replay_of_big_bins(maxsearch):
    """High-duplication reads (31 start positions over a 130 bp genome, both strands, 1 % substitutions, a few poly-A reads
    that stay in the bins) against the genome and its shuffled pieces as contigs, with a small maxsearch: most bins are
    larger than the scan window, so the outcome depends on the order in which reads leave the bins."""
    import oracle_lib as O
    L = 100
    genome = synth.make_genome(130, 77)
    reads = synth.make_reads(1500, L, 130, seed=77, genome=genome, special=0.01)
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    contigs = [genome, comp[genome[::-1]], genome[:115], genome[12:]]
    refs = np.concatenate(contigs)
    off = np.concatenate([[0], np.cumsum([len(c) for c in contigs])]).astype(np.uint64)
    S = O.Stage1(O.resolve_params(L), reads)
    sg = np.nonzero(S.cls == 0)[0].astype(np.uint32)
    with api.Context(api.resolve_params(L)) as ctx:
        ctx.for_reads(reads)
        for thr in (4, 8, 30):
            want = S.realign(sg, refs, off, thr, maxsearch)
            got = ctx.realign(sg, refs, off, thr, maxsearch)
            assert np.array_equal(got.claim_y, want["claim_y"]), f"thr {thr}: {len(got.claim_y)} vs {len(want['claim_y'])} claims"
            assert np.array_equal(got.claim_contig, want["claim_contig"]) and np.array_equal(got.claim_sg, want["claim_sg"])
            sg = sg[want["flag"] == 0]
            if len(sg) == 0:
                break
    S.close()


@pytest.mark.parametrize("n,distinct,nbuckets", [(200000, 4000, 300), (60000, 40, 7), (500000, 200000, 16384), (30000, 3, 1)])
def test_index_posting_order_in_big_buckets(n, distinct, nbuckets):
    """Buckets far above 64 tuples with many equal minimizers: the posting order is whatever the reference's unstable
    in-place radix sort leaves (ksort.h:108-157); the oracle replays it on the CPU."""
    import oracle_lib as O
    rng = np.random.default_rng(n + distinct)
    bks = rng.choice(1 << 14, size=nbuckets, replace=False).astype(np.uint64)
    keys = (rng.integers(0, 1 << 48, size=distinct, dtype=np.uint64) << np.uint64(14)) | bks[rng.integers(0, nbuckets, size=distinct)]
    xy = np.zeros((n, 2), dtype=np.uint64)
    xy[:, 0] = keys[rng.integers(0, distinct, size=n)]
    xy[:, 1] = np.arange(n, dtype=np.uint64)
    off, flat = O.bucket_major(xy)
    with api.Context(api.resolve_params(100)) as ctx:
        ix, ox = ctx.idx_build(flat, off), O.Index(flat, off)
        okeys, st, post = ox.flat()
        assert ix.stats() == (ox.n_keys, ox.n_post)
        for i in rng.choice(len(okeys), size=min(len(okeys), 3000), replace=False):
            assert np.array_equal(ix.get(int(okeys[i])), post[int(st[i]):int(st[i + 1])]), f"key {int(okeys[i]):#x}"
        ix.close()
        ox.close()


def test_empty_input():
    with api.Context(api.resolve_params(100)) as ctx:
        rr = ctx.for_reads(np.zeros((0, 100), dtype=np.uint8))
        assert rr.n_sketched == 0 and len(rr.cls) == 0
        br = ctx.for_bucket()
        assert len(br.cl_n) == 0 and len(br.sg) == 0


def test_bad_characters_are_rejected_loudly():
    reads = synth.make_reads(1000, 100, 5000, seed=1)
    reads[17, 5] = ord("x")
    with api.Context(api.resolve_params(100)) as ctx:
        with pytest.raises(api.McbError):
            ctx.for_reads(reads)


def test_host_buffer_variants_agree():
    """mcb_for_reads (contiguous rows), mcb_for_reads_ptrs (scattered strings) give identical tuples."""
    name, n, L, G, seed, special, env = CASES[0]
    reads = synth.make_reads(n, L, G, seed=seed, special=special)
    with api.Context(params_for(L, env)) as ctx:
        ctx.for_reads(reads)
        t0 = ctx.debug_read_tuples(n)
    bufs = [np.frombuffer(reads[i].tobytes() + b"\0", dtype=np.uint8).copy() for i in range(n)]
    ptrs = np.array([b.ctypes.data for b in bufs], dtype=np.uint64)
    with api.Context(params_for(L, env)) as ctx:
        ctx.for_reads_ptrs(ptrs, n_threads=4)
        t1 = ctx.debug_read_tuples(n)
    assert np.array_equal(t0, t1)


if __name__ == "__main__":
    import sys
    import traceback
    sel = sys.argv[1:] or [c[0] for c in CASES]
    rc = 0
    for case in CASES:
        if case[0] not in sel:
            continue
        log = []
        try:
            run_case(case, log)
            print(f"[{case[0]}] PASS")
        except Exception:
            rc = 1
            print(f"[{case[0]}] FAIL")
            traceback.print_exc()
        print("\n".join("   " + l for l in log))
    sys.exit(rc)


# ---------------------------------------------------------------- N1: the contig merge (combine_cluster, kthread_cb.c:570-630)
@pytest.mark.parametrize("name", refdump.golden_names())
def test_combine_matches_reference_dump(name):
    """mcb_combine on the device-resident seed contigs against the contigs the unmodified reference holds after combine_cluster
    (num_thr = 1; c_cl_* of the committed fixtures): members, their order, consensus strings, number of iterations."""
    reads, meta, d, _ = refdump.load_golden(name)
    env = meta["env"]
    p = api.resolve_params(meta["L"], k=int(env.get("MC_K", 0)), e=int(env.get("MC_E", 0)), w=int(env.get("MC_W", 0)), m=int(env.get("MC_M", 0)))
    cbthr = int(env.get("MC_CBTHR", 0)) or 2 * p.diff_threshold
    c = d.clusters("c_cl")
    with api.Context(p) as ctx:
        ctx.for_reads(reads)
        ctx.for_bucket()
        r = ctx.combine(cbthr)
    assert r.iterations == d.n_idx()
    assert np.array_equal(r.cl_n, c["n"].astype(np.uint32))
    assert np.array_equal(r.cl_a, c["a"]) and np.array_equal(r.cl_a_off, c["a_off"])
    assert np.array_equal(r.cl_ref, c["ref"]) and np.array_equal(r.cl_ref_off, c["ref_off"])


@pytest.mark.parametrize("n,L,G,seed,opts", [(40000, 100, 200000, 5, {}), (30000, 150, 150000, 6, {"w": 20}), (25000, 75, 120000, 7, {}),
                                             (30000, 100, 9000, 8, {}),          # 300x coverage: long contigs, hundreds of members, many partners
                                             (20000, 100, 100000, 9, {"k": 24, "e": 6, "m": 3, "w": 10})])
def test_combine_matches_oracle(n, L, G, seed, opts):
    import oracle_lib as O
    reads = synth.make_reads(n, L, G, seed=seed, special=0.005)
    S = O.Stage1(O.resolve_params(L, **opts), reads)
    p = api.resolve_params(L, **opts)
    for cbthr in (2 * p.diff_threshold, 3):
        want = S.combine(cbthr)
        with api.Context(p) as ctx:
            ctx.for_reads(reads)
            ctx.for_bucket()
            r = ctx.combine(cbthr)
        assert r.iterations == want["iterations"], (r.iterations, want["iterations"])
        assert np.array_equal(r.cl_n, want["cl_n"]) and np.array_equal(r.cl_a, want["cl_a"]), "members differ"
        assert np.array_equal(r.cl_ref, want["cl_ref"]) and np.array_equal(r.cl_ref_off, want["cl_ref_off"]), "consensus strings differ"
        assert r.n_merges == int(want["iter_merges"].sum())
    S.close()


@pytest.mark.parametrize("knob", [{}, {"MCB_CB_SLOTCAP": "2"}, {"MCB_CB_TWOPASS": "1"}], ids=["slots", "slot_overflow_fallback", "two_pass"])
def test_combine_sketch_paths_agree_on_repeats(knob, monkeypatch):
    """The merge sketches its contigs in one pass that parks the tuples of every stretch in a fixed slot; a stretch that needs more
    (forced here with a two-tuple slot) falls back to the counting + emitting passes.  Tandem repeats put identical k-mers into one
    window (the tie scans of mm_sketch_lh_ori), deep coverage makes long contigs of many stretches."""
    import oracle_lib as O
    for k, v in knob.items():
        monkeypatch.setenv(k, v)
    L = 100
    genome = _tandem_genome(61, 12000)
    reads = synth.make_reads(30000, L, len(genome), seed=61, sub_rate=0.004, genome=genome)
    S = O.Stage1(O.resolve_params(L), reads)
    p = api.resolve_params(L)
    want = S.combine(2 * p.diff_threshold)
    with api.Context(p) as ctx:
        ctx.for_reads(reads)
        ctx.for_bucket()
        r = ctx.combine(2 * p.diff_threshold)
        r2 = (ctx.for_reads(reads), ctx.for_bucket(), ctx.combine(2 * p.diff_threshold))[2]      # a second run on the same context starts a fresh count arena
    for got in (r, r2):
        assert got.iterations == want["iterations"]
        assert np.array_equal(got.cl_n, want["cl_n"]) and np.array_equal(got.cl_a, want["cl_a"]), "members differ"
        assert np.array_equal(got.cl_ref, want["cl_ref"]) and np.array_equal(got.cl_ref_off, want["cl_ref_off"]), "consensus strings differ"
    S.close()
