"""CPU tests (no GPU): the oracle restatement (oracle/mc_oracle.c) against the reference.

Pins, in order of strength:
  1. the known-answer vectors of SURVEY.md §8c (produced by linking the reference's own sketch.o),
  2. the committed golden fixtures tests/golden/*.npz = state dumps of the unmodified reference (num_thr=1),
  3. where oracle/_ref/libmcref_units.so exists (built from /root/reference/src), the reference's own hash64 /
     mm_sketch_two / mm_sketch_lh_ori / radix_sort_128x called through ctypes on random inputs.
Everything is bit-exact (integer work)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
import refdump

R0 = b"ATAGATGCAGACCTCAATGCGAGAGCCCGCTGCGCTATCATTCTGCAAGAATTGCGGACCCGCTCGAACCACTCGGGGCTCCATTTAGGACGACACCTCT"


# ---------------------------------------------------------------- 1. known answers
def test_kat_sketch_two():
    assert O.sketch_two(R0, 31, 7) == (0x004ed5956ad8b716, 0x0000000700000091)
    assert O.sketch_two(R0, 30, 7) == (0x000eefcbe0d3aa21, 0x00000007000000c3)
    assert O.sketch_two(R0, 17, 7) == (0x000000000b4b3d21, 0x00000007000000a6)
    # every second 8-mer of (ACGT)x10 is its own reverse complement: skipped without counting (sketch.c:265-272)
    assert O.sketch_two(b"ACGT" * 10, 8, 1) == (0x0000000000000248, 0x0000000100000010)


def test_kat_sketch_lh():
    want = [(0x01d65b0eb74b7ea6, 36, 0), (0x029096f0812e5494, 42, 0), (0x06994f201d9082cc, 46, 0), (0x06faf45bf305c061, 53, 0), (0x0649e8720f8749cf, 68, 0),
            (0x017b5bdaa5ab7658, 70, 1), (0x00a856f4aa12dd96, 71, 1), (0x004ed5956ad8b716, 72, 1), (0x011cd93786116d8f, 82, 0)]
    got, n = O.sketch_lh(R0, 19, 31, 0x500)
    assert n == 9
    assert [(int(x), int(y)) for x, y in got] == [(x, 0x500 << 32 | p << 1 | s) for x, p, s in want]
    # the cap truncates the output but not the count (first_mininum semantics of kthread_bucket.c:463)
    got6, n6 = O.sketch_lh(R0, 19, 31, 0x500, cap=6)
    assert n6 == 9 and np.array_equal(got6, got[:6])


def test_kat_product_host_helpers():
    """The library's host-side boundary helpers (no GPU needed) agree with the same vectors."""
    from minicom_b200 import api
    assert api.sketch_two_host(R0, 31, 7) == (0x004ed5956ad8b716, 0x0000000700000091)
    assert api.sketch_two_host(b"ACGT" * 10, 8, 1) == (0x0000000000000248, 0x0000000100000010)
    got = api.sketch_lh_host(R0, 19, 31, 0x500)
    ora, _ = O.sketch_lh(R0, 19, 31, 0x500)
    assert np.array_equal(got, ora)
    rng = np.random.default_rng(5)
    for k in (10, 17, 25, 31):
        mask = (1 << 2 * k) - 1
        for key in rng.integers(0, 1 << 62, size=50, dtype=np.uint64):
            assert api.hash64(int(key) & mask, mask) == O.hash64(int(key) & mask, mask)


# ---------------------------------------------------------------- 3. the reference itself through ctypes (when built here)
def _ref_units():
    p = os.path.join(refdump.REF_DIR, "libmcref_units.so")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref/libmcref_units.so not built (needs /root/reference; oracle/ref/build_ref.sh)")
    L = C.CDLL(p)
    L.ref_hash64.restype = C.c_uint64
    L.ref_hash64.argtypes = [C.c_uint64, C.c_uint64]
    L.ref_sketch_two.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_uint32, C.c_void_p]
    L.ref_sketch_lh_ori.restype = C.c_int64
    L.ref_sketch_lh_ori.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_int64]
    L.ref_radix_sort_128x.argtypes = [C.c_void_p, C.c_int64]
    return L


def _rand_seq(rng, n, p_n=0.0, low_complexity=False):
    alpha = b"AC" if low_complexity else b"ACGT"
    s = bytearray(alpha[i] for i in rng.integers(0, len(alpha), size=n))
    for i in np.nonzero(rng.random(n) < p_n)[0]:
        s[i] = ord("N")
    return bytes(s)


def test_reference_units_hash_and_sketch():
    R = _ref_units()
    rng = np.random.default_rng(11)
    for k in (10, 16, 17, 24, 30, 31):
        mask = (1 << 2 * k) - 1
        for key in rng.integers(0, 1 << 62, size=200, dtype=np.uint64):
            key = int(key) & mask
            assert O.hash64(key, mask) == R.ref_hash64(key, mask)
    for trial in range(300):
        n = int(rng.integers(20, 400))
        k = int(rng.integers(10, min(31, n) + 1))
        s = _rand_seq(rng, n, low_complexity=trial % 5 == 0)
        out = (C.c_uint64 * 2)()
        R.ref_sketch_two(s, n, k, trial, out)
        assert O.sketch_two(s, k, trial) == (int(out[0]), int(out[1])), (s, k)
    for trial in range(300):
        n = int(rng.integers(40, 600))
        k = int(rng.integers(10, 32))
        w = int(rng.integers(1, 50))
        s = _rand_seq(rng, n, p_n=0.01 if trial % 3 == 0 else 0.0, low_complexity=trial % 7 == 0)
        buf = np.zeros((2048, 2), dtype=np.uint64)
        cnt = R.ref_sketch_lh_ori(s, n, w, k, trial, buf.ctypes.data, 2048)
        got, n_got = O.sketch_lh(s, w, k, trial, cap=2048)
        assert n_got == cnt and np.array_equal(got, buf[:cnt]), (s, w, k)


def test_reference_units_unstable_sort_order():
    """radix_sort_128x leaves equal keys in an order that depends on its cycle-leader walk (ksort.h:131-145)."""
    R = _ref_units()
    rng = np.random.default_rng(3)
    for n, distinct in ((5, 3), (64, 8), (65, 8), (300, 20), (1000, 7), (5000, 300), (20000, 50)):
        keys = rng.integers(0, 1 << 62, size=distinct, dtype=np.uint64)
        xy = np.zeros((n, 2), dtype=np.uint64)
        xy[:, 0] = keys[rng.integers(0, distinct, size=n)]
        xy[:, 1] = np.arange(n, dtype=np.uint64)
        want = xy.copy()
        R.ref_radix_sort_128x(want.ctypes.data, n)
        got = O.radix_sort_x(xy)
        assert np.array_equal(got, want), (n, distinct)
        if n > 64:
            stable = xy[np.argsort(xy[:, 0], kind="stable")]
            assert np.array_equal(got[:, 0], stable[:, 0])


# ---------------------------------------------------------------- 2. golden fixtures
def params_for(L, env):
    return O.resolve_params(L, k=int(env.get("MC_K", 0)), e=int(env.get("MC_E", 0)), w=int(env.get("MC_W", 0)), m=int(env.get("MC_M", 0)),
                            max_rounds=int(env.get("MC_MAXROUNDS", 0)))


@pytest.mark.parametrize("name", refdump.golden_names())
def test_oracle_matches_reference_dump(name):
    reads, meta, dump, _ = refdump.load_golden(name)
    n, L, env = meta["n"], meta["L"], meta["env"]
    S = O.Stage1(params_for(L, env), reads)
    # --- kt_for_reads
    for code, fn in ((1, "r_allA.u32"), (2, "r_allT.u32"), (3, "r_allN.u32"), (4, "r_fpA.u32"), (5, "r_fpT.u32"), (6, "r_fpN.u32"), (7, "r_Nfile.u32")):
        assert np.array_equal(np.nonzero(S.cls == code)[0].astype(np.uint32), dump.arr(fn)), fn
    valid = S.tuples[:, 0] != O.MAXU64
    off, mine = O.bucket_major(S.tuples[valid])
    roff, rxy = dump.buckets("r_B0")
    assert np.array_equal(off, roff) and np.array_equal(mine, rxy), "B[0] tuples differ"
    seqs = dump.seqs()
    sk = np.nonzero(S.cls == 0)[0]
    want = np.frombuffer(b"".join(seqs[i] for i in sk), dtype=np.uint8).reshape(len(sk), L)
    assert np.array_equal(S.rows[sk], want), "N-replaced reads differ"
    # --- kt_for_bucket
    cl = dump.clusters("b_cl")
    assert np.array_equal(S.cl_n.astype(np.uint64), cl["n"])
    assert np.array_equal(S.cl_a, cl["a"]) and np.array_equal(S.cl_a_off, cl["a_off"])
    assert np.array_equal(S.cl_ref_off, cl["ref_off"]) and np.array_equal(S.cl_ref, cl["ref"])
    assert np.array_equal(S.sg, dump.arr("b_sg.u32")), "singles (order) differ"
    off, mine = O.bucket_major(S.mi)
    roff, rxy = dump.buckets("b_mi")
    assert np.array_equal(off, roff) and np.array_equal(mine, rxy), "mi[0] tuples differ"
    # --- every mm_idx_generation of the host merge
    assert dump.n_idx() >= 1
    for j in range(dump.n_idx()):
        off, xy = dump.buckets(f"i{j}_in")
        ix = O.Index(xy, off)
        post = dump.postings(j)
        assert ix.n_keys == len(post) and ix.n_post == len(xy)
        for x, ys in post:
            assert np.array_equal(ix.get(x), ys), f"index {j} key {x:#x}"
        ix.close()
    # --- every realign_hash round
    cc = dump.clusters("c_cl")
    maxsearch = 2000 if len(S.sg) <= 5000000 else 500
    assert dump.n_realign() >= 1
    for j in range(dump.n_realign()):
        sg = dump.arr(f"h{j}_sg.u32")
        thr = int(dump.arr(f"h{j}_thr.u64")[0])
        r = S.realign(sg, cc["ref"], cc["ref_off"], thr, maxsearch, int(env.get("MC_S", 0)))
        want_cnt = dump.arr(f"h{j}_app_cnt.u64")
        assert np.array_equal(np.bincount(r["claim_contig"], minlength=len(want_cnt)).astype(np.uint64), want_cnt)
        assert np.array_equal(r["claim_y"], dump.arr(f"h{j}_app_y.u64")), f"round {j}: claims / append order differ"
        assert np.array_equal(sg[r["fpA_sg"]], dump.arr(f"h{j}_fpA.u32")) and np.array_equal(sg[r["fpT_sg"]], dump.arr(f"h{j}_fpT.u32"))
        assert np.array_equal(r["flag"], dump.arr(f"h{j}_flag.u8"))
    S.close()


BIG_CASES = [
    ("20k_special", 20000, 100, 100000, 3, 0.01, {}),
    ("30k_k25_e6", 30000, 100, 150000, 7, 0.005, {"MC_K": 25, "MC_E": 6, "MC_M": 4, "MC_W": 12, "MC_S": 3, "MC_STEP": 3, "MC_EMAX": 30}),
]


@pytest.mark.parametrize("case", BIG_CASES, ids=[c[0] for c in BIG_CASES])
def test_oracle_matches_reference_run_here(case):
    """Larger sets (index buckets above 64 tuples, several realign rounds): the reference binary is run here (or its
    cached dump reused); skipped where neither exists."""
    name, n, L, G, seed, special, env = case
    if not (refdump.have_reference(L) or refdump.have_cached(n, L, G, seed, special, "sg", env)):
        pytest.skip("reference binary not built and no cached dump")
    reads, dump = refdump.cached_reference(n, L, G, seed, special, "sg", env)
    S = O.Stage1(params_for(L, env), reads)
    cl = dump.clusters("b_cl")
    assert np.array_equal(S.cl_a, cl["a"]) and np.array_equal(S.cl_ref, cl["ref"]) and np.array_equal(S.sg, dump.arr("b_sg.u32"))
    for j in range(dump.n_idx()):
        off, xy = dump.buckets(f"i{j}_in")
        ix = O.Index(xy, off)
        keys, st, post = ix.flat()
        want = dump.postings(j)
        assert np.array_equal(keys, np.array([x for x, _ in want], dtype=np.uint64))
        assert np.array_equal(post, np.concatenate([ys for _, ys in want])), f"index {j}: posting order differs"
        ix.close()
    cc = dump.clusters("c_cl")
    for j in range(dump.n_realign()):
        sg = dump.arr(f"h{j}_sg.u32")
        r = S.realign(sg, cc["ref"], cc["ref_off"], int(dump.arr(f"h{j}_thr.u64")[0]), 2000, int(env.get("MC_S", 0)))
        assert np.array_equal(r["claim_y"], dump.arr(f"h{j}_app_y.u64")), f"round {j}"
        assert np.array_equal(r["flag"], dump.arr(f"h{j}_flag.u8"))
    S.close()


def test_oracle_maxsearch_window_is_sequential():
    """Bins larger than maxsearch: only the last `maxsearch` live entries are scanned, and removals slide that
    window (kthread_hash_realign.c:388, bbhashdict.c:33-67).  20 identical singles, maxsearch 4: every window of the
    contig claims 4 of them (highest sg index first) until none is left."""
    rng = np.random.default_rng(9)
    L = 100
    core = rng.integers(0, 4, size=L)
    read = np.frombuffer(b"ACGT", dtype=np.uint8)[core]
    reads = np.tile(read, (20, 1))
    p = O.resolve_params(L)
    S = O.Stage1(p, reads)
    contig = np.concatenate([read, read])          # windows 0 and 100 match exactly
    r = S.realign(np.arange(20, dtype=np.uint32), contig, np.array([0, 2 * L], dtype=np.uint64), 4, 4)
    got = [(int(y >> np.uint64(32)), int((y & np.uint64(0xFFFFFFFF)) >> np.uint64(1))) for y in r["claim_y"]]
    # window 0, forward dictionaries l = 0..4: each probe scans the last 4 live entries of the (single) bin
    assert got == [(i, 0) for i in range(19, -1, -1)]
    assert r["flag"].all()
    S.close()


# ---------------------------------------------------------------- boundary: the C-ABI library loads and exports what the header declares
def test_cabi_exports_match_header():
    import re
    from minicom_b200 import api
    hdr = open(os.path.join(refdump.ROOT, "include", "minicom_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(mcb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = api.load_library()
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"libminicom_b200.so lacks {missing}"
    assert sorted(api.EXPORTS) == declared, "api.EXPORTS and include/minicom_b200.h disagree"


def test_resolve_params_matches_reference_rules():
    from minicom_b200 import api
    for L in (36, 69, 70, 75, 79, 80, 100, 101, 150, 250):
        for k in (0, 17, 25):
            for w in (0, 12):
                for mr in (0, 10, 35, 60):
                    a = api.resolve_params(L, k=k, e=0, w=w, m=0, max_rounds=mr)
                    o = O.resolve_params(L, k=k, w=w, max_rounds=mr)
                    assert (a.readlen, a.k, a.b, a.rw, a.first_mininum, a.diff_threshold, a.max_rounds) == \
                           (o.readlen, o.k, o.b, o.rw, o.first_mininum, o.diff_threshold, o.max_rounds)


def test_no_gpu_means_loud_failure():
    """Without a CUDA device the product refuses to work (no CPU fallback)."""
    from minicom_b200 import api
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(api.McbError):
        api.Context(api.resolve_params(100))


@pytest.mark.parametrize("name", ["g100_default", "g100_order", "g75_default", "g150_default", "g100_opts"])
def test_reference_directory_round_trips_through_decompress(name):
    """Pins the round-trip procedure itself (tests/refdump.py:roundtrip, used by the GPU drop-in tests) on the committed
    reference output directories: `minicom -d` = the reference's decompress program (minicom:383-389)."""
    import tempfile
    import refdump
    if not os.path.exists(os.path.join(refdump.REF_DIR, "decompress")):
        pytest.skip("oracle/_ref/decompress not built")
    reads, meta, _, out = refdump.load_golden(name)
    with tempfile.TemporaryDirectory() as wd:
        d = os.path.join(wd, "out")
        os.makedirs(d)
        for f, b in out.items():
            with open(os.path.join(d, f), "wb") as fh:
                fh.write(b)
        r = refdump.roundtrip(d, wd, meta["mode"], reads)
    assert r["reads"] == len(reads)
