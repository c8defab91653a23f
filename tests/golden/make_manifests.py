#!/usr/bin/env python
"""TEST INFRASTRUCTURE: digests of the UNMODIFIED reference's results at the BASELINE.json sizes -> tests/golden/manifest_<cfg>.json.

Each manifest = one run of oracle/_ref/minicom_ref_L<readlen>_<mode> (built from /root/reference/src by oracle/ref/build_ref.sh)
with num_thr = 1 — the only deterministic configuration of the reference — on the seeded synthetic read set of a
BASELINE config, with the link-time state dumps of oracle/ref/mcref_wrap.cpp.  The raw results are gigabytes; the manifest keeps
the SHA-256 of every file of the pre-back-end output directory and of every canonical state array (minicom_b200/parity.py), so the
GPU tests and bench.py can prove byte-identity at the stated sizes on a box where neither the reference sources nor its outputs
exist.  A 10 M-read config takes ~20 min of one CPU core here, C4 (50 M x 150 bp) a few hours.

usage:  python tests/golden/make_manifests.py C2 [C3 C4s C4 ...]      (work directory: $MCB_MANIFEST_DIR or /tmp/mcb_manifest)
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refdump  # noqa: E402
from minicom_b200 import parity, synth  # noqa: E402

# name -> reads per file, read length, genome, seed, mode, reference options.  The read sets are the ones bench.py synthesizes
# (seed 1 on rank 0); C3 = two files, file 2 from seed + 1000 over the same genome.
CONFIGS = {
    "C1o": dict(n=1_000_000, L=100, G=5_000_000, seed=1, mode="order", env={}),
    "C2": dict(n=10_000_000, L=100, G=50_000_000, seed=1, mode="order", env={}),
    "C3": dict(n=5_000_000, L=101, G=50_000_000, seed=1, mode="pe", env={}),
    "C3s": dict(n=500_000, L=101, G=5_000_000, seed=1, mode="pe", env={}),
    "C4t": dict(n=1_000_000, L=150, G=7_500_000, seed=1, mode="sg", env={"MC_W": 20, "MC_S": 4, "MC_EMAX": 40, "MC_STEP": 2}),
    "C4s": dict(n=10_000_000, L=150, G=75_000_000, seed=1, mode="sg", env={"MC_W": 20, "MC_S": 4, "MC_EMAX": 40, "MC_STEP": 2}),
    "C4": dict(n=50_000_000, L=150, G=375_000_000, seed=1, mode="sg", env={"MC_W": 20, "MC_S": 4, "MC_EMAX": 40, "MC_STEP": 2}),
}
WORK = os.environ.get("MCB_MANIFEST_DIR", "/tmp/mcb_manifest")


def make_reads(c):
    if c["mode"] == "pe":
        genome = synth.make_genome(c["G"], c["seed"])
        return synth.make_reads(c["n"], c["L"], c["G"], seed=c["seed"], genome=genome), synth.make_reads(c["n"], c["L"], c["G"], seed=c["seed"] + 1000, genome=genome)
    return synth.make_reads(c["n"], c["L"], c["G"], seed=c["seed"]), None


def run(name):
    c = CONFIGS[name]
    wd = os.path.join(WORK, name)
    if not os.path.exists(os.path.join(wd, "ok")):
        reads, reads2 = make_reads(c)
        t0 = time.time()
        refdump.run_reference(reads, wd, mode=c["mode"], env_opts=c["env"], threads=1, dump="noseqs", reads2=reads2)
        print(f"{name}: reference ran {time.time() - t0:.0f}s", flush=True)
        with open(os.path.join(wd, "ok"), "w") as f:
            f.write("ok")
    return wd


def file_sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def state_digests(d: refdump.Dump, n_reads: int) -> dict:
    out = {}
    cls = np.zeros(n_reads, dtype=np.uint8)
    for code, fn in ((1, "r_allA.u32"), (2, "r_allT.u32"), (3, "r_allN.u32"), (4, "r_fpA.u32"), (5, "r_fpT.u32"), (6, "r_fpN.u32"), (7, "r_Nfile.u32")):
        cls[d.arr(fn).astype(np.int64)] = code
    t = d.arr("r_B0_xy.u64").reshape(-1, 2)
    t = t[np.argsort(t[:, 1] >> np.uint64(32), kind="stable")]                  # one tuple per sketched read: read-id order
    out.update({"cls": parity.digest(cls), "tuples": parity.digest(t)})
    b = d.clusters("b_cl")
    out.update({"seed.cl_n": parity.digest(b["n"].astype(np.uint32)), "seed.cl_a": parity.digest(b["a"]), "seed.cl_ref": parity.digest(b["ref"]),
                "seed.cl_reflen": parity.digest(np.diff(b["ref_off"].astype(np.int64)).astype(np.uint64)), "sg": parity.digest(d.arr("b_sg.u32")),
                "seed.mi": parity.digest(d.arr("b_mi_xy.u64").reshape(-1, 2))})
    for j in range(d.n_idx()):
        x = d.arr(f"i{j}_in_xy.u64").reshape(-1, 2)[:, 0]
        order = np.lexsort((x, x & parity.NB_MASK))                             # bucket-major, keys ascending = the dump's key order
        xs = x[order]
        head = np.ones(len(xs), dtype=bool)
        head[1:] = xs[1:] != xs[:-1]
        keys = xs[head]
        cnt = np.diff(np.concatenate([np.nonzero(head)[0], [len(xs)]])).astype(np.int64)
        raw = d.arr(f"i{j}_post.u64")
        rec = np.concatenate([[0], np.cumsum(2 + cnt)])[:-1]
        assert len(raw) == int((2 + cnt).sum()) and np.array_equal(raw[rec], keys) and np.array_equal(raw[rec + 1], cnt.astype(np.uint64)), f"index dump {j} has an unexpected layout"
        body = np.ones(len(raw), dtype=bool)
        body[rec] = False
        body[rec + 1] = False
        out.update({f"idx{j}.keys": parity.digest(keys), f"idx{j}.cnt": parity.digest(cnt.astype(np.uint32)), f"idx{j}.post": parity.digest(raw[body])})
    c = d.clusters("c_cl")
    out.update(parity.contig_digests(c["n"], c["a"], c["ref"], np.diff(c["ref_off"].astype(np.int64)).astype(np.uint64)))
    for j in range(d.n_realign()):
        out.update({f"h{j}.sg": parity.digest(d.arr(f"h{j}_sg.u32")), f"h{j}.app_cnt": parity.digest(d.arr(f"h{j}_app_cnt.u64")), f"h{j}.app_y": parity.digest(d.arr(f"h{j}_app_y.u64")),
                    f"h{j}.flag": parity.digest(d.arr(f"h{j}_flag.u8")), f"h{j}.fpA": parity.digest(d.arr(f"h{j}_fpA.u32")), f"h{j}.fpT": parity.digest(d.arr(f"h{j}_fpT.u32"))})
    return out


def main():
    for name in sys.argv[1:] or ["C2"]:
        c = CONFIGS[name]
        wd = run(name)
        d = refdump.Dump(os.path.join(wd, "dump"))
        with open(os.path.join(wd, "timing.json")) as f:
            timing = json.load(f)
        n_total = c["n"] * (2 if c["mode"] == "pe" else 1)
        man = {"config": dict(c, name=name), "reference": {"num_thr": 1, "binary": f"oracle/_ref/minicom_ref_L{c['L']}_{c['mode']}", "timing_s": timing},
               "counts": {"reads": n_total, "seed_contigs": int(len(d.arr("b_cl_n.u64"))), "singles_stage1": int(len(d.arr("b_sg.u32"))), "contigs": int(len(d.arr("c_cl_n.u64"))),
                          "contig_bases": int(len(d.arr("c_cl_ref.u8"))), "index_builds": d.n_idx(), "realign_rounds": d.n_realign(),
                          "claims": [int(len(d.arr(f"h{j}_app_y.u64"))) for j in range(d.n_realign())]},
               "out": {f: file_sha(os.path.join(wd, "out", f)) for f in sorted(os.listdir(os.path.join(wd, "out")))},
               "state": state_digests(d, n_total)}
        path = os.path.join(HERE, f"manifest_{name}.json")
        with open(path, "w") as f:
            json.dump(man, f, indent=1, sort_keys=True)
        print(f"{name}: {len(man['out'])} files, {len(man['state'])} state arrays -> {path}", flush=True)


if __name__ == "__main__":
    main()
