"""GPU end-to-end: the drop-in executable (the reference's kept host objects + dropin/mcb_dropin.cpp +
libminicom_b200.so) must write a byte-identical pre-back-end directory to the reference built with num_thr=1, and the
directory must round-trip through the reference's own decompressor."""
import filecmp
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

import refdump

pytestmark = pytest.mark.gpu

CASES = [
    ("20k_special", 20000, 100, 100000, 3, 0.01, {}),
    ("60k_plain", 60000, 100, 300000, 5, 0.0, {}),
    ("30k_k25_e6", 30000, 100, 150000, 7, 0.005, {"MC_K": 25, "MC_E": 6, "MC_M": 4, "MC_W": 12, "MC_S": 3, "MC_STEP": 3, "MC_EMAX": 30}),
    # BASELINE.json configs[0]: 1M x 100 bp, 5 Mbp genome, defaults (index buckets exceed 64 tuples here: unstable-sort order matters)
    ("C1_1M", 1000000, 100, 5000000, 1, 0.0, {}),
]


def compare_dirs(a, b):
    fa, fb = sorted(os.listdir(a)), sorted(os.listdir(b))
    assert fa == fb, f"file sets differ: {set(fa) ^ set(fb)}"
    bad = [f for f in fa if not filecmp.cmp(os.path.join(a, f), os.path.join(b, f), shallow=False)]
    assert not bad, f"files differ: {bad}"
    return fa


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_dropin_directory_is_byte_identical(case):
    name, n, L, G, seed, special, env = case
    exe = refdump.dropin_binary(L, "sg")
    if not os.path.exists(exe):
        pytest.fail(f"{exe} not built (dropin/build_dropin.sh {L} sg)")
    big = n > 200000
    if big and not refdump.have_cached(n, L, G, seed, special, "sg", env, dump=False) and not os.environ.get("MCB_RUN_BIG_REFERENCE"):
        pytest.skip("reference output for the large case is not cached (takes ~1 min of CPU: set MCB_RUN_BIG_REFERENCE=1)")
    reads, got = refdump.cached_reference(n, L, G, seed, special, "sg", env, dump=not big)
    ref_out = got if big else os.path.join(os.path.dirname(got.path), "out")
    with tempfile.TemporaryDirectory() as wd:
        r = refdump.run_reference(reads, wd, mode="sg", env_opts=env, dump=False, exe=exe)
        files = compare_dirs(ref_out, r["out"])
        rt = refdump.roundtrip(r["out"], wd, "sg", reads)                   # `minicom -d` gives the reads back (multiset)
        print(name, "identical files:", len(files), "round trip:", rt, "timing:", r["timing"])


@pytest.mark.parametrize("name", refdump.golden_names())
def test_dropin_matches_golden_output_directory(name):
    """The committed fixtures hold the reference's output directory (file name -> bytes): sg mode at L=75/100/150, an option
    sweep and ORDER mode.  Needs neither /root/reference nor the reference binary."""
    reads, meta, _, want = refdump.load_golden(name)
    exe = refdump.dropin_binary(meta["L"], meta["mode"])
    if not os.path.exists(exe):
        pytest.fail(f"{exe} not built (dropin/build_dropin.sh {meta['L']} {meta['mode']})")
    with tempfile.TemporaryDirectory() as wd:
        r = refdump.run_reference(reads, wd, mode=meta["mode"], env_opts=meta["env"], dump=False, exe=exe)
        got = {f: open(os.path.join(r["out"], f), "rb").read() for f in os.listdir(r["out"])}
        rt = refdump.roundtrip(r["out"], wd, meta["mode"], reads)          # exact order under -p, multiset otherwise
    assert sorted(got) == sorted(want), f"file sets differ: {set(got) ^ set(want)}"
    bad = [f for f in want if got[f] != want[f]]
    assert not bad, f"files differ: {bad}"
    assert rt["reads"] == len(reads)


def test_dropin_paired_end_mode():
    """-1/-2: two files, _PE build (kthread_dump_pe.c on the host side).  The reference binary is run here (it is a plain
    x86-64 executable that travels with the snapshot); skipped when it was never built."""
    L, n = 101, 12000
    ref, exe = refdump.ref_binary(L, "pe"), refdump.dropin_binary(L, "pe")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/minicom_ref_L101_pe not built")
    if not os.path.exists(exe):
        pytest.fail(f"{exe} not built")
    from minicom_b200 import synth
    genome = synth.make_genome(60000, 51)
    r1 = synth.make_reads(n, L, 60000, seed=51, special=0.01, genome=genome)
    r2 = synth.make_reads(n, L, 60000, seed=52, special=0.01, genome=genome)
    with tempfile.TemporaryDirectory() as wa, tempfile.TemporaryDirectory() as wb:
        a = refdump.run_reference(r1, wa, mode="pe", dump=False, reads2=r2)
        b = refdump.run_reference(r1, wb, mode="pe", dump=False, reads2=r2, exe=exe)
        files = compare_dirs(a["out"], b["out"])
        rt = refdump.roundtrip(b["out"], wb, "pe", r1, r2)
        print("PE identical files:", len(files), "round trip:", rt)
