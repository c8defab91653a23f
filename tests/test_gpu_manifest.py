"""GPU end-to-end at the BASELINE.json sizes: the drop-in executable (reference host objects + dropin/mcb_dropin.cpp +
libminicom_b200.so), run with num_thr = 1, must write a pre-back-end directory whose every file has the SHA-256 the
single-threaded reference produced for the same seeded reads (tests/golden/manifest_<cfg>.json, made by
tests/golden/make_manifests.py where the reference exists), and the directory must round-trip through the reference's
decompressor (`minicom -d`).  Needs neither /root/reference nor the reference's outputs on the GPU box.

  C1o / C3s / C4t   1 M-read versions of configs 2, 3, 4 (order-preserving; paired-end; 150 bp with -w/-s/-E/-S)      always
  C2 / C3 / C4s     configs 2 and 3 at their stated sizes (10 M x 100 bp -p; 2 x 5 M x 101 bp -1/-2) and config 4's options
                    at 10 M x 150 bp — minutes each, most of it the reference's own single-threaded host stages            unless MCB_SKIP_BIG=1
  C4                config 4 at 50 M x 150 bp: run by hand (MCB_RUN_C4=1), ~15 min of host time around the GPU calls
"""
import hashlib
import json
import os
import sys
import tempfile
import time

import pytest

import refdump

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(refdump.ROOT, "tests", "golden")
sys.path.insert(0, GOLDEN)


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def _cases():
    out = []
    for name in ("C1o", "C3s", "C4t", "C2", "C3", "C4s", "C4"):
        if os.path.exists(os.path.join(GOLDEN, f"manifest_{name}.json")):
            out.append(name)
    return out


@pytest.mark.parametrize("name", _cases())
def test_dropin_directory_matches_reference_manifest(name):
    with open(os.path.join(GOLDEN, f"manifest_{name}.json")) as f:
        man = json.load(f)
    c = man["config"]
    big = c["n"] * (2 if c["mode"] == "pe" else 1) > 2_000_000
    if big and os.environ.get("MCB_SKIP_BIG"):
        pytest.skip("MCB_SKIP_BIG set")
    if name == "C4" and not os.environ.get("MCB_RUN_C4"):
        pytest.skip("50 M x 150 bp: set MCB_RUN_C4=1 (about 15 minutes, mostly the reference's single-threaded host stages)")
    exe = refdump.dropin_binary(c["L"], c["mode"])
    if not os.path.exists(exe):
        pytest.fail(f"{exe} not built (dropin/build_dropin.sh {c['L']} {c['mode']})")
    import make_manifests
    t0 = time.time()
    reads, reads2 = make_manifests.make_reads(c)
    with tempfile.TemporaryDirectory() as wd:
        r = refdump.run_reference(reads, wd, mode=c["mode"], env_opts=c["env"], threads=1, dump=False, reads2=reads2, exe=exe)
        got = {f: _sha(os.path.join(r["out"], f)) for f in sorted(os.listdir(r["out"]))}
        assert sorted(got) == sorted(man["out"]), f"file sets differ: {set(got) ^ set(man['out'])}"
        bad = [f for f in got if got[f] != man["out"][f]]
        assert not bad, f"{name}: files differ from the reference's: {bad}"
        rt = refdump.roundtrip(r["out"], wd, c["mode"], reads, reads2)
    print(f"{name}: {len(got)} files byte-identical to the reference (num_thr=1); round trip {rt}; {time.time() - t0:.0f}s; front end "
          f"{sum(r['timing'][k] for k in ('kt_for_reads', 'kt_for_bucket', 'mm_idx_generation', 'realign_hash')):.2f}s inside the entry points")


@pytest.mark.parametrize("knob", [{"MCB_HOST_MERGE": "1"}, {"MCB_HOST_DUMP": "1"}, {"MCB_ASCII_READS": "1"}, {"MCB_HOST_MERGE": "1", "MCB_HOST_DUMP": "1", "MCB_ASCII_READS": "1"}],
                         ids=["host_merge", "host_dump", "ascii_reads", "round1_paths"])
def test_dropin_alternative_paths_write_the_same_directory(knob):
    """The drop-in's switches between the device and the reference's host code for the stages next to the path (contig merge N1,
    dump worker N2, FASTQ reader N3) must not change a byte: C1o (1 M reads, order-preserving) against the reference manifest with
    each of them flipped.  (MCB_HOST_MERGE=1 is what the sharded bench's recording run uses.)"""
    with open(os.path.join(GOLDEN, "manifest_C1o.json")) as f:
        man = json.load(f)
    c = man["config"]
    exe = refdump.dropin_binary(c["L"], c["mode"])
    if not os.path.exists(exe):
        pytest.fail(f"{exe} not built")
    import make_manifests
    reads, reads2 = make_manifests.make_reads(c)
    with tempfile.TemporaryDirectory() as wd:
        r = refdump.run_reference(reads, wd, mode=c["mode"], env_opts=dict(c["env"], **knob), threads=1, dump=False, reads2=reads2, exe=exe)
        got = {f: _sha(os.path.join(r["out"], f)) for f in sorted(os.listdir(r["out"]))}
    assert sorted(got) == sorted(man["out"])
    bad = [f for f in got if got[f] != man["out"][f]]
    assert not bad, f"{knob}: files differ from the reference's: {bad}"
