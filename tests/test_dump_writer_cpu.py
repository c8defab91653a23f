"""N2, the marshalling half (dropin/mcb_dump_writer.h: which file gets which bytes), checked on the CPU against the unmodified
reference: oracle/_ref/minicom_ref_L<L>_<mode>_dumpcheck is the reference with its kt_dump_[pe_]for replaced by the writer + the
oracle's print_encode (oracle/ref/mcref_dumpcheck.cpp); it must write the directory the plain reference writes, byte for byte,
in default, order-preserving and paired-end mode, with one and with several dump threads.  (The product's drop-in uses the same
writer with the device encoder: tests/test_gpu_dropin.py, tests/test_gpu_manifest.py.)"""
import filecmp
import os
import tempfile

import pytest

import refdump
from minicom_b200 import synth


def _compare(a, b):
    fa, fb = sorted(os.listdir(a)), sorted(os.listdir(b))
    assert fa == fb, f"file sets differ: {set(fa) ^ set(fb)}"
    bad = [f for f in fa if not filecmp.cmp(os.path.join(a, f), os.path.join(b, f), shallow=False)]
    assert not bad, f"files differ: {bad}"
    return fa


@pytest.mark.parametrize("L,mode,threads", [(100, "sg", 1), (100, "order", 1), (101, "pe", 1), (100, "order", 3)])
def test_dump_writer_reproduces_the_reference_directory(L, mode, threads):
    ref = refdump.ref_binary(L, mode)
    chk = ref + "_dumpcheck"
    if not (os.path.exists(ref) and os.path.exists(chk)):
        pytest.skip("oracle/_ref/*_dumpcheck not built (needs /root/reference: oracle/ref/build_ref.sh <L> <mode> dumpcheck)")
    n, G = 20000, 60000
    genome = synth.make_genome(G, 71)
    r1 = synth.make_reads(n, L, G, seed=71, special=0.02, genome=genome)
    r2 = synth.make_reads(n, L, G, seed=72, special=0.02, genome=genome) if mode == "pe" else None
    with tempfile.TemporaryDirectory() as wa, tempfile.TemporaryDirectory() as wb:
        a = refdump.run_reference(r1, wa, mode=mode, threads=threads, dump=False, reads2=r2)
        b = refdump.run_reference(r1, wb, mode=mode, threads=threads, dump=False, reads2=r2, exe=chk)
        if threads == 1:
            files = _compare(a["out"], b["out"])
            assert any(f.startswith("dif_char.txt") for f in files) and os.path.getsize(os.path.join(b["out"], "dif_char.txt.0")) > 1000
        else:
            # the reference's contig stages are not deterministic with several threads: only the file set and the
            # round trip can be compared
            assert sorted(os.listdir(a["out"])) == sorted(os.listdir(b["out"]))
            rt = refdump.roundtrip(b["out"], wb, mode, r1, r2, threads=threads)
            assert rt["reads"] == n
