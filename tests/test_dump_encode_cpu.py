"""N2, the part restated so far: print_encode's per-read diff encoding (kthread_dump.c:66-118).  The oracle's
mco_print_encode is pinned against the reference's own dif_char.txt: the committed fixtures hold the reference's contigs after the
merge, the reads every realign round appended and the output directory, so the final clusters, their dump order (cmpcluster2 /
cmpcluster3 via the stable qsort of glibc) and therefore every line of dif_char.txt.0 can be rebuilt and compared.  No GPU."""
import numpy as np
import pytest

import oracle_lib as O
import refdump


def final_clusters(reads, meta, d):
    """(ref string, members in dump order) of every cluster the reference dumps, rebuilt from the state dumps"""
    c = d.clusters("c_cl")
    n_c = len(c["n"])
    members = [list(c["a"][int(c["a_off"][i]):int(c["a_off"][i + 1])]) for i in range(n_c)]
    key2 = lambda y: ((int(y) & 0xFFFFFFFF) >> 1, int(y) & 1)                       # cmpcluster2 (kthread_cb.c:54-69)
    key3 = lambda y: ((int(y) & 0xFFFFFFFF) >> 1, int(y) >> 32)                     # cmpcluster3 (ORDER / _PE, :72-85)
    for j in range(d.n_realign()):
        for m in members:
            m.sort(key=key2)                                                        # realign_hash_search re-sorts before it appends (kthread_hash_realign.c:318)
        cnt, ys = d.arr(f"h{j}_app_cnt.u64"), d.arr(f"h{j}_app_y.u64")
        o = 0
        for i in range(n_c):
            members[i].extend(ys[o:o + int(cnt[i])])
            o += int(cnt[i])
    final = key3 if meta["mode"] in ("order", "pe") else key2
    for m in members:
        m.sort(key=final)                                                           # print_encode (kthread_dump.c:34 / :127)
    refs = [c["ref"][int(c["ref_off"][i]):int(c["ref_off"][i + 1])].tobytes() for i in range(n_c)]
    return refs, members


def original_reads(reads, d):
    """the reads as the dump stage sees them: N put back (the input rows still hold them)"""
    return reads


@pytest.mark.parametrize("name", refdump.golden_names())
def test_oracle_print_encode_reproduces_the_reference_dif_char_file(name):
    reads, meta, d, out = refdump.load_golden(name)
    L = meta["L"]
    refs, members = final_clusters(reads, meta, d)
    lines = []
    for ref, mem in zip(refs, members):
        for y in mem:
            rid, pos, direction = int(y) >> 32, (int(y) & 0xFFFFFFFF) >> 1, int(y) & 1
            lines.append(O.print_encode(reads[rid].tobytes(), direction, ref[pos:pos + L]))
    want = out["dif_char.txt.0"]
    got = b"".join(l + b"\n" for l in lines)
    assert len(lines) > 1000
    assert got == want, "oracle restatement of print_encode differs from the reference's dif_char.txt"


def test_print_encode_known_cases():
    L = 12
    ref = b"ACGTACGTACGT"
    assert O.print_encode(ref, 0, ref) == b"0"
    assert O.print_encode(b"TCGTACGTACGT", 0, ref) == b"T"                          # mismatch first: nothing before it, trailing run dropped
    assert O.print_encode(b"ACGTACGTACGA", 0, ref) == b"11A"                        # run of 11, then the character
    assert O.print_encode(b"AGGTACGTACGT", 0, ref) == b"AG"                         # run of 1 is copied, not counted
    assert O.print_encode(b"ACNTACGNACGT", 0, ref) == b"2N4N"                       # N never equals the consensus
    assert O.print_encode(b"ACGTACGTACGT", 1, ref) == b"0"                          # ACGTACGTACGT is its own reverse complement
    assert O.print_encode(b"ACGTACGTACGA", 1, ref) == b"T"                          # rc = TCGTACGTACGT
