"""N3: the FASTQ -> packed-reads parser of libminicom_b200.so (csrc/mcb_fastq.cu, host code) against
  1. the oracle's restatement of kseq_read / bseq_read (oracle/mc_oracle.c: mco_kseq_all, mco_pack_row), and
  2. where oracle/_ref/libmcref_units.so exists, the reference's own bseq_read (bseq.c compiled from /root/reference/src),
on well-formed and awkward files: comments, CR LF, multi-line records, FASTA records, blank lines, junk before the first header,
no final newline, truncated records, gzip.  No GPU needed (the parser is host code; the rows are simply not page-locked here)."""
import ctypes as C
import gzip
import os
import tempfile

import numpy as np
import pytest

import oracle_lib as O
import refdump
from minicom_b200 import api, synth


def _ref_units():
    p = os.path.join(refdump.REF_DIR, "libmcref_units.so")
    if not os.path.exists(p):
        return None
    R = C.CDLL(p)
    if not hasattr(R, "ref_bseq_read"):
        return None
    R.ref_bseq_read.restype = C.c_int64
    R.ref_bseq_read.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int64]
    return R


def _reference_rows(data: bytes, L: int, gz=False):
    R = _ref_units()
    if R is None:
        return None
    with tempfile.NamedTemporaryFile(suffix=".fq.gz" if gz else ".fq", delete=False) as f:
        f.write(gzip.compress(data) if gz else data)
    try:
        rows = np.zeros((len(data) // L + 1, L), dtype=np.uint8)
        n = R.ref_bseq_read(f.name.encode(), L, rows.ctypes.data, len(rows))
        assert n >= 0
        return rows[:n]
    finally:
        os.unlink(f.name)


def _ours(data: bytes, L: int, threads=2):
    rs = api.ReadSet(L)
    rows = rs.add_fastq_buffer(data, threads, want_ascii=True)
    packed, nrid, nmask = (a.copy() for a in rs.arrays())
    rs.close()
    return rows, packed, nrid, nmask


def _check(data: bytes, L: int, with_reference=True):
    want = O.kseq_all(data)
    if any(len(s) != L for s in want):                      # bseq.c:54-57: the reference exits
        with pytest.raises(api.McbError, match="Length of reads are different"):
            _ours(data, L)
        return None
    want_rows = np.frombuffer(b"".join(want), dtype=np.uint8).reshape(len(want), L)
    rows, packed, nrid, nmask = _ours(data, L)
    assert rows.shape == want_rows.shape and np.array_equal(rows, want_rows)
    op, om, st = O.pack_rows(want_rows)
    assert (st >= 0).all()
    assert np.array_equal(packed, op)
    assert np.array_equal(nrid, np.nonzero(st == 1)[0]) and np.array_equal(nmask, om[st == 1])
    if with_reference:
        ref = _reference_rows(data, L)
        if ref is not None:
            assert np.array_equal(ref, want_rows), "oracle restatement of kseq_read differs from the reference's bseq_read"
    return rows


def _seqs(n, L, seed, special=0.05):
    return [bytes(r) for r in synth.make_reads(n, L, max(1000, 5 * n), seed=seed, special=special)]


def test_plain_four_line_fastq():
    L = 100
    s = _seqs(300, L, 1)
    data = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, q, b"I" * L) for i, q in enumerate(s))
    rows = _check(data, L)
    assert len(rows) == 300


def test_comments_crlf_blank_lines_and_no_final_newline():
    L = 75
    s = _seqs(40, L, 2)
    recs = []
    for i, q in enumerate(s):
        nl = b"\r\n" if i % 3 == 0 else b"\n"
        recs.append(b"@read.%d len=%d\tx" % (i, L) + nl + (b"\n" if i % 5 == 0 else b"") + q + nl + b"+r%d" % i + nl + b"#" * L + nl + (b"\n\n" if i % 7 == 0 else b""))
    data = b"".join(recs)
    assert len(_check(data, L)) == 40
    assert len(_check(data.rstrip(b"\r\n"), L)) == 40
    assert len(_check(b"junk line\nmore junk\n" + data, L)) == 40


def test_multi_line_records_and_fasta():
    L = 100
    s = _seqs(30, L, 3)
    recs = []
    for i, q in enumerate(s):
        if i % 3 == 0:          # FASTQ, sequence and quality over several lines
            recs.append(b"@m%d\n" % i + q[:37] + b"\n" + q[37:80] + b"\n\n" + q[80:] + b"\n+\n" + b"I" * 50 + b"\n" + b"I" * 50 + b"\n")
        elif i % 3 == 1:        # FASTA
            recs.append(b">f%d desc\n" % i + q[:60] + b"\n" + q[60:] + b"\n")
        else:
            recs.append(b"@s%d\n%s\n+\n%s\n" % (i, q, b"5" * L))
    data = b"".join(recs)
    assert len(_check(data, L)) == 30
    # quality characters '@' and '+' at line starts must not start records
    tricky = b"@t0\n" + s[0] + b"\n+\n" + b"@" + b"I" * (L - 1) + b"\n@t1\n" + s[1] + b"\n+\n" + b"+" * L + b"\n"
    assert len(_check(tricky, L)) == 2


def test_truncated_and_malformed_records_end_the_file_like_kseq():
    L = 100
    s = _seqs(6, L, 4)
    ok = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, q, b"I" * L) for i, q in enumerate(s[:4]))
    assert len(_check(ok + b"@r4\n" + s[4] + b"\n+\n", L)) == 4                           # no quality line: record dropped
    assert len(_check(ok + b"@r4\n" + s[4] + b"\n+", L)) == 4
    assert len(_check(ok + b"@r4\n" + s[4] + b"\n+\n" + b"I" * 40 + b"\n", L)) == 4           # short quality at end of file
    # quality longer than the sequence: kseq_read fails there and bseq_read stops, later records are never read
    assert len(_check(ok[:2 * (2 * L + 9)] + b"@bad\n" + s[4] + b"\n+\n" + b"I" * (L + 3) + b"\n" + ok, L)) == 2
    assert len(_check(ok + b"@r4\n" + s[4], L)) == 5                                       # header + sequence, then end of file (FASTA-style)
    assert len(_check(b"", L)) == 0
    assert len(_check(b"no header at all\n", L)) == 0


def test_wrong_length_and_bad_characters_are_refused():
    L = 100
    s = _seqs(5, L, 5, special=0.0)
    data = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, q if i != 3 else q[:-1], b"I" * (L if i != 3 else L - 1)) for i, q in enumerate(s))
    assert _check(data, L) is None
    bad = bytearray(b"".join(b"@r%d\n%s\n+\n%s\n" % (i, q, b"I" * L) for i, q in enumerate(s)))
    bad[bad.index(b"\n") + 10] = ord("a")                   # lower case: process_reads counts only upper-case ACGTN
    with pytest.raises(api.McbError, match="characters other than"):
        _ours(bytes(bad), L)
    rs = api.ReadSet(L)
    rows = np.frombuffer(b"".join(s), dtype=np.uint8).reshape(5, L).copy()
    rows[2, 50] = ord("R")
    with pytest.raises(api.McbError, match="read 2 contains"):
        rs.add_rows(rows)
    assert len(rs) == 0


@pytest.mark.parametrize("L", [36, 64, 101, 150, 256])
def test_packing_matches_the_oracle_layout(L):
    reads = synth.make_reads(2000, L, 20000, seed=L, special=0.1)
    rs = api.ReadSet(L)
    rs.add_rows(reads[:700], 3)
    rs.add_rows(reads[700:], 2)                             # appending keeps rows, ids of the N table are global
    packed, nrid, nmask = rs.arrays()
    op, om, st = O.pack_rows(reads)
    assert np.array_equal(packed, op)
    assert np.array_equal(nrid, np.nonzero(st == 1)[0]) and np.array_equal(nmask, om[st == 1])
    assert (st == 1).sum() > 0


def test_gzip_and_plain_files_and_appending_a_second_file():
    L = 101
    a, b = _seqs(500, L, 6), _seqs(400, L, 7)
    fa = b"".join(b"@a%d/1\n%s\n+\n%s\n" % (i, q, b"I" * L) for i, q in enumerate(a))
    fb = b"".join(b"@b%d/2\n%s\n+\n%s\n" % (i, q, b"I" * L) for i, q in enumerate(b))
    with tempfile.TemporaryDirectory() as d:
        p1, p2 = os.path.join(d, "r1.fastq"), os.path.join(d, "r2.fastq.gz")
        open(p1, "wb").write(fa)
        open(p2, "wb").write(gzip.compress(fb))
        rs = api.ReadSet(L)
        r1 = rs.add_fastq(p1, 2, want_ascii=True)
        r2 = rs.add_fastq(p2, 2, want_ascii=True)             # bseq_read_second (-2): appended behind the first file
    want = np.frombuffer(b"".join(a + b), dtype=np.uint8).reshape(900, L)
    assert np.array_equal(np.vstack([r1, r2]), want)
    op, om, st = O.pack_rows(want)
    packed, nrid, nmask = rs.arrays()
    assert np.array_equal(packed, op) and np.array_equal(nrid, np.nonzero(st == 1)[0]) and np.array_equal(nmask, om[st == 1])
    ref = _reference_rows(fb, L, gz=True)
    if ref is not None:
        assert np.array_equal(ref, want[500:])
    with pytest.raises(api.McbError, match="cannot open"):
        rs.add_fastq("/nonexistent/file.fastq")


def test_random_formatting_fuzz():
    rng = np.random.default_rng(99)
    L = 40
    for trial in range(60):
        s = _seqs(int(rng.integers(1, 12)), L, 1000 + trial)
        out = []
        for i, q in enumerate(s):
            nl = b"\r\n" if rng.random() < 0.2 else b"\n"
            cut = sorted(rng.integers(1, L, size=int(rng.integers(0, 3))).tolist())
            parts = [q[a:b] for a, b in zip([0] + cut, cut + [L])]
            seq = nl.join(p for p in parts if p) + nl
            kind = rng.random()
            if kind < 0.25:
                out.append(b">f%d" % i + nl + seq)
            else:
                qual = bytes(rng.integers(33, 74, size=L).astype(np.uint8))
                if rng.random() < 0.1:
                    qual = qual[:int(rng.integers(1, L))]                      # malformed
                qcut = int(rng.integers(1, L)) if rng.random() < 0.3 else L
                out.append(b"@q%d c" % i + nl + seq + b"+" + nl + qual[:qcut] + (nl + qual[qcut:] if qcut < len(qual) else b"") + nl)
            if rng.random() < 0.2:
                out.append(nl)
        data = b"".join(out)
        if rng.random() < 0.3:
            data = data.rstrip(b"\r\n")
        _check(data, L, with_reference=trial % 4 == 0)
