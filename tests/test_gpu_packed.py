"""N3 on the device: mcb_for_reads_packed (kt_for_reads on the 2-bit rows the host parser makes, csrc/mcb_fastq.cu) must give
what mcb_for_reads gives on the ASCII rows — classes, N table, replacement, tuples, the packed read table — and both must equal
the oracle; kt_for_bucket downstream is unchanged.  Through the C-ABI (ctypes)."""
import numpy as np
import pytest

import oracle_lib as O
from minicom_b200 import api, synth

pytestmark = pytest.mark.gpu

CASES = [
    ("L100_special", 30000, 100, 150000, 71, 0.03, {}),
    ("L36_k12", 6000, 36, 20000, 72, 0.02, {"k": 12}),
    ("L64", 6000, 64, 30000, 73, 0.02, {}),
    ("L101", 9000, 101, 50000, 74, 0.02, {}),
    ("L150_w20", 8000, 150, 60000, 75, 0.02, {"w": 20}),
    ("L256_e8", 3000, 256, 40000, 76, 0.02, {"e": 8}),
    ("L100_k30_even", 5000, 100, 30000, 77, 0.02, {"k": 30}),
    ("n1", 1, 100, 1000, 78, 0.0, {}),
    ("n129", 129, 100, 600, 79, 0.05, {}),
    ("chunks", 2_300_000, 100, 2_000_000, 80, 0.002, {}),        # more than one upload chunk (2^21 reads)
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_packed_reads_give_the_same_results_as_ascii_rows(case):
    name, n, L, G, seed, special, opts = case
    reads = synth.make_reads(n, L, G, seed=seed, special=special)
    rs = api.ReadSet(L)
    rs.add_rows(reads, 4)
    with api.Context(api.resolve_params(L, **opts)) as a, api.Context(api.resolve_params(L, **opts)) as b:
        ra = a.for_reads(reads)
        rb = b.for_reads_packed(rs)
        for f in ("cls", "nread_rid", "nread_repl", "nread_off", "npos"):
            assert np.array_equal(getattr(ra, f), getattr(rb, f)), f
        assert ra.n_sketched == rb.n_sketched
        ta, tb = a.debug_read_tuples(n), b.debug_read_tuples(n)
        assert np.array_equal(ta, tb)
        assert np.array_equal(a.debug_unpack_reads(n), b.debug_unpack_reads(n))
        if n <= 40000:
            S = O.Stage1(O.resolve_params(L, **opts), reads)
            assert np.array_equal(rb.cls, S.cls) and np.array_equal(tb, S.tuples)
            S.close()
        ba, bb = a.for_bucket(), b.for_bucket()
        for f in ("cl_n", "cl_a", "cl_ref", "cl_ref_off", "sg", "mi_cnt"):
            assert np.array_equal(getattr(ba, f), getattr(bb, f)), f
        m = a.params.first_mininum
        valid = np.arange(m)[None, :] < ba.mi_cnt[:, None]          # slots beyond mi_cnt are not written
        assert np.array_equal(ba.mi[valid], bb.mi[valid]), "mi"
        # the plain-array form of the call, from pageable memory
        packed, nrid, nmask = (x.copy() for x in rs.arrays())
        rc = b.for_reads_packed(packed, nrid, nmask)
        assert np.array_equal(rc.cls, ra.cls) and np.array_equal(b.debug_read_tuples(n), ta)


def test_fastq_file_to_device():
    """the whole N3 path: file -> mcb_readset_add_fastq -> mcb_for_reads_packed, against the oracle on the rows"""
    import os
    import tempfile
    L, n = 100, 20000
    reads = synth.make_reads(n, L, 100000, seed=81, special=0.02)
    with tempfile.TemporaryDirectory() as d:
        fq = os.path.join(d, "in.fastq")
        synth.write_fastq(fq, reads)
        rs = api.ReadSet(L)
        rows = rs.add_fastq(fq, 4, want_ascii=True)
    assert np.array_equal(rows, reads)
    S = O.Stage1(O.resolve_params(L), reads)
    with api.Context(api.resolve_params(L)) as ctx:
        r = ctx.for_reads_packed(rs)
        assert np.array_equal(r.cls, S.cls) and np.array_equal(ctx.debug_read_tuples(n), S.tuples)
        br = ctx.for_bucket()
        assert np.array_equal(br.cl_a, S.cl_a) and np.array_equal(br.cl_ref, S.cl_ref) and np.array_equal(br.sg, S.sg)
    S.close()


def test_malformed_packed_input_is_refused():
    L, n = 100, 1000
    reads = synth.make_reads(n, L, 5000, seed=82, special=0.05)
    rs = api.ReadSet(L)
    rs.add_rows(reads)
    packed, nrid, nmask = (x.copy() for x in rs.arrays())
    assert len(nrid) > 2
    with api.Context(api.resolve_params(L)) as ctx:
        bad = packed.copy()
        bad[5, 3] |= np.uint64(1) << np.uint64(40)            # bits beyond base 99 (word 3 holds bases 96..99)
        with pytest.raises(api.McbError, match="malformed"):
            ctx.for_reads_packed(bad, nrid, nmask)
        with pytest.raises(api.McbError, match="malformed"):
            ctx.for_reads_packed(packed, nrid[::-1].copy(), nmask)      # not ascending
        m2 = nmask.copy()
        m2[0] = 0                                                # listed, but no N
        with pytest.raises(api.McbError, match="malformed"):
            ctx.for_reads_packed(packed, nrid, m2)
        r = ctx.for_reads_packed(packed, nrid, nmask)            # the context is still usable
        assert r.n_sketched > 0
