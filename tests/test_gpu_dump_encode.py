"""N2 first slice on the device: mcb_dump_encode (print_encode's per-read diff encoding, kthread_dump.c:66-118) must reproduce the
reference's own dif_char.txt for the final clusters of the committed fixtures, and the oracle on random members."""
import numpy as np
import pytest

import oracle_lib as O
import refdump
from minicom_b200 import api, synth
from test_dump_encode_cpu import final_clusters

pytestmark = pytest.mark.gpu


def _flat(refs, members):
    moff = np.zeros(len(members) + 1, dtype=np.uint64)
    np.cumsum([len(m) for m in members], out=moff[1:])
    roff = np.zeros(len(refs) + 1, dtype=np.uint64)
    np.cumsum([len(r) for r in refs], out=roff[1:])
    mem = np.array([y for m in members for y in m], dtype=np.uint64)
    return mem, moff, np.frombuffer(b"".join(refs), dtype=np.uint8), roff


@pytest.mark.parametrize("name", refdump.golden_names())
@pytest.mark.parametrize("packed_input", [False, True], ids=["ascii", "packed"])
def test_dump_encode_reproduces_the_reference_dif_char_file(name, packed_input):
    reads, meta, d, out = refdump.load_golden(name)
    env = meta["env"]
    p = api.resolve_params(meta["L"], k=int(env.get("MC_K", 0)), e=int(env.get("MC_E", 0)), w=int(env.get("MC_W", 0)), m=int(env.get("MC_M", 0)))
    refs, members = final_clusters(reads, meta, d)
    mem, moff, rflat, roff = _flat(refs, members)
    with api.Context(p) as ctx:
        if packed_input:
            rs = api.ReadSet(meta["L"])
            rs.add_rows(reads)
            ctx.for_reads_packed(rs)
        else:
            ctx.for_reads(reads)
        eo, enc = ctx.dump_encode(mem, moff, rflat, roff)
    got = b"".join(enc[int(eo[i]):int(eo[i + 1])] + b"\n" for i in range(len(mem)))
    assert got == out["dif_char.txt.0"]


def test_dump_encode_matches_oracle_on_random_members_and_uses_the_contigs_on_the_device():
    L, n = 100, 30000
    reads = synth.make_reads(n, L, 60000, seed=91, special=0.03)
    with api.Context(api.resolve_params(L)) as ctx:
        ctx.for_reads(reads)
        ctx.for_bucket()
        cr = ctx.combine(8)                                   # the merged contigs are the ones the context holds from here on
        nc = len(cr.cl_n)
        assert nc > 100
        eo, enc = ctx.dump_encode(cr.cl_a, cr.cl_a_off)       # refs = None: contigs on the device
        for i in np.random.default_rng(5).integers(0, len(cr.cl_a), size=3000):
            c = int(np.searchsorted(cr.cl_a_off, i, side="right")) - 1
            y = int(cr.cl_a[i])
            rid, pos, direction = y >> 32, (y & 0xFFFFFFFF) >> 1, y & 1
            ref = cr.cl_ref[int(cr.cl_ref_off[c]):int(cr.cl_ref_off[c + 1])].tobytes()
            assert enc[int(eo[i]):int(eo[i + 1])] == O.print_encode(reads[rid].tobytes(), direction, ref[pos:pos + L])
        # members that do not fit their contig are refused
        bad = cr.cl_a.copy()
        bad[0] = (int(bad[0]) & ~0xFFFFFFFF) | (5000 << 1)
        with pytest.raises(api.McbError, match="outside their contig"):
            ctx.dump_encode(bad, cr.cl_a_off)
