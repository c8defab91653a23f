"""The two places where the CUDA kernels do NOT follow the reference's control flow literally, restated in plain Python and
checked against the oracle (itself pinned to the reference) on random and adversarial strings.  No GPU needed: these tests
guard the ALGORITHMS; tests/test_gpu_parity.py checks the kernels.

 1. mcb_sketch_two_wide (minicom_b200/csrc/mcb_common.cuh): mm_sketch_two (sketch.c:238-289) without its counter `l` —
    "hash a position iff at least k bases were read and the k-mer is not its own reverse complement" — and, for odd k,
    without the symmetry test at all.
 2. k_sketch_lh2 (minicom_b200/csrc/mcb_stage1.cu): mm_sketch_lh_ori (sketch.c:116-165) with the window minimum taken
    from block suffix minima + a running prefix minimum instead of rescanning the ring, and the identical-k-mer scans run
    only when a "same hash elsewhere" bit is set.
"""
import numpy as np
import pytest

import oracle_lib as O

CODE = {65: 0, 67: 1, 71: 2, 84: 3}          # A C G T
FULL = (1 << 64) - 1


def sketch_two_simplified(seq: bytes, k: int, rid: int):
    mask, shift1 = (1 << (2 * k)) - 1, 2 * (k - 1)
    f = r = 0
    best, bp, bz = FULL, 0, 0
    for i, ch in enumerate(seq):
        c = CODE[ch]
        f = (f << 2 | c) & mask
        r = (r >> 2) | ((3 ^ c) << shift1)
        if i < k - 1:
            assert f != r, "a partial window can never be symmetric"
            continue
        if f == r:
            assert k % 2 == 0, "an odd-length k-mer can never be symmetric"
            continue
        z = 0 if f < r else 1
        h = O.hash64(r if z else f, mask)
        if h < best:
            best, bp, bz = h, i, z
    return best, (rid << 32) | (bp << 1) | bz


def _strings(rng, n, L):
    out = []
    for _ in range(n):
        kind = rng.integers(0, 4)
        if kind == 0:
            s = rng.integers(0, 4, size=L)
        elif kind == 1:                                   # tandem repeat (palindromic units included)
            u = rng.integers(0, 4, size=int(rng.integers(1, 9)))
            s = np.tile(u, L // len(u) + 1)[:L]
        elif kind == 2:                                   # reverse-complement palindrome around the middle
            h = rng.integers(0, 4, size=L // 2)
            s = np.concatenate([h, (3 - h)[::-1], rng.integers(0, 4, size=L - 2 * (L // 2))])
        else:                                             # low complexity: long runs of one base
            s = np.repeat(rng.integers(0, 4, size=L // 7 + 1), 7)[:L]
        out.append(bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[s]))
    return out


@pytest.mark.parametrize("k", [2, 3, 8, 15, 16, 17, 20, 30, 31])
def test_sketch_two_without_the_counter(k):
    rng = np.random.default_rng(1000 + k)
    for s in _strings(rng, 60, 100) + _strings(rng, 20, max(k, 33)):
        want = O.sketch_two(s, k, 5)
        got = sketch_two_simplified(s, k, 5)
        if want[0] == FULL:                               # no k-mer at all (every window symmetric): position fields are unspecified
            assert got[0] == FULL
        else:
            assert got == want, (k, s)


def sketch_lh_block_minimum(seq: bytes, w: int, k: int, rid: int):
    """k_sketch_lh2, statement by statement (without the early exit after m outputs)."""
    mask, shift1 = (1 << (2 * k)) - 1, 2 * (k - 1)
    NOP = 0xFFFFFFFF
    rx, rp = [FULL] * w, [NOP] * w
    ss = [w - 1] * w                                      # suffix-minimum slot | tie << 7
    out = []
    fw = rv = 0
    mn_x, mn_p = FULL, NOP
    px, pslot, ptie = FULL, 0, False
    l = bp = mp = 0
    emit = lambda hx, p: out.append((hx, (rid << 32) | p))
    for i, ch in enumerate(seq):
        cc = CODE.get(ch, 4)
        ix, ip = FULL, NOP
        if cc < 4:
            fw = (fw << 2 | cc) & mask
            rv = (rv >> 2) | ((3 ^ cc) << shift1)
            if fw == rv:
                continue
            z = 0 if fw < rv else 1
            l += 1
            if l >= k:
                ix, ip = O.hash64(rv if z else fw, mask), (i << 1) | z
        else:
            l = 0
        rx[bp], rp[bp] = ix, ip
        if ix <= px:
            ptie, px, pslot = (ix == px), ix, bp
        if l == w + k - 1:
            for j in list(range(bp + 1, w)) + list(range(0, bp)):
                if mn_x == rx[j] and rp[j] != mn_p:
                    emit(rx[j], rp[j])
        if ix <= mn_x:
            if l >= w + k:
                emit(mn_x, mn_p)
            mn_x, mn_p, mp = ix, ip, bp
        elif bp == mp:
            if l >= w + k - 1:
                emit(mn_x, mn_p)
            nx, ns, tie = px, pslot, ptie
            if bp + 1 < w:
                sv = ss[bp + 1]
                sslot = sv & 0x7F
                sx = rx[sslot]
                if sx < px:
                    nx, ns, tie = sx, sslot, bool(sv >> 7)
                elif sx == px:
                    tie = True
            mn_x, mp, mn_p = nx, ns, rp[ns]
            if tie and l >= w + k - 1:
                for j in list(range(bp + 1, w)) + list(range(0, bp + 1)):
                    if mn_x == rx[j] and mn_p != rp[j]:
                        emit(rx[j], rp[j])
        bp += 1
        if bp == w:
            bp = 0
            sx, sv = rx[w - 1], w - 1
            ss[w - 1] = sv
            for j in range(w - 2, 0, -1):
                v = rx[j]
                if v < sx:
                    sx, sv = v, j
                elif v == sx:
                    sv |= 0x80
                ss[j] = sv
            px, pslot, ptie = FULL, 0, False
    if mn_x != FULL:
        emit(mn_x, mn_p)
    return out


@pytest.mark.parametrize("w,k", [(1, 5), (2, 4), (3, 7), (5, 6), (7, 20), (19, 31), (19, 30), (40, 15), (100, 11)])
def test_sketch_lh_with_block_minima(w, k):
    rng = np.random.default_rng(2000 + 37 * w + k)
    strings = _strings(rng, 24, 260) + _strings(rng, 8, w + k + 3) + _strings(rng, 4, max(1, k - 1))
    with_n = []
    for s in _strings(rng, 8, 260):                       # ambiguous bases reset the run length but keep their ring slot
        b = bytearray(s)
        for p in rng.integers(0, len(b), size=4):
            b[p] = ord("N")
        with_n.append(bytes(b))
    for s in strings + with_n:
        want, n = O.sketch_lh(s, w, k, 0x700)
        got = sketch_lh_block_minimum(s, w, k, 0x700)
        assert n == len(got), (w, k, s)
        assert [tuple(int(v) for v in row) for row in want] == got, (w, k, s)


def test_oracle_agrees_with_the_reference_on_the_adversarial_strings():
    """The strings above are nastier than the random ones the oracle was pinned with (tests/test_oracle.py): tandem repeats,
    reverse-complement palindromes, single-base runs.  Where the reference units library is built (oracle/_ref), pin the
    oracle on them directly against the reference's own mm_sketch_two / mm_sketch_lh_ori."""
    import ctypes as C
    from test_oracle import _ref_units
    R = _ref_units()
    rng = np.random.default_rng(77)
    for trial, s in enumerate(_strings(rng, 120, 200)):
        k = int(rng.integers(4, 32))
        out = (C.c_uint64 * 2)()
        R.ref_sketch_two(s, len(s), k, trial, out)
        assert O.sketch_two(s, k, trial) == (int(out[0]), int(out[1])), (s, k)
        w = int(rng.integers(1, 40))
        buf = np.zeros((4096, 2), dtype=np.uint64)
        cnt = R.ref_sketch_lh_ori(s, len(s), w, k, trial, buf.ctypes.data, 4096)
        got, n_got = O.sketch_lh(s, w, k, trial, cap=4096)
        assert n_got == cnt and np.array_equal(got, buf[:cnt]), (s, w, k)


# ---------------------------------------------------------------- 3. top-aligned arithmetic (mcb_hash64_ta, MCB_TA_ROLL, mcb_common.cuh)
# For 16 < k < 32 the device keeps k-mers and hashes in the TOP 2k bits of a pair of 32-bit halves (value << (64 - 2k)): hash64's
# "& mask" after each multiplication becomes the wrap-around of 64-bit arithmetic, x ^= x >> s needs the shifted-in low bits
# cleared, and some shifts are written as multiplications by powers of two (IMAD / IMAD.HI on the FMA pipe).  Restated here with
# Python integers, 32 bits at a time exactly as the kernel does it, and compared with the oracle's hash64 / the plain roll.
M32 = (1 << 32) - 1


def _umulhi(a, b):
    return ((a * b) >> 32) & M32


def _madhi(a, b, c):
    return (_umulhi(a, b) + c) & M32


def hash64_top_aligned(lo, hi, k):
    sh = 64 - 2 * k
    lm = (~((1 << sh) - 1)) & M32
    m24, m14, m28 = 1 << 8, 1 << 18, 1 << 4
    sub = ((1 << 64) - (1 << sh)) & FULL
    t = (lo * 0x1FFFFF + sub) & FULL
    hi = (hi * 0x1FFFFF + (t >> 32)) & M32
    lo = t & M32
    fl, u = _madhi(lo, m24, (hi * m24) & M32), _umulhi(hi, m24)                  # x ^= x >> 24: funnel shift as IMAD + IMAD.HI
    lo, hi = lo ^ (fl & lm), hi ^ u
    t = lo * 265
    hi = (hi * 265 + (t >> 32)) & M32
    lo = t & M32
    fl, u = (((hi << 32) | lo) >> 14) & M32, _umulhi(hi, m14)                    # x ^= x >> 14
    lo, hi = lo ^ (fl & lm), hi ^ u
    t = lo * 21
    hi = (hi * 21 + (t >> 32)) & M32
    lo = t & M32
    fl, u = (((hi << 32) | lo) >> 28) & M32, _umulhi(hi, m28)                    # x ^= x >> 28
    lo, hi = lo ^ (fl & lm), hi ^ u
    t = lo * 0x80000001
    hi = (hi * 0x80000001 + (t >> 32)) & M32
    lo = t & M32
    return (hi << 32) | lo


@pytest.mark.parametrize("k", list(range(17, 32)))
def test_top_aligned_hash_and_roll_equal_the_plain_arithmetic(k):
    rng = np.random.default_rng(k)
    mask, sh = (1 << (2 * k)) - 1, 64 - 2 * k
    keys = [0, 1, mask, mask - 1, 1 << (2 * k - 1)] + [int(x) & mask for x in rng.integers(0, 1 << 62, size=3000)]
    for key in keys:
        kp = (key << sh) & FULL
        got = hash64_top_aligned(kp & M32, kp >> 32, k)
        assert got & ((1 << sh) - 1) == 0, "a top-aligned hash has empty low bits"
        assert got >> sh == O.hash64(key, mask)
    # the roll (sketch.c:257-259): f = (f << 2 | c) & mask, r = (r >> 2) | ((3 ^ c) << 2(k-1)), on the halves
    lm, csh, p30, neg30 = (~((1 << sh) - 1)) & M32, 1 << sh, 1 << 30, (-(1 << 30)) & M32
    f = r = 0
    flo = fhi = rlo = rhi = 0
    for c in rng.integers(0, 4, size=2000).tolist():
        f = (f << 2 | c) & mask
        r = (r >> 2) | ((3 ^ c) << (2 * (k - 1)))
        fhi = ((fhi << 2) | (flo >> 30)) & M32
        flo = (flo * 4 + c * csh) & M32
        rlo = ((((rhi << 32) | rlo) >> 2) & M32) & lm
        rhi = _madhi(rhi, p30, (c * neg30 + 0xC0000000) & M32)
        assert ((fhi << 32) | flo) == (f << sh) and ((rhi << 32) | rlo) == (r << sh)
