"""Seeded synthetic read sets of the shapes BASELINE.json names.

A read set is a pure function of (n_reads, read_len, genome_len, seed, ...):
uniform random genome over ACGT; each read = uniform start, `read_len` bases,
each base substituted with probability `sub_rate` to a uniform *different*
base, reverse-complemented with probability 0.5.  `special` injects the read
classes kthread_reads.c:84-226 distinguishes (all-A / all-T / all-N, mostly
A/T/N, few-N, many-N) for the correctness sets; the throughput sets have none.

Reads are returned as an (n, L) uint8 array of ASCII codes (the reference's
in-memory representation is one NUL-terminated ASCII string per read,
bseq.c:38-66).
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


def make_genome(genome_len: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng([seed, 0x67656E])
    return _ACGT[rng.integers(0, 4, size=genome_len, dtype=np.uint8)]


def make_reads(n_reads: int, read_len: int, genome_len: int, seed: int = 1,
               sub_rate: float = 0.01, special: float = 0.0,
               genome: np.ndarray | None = None) -> np.ndarray:
    """(n_reads, read_len) uint8 ASCII array."""
    if genome is None:
        genome = make_genome(genome_len, seed)
    rng = np.random.default_rng([seed, 0x726561])
    out = np.empty((n_reads, read_len), dtype=np.uint8)
    chunk = 1 << 20
    idx = np.arange(read_len, dtype=np.int64)
    for lo in range(0, n_reads, chunk):
        hi = min(n_reads, lo + chunk)
        m = hi - lo
        start = rng.integers(0, genome_len - read_len + 1, size=m, dtype=np.int64)
        r = genome[start[:, None] + idx[None, :]]
        sub = rng.random((m, read_len)) < sub_rate
        if sub.any():
            # substitute by a uniform different base: rotate the code by 1..3
            code = np.searchsorted(_ACGT, r[sub])          # A0 C1 G2 T3 (ACGT is sorted)
            rot = rng.integers(1, 4, size=code.shape[0])
            r[sub] = _ACGT[(code + rot) & 3]
        rc = rng.random(m) < 0.5
        r[rc] = _COMP[r[rc][:, ::-1]]
        out[lo:hi] = r
    if special > 0:
        _inject_special(out, rng, special)
    return out


def _inject_special(reads: np.ndarray, rng: np.random.Generator, frac: float) -> None:
    n, L = reads.shape
    k = max(8, int(n * frac))
    rows = rng.choice(n, size=min(k, n), replace=False)
    for j, r in enumerate(rows):
        kind = j % 10
        if kind == 0:
            reads[r] = ord("A")
        elif kind == 1:
            reads[r] = ord("T")
        elif kind == 2:
            reads[r] = ord("N")
        elif kind == 3:          # mostly A, a few other bases
            reads[r] = ord("A")
            pos = rng.choice(L, size=int(rng.integers(1, 4)), replace=False)
            reads[r, pos] = _ACGT[rng.integers(1, 4, size=pos.size)]
        elif kind == 4:          # mostly T
            reads[r] = ord("T")
            pos = rng.choice(L, size=int(rng.integers(1, 4)), replace=False)
            reads[r, pos] = _ACGT[rng.integers(0, 3, size=pos.size)]
        elif kind == 5:          # mostly N
            reads[r] = ord("N")
            pos = rng.choice(L, size=int(rng.integers(1, 4)), replace=False)
            reads[r, pos] = _ACGT[rng.integers(0, 4, size=pos.size)]
        elif kind == 6:          # a few N in a normal read
            pos = rng.choice(L, size=int(rng.integers(1, 6)), replace=False)
            reads[r, pos] = ord("N")
        elif kind == 7:          # many N (> 0.4 L) but not "mostly N"
            pos = rng.choice(L, size=int(0.4 * L) + 2 + int(rng.integers(0, 5)), replace=False)
            reads[r, pos] = ord("N")
        elif kind == 8:          # near poly-A beyond the stage-1 threshold (stage-2 diversion, bbhashdict.c:157-186)
            reads[r] = ord("A")
            pos = rng.choice(L, size=int(rng.integers(5, 9)), replace=False)
            reads[r, pos] = _ACGT[rng.integers(1, 4, size=pos.size)]
        else:                    # near poly-T likewise, with an N
            reads[r] = ord("T")
            pos = rng.choice(L, size=int(rng.integers(5, 9)), replace=False)
            reads[r, pos] = _ACGT[rng.integers(0, 3, size=pos.size)]
            reads[r, int(rng.integers(0, L))] = ord("N")


def fastq_record_bytes(L: int) -> int:
    return 3 + L + 3 + L + 1


def write_fastq(path: str, reads: np.ndarray, first_record: int | None = None) -> None:
    """4-line FASTQ with dummy names/qualities (all the reference reads is the sequence line).
    first_record: write into an EXISTING file at that record index (records have a fixed size), so that several processes
    can fill one file side by side."""
    n, L = reads.shape
    head = b"@r\n"
    tail = b"\n+\n" + b"I" * L + b"\n"
    rec = len(head) + L + len(tail)
    with open(path, "wb" if first_record is None else "r+b") as f:
        if first_record is not None:
            f.seek(first_record * rec)
        chunk = 1 << 18
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            buf = np.empty((hi - lo, rec), dtype=np.uint8)
            buf[:, :len(head)] = np.frombuffer(head, dtype=np.uint8)
            buf[:, len(head):len(head) + L] = reads[lo:hi]
            buf[:, len(head) + L:] = np.frombuffer(tail, dtype=np.uint8)
            f.write(buf.tobytes())


# The five BASELINE.json shapes: (n_reads, read_len, genome_len)
SHAPES = {
    "C1": (1_000_000, 100, 5_000_000),
    "C2": (10_000_000, 100, 50_000_000),
    "C3": (5_000_000, 101, 50_000_000),      # per file; two files
    "C4": (50_000_000, 150, 375_000_000),
    "C5": (200_000_000, 100, 1_000_000_000),
}
