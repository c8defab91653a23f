"""TEST INFRASTRUCTURE: ctypes view of oracle/_build/libmc_oracle.so (oracle/mc_oracle.c, the CPU restatement of the
reference's front end).  Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "libmc_oracle.so")
MAXU64 = np.uint64(0xFFFFFFFFFFFFFFFF)


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("readlen", "k", "b", "rw", "first_mininum", "diff_threshold", "max_rounds")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "oracle", "mc_oracle.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(LIB)
        L.mco_hash64.restype = C.c_uint64
        L.mco_hash64.argtypes = [C.c_uint64, C.c_uint64]
        L.mco_sketch_two.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_uint32, C.c_void_p]
        L.mco_sketch_lh.restype = C.c_int64
        L.mco_sketch_lh.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_int64]
        L.mco_radix_sort_x.argtypes = [C.c_void_p, C.c_int64]
        L.mco_classify.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p]
        L.mco_encode_byte_ok.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.mco_kseq_all.restype = C.c_int64
        L.mco_kseq_all.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int64]
        L.mco_pack_row.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.mco_print_encode.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p]
        L.mco_stage1_run.restype = C.c_void_p
        L.mco_stage1_run.argtypes = [C.POINTER(Params), C.c_void_p, C.c_uint64]
        L.mco_stage1_free.argtypes = [C.c_void_p]
        L.mco_stage1_field.restype = C.c_void_p
        L.mco_stage1_field.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
        L.mco_stage1_rounds.argtypes = [C.c_void_p]
        L.mco_stage1_sketched.restype = C.c_uint64
        L.mco_stage1_sketched.argtypes = [C.c_void_p]
        L.mco_idx_build.restype = C.c_void_p
        L.mco_idx_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.mco_idx_get.restype = C.c_void_p
        L.mco_idx_get.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]
        L.mco_idx_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.mco_idx_keys.restype = C.c_void_p
        L.mco_idx_keys.argtypes = [C.c_void_p]
        L.mco_idx_postings.restype = C.c_void_p
        L.mco_idx_postings.argtypes = [C.c_void_p]
        L.mco_idx_key_starts.argtypes = [C.c_void_p, C.c_void_p]
        L.mco_idx_free.argtypes = [C.c_void_p]
        L.mco_realign.restype = C.c_void_p
        L.mco_realign.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int]
        L.mco_realign_field.restype = C.c_void_p
        L.mco_realign_field.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.c_uint64]
        L.mco_realign_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.mco_realign_free.argtypes = [C.c_void_p]
        L.mco_combine.restype = C.c_void_p
        L.mco_combine.argtypes = [C.c_void_p, C.c_int]
        L.mco_combine_field.restype = C.c_void_p
        L.mco_combine_field.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
        L.mco_combine_iterations.argtypes = [C.c_void_p]
        L.mco_combine_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def resolve_params(L, k=0, e=0, w=0, m=0, max_rounds=0):
    """minicommain.c:92-127, preprocess.c:89-107 (same rules as mcb_resolve_params, restated for the checker)."""
    kk = k if k > 0 else (17 if L < 80 else 31)
    mr = 35
    if 0 < max_rounds < 35:
        mr = max_rounds
    rw = w if w > 0 else (L // 2 - kk if L >= 70 else 3)
    return Params(L, kk, 14, rw, m if m > 0 else 6, e if e > 0 else 4, mr)


def hash64(key, mask):
    return int(lib().mco_hash64(key, mask))


def sketch_two(seq: bytes, k: int, rid: int):
    out = (C.c_uint64 * 2)()
    lib().mco_sketch_two(seq, len(seq), k, rid, out)
    return int(out[0]), int(out[1])


def sketch_lh(seq: bytes, w: int, k: int, rid: int, cap: int = 1 << 16):
    out = np.zeros((cap, 2), dtype=np.uint64)
    n = int(lib().mco_sketch_lh(seq, len(seq), w, k, rid, out.ctypes.data, cap))
    return out[:min(n, cap)].copy(), n


def radix_sort_x(xy: np.ndarray):
    a = np.ascontiguousarray(xy, dtype=np.uint64).copy()
    lib().mco_radix_sort_x(a.ctypes.data, len(a))
    return a


def classify(seq: bytes, e: int):
    repl = C.create_string_buffer(2)
    c = lib().mco_classify(seq, len(seq), e, repl)
    return c, repl.raw[:1] if repl.raw[0] else b""


def kseq_all(data: bytes):
    """every sequence kseq_read / bseq_read would deliver from the file contents `data` (list of bytes)"""
    seqs = np.zeros(len(data) + 1, dtype=np.uint8)
    lens = np.zeros(len(data) // 2 + 2, dtype=np.uint32)
    n = lib().mco_kseq_all(data, len(data), seqs.ctypes.data, len(seqs), lens.ctypes.data, len(lens))
    out, o = [], 0
    for i in range(n):
        out.append(seqs[o:o + int(lens[i])].tobytes())
        o += int(lens[i])
    return out


def pack_rows(rows: np.ndarray):
    """(packed u64[n][WS], mask u64[n][WS], status[n]) of ASCII rows in the device layout; status 0 / 1 (has N) / -1 (bad character)"""
    n, L = rows.shape
    ws = (((L + 31) // 32) + 1) & ~1
    packed, mask, st = np.zeros((n, ws), np.uint64), np.zeros((n, ws), np.uint64), np.zeros(n, np.int32)
    f = lib().mco_pack_row
    for i in range(n):
        st[i] = f(rows[i].tobytes(), L, ws, packed[i].ctypes.data, mask[i].ctypes.data)
    return packed, mask, st


def print_encode(read: bytes, direction: int, ref_window: bytes) -> bytes:
    """the line print_encode writes to dif_char.txt for one member (kthread_dump.c:85-118)"""
    out = C.create_string_buffer(len(read) + 2)
    n = lib().mco_print_encode(read, direction, ref_window, len(read), out)
    return out.raw[:n]


def encode_byte_ok(read_oriented: bytes, ref_window: bytes):
    return bool(lib().mco_encode_byte_ok(read_oriented, ref_window, len(read_oriented)))


def _field(fn, h, what, dtype, *extra, per=1):
    cnt = C.c_uint64(0)
    p = fn(h, what, C.byref(cnt), *extra)
    n = cnt.value * per
    if n == 0 or not p:
        return np.zeros(0, dtype=dtype)
    isz = np.dtype(dtype).itemsize
    return np.frombuffer((C.c_char * (n * isz)).from_address(p), dtype=dtype).copy()


class Stage1:
    """kt_for_reads + kt_for_bucket of the restatement."""

    def __init__(self, params: Params, reads: np.ndarray):
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        self.params, self.n, self.L = params, reads.shape[0], reads.shape[1]
        self._h = lib().mco_stage1_run(C.byref(params), reads.ctypes.data, self.n)
        f = lambda w, dt: _field(lib().mco_stage1_field, self._h, w, dt)  # noqa: E731
        self.cls = f(0, np.uint8)
        self.tuples = _field(lib().mco_stage1_field, self._h, 1, np.uint64, per=2).reshape(-1, 2)
        self.rows = f(2, np.uint8).reshape(self.n, self.L)
        self.cl_n = f(3, np.uint32)
        self.cl_a_off = f(4, np.uint64)
        self.cl_a = f(5, np.uint64)
        self.cl_ref_off = f(6, np.uint64)
        self.cl_ref = f(7, np.uint8)
        self.sg = f(8, np.uint32)
        self.mi = _field(lib().mco_stage1_field, self._h, 9, np.uint64, per=2).reshape(-1, 2)
        self.rounds = int(lib().mco_stage1_rounds(self._h))
        self.n_sketched_total = int(lib().mco_stage1_sketched(self._h))

    def realign(self, sg, refs, ref_off, threshold, maxsearch, ininumdict=0):
        sg = np.ascontiguousarray(sg, dtype=np.uint32)
        refs = np.ascontiguousarray(refs, dtype=np.uint8)
        ref_off = np.ascontiguousarray(ref_off, dtype=np.uint64)
        h = lib().mco_realign(self._h, sg.ctypes.data, len(sg), refs.ctypes.data, ref_off.ctypes.data, len(ref_off) - 1, threshold, maxsearch, ininumdict)
        f = lambda w, dt: _field(lib().mco_realign_field, h, w, dt, len(sg))  # noqa: E731
        cnt = np.zeros(4, dtype=np.uint64)
        lib().mco_realign_counters(h, cnt.ctypes.data)
        out = {"claim_contig": f(0, np.uint32), "claim_sg": f(1, np.uint32), "claim_y": f(2, np.uint64), "fpA_sg": f(3, np.uint32), "fpT_sg": f(4, np.uint32),
               "flag": f(5, np.uint8), "n_windows": int(cnt[0]), "n_probes": int(cnt[1]), "n_candidates": int(cnt[2]), "numdict": int(cnt[3])}
        lib().mco_realign_free(h)
        return out

    def combine(self, cbthreshold: int):
        """combine_cluster (kthread_cb.c:570-630) over the seed contigs of this run.  Returns a dict with the final contigs, the
        number of iterations, and the tuples handed to every mm_idx_generation in push order."""
        h = lib().mco_combine(self._h, cbthreshold)
        f = lambda w, dt: _field(lib().mco_combine_field, h, w, dt)  # noqa: E731
        out = {"cl_n": f(0, np.uint32), "cl_a_off": f(1, np.uint64), "cl_a": f(2, np.uint64), "cl_ref_off": f(3, np.uint64), "cl_ref": f(4, np.uint8),
               "iter_off": f(5, np.uint64), "iter_tuples": _field(lib().mco_combine_field, h, 6, np.uint64, per=2).reshape(-1, 2), "iter_merges": f(7, np.uint32),
               "iterations": int(lib().mco_combine_iterations(h))}
        lib().mco_combine_free(h)
        return out

    def close(self):
        if self._h:
            lib().mco_stage1_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


class Index:
    def __init__(self, xy: np.ndarray, bucket_off: np.ndarray, b: int = 14):
        xy = np.ascontiguousarray(xy, dtype=np.uint64)
        off = np.ascontiguousarray(bucket_off, dtype=np.uint64)
        self._h = lib().mco_idx_build(xy.ctypes.data, off.ctypes.data, b)
        nk, npost = C.c_uint64(0), C.c_uint64(0)
        lib().mco_idx_stats(self._h, C.byref(nk), C.byref(npost))
        self.n_keys, self.n_post = nk.value, npost.value

    def get(self, x):
        n = C.c_int(0)
        p = lib().mco_idx_get(self._h, int(x), C.byref(n))
        if not p or n.value == 0:
            return np.zeros(0, dtype=np.uint64)
        return np.frombuffer((C.c_char * (n.value * 8)).from_address(p), dtype=np.uint64).copy()

    def flat(self):
        """(keys, starts, postings) in (bucket, key) order."""
        nk, npost = self.n_keys, self.n_post
        keys = np.frombuffer((C.c_char * (nk * 8)).from_address(lib().mco_idx_keys(self._h)), dtype=np.uint64).copy() if nk else np.zeros(0, np.uint64)
        post = np.frombuffer((C.c_char * (npost * 8)).from_address(lib().mco_idx_postings(self._h)), dtype=np.uint64).copy() if npost else np.zeros(0, np.uint64)
        st = np.zeros(nk + 1, dtype=np.uint64)
        lib().mco_idx_key_starts(self._h, st.ctypes.data)
        return keys, st, post

    def close(self):
        if self._h:
            lib().mco_idx_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


def bucket_major(tuples, b=14):
    """stable partition of (n,2) tuples by bucket = x & (2^b-1): what per-bucket pushes in the same order would give."""
    bk = (tuples[:, 0] & np.uint64((1 << b) - 1)).astype(np.int64)
    order = np.argsort(bk, kind="stable")
    cnt = np.bincount(bk, minlength=1 << b).astype(np.uint64)
    off = np.zeros((1 << b) + 1, dtype=np.uint64)
    np.cumsum(cnt, out=off[1:])
    return off, tuples[order]
