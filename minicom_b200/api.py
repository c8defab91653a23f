"""ctypes binding of libminicom_b200.so (include/minicom_b200.h) for tests, bench.py and smoke().

This is plumbing only: every function forwards to the C-ABI entry point of the same name and copies the
context-owned result arrays into numpy arrays.  There is no Python or CPU implementation of the path here; if the
library (or a CUDA device) is missing the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libminicom_b200.so")

CLS_NAMES = ["sketched", "allA", "allT", "allN", "fpA", "fpT", "fpN", "Nfile"]


class McbError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("readlen", C.c_int32), ("k", C.c_int32), ("b", C.c_int32), ("rw", C.c_int32),
                ("first_mininum", C.c_int32), ("diff_threshold", C.c_int32), ("max_rounds", C.c_int32),
                ("device", C.c_int32)]


class _ReadsResult(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("cls", C.c_void_p), ("n_nreads", C.c_uint64), ("nread_rid", C.c_void_p),
                ("nread_repl", C.c_void_p), ("nread_off", C.c_void_p), ("npos", C.c_void_p), ("n_sketched", C.c_uint64)]


class _ReadsetView(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("readlen", C.c_int32), ("row_words", C.c_int32), ("packed", C.c_void_p), ("n_nreads", C.c_uint64),
                ("nread_rid", C.c_void_p), ("nmask", C.c_void_p)]


class _EncodeResult(C.Structure):
    _fields_ = [("n_members", C.c_uint64), ("n_bytes", C.c_uint64), ("enc_off", C.c_void_p), ("enc", C.c_void_p)]


class _BucketResult(C.Structure):
    _fields_ = [("n_clusters", C.c_uint64), ("cl_n", C.c_void_p), ("cl_a_off", C.c_void_p), ("cl_a", C.c_void_p),
                ("cl_ref_off", C.c_void_p), ("cl_ref", C.c_void_p), ("n_sg", C.c_uint64), ("sg", C.c_void_p),
                ("mi_cnt", C.c_void_p), ("mi", C.c_void_p), ("rounds", C.c_int32), ("n_sketched_total", C.c_uint64),
                ("n_grouped", C.c_uint64)]


class _CombineResult(C.Structure):
    _fields_ = [("n_clusters", C.c_uint64), ("cl_n", C.c_void_p), ("cl_a_off", C.c_void_p), ("cl_a", C.c_void_p), ("cl_ref_off", C.c_void_p),
                ("cl_ref", C.c_void_p), ("iterations", C.c_int32), ("n_merges", C.c_uint64), ("n_index_tuples", C.c_uint64)]


class _RealignResult(C.Structure):
    _fields_ = [("n_claims", C.c_uint64), ("claim_contig", C.c_void_p), ("claim_sg", C.c_void_p), ("claim_y", C.c_void_p),
                ("claim_prio", C.c_void_p), ("n_fpA", C.c_uint64), ("n_fpT", C.c_uint64), ("fpA_sg", C.c_void_p), ("fpT_sg", C.c_void_p),
                ("n_windows", C.c_uint64), ("n_probes", C.c_uint64), ("n_candidates", C.c_uint64),
                ("n_dict_keys", C.c_uint64), ("numdict", C.c_int32)]


EXPORTS = [
    "mcb_resolve_params", "mcb_create", "mcb_destroy", "mcb_last_error", "mcb_version",
    "mcb_for_reads", "mcb_for_reads_ptrs", "mcb_for_reads_device", "mcb_for_reads_packed", "mcb_for_reads_packed_device",
    "mcb_dump_encode", "mcb_readset_create", "mcb_readset_destroy", "mcb_readset_add_fastq", "mcb_readset_add_fastq_buffer", "mcb_readset_add_rows", "mcb_readset_get",
    "mcb_debug_read_tuples", "mcb_debug_sketch_two", "mcb_debug_unpack_reads",
    "mcb_for_bucket", "mcb_for_bucket_keep", "mcb_idx_build", "mcb_idx_build_scattered", "mcb_idx_get", "mcb_idx_destroy", "mcb_idx_stats", "mcb_idx_arrays",
    "mcb_combine", "mcb_realign", "mcb_sketch_lh_host", "mcb_sketch_two_host", "mcb_hash64",
    "mcb_timers_enable", "mcb_timers_reset", "mcb_timer_get", "mcb_timers_dump", "mcb_kernel_launches",
    "mcb_round_control_init", "mcb_round_control_begin", "mcb_round_control_end",
    "mcb_group_create", "mcb_group_destroy", "mcb_group_size", "mcb_group_context", "mcb_group_for_reads_ptrs", "mcb_group_for_bucket",
    "mcb_group_idx_build_scattered", "mcb_group_realign", "mcb_group_timers_dump",
    "mcb_shard_unique_id", "mcb_shard_init", "mcb_shard_attach", "mcb_shard_begin", "mcb_shard_for_bucket", "mcb_shard_realign",
]
NCCL_ID_BYTES = 128


class RoundControl(C.Structure):
    _fields_ = [("round", C.c_int32), ("last_rounds", C.c_int32), ("pre_members", C.c_int64)]

_lib = None


def load_library() -> C.CDLL:
    """Load the CUDA library.  Fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise McbError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(or `make -C minicom_b200/csrc`)")
    lib = C.CDLL(LIB_PATH)
    lib.mcb_last_error.restype = C.c_char_p
    lib.mcb_version.restype = C.c_char_p
    lib.mcb_resolve_params.argtypes = [C.POINTER(Params)] + [C.c_int] * 6
    lib.mcb_resolve_params.restype = None
    lib.mcb_create.argtypes = [C.POINTER(Params), C.POINTER(C.c_void_p)]
    lib.mcb_destroy.argtypes = [C.c_void_p]
    lib.mcb_destroy.restype = None
    lib.mcb_for_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(_ReadsResult)]
    lib.mcb_for_reads_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(_ReadsResult)]
    lib.mcb_for_reads_ptrs.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.POINTER(_ReadsResult)]
    lib.mcb_for_reads_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(_ReadsResult)]
    lib.mcb_for_reads_packed_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(_ReadsResult)]
    lib.mcb_readset_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.mcb_readset_destroy.argtypes = [C.c_void_p]
    lib.mcb_readset_destroy.restype = None
    lib.mcb_readset_add_fastq.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.mcb_readset_add_fastq_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.mcb_readset_add_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
    lib.mcb_readset_get.argtypes = [C.c_void_p, C.POINTER(_ReadsetView)]
    lib.mcb_readset_get.restype = None
    lib.mcb_dump_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(_EncodeResult)]
    lib.mcb_debug_read_tuples.argtypes = [C.c_void_p, C.c_void_p]
    lib.mcb_debug_sketch_two.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    lib.mcb_debug_unpack_reads.argtypes = [C.c_void_p, C.c_void_p]
    lib.mcb_for_bucket.argtypes = [C.c_void_p, C.POINTER(_BucketResult)]
    lib.mcb_for_bucket_keep.argtypes = [C.c_void_p, C.POINTER(_BucketResult)]
    lib.mcb_idx_build.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.mcb_idx_build_scattered.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.mcb_idx_get.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]
    lib.mcb_idx_get.restype = C.c_void_p
    lib.mcb_idx_destroy.argtypes = [C.c_void_p]
    lib.mcb_idx_destroy.restype = None
    lib.mcb_idx_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.mcb_idx_stats.restype = None
    lib.mcb_idx_arrays.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 4
    lib.mcb_idx_arrays.restype = None
    lib.mcb_combine.argtypes = [C.c_void_p, C.c_int, C.POINTER(_CombineResult)]
    lib.mcb_realign.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64,
                                C.c_int, C.c_int, C.c_int, C.POINTER(_RealignResult)]
    lib.mcb_sketch_lh_host.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_int64]
    lib.mcb_sketch_lh_host.restype = C.c_int64
    lib.mcb_sketch_two_host.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_uint32, C.c_void_p]
    lib.mcb_sketch_two_host.restype = None
    lib.mcb_hash64.argtypes = [C.c_uint64, C.c_uint64]
    lib.mcb_hash64.restype = C.c_uint64
    lib.mcb_timers_enable.argtypes = [C.c_void_p, C.c_int]
    lib.mcb_timers_enable.restype = None
    lib.mcb_timers_reset.argtypes = [C.c_void_p]
    lib.mcb_timers_reset.restype = None
    lib.mcb_timer_get.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint64)]
    lib.mcb_timer_get.restype = C.c_double
    lib.mcb_timers_dump.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.mcb_timers_dump.restype = C.c_size_t
    lib.mcb_kernel_launches.argtypes = [C.c_void_p]
    lib.mcb_kernel_launches.restype = C.c_uint64
    lib.mcb_round_control_init.argtypes = [C.POINTER(RoundControl)]
    lib.mcb_round_control_init.restype = None
    lib.mcb_round_control_begin.argtypes = [C.POINTER(RoundControl), C.c_int, C.c_int]
    lib.mcb_round_control_end.argtypes = [C.POINTER(RoundControl), C.c_uint64]
    lib.mcb_shard_unique_id.argtypes = [C.c_void_p]
    lib.mcb_shard_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.mcb_shard_attach.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.mcb_shard_begin.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
    lib.mcb_shard_for_bucket.argtypes = [C.c_void_p, C.POINTER(_BucketResult), C.c_void_p, C.c_int]
    lib.mcb_shard_realign.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64,
                                      C.c_int, C.c_int, C.c_int, C.POINTER(_RealignResult)]
    _lib = lib
    return lib


# Result arrays live in context-owned pinned memory that the next call of the same entry point overwrites, so the wrappers
# copy them out by default.  Callers that consume a result before the next call (bench.py) can switch the copies off.
VIEW_COPY = True


def _view(ptr, n, dtype):
    if not n or not ptr:
        return np.zeros(0, dtype=dtype)
    nbytes = int(n) * np.dtype(dtype).itemsize
    buf = (C.c_char * nbytes).from_address(ptr)
    a = np.frombuffer(buf, dtype=dtype)
    return a.copy() if VIEW_COPY else a


def resolve_params(readlen, k=0, e=0, w=0, m=0, max_rounds=0, device=0) -> Params:
    p = Params()
    load_library().mcb_resolve_params(C.byref(p), readlen, k, e, w, m, max_rounds)
    p.device = device
    return p


def hash64(key: int, mask: int) -> int:
    return int(load_library().mcb_hash64(key, mask))


def sketch_two_host(seq: bytes, k: int, rid: int):
    out = np.zeros(2, dtype=np.uint64)
    load_library().mcb_sketch_two_host(seq, len(seq), k, rid, out.ctypes.data)
    return int(out[0]), int(out[1])


def sketch_lh_host(seq: bytes, w: int, k: int, rid: int) -> np.ndarray:
    lib = load_library()
    cap = max(16, len(seq))
    out = np.zeros((cap, 2), dtype=np.uint64)
    n = lib.mcb_sketch_lh_host(seq, len(seq), w, k, rid, out.ctypes.data, cap)
    if n > cap:
        out = np.zeros((n, 2), dtype=np.uint64)
        n = lib.mcb_sketch_lh_host(seq, len(seq), w, k, rid, out.ctypes.data, n)
    return out[:n].copy()


@dataclass
class ReadsResult:
    cls: np.ndarray
    nread_rid: np.ndarray
    nread_repl: np.ndarray
    nread_off: np.ndarray
    npos: np.ndarray
    n_sketched: int


@dataclass
class BucketResult:
    cl_n: np.ndarray
    cl_a_off: np.ndarray
    cl_a: np.ndarray
    cl_ref_off: np.ndarray
    cl_ref: np.ndarray
    sg: np.ndarray
    mi_cnt: np.ndarray
    mi: np.ndarray          # (n_clusters, m, 2)
    rounds: int
    n_sketched_total: int
    n_grouped: int


@dataclass
class CombineResult:
    cl_n: np.ndarray
    cl_a_off: np.ndarray
    cl_a: np.ndarray
    cl_ref_off: np.ndarray
    cl_ref: np.ndarray
    iterations: int
    n_merges: int


@dataclass
class RealignResult:
    claim_contig: np.ndarray
    claim_sg: np.ndarray
    claim_y: np.ndarray
    claim_prio: np.ndarray
    fpA_sg: np.ndarray
    fpT_sg: np.ndarray
    n_windows: int
    n_probes: int
    n_candidates: int
    n_dict_keys: int
    numdict: int


class ReadSet:
    """mcb_readset: reads parsed from FASTQ (or packed from rows) into the device layout, on the host.
    Host code only; needs no GPU (the rows are page-locked when there is one)."""

    def __init__(self, readlen: int):
        self.lib = load_library()
        self.readlen = readlen
        h = C.c_void_p(0)
        self._check(self.lib.mcb_readset_create(readlen, C.byref(h)))
        self._h = h

    def _check(self, rc):
        if rc != 0:
            raise McbError(f"minicom_b200 error {rc}: {self.lib.mcb_last_error().decode()}")

    def _add(self, call, want_ascii):
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        a, n = C.c_void_p(0), C.c_uint64(0)
        self._check(call(C.byref(a) if want_ascii else None, C.byref(n)))
        if not want_ascii:
            return int(n.value)
        L = self.readlen
        rows = np.zeros((n.value, L), dtype=np.uint8)
        if n.value:
            blob = np.frombuffer((C.c_char * (n.value * (L + 1))).from_address(a.value), dtype=np.uint8).reshape(n.value, L + 1)
            assert (blob[:, L] == 0).all()
            rows[:] = blob[:, :L]
            libc.free(a)
        return rows

    def add_fastq(self, path: str, n_threads: int = 1, want_ascii: bool = False):
        """appends every record of the file; returns the number of reads added, or (want_ascii) their rows as the host stages see them"""
        return self._add(lambda a, n: self.lib.mcb_readset_add_fastq(self._h, path.encode(), n_threads, a, n), want_ascii)

    def add_fastq_buffer(self, data: bytes, n_threads: int = 1, want_ascii: bool = False):
        return self._add(lambda a, n: self.lib.mcb_readset_add_fastq_buffer(self._h, data, len(data), n_threads, a, n), want_ascii)

    def add_rows(self, rows: np.ndarray, n_threads: int = 1) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        assert rows.ndim == 2 and rows.shape[1] == self.readlen
        self._check(self.lib.mcb_readset_add_rows(self._h, rows.ctypes.data, rows.shape[0], n_threads))

    def view(self) -> _ReadsetView:
        v = _ReadsetView()
        self.lib.mcb_readset_get(self._h, C.byref(v))
        return v

    def arrays(self):
        """(packed u64[n][WS], nread_rid u32[nn], nmask u64[nn][WS]) as numpy views of the read set's own memory"""
        v = self.view()
        ws = v.row_words

        def arr(ptr, count, dt):
            if not count or not ptr:
                return np.zeros(0, dtype=dt)
            return np.frombuffer((C.c_char * (int(count) * np.dtype(dt).itemsize)).from_address(ptr), dtype=dt)
        return arr(v.packed, v.n_reads * ws, np.uint64).reshape(-1, ws), arr(v.nread_rid, v.n_nreads, np.uint32), arr(v.nmask, v.n_nreads * ws, np.uint64).reshape(-1, ws)

    def __len__(self):
        return int(self.view().n_reads)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.mcb_readset_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


class Index:
    def __init__(self, lib, handle):
        self._lib, self._h = lib, handle

    def get(self, x: int) -> np.ndarray:
        n = C.c_int(0)
        p = self._lib.mcb_idx_get(self._h, int(x), C.byref(n))
        return _view(p, n.value, np.uint64)

    def stats(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._lib.mcb_idx_stats(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def arrays(self):
        """(keys, kstart, post) of the whole index (copies)"""
        nk, npost = self.stats()
        k, ks, po = C.c_void_p(0), C.c_void_p(0), C.c_void_p(0)
        self._lib.mcb_idx_arrays(self._h, C.byref(k), C.byref(ks), C.byref(po), None)
        return _view(k.value, nk, np.uint64), _view(ks.value, nk + 1, np.uint32), _view(po.value, npost, np.uint64)

    def close(self):
        if self._h:
            self._lib.mcb_idx_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


class Context:
    """One device context = one run of the front end (mcb_ctx)."""

    def __init__(self, params: Params):
        self.lib = load_library()
        self.params = params
        h = C.c_void_p(0)
        self._check(self.lib.mcb_create(C.byref(params), C.byref(h)))
        self._h = h

    def _check(self, rc):
        if rc != 0:
            raise McbError(f"minicom_b200 error {rc}: {self.lib.mcb_last_error().decode()}")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.mcb_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- kt_for_reads
    def _reads_result(self, r: _ReadsResult) -> ReadsResult:
        nn = r.n_nreads
        off = _view(r.nread_off, nn + 1, np.uint64)
        return ReadsResult(_view(r.cls, r.n_reads, np.uint8), _view(r.nread_rid, nn, np.uint32), _view(r.nread_repl, nn, np.uint8),
                           off, _view(r.npos, int(off[-1]) if len(off) else 0, np.uint32), int(r.n_sketched))

    def for_reads(self, rows: np.ndarray) -> ReadsResult:
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        assert rows.ndim == 2 and rows.shape[1] == self.params.readlen
        r = _ReadsResult()
        self._check(self.lib.mcb_for_reads(self._h, rows.ctypes.data, rows.shape[0], C.byref(r)))
        return self._reads_result(r)

    def for_reads_packed(self, packed, nread_rid=None, nmask=None) -> ReadsResult:
        """packed: a ReadSet, or u64[n][WS] rows with the side table of the reads that contain N"""
        r = _ReadsResult()
        if isinstance(packed, ReadSet):
            v = packed.view()
            assert v.readlen == self.params.readlen
            self._check(self.lib.mcb_for_reads_packed(self._h, v.packed, v.n_reads, v.nread_rid, v.nmask, v.n_nreads, C.byref(r)))
            return self._reads_result(r)
        packed = np.ascontiguousarray(packed, dtype=np.uint64)
        nread_rid = np.ascontiguousarray(nread_rid if nread_rid is not None else np.zeros(0), dtype=np.uint32)
        nmask = np.ascontiguousarray(nmask if nmask is not None else np.zeros(0), dtype=np.uint64)
        self._check(self.lib.mcb_for_reads_packed(self._h, packed.ctypes.data, packed.shape[0], nread_rid.ctypes.data, nmask.ctypes.data, len(nread_rid), C.byref(r)))
        return self._reads_result(r)

    def for_reads_packed_device(self, d_packed: int, n: int, d_nread_rid: int = 0, d_nmask: int = 0, n_nreads: int = 0) -> ReadsResult:
        """the same with the three arrays resident in device memory (raw device pointers)"""
        r = _ReadsResult()
        self._check(self.lib.mcb_for_reads_packed_device(self._h, d_packed, n, d_nread_rid, d_nmask, n_nreads, C.byref(r)))
        return self._reads_result(r)

    def for_reads_device(self, dptr: int, n: int) -> ReadsResult:
        r = _ReadsResult()
        self._check(self.lib.mcb_for_reads_device(self._h, dptr, n, C.byref(r)))
        return self._reads_result(r)

    def for_reads_ptrs(self, ptrs: np.ndarray, n_threads: int = 1) -> ReadsResult:
        ptrs = np.ascontiguousarray(ptrs, dtype=np.uint64)
        r = _ReadsResult()
        self._check(self.lib.mcb_for_reads_ptrs(self._h, ptrs.ctypes.data, 8, len(ptrs), n_threads, C.byref(r)))
        return self._reads_result(r)

    def debug_read_tuples(self, n: int) -> np.ndarray:
        out = np.zeros((n, 2), dtype=np.uint64)
        self._check(self.lib.mcb_debug_read_tuples(self._h, out.ctypes.data))
        return out

    def debug_sketch_two(self, rids: np.ndarray, k: int) -> np.ndarray:
        rids = np.ascontiguousarray(rids, dtype=np.uint32)
        out = np.zeros((len(rids), 2), dtype=np.uint64)
        self._check(self.lib.mcb_debug_sketch_two(self._h, rids.ctypes.data, len(rids), k, out.ctypes.data))
        return out

    def debug_unpack_reads(self, n: int) -> np.ndarray:
        out = np.zeros((n, self.params.readlen), dtype=np.uint8)
        self._check(self.lib.mcb_debug_unpack_reads(self._h, out.ctypes.data))
        return out

    # -- print_encode's per-read encoding (N2, first slice)
    def dump_encode(self, members, member_off, refs=None, ref_off=None):
        """(enc_off u64[n+1], enc bytes) for members y = rid<<32 | pos<<1 | dir given contig-major with member_off[n_contigs+1]"""
        members = np.ascontiguousarray(members, dtype=np.uint64)
        member_off = np.ascontiguousarray(member_off, dtype=np.uint64)
        r = _EncodeResult()
        if refs is None:
            self._check(self.lib.mcb_dump_encode(self._h, members.ctypes.data, member_off.ctypes.data, None, None, len(member_off) - 1, C.byref(r)))
        else:
            refs = np.ascontiguousarray(refs, dtype=np.uint8)
            ref_off = np.ascontiguousarray(ref_off, dtype=np.uint64)
            self._check(self.lib.mcb_dump_encode(self._h, members.ctypes.data, member_off.ctypes.data, refs.ctypes.data, ref_off.ctypes.data, len(ref_off) - 1, C.byref(r)))
        return _view(r.enc_off, r.n_members + 1, np.uint64), _view(r.enc, r.n_bytes, np.uint8).tobytes()

    # -- kt_for_bucket
    def for_bucket(self) -> BucketResult:
        r = _BucketResult()
        self._check(self.lib.mcb_for_bucket(self._h, C.byref(r)))
        nc, m = r.n_clusters, self.params.first_mininum
        a_off = _view(r.cl_a_off, nc + 1, np.uint64)
        r_off = _view(r.cl_ref_off, nc + 1, np.uint64)
        return BucketResult(_view(r.cl_n, nc, np.uint32), a_off, _view(r.cl_a, int(a_off[-1]) if nc else 0, np.uint64), r_off,
                            _view(r.cl_ref, int(r_off[-1]) if nc else 0, np.uint8), _view(r.sg, r.n_sg, np.uint32),
                            _view(r.mi_cnt, nc, np.uint8), _view(r.mi, nc * m * 2, np.uint64).reshape(nc, m, 2),
                            int(r.rounds), int(r.n_sketched_total), int(r.n_grouped))

    # -- combine_cluster (the contig merge, on the seed contigs mcb_for_bucket left on the device)
    def combine(self, cbthreshold: int) -> CombineResult:
        r = _CombineResult()
        self._check(self.lib.mcb_combine(self._h, cbthreshold, C.byref(r)))
        nc = r.n_clusters
        a_off = _view(r.cl_a_off, nc + 1, np.uint64)
        r_off = _view(r.cl_ref_off, nc + 1, np.uint64)
        return CombineResult(_view(r.cl_n, nc, np.uint32), a_off, _view(r.cl_a, int(a_off[-1]) if nc else 0, np.uint64), r_off,
                             _view(r.cl_ref, int(r_off[-1]) if nc else 0, np.uint8), int(r.iterations), int(r.n_merges))

    # -- mm_idx_generation
    def idx_build(self, tuples: np.ndarray, bucket_off: np.ndarray) -> Index:
        tuples = np.ascontiguousarray(tuples, dtype=np.uint64)
        bucket_off = np.ascontiguousarray(bucket_off, dtype=np.uint64)
        assert len(bucket_off) == (1 << self.params.b) + 1
        h = C.c_void_p(0)
        self._check(self.lib.mcb_idx_build(self._h, tuples.ctypes.data, bucket_off.ctypes.data, C.byref(h)))
        return Index(self.lib, h)

    # -- realign_hash
    def realign(self, sg: np.ndarray, refs: np.ndarray, ref_off: np.ndarray, threshold: int, maxsearch: int, ininumdict: int = 0) -> RealignResult:
        sg = np.ascontiguousarray(sg, dtype=np.uint32)
        r = _RealignResult()
        if refs is None:      # reuse the contigs of the previous call (see include/minicom_b200.h)
            self._check(self.lib.mcb_realign(self._h, sg.ctypes.data, len(sg), None, None, 0, threshold, maxsearch, ininumdict, C.byref(r)))
        else:
            refs = np.ascontiguousarray(refs, dtype=np.uint8)
            ref_off = np.ascontiguousarray(ref_off, dtype=np.uint64)
            self._check(self.lib.mcb_realign(self._h, sg.ctypes.data, len(sg), refs.ctypes.data, ref_off.ctypes.data, len(ref_off) - 1,
                                             threshold, maxsearch, ininumdict, C.byref(r)))
        return self._realign_result(r)

    def _bucket_result(self, r: _BucketResult) -> BucketResult:
        nc, m = r.n_clusters, self.params.first_mininum
        a_off = _view(r.cl_a_off, nc + 1, np.uint64)
        r_off = _view(r.cl_ref_off, nc + 1, np.uint64)
        return BucketResult(_view(r.cl_n, nc, np.uint32), a_off, _view(r.cl_a, int(a_off[-1]) if nc else 0, np.uint64), r_off,
                            _view(r.cl_ref, int(r_off[-1]) if nc else 0, np.uint8), _view(r.sg, r.n_sg, np.uint32),
                            _view(r.mi_cnt, nc, np.uint8), _view(r.mi, nc * m * 2, np.uint64).reshape(nc, m, 2),
                            int(r.rounds), int(r.n_sketched_total), int(r.n_grouped))

    def _realign_result(self, r: _RealignResult) -> RealignResult:
        n = r.n_claims
        return RealignResult(_view(r.claim_contig, n, np.uint32), _view(r.claim_sg, n, np.uint32), _view(r.claim_y, n, np.uint64),
                             _view(r.claim_prio, n, np.uint64), _view(r.fpA_sg, r.n_fpA, np.uint32), _view(r.fpT_sg, r.n_fpT, np.uint32),
                             int(r.n_windows), int(r.n_probes), int(r.n_candidates), int(r.n_dict_keys), int(r.numdict))

    # -- the same path over several GPUs (include/minicom_b200.h, "the same path over the GPUs of one box")
    def shard_init(self, unique_id: bytes, rank: int, n_ranks: int):
        assert len(unique_id) == NCCL_ID_BYTES
        self._check(self.lib.mcb_shard_init(self._h, unique_id, rank, n_ranks))

    def shard_begin(self, n_total: int, rid_base: int):
        self._check(self.lib.mcb_shard_begin(self._h, n_total, rid_base))

    def shard_for_bucket(self, cap_rounds: int = 64):
        """(this rank's BucketResult, per-round counts (n_rounds, 4): contigs, members, consensus bytes, singles)"""
        r = _BucketResult()
        rc = np.zeros(4 * cap_rounds, dtype=np.uint64)
        self._check(self.lib.mcb_shard_for_bucket(self._h, C.byref(r), rc.ctypes.data, cap_rounds))
        return self._bucket_result(r), rc.reshape(cap_rounds, 4)[:int(r.rounds)].copy()

    def shard_realign(self, sg, sg_index, n_sg_total, refs, ref_off, threshold, maxsearch, ininumdict=0) -> RealignResult:
        sg = np.ascontiguousarray(sg, dtype=np.uint32)
        sg_index = np.ascontiguousarray(sg_index, dtype=np.uint32)
        assert len(sg) == len(sg_index)
        r = _RealignResult()
        if refs is None:
            self._check(self.lib.mcb_shard_realign(self._h, sg.ctypes.data, sg_index.ctypes.data, len(sg), n_sg_total, None, None, 0,
                                                   threshold, maxsearch, ininumdict, C.byref(r)))
        else:
            refs = np.ascontiguousarray(refs, dtype=np.uint8)
            ref_off = np.ascontiguousarray(ref_off, dtype=np.uint64)
            self._check(self.lib.mcb_shard_realign(self._h, sg.ctypes.data, sg_index.ctypes.data, len(sg), n_sg_total, refs.ctypes.data, ref_off.ctypes.data,
                                                   len(ref_off) - 1, threshold, maxsearch, ininumdict, C.byref(r)))
        return self._realign_result(r)

    # -- measurement
    def timers_enable(self, on=True):
        self.lib.mcb_timers_enable(self._h, 1 if on else 0)

    def timers_reset(self):
        self.lib.mcb_timers_reset(self._h)

    def timers(self) -> dict:
        need = self.lib.mcb_timers_dump(self._h, None, 0)
        buf = C.create_string_buffer(need + 16)
        self.lib.mcb_timers_dump(self._h, buf, need + 16)
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, cnt = line.rsplit(" ", 2)
            out[name] = (float(ms), int(cnt))
        return out

    def kernel_launches(self) -> int:
        return int(self.lib.mcb_kernel_launches(self._h))
