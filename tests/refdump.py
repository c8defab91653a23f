"""TEST INFRASTRUCTURE: run the reference binary built under oracle/_ref (oracle/ref/build_ref.sh) on a seeded
synthetic read set and load the state dumps its link-time wrappers write (oracle/ref/mcref_wrap.cpp).

Results are cached under tests/_cache/<key>/ (git-ignored; it travels to the GPU box with the snapshot, so a cache
filled in the build container is reused there; when it is missing the binary is simply run again — it is a plain
x86-64 executable and does not need /root/reference at run time).
"""
from __future__ import annotations

import hashlib
import json
import os
import resource
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from minicom_b200 import synth  # noqa: E402

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
CACHE = os.path.join(ROOT, "tests", "_cache")


def ref_binary(L: int, mode: str = "sg") -> str:
    return os.path.join(REF_DIR, f"minicom_ref_L{L}_{mode}")


def have_reference(L: int, mode: str = "sg") -> bool:
    return os.path.exists(ref_binary(L, mode))


def _unlimit_stack():
    try:
        resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))
    except (ValueError, OSError):
        pass


def dropin_binary(L: int, mode: str = "sg") -> str:
    return os.path.join(ROOT, "dropin", "_build", f"minicom_b200_L{L}_{mode}")


def run_reference(reads: np.ndarray, workdir: str, mode: str = "sg", env_opts: dict | None = None, threads: int = 1,
                  dump: bool = True, reads2: np.ndarray | None = None, exe: str | None = None) -> dict:
    """Runs the reference (or, with `exe`, the drop-in executable) on `reads` inside workdir;
    returns {'timing': {...}, 'dump': dir, 'out': dir}."""
    L = reads.shape[1]
    exe = exe or ref_binary(L, mode)
    if not os.path.exists(exe):
        raise FileNotFoundError(f"{exe} missing: run oracle/ref/build_ref.sh {L} {mode} where /root/reference exists")
    os.makedirs(workdir, exist_ok=True)
    fq = os.path.join(workdir, "in.fastq")
    synth.write_fastq(fq, reads)
    args = [exe, fq]
    if reads2 is not None:
        fq2 = os.path.join(workdir, "in2.fastq")
        synth.write_fastq(fq2, reads2)
        args.append(fq2)
    out = os.path.join(workdir, "out")
    dumpdir = os.path.join(workdir, "dump")
    tmpd = os.path.join(workdir, "tmp") + "/"
    for d in (out, dumpdir, tmpd):
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
    args.append(out)
    env = dict(os.environ)
    env.update({"MC_T": str(threads), "MC_TMPDIR": tmpd, "MC_TIMING": os.path.join(workdir, "timing.json"), "OMP_NUM_THREADS": str(threads),
                "MCB_TIMING": os.path.join(workdir, "timing.json")})
    if dump:
        env.update({"MC_DUMP": dumpdir})
        if dump != "noseqs":             # the N-replaced reads as text: as large as the input
            env["MC_DUMP_SEQS"] = "1"
    for k, v in (env_opts or {}).items():
        env[k] = str(v)
    p = subprocess.run(args, env=env, cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, preexec_fn=_unlimit_stack)
    if p.returncode != 0:
        raise RuntimeError(f"reference failed ({p.returncode}): {p.stdout.decode()[-2000:]}")
    with open(os.path.join(workdir, "timing.json")) as f:
        timing = json.load(f)
    os.remove(fq)
    return {"timing": timing, "dump": dumpdir, "out": out, "log": p.stdout.decode()}


def _cache_dir(n_reads, L, G, seed, special, mode, env_opts, dump):
    key = json.dumps([n_reads, L, G, seed, special, mode, sorted((env_opts or {}).items())] + ([] if dump else ["nodump"]))
    return key, os.path.join(CACHE, hashlib.sha1(key.encode()).hexdigest()[:16])


def have_cached(n_reads, L, G, seed, special=0.0, mode="sg", env_opts=None, dump=True) -> bool:
    return os.path.exists(os.path.join(_cache_dir(n_reads, L, G, seed, special, mode, env_opts, dump)[1], "ok"))


def cached_reference(n_reads: int, L: int, G: int, seed: int, special: float = 0.0, mode: str = "sg", env_opts: dict | None = None,
                     dump: bool = True):
    """(reads, Dump) for a seeded synthetic set; the reference is run once per key.  With dump=False only the
    output directory is kept (for the large cases) and (reads, out_dir) is returned."""
    key, wd = _cache_dir(n_reads, L, G, seed, special, mode, env_opts, dump)
    reads = synth.make_reads(n_reads, L, G, seed=seed, special=special)
    if not os.path.exists(os.path.join(wd, "ok")):
        shutil.rmtree(wd, ignore_errors=True)
        run_reference(reads, wd, mode=mode, env_opts=env_opts, dump=dump)
        with open(os.path.join(wd, "ok"), "w") as f:
            f.write(key)
    if not dump:
        return reads, os.path.join(wd, "out")
    return reads, Dump(os.path.join(wd, "dump"))


class Dump:
    """Accessors over the flat little-endian arrays written by mcref_wrap.cpp."""

    def __init__(self, path: str):
        self.path = path

    def has(self, name):
        return os.path.exists(os.path.join(self.path, name))

    def arr(self, name, dtype=None):
        if dtype is None:
            dtype = {"u64": np.uint64, "u32": np.uint32, "u8": np.uint8}[name.rsplit(".", 1)[1]]
        return np.fromfile(os.path.join(self.path, name), dtype=dtype)

    def buckets(self, prefix):
        cnt = self.arr(prefix + "_counts.u64")
        xy = self.arr(prefix + "_xy.u64").reshape(-1, 2)
        off = np.zeros(len(cnt) + 1, dtype=np.uint64)
        np.cumsum(cnt, out=off[1:])
        return off, xy

    def clusters(self, prefix):
        n = self.arr(prefix + "_n.u64")
        a = self.arr(prefix + "_a.u64")
        rl = self.arr(prefix + "_reflen.u64")
        ref = self.arr(prefix + "_ref.u8")
        a_off = np.zeros(len(n) + 1, dtype=np.uint64)
        np.cumsum(n, out=a_off[1:])
        r_off = np.zeros(len(n) + 1, dtype=np.uint64)
        np.cumsum(rl, out=r_off[1:])
        return {"n": n, "a": a, "a_off": a_off, "ref": ref, "ref_off": r_off, "pert": self.arr(prefix + "_pert.u64")}

    def postings(self, j):
        """list of (x, ys) in dump order (bucket-major, keys ascending)."""
        raw = self.arr(f"i{j}_post.u64")
        out, i = [], 0
        while i < len(raw):
            x, n = int(raw[i]), int(raw[i + 1])
            out.append((x, raw[i + 2:i + 2 + n]))
            i += 2 + n
        return out

    def n_idx(self):
        j = 0
        while self.has(f"i{j}_post.u64"):
            j += 1
        return j

    def n_realign(self):
        j = 0
        while self.has(f"h{j}_thr.u64"):
            j += 1
        return j

    def raw(self, name) -> bytes:
        with open(os.path.join(self.path, name), "rb") as f:
            return f.read()

    def seqs(self):
        return [l for l in self.raw("r_seqs.txt").split(b"\n") if l]

    def npos(self, n_reads):
        raw = self.arr("r_npos.u32")
        out, i = [], 0
        for _ in range(n_reads):
            c = int(raw[i])
            out.append(raw[i + 1:i + 1 + c])
            i += 1 + c
        return out


class GoldenDump(Dump):
    """Same accessors over a committed fixture tests/golden/<name>.npz (written by tests/golden/make_golden.py)."""

    def __init__(self, npz):
        self.npz = npz
        self.path = None

    def has(self, name):
        return "dump/" + name in self.npz.files

    def arr(self, name, dtype=None):
        if dtype is None:
            dtype = {"u64": np.uint64, "u32": np.uint32, "u8": np.uint8}[name.rsplit(".", 1)[1]]
        return np.frombuffer(self.npz["dump/" + name].tobytes(), dtype=dtype)

    def raw(self, name) -> bytes:
        return self.npz["dump/" + name].tobytes()


GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))


def load_golden(name):
    """(reads, meta dict, GoldenDump, {output file name: bytes})"""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = eval(z["meta"].tobytes().decode(), {"__builtins__": {}})
    out = {k[4:]: z[k].tobytes() for k in z.files if k.startswith("out/")}
    return z["reads"], meta, GoldenDump(z), out


def roundtrip(outdir: str, workdir: str, mode: str, reads: np.ndarray, reads2: np.ndarray | None = None, threads: int = 1) -> dict:
    """Feeds a pre-back-end output directory (the reference's or the drop-in's) to the reference's own decompressor
    (oracle/_ref/decompress, built from /root/reference/src/decompress.c; CLI as in minicom:383) and checks what `minicom -d`
    promises: default mode gives the input reads back as a multiset (minicom:389 concatenates the part files), -p gives them
    back in the original order, -1/-2 gives the read PAIRS back as a multiset."""
    import subprocess
    dec = os.path.join(REF_DIR, "decompress")
    if not os.path.exists(dec):
        raise FileNotFoundError(f"{dec} missing: run oracle/ref/build_ref.sh where /root/reference exists")
    d = os.path.join(workdir, "dec_in")
    shutil.rmtree(d, ignore_errors=True)
    shutil.copytree(outdir, d)
    res, res2 = os.path.join(workdir, "dec.reads"), os.path.join(workdir, "dec2.reads")
    args = [dec, d, res, "true" if mode == "pe" else "false", "true" if mode == "order" else "false", str(threads)] + ([res2] if mode == "pe" else [])
    p = subprocess.run(args, cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, preexec_fn=_unlimit_stack)
    if p.returncode != 0:
        raise RuntimeError(f"decompress failed ({p.returncode}): {p.stdout.decode()[-2000:]}")

    def lines(path):
        with open(path, "rb") as f:
            return np.frombuffer(f.read(), dtype=np.uint8)

    def as_rows(blob, L):
        assert blob.size % (L + 1) == 0, "decoded file is not a whole number of reads"
        rows = blob.reshape(-1, L + 1)
        assert (rows[:, L] == 10).all(), "decoded reads are not newline-terminated rows"
        return rows[:, :L]

    def sorted_rows(a):
        return a[np.lexsort(a.T[::-1])]

    L = reads.shape[1]
    parts = lambda names: np.concatenate([as_rows(lines(os.path.join(d, f)), L) for f in names])          # noqa: E731
    results = sorted((f for f in os.listdir(d) if f.startswith("result_") and f.endswith(".seq")), key=lambda f: int(f[7:-4]))
    if mode == "order":
        got = as_rows(lines(res), L)
        assert got.shape == reads.shape and np.array_equal(got, reads), "order-preserving round trip: decoded reads differ from the input"
        return {"reads": len(got), "kind": "exact order"}
    if mode == "pe":
        # file 1 comes back as a multiset in cluster order (catsh.sh, decompress.c:1298-1308) and argv[6] holds the mates line
        # for line (decompress.c:1311-1315): the PAIRS are what must survive
        got1 = parts(["aatt.fasta", "single_dec.fasta"] + results)
        got2 = as_rows(lines(res2), L)
        assert got1.shape == reads.shape and got2.shape == reads2.shape, "paired-end round trip: read counts differ"
        assert np.array_equal(sorted_rows(np.hstack([got1, got2])), sorted_rows(np.hstack([reads, reads2]))), "paired-end round trip: decoded pairs differ from the input pairs"
        return {"reads": len(got1), "kind": "multiset of pairs"}
    got = parts(["aatt.fasta", "single_N.seq", "single_dec.fasta"] + results)   # minicom:389
    assert got.shape == reads.shape and np.array_equal(sorted_rows(got), sorted_rows(reads)), "round trip: decoded multiset differs from the input"
    return {"reads": len(got), "kind": "multiset"}
