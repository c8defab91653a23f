"""Multi-GPU parity: the sharded front end (minicom_b200/shard.py) against the single-GPU path, bit for bit.
Needs two GPUs (run with `gpurun --gpus 2`); skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    """devices the CUDA driver shows (no torch import in the test process)"""
    import ctypes
    try:
        cuda = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        if cuda.cuInit(0) != 0 or cuda.cuDeviceGetCount(ctypes.byref(n)) != 0:
            return 0
        return n.value
    except OSError:
        return 0


@pytest.mark.parametrize("world,extra", [(2, []), (2, ["--reads", "20011", "--readlen", "75", "--genome", "90000"]),
                                         (2, ["--bigbins"])])          # dictionary bins above maxsearch: the sequential replay, sharded
def test_sharded_front_end_matches_single_gpu(world, extra):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29611", "-m", "minicom_b200.shard_check"] + extra
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    out = p.stdout.decode()
    print(out[-4000:])
    assert p.returncode == 0 and "SHARD CHECK OK" in out


@pytest.mark.parametrize("name", ["C1o", "C3s", "C4t"])
def test_dropin_over_two_gpus_matches_reference_manifest(name):
    """The reference's own host program (minicommain.c behind dropin/mcb_dropin.cpp) with MCB_DEVICES=0,1: kt_for_reads,
    kt_for_bucket, every mm_idx_generation and every realign_hash run sharded over two GPUs inside the library, and the output
    directory must still carry the SHA-256s of the single-threaded reference (tests/golden/manifest_<name>.json)."""
    import hashlib
    import json
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_manifests
    import refdump
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    with open(os.path.join(ROOT, "tests", "golden", f"manifest_{name}.json")) as f:
        man = json.load(f)
    c = man["config"]
    exe = refdump.dropin_binary(c["L"], c["mode"])
    if not os.path.exists(exe):
        pytest.fail(f"{exe} not built")
    reads, reads2 = make_manifests.make_reads(c)
    with tempfile.TemporaryDirectory() as wd:
        r = refdump.run_reference(reads, wd, mode=c["mode"], env_opts=dict(c["env"], MCB_DEVICES="0,1"), threads=1, dump=False, reads2=reads2, exe=exe)
        got = {f: hashlib.sha256(open(os.path.join(r["out"], f), "rb").read()).hexdigest() for f in sorted(os.listdir(r["out"]))}
        assert sorted(got) == sorted(man["out"]), f"file sets differ: {set(got) ^ set(man['out'])}"
        bad = [f for f in got if got[f] != man["out"][f]]
        assert not bad, f"{name} over 2 GPUs: files differ from the reference's: {bad}"
    print(f"{name}: {len(got)} files byte-identical to the reference with MCB_DEVICES=0,1; timing {r['timing']}")
