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
