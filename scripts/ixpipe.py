"""Wall clock of mcb_idx_build on synthetic bucket-major tuples in pinned memory (pipelined build diagnostics)."""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minicom_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
rng = np.random.default_rng(5)
x = rng.integers(0, 1 << 40, size=n // 3 + 1, dtype=np.uint64).repeat(3)[:n]
rng.shuffle(x)
order = np.argsort(x & np.uint64(16383), kind="stable")
x = x[order]
y = rng.integers(0, 1 << 50, size=n, dtype=np.uint64)
xy = np.stack([x, y], axis=1).reshape(-1)
off = np.zeros(16385, dtype=np.uint64)
off[1:] = np.cumsum(np.bincount((x & np.uint64(16383)).astype(np.int64), minlength=16384))
t = torch.from_numpy(xy).pin_memory(); xy = t.numpy()
C = api.C
with api.Context(api.resolve_params(100)) as ctx:
    for it in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h = C.c_void_p(0)
        ctx._check(ctx.lib.mcb_idx_build(ctx._h, xy.ctypes.data, off.ctypes.data, C.byref(h)))
        dt = (time.perf_counter() - t0) * 1e3
        nk, npost = C.c_uint64(0), C.c_uint64(0)
        ctx.lib.mcb_idx_stats(h, C.byref(nk), C.byref(npost))
        ctx.lib.mcb_idx_destroy(h)
        print(f"iter {it}: {dt:.3f} ms  keys {nk.value} post {npost.value}  (h2d {n*16/1e6:.0f} MB, d2h {(nk.value*12+npost.value*8)/1e6:.0f} MB)", file=sys.stderr)
