import sys, time, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from minicom_b200 import api
import oracle_lib as O
rng = np.random.default_rng(1)
for n in (1_000_000, 6_800_000):
    x = rng.integers(0, 1 << 62, size=n, dtype=np.uint64)
    xy = np.stack([x, np.arange(n, dtype=np.uint64)], axis=1)
    off, flat = O.bucket_major(xy)
    with api.Context(api.resolve_params(100)) as ctx:
        ctx.timers_enable(True)
        for rep in range(3):
            ctx.timers_reset()
            ix = ctx.idx_build(flat, off); ix.close()
            tm = ctx.timers()
            print(n, rep, {k: round(v[0], 3) for k, v in tm.items() if k.startswith("k:") or k == "idx_build"}, flush=True)
