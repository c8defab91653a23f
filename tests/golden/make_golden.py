#!/usr/bin/env python
"""TEST INFRASTRUCTURE: generates the committed golden fixtures tests/golden/*.npz.

Each fixture = one run of the UNMODIFIED reference (oracle/_ref/minicom_ref_L<readlen>_<mode>, built from
/root/reference/src by oracle/ref/build_ref.sh with num_thr=1, the only deterministic configuration) on a small seeded
synthetic read set, with the link-time state dumps of oracle/ref/mcref_wrap.cpp: bucket tuples after kt_for_reads,
read classes, N bookkeeping, N-replaced reads, seed contigs / singles / index tuples after kt_for_bucket, the inputs
and posting lists of every mm_idx_generation, the contigs after the host merge, and per realign_hash round the
singles, the claims in append order, sg_flag and the near-poly-A/T diversions.  The pre-back-end output directory is
stored too (file name -> bytes), so byte-identity of the drop-in can be checked where the reference is absent.

Run it where /root/reference exists:  python tests/golden/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refdump  # noqa: E402
from minicom_b200 import synth  # noqa: E402

# name -> (n_reads, L, genome, seed, special fraction, mode, reference env options)
CASES = {
    "g100_default": (3000, 100, 15000, 21, 0.02, "sg", {}),
    "g100_opts": (2500, 100, 12000, 22, 0.01, "sg", {"MC_K": 25, "MC_E": 6, "MC_M": 4, "MC_W": 12, "MC_S": 3, "MC_STEP": 3, "MC_EMAX": 30}),
    "g75_default": (2500, 75, 12000, 23, 0.01, "sg", {}),
    "g150_default": (2000, 150, 15000, 24, 0.01, "sg", {}),
    "g100_order": (2000, 100, 10000, 25, 0.01, "order", {}),
}


def main():
    for name, (n, L, G, seed, special, mode, env) in CASES.items():
        if not refdump.have_reference(L, mode):
            print(f"skip {name}: oracle/_ref/minicom_ref_L{L}_{mode} not built")
            continue
        reads = synth.make_reads(n, L, G, seed=seed, special=special)
        with tempfile.TemporaryDirectory() as wd:
            r = refdump.run_reference(reads, wd, mode=mode, env_opts=env, dump=True)
            blob = {"reads": reads, "meta": np.frombuffer(repr({"n": n, "L": L, "G": G, "seed": seed, "special": special, "mode": mode, "env": env}).encode(), dtype=np.uint8)}
            for fn in sorted(os.listdir(r["dump"])):
                blob["dump/" + fn] = np.fromfile(os.path.join(r["dump"], fn), dtype=np.uint8)
            for fn in sorted(os.listdir(r["out"])):
                blob["out/" + fn] = np.fromfile(os.path.join(r["out"], fn), dtype=np.uint8)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **blob)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
