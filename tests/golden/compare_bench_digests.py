#!/usr/bin/env python
"""TEST INFRASTRUCTURE: compare the digest VALUES a bench.py line carries (parity.values, written when no manifest existed at run
time) with a reference manifest made afterwards by make_manifests.py — no GPU needed.
usage: python tests/golden/compare_bench_digests.py profiles/r2_bench_C4_N1_n.json tests/golden/manifest_C4.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from minicom_b200 import parity  # noqa: E402


def main():
    with open(sys.argv[1]) as f:
        line = json.load(f)
    with open(sys.argv[2]) as f:
        man = json.load(f)
    got = line["parity"].get("values")
    if not got:
        raise SystemExit("the bench line carries no digest values (it was compared with a manifest at run time: see its parity field)")
    res = parity.compare(got, man["state"])
    res.update({"bench_line": sys.argv[1], "manifest": sys.argv[2], "counts_bench": {"claims": line["parity"]["claims"], "contigs": line["parity"]["contigs"], "iterations": line["parity"]["iterations"]},
                "counts_reference": {k: man["counts"][k] for k in ("claims", "contigs", "index_builds") if k in man["counts"]}})
    print(json.dumps(res, indent=1))
    return 0 if res["status"] == "ok" else 1


if __name__ == "__main__":
    sys.exit(main())
