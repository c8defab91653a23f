#!/usr/bin/env python
"""One line per launch of an `ncu --page raw --csv` export (scripts/gpu_profile.sh), plus per-kernel means.
usage: ncu_summary.py <raw.csv> [--md out.md] [--traffic traffic.json] [--instructions instructions.json --workload C2]"""
import argparse
import collections
import csv
import json
import re

COLS = {"ms": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
        "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed", "rd_pct": "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "wr_pct": "dram__bytes_write.sum.pct_of_peak_sustained_elapsed", "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "occ": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread", "winst": "smsp__inst_executed.sum",
        "lanes": "smsp__thread_inst_executed_per_inst_executed.ratio", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1hit": "l1tex__t_sector_hit_rate.pct"}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    return name.strip()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw")
    ap.add_argument("--md")
    ap.add_argument("--traffic")
    ap.add_argument("--instructions")
    ap.add_argument("--workload", default="C2")
    a = ap.parse_args()
    with open(a.raw) as f:
        r = csv.reader(l for l in f if not l.startswith("=="))
        head, units = next(r), next(r)
        idx = {}
        for k, v in COLS.items():                                   # exact name, else a section-prefixed copy of the metric
            hit = [i for i, h in enumerate(head) if h == v] or [i for i, h in enumerate(head) if h.endswith("." + v)]
            if hit:
                idx[k] = hit[0]
        ki, gi, bi = head.index("Kernel Name"), head.index("Grid Size"), head.index("Block Size")
        rows = []
        for row in r:
            d = {"kernel": short(row[ki]), "grid": row[gi], "block": row[bi]}
            for k, i in idx.items():
                try:
                    d[k] = float(row[i].replace(",", "")) * SCALE.get(units[i], 1)
                except ValueError:
                    d[k] = float("nan")
            for k in COLS:
                d.setdefault(k, float("nan"))
            if d["dram_pct"] != d["dram_pct"]:                      # the section-prefixed column is empty in some exports: read % + write %
                d["dram_pct"] = d["rd_pct"] + d["wr_pct"]
            rows.append(d)
    out = ["| # | kernel | grid | ms | DRAM MB (rd+wr) | DRAM % | SM % | issue % | lanes/inst | occupancy % | regs | M warp inst | L1 hit % |", "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for i, d in enumerate(rows):
        out.append(f"| {i} | `{d['kernel']}` | {d['grid']} | {d['ms']:.4f} | {(d['rd'] + d['wr']) / 1e6:.1f} | {d['dram_pct']:.1f} | {d['sm_pct']:.1f} | {d['issue']:.1f} | {d['lanes']:.1f} | {d['occ']:.1f} | {d['regs']:.0f} | {d['winst'] / 1e6:.2f} | {d['l1hit']:.1f} |")
    by = collections.defaultdict(list)
    for d in rows:
        by[d["kernel"]].append(d)
    out += ["", "Per kernel (all captured launches):", "", "| kernel | launches | total ms | mean DRAM MB | time-weighted DRAM % | time-weighted issue % | lanes/inst |", "|---|---|---|---|---|---|---|"]
    for k, v in sorted(by.items(), key=lambda kv: -sum(d["ms"] for d in kv[1])):
        t = sum(d["ms"] for d in v)
        tw = lambda f: sum(d[f] * d["ms"] for d in v) / t if t else 0        # noqa: E731
        out.append(f"| `{k}` | {len(v)} | {t:.3f} | {sum(d['rd'] + d['wr'] for d in v) / len(v) / 1e6:.1f} | {tw('dram_pct'):.1f} | {tw('issue'):.1f} | {tw('lanes'):.1f} |")
    text = "\n".join(out) + "\n"
    if a.md:
        with open(a.md, "w") as f:
            f.write(text)
    else:
        print(text)
    strip = lambda k: re.sub(r"<.*", "", k)[2:] if k.startswith("k_") else k     # noqa: E731
    if a.traffic:
        tr = {}
        for k, v in by.items():
            big = max(v, key=lambda d: d["ms"])                     # the largest launch of the kernel: what a full-size call moves
            tr.setdefault(strip(k), 0)
            tr[strip(k)] = max(tr[strip(k)], int(big["rd"] + big["wr"]))
        with open(a.traffic, "w") as f:
            json.dump({"_comment": f"dram__bytes_read.sum + dram__bytes_write.sum of the LARGEST captured launch of each kernel, workload {a.workload}, from {a.raw}; read by bench.py for roofline.traffic", **dict(sorted(tr.items()))}, f, indent=1)
    if a.instructions:
        ins = {"_comment": f"smsp__inst_executed.sum of the largest captured launch, workload {a.workload}, from {a.raw}; read by bench.py for roofline.issue", "workload": a.workload}
        for k, v in by.items():
            ins[strip(k)] = max(ins.get(strip(k), 0), int(max(v, key=lambda d: d["ms"])["winst"]))
        with open(a.instructions, "w") as f:
            json.dump(ins, f, indent=1)


if __name__ == "__main__":
    main()
