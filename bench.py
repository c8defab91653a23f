#!/usr/bin/env python
"""bench.py — reads/s of the sketch + index + overlap stage (BASELINE.json metric) on N B200s.

One "step" = one pass of the hot path over one batch of synthetic reads: every call the reference's host code makes
into the replaced functions during one compression run — kt_for_reads, kt_for_bucket, each mm_idx_generation, each
realign_hash of the -e/-S/-E schedule — issued through the C-ABI of libminicom_b200.so.

The host stages that sit BETWEEN those calls (contig merge, kthread_cb.c) are not part of the path; their outputs are
needed as inputs, so the workload is prepared once, untimed, by running the drop-in executable (reference host objects +
this library) with MCB_RECORD: it writes the host-side inputs of every index build / realign call, and each timed step
replays exactly that call sequence.  At N = 1 the recording run uses num_thr = 1, the reference's only deterministic
configuration, so a step has ONE right answer: before the timed region the results of one step (read classes, tuples, seed
contigs, singles, every index, the claims of every threshold round) are digested and compared with the manifest of the
single-threaded reference for the workload (tests/golden/manifest_<workload>.json) -> `parity` in the JSON line.  At N > 1 the
sharded results are merged and compared with one GPU running the whole job on the same inputs.

Two timed regions per run:
  value : device time (CUDA events inside the library, on its stream) of the four entry points with the reads resident
          in HBM; host<->device copies excluded.
  e2e   : wall clock of the same calls through the host-buffer C-ABI (pinned host rows in, results copied back to host
          memory every step).
`--impl reference` times the reference's own multithreaded CPU implementation (oracle/_ref, built from the unmodified
sources) on the SAME workload, all host threads; a step is minutes of CPU, so that arm caps itself at one timed step.
`cpu_baseline` inside our own line is the same binary on a bounded sample (named in its fields).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# name -> (reads per GPU, read length, genome length, mode, reference options)
WORKLOADS = {
    "C1": (1_000_000, 100, 5_000_000, "sg", {}),
    "C2": (10_000_000, 100, 50_000_000, "order", {}),
    "tiny": (100_000, 100, 500_000, "order", {}),
    # BASELINE.json configs[3] (150 bp, non-default -w/-s/-E, stepping -S): C4 is the named size, C4s a fifth of it
    "C4s": (10_000_000, 150, 75_000_000, "sg", {"MC_W": 20, "MC_S": 4, "MC_EMAX": 40, "MC_STEP": 2}),
    "C4": (50_000_000, 150, 375_000_000, "sg", {"MC_W": 20, "MC_S": 4, "MC_EMAX": 40, "MC_STEP": 2}),
}
# bounded sample of the workload for the CPU arm: same shape and coverage, fewer reads
CPU_SAMPLE = {"C2": (1_000_000, 100, 5_000_000), "C1": (1_000_000, 100, 5_000_000), "tiny": (100_000, 100, 500_000),
              "C4s": (500_000, 150, 3_750_000), "C4": (500_000, 150, 3_750_000)}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Native libraries (NCCL's version banner, the reference binary) also write to
# file descriptor 1, so the real stdout is set aside and everything else is sent to stderr.
_REAL_STDOUT = None


def claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def unlimit_stack():
    import resource
    try:
        resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))
    except (ValueError, OSError):
        pass


def run_binary(exe, fastq, workdir, env_extra, threads, fastq2=None):
    out, tmpd = os.path.join(workdir, "out"), os.path.join(workdir, "tmp") + "/"
    for d in (out, tmpd):
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
    timing = os.path.join(workdir, "timing.json")
    env = dict(os.environ)
    env.update({"MC_T": str(threads), "OMP_NUM_THREADS": str(threads), "MC_TMPDIR": tmpd, "MC_TIMING": timing, "MCB_TIMING": timing})
    env.update({k: str(v) for k, v in env_extra.items()})
    t0 = time.time()
    p = subprocess.run([exe, fastq] + ([fastq2] if fastq2 else []) + [out], env=env, cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, preexec_fn=unlimit_stack)
    if p.returncode != 0:
        raise RuntimeError(f"{exe} failed ({p.returncode}): {p.stdout.decode()[-3000:]}")
    with open(timing) as f:
        t = json.load(f)
    t["wall_total"] = time.time() - t0
    shutil.rmtree(out, ignore_errors=True)
    return t


def front_end_seconds(t):
    return t["kt_for_reads"] + t["kt_for_bucket"] + t["mm_idx_generation"] + t["realign_hash"]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.stop_flag, self.gpu = [], False, gpu_index
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=5).stdout.decode().strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.15)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop_flag = True
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows if len(r) > 3 + i)]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(kernel, c):
    """(algorithmic bytes per launch, launches per step) of one kernel: SURVEY.md 8(d) per-unit figures x the units one
    launch processes ("roofline accounting").  Kernels the survey has no term for get the bytes they cannot
    avoid moving (inputs read once + outputs written once)."""
    L4 = (c["L"] + 3) // 4
    rounds = c["rounds"]
    nr = max(1, len(rounds))
    if kernel in ("k:pack_classify_sketch", "k:classify_sketch_packed"):        # packed read + tuple per read
        return c["N"] * (L4 + 16), 1
    if kernel in ("k:sort_scatter", "k:sort_hist"):
        return None, None                          # filled by the caller: 32 B x tuples, spread over the passes
    if kernel == "k:consensus":                   # count + verify: 2 x packed read per grouped read
        return 2 * L4 * c["N_grp"] / max(1, c["bucket_rounds"]), c["bucket_rounds"]
    if kernel == "k:index_sort":                  # one read + one write of every 16-byte tuple
        return 32 * c["T_cb"] / max(1, c["n_idx"]), c["n_idx"]
    if kernel == "k:sketch_lh":                   # consensus characters in, first m tuples out
        return (c["seed_ref_bytes"] + 16 * c["m"] * c["clusters"]) / max(1, c["bucket_rounds"]), c["bucket_rounds"]
    if kernel in ("k:s2_kmer_sort_scatter", "k:s2_kmer_sort_hist"):   # 8-byte table entries, read + written once per pass
        return 16 * rounds[0]["R"] if rounds else None, None
    if kernel == "k:s2_probe_table":              # per single: packed read + 12 B per dictionary (key + bin slot)
        return sum(r["S"] * (L4 + 12 * r["nd"]) for r in rounds) / nr, nr
    if kernel == "k:s2_verify":                   # per candidate: the window and the single
        return sum(2 * L4 * r["C"] for r in rounds) / nr, nr
    if kernel == "k:s2_singles":
        return sum(r["S"] * 2 * L4 for r in rounds) / nr, nr
    return None, None


def issue_roofline(kernel, workload, avg_launch_s, sm_count, sm_mhz, path=None):
    """The dominant kernel against the INSTRUCTION-ISSUE roofline (it is not bandwidth bound): warp instructions
    per launch from the committed ncu capture (profiles/instructions.json, same workload) / the live CUDA-event launch time,
    against sm_count x 4 schedulers x the SM clock sampled during the timed region.  None when the capture has no figure."""
    path = path or os.path.join(ROOT, "profiles", "instructions.json")
    if not os.path.exists(path) or not avg_launch_s:
        return None
    with open(path) as f:
        tab = json.load(f)
    if tab.get("workload") != workload or kernel not in tab:
        return None
    peak = sm_count * 4 * (sm_mhz or 1965.0) * 1e6
    achieved = tab[kernel] / avg_launch_s
    return {"unit": "warp instructions/s", "achieved": round(achieved, 1), "peak": peak, "frac": round(achieved / peak, 4),
            "warp_instructions": tab[kernel], "source": "profiles/instructions.json (ncu smsp__inst_executed.sum, same workload, per step or per launch as the caller's time is) / live device time"}


def opts_text(ref_env):
    """the reference options of a workload in minicom's own flag names"""
    names = {"MC_K": "-k", "MC_E": "-e", "MC_M": "-m", "MC_W": "-w", "MC_S": "-s", "MC_EMAX": "-E", "MC_STEP": "-S"}
    extra = " ".join(f"{names[k]} {v}" for k, v in ref_env.items() if k in names)
    return "defaults k=31 m=6 e=4" + (f" with {extra}" if extra else "")


def bench_ours(args):
    import torch
    import torch.distributed as dist
    from minicom_b200 import api, synth
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, L, G, mode, ref_env = WORKLOADS[args.workload]
    if args.reads:
        n, G = args.reads, max(1000, args.reads * 5)
    threads = args.host_threads or min(os.cpu_count() or 1, 32) // max(1, world) or 1
    # the host stages of the recording run (contig merge) are deterministic with one thread only
    rec_threads = args.record_threads or 1
    exe = os.path.join(ROOT, "dropin", "_build", f"minicom_b200_L{L}_{mode}")
    if not os.path.exists(exe):
        raise SystemExit(f"{exe} missing: run __graft_entry__.build() where /root/reference exists (no CPU fallback)")
    # ---- prepare (untimed): synthetic reads, one drop-in run that records the call sequence
    t0 = time.time()
    reads = synth.make_reads(n, L, G, seed=1 + rank)          # every rank owns an independent read set (weak scaling)
    wd = tempfile.mkdtemp(prefix=f"mcb_bench_r{rank}_")
    rec = os.path.join(wd, "rec")
    os.makedirs(rec)
    fq = os.path.join(wd, "in.fastq")
    synth.write_fastq(fq, reads)
    log(f"[rank {rank}] workload {args.workload}: {n} x {L} bp synthesized in {time.time() - t0:.1f}s; running the drop-in executable once (num_thr={rec_threads}) to record the call sequence")
    dt = run_binary(exe, fq, wd, dict(ref_env, MCB_RECORD=rec, MCB_DEVICE=local, MCB_HOST_MERGE=1), rec_threads)
    os.remove(fq)
    log(f"[rank {rank}] drop-in run: front end {front_end_seconds(dt):.3f}s wall inside the entry points (first call includes CUDA context creation); whole program {dt['wall_total']:.1f}s")
    idx_calls, realign_calls = [], []
    keep_alive = []

    def pin(a):
        """the step's inputs live in page-locked host memory (bench contract): the copies inside the timed region are plain DMA"""
        t = torch.from_numpy(a).pin_memory()
        keep_alive.append(t)
        return t.numpy()

    j = 0
    while os.path.exists(os.path.join(rec, f"idx_{j}.off.u64")):
        idx_calls.append((pin(np.fromfile(os.path.join(rec, f"idx_{j}.xy.u64"), dtype=np.uint64)), np.fromfile(os.path.join(rec, f"idx_{j}.off.u64"), dtype=np.uint64)))
        j += 1
    j = 0
    while os.path.exists(os.path.join(rec, f"realign_{j}.meta.u64")):
        meta = np.fromfile(os.path.join(rec, f"realign_{j}.meta.u64"), dtype=np.uint64)
        realign_calls.append((pin(np.fromfile(os.path.join(rec, f"realign_{j}.sg.u32"), dtype=np.uint32)), pin(np.fromfile(os.path.join(rec, f"realign_{j}.refs.u8"), dtype=np.uint8)),
                              np.fromfile(os.path.join(rec, f"realign_{j}.off.u64"), dtype=np.uint64), int(meta[0]), int(meta[1]), int(meta[2])))
        j += 1
    shutil.rmtree(wd, ignore_errors=True)
    same_contigs = [j > 0 and np.array_equal(realign_calls[j][1], realign_calls[j - 1][1]) and np.array_equal(realign_calls[j][2], realign_calls[j - 1][2])
                    for j in range(len(realign_calls))]
    # pinned host rows (e2e arm) and device-resident rows (value arm)
    rows_pinned = torch.empty((n, L), dtype=torch.uint8, pin_memory=True)
    rows_pinned.numpy()[:] = reads
    rows_dev = rows_pinned.to("cuda", non_blocking=False)
    del reads
    params = api.resolve_params(L, device=local, **{k: int(ref_env[e]) for k, e in (("k", "MC_K"), ("e", "MC_E"), ("w", "MC_W"), ("m", "MC_M")) if e in ref_env})
    ctx = api.Context(params)
    ctx.timers_enable(True)
    counters = {}

    wall = {}

    def step(device_resident):
        C = api.C
        w = wall.setdefault("device" if device_resident else "host", {"for_reads": 0.0, "for_bucket": 0.0, "idx_build": 0.0, "realign": 0.0})
        t = time.perf_counter()
        rr = api._ReadsResult()
        if device_resident:
            ctx._check(ctx.lib.mcb_for_reads_device(ctx._h, rows_dev.data_ptr(), n, C.byref(rr)))
        else:
            ctx._check(ctx.lib.mcb_for_reads(ctx._h, rows_pinned.data_ptr(), n, C.byref(rr)))
        w["for_reads"] += time.perf_counter() - t
        t = time.perf_counter()
        br = api._BucketResult()
        ctx._check(ctx.lib.mcb_for_bucket(ctx._h, C.byref(br)))
        w["for_bucket"] += time.perf_counter() - t
        t = time.perf_counter()
        nc = int(br.n_clusters)
        ref_bytes = int(C.cast(br.cl_ref_off, C.POINTER(C.c_uint64))[nc]) if nc else 0
        mem = int(C.cast(br.cl_a_off, C.POINTER(C.c_uint64))[nc]) if nc else 0
        idx_d2h = 0
        for xy, off in idx_calls:
            h = C.c_void_p(0)
            ctx._check(ctx.lib.mcb_idx_build(ctx._h, xy.ctypes.data, off.ctypes.data, C.byref(h)))
            nk, npost = C.c_uint64(0), C.c_uint64(0)
            ctx.lib.mcb_idx_stats(h, C.byref(nk), C.byref(npost))
            idx_d2h += nk.value * 12 + npost.value * 8 + len(off) * 4
            ctx.lib.mcb_idx_destroy(h)
        w["idx_build"] += time.perf_counter() - t
        t = time.perf_counter()
        rounds = []
        for j, (sg, refs, off, thr, ms, nd) in enumerate(realign_calls):
            r = api._RealignResult()
            if same_contigs[j]:     # what the drop-in shim does: later rounds of the schedule reuse the contigs of the first
                ctx._check(ctx.lib.mcb_realign(ctx._h, sg.ctypes.data, len(sg), None, None, len(off) - 1, thr, ms, nd, C.byref(r)))
            else:
                ctx._check(ctx.lib.mcb_realign(ctx._h, sg.ctypes.data, len(sg), refs.ctypes.data, off.ctypes.data, len(off) - 1, thr, ms, nd, C.byref(r)))
            rounds.append({"S": len(sg), "R": int(len(refs)), "W": int(r.n_windows), "C": int(r.n_candidates), "nd": int(r.numdict), "claims": int(r.n_claims),
                           "probes": int(r.n_probes), "polyAT": int(r.n_fpA + r.n_fpT)})
        w["realign"] += time.perf_counter() - t
        counters.update({"N": n, "L": L, "m": int(params.first_mininum), "n_idx": len(idx_calls), "seed_ref_bytes": ref_bytes,
                         "N_sk": int(br.n_sketched_total), "N_grp": int(br.n_grouped), "bucket_rounds": int(br.rounds), "clusters": nc,
                         "singles_stage1": int(br.n_sg), "T_cb": int(sum(len(x[0]) // 2 for x in idx_calls)), "rounds": rounds})
        # bytes crossing PCIe in this step, counted from the arrays the library copies (inputs in, results out)
        h2d = (0 if device_resident else n * L) + sum(x[0].nbytes + x[1].nbytes for x in idx_calls) + sum(c[0].nbytes + (0 if same_contigs[j] else c[1].nbytes + 3 * c[2].nbytes) for j, c in enumerate(realign_calls))
        nn = int(rr.n_nreads)
        d2h = n + nn * (5 + params_ws * 8)                                                        # classes, N side table
        d2h += nc * (4 + 16 + 1 + 16 * params.first_mininum) + mem * 8 + ref_bytes + int(br.n_sg) * 4   # seed contigs, singles, index tuples
        d2h += idx_d2h + sum(r["claims"] * 16 + r["polyAT"] * 4 for r in rounds)
        return h2d, d2h

    params_ws = (((L + 31) // 32) + 1) & ~1

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def dev_ms(tm):
        return sum(tm.get(k, (0.0, 0))[0] for k in ("for_reads", "for_bucket", "idx_build", "realign"))

    par = parity_step(ctx, args, params, rows_pinned.numpy(), idx_calls, realign_calls, same_contigs, rec_threads) if world == 1 else {"status": "unchecked", "note": "replica mode"}
    log(f"[rank {rank}] parity: {json.dumps(par)}")
    for _ in range(args.warmup):
        step(True)
        step(False)
    # ---- timed region 1: device-resident ("value")
    wall.clear()
    ctx.timers_reset()
    barrier()
    with ClockSampler(local) as clk:
        w0 = time.perf_counter()
        for _ in range(args.steps):
            step(True)
        barrier()
        wall_dev_arm = time.perf_counter() - w0
        tm = ctx.timers()
        launches = ctx.kernel_launches()
        # ---- timed region 2: host buffers in, host results out ("e2e")
        ctx.timers_reset()
        barrier()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            h2d, d2h = step(False)
        barrier()
        wall_e2e = time.perf_counter() - w0
        tm_e2e = ctx.timers()
    ms_dev = dev_ms(tm) / args.steps
    vals = torch.tensor([ms_dev, wall_e2e / args.steps * 1e3, wall_dev_arm / args.steps * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    ms_dev_max, ms_e2e_max, ms_wall_dev_max = (float(x) for x in vals.cpu())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (live CUDA-event times of the value arm)
    peak, peak_src = measured_peaks()
    kern = {k: v for k, v in tm.items() if k.startswith("k:")}
    dom = max(kern, key=lambda k: kern[k][0])
    dom_ms, dom_cnt = kern[dom]
    ab, per_step_launches = algorithmic_bytes(dom, counters)
    if ab is None and dom in ("k:sort_scatter", "k:sort_hist"):
        ab = 32.0 * counters["N_sk"] / max(1, dom_cnt / args.steps)
    if dom in ("k:s2_kmer_sort_scatter", "k:s2_kmer_sort_hist") and ab:
        pass                                                  # per pass (= per launch): every entry read and written once
    avg_launch_s = dom_ms / max(1, dom_cnt) / 1e3
    achieved = (ab / avg_launch_s / 1e9) if ab else None
    roof = {"bound": "hbm", "kernel": dom[2:], "achieved": round(achieved, 2) if achieved else None, "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 5) if achieved else None, "traffic": None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": ab, "avg_launch_ms": dom_ms / max(1, dom_cnt), "share_of_device_time": round(dom_ms / max(1e-9, dev_ms(tm)), 4)}
    # the whole path against SURVEY.md 8(d)'s formula A (what the reference's algorithm has to touch, per step)
    L4 = (L + 3) // 4
    A = (counters["N_sk"] * (L4 + 16) + 32 * (counters["N_sk"] + counters["T_cb"]) + 2 * L4 * counters["N_grp"]
         + sum(r["S"] * (L4 + 12 * r["nd"]) + (r["R"] + 3) // 4 + 8 * (2 * r["nd"] - 1) * r["W"] + L4 * r["C"] for r in counters["rounds"]))
    roof["whole_path"] = {"algorithmic_bytes_per_step": int(A), "achieved": round(A / (ms_dev_max / 1e3) / 1e9, 2), "frac": round(A / (ms_dev_max / 1e3) / 1e9 / peak, 5),
                          "note": "SURVEY 8(d) formula A over the device time of the step; its Stage-2 term charges 8 B for each of the reference's W*(2nd-1) window probes, which the inverted join never issues"}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        with open(prof) as f:
            roof["traffic"] = json.load(f).get(dom[2:])
    roof["issue"] = issue_roofline(dom[2:], args.workload, avg_launch_s, torch.cuda.get_device_properties(local).multi_processor_count, clk.summary().get("sm_mhz"))
    cpu = cpu_baseline(args.workload, 1, quiet=True) if not args.no_cpu_baseline else None
    value = n * world / (ms_dev_max / 1e3)
    line = {
        "metric": "reads/sec for sketch+index+overlap stage", "value": round(value, 1), "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_dev_max, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {n} x {L} bp reads, {G} bp random genome, 1% substitutions, mode {mode}, {opts_text(ref_env)}" + ("" if world == 1 else "; independent read set per GPU, no exchange"),
                   "l2": "inputs larger than L2 (reads %.0f MB per step)" % (n * L / 1e6), "timing": "value = CUDA-event device time of the four entry points (reads resident in HBM); e2e = wall clock through the host-buffer C-ABI",
                   "bases_per_s": round(value * L, 1), "wall_ms_per_step_device_arm": round(ms_wall_dev_max, 3),
                   "counters": {k: v for k, v in counters.items() if k != "rounds"}, "realign_rounds": counters["rounds"],
                   "device_ms_by_entry_point": {k: round(tm[k][0] / args.steps, 4) for k in ("for_reads", "for_bucket", "idx_build", "realign") if k in tm},
                   "kernel_ms_per_step": {k[2:]: round(v[0] / args.steps, 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])},
                   "host_wall_ms_per_step": {a: {k: round(v / args.steps * 1e3, 3) for k, v in d.items()} for a, d in wall.items()},
                   "dropin_first_run_front_end_s": round(front_end_seconds(dt), 4), "host_threads": threads},
        "e2e": {"value": round(n * world / (ms_e2e_max / 1e3), 1), "unit": "reads/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": round(ms_e2e_max, 3),
                "copy_ms_per_step": {k: round(tm_e2e[k][0] / args.steps, 3) for k in ("h2d", "d2h") if k in tm_e2e}},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
        "roofline": roof,
        "cpu_baseline": cpu,
        "parity": par,
    }
    emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def bench_pipeline(args):
    """N = 1 (default): the whole front end stays on the device.  A step = kt_for_reads, kt_for_bucket, combine_cluster (the contig
    merge: every mm_idx_generation of the run happens inside it, on device-resident tuples), and every realign_hash of the
    -e/-S/-E schedule against the merged contigs, which never leave the device.  The only recorded inputs are the singles of every
    threshold round (the reference's host computes them with updateSingle(), preprocess.c:243-255), taken from one untimed run of
    the drop-in executable; that run is deterministic at any thread count now that the merge is on the device."""
    import torch
    from minicom_b200 import api, parity, synth
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    n, L, G, mode, ref_env = WORKLOADS[args.workload]
    if args.reads:
        n, G = args.reads, max(1000, args.reads * 5)
    threads = args.host_threads or min(os.cpu_count() or 1, 32)
    exe = os.path.join(ROOT, "dropin", "_build", f"minicom_b200_L{L}_{mode}")
    if not os.path.exists(exe):
        raise SystemExit(f"{exe} missing: run __graft_entry__.build() where /root/reference exists (no CPU fallback)")
    t0 = time.time()
    reads = synth.make_reads(n, L, G, seed=1)
    wd = tempfile.mkdtemp(prefix="mcb_bench_")
    rec = os.path.join(wd, "rec")
    os.makedirs(rec)
    fq = os.path.join(wd, "in.fastq")
    synth.write_fastq(fq, reads)
    log(f"workload {args.workload}: {n} x {L} bp synthesized in {time.time() - t0:.1f}s; running the drop-in executable once (num_thr={threads}) to record the singles of every threshold round")
    dt = run_binary(exe, fq, wd, dict(ref_env, MCB_RECORD=rec, MCB_DEVICE=local), threads)
    os.remove(fq)
    log(f"drop-in run: front end {front_end_seconds(dt):.3f}s wall inside the entry points + contig merge {dt.get('combine_cluster', 0):.3f}s; whole program {dt['wall_total']:.1f}s")
    keep_alive = []

    def pin(a):
        t = torch.from_numpy(a).pin_memory()
        keep_alive.append(t)
        return t.numpy()

    _, realign_calls = load_recorded(rec, None)
    shutil.rmtree(wd, ignore_errors=True)
    rounds_in = [(pin(sg), int(len(off) - 1), thr, ms, nd) for sg, refs, off, thr, ms, nd in realign_calls]
    rec_contigs = (realign_calls[0][1], realign_calls[0][2]) if realign_calls else None
    del realign_calls
    params = api.resolve_params(L, device=local, **{k: int(ref_env[e]) for k, e in (("k", "MC_K"), ("e", "MC_E"), ("w", "MC_W"), ("m", "MC_M")) if e in ref_env})
    ws = (((L + 31) // 32) + 1) & ~1
    packed_input = not args.ascii
    if packed_input:
        # the reads as the library's FASTQ reader hands them to kt_for_reads (csrc/mcb_fastq.cu): 2-bit rows in
        # page-locked memory + the side table of the reads with N; made here, untimed, like the FASTQ load the metric excludes (8d)
        rs = api.ReadSet(L)
        t0 = time.time()
        rs.add_rows(reads, threads)
        log(f"reads packed on the host by {threads} threads in {time.time() - t0:.2f}s ({n * ws * 8 / 1e6:.0f} MB page-locked)")
        rsv = rs.view()
        pk, nrid_h, nmask_h = rs.arrays()
        pk_dev, nrid_dev, nmask_dev = torch.from_numpy(pk).to("cuda"), torch.from_numpy(nrid_h.view(np.int32).copy()).to("cuda"), torch.from_numpy(nmask_h.view(np.int64).copy()).to("cuda")
        rows_pinned = rows_dev = None
        reads_h2d = n * ws * 8 + int(rsv.n_nreads) * (4 + ws * 8)
    else:
        rows_pinned = torch.empty((n, L), dtype=torch.uint8, pin_memory=True)
        rows_pinned.numpy()[:] = reads
        rows_dev = rows_pinned.to("cuda", non_blocking=False)
        reads_h2d = n * L
    del reads

    def load_reads(ctx_):
        return ctx_.for_reads_packed(rs) if packed_input else ctx_.for_reads(rows_pinned.numpy())
    cbthr = int(ref_env.get("MC_CBTHR", 0)) or 2 * int(params.diff_threshold)          # minicommain.c:122-126
    ctx = api.Context(params)
    ctx.timers_enable(True)
    m = int(params.first_mininum)
    # ---- parity (untimed, copying API)
    got = {}
    rr = load_reads(ctx)
    tuples = ctx.debug_read_tuples(n)
    br = ctx.for_bucket()
    mi_valid = br.mi[np.arange(m)[None, :] < br.mi_cnt[:, None]]
    got.update(parity.stage1_digests(rr.cls, tuples, br.cl_n, br.cl_a, br.cl_ref, np.diff(br.cl_ref_off.astype(np.int64)), br.sg, mi_valid))
    del tuples
    bm = parity.bucket_major(mi_valid)                                                  # the first mm_idx_generation, rebuilt from the host side: posting order
    off0 = np.zeros((1 << 14) + 1, dtype=np.uint64)
    np.cumsum(np.bincount((bm[:, 0] & parity.NB_MASK).astype(np.int64), minlength=1 << 14), out=off0[1:])
    ix = ctx.idx_build(bm, off0)
    got.update(parity.index_digests(0, *ix.arrays()))
    ix.close()
    del bm, mi_valid, br
    load_reads(ctx)
    ctx.for_bucket()
    cr = ctx.combine(cbthr)
    got.update(parity.contig_digests(cr.cl_n, cr.cl_a, cr.cl_ref, np.diff(cr.cl_ref_off.astype(np.int64))))
    consistent = rec_contigs is not None and np.array_equal(cr.cl_ref, rec_contigs[0]) and np.array_equal(cr.cl_ref_off, rec_contigs[1])
    n_contigs, claims = len(cr.cl_n), []
    for j, (sg, nc, thr, ms, nd) in enumerate(rounds_in):
        r = ctx.realign(sg, None, None, thr, ms, nd)
        got.update(parity.realign_digests(j, sg, n_contigs, r.claim_contig, r.claim_sg, r.claim_y, r.fpA_sg, r.fpT_sg))
        claims.append(len(r.claim_y))
    par = {"status": "unchecked", "digests": len(got), "claims": claims, "iterations": cr.iterations, "contigs": n_contigs, "recorded_contigs_equal_merged": bool(consistent)}
    path = os.path.join(ROOT, "tests", "golden", f"manifest_{args.workload}.json")
    if not args.reads and os.path.exists(path):
        with open(path) as f:
            man = json.load(f)
        par.update(parity.compare(got, man["state"]))
        par.update({"manifest": os.path.relpath(path, ROOT), "against": "unmodified reference, num_thr=1, same seeded reads (tests/golden/make_manifests.py)",
                    "covers": "read classes, minimizer tuples, seed contigs, singles, index tuples, the first index (keys + posting order), the contigs after the device merge "
                              "(members, order, consensus), and per threshold round the singles, claims in append order, sg_flag and poly-A/T diversions"})
    else:
        # no manifest committed for this workload (yet): the digests themselves go into the line, so that they can be compared with a
        # reference manifest made later (tests/golden/make_manifests.py) without another GPU run
        par.update({"note": "no reference manifest for this workload at run time; digest values included", "values": got})
    log(f"parity: {json.dumps({k: v for k, v in par.items() if k != 'values'})}")
    del cr, got
    api.VIEW_COPY = False
    counters, wall = {}, {}

    def step(device_resident):
        C = api.C
        w = wall.setdefault("device" if device_resident else "host", {"for_reads": 0.0, "for_bucket": 0.0, "combine": 0.0, "realign": 0.0})
        t = time.perf_counter()
        rr = api._ReadsResult()
        if packed_input and device_resident:
            ctx._check(ctx.lib.mcb_for_reads_packed_device(ctx._h, pk_dev.data_ptr(), n, nrid_dev.data_ptr(), nmask_dev.data_ptr(), rsv.n_nreads, C.byref(rr)))
        elif packed_input:
            ctx._check(ctx.lib.mcb_for_reads_packed(ctx._h, rsv.packed, n, rsv.nread_rid, rsv.nmask, rsv.n_nreads, C.byref(rr)))
        elif device_resident:
            ctx._check(ctx.lib.mcb_for_reads_device(ctx._h, rows_dev.data_ptr(), n, C.byref(rr)))
        else:
            ctx._check(ctx.lib.mcb_for_reads(ctx._h, rows_pinned.data_ptr(), n, C.byref(rr)))
        w["for_reads"] += time.perf_counter() - t
        t = time.perf_counter()
        br = api._BucketResult()
        ctx._check(ctx.lib.mcb_for_bucket_keep(ctx._h, C.byref(br)))
        w["for_bucket"] += time.perf_counter() - t
        t = time.perf_counter()
        cr = api._CombineResult()
        ctx._check(ctx.lib.mcb_combine(ctx._h, cbthr, C.byref(cr)))
        nc = int(cr.n_clusters)
        mem = int(C.cast(cr.cl_a_off, C.POINTER(C.c_uint64))[nc]) if nc else 0
        ref_bytes = int(C.cast(cr.cl_ref_off, C.POINTER(C.c_uint64))[nc]) if nc else 0
        w["combine"] += time.perf_counter() - t
        t = time.perf_counter()
        rounds = []
        for sg, nc_rec, thr, ms, nd in rounds_in:
            r = api._RealignResult()
            ctx._check(ctx.lib.mcb_realign(ctx._h, sg.ctypes.data, len(sg), None, None, nc, thr, ms, nd, C.byref(r)))
            rounds.append({"S": len(sg), "R": ref_bytes, "W": int(r.n_windows), "C": int(r.n_candidates), "nd": int(r.numdict), "claims": int(r.n_claims),
                           "probes": int(r.n_probes), "polyAT": int(r.n_fpA + r.n_fpT)})
        w["realign"] += time.perf_counter() - t
        counters.update({"N": n, "L": L, "m": m, "n_idx": int(cr.iterations), "seed_ref_bytes": 0, "N_sk": int(br.n_sketched_total), "N_grp": int(br.n_grouped),
                         "bucket_rounds": int(br.rounds), "clusters": int(br.n_clusters), "singles_stage1": int(br.n_sg), "contigs": nc, "merges": int(cr.n_merges),
                         "merge_iterations": int(cr.iterations), "T_cb": int(cr.n_index_tuples), "rounds": rounds})
        nn = int(rr.n_nreads)
        h2d = (0 if device_resident else reads_h2d) + sum(x[0].nbytes for x in rounds_in)
        d2h = n + nn * (5 + (0 if packed_input else ws * 8)) + int(br.n_sg) * 4 + nc * (4 + 16) + mem * 8 + ref_bytes + sum(r["claims"] * 24 + r["polyAT"] * 4 for r in rounds)
        return h2d, d2h

    def metric_ms(tm):         # kt_for_reads + kt_for_bucket + all mm_idx_generation + all realign_hash
        return sum(tm.get(k, (0.0, 0))[0] for k in ("for_reads", "for_bucket", "idx_build", "realign"))

    for _ in range(args.warmup):
        step(True)
        step(False)
    wall.clear()
    ctx.timers_reset()
    torch.cuda.synchronize()
    with ClockSampler(local) as clk:
        w0 = time.perf_counter()
        for _ in range(args.steps):
            step(True)
        torch.cuda.synchronize()
        wall_dev_arm = time.perf_counter() - w0
        tm = ctx.timers()
        launches = ctx.kernel_launches()
        ctx.timers_reset()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            h2d, d2h = step(False)
        torch.cuda.synchronize()
        wall_e2e = time.perf_counter() - w0
        tm_e2e = ctx.timers()
    ms_dev = metric_ms(tm) / args.steps
    ms_merge = tm.get("combine", (0.0, 0))[0] / args.steps
    ms_e2e = wall_e2e / args.steps * 1e3
    ms_e2e_metric = sum(wall["host"][k] for k in ("for_reads", "for_bucket", "realign")) / args.steps * 1e3 + tm_e2e.get("idx_build", (0.0, 0))[0] / args.steps
    peak, peak_src = measured_peaks()
    kern = {k: v for k, v in tm.items() if k.startswith("k:")}
    path_kern = {k: v for k, v in kern.items() if not k.startswith("k:cb_")}              # the merge's own kernels are outside the metric
    dom = max(path_kern, key=lambda k: path_kern[k][0])
    dom_ms, dom_cnt = kern[dom]
    # algorithmic bytes of all launches of the kernel in one step / its device time in one step (= bytes per launch / average launch
    # duration when every launch does the same amount of work; several kernels launch once per round or per work list)
    ab, ab_launches = algorithmic_bytes(dom, counters)
    if ab is None and dom in ("k:sort_scatter", "k:sort_hist"):
        ab, ab_launches = 32.0 * counters["N_sk"], 1                 # the tuple sort: every 16-byte tuple read and written once
    step_bytes = ab * (ab_launches or 1) if ab else None
    launches_per_step = dom_cnt / args.steps
    avg_launch_s = dom_ms / max(1, dom_cnt) / 1e3
    achieved = (step_bytes / (dom_ms / args.steps / 1e3) / 1e9) if step_bytes else None
    roof = {"bound": "hbm", "kernel": dom[2:], "achieved": round(achieved, 2) if achieved else None, "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 5) if achieved else None, "traffic": None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": step_bytes / launches_per_step if step_bytes else None, "launches_per_step": launches_per_step,
            "avg_launch_ms": dom_ms / max(1, dom_cnt), "share_of_device_time": round(dom_ms / max(1e-9, metric_ms(tm)), 4)}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof) and args.workload == "C2":
        with open(prof) as f:
            per_step = json.load(f).get(dom[2:])                       # DRAM bytes of all launches of the kernel in one step (committed ncu capture)
        if per_step:
            roof["traffic"] = per_step / max(1.0, launches_per_step)
            roof["traffic_per_step"] = per_step
            roof["algorithmic_bytes_per_step"] = step_bytes
    # the same kernel against the instruction-issue roofline: warp instructions of its launches in one step (committed capture) / its live time per step
    roof["issue"] = issue_roofline(dom[2:], args.workload, dom_ms / args.steps / 1e3, torch.cuda.get_device_properties(local).multi_processor_count, clk.summary().get("sm_mhz"))
    cpu = cpu_baseline(args.workload, 1, quiet=True) if not args.no_cpu_baseline else None
    value = n / (ms_dev / 1e3)
    line = {
        "metric": "reads/sec for sketch+index+overlap stage", "value": round(value, 1), "unit": "reads/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_dev, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {n} x {L} bp reads, {G} bp random genome, 1% substitutions, mode {mode}, {opts_text(ref_env)}",
                   "step": "kt_for_reads, kt_for_bucket, combine_cluster on the device (all mm_idx_generation calls inside it), every realign_hash of the schedule; "
                           "seed contigs, index tuples, indexes and merged contigs never leave the device",
                   "reads_input": ("2-bit packed rows (%d B/read) + N side table, as the library's FASTQ reader (mcb_readset_add_fastq, SURVEY 8f N3) leaves them in page-locked memory" % (ws * 8)) if packed_input
                                  else "ASCII rows (%d B/read)" % L,
                   "l2": "inputs larger than L2 (reads %.0f MB per step)" % ((n * ws * 8 if packed_input else n * L) / 1e6),
                   "timing": "value = CUDA-event device time of kt_for_reads + kt_for_bucket + all index builds + all realign_hash rounds (SURVEY 8d: the contig merge is not part of the metric; "
                             "its device time is merge_ms_per_step and value_with_merge includes it), reads resident in HBM; e2e = wall clock of the whole step, merge included, through the host-buffer C-ABI",
                   "bases_per_s": round(value * L, 1), "merge_ms_per_step": round(ms_merge, 4), "value_with_merge": round(n / ((ms_dev + ms_merge) / 1e3), 1),
                   "wall_ms_per_step_device_arm": round(wall_dev_arm / args.steps * 1e3, 3),
                   "counters": {k: v for k, v in counters.items() if k != "rounds"}, "realign_rounds": counters["rounds"],
                   "device_ms_by_entry_point": {k: round(tm[k][0] / args.steps, 4) for k in ("for_reads", "for_bucket", "idx_build", "combine", "realign") if k in tm},
                   "kernel_ms_per_step": {k[2:]: round(v[0] / args.steps, 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])},
                   "host_wall_ms_per_step": {a: {k: round(v / args.steps * 1e3, 3) for k, v in d.items()} for a, d in wall.items()},
                   "dropin_run_s": {"front_end": round(front_end_seconds(dt), 3), "contig_merge": round(dt.get("combine_cluster", 0), 3), "whole_program": round(dt["wall_total"], 1)},
                   "host_threads": threads},
        "e2e": {"value": round(n / (ms_e2e_metric / 1e3), 1), "unit": "reads/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": round(ms_e2e_metric, 3),
                "copy_ms_per_step": {k: round(tm_e2e[k][0] / args.steps, 3) for k in ("h2d", "d2h") if k in tm_e2e},
                "whole_step_with_merge": {"value": round(n / (ms_e2e / 1e3), 1), "ms_per_step": round(ms_e2e, 3), "merge_wall_ms_per_step": round(wall["host"]["combine"] / args.steps * 1e3, 3)},
                "note": "the metric end to end (SURVEY 8d, and what the reference arm times): host wall clock of the kt_for_reads, kt_for_bucket and realign_hash calls through the host-buffer C-ABI "
                        "(packed reads uploaded from page-locked memory, results copied back) + the CUDA-event time of the index builds, which run inside the device merge; "
                        "whole_step_with_merge adds the contig merge (combine_cluster, outside the metric), i.e. the wall clock of the entire step; byte counts are the whole step's"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
        "roofline": roof,
        "cpu_baseline": cpu,
        "parity": par,
    }
    emit(line)
    ctx.close()


def parity_step(ctx, args, params, rows, idx_calls, realign_calls, same_contigs, rec_threads):
    """One untimed pass of the step through the copying API; every result is digested (minicom_b200/parity.py) and compared with the
    manifest of the single-threaded reference for this workload."""
    from minicom_b200 import parity
    got = {}
    rr = ctx.for_reads(rows)
    tuples = ctx.debug_read_tuples(len(rows))
    br = ctx.for_bucket()
    m = int(params.first_mininum)
    got.update(parity.stage1_digests(rr.cls, tuples, br.cl_n, br.cl_a, br.cl_ref, np.diff(br.cl_ref_off.astype(np.int64)), br.sg,
                                     br.mi[np.arange(m)[None, :] < br.mi_cnt[:, None]]))
    del tuples
    for j, (xy, off) in enumerate(idx_calls):
        ix = ctx.idx_build(xy, off)
        got.update(parity.index_digests(j, *ix.arrays()))
        ix.close()
    if realign_calls:
        got["contigs.ref"] = parity.digest(realign_calls[0][1])
        got["contigs.reflen"] = parity.digest(np.diff(realign_calls[0][2].astype(np.int64)).astype(np.uint64))
    claims = []
    for j, (sg, refs, off, thr, ms, nd) in enumerate(realign_calls):
        r = ctx.realign(sg, None if same_contigs[j] else refs, None if same_contigs[j] else off, thr, ms, nd)
        got.update(parity.realign_digests(j, sg, len(off) - 1, r.claim_contig, r.claim_sg, r.claim_y, r.fpA_sg, r.fpT_sg))
        claims.append(len(r.claim_y))
    path = os.path.join(ROOT, "tests", "golden", f"manifest_{args.workload}.json")
    out = {"status": "unchecked", "digests": len(got), "claims": claims}
    if args.reads or rec_threads != 1 or not os.path.exists(path):
        out["note"] = "no manifest applies (size override, multi-threaded recording, or none committed for this workload)"
        return out
    with open(path) as f:
        man = json.load(f)
    out.update(parity.compare(got, man["state"]))
    out.update({"manifest": os.path.relpath(path, ROOT), "against": "unmodified reference, num_thr=1, same seeded reads (tests/golden/make_manifests.py)",
                "covers": "read classes, minimizer tuples, seed contigs, singles, index tuples, every index (keys + posting order), the recorded contigs, "
                          "and per threshold round the singles, claims in append order, sg_flag and poly-A/T diversions"})
    return out


def load_recorded(rec, pin):
    idx_calls, realign_calls = [], []
    j = 0
    while os.path.exists(os.path.join(rec, f"idx_{j}.off.u64")):
        idx_calls.append((np.fromfile(os.path.join(rec, f"idx_{j}.xy.u64"), dtype=np.uint64), np.fromfile(os.path.join(rec, f"idx_{j}.off.u64"), dtype=np.uint64)))
        j += 1
    j = 0
    while os.path.exists(os.path.join(rec, f"realign_{j}.meta.u64")):
        meta = np.fromfile(os.path.join(rec, f"realign_{j}.meta.u64"), dtype=np.uint64)
        realign_calls.append((np.fromfile(os.path.join(rec, f"realign_{j}.sg.u32"), dtype=np.uint32), np.fromfile(os.path.join(rec, f"realign_{j}.refs.u8"), dtype=np.uint8),
                              np.fromfile(os.path.join(rec, f"realign_{j}.off.u64"), dtype=np.uint64), int(meta[0]), int(meta[1]), int(meta[2])))
        j += 1
    return idx_calls, realign_calls


def bench_sharded(args):
    """N > 1: ONE job over N x n reads, sharded as SURVEY.md 8e says (csrc/mcb_shard.cu, minicom_b200/shard.py): reads by read-id
    range; every round of kt_for_bucket each tuple travels with the packed row of its read to the owner of its bucket (NCCL
    grouped send/recv inside the library); index builds by bucket range; Stage 2 on the rank that owns the single, against all
    contigs.  The host-side inputs (what the contig merger hands to mm_idx_generation / realign_hash) come from one untimed
    drop-in run of the whole job on rank 0.  Before the timed region the ranks' results of one step are merged and compared,
    digest by digest, with ONE GPU running the whole job on the same inputs -> `parity`."""
    import torch
    import torch.distributed as dist
    from minicom_b200 import api, parity, shard, synth
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)       # launcher plumbing: barriers, the max over ranks, the unique id
    n, L, G, mode, ref_env = WORKLOADS[args.workload]
    if args.reads:
        n, G = args.reads, max(1000, args.reads * 5)
    n_total, G_total = n * world, G * world
    threads = args.host_threads or min(os.cpu_count() or 1, 32)
    exe = os.path.join(ROOT, "dropin", "_build", f"minicom_b200_L{L}_{mode}")
    if not os.path.exists(exe):
        raise SystemExit(f"{exe} missing: run __graft_entry__.build() where /root/reference exists (no CPU fallback)")
    # ---- prepare (untimed)
    t0 = time.time()
    genome = synth.make_genome(G_total, seed=1)
    reads = synth.make_reads(n, L, G_total, seed=101 + rank, genome=genome)      # rank r holds read ids [r*n, (r+1)*n) of the job
    del genome
    box = [tempfile.mkdtemp(prefix="mcb_bench_shard_") if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    wd = box[0]
    # all ranks fill ONE FASTQ side by side (fixed-size records): no concatenation pass
    fq = os.path.join(wd, "in.fastq")
    if rank == 0:
        with open(fq, "wb") as f:
            f.truncate(n_total * synth.fastq_record_bytes(L))
    dist.barrier()
    synth.write_fastq(fq, reads, first_record=rank * n)
    dist.barrier()
    rec = os.path.join(wd, "rec")
    if rank == 0:
        os.makedirs(rec)
        log(f"[rank 0] job: {n_total} x {L} bp over {world} GPUs, inputs ready in {time.time() - t0:.1f}s; one untimed drop-in run records the call sequence")
        dt = run_binary(exe, fq, wd, dict(ref_env, MCB_RECORD=rec, MCB_DEVICE=local, MCB_HOST_MERGE=1), threads)     # host merge: its mm_idx_generation calls are what the ranks replay
        os.remove(fq)
        log(f"[rank 0] drop-in run: front end {front_end_seconds(dt):.3f}s inside the entry points; whole program {dt['wall_total']:.1f}s")
    dist.barrier()
    idx_calls, realign_calls = load_recorded(rec, None)
    dist.barrier()
    keep = []

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        keep.append(t)
        return t.numpy()

    ws = (((L + 31) // 32) + 1) & ~1
    if args.ascii:
        rows_pinned = torch.empty((n, L), dtype=torch.uint8, pin_memory=True)
        rows_pinned.numpy()[:] = reads
        rows_dev = rows_pinned.to(dev)
        in_host, in_dev, reads_h2d = rows_pinned.numpy(), rows_dev, n * L
    else:
        # this rank's reads as the library's FASTQ reader leaves them: 2-bit rows in page-locked memory + N side table
        rs = api.ReadSet(L)
        rs.add_rows(reads, max(1, threads // world))
        pk, nrid_h, nmask_h = rs.arrays()
        keep_dev = (torch.from_numpy(pk).to(dev), torch.from_numpy(nrid_h.view(np.int32).copy()).to(dev), torch.from_numpy(nmask_h.view(np.int64).copy()).to(dev))
        in_host, in_dev = rs, (keep_dev[0].data_ptr(), keep_dev[1].data_ptr(), keep_dev[2].data_ptr(), len(nrid_h))
        reads_h2d = n * ws * 8 + len(nrid_h) * (4 + ws * 8)
    params = api.resolve_params(L, device=local, **{k: int(ref_env[e]) for k, e in (("k", "MC_K"), ("e", "MC_E"), ("w", "MC_W"), ("m", "MC_M")) if e in ref_env})
    ctx = api.Context(params)
    ctx.timers_enable(True)
    fe = shard.ShardedFrontEnd(ctx, rank, world, shard.make_unique_id(dist))
    m = int(params.first_mininum)
    # ---- parity pass (untimed, copying API): this rank's results -> files; rank 0 merges them and compares with one GPU
    pdir = os.path.join(wd, "parity")
    if rank == 0:
        os.makedirs(pdir)
    dist.barrier()
    rr, part = fe.stage1(in_dev, n_total, True)
    # this rank's share of every recorded call
    b0, b1 = shard.bucket_range(rank, world)
    my_idx = []
    for xy, off in idx_calls:
        lo, hi = int(off[b0]), int(off[b1])
        mine = np.full(shard.NB + 1, hi - lo, dtype=np.uint64)
        mine[:b0] = 0
        mine[b0:b1 + 1] = off[b0:b1 + 1] - np.uint64(lo)
        my_idx.append((pin(xy.reshape(-1, 2)[lo:hi]), mine))
    my_realign = []
    for j, (sg, refs, off, thr, ms, nd) in enumerate(realign_calls):
        same = j > 0 and np.array_equal(refs, realign_calls[j - 1][1]) and np.array_equal(off, realign_calls[j - 1][2])
        mine, pos = fe.select(sg)                                  # the singles of the job's list that this rank produced in Stage 1
        my_realign.append((pin(mine), pin(pos), len(sg), None if same else pin(refs), None if same else off, thr, ms, nd))
    blob = {"cls": rr.cls, "cl_n": part.cl_n, "cl_a": part.cl_a, "cl_ref": part.cl_ref, "cl_reflen": part.cl_reflen, "sg": part.sg, "mi": part.mi, "mi_cnt": part.mi_cnt, "rounds": part.rounds}
    for j, (xy, off) in enumerate(my_idx):
        ix = ctx.idx_build(xy, off)
        k, ks, po = ix.arrays()
        blob.update({f"ik{j}": k, f"ic{j}": np.diff(ks.astype(np.int64)).astype(np.uint32), f"ip{j}": po})
        ix.close()
    for j, (sgl, pos, n_job, refs, off, thr, ms, nd) in enumerate(my_realign):
        r = fe.realign(sgl, pos, n_job, refs, off, thr, ms, nd)
        blob.update({f"cc{j}": r.claim_contig, f"cs{j}": r.claim_sg, f"cy{j}": r.claim_y, f"cp{j}": r.claim_prio, f"fa{j}": r.fpA_sg, f"ft{j}": r.fpT_sg})
    np.savez(os.path.join(pdir, f"rank{rank}.npz"), **blob)
    np.save(os.path.join(pdir, f"rows{rank}.npy"), reads)
    del reads, blob
    dist.barrier()
    par = None
    if rank == 0:
        zs = [np.load(os.path.join(pdir, f"rank{q}.npz")) for q in range(world)]
        got = {}
        mg = shard.merge_stage1([shard.Stage1Part(z["cl_n"], z["cl_a"], z["cl_ref"], z["cl_reflen"], z["sg"], z["mi"], z["mi_cnt"], z["rounds"]) for z in zs])
        got.update({"cls": parity.digest(np.concatenate([z["cls"] for z in zs])), "seed.cl_n": parity.digest(mg.cl_n), "seed.cl_a": parity.digest(mg.cl_a),
                    "seed.cl_ref": parity.digest(mg.cl_ref), "seed.cl_reflen": parity.digest(mg.cl_reflen), "sg": parity.digest(mg.sg),
                    "seed.mi": parity.digest(parity.bucket_major(mg.mi[np.arange(m)[None, :] < mg.mi_cnt[:, None]]))})
        for j in range(len(idx_calls)):
            got.update({f"idx{j}.keys": parity.digest(np.concatenate([z[f"ik{j}"] for z in zs])), f"idx{j}.cnt": parity.digest(np.concatenate([z[f"ic{j}"] for z in zs])),
                        f"idx{j}.post": parity.digest(np.concatenate([z[f"ip{j}"] for z in zs]))})
        for j, (sg, refs, off, thr, ms, nd) in enumerate(realign_calls):
            gc, gs, gy, _ = shard.merge_claims([(z[f"cc{j}"], z[f"cs{j}"], z[f"cy{j}"], z[f"cp{j}"]) for z in zs])
            got.update(parity.realign_digests(j, sg, len(off) - 1, gc, gs, gy, np.sort(np.concatenate([z[f"fa{j}"] for z in zs])), np.sort(np.concatenate([z[f"ft{j}"] for z in zs]))))
        # the whole job on ONE GPU, same inputs
        all_rows = np.concatenate([np.load(os.path.join(pdir, f"rows{q}.npy")) for q in range(world)])
        one = api.Context(params)
        want = {}
        r1 = one.for_reads(all_rows)
        del all_rows
        bq = one.for_bucket()
        want.update(parity.stage1_digests(r1.cls, np.zeros((0, 2), np.uint64), bq.cl_n, bq.cl_a, bq.cl_ref, np.diff(bq.cl_ref_off.astype(np.int64)), bq.sg,
                                          bq.mi[np.arange(m)[None, :] < bq.mi_cnt[:, None]]))
        want.pop("tuples")
        for j, (xy, off) in enumerate(idx_calls):
            ix = one.idx_build(xy, off)
            want.update(parity.index_digests(j, *ix.arrays()))
            ix.close()
        for j, (sg, refs, off, thr, ms, nd) in enumerate(realign_calls):
            same = j > 0 and np.array_equal(refs, realign_calls[j - 1][1]) and np.array_equal(off, realign_calls[j - 1][2])
            r = one.realign(sg, None if same else refs, None if same else off, thr, ms, nd)
            want.update(parity.realign_digests(j, sg, len(off) - 1, r.claim_contig, r.claim_sg, r.claim_y, r.fpA_sg, r.fpT_sg))
        one.close()
        par = parity.compare(got, want)
        par.update({"against": f"one GPU running the whole job of {n_total} reads on the same inputs (library single-GPU path, itself checked against the reference manifests at N=1)",
                    "covers": "read classes, seed contigs, singles, index tuples, every index (keys + posting order), per threshold round claims in append order, sg_flag, poly-A/T diversions"})
        log(f"[rank 0] parity: {json.dumps(par)}")
        shutil.rmtree(wd, ignore_errors=True)
    dist.barrier()
    T_cb = int(sum(len(x[0]) // 2 for x in idx_calls))
    R_total = int(len(realign_calls[0][1])) if realign_calls else 0
    del idx_calls, realign_calls
    api.VIEW_COPY = False          # results are consumed (counted) before the next call: no second copy on the host
    stats = {}

    wall = {}

    def phase_barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def step(device_resident):
        w = wall.setdefault("device" if device_resident else "host", {"stage1": 0.0, "idx_build": 0.0, "realign": 0.0})
        # The phases with collectives start together: the index builds in between move their tuples and results over PCIe, whose
        # speed differs between ranks (shared host bridges), and without the barrier that skew would be booked as collective time
        # on the ranks that wait.  The barrier itself is outside every timer; the wall clock (e2e) contains it.
        phase_barrier()
        t = time.perf_counter()
        rr, part = fe.stage1(in_dev if device_resident else in_host, n_total, device_resident, keep_mask=True)
        w["stage1"] += time.perf_counter() - t
        t = time.perf_counter()
        for xy, off in my_idx:
            ctx.idx_build(xy, off).close()
        w["idx_build"] += time.perf_counter() - t
        phase_barrier()
        t = time.perf_counter()
        claims, rounds, d2h = 0, [], 0
        for sgl, pos, n_job, refs, off, thr, ms, nd in my_realign:
            r = fe.realign(sgl, pos, n_job, refs, off, thr, ms, nd)
            claims += len(r.claim_y)
            rounds.append({"S": len(sgl), "R": R_total, "W": int(r.n_windows), "C": int(r.n_candidates), "nd": int(r.numdict)})
            d2h += len(r.claim_y) * 24 + (len(r.fpA_sg) + len(r.fpT_sg)) * 4
        w["realign"] += time.perf_counter() - t
        d2h += len(rr.cls) + part.cl_n.nbytes + part.cl_a.nbytes + part.cl_ref.nbytes + len(part.cl_n) * (16 + 1 + 16 * m) + part.sg.nbytes
        d2h += sum(len(xy) * 8 + len(xy) * 12 + (shard.NB + 1) * 4 for xy, _ in my_idx)      # postings + (at most) one key/start per tuple + bucket table
        stats.update({"seed_contigs_rank0": int(len(part.cl_n)), "singles_rank0": int(len(part.sg)), "claims_rank0": claims, "bucket_rounds": int(len(part.rounds)),
                      "d2h": int(d2h),
                      "counters": {"N": n, "L": L, "m": m, "n_idx": len(my_idx), "seed_ref_bytes": int(part.cl_ref.nbytes), "N_sk": int(rr.n_sketched), "N_grp": int(len(part.cl_a)),
                                   "bucket_rounds": int(len(part.rounds)), "clusters": int(len(part.cl_n)), "T_cb": int(sum(len(xy) for xy, _ in my_idx)), "rounds": rounds}})

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def dev_ms(tm):
        """entry points (library CUDA events on its stream) + the collectives the library issues between them"""
        return sum(tm.get(k, (0.0, 0))[0] for k in ("for_reads", "for_bucket", "idx_build", "realign")) + coll_ms(tm)

    def coll_ms(tm):
        return sum(v[0] for k, v in tm.items() if k.startswith("nccl:"))

    for _ in range(args.warmup):
        step(True)
        step(False)
    wall.clear()
    ctx.timers_reset()
    barrier()
    with ClockSampler(local) as clk:
        w0 = time.perf_counter()
        for _ in range(args.steps):
            step(True)
        barrier()
        wall_dev_arm = time.perf_counter() - w0
        tm = ctx.timers()
        launches = ctx.kernel_launches()
        ctx.timers_reset()
        barrier()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            step(False)
        barrier()
        wall_e2e = time.perf_counter() - w0
        tm_e2e = ctx.timers()
    ms_dev = dev_ms(tm) / args.steps
    vals = torch.tensor([ms_dev, wall_e2e / args.steps * 1e3, wall_dev_arm / args.steps * 1e3, coll_ms(tm) / args.steps], dtype=torch.float64, device=dev)
    dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    ms_dev_max, ms_e2e_max, ms_wall_dev_max, coll_max = (float(x) for x in vals.cpu())
    per_rank = [None] * world
    dist.all_gather_object(per_rank, {"ms": round(ms_dev, 3), "entry": {k: round(tm[k][0] / args.steps, 3) for k in ("for_reads", "for_bucket", "idx_build", "realign") if k in tm},
                                      "coll": round(coll_ms(tm) / args.steps, 3), "sort_lsd_fallbacks": tm.get("sort_lsd_fallbacks", (0, 0))[1],
                                      "top": {k[2:]: round(v[0] / args.steps, 3) for k, v in sorted(((k, v) for k, v in tm.items() if k.startswith("k:")), key=lambda kv: -kv[1][0])[:8]}})
    if rank == 0:
        peak, peak_src = measured_peaks()
        kern = {k: v for k, v in tm.items() if k.startswith("k:")}
        dom = max(kern, key=lambda k: kern[k][0])
        value = n_total / (ms_dev_max / 1e3)
        dom_ms, dom_cnt = kern[dom]
        ab, _ = algorithmic_bytes(dom, stats["counters"])
        if ab is None and dom in ("k:sort_scatter", "k:sort_hist"):
            ab = 32.0 * stats["counters"]["N_sk"] / max(1, dom_cnt / args.steps)
        avg_s = dom_ms / max(1, dom_cnt) / 1e3
        roof = {"bound": "hbm", "kernel": dom[2:], "achieved": round(ab / avg_s / 1e9, 2) if ab else None, "peak": peak, "unit": "GB/s",
                "frac": round(ab / avg_s / 1e9 / peak, 5) if ab else None, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ab, "avg_launch_ms": dom_ms / max(1, dom_cnt), "note": "rank 0's dominant kernel and units"}
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            with open(prof) as f:
                roof["traffic"] = json.load(f).get(dom[2:])
        stats.pop("counters", None)
        sent = int(tm.get("nccl_bytes_sent", (0.0, 0))[0])
        line = {
            "metric": "reads/sec for sketch+index+overlap stage", "value": round(value, 1), "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_dev_max, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"{args.workload} x {world}: ONE job of {n_total} x {L} bp reads ({n} per GPU), {G_total} bp random genome, 1% substitutions, mode {mode}, {opts_text(ref_env)}",
                       "sharding": "reads by read-id range; every round each tuple + the packed row of its read go to the owner of its bucket (bucket*G>>14) in one grouped NCCL send/recv "
                                   "inside the library; index builds by bucket range; Stage 2 on the rank that owns the single, against all contigs (no claim crosses ranks); "
                                   "results bit-identical to one GPU (see parity)",
                       "reads_input": "ASCII rows" if args.ascii else "2-bit packed rows + N side table (the library's FASTQ reader, SURVEY 8f N3)",
                       "l2": "inputs larger than L2 (reads %.0f MB per GPU per step)" % (reads_h2d / 1e6),
                       "timing": "value = max over ranks of (CUDA-event time of the library entry points + CUDA-event time of the NCCL collectives it issues), reads resident in HBM; e2e = max over ranks of the wall clock with pinned host inputs and results copied back",
                       "bases_per_s": round(value * L, 1), "wall_ms_per_step_device_arm": round(ms_wall_dev_max, 3), "collective_ms_per_step": round(coll_max, 3),
                       "collective_ms_per_step_rank0": {k.split(":", 1)[1]: round(v[0] / args.steps, 4) for k, v in tm.items() if k.startswith("nccl:") or k.startswith("nccl_in:")},
                       "host_wall_ms_per_step_rank0": {a: {k: round(v / args.steps * 1e3, 3) for k, v in d.items()} for a, d in wall.items()},
                       "per_rank": per_rank,
                       "nccl_bytes_sent_per_step_rank0": int(sent // max(1, args.steps)), "T_cb": T_cb, "contig_bases": R_total, "rank0": stats,
                       "device_ms_by_entry_point_rank0": {k: round(tm[k][0] / args.steps, 4) for k in ("for_reads", "for_bucket", "idx_build", "realign") if k in tm},
                       "kernel_ms_per_step_rank0": {k[2:]: round(v[0] / args.steps, 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])[:14]}, "host_threads": threads},
            "e2e": {"value": round(n_total / (ms_e2e_max / 1e3), 1), "unit": "reads/s", "h2d_bytes_per_step": int(reads_h2d + sum(x[0].nbytes + x[1].nbytes for x in my_idx) + sum(c[0].nbytes + c[1].nbytes + (c[3].nbytes if c[3] is not None else 0) for c in my_realign)),
                    "d2h_bytes_per_step": stats.get("d2h"), "ms_per_step": round(ms_e2e_max, 3), "copy_ms_per_step_rank0": {k: round(tm_e2e[k][0] / args.steps, 3) for k in ("h2d", "d2h") if k in tm_e2e},
                    "note": "byte counts are rank 0's; each rank hands its own results to its host (merging the ranks' lists is the caller's, as in shard.merge_*)"},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": roof,
            "cpu_baseline": None,
            "parity": par,
        }
        emit(line)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


def cpu_baseline(workload, steps, quiet=False, warmup=0, full=False):
    """The reference's own CPU implementation (oracle/_ref, unmodified sources), all host threads.  full=False: a bounded sample of
    the workload (same shape and coverage, fewer reads) for the `cpu_baseline` field of our own line; full=True: the workload itself."""
    from minicom_b200 import synth
    wn, wL, wG, mode, ref_env = WORKLOADS[workload]
    n, L, G = (wn, wL, wG) if full else CPU_SAMPLE[workload]
    exe = os.path.join(ROOT, "oracle", "_ref", f"minicom_ref_L{L}_{mode}")
    if not os.path.exists(exe):
        return {"value": None, "unit": "reads/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {exe} not built"}
    cores = os.cpu_count() or 1
    threads = min(cores, 64)
    wd = tempfile.mkdtemp(prefix="mcb_cpu_")
    reads = synth.make_reads(n, L, G, seed=1)
    fq = os.path.join(wd, "in.fastq")
    synth.write_fastq(fq, reads)
    del reads
    secs, last = [], None
    for i in range(warmup + steps):
        t = run_binary(exe, fq, wd, ref_env, threads)
        last = t
        if i >= warmup:
            secs.append(front_end_seconds(t))
        if not quiet:
            log(f"[reference] step {i}: front end {front_end_seconds(t):.2f}s (reads {t['kt_for_reads']:.2f} bucket {t['kt_for_bucket']:.2f} idx {t['mm_idx_generation']:.2f} realign {t['realign_hash']:.2f}), host merge {t.get('host_combine', 0):.2f}s")
    shutil.rmtree(wd, ignore_errors=True)
    s = sum(secs) / len(secs)
    what = "the workload itself" if (n, L, G) == (wn, wL, wG) else f"a bounded sample of the workload ({n} of {wn} reads, same read length and 20x coverage)"
    return {"value": round(n / s, 1), "unit": "reads/s", "cores": threads, "kind": "reference",
            "sample": f"{what}: {n} x {L} bp reads over a {G} bp genome, -t {threads}, wall time inside kt_for_reads+kt_for_bucket+mm_idx_generation+realign_hash = {s:.2f}s",
            "sample_reads": n, "workload_reads": wn, "same_config": (n, L, G) == (wn, wL, wG), "seconds": round(s, 3),
            "phases_s": {k: round(last[k], 3) for k in ("kt_for_reads", "kt_for_bucket", "mm_idx_generation", "realign_hash", "host_combine") if k in last}}


def bench_reference(args):
    """The reference arm runs the SAME workload as ours (C2: 10 M reads).  One step is minutes of CPU, so whatever --steps / --warmup
    say, it times ONE step without warm-up (the reference is a cold-start batch program anyway) and reports steps = 1."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", 1))
    steps = max(1, min(args.steps, args.ref_steps))
    cpu = cpu_baseline(args.workload, steps, warmup=0, full=not args.ref_sample)
    wn, wL, wG, mode, _ = WORKLOADS[args.workload]
    line = {"impl": "reference", "metric": "reads/sec for sketch+index+overlap stage", "value": cpu["value"], "unit": "reads/s", "n_gpus": world, "steps": steps,
            "warmup": 0, "ms_per_step": cpu["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {cpu['sample_reads']} x {wL} bp reads, {wG if cpu['same_config'] else CPU_SAMPLE[args.workload][2]} bp random genome, 1% substitutions, mode {mode}, {opts_text(WORKLOADS[args.workload][4])}",
                       "same_config_as_ours": cpu["same_config"],
                       "note": f"unmodified reference (oracle/_ref), -t {cpu['cores']}; steps capped at {steps} (requested {args.steps}) because one step is minutes of CPU; "
                               "the reference is a single-process CPU program, so its value does not grow with n_gpus"},
            "cpu_baseline": cpu, "e2e": {"value": cpu["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="override reads per GPU (debug)")
    ap.add_argument("--host-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--record-threads", type=int, default=0, help="num_thr of the untimed recording run at N=1 (default 1: deterministic, manifest parity applies)")
    ap.add_argument("--ref-steps", type=int, default=1, help="--impl reference: cap on the timed steps (one step of C2 is minutes of CPU)")
    ap.add_argument("--ref-sample", action="store_true", help="--impl reference: time the bounded sample instead of the full workload")
    ap.add_argument("--replay", action="store_true", help="N = 1: replay the host merger's recorded mm_idx_generation / realign_hash calls (contig merge on the host, num_thr=1 recording) instead of the device-resident pipeline")
    ap.add_argument("--ascii", action="store_true", help="N = 1: hand kt_for_reads the reads as ASCII rows (mcb_for_reads) instead of the packed rows of the library's FASTQ reader")
    ap.add_argument("--replicas", action="store_true", help="N > 1: run N independent single-GPU jobs instead of one sharded job")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        bench_reference(args)
    elif int(os.environ.get("WORLD_SIZE", 1)) > 1 and not args.replicas:
        bench_sharded(args)
    elif args.replay or int(os.environ.get("WORLD_SIZE", 1)) > 1:
        bench_ours(args)
    else:
        bench_pipeline(args)


if __name__ == "__main__":
    main()
