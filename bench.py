#!/usr/bin/env python
"""bench.py — read-bases/s through the k-mer→pileup path (BASELINE.json metric).

A step = one whole sample through the path: count k-mers of R1 and R2 (the KMC3 replacement), map the
counted k-mers to the 4-strain SARS-CoV-2 db, select the reference, score variants, read the result
back.  Workload at every N: BASELINE config C2 (SARS-CoV-2 single sample, 4-strain k=21 db, synthetic
150 bp PE reads at 10,000x with planted SNVs/iSNVs), one such sample per GPU per step (sample-per-GPU,
no collective: weak scaling).

  value    : bases/s with the reads already resident in HBM (device timed, CUDA events on the ctx stream)
  e2e      : bases/s through the public API from pinned HOST buffers (H2D + result D2H inside the timing)
  roofline : the scan kernel (pack + seed/extend): algorithmic bytes = bases + 4 B/read offsets per launch
  cpu_baseline / --impl reference : the oracle (restated reference + KMC contract; the Rust reference
             and KMC3 cannot be built here) on the host cores, all threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

OUT = sys.stdout
METRIC = "read_bases_per_sec_kmer_to_pileup"
UNIT = "bases/s"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(sample_index, depth):
    from bronko_b200 import sim
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[0]), depth, sim.SEED0 + sample_index)
    return [(r1, o1), (r2, o2)]


def oracle_step(oi, files, threads):
    import bronko_b200
    from util import oracle_sample
    counts, s = oracle_sample(oi, files, bronko_b200.CallArgs(), threads=threads)
    return s


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The Rust binary + KMC3 cannot be
    built in this image (no cargo/rustc/kmc, no network), so this times the oracle port (oracle/), all host
    threads, on a bounded sample of the same workload per step."""
    if rank != 0:
        return
    from oracle import oracle as O
    from bronko_b200 import sim
    cores = os.cpu_count() or 1
    sample_depth = min(args.depth, args.ref_depth)
    files = make_workload(0, sample_depth)
    n_bases = sum(len(b) for b, _ in files)
    oi = O.Index.build(21, [sim.genome_path(n) for n in sim.SARS4])
    for _ in range(max(1, min(args.warmup, 1))):
        oracle_step(oi, files, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(oi, files, cores)
    dt = time.perf_counter() - t0
    v = n_bases * args.steps / dt
    sample = "SARS-CoV-2 %dx 150bp PE (%d bases/step), oracle port, %d threads" % (sample_depth, n_bases, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "C2: SARS-CoV-2 single sample vs 4-strain k=21 db, 150bp PE, bounded sample at %dx" % sample_depth},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), file=OUT, flush=True)


def main():
    # stdout carries exactly ONE line, the JSON: everything libraries print to fd 1 (NCCL's version banner under
    # torchrun, ...) is sent to stderr, and the line is written to the saved descriptor
    global OUT
    OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--depth", type=int, default=10000, help="coverage of the C2 sample (BASELINE: 10,000x)")
    ap.add_argument("--ref-depth", type=int, default=2000, help="bounded sample depth for the CPU arm")
    ap.add_argument("--cpu-depth", type=int, default=2000, help="bounded sample depth for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="developer runs: skip the host-buffer leg (e2e is null)")
    ap.add_argument("--in-flight", type=int, default=4,
                    help="samples in flight per GPU (one bk_ctx + stream each); 1 = strictly one sample at a time")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import bronko_b200
    from bronko_b200 import sim
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    # S contexts per GPU keep S samples in flight: a sample's sequential tail (the exact noise chains run on
    # three warps) overlaps the next samples' streaming kernels.  Every step is still one whole sample.
    S = max(1, args.in_flight)
    ctxs = []
    for i in range(S):
        c = bronko_b200.Bronko(local_rank)    # raises without the CUDA library / a B200: no fallback
        if i == 0:
            c.build_index(21, [sim.genome_path(n) for n in sim.SARS4])
        else:
            c.share_index(ctxs[0])            # one copy of the db per GPU
        ctxs.append(c)
    ctx = ctxs[0]
    files = make_workload(rank, args.depth)
    n_bases = sum(len(b) for b, _ in files)
    n_reads = sum(len(o) - 1 for _, o in files)

    # device-resident inputs (value) and pinned host inputs (e2e); read-only, shared by the contexts
    dev, pinned = [], []
    for b, o in files:
        pad = np.concatenate([b, np.full(64, ord("*"), dtype=np.uint8)])
        tb = torch.from_numpy(pad).cuda()
        to = torch.from_numpy(o.view(np.int32)).cuda()
        dev.append((tb, to, len(o) - 1, len(b)))
        hb = torch.from_numpy(pad).pin_memory()
        ho = torch.from_numpy(o.view(np.int32).copy()).pin_memory()
        pinned.append((hb, ho, len(o) - 1))
    cargs = bronko_b200.CallArgs()
    last = {}
    stage_acc = {}
    acc_lock = threading.Lock()

    def step_device(c):
        c.begin(cargs)
        for slot, (tb, to, n, nb) in enumerate(dev):
            c.push_device(slot, tb.data_ptr(), to.data_ptr(), n, nb, 150)
        return c.finish()

    def step_e2e(c):
        c.begin(cargs)
        for slot, (hb, ho, n) in enumerate(pinned):
            c.push_ptr(slot, hb.data_ptr(), ho.data_ptr(), n)
        r = c.finish()
        _ = r.variants
        return r

    def run_steps(n_steps, fn, collect=False):
        """n_steps samples over the S contexts (thread ci takes steps ci, ci+S, ...; ctypes drops the GIL)."""
        def work(ci):
            for _ in range(ci, n_steps, S):
                r = fn(ctxs[ci])
                last["res"] = r
                if collect:
                    t = ctxs[ci].stage_times()
                    with acc_lock:
                        for k_, v_ in t.items():
                            stage_acc[k_] = stage_acc.get(k_, 0.0) + v_
        th = [threading.Thread(target=work, args=(ci,)) for ci in range(S)]
        for t in th:
            t.start()
        for t in th:
            t.join()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: inputs resident in HBM, device-timed ------------------------------------------
    run_steps(args.warmup * S, step_device)
    # single-sample latency and per-kernel times with nothing else in flight (the roofline of the scan kernel is
    # quoted on the kernel timed alone; under S-in-flight other samples' kernels share the SMs)
    torch.cuda.synchronize()
    # clocks / throttle reasons: sampled from here, through the device-timed region, to the end of an un-timed leg of the
    # same load behind it (the timed region alone, ~0.1 s, is shorter than nvidia-smi's start-up)
    clocks = ClockSampler(local_rank)
    clocks.start()
    alone = {}
    n_alone = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(n_alone):
        step_device(ctx)
        for k_, v_ in ctx.stage_times().items():
            alone[k_] = alone.get(k_, 0.0) + v_ / n_alone
    latency_ms = (time.perf_counter() - t0) * 1e3 / n_alone
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_steps(args.steps, step_device, collect=True)
    torch.cuda.synchronize()
    e1.record()
    barrier()
    res = last["res"]
    t_end = time.perf_counter() + 0.5
    while time.perf_counter() < t_end:                     # same work, same samples in flight, not timed
        run_steps(max(S, args.steps // 4), step_device)
    torch.cuda.synchronize()
    clk = clocks.stop()
    clk["window"] = "single-sample latency leg + device-timed region + 0.5 s of the same load behind it"
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    total_bases = n_bases * world
    value = total_bases / (ms_step * 1e-3)

    # ---- e2e: pinned host buffers through the public API, wall clock incl. H2D + result D2H ------
    if not args.no_e2e:
        run_steps(2 * S, step_e2e)
        barrier()
        t0 = time.perf_counter()
        run_steps(args.steps, step_e2e)
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e_value = total_bases * args.steps / e2e_s
    else:
        e2e_value = None
    h2d = sum(hb.numel() - 64 + ho.numel() * 4 for hb, ho, _ in pinned)
    d2h = int(len(res.variants) * 72 + 120 + 2 * 4 * 16)

    # ---- roofline of the dominant kernel (scan) -------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    scan_launches = max(1, int(round(alone["scan_launches"])))
    scan_ms = alone["scan_ms"] / scan_launches
    alg_bytes = (n_bases + 4 * (n_reads + 2)) / 2.0            # per launch: one file's bases + u32 offsets
    achieved = alg_bytes / (scan_ms * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of one k_scan launch at this workload, from the committed
    # `ncu --set full` capture (profiles/r01_ncu_full_summary.csv: 158.05 MB + 4.95 / 5.42 MB, two launches); other depths: not captured
    traffic = 163.2e6 if args.depth == 10000 else None
    alone_total = max(alone.get("total_ms", 0.0), 1e-9)
    roofline = {"bound": "hbm", "kernel": "k_scan", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": alg_bytes, "avg_launch_ms": scan_ms,
                "share_of_step": alone["scan_ms"] / alone_total,
                "timed": "CUDA events around each k_scan launch, one sample in flight (kernel timed alone), %d samples" % n_alone,
                "note": "k_scan is the kernel that streams the reads (the HBM-bound stage of SURVEY.md 8d); the other "
                        "stages are latency-bound random access (leftover / map) or a sequential FP64 chain (noise): "
                        "see stage_ms_single_sample and profiles/r01_kernel_share.txt"}
    stages = {k_: (v_ / args.steps) for k_, v_ in stage_acc.items()}

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on a bounded sample ----------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        cd = min(args.depth, args.cpu_depth)
        cfiles = make_workload(0, cd)
        cb = sum(len(b) for b, _ in cfiles)
        oi = O.Index.build(21, [sim.genome_path(n) for n in sim.SARS4])
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            oracle_step(oi, cfiles, cores)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        cpu = {"value": cb / best, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "SARS-CoV-2 %dx 150bp PE (%d bases), oracle port of bronko+KMC contract, %d threads, best of 2" % (cd, cb, cores)}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": "C2: SARS-CoV-2 single sample vs 4-strain k=21 db, 150bp PE at %dx, one sample per GPU per step" % args.depth,
                       "bases_per_step_per_gpu": n_bases, "reads_per_step_per_gpu": n_reads, "k": 21,
                       "l2": "inputs (%.0f MB/step) exceed the 126 MB L2; no explicit flush" % (n_bases / 1e6),
                       "parallelism": "sample-per-GPU x%d, no collective" % world,
                       "samples_in_flight_per_gpu": S},
            "latency_ms_single_sample": latency_ms,
            "e2e": None if e2e_value is None else {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(stage_acc["launches"]),
            "roofline": roofline,
            "roofline_path": {"alg_bytes_per_step": 1.03 * n_bases, "achieved_gbs": 1.03 * n_bases * world / (ms_step * 1e-3) / 1e9,
                              "frac_of_hbm": 1.03 * n_bases / (ms_step * 1e-3) / 1e9 / peak,
                              "note": "whole path, SURVEY.md 8d: 1.03 algorithmic bytes per read base"},
            "cpu_baseline": cpu, "clocks": clk,
            "stage_ms_per_step_in_flight": stages, "stage_ms_single_sample": alone,
            "result_check": {"best_genome": int(res.best_genome), "n_variants": int(len(res.variants))},
        }
        print(json.dumps(out), file=OUT, flush=True)
    for c in ctxs:
        c.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
