#!/usr/bin/env python
"""bench.py — read-bases/s through the k-mer→pileup path, samples/min from FASTQ(.gz) (BASELINE.json metric).

A step = one whole sample through the path: count k-mers of R1 and R2 (the KMC3 replacement), map the
counted k-mers to the 4-strain SARS-CoV-2 db, select the reference, score variants, read the result
back.  Workload at every N: BASELINE config C2 (SARS-CoV-2 single sample, 4-strain k=21 db, synthetic
150 bp PE reads at 10,000x with planted SNVs/iSNVs), one such sample per GPU per step (sample-per-GPU,
no collective: weak scaling).

  value    : bases/s with the reads already resident in HBM (device timed, CUDA events)
  e2e      : bases/s through the public API from pinned HOST buffers (H2D + result D2H inside the timing)
  roofline : the scan kernel (pack + seed/extend): algorithmic bytes = bases + 4 B/read offsets per launch;
             roofline_path: the whole path against SURVEY.md §8d's 1.03 B/base
  fastq    : samples/min from FASTQ.gz files on disk to a VCF on disk (decode + path + writer), GPU arm and CPU arm
  h2d      : the box's host→device ceiling for the same pinned buffers (what bounds e2e)
  sharded  : BASELINE config C3 — ONE ultra-deep sample (default 10^6x), read chunks sharded over the ranks, counts
             merged over NCCL inside the library; checked bit for bit against the same sample run unsharded on rank 0
             (a failed check fails the bench: exit code 1)
  cpu_baseline / --impl reference : the oracle (restated reference + KMC contract; the Rust reference
             and KMC3 cannot be built here) on the host cores, all threads, on the SAME config.
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


def workload_name(depth):
    return "C2: SARS-CoV-2 single sample vs 4-strain k=21 db, 150bp PE at %dx, one sample per GPU per step" % depth


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def committed_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the committed `ncu --set full` summary
    of this round (profiles/r02_ncu_full_summary.csv; same workload) — an offline capture, labelled as such; None if
    the file or the kernel is missing."""
    import csv
    for name in ("r02_ncu_full_summary.csv", "r01_ncu_full_summary.csv"):
        p = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(p):
            continue
        try:
            rows = [r for r in csv.reader(open(p)) if r]
            hdr = rows[0]
            ir, iw = [i for i, h in enumerate(hdr) if h.startswith("dram__bytes_read.sum")][0], [i for i, h in enumerate(hdr) if h.startswith("dram__bytes_write.sum")][0]
            vals = [(float(r[ir]) + float(r[iw])) * 1e6 for r in rows[1:] if r[0].replace("void ", "").startswith(kernel)]
            if vals:
                return sum(vals) / len(vals), "profiles/" + name
        except Exception:
            pass
    return None, None


REF_BUDGET_S = 100.0        # --impl reference: warm-up + K steps of the CPU port stay inside this (plus ~20 s for its FASTQ leg)


class ClockSampler:
    """SM clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).  ONE sampler per box (rank 0, every
    GPU of the job).  The numbers come from NVML inside this process (what nvidia-smi prints, without forking a poller
    that asks every GPU for its power draw every 50 ms — at eight GPUs that poller held the driver long enough to show in
    the ranks' launch rates); nvidia-smi is the fallback when the NVML module is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_indices, period_s=0.1):
        self.gpus, self.proc, self.lines, self.period = list(gpu_indices), None, [], period_s
        self.sm, self.mx, self.reasons, self.nv, self.thread, self.stop_flag = [], [], set(), None, None, False
        self.source = None

    def _physical(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            try:
                return [int(ids[g]) for g in self.gpus]
            except (ValueError, IndexError):
                return None                      # UUIDs: let nvidia-smi sort it out
        return self.gpus

    def start(self):
        try:
            if os.environ.get("BK_BENCH_SAMPLER") == "smi":      # developer runs: the forked poller, for comparison
                raise RuntimeError("forced")
            import pynvml
            phys = self._physical()
            if phys is None:
                raise RuntimeError("device ids")
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handles = [pynvml.nvmlDeviceGetHandleByIndex(g) for g in phys]
            self.get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self.mx = [float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)) for h in self.handles]
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ",".join(str(g) for g in self.gpus), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, n = self.nv, 0
        while not self.stop_flag:
            for h in self.handles:
                try:
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    if n % 2 == 0:                           # (the reasons are sticky enough for every other poll)
                        r = int(self.get_reasons(h))
                        for bit, name in self.REASONS:
                            if r & bit:
                                self.reasons.add(name)
                except Exception:
                    pass
            n += 1
            time.sleep(self.period)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nv is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "gpus": list(self.gpus), "source": "nvml"}
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
                "reasons": sorted(reasons), "samples": len(sm), "gpus": list(self.gpus), "source": "nvidia-smi"}


def make_workload(sample_index, depth):
    from bronko_b200 import sim
    r1, o1, r2, o2, _ = sim.simulate_pairs(sim.load_genome(sim.SARS4[0]), depth, sim.SEED0 + sample_index)
    return [(r1, o1), (r2, o2)]


def oracle_step(oi, files, threads):
    import bronko_b200
    from util import oracle_sample
    counts, s = oracle_sample(oi, files, bronko_b200.CallArgs(), threads=threads)
    return s


# ---- FASTQ files of the sample (fixed-width headers: the text is built with numpy, no per-read Python) ---------------
def fastq_text(bases, off, mate):
    n = len(off) - 1
    L = int(off[1] - off[0]) if n else 0
    assert n == 0 or bool((np.diff(off.astype(np.int64)) == L).all())
    idx = np.arange(n)
    hdr = np.empty((n, 13), dtype=np.uint8)
    hdr[:, 0:4] = np.frombuffer(b"@s0_", dtype=np.uint8)
    for d in range(7):
        hdr[:, 4 + d] = ord("0") + (idx // 10 ** (6 - d)) % 10
    hdr[:, 11] = ord("/"); hdr[:, 12] = ord("0") + mate
    rec = np.empty((n, 13 + 1 + L + 3 + L + 1), dtype=np.uint8)
    rec[:, :13] = hdr
    rec[:, 13] = 10
    rec[:, 14:14 + L] = bases[:n * L].reshape(n, L)
    rec[:, 14 + L:17 + L] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    rec[:, 17 + L:17 + 2 * L] = ord("I")
    rec[:, 17 + 2 * L] = 10
    return rec.reshape(-1)


def write_fastq_gz(path, text, bgzf, threads=8):
    """plain: one gzip member (what `gzip` writes).  bgzf: BGZF blocks (what bgzip / htslib write; every block is a
    complete gzip member <= 64 KiB, so the file is also a valid multi-member .gz for any reader)."""
    import zlib
    from concurrent.futures import ThreadPoolExecutor
    raw = memoryview(text)
    if not bgzf:
        c = zlib.compressobj(1, zlib.DEFLATED, 31)
        with open(path, "wb") as f:
            for i in range(0, len(raw), 64 << 20):
                f.write(c.compress(raw[i:i + (64 << 20)]))
            f.write(c.flush())
        return
    B = 65280

    def block(i):
        chunk = raw[i:i + B]
        co = zlib.compressobj(1, zlib.DEFLATED, -15)
        body = co.compress(chunk) + co.flush()
        bsize = len(body) + 25
        head = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + (bsize).to_bytes(2, "little")
        return head + body + (zlib.crc32(chunk) & 0xFFFFFFFF).to_bytes(4, "little") + len(chunk).to_bytes(4, "little")
    with ThreadPoolExecutor(threads) as ex, open(path, "wb") as f:
        for blk in ex.map(block, range(0, len(raw), B)):
            f.write(blk)
        f.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))     # BGZF EOF marker


def prepare_fastq(files, tmpdir, tag, bgzf):
    from concurrent.futures import ThreadPoolExecutor
    paths = [os.path.join(tmpdir, "%s_R%d.fastq.gz" % (tag, m + 1)) for m in range(len(files))]
    with ThreadPoolExecutor(2) as ex:
        list(ex.map(lambda m: write_fastq_gz(paths[m], fastq_text(files[m][0], files[m][1], m + 1), bgzf), range(len(files))))
    return paths


def cpu_fastq_sample(oi, paths, threads, out_vcf):
    """The reference's CPU flow for one sample: decode both files, count, map, select, score, write the VCF."""
    import bronko_b200
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    from util import oracle_params
    args = bronko_b200.CallArgs()
    with ThreadPoolExecutor(2) as ex:
        reads = list(ex.map(O.read_fastq, paths))                  # (rayon::join of the two KMC children, src/call.rs:302-307)
    counts = [O.Counts.count(args.kmer, b, off, args.min_kmers, 1000000, max(1, threads // 2)) for b, off in reads]
    s = O.Sample(oi, oracle_params(args), counts)
    with open(out_vcf, "w") as f:
        f.write(s.vcf_text(paths[0]))
    return s


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the SAME config.  The Rust binary + KMC3
    cannot be built in this image (no cargo/rustc/kmc, no network), so this times the oracle port (oracle/), all host
    threads; a step is the whole C2 sample, exactly what the GPU arm's step is."""
    if rank != 0:
        return
    import tempfile
    from oracle import oracle as O
    from bronko_b200 import sim
    cores = os.cpu_count() or 1
    files = make_workload(0, args.depth)
    full_bases = sum(len(b) for b, _ in files)
    oi = O.Index.build(21, [sim.genome_path(n) for n in sim.SARS4])
    # A step is a BOUNDED sample of the workload: the whole C2 sample takes the port ~3 s, so with the driver's K the
    # steps take the first n read pairs of it, n sized so that warm-up + K steps stay inside REF_BUDGET_S (the whole
    # sample when it fits).  Bases per second barely depend on n: the port's time is counting, linear in the reads.
    t0 = time.perf_counter()
    oracle_step(oi, files, cores)                                     # untimed: page-in + the rate of the full sample
    t_full = time.perf_counter() - t0
    frac = min(1.0, max(0.02, REF_BUDGET_S / (t_full * max(1, args.steps + min(args.warmup, 1)))))
    step_files = files
    if frac < 1.0:
        step_files = []
        for b, o in files:
            n = max(1, int((len(o) - 1) * frac))
            step_files.append((b[:int(o[n])], o[:n + 1]))
    n_bases = sum(len(b) for b, _ in step_files)
    for _ in range(max(0, min(args.warmup, 1))):
        oracle_step(oi, step_files, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(oi, step_files, cores)
    dt = time.perf_counter() - t0
    v = n_bases * args.steps / dt
    fq = None
    if not args.no_fastq:
        with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
            paths = prepare_fastq(files, td, "ref", bgzf=False)
            t1 = time.perf_counter()
            reps = 2
            for _ in range(reps):
                cpu_fastq_sample(oi, paths, cores, os.path.join(td, "ref.vcf"))
            fq = {"samples_per_min": 60.0 * reps / (time.perf_counter() - t1), "input": "plain single-member .fastq.gz, R1 + R2",
                  "threads": cores, "samples": reps}
    sample = "SARS-CoV-2 %dx 150bp PE: the first %.0f %% of the read pairs of the C2 sample per step (%d of %d bases), oracle port, %d threads" % (
        args.depth, 100.0 * n_bases / full_bases, n_bases, full_bases, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args.depth)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fastq": fq,
    }), file=OUT, flush=True)


# ---- BASELINE config C3: one ultra-deep sample, read chunks sharded over the ranks -----------------------------------
def sharded_leg(args, rank, world, local_rank, owner_ctx, dist, torch):
    import bronko_b200
    from bronko_b200 import sim
    from bronko_b200.dist import init_sharded, leave_sharded
    dev = torch.device("cuda", local_rank)
    genome = sim.load_genome(sim.SARS4[0])
    L = len(genome)
    n_chunks = args.shard_chunks
    while n_chunks % world:
        n_chunks += 1
    chunk_pairs = int(round(args.shard_depth / n_chunks * L / 300))
    total_depth = chunk_pairs * n_chunks * 300 / L
    plan = sim.plant_for(genome, sim.SEED0 + 777)
    seed_of = lambda c: (sim.SEED0 + 777) * 1000 + c          # noqa: E731
    mine = list(range(rank * n_chunks // world, (rank + 1) * n_chunks // world))
    chunks = {c: sim.simulate_pairs_torch(genome, chunk_pairs, seed_of(c), dev, plan) for c in mine}
    torch.cuda.synchronize()
    n_bases_chunk = 2 * chunk_pairs * 150
    total_bases = n_bases_chunk * n_chunks
    cargs = bronko_b200.CallArgs()

    def push_chunk(c, ch):
        r1, r2, off = ch
        c.push_device(0, r1.data_ptr(), off.data_ptr(), chunk_pairs, chunk_pairs * 150, 150)
        c.push_device(1, r2.data_ptr(), off.data_ptr(), chunk_pairs, chunk_pairs * 150, 150)

    ctx = bronko_b200.Bronko(local_rank)
    ctx.share_index(owner_ctx)
    if world > 1:
        init_sharded(ctx)

    def step():
        ctx.begin(cargs)
        for c in mine:
            push_chunk(ctx, chunks[c])
        return ctx.finish()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    for _ in range(max(1, args.shard_warmup)):
        res = step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    acc = {}
    for _ in range(args.shard_steps):
        res = step()
        for k_, v_ in ctx.stage_times().items():
            acc[k_] = acc.get(k_, 0.0) + v_ / args.shard_steps
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.shard_steps
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    out = {"config": "C3: ONE SARS-CoV-2 sample at %.0fx (%d read pairs, %.3g bases), %d chunks of %d pairs, read chunks sharded over %d rank(s)"
                     % (total_depth, chunk_pairs * n_chunks, total_bases, n_chunks, chunk_pairs, world),
           "n_ranks": world, "depth": total_depth, "bases": total_bases, "ms_per_sample": ms, "value": total_bases / (ms * 1e-3), "unit": UNIT,
           "steps": args.shard_steps, "warmup": max(1, args.shard_warmup), "inputs": "resident in HBM, generated on the device (torch Philox, seed per chunk)",
           "collectives": {"calls_per_sample": int(round(acc.get("coll_calls", 0))), "ms_per_sample_rank0": acc.get("coll_ms", 0.0),
                           "transport": "NCCL issued by the library on the context's stream (libnccl via dlopen)" if world > 1 else "none (one rank)"},
           "stage_ms_rank0": {k_: acc[k_] for k_ in ("scan_ms", "leftover_ms", "finalize_ms", "map_ms", "score_ms", "coll_ms", "total_ms") if k_ in acc},
           "result": {"best_genome": int(res.best_genome), "n_variants": int(len(res.variants)), "kmc_R1": [int(x) for x in res.kmc_stats(0)]}}
    # the check: the same sample, unsharded, on rank 0 (chunks of other ranks are regenerated from their seeds)
    ok = True
    if world > 1 and not args.no_shard_check:
        if rank == 0:
            whole = bronko_b200.Bronko(local_rank)
            whole.share_index(owner_ctx)
            t0 = time.perf_counter()
            # every chunk of the sample resident on this GPU (the other ranks' chunks regenerated from their seeds)
            allc = {c: (chunks[c] if c in chunks else sim.simulate_pairs_torch(genome, chunk_pairs, seed_of(c), dev, plan)) for c in range(n_chunks)}
            torch.cuda.synchronize()               # the chunks are complete before the library's streams read them
            for _ in range(2):                     # (the first run allocates)
                whole.begin(cargs)
                for c in range(n_chunks):
                    push_chunk(whole, allc[c])
                ref = whole.finish()
            unsharded_ms = whole.stage_times()["total_ms"]
            del allc
            checks = {
                "variants": ref.variants.tobytes() == res.variants.tobytes(),
                "pileup": bool((ref.pileup() == res.pileup()).all()),
                "kmc": [ref.kmc_stats(f) for f in range(2)] == [res.kmc_stats(f) for f in range(2)],
                "tallies": all(ref.mapping_data(f).tobytes() == res.mapping_data(f).tobytes() for f in range(2)),
                "best_genome": ref.best_genome == res.best_genome,
                "noise_max": bool(np.array_equal(ref.noise_max(), res.noise_max())),
                "summary": (ref.num_major_variants, ref.num_minor_variants, ref.breadth_coverage, ref.depth_coverage, ref.num_unmapped_kmers) ==
                           (res.num_major_variants, res.num_minor_variants, res.breadth_coverage, res.depth_coverage, res.num_unmapped_kmers),
            }
            ok = all(checks.values())
            out["bit_equal"] = ok
            out["bit_equal_checks"] = checks
            out["unsharded_same_sample_ms_rank0"] = unsharded_ms
            out["efficiency_vs_unsharded_device_time"] = unsharded_ms / (ms * world)
            out["check_wall_s"] = time.perf_counter() - t0
            whole.close()
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.broadcast(flag, src=0)
        ok = bool(flag.item())
    if world > 1:
        leave_sharded(ctx)
    ctx.close()
    del chunks
    torch.cuda.empty_cache()
    return out, ok


def _sync_ctx(c):
    """Wait for the counting kernels a context has enqueued so far (both file slots)."""
    import torch
    for slot in (0, 1):
        h = c._lib.bk_stream_slot(c.h, slot)
        if h:
            torch.cuda.ExternalStream(int(h)).synchronize()


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
    ap.add_argument("--depth", type=int, default=10000, help="coverage of the C2 sample (BASELINE: 10,000x); both arms and cpu_baseline use it")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="developer runs: skip the host-buffer leg (e2e is null)")
    ap.add_argument("--no-fastq", action="store_true", help="skip the FASTQ.gz → VCF legs (samples/min)")
    ap.add_argument("--no-sharded", action="store_true", help="skip the C3 leg (one ultra-deep sample sharded over the ranks)")
    ap.add_argument("--no-shard-check", action="store_true", help="skip the unsharded re-run of the C3 sample on rank 0")
    ap.add_argument("--shard-depth", type=float, default=1e6, help="total coverage of the C3 sample (BASELINE: 1M x)")
    ap.add_argument("--shard-chunks", type=int, default=40, help="read chunks of the C3 sample (rounded up to a multiple of the ranks)")
    ap.add_argument("--shard-steps", type=int, default=3)
    ap.add_argument("--shard-warmup", type=int, default=1)
    ap.add_argument("--in-flight", type=int, default=6,
                    help="samples in flight per GPU (one bk_ctx + streams each); 1 = strictly one sample at a time")
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
        import datetime
        # (a rank that dies must not leave the others waiting for ever in a collective)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=240))
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
    # the same reads as 2-bit words (bk_reads_pack — what a host decode stage hands over when the bytes have to cross PCIe)
    packed_in = []
    if not args.no_e2e:
        for b, o in files:
            pk, poff, rest, roff = bronko_b200.pack_reads(b, o)
            assert len(roff) == 1, "synthetic reads hold ACGT only"
            packed_in.append((torch.from_numpy(pk.view(np.int32)).pin_memory(), torch.from_numpy(poff.view(np.int32).copy()).pin_memory(), len(poff) - 1))
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

    def step_e2e_packed(c):
        c.begin(cargs)
        for slot, (hp, ho, n) in enumerate(packed_in):
            c.push_packed_ptr(slot, hp.data_ptr(), ho.data_ptr(), n)
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
    # every rank watches its own GPU through NVML (two light queries per 50 ms inside the process); rank 0 merges
    clocks = ClockSampler([local_rank])
    clocks.start()
    alone = {}
    n_alone = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(n_alone):
        step_device(ctx)
        for k_, v_ in ctx.stage_times().items():
            alone[k_] = alone.get(k_, 0.0) + v_ / n_alone
    latency_ms = (time.perf_counter() - t0) * 1e3 / n_alone
    # the timed region runs without the per-stage events (bk_stage_timing: instrumentation, ~20 driver calls per sample on
    # the one host thread a context has); the stage times of samples in flight come from the un-timed leg behind it
    for c in ctxs:
        c.set_stage_timing(False)
    run_steps(S, step_device)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import resource
    ru0 = resource.getrusage(resource.RUSAGE_SELF)
    e0.record()
    run_steps(args.steps, step_device)
    torch.cuda.synchronize()
    e1.record()
    ru1 = resource.getrusage(resource.RUSAGE_SELF)
    host_cpu_ms = {"user": (ru1.ru_utime - ru0.ru_utime) * 1e3 / args.steps, "sys": (ru1.ru_stime - ru0.ru_stime) * 1e3 / args.steps,
                   "note": "CPU time of this process (all threads) per step of the timed region"}
    barrier()
    res = last["res"]
    t_end = time.perf_counter() + 0.5
    while time.perf_counter() < t_end:                     # same work, same samples in flight, not timed
        run_steps(max(S, args.steps // 4), step_device)
    torch.cuda.synchronize()
    for c in ctxs:
        c.set_stage_timing(True)
    n_collect = max(S, args.steps // 4)
    run_steps(n_collect, step_device, collect=True)
    torch.cuda.synchronize()
    clk = clocks.stop()
    if dist is not None:                                   # lowest per-GPU median, highest maximum, any reason seen anywhere
        names = [n for _, n in ClockSampler.REASONS]
        t = torch.tensor([-(clk.get("sm_mhz") or 0.0), clk.get("sm_max_mhz") or 0.0] + [1.0 if n in clk.get("reasons", []) else 0.0 for n in names],
                         dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cnt = torch.tensor([float(clk.get("samples") or 0)], dtype=torch.float64, device="cuda")
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        v = t.tolist()
        clk = {"sm_mhz": -v[0] or None, "sm_max_mhz": v[1] or None, "reasons": [n for n, f in zip(names, v[2:]) if f > 0] + [r for r in clk.get("reasons", []) if r not in names],
               "samples": int(cnt.item()), "gpus": list(range(world)), "source": clk.get("source"),
               "merge": "sm_mhz = lowest per-GPU median under load, sm_max_mhz = highest maximum, reasons = seen on any GPU"}
    clk["window"] = "single-sample latency leg + device-timed region + 0.5 s of the same load behind it; every rank samples its own GPU"
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    total_bases = n_bases * world
    value = total_bases / (ms_step * 1e-3)

    # ---- e2e: pinned host buffers through the public API, wall clock incl. H2D + result D2H ------
    h2d = sum(hb.numel() - 64 + ho.numel() * 4 for hb, ho, _ in pinned)
    h2d_probe = None
    if not args.no_e2e:
        run_steps(2 * S, step_e2e)
        barrier()
        t0 = time.perf_counter()
        run_steps(args.steps, step_e2e)
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e_ascii_value = total_bases * args.steps / e2e_s
        run_steps(2 * S, step_e2e_packed)
        barrier()
        t0 = time.perf_counter()
        run_steps(args.steps, step_e2e_packed)
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e_value = total_bases * args.steps / e2e_s
        # the ceiling e2e runs into: the same pinned buffers copied to the device and nothing else, all ranks at once
        dst = [torch.empty_like(tb) for tb, _, _, _ in dev]
        cs = torch.cuda.Stream()
        reps = 10
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(cs):
            for _ in range(reps):
                for d_, (hb, _, _) in zip(dst, pinned):
                    d_.copy_(hb, non_blocking=True)
        cs.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        nbytes = sum(hb.numel() for hb, _, _ in pinned)
        h2d_probe = {"gbs_per_gpu": nbytes * reps / dt / 1e9, "gbs_total": nbytes * reps * world / dt / 1e9, "ranks_copying": world,
                     "bytes_per_copy": nbytes, "bases_per_s_if_only_h2d": total_bases * reps / dt,
                     "note": "cudaMemcpyAsync of the sample's pinned ASCII buffers, all ranks concurrently, max over ranks: the upper bound of e2e on this box"}
        del dst
    else:
        e2e_value = e2e_ascii_value = None
    h2d_packed = sum(hp.numel() * 4 + ho.numel() * 4 for hp, ho, _ in packed_in)
    d2h = int(len(res.variants) * 72 + 120 + 2 * 4 * 16)

    # ---- roofline of the streaming kernel (scan) --------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    scan_launches = max(1, int(round(alone["scan_launches"])))
    scan_ms = alone["scan_ms"] / scan_launches
    alg_bytes = (n_bases + 4 * (n_reads + 2)) / 2.0            # per launch: one file's bases + u32 offsets
    achieved = alg_bytes / (scan_ms * 1e-3) / 1e9
    traffic, traffic_src = committed_traffic("k_scan") if args.depth == 10000 else (None, None)
    stage_keys = ("scan_ms", "leftover_ms", "finalize_ms", "map_ms", "score_ms")
    kernel_sum = sum(alone[k_] for k_ in stage_keys)
    roofline = {"bound": "hbm", "kernel": "k_scan", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "alg_bytes_per_launch": alg_bytes, "avg_launch_ms": scan_ms,
                "share_of_kernel_time": alone["scan_ms"] / max(kernel_sum, 1e-9),
                "stage_shares": {k_: alone[k_] / max(kernel_sum, 1e-9) for k_ in stage_keys},
                "timed": "CUDA events around each k_scan launch on its stream, one sample in flight, %d samples" % n_alone,
                "note": "k_scan streams the reads (the HBM-bound stage of SURVEY.md 8d: 1 B/base); the other stages are "
                        "latency-bound random access (leftover / bins / map) or a sequential FP64 chain (noise) whose "
                        "algorithmic bytes are ~3 % of the path's: the whole path against its bytes is roofline_path"}
    stages = {k_: (v_ / n_collect) for k_, v_ in stage_acc.items()}

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on the SAME sample --------------
    cpu = None
    oi = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        oi = O.Index.build(21, [sim.genome_path(n) for n in sim.SARS4])
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            oracle_step(oi, files, cores)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        cpu = {"value": n_bases / best, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "the same C2 sample: SARS-CoV-2 %dx 150bp PE (%d bases), oracle port of bronko+KMC contract, %d threads, best of 2" % (args.depth, n_bases, cores)}

    # ---- samples/min from FASTQ.gz on disk to a VCF on disk (BASELINE metric, second half) ----------
    fastq = None
    if not args.no_fastq and rank == 0:
        fastq = fastq_legs(args, ctxs, files, oi if world == 1 else None)
    if dist is not None:
        dist.barrier()

    # ---- C3: the read-sharded ultra-deep sample ----------------------------------------------------
    sharded, shard_ok = None, True
    if not args.no_sharded:
        for c in ctxs[1:]:
            c.close()                          # (their scratch is not needed any more)
        ctxs = ctxs[:1]
        sharded, shard_ok = sharded_leg(args, rank, world, local_rank, ctx, dist, torch)

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(args.depth),
                       "bases_per_step_per_gpu": n_bases, "reads_per_step_per_gpu": n_reads, "k": 21,
                       "l2": "inputs (%.0f MB/step) exceed the 126 MB L2; no explicit flush" % (n_bases / 1e6),
                       "parallelism": "sample-per-GPU x%d, no collective" % world,
                       "samples_in_flight_per_gpu": S},
            "latency_ms_single_sample": latency_ms,
            "e2e": None if e2e_value is None else {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_packed, "d2h_bytes_per_step": d2h,
                                                   "input": "2-bit packed reads + u32 read offsets in pinned host memory (bk_reads_push_packed), unpacked on the device"},
            "e2e_ascii": None if e2e_value is None else {"value": e2e_ascii_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                                         "input": "ASCII bases + u32 read offsets in pinned host memory (bk_reads_push): bound by the H2D rate, see h2d_ceiling"},
            "h2d_ceiling": h2d_probe,
            "gpu_launches": int(round(stage_acc["launches"] / n_collect * args.steps)) * world,   # kernels of the timed region, all ranks (the same count every sample and rank)
            "gpu_launches_per_rank": int(round(stage_acc["launches"] / n_collect * args.steps)),
            "host_cpu_ms_per_step": host_cpu_ms,
            "roofline": roofline,
            "roofline_path": {"alg_bytes_per_step": 1.03 * n_bases, "achieved_gbs": 1.03 * n_bases * world / (ms_step * 1e-3) / 1e9,
                              "frac_of_hbm": 1.03 * n_bases / (ms_step * 1e-3) / 1e9 / peak,
                              "note": "whole path, SURVEY.md 8d: 1.03 algorithmic bytes per read base"},
            "cpu_baseline": cpu, "clocks": clk,
            "fastq": fastq, "sharded": sharded,
            "stage_ms_per_step_in_flight": stages, "stage_ms_single_sample": alone,
            "result_check": {"best_genome": int(res.best_genome), "n_variants": int(len(res.variants))},
        }
        print(json.dumps(out), file=OUT, flush=True)
    for c in ctxs:
        c.close()
    if dist is not None:
        dist.destroy_process_group()
    if not shard_ok:
        sys.stderr.write("bench.py: the sharded sample is NOT bit-equal to the unsharded run\n")
        sys.exit(1)


def fastq_legs(args, ctxs, files, oi):
    """Rank 0: the C2 sample as R1/R2 .fastq.gz files on disk → VCF on disk, samples/min.  GPU arm: several samples in
    flight (decode of the next samples overlaps the GPU), plain single-member gzip (host inflate) and BGZF; CPU arm
    (N = 1 only): the oracle flow from the same plain files."""
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    import bronko_b200
    out = {}
    cargs = bronko_b200.CallArgs()
    cores = os.cpu_count() or 1
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        for kind, bgzf in (("plain_gz", False), ("bgzf", True)):
            paths = prepare_fastq(files, td, kind, bgzf)
            size = sum(os.path.getsize(p) for p in paths)

            def one(ci, i, paths=paths, kind=kind):
                c = ctxs[ci]
                c.begin(cargs)
                for slot, p in enumerate(paths):
                    c.push_fastq(slot, p)
                r = c.finish()
                r.write_vcf(paths[0], os.path.join(td, "%s_%d.vcf" % (kind, ci)))
                return len(r.variants)
            nv = [one(0, 0)]                                           # warm-up (page cache, allocations)
            per_ctx = 8 if bgzf else 2                                 # (a BGZF sample takes ~10 ms, a plain-gzip one ~1.4 s of host inflate per context)
            n = per_ctx * len(ctxs)

            def worker(ci):                                            # one host thread per context (a context is not thread-safe)
                for i in range(per_ctx):
                    one(ci, i)
            th = [threading.Thread(target=worker, args=(ci,)) for ci in range(len(ctxs))]
            t0 = time.perf_counter()
            for t in th:
                t.start()
            for t in th:
                t.join()
            dt = time.perf_counter() - t0
            out[kind] = {"samples_per_min": 60.0 * n / dt, "samples": n, "in_flight": len(ctxs), "compressed_bytes_per_sample": size,
                         "n_variants": nv[0], "decode": ctxs[0].decode_info()}
        if oi is not None and not args.no_cpu_baseline:
            paths = prepare_fastq(files, td, "cpu", False)
            t0 = time.perf_counter()
            cpu_fastq_sample(oi, paths, cores, os.path.join(td, "cpu.vcf"))
            out["cpu_port_plain_gz"] = {"samples_per_min": 60.0 / (time.perf_counter() - t0), "threads": cores, "samples": 1}
    out["note"] = "one C2 sample (R1 + R2 .fastq.gz on /dev/shm) → decode → k-mer→pileup path → VCF file; GPU arm keeps several samples in flight on one GPU"
    return out


if __name__ == "__main__":
    main()
