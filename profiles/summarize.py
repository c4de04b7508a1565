#!/usr/bin/env python
"""Turn the ncu reports brought back from the GPU box (gpurun_out/*.ncu-rep, not tracked) into the small text
summaries committed here.  Usage: python profiles/summarize.py gpurun_out/full_r1s.ncu-rep gpurun_out/launches_r1s.csv r01"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    hdr, units, rows = raw(rep)
    with open("profiles/%s_ncu_full_summary.csv" % tag, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + ["%s [%s]" % (k, units[hdr.index(k)]) for k in KEYS if k in hdr])
        for r in rows:
            d = dict(zip(hdr, r))
            w.writerow([d["Kernel Name"].split("(")[0]] + [d[k] for k in KEYS if k in hdr])
    # share of the step per kernel from the launch list (gpu__time_duration per launch, last sample of the run)
    lr = list(csv.reader(open(launches)))
    h0 = [i for i, r in enumerate(lr) if r and r[0] == "ID"][0]
    H = lr[h0]
    ik, iv = H.index("Kernel Name"), H.index("Metric Value")
    per = collections.OrderedDict()
    for r in lr[h0 + 1:]:
        if len(r) > iv:
            per.setdefault(r[ik].split("(")[0], []).append(float(r[iv].replace(",", "")) / 1e3)
    with open("profiles/%s_kernel_share.txt" % tag, "w") as f:
        f.write("median gpu__time_duration per launch (us), launches in the list, share of one sample's kernel time\n")
        n_samples = max(1, len(per.get("k_select", [1])))
        tot = sum(sorted(v)[len(v) // 2] * len(v) / n_samples for v in per.values())
        for k, v in per.items():
            med = sorted(v)[len(v) // 2]
            f.write("%-28s %9.1f us  x%-3d  %5.1f %%\n" % (k, med, len(v) // n_samples, 100 * med * len(v) / n_samples / tot))
        f.write("sum per sample: %.1f us (cold-cache, serialised by the profiler)\n" % tot)


if __name__ == "__main__":
    main()
