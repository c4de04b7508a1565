#!/usr/bin/env python
"""tools/launch_share.py LAUNCHES.csv — per kernel: median gpu__time_duration per launch, launches per sample, share of one
sample's kernel time, from an `ncu --metrics gpu__time_duration.sum --csv` launch list of a bench run."""
import collections
import csv
import sys


def main():
    lr = list(csv.reader(open(sys.argv[1])))
    h0 = [i for i, r in enumerate(lr) if r and r[0] == "ID"][0]
    H = lr[h0]
    ik, iv = H.index("Kernel Name"), H.index("Metric Value")
    per = collections.OrderedDict()
    for r in lr[h0 + 1:]:
        if len(r) > iv:
            per.setdefault(r[ik].split("(")[0], []).append(float(r[iv].replace(",", "")) / 1e3)
    n_samples = max(1, len(per.get("k_select", [1])))
    tot = sum(sorted(v)[len(v) // 2] * len(v) / n_samples for v in per.values())
    print("median gpu__time_duration per launch (us), launches per sample, share of one sample's kernel time")
    for k, v in per.items():
        med = sorted(v)[len(v) // 2]
        print("%-34s %9.1f us  x%-4.1f %5.1f %%" % (k, med, len(v) / n_samples, 100 * med * len(v) / n_samples / tot))
    print("sum per sample: %.1f us over %d samples (cold-cache, serialised by the profiler)" % (tot, n_samples))


if __name__ == "__main__":
    main()
