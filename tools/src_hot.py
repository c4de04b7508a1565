#!/usr/bin/env python
"""tools/src_hot.py REPORT.ncu-rep KERNEL_REGEX [N] — the N source lines of a kernel with the most warp-stall samples
(ncu --page source --print-source cuda,sass of a --set full --import-source on capture)."""
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H = [r for r in rows if r and r[0] == "Line No"][0]
    i_s, i_i = H.index("# Samples"), H.index("Instructions Executed")
    file, agg, seen = None, [], set()
    for r in rows:
        if r and r[0] == "File Path":
            file = r[1].split("/")[-1]
        if r and r[0] == "Kernel Name" and agg:
            break                                               # first launch only
        if len(r) > i_i and r[0].isdigit():
            try:
                agg.append((int(r[i_s] or 0), file, int(r[0]), r[1].strip()[:120], int(r[i_i] or 0)))
            except ValueError:
                pass
    tot = sum(a[0] for a in agg) or 1
    print("total samples", tot, " total warp instructions", sum(a[4] for a in agg))
    for a in sorted(agg, reverse=True)[:n]:
        print("%6d %5.1f%% %s:%d  inst=%d  %s" % (a[0], 100 * a[0] / tot, a[1], a[2], a[4], a[3]))


if __name__ == "__main__":
    main()
