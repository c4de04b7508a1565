#!/usr/bin/env python
"""Median gpu__time_duration per kernel from an ncu launch-list CSV (profiles/README.md)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[h]
ik, iv = H.index("Kernel Name"), H.index("Metric Value")
d = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) > iv:
        d.setdefault(r[ik].split("(")[0], []).append(float(r[iv].replace(",", "")) / 1e3)
for k, v in d.items():
    print("%-28s n=%3d median %8.1f us" % (k, len(v), sorted(v)[len(v) // 2]))
