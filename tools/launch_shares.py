#!/usr/bin/env python
"""Per-kernel totals and shares of an ncu launch list (`--metrics gpu__time_duration.sum --csv`)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-64s n=%4d total %9.3f ms mean %8.1f us share %5.1f %%" % (k[:64], n, t / 1e6, t / n / 1e3, 100 * t / tot))
