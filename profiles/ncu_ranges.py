#!/usr/bin/env python
"""Share of executed warp-instructions / stall samples per source-line RANGE of an ncu report (with -lineinfo + --import-source).
usage: ncu_ranges.py report.ncu-rep name=lo-hi[,lo-hi...] ..."""
import csv, subprocess, sys
rep = sys.argv[1]
ranges = []
for a in sys.argv[2:]:
    n, r = a.split("=")
    ranges.append((n, [tuple(int(x) for x in p.split("-")) for p in r.split(",")]))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None; per = {}
for r in rows:
    if len(r) > 10 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-": continue
    g = lambda name: float(r[hdr.index(name)] or 0)
    a = per.setdefault(int(r[0]), [0, 0, 0]); a[0] += g("Instructions Executed"); a[1] += g("Thread Instructions Executed"); a[2] += g("Warp Stall Sampling (All Samples)")
ti = sum(a[0] for a in per.values()); ts = sum(a[2] for a in per.values())
print("total warp-instructions %.0f, stall samples %.0f" % (ti, ts))
seen = set()
for n, rs in ranges:
    i = t = s = 0
    for line, a in per.items():
        if any(lo <= line <= hi for lo, hi in rs): i += a[0]; t += a[1]; s += a[2]; seen.add(line)
    print("%-14s %5.1f%% inst  %5.1f%% stall  lanes %4.1f  (%.1f M warp-inst)" % (n, 100 * i / ti, 100 * s / max(ts, 1), t / max(i, 1), i / 1e6))
i = sum(a[0] for l, a in per.items() if l not in seen); s = sum(a[2] for l, a in per.items() if l not in seen)
print("%-14s %5.1f%% inst  %5.1f%% stall  lines: %s" % ("(other)", 100 * i / ti, 100 * s / max(ts, 1), sorted(l for l in per if l not in seen and per[l][0] > ti * 0.002)))
