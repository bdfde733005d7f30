#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: share of executed instructions, of stall samples, lane
utilisation and the dominant stall reasons.  usage: ncu_hotspots.py report.ncu-rep [min_pct]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
agg = {}
hdr = None
for r in rows:
    if len(r) > 10 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":      # cuda lines have '-' as the address
        continue
    g = lambda name: float(r[hdr.index(name)] or 0)
    line = int(r[0]); src = r[1].strip()
    a = agg.setdefault(line, {"src": src, "inst": 0, "tinst": 0, "samp": 0, "st": {}})
    a["inst"] += g("Instructions Executed"); a["tinst"] += g("Thread Instructions Executed"); a["samp"] += g("Warp Stall Sampling (All Samples)")
    for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_branch_resolving", "stall_mio", "stall_lg", "stall_math", "stall_not_selected", "stall_no_inst"):
        a["st"][k] = a["st"].get(k, 0) + g(k)
ti = sum(a["inst"] for a in agg.values()); ts = sum(a["samp"] for a in agg.values())
print("total warp-instructions %.0f, stall samples %.0f" % (ti, ts))
for line in sorted(agg):
    a = agg[line]
    if a["inst"] > ti * thr / 100 or a["samp"] > ts * thr / 100:
        top = sorted(a["st"].items(), key=lambda kv: -kv[1])[:2]
        print("L%-4d %5.1f%% inst %5.1f%% stall  lanes %4.1f  %-34s | %s" % (line, 100 * a["inst"] / ti, 100 * a["samp"] / max(ts, 1), a["tinst"] / max(a["inst"], 1),
              " ".join("%s=%d" % (k.replace("stall_", ""), v) for k, v in top if v), a["src"][:110]))
