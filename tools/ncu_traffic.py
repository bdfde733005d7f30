#!/usr/bin/env python
"""DRAM bytes read + written by the one kernel launch in an ncu report -> JSON (what bench.py reports as roofline.traffic)."""
import csv
import json
import subprocess
import sys

rep, note = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]


def val(k):
    v = float(r[hdr.index(k)].replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[hdr.index(k)], 1)


gi = hdr.index("Grid Size") if "Grid Size" in hdr else None
out = {"kernel": r[hdr.index("Kernel Name")], "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
       "duration_ns": val("gpu__time_duration.sum") if "gpu__time_duration.sum" in hdr else None, "grid": r[gi] if gi is not None else None, "source": note}
print(json.dumps(out, indent=1))
