#!/usr/bin/env python
"""One-paragraph summary of an ncu --set full report (first kernel in it): duration, DRAM traffic, issue rate,
lane utilisation, occupancy limiters, top stall reasons.  usage: summarize.py report.ncu-rep [label]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
label = sys.argv[2] if len(sys.argv) > 2 else rep
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
g = lambda k: r[hdr.index(k)] if k in hdr else "n/a"
u = lambda k: units[hdr.index(k)] if k in hdr else ""
stalls = sorted(((float(r[i] or 0), h.split("issue_stalled_")[1].split("_per_")[0]) for i, h in enumerate(hdr)
                 if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "selected" not in h), reverse=True)[:4]
print("### %s" % label)
print("* kernel `%s`, grid %s x block %s, %s regs/thread, occupancy limit: smem %s blocks / regs %s blocks" % (
    g("Kernel Name")[:70], g("Grid Size"), g("Block Size"), g("launch__registers_per_thread"), g("launch__occupancy_limit_shared_mem"), g("launch__occupancy_limit_registers")))
print("* duration %s %s (ncu, cold cache, serialised); DRAM read %s %s + write %s %s; DRAM throughput %s %% of peak" % (
    g("gpu__time_duration.sum"), u("gpu__time_duration.sum"), g("dram__bytes_read.sum"), u("dram__bytes_read.sum"), g("dram__bytes_write.sum"), u("dram__bytes_write.sum"),
    g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")))
print("* warp instructions %s; issue slots busy %s %%; active lanes per instruction %s / 32; warps resident per SM %s" % (
    g("smsp__inst_executed.sum"), g("smsp__issue_active.avg.pct_of_peak_sustained_active"), g("smsp__thread_inst_executed_per_inst_executed.ratio"), g("sm__warps_active.avg.per_cycle_active")))
print("* stall cycles per issued instruction: " + ", ".join("%s %.2f" % (n, v) for v, n in stalls))
print("* shared-memory bank conflicts %s; L2 hit rate %s %%" % (g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"), g("lts__t_sector_hit_rate.pct")))
