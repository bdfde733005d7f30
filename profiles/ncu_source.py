#!/usr/bin/env python
"""Per-source-line view of an ncu report (kernel compiled with -lineinfo, captured with --import-source on).

  ncu_source.py hotspots report.ncu-rep [min_pct]                 lines holding >= min_pct of the executed warp-instructions or stall samples
  ncu_source.py ranges   report.ncu-rep name=lo-hi[,lo-hi...] ... shares per line range of the MAIN source file

Lines are keyed by (file, line): CUDA header lines inlined into the kernel are reported under their own file instead of being
merged into the kernel's.  Before anything is printed the source text embedded in the report is compared with the files on
disk: a report captured from a different build than the tree it is read in (round 1's v9 hot-spot file was: every line
number was off by 72) is refused unless --force is given.
"""
import csv
import os
import subprocess
import sys

STALLS = ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_branch_resolving", "stall_mio", "stall_lg", "stall_math", "stall_not_selected", "stall_no_inst",
          "stall_dispatch", "stall_drain", "stall_imc", "stall_lg_throttle", "stall_membar", "stall_mio_throttle", "stall_sleeping", "stall_tex_throttle")


def load(rep, kernel_filter=None):
    """-> {(file, line): {"src", "inst", "tinst", "samp", "st": {...}}}, function name"""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    agg, hdr, cur_file, func = {}, None, None, None
    for r in csv.reader(txt.splitlines()):
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1]; continue
        if len(r) == 2 and r[0] == "Function Name":
            func = r[1]; continue
        if len(r) > 10 and r[0] == "Line No":
            hdr = r; continue
        if hdr is None or len(r) < len(hdr) or r[2] != "-":          # cuda lines have '-' as the address
            continue
        g = lambda name: float(r[hdr.index(name)] or 0) if name in hdr else 0.0  # noqa: E731
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, {"src": r[1].strip(), "inst": 0, "tinst": 0, "samp": 0, "st": {}})
        a["inst"] += g("Instructions Executed"); a["tinst"] += g("Thread Instructions Executed"); a["samp"] += g("Warp Stall Sampling (All Samples)")
        for k in STALLS:
            a["st"][k] = a["st"].get(k, 0) + g(k)
    return agg, func


def check_against_tree(agg):
    """every line that executed instructions must read the same in the report and on disk"""
    bad, files = [], {}
    for (f, line), a in agg.items():
        if not a["inst"] or f is None or not os.path.exists(f):
            continue
        if f not in files:
            files[f] = open(f, errors="replace").read().splitlines()
        disk = files[f][line - 1].strip() if line - 1 < len(files[f]) else "<beyond end of file>"
        norm = lambda t: t.replace('"', "").replace(" ", "")            # the CSV export drops / doubles quotes inside source text  # noqa: E731
        if norm(disk) != norm(a["src"]):
            bad.append((f, line, a["src"][:80], disk[:80]))
    return bad


def main():
    args = [a for a in sys.argv[1:] if a != "--force"]
    force = "--force" in sys.argv
    mode, rep = args[0], args[1]
    agg, func = load(rep)
    bad = check_against_tree(agg)
    if bad and not force:
        print("REFUSED: %d executed source lines of %s differ from the files on disk — the report was not captured from this tree.  First mismatches:" % (len(bad), rep))
        for f, line, a, b in bad[:5]:
            print("  %s:%d\n    report: %s\n    disk:   %s" % (f, line, a, b))
        sys.exit(2)
    ti = sum(a["inst"] for a in agg.values()); ts = sum(a["samp"] for a in agg.values())
    main_file = max({f for f, _ in agg}, key=lambda f: sum(a["inst"] for (ff, _), a in agg.items() if ff == f))
    print("kernel %s" % func)
    print("total warp-instructions %.0f, stall samples %.0f; source check against the tree: %s" % (ti, ts, "ok" if not bad else "FORCED, %d lines differ" % len(bad)))
    if mode == "hotspots":
        thr = float(args[2]) if len(args) > 2 else 1.5
        for (f, line) in sorted(agg, key=lambda k: (k[0] != main_file, k[0] or "", k[1])):
            a = agg[(f, line)]
            if a["inst"] > ti * thr / 100 or a["samp"] > ts * thr / 100:
                top = sorted(a["st"].items(), key=lambda kv: -kv[1])[:2]
                tag = "L%-4d" % line if f == main_file else "%s:%d" % (os.path.basename(f or "?"), line)
                print("%-22s %5.1f%% inst %5.1f%% stall  lanes %4.1f  %-34s | %s" % (tag, 100 * a["inst"] / ti, 100 * a["samp"] / max(ts, 1), a["tinst"] / max(a["inst"], 1),
                      " ".join("%s=%d" % (k.replace("stall_", ""), v) for k, v in top if v), a["src"][:110]))
    else:
        seen = set()
        for spec in args[2:]:
            n, r = spec.split("=")
            rs = [tuple(int(x) for x in p.split("-")) for p in r.split(",")]
            i = t = s = 0
            for (f, line), a in agg.items():
                if f == main_file and any(lo <= line <= hi for lo, hi in rs):
                    i += a["inst"]; t += a["tinst"]; s += a["samp"]; seen.add((f, line))
            print("%-16s %5.1f%% inst  %5.1f%% stall  lanes %4.1f  (%.2f M warp-inst)" % (n, 100 * i / ti, 100 * s / max(ts, 1), t / max(i, 1), i / 1e6))
        oi = sum(a["inst"] for k, a in agg.items() if k not in seen and k[0] == main_file); os_ = sum(a["samp"] for k, a in agg.items() if k not in seen and k[0] == main_file)
        hi = sum(a["inst"] for k, a in agg.items() if k[0] != main_file); hs = sum(a["samp"] for k, a in agg.items() if k[0] != main_file)
        print("%-16s %5.1f%% inst  %5.1f%% stall  lines: %s" % ("(other lines)", 100 * oi / ti, 100 * os_ / max(ts, 1), sorted(l for (f, l), a in agg.items() if (f, l) not in seen and f == main_file and a["inst"] > ti * 0.002)))
        print("%-16s %5.1f%% inst  %5.1f%% stall  (CUDA headers inlined into the kernel)" % ("(other files)", 100 * hi / ti, 100 * hs / max(ts, 1)))


if __name__ == "__main__":
    main()
