#!/usr/bin/env python
"""Kernel A/B bench (development aid): device-resident pipeline timings per count-kernel variant and workload.
usage: python tools/kbench.py [--variants 1,2] [--steps 10] [--panel]   (variants = MD_GEN values: candidate generator of count_warp)"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from methyldackel_b200 import _abi as A  # noqa: E402
from methyldackel_b200 import api  # noqa: E402


def dataset(name, args):
    cache = os.environ.get("MDBENCH_CACHE", "/tmp/mdbench"); os.makedirs(cache, exist_ok=True)
    p = os.path.join(cache, name)
    if not os.path.exists(p + ".bam.bai"):
        subprocess.run([os.path.join(ROOT, "methyldackel_b200", "lib", "mdsynth"), "--out", p] + args, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="0,1")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--panel", action="store_true")
    ap.add_argument("--envs", default="", help="semicolon list of extra env settings to sweep, e.g. 'MD_PREFETCH=0;MD_PREFETCH=1'")
    ap.add_argument("--only", default="", help="comma list of option-set names (cpg,all,var) to run")
    a = ap.parse_args()
    sets = [("c2", ["--contigs", "chr1:10000000", "--depth", "30"], "chr1")]
    if a.panel:
        sets.append(("panel", ["--contigs", "amp:150000", "--depth", "2000", "--isize-mean", "180", "--isize-sd", "25", "--isize-min", "150", "--isize-max", "300"], "amp"))
    cfgs = [("cpg", A.default_config()), ("all", A.default_config(keepCHG=1, keepCHH=1)), ("var", A.default_config(minOppositeDepth=5, maxVariantFrac=0.25))]
    for name, sargs, contig in sets:
        p = dataset("kb_" + name, sargs)
        b = api.BamFile(p + ".bam")
        ref = api.fetch_contig(p + ".fa", contig)
        soa = b.read_region(0)
        for cname, cfg in cfgs:
            if a.only and cname not in a.only.split(","):
                continue
            base = None
            for v, envset in [(v, e) for v in a.variants.split(",") for e in (a.envs.split(";") if a.envs else [""])]:
                os.environ["MD_GEN"] = v
                for kv in filter(None, envset.split(",")):
                    os.environ[kv.split("=")[0]] = kv.split("=")[1]
                g = api.GpuContext(cfg)
                g.load_contig(0, ref)
                d = g.upload(soa)
                for _ in range(3):
                    st = g.extract_tile_device(0, 0, len(ref), d)
                tp = tc = 0.0
                for _ in range(a.steps):
                    st = g.extract_tile_device(0, 0, len(ref), d)
                    t = g.last_timing(); tp += t[1]; tc += t[2]
                calls, n = g.fetch_calls(len(ref) + 16)
                import ctypes as C
                sig = hash(bytes(C.string_at(calls, n * 16)))
                if base is None:
                    base = sig
                print(json.dumps({"set": name, "cfg": cname, "variant": int(v), "env": envset, "reads": soa.n_reads, "prep_ms": round(tp / a.steps, 4), "count_ms": round(tc / a.steps, 4),
                                  "M_aln_s": round(soa.n_reads / ((tp + tc) / a.steps) / 1e3, 1), "calls": st.n_calls, "same_as_first": sig == base}), flush=True)
                g.free(d); g.close()
        b.close()


if __name__ == "__main__":
    main()
