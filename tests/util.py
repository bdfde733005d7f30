import glob
import os
import subprocess


def run_ref(ref_bin, sub, args, fa, bam, prefix, extra_pos=()):
    cmd = [ref_bin, sub] + list(args) + [fa, bam] + list(extra_pos)
    if sub == "extract":
        cmd += ["-o", prefix]
    return subprocess.run(cmd, capture_output=True, text=True)


def outputs(prefix):
    return sorted(f for f in glob.glob(prefix + "*") if not f.endswith((".stdout", ".stderr")))


def compare_outputs(ref_prefix, new_prefix):
    """Byte-compares every file the reference wrote with its counterpart; the track-header line embeds the
    output prefix (extract.c:563), so the prefixes are normalised first. Returns a list of problems."""
    problems = []
    refs = outputs(ref_prefix)
    if not refs:
        problems.append("reference produced no files for " + ref_prefix)
    for f in refs:
        g = new_prefix + f[len(ref_prefix):]
        if not os.path.exists(g):
            problems.append("missing " + g)
            continue
        a = open(f).read().replace(ref_prefix, "PREFIX")
        b = open(g).read().replace(new_prefix, "PREFIX")
        if a != b:
            la, lb = a.splitlines(), b.splitlines()
            first = next((i for i, (x, y) in enumerate(zip(la, lb)) if x != y), min(len(la), len(lb)))
            problems.append("%s differs: %d vs %d lines, first difference at line %d: %r vs %r" % (
                os.path.basename(f), len(la), len(lb), first, la[first] if first < len(la) else None, lb[first] if first < len(lb) else None))
    news = outputs(new_prefix)
    if len(news) != len(refs):
        problems.append("file sets differ: %s vs %s" % ([os.path.basename(x) for x in refs], [os.path.basename(x) for x in news]))
    return problems
