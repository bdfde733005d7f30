#!/usr/bin/env python
"""Regenerates tests/golden/expected/ from oracle/_ref/MethylDackel (the reference's own C sources built by
oracle/Makefile).  Run it in the authoring container, where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

Each case directory holds the files the reference wrote for one command of tests/cases.py on the fixture
BAMs in tests/golden/fixtures/ (copies of the reference's tests/*.bam, *.bai, *.fa), with the output
prefix normalised to PREFIX, plus stdout.  tests/test_golden.py compares against these without needing
the reference binary."""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "MethylDackel")


def main():
    out = os.path.join(HERE, "expected")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    todo = [(c[0], c[1], c[2], c[3]) for c in cases.REFERENCE_TESTS] + list(cases.FIXTURE_EXTRA)
    for name, args, fa, bam in todo:
        d = os.path.join(out, name)
        os.makedirs(d)
        prefix = os.path.join(d, "PREFIX")
        r = subprocess.run([REF, "extract"] + args + [cases.fx(fa), cases.fx(bam), "-o", prefix], capture_output=True, text=True, check=True)
        open(os.path.join(d, "stdout"), "w").write(r.stdout)
        for f in glob.glob(prefix + "*"):
            txt = open(f).read().replace(prefix, "PREFIX")
            open(f, "w").write(txt)
    # -l BED and perRead on the fixtures
    for name, args, fa, bam in cases.FIXTURE_BED:
        d = os.path.join(out, name)
        os.makedirs(d)
        prefix = os.path.join(d, "PREFIX")
        r = subprocess.run([REF, "extract"] + cases.fx_bed(args) + [cases.fx(fa), cases.fx(bam), "-o", prefix], capture_output=True, text=True, check=True)
        open(os.path.join(d, "stdout"), "w").write(r.stdout)
        for f in glob.glob(prefix + "*"):
            txt = open(f).read().replace(prefix, "PREFIX")
            open(f, "w").write(txt)
    for name, args, fa, bam in cases.FIXTURE_PERREAD:
        d = os.path.join(out, name)
        os.makedirs(d)
        subprocess.run([REF, "perRead"] + args + ["-o", os.path.join(d, "perRead.txt"), cases.fx(fa), cases.fx(bam)], capture_output=True, text=True, check=True)
    # mbias on a fixture: --txt table and the suggestion line
    d = os.path.join(out, "mbias_cg")
    os.makedirs(d)
    r = subprocess.run([REF, "mbias", "--txt", "-q", "2", cases.fx("cg100.fa"), cases.fx("cg_aln.bam"), os.path.join(d, "svg")], capture_output=True, text=True, check=True)
    open(os.path.join(d, "stdout"), "w").write(r.stdout)
    open(os.path.join(d, "suggestion"), "w").write("".join(l + "\n" for l in r.stderr.splitlines() if l.startswith("Suggested inclusion options:")))
    for f in glob.glob(os.path.join(d, "svg*")):
        os.unlink(f)
    print("wrote", len(os.listdir(out)), "cases to", out)


if __name__ == "__main__":
    main()
