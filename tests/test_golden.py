"""Committed golden vectors (tests/golden/expected, produced by tests/golden/make_golden.py from the reference
build): CPU tier checks the oracle port + host logic against them; the GPU tier checks the drop-in binary."""
import glob
import os
import subprocess
import sys

import pytest

import cases
import oracle_binding as ob

EXP = os.path.join(cases.ROOT, "tests", "golden", "expected")
CASES = [(c[0], c[1], c[2], c[3]) for c in cases.REFERENCE_TESTS] + list(cases.FIXTURE_EXTRA)
NEW_BIN = os.path.join(cases.ROOT, "methyldackel_b200", "lib", "MethylDackel")


def _check(case_dir, prefix):
    exp_files = sorted(f for f in glob.glob(os.path.join(case_dir, "PREFIX*")))
    assert exp_files
    for f in exp_files:
        g = prefix + os.path.basename(f)[len("PREFIX"):]
        assert os.path.exists(g), g
        assert open(g).read().replace(prefix, "PREFIX") == open(f).read(), os.path.basename(f)
    assert len(glob.glob(prefix + "*")) == len(exp_files)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_golden_cpu(built, tmp_path, case):
    name, args, fa, bam = case
    prefix = str(tmp_path / "out")
    assert ob.run_host_main("extract", list(args) + [cases.fx(fa), cases.fx(bam), "-o", prefix], ob.OracleBackend()) == 0
    _check(os.path.join(EXP, name), prefix)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_golden_gpu(built, tmp_path, case):
    name, args, fa, bam = case
    prefix = str(tmp_path / "out")
    r = subprocess.run([NEW_BIN, "extract"] + list(args) + [cases.fx(fa), cases.fx(bam), "-o", prefix], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout == open(os.path.join(EXP, name, "stdout")).read()
    _check(os.path.join(EXP, name), prefix)


@pytest.mark.gpu
def test_golden_mbias_gpu(built, tmp_path):
    r = subprocess.run([NEW_BIN, "mbias", "--txt", "-q", "2", cases.fx("cg100.fa"), cases.fx("cg_aln.bam"), str(tmp_path / "svg")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout == open(os.path.join(EXP, "mbias_cg", "stdout")).read()
    sug = "".join(l + "\n" for l in r.stderr.splitlines() if l.startswith("Suggested inclusion options:"))
    assert sug == open(os.path.join(EXP, "mbias_cg", "suggestion")).read()


# ---- -l BED and perRead golden vectors
@pytest.mark.parametrize("case", cases.FIXTURE_BED, ids=[c[0] for c in cases.FIXTURE_BED])
def test_golden_bed_cpu(built, tmp_path, case):
    name, args, fa, bam = case
    prefix = str(tmp_path / "out")
    assert ob.run_host_main("extract", cases.fx_bed(args) + [cases.fx(fa), cases.fx(bam), "-o", prefix], ob.OracleBackend()) == 0
    _check(os.path.join(EXP, name), prefix)


@pytest.mark.gpu
@pytest.mark.parametrize("case", cases.FIXTURE_BED, ids=[c[0] for c in cases.FIXTURE_BED])
def test_golden_bed_gpu(built, tmp_path, case):
    name, args, fa, bam = case
    prefix = str(tmp_path / "out")
    r = subprocess.run([NEW_BIN, "extract"] + cases.fx_bed(args) + [cases.fx(fa), cases.fx(bam), "-o", prefix], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout == open(os.path.join(EXP, name, "stdout")).read()
    _check(os.path.join(EXP, name), prefix)


@pytest.mark.parametrize("case", cases.FIXTURE_PERREAD, ids=[c[0] for c in cases.FIXTURE_PERREAD])
def test_golden_perread_cpu(built, tmp_path, case):
    name, args, fa, bam = case
    out = str(tmp_path / "pr.txt")
    assert ob.run_host_main("perRead", list(args) + ["-o", out, cases.fx(fa), cases.fx(bam)], ob.OracleBackend()) == 0
    assert open(out).read() == open(os.path.join(EXP, name, "perRead.txt")).read()


@pytest.mark.gpu
@pytest.mark.parametrize("case", cases.FIXTURE_PERREAD, ids=[c[0] for c in cases.FIXTURE_PERREAD])
def test_golden_perread_gpu(built, tmp_path, case):
    name, args, fa, bam = case
    out = str(tmp_path / "pr.txt")
    r = subprocess.run([NEW_BIN, "perRead"] + list(args) + ["-o", out, cases.fx(fa), cases.fx(bam)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out).read() == open(os.path.join(EXP, name, "perRead.txt")).read()
