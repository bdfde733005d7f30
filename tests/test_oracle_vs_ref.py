"""CPU tier (no GPU): pins the oracle port and the product's HOST logic against oracle/_ref
(the reference's own C sources built on oracle/htslib_shim), on the reference's fixture BAMs and on
synthetic BAMs.  The host driver is run with the oracle port bound as its device back end
(tests/oracle_binding.py), so option parsing, BAM decode, tiling, chunk replay and formatting are
all exercised exactly as the GPU run will exercise them."""
import os

import pytest

import cases
import oracle_binding as ob
from util import run_ref, compare_outputs


def _both(built, tmp_path, name, args, fa, bam):
    refp, newp = str(tmp_path / (name + "_ref")), str(tmp_path / (name + "_new"))
    r = run_ref(built["ref_bin"], "extract", args, fa, bam, refp)
    assert r.returncode == 0, r.stderr
    rc = ob.run_host_main("extract", list(args) + [fa, bam, "-o", newp], ob.OracleBackend())
    assert rc == 0
    return refp, newp


@pytest.mark.parametrize("case", cases.REFERENCE_TESTS, ids=[c[0] for c in cases.REFERENCE_TESTS])
def test_reference_testsuite_counts_and_bytes(built, tmp_path, case):
    name, args, fa, bam, counts = case
    refp, newp = _both(built, tmp_path, name, args, cases.fx(fa), cases.fx(bam))
    for suffix, n in cases.counts_for(case).items():
        assert sum(1 for _ in open(refp + suffix)) == n
    assert compare_outputs(refp, newp) == []


@pytest.mark.parametrize("case", [pytest.param(c, marks=pytest.mark.xfail(strict=True, reason="upstream asserts 12 lines; the sources in /root/reference give 11 on every "
                                                                                                   "implementation here (tests/test_reference_test8.py); unresolved without a real htslib build"))
                                  if c[0] in cases.DISPUTED else c for c in cases.REFERENCE_TESTS], ids=[c[0] for c in cases.REFERENCE_TESTS])
def test_upstream_asserted_line_counts(built, tmp_path, case):
    """the numbers exactly as /root/reference/tests/test.py asserts them, against oracle/_ref"""
    name, args, fa, bam, counts = case
    refp = str(tmp_path / (name + "_ref"))
    assert run_ref(built["ref_bin"], "extract", args, cases.fx(fa), cases.fx(bam), refp).returncode == 0
    for suffix, n in counts.items():
        assert sum(1 for _ in open(refp + suffix)) == n, "reference build disagrees with tests/test.py"


@pytest.mark.parametrize("case", cases.FIXTURE_EXTRA, ids=[c[0] for c in cases.FIXTURE_EXTRA])
def test_fixture_extra(built, tmp_path, case):
    name, args, fa, bam = case
    refp, newp = _both(built, tmp_path, name, args, cases.fx(fa), cases.fx(bam))
    assert compare_outputs(refp, newp) == []


@pytest.mark.parametrize("opts", cases.SYNTH_OPTION_SETS, ids=[cases.slug(o) for o in cases.SYNTH_OPTION_SETS])
def test_synthetic_noisy(built, synth, tmp_path, opts):
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    refp, newp = _both(built, tmp_path, "s", opts, p + ".fa", p + ".bam")
    assert compare_outputs(refp, newp) == []


@pytest.mark.parametrize("opts", [[], ["--CHG", "--CHH", "--mergeContext"], ["--cytosine_report", "--CHG", "--CHH"],
                                  ["--minOppositeDepth", "3", "--maxVariantFrac", "0.2", "--CHG"]], ids=["default", "merge", "cyt", "variant"])
def test_synthetic_bismark_nondirectional(built, synth, tmp_path, opts):
    p = synth("bismark", "--contigs", "chrA:40000", "--depth", "40", "--bismark-tags", "--nondirectional", "0.3", "--single-frac", "0.1",
              "--isize-mean", "200", "--isize-sd", "30", "--read-seed", "99")
    refp, newp = _both(built, tmp_path, "b", opts, p + ".fa", p + ".bam")
    assert compare_outputs(refp, newp) == []


def test_deep_overlap_panel(built, synth, tmp_path):
    """config-4 shape in miniature: deep, short inserts, heavy mate overlap"""
    p = synth("panel", "--contigs", "amp:3000", "--depth", "1500", "--isize-mean", "180", "--isize-sd", "25", "--isize-min", "150", "--isize-max", "300")
    refp, newp = _both(built, tmp_path, "p", ["--CHG", "--CHH"], p + ".fa", p + ".bam")
    assert compare_outputs(refp, newp) == []


@pytest.mark.parametrize("opts", [["--noSVG"], ["--noSVG", "--CHG", "--CHH"], ["--noSVG", "--nOT", "3,3,3,3", "--chunkSize", "2500", "--CHG"],
                                  ["--noSVG", "-r", "chr1:1000-30000"]], ids=["cpg", "all", "trim_chunk", "region"])
def test_mbias_txt(built, synth, tmp_path, opts):
    import subprocess, sys, ctypes
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    r = subprocess.run([built["ref_bin"], "mbias"] + opts + [p + ".fa", p + ".bam"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # the host driver prints to this process's stdout: run it in a child to capture
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import oracle_binding as ob; "
            "sys.exit(ob.run_host_main('mbias', %r, ob.OracleBackend()))") % (cases.ROOT, os.path.join(cases.ROOT, "tests"), opts + [p + ".fa", p + ".bam"])
    n = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert n.returncode == 0, n.stderr
    assert n.stdout == r.stdout
    assert len(r.stdout.splitlines()) > 50


def test_mbias_suggestion_line(built, synth, tmp_path):
    import subprocess, sys
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    r = subprocess.run([built["ref_bin"], "mbias", p + ".fa", p + ".bam", str(tmp_path / "svgref")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import oracle_binding as ob; "
            "sys.exit(ob.run_host_main('mbias', %r, ob.OracleBackend()))") % (cases.ROOT, os.path.join(cases.ROOT, "tests"), [p + ".fa", p + ".bam", str(tmp_path / "svgnew")])
    n = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    ref_line = [l for l in r.stderr.splitlines() if l.startswith("Suggested inclusion options:")]
    new_line = [l for l in n.stderr.splitlines() if l.startswith("Suggested inclusion options:")]
    assert ref_line and ref_line == new_line
    _same_svgs(tmp_path, "svgref", "svgnew", expect=["OB", "OT"])


def _same_svgs(tmp_path, ref_prefix, new_prefix, expect=None):
    """the M-bias plots (makeSVGs, svg.c:302-437) byte for byte: same set of <prefix>_<strand>.svg files, same content"""
    ref_files = sorted(f[len(ref_prefix):] for f in os.listdir(str(tmp_path)) if f.startswith(ref_prefix + "_") and f.endswith(".svg"))
    new_files = sorted(f[len(new_prefix):] for f in os.listdir(str(tmp_path)) if f.startswith(new_prefix + "_") and f.endswith(".svg"))
    assert ref_files == new_files and ref_files
    if expect is not None:
        assert ref_files == ["_%s.svg" % e for e in expect]
    for f in ref_files:
        assert open(str(tmp_path / (ref_prefix + f))).read() == open(str(tmp_path / (new_prefix + f))).read(), f


@pytest.mark.parametrize("name,synth_args,opts", [
    ("nondirectional", ["--contigs", "chrA:40000", "--depth", "40", "--bismark-tags", "--nondirectional", "0.3", "--single-frac", "0.1", "--isize-mean", "200", "--isize-sd", "30", "--read-seed", "99"], ["--CHG", "--CHH"]),
    ("short_reads", ["--contigs", "chr1:30000", "--depth", "30", "--readlen", "36", "--isize-mean", "120", "--isize-sd", "20", "--isize-min", "40", "--isize-max", "250"], ["--noCpG", "--CHH", "--nOT", "2,2,3,3"]),
    ("long_reads", ["--contigs", "chr1:60000", "--depth", "20", "--readlen", "251", "--isize-mean", "400", "--isize-sd", "50", "--isize-min", "260", "--isize-max", "700"], ["--CHG"]),
], ids=["nondirectional_all_four_strands", "short_reads_x5_ticks", "long_reads_x10_ticks"])
def test_mbias_svgs(built, synth, tmp_path, name, synth_args, opts):
    import subprocess, sys
    p = synth("svg_" + name, *synth_args)
    r = subprocess.run([built["ref_bin"], "mbias"] + opts + [p + ".fa", p + ".bam", str(tmp_path / "svgref")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import oracle_binding as ob; "
            "sys.exit(ob.run_host_main('mbias', %r, ob.OracleBackend()))") % (cases.ROOT, os.path.join(cases.ROOT, "tests"), opts + [p + ".fa", p + ".bam", str(tmp_path / "svgnew")])
    n = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert n.returncode == 0, n.stderr
    assert n.stdout == r.stdout
    assert [l for l in n.stderr.splitlines() if l.startswith("Suggested")] == [l for l in r.stderr.splitlines() if l.startswith("Suggested")]
    _same_svgs(tmp_path, "svgref", "svgnew")


def test_mbias_svg_fixture(built, tmp_path):
    """the reference's own fixture (4 records): the plot of a strand with a single read pair"""
    import subprocess, sys
    args = ["-q", "2", cases.fx("cg100.fa"), cases.fx("cg_aln.bam")]
    r = subprocess.run([built["ref_bin"], "mbias"] + args + [str(tmp_path / "svgref")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import oracle_binding as ob; "
            "sys.exit(ob.run_host_main('mbias', %r, ob.OracleBackend()))") % (cases.ROOT, os.path.join(cases.ROOT, "tests"), args + [str(tmp_path / "svgnew")])
    n = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert n.returncode == 0, n.stderr
    _same_svgs(tmp_path, "svgref", "svgnew")


def test_parallel_decode_many_small_jobs(built, synth, tmp_path):
    """forces the multi-threaded decoder to cut the BAM into one job per BGZF block, so that records straddle
    nearly every job boundary (stitch path), with 5 decode threads"""
    import subprocess, sys
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    refp, newp = str(tmp_path / "ref"), str(tmp_path / "new")
    r = run_ref(built["ref_bin"], "extract", ["--CHG", "--mergeContext"], p + ".fa", p + ".bam", refp)
    assert r.returncode == 0
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import oracle_binding as ob; "
            "sys.exit(ob.run_host_main('extract', %r, ob.OracleBackend()))") % (cases.ROOT, os.path.join(cases.ROOT, "tests"),
                                                                               ["--CHG", "--mergeContext", "-@", "5", p + ".fa", p + ".bam", "-o", newp])
    n = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, MD_DECODE_JOB_BYTES="1"))
    assert n.returncode == 0, n.stderr
    assert compare_outputs(refp, newp) == []


@pytest.mark.parametrize("sub", ["extract", "mbias"])
@pytest.mark.parametrize("nq", [4, 12, 40], ids=["2bit", "4bit", "8bit"])
def test_small_tiles_carry_and_encodings(built, synth, tmp_path, nq, sub):
    """tiles of 300 alignments: every tile cut carries straddling reads into the next tile, re-encoded into that tile's
    phred alphabet; jobs of one BGZF block each (speculative record chains adopted or re-walked)"""
    import subprocess, sys
    p = synth("st%d" % nq, "--contigs", "chr1:40000,chr2:9000", "--depth", "30", "--quals", str(nq), "--lower-frac", "0.02")
    refp, newp = str(tmp_path / "ref"), str(tmp_path / "new")
    opts = ["--CHG", "--CHH", "--mergeContext"] if sub == "extract" else ["--CHH", "--txt", "--noSVG"]
    tail_ref = [p + ".fa", p + ".bam"] + ([] if sub == "mbias" else [])
    if sub == "extract":
        r = run_ref(built["ref_bin"], "extract", opts, p + ".fa", p + ".bam", refp)
        assert r.returncode == 0
        argv = opts + ["-@", "3", p + ".fa", p + ".bam", "-o", newp]
    else:
        r = subprocess.run([built["ref_bin"], "mbias"] + opts + tail_ref, capture_output=True, text=True)
        assert r.returncode == 0
        argv = opts + ["-@", "3", p + ".fa", p + ".bam"]
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import oracle_binding as ob; "
            "sys.exit(ob.run_host_main(%r, %r, ob.OracleBackend()))") % (cases.ROOT, os.path.join(cases.ROOT, "tests"), sub, argv)
    n = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, MD_DECODE_JOB_BYTES="1", MD_TILE_READS="300"))
    assert n.returncode == 0, n.stderr
    if sub == "extract":
        assert compare_outputs(refp, newp) == []
    else:
        assert r.stdout == n.stdout and len(r.stdout.splitlines()) > 10


@pytest.mark.parametrize("name,args", [
    ("len250", ["--contigs", "chr1:80000", "--depth", "20", "--readlen", "250", "--isize-mean", "400", "--isize-sd", "80", "--isize-min", "250", "--isize-max", "900"]),
    ("len20000", ["--contigs", "chr1:300000", "--depth", "6", "--readlen", "20000", "--isize-mean", "30000", "--isize-sd", "4000", "--isize-min", "20000", "--isize-max", "45000"]),
], ids=["250bp", "20kb"])
def test_synthetic_read_lengths(built, synth, tmp_path, name, args):
    p = synth(name, *args)
    refp, newp = _both(built, tmp_path, name, ["--CHG", "--CHH"], p + ".fa", p + ".bam")
    assert compare_outputs(refp, newp) == []


@pytest.mark.parametrize("nq", [4, 12, 40], ids=["2bit", "4bit", "8bit"])
def test_phred_encodings_cpu(built, synth, tmp_path, nq):
    """host packer + oracle port: 2-bit / 4-bit / plain phred tiles all reproduce the reference"""
    p = synth("q%d" % nq, "--contigs", "chr1:50000", "--depth", "25", "--quals", str(nq))
    refp, newp = _both(built, tmp_path, "q", ["--CHG", "--CHH", "-p", "9"], p + ".fa", p + ".bam")
    assert compare_outputs(refp, newp) == []


@pytest.mark.parametrize("opts", cases.BED_OPTION_SETS, ids=[cases.slug(o) for o in cases.BED_OPTION_SETS])
def test_bed_regions(built, synth, tmp_path, opts):
    """-l / --keepStrand (bed.c; extract.c:353-367, 402-405, 425; common.c:432-439): host BED parsing + chunk skipping with the port's
    restatement of the three BED tests, against the reference build"""
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    refp, newp = _both(built, tmp_path, "bed", cases.with_bed(opts, tmp_path), p + ".fa", p + ".bam")
    assert compare_outputs(refp, newp) == []
    assert sum(1 for _ in open([f for f in __import__("glob").glob(refp + "*")][0])) > 100


@pytest.mark.parametrize("opts", cases.BED_MBIAS_SETS, ids=["plain", "strand_allctx"])
def test_bed_mbias(built, synth, tmp_path, opts):
    import subprocess
    import sys
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    o = cases.with_bed(opts, tmp_path)
    r = subprocess.run([built["ref_bin"], "mbias"] + o + [p + ".fa", p + ".bam"], capture_output=True, text=True)
    code = ("import sys; sys.path[:0]=[%r,%r]; import oracle_binding as ob; "
            "sys.exit(ob.run_host_main('mbias', %r, ob.OracleBackend()))") % (cases.ROOT, os.path.join(cases.ROOT, "tests"), o + [p + ".fa", p + ".bam"])
    n = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and n.returncode == 0, (r.stderr, n.stderr)
    assert n.stdout == r.stdout and len(r.stdout.splitlines()) > 50


def test_bed_malformed_is_refused(built, synth, tmp_path):
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    bad = str(tmp_path / "bad.bed")
    open(bad, "w").write("chrUnknown\t1\t2\n")
    assert ob.run_host_main("extract", ["-l", bad, p + ".fa", p + ".bam", "-o", str(tmp_path / "x")], ob.OracleBackend()) == 1   # extract.c:1473-1476


def _perread_both(built, tmp_path, args, fa, bam):
    import subprocess
    refp, newp = str(tmp_path / "pr_ref.txt"), str(tmp_path / "pr_new.txt")
    r = subprocess.run([built["ref_bin"], "perRead"] + list(args) + ["-o", refp, fa, bam], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert ob.run_host_main("perRead", list(args) + ["-o", newp, fa, bam], ob.OracleBackend()) == 0
    return open(refp).read(), open(newp).read()


@pytest.mark.parametrize("opts", cases.PERREAD_SETS, ids=[cases.slug(o) for o in cases.PERREAD_SETS])
def test_perread_synthetic(built, synth, tmp_path, opts):
    """perRead (perRead.c:37-94, 183-196): host driver + the port's restatement of processRead against the reference build"""
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    a, b = _perread_both(built, tmp_path, cases.with_bed(opts, tmp_path), p + ".fa", p + ".bam")
    assert a == b and len(a.splitlines()) > 500


@pytest.mark.parametrize("fx", cases.PERREAD_FIXTURES, ids=[f[1] for f in cases.PERREAD_FIXTURES])
def test_perread_fixtures(built, tmp_path, fx):
    fa, bam, args = fx
    a, b = _perread_both(built, tmp_path, args, cases.fx(fa), cases.fx(bam))
    assert a == b and a


def test_perread_indels_clips_long_reads(built, synth, tmp_path):
    """odd read length (pad nibble behind the sequence), indels and clips next to low-phred bases: the skipped-base quirk crosses CIGAR op ends"""
    p = synth("len75", "--contigs", "chr1:50000", "--depth", "25", "--readlen", "75", "--isize-mean", "160", "--isize-sd", "40", "--isize-min", "75", "--isize-max", "400")
    a, b = _perread_both(built, tmp_path, ["-p", "12", "-q", "0"], p + ".fa", p + ".bam")
    assert a == b and len(a.splitlines()) > 500


@pytest.mark.parametrize("opts", [["--noSVG", "--minConversionEfficiency", "0.97", "--CHH"], ["--noSVG", "--minConversionEfficiency", "0.9", "--CHG", "--nOT", "3,3,3,3"]],
                         ids=["ce097_chh", "ce09_chg_trim"])
def test_mbias_conversion_efficiency(built, synth, tmp_path, opts):
    """mbias --minConversionEfficiency: the filter sees the chunk's own window contig[localPos, localEnd] (MBias.c:147,154-156); one
    chunk per contig keeps every read inside it (see cases.SYNTH_OPTION_SETS)"""
    import subprocess
    import sys
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    r = subprocess.run([built["ref_bin"], "mbias"] + opts + [p + ".fa", p + ".bam"], capture_output=True, text=True)
    r0 = subprocess.run([built["ref_bin"], "mbias", "--noSVG"] + opts[3:] + [p + ".fa", p + ".bam"], capture_output=True, text=True)
    code = ("import sys; sys.path[:0]=[%r,%r]; import oracle_binding as ob; "
            "sys.exit(ob.run_host_main('mbias', %r, ob.OracleBackend()))") % (cases.ROOT, os.path.join(cases.ROOT, "tests"), opts + [p + ".fa", p + ".bam"])
    n = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and n.returncode == 0, (r.stderr, n.stderr)
    assert n.stdout == r.stdout and len(r.stdout.splitlines()) > 50
    assert r.stdout != r0.stdout                      # the filter does drop alignments on this data set


BED_VARIANTS = {
    "unsorted_nested": "chr1\t30000\t31000\nchr1\t100\t50000\nchr1\t200\t300\nchr2\t5\t6\nchr1\t100\t200\n",
    "headers_then_blank_line_ends_file": "browser position chr1\ntrack name=x\n# note\nchr1\t1000\t2000\tn\t0\t-\nchr2\t100\t14000\tm\t0\t+\n\nchr1\t30000\t40000\n",
    "past_contig_end": "chr1\t59000\t70000\nchr2\t14990\t99999\n",
    "short_columns_keepstrand": "chr1\t1000\t9000\tname\nchr1\t20000\t29000\tname\t5\nchr2\t0\t15000\tx\t1\t-\textra\n",
    "adjacent_and_single_base": "chr1\t500\t501\nchr1\t501\t502\nchr1\t502\t600\nchr1\t600\t601\n",
}


@pytest.mark.parametrize("name", sorted(BED_VARIANTS))
@pytest.mark.parametrize("gz", [False, True], ids=["plain", "gzip"])
def test_bed_file_variants(built, synth, tmp_path, name, gz):
    """parseBED leniencies (bed.c:91-236): header lines, blank lines, unsorted and nested regions, ends past the contig, missing
    strand columns under --keepStrand, gzip-compressed input"""
    import gzip
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    f = str(tmp_path / ("r.bed.gz" if gz else "r.bed"))
    (gzip.open(f, "wt") if gz else open(f, "w")).write(BED_VARIANTS[name])
    refp, newp = _both(built, tmp_path, "bv", ["-l", f, "--keepStrand", "--CHH", "--chunkSize", "4000", "--cytosine_report"], p + ".fa", p + ".bam")
    assert compare_outputs(refp, newp) == []
