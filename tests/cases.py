"""Shared case tables for the parity tests."""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FX = os.path.join(ROOT, "tests", "golden", "fixtures")


def fx(n):
    return os.path.join(FX, n)


# The 15 commands of the reference's tests/test.py (file:line in comments) with the line counts it asserts — all 15 exactly
# as upstream holds them.  Test 8 (12 lines) is not reproduced by the reference's own sources as they lie in /root/reference
# (11 lines by oracle/_ref, by the oracle port, by the CUDA path and by the independent Python model tests/pymodel.py);
# DISPUTED records that, tests/test_reference_test8.py holds the derivation, and the upstream number stays in the table as a
# strict xfail until a real-htslib build says otherwise.
REFERENCE_TESTS = [
    ("t01", ["-q", "2"], "ct100.fa", "ct_aln.bam", {"_CpG.bedGraph": 1}),                                   # test.py:18
    ("t02", ["-q", "2"], "cg100.fa", "cg_aln.bam", {"_CpG.bedGraph": 49}),                                  # test.py:25 (>1)
    ("t03", ["-q", "10"], "cg100.fa", "cg_aln.bam", {"_CpG.bedGraph": 1}),                                  # test.py:35
    ("t04", ["--methylKit", "--CHH", "--CHG", "-q", "2"], "cg100.fa", "cg_aln.bam",
     {"_CpG.methylKit": 49, "_CHG.methylKit": 1, "_CHH.methylKit": 2}),                                      # test.py:45
    ("t05", ["--minDepth", "2", "-q", "2"], "cg100.fa", "cg_aln.bam", {"_CpG.bedGraph": 1}),                # test.py:60
    ("t06", ["--ignoreFlags", "0xD00", "-q", "2"], "cg100.fa", "cg_aln.bam", {"_CpG.bedGraph": 49}),        # test.py:68
    ("t07", ["--requireFlags", "0xD00", "-q", "2"], "cg100.fa", "cg_aln.bam", {"_CpG.bedGraph": 49}),       # test.py:76
    ("t08", ["--nOT", "50,50,40,40", "-q", "2"], "cg100.fa", "cg_aln.bam", {"_CpG.bedGraph": 12}),          # test.py:84-87
    ("t09", ["-p", "1", "-q", "0", "--minOppositeDepth", "3", "--maxVariantFrac", "0.25"], "cg100.fa", "cg_with_variants.bam",
     {"_CpG.bedGraph": 48}),                                                                                  # test.py:92
    ("t10", [], "chgchh.fa", "chgchh_aln.bam", {"_CpG.bedGraph": 2}),                                        # test.py:101
    ("t11", ["-q", "5"], "chgchh.fa", "chgchh_aln.bam", {"_CpG.bedGraph": 3}),                               # test.py:109
    ("t12", ["-q", "5", "--minConversionEfficiency", "0.9"], "chgchh.fa", "chgchh_aln.bam", {"_CpG.bedGraph": 2}),         # test.py:117
    ("t13", ["-q", "5", "--minConversionEfficiency", "1.0"], "chgchh.fa", "chgchh_aln.bam", {"_CpG.bedGraph": 1}),         # test.py:125
    ("t14", ["-q", "1"], "cg100.fa", "NH.bam", {"_CpG.bedGraph": 1}),                                        # test.py:133
    ("t15", ["--ignoreNH", "-q", "1"], "cg100.fa", "NH.bam", {"_CpG.bedGraph": 49}),                         # test.py:141
]

DISPUTED = {"t08": {"_CpG.bedGraph": 11}}


def counts_for(case):
    """line counts every implementation in this repository produces (== upstream's except for DISPUTED)"""
    return DISPUTED.get(case[0], case[4])


FIXTURE_EXTRA = [
    ("x_cyt", ["-q", "2", "--cytosine_report", "--CHG", "--CHH"], "cg100.fa", "cg_aln.bam"),
    ("x_mrg", ["-q", "2", "--mergeContext", "--CHG"], "cg100.fa", "cg_aln.bam"),
    ("x_all", ["-q", "5", "--CHG", "--CHH"], "chgchh.fa", "chgchh_aln.bam"),
    ("x_allmrg", ["-q", "0", "--CHG", "--CHH", "--mergeContext"], "chgchh.fa", "chgchh_aln.bam"),
    ("x_F0", ["-q", "0", "-F", "0", "--keepDupes", "-p", "1"], "cg100.fa", "cg_aln.bam"),
    ("x_var", ["-p", "1", "-q", "0", "--minOppositeDepth", "1", "--maxVariantFrac", "0.1", "--mergeContext"], "cg100.fa", "cg_with_variants.bam"),
]

# option sets exercised on synthetic BAMs (each is diffed byte-for-byte against oracle/_ref)
SYNTH_OPTION_SETS = [
    [],
    ["--fraction"], ["--counts"], ["--logit"],
    ["--CHG", "--CHH"],
    ["--mergeContext", "--CHG", "--CHH"],
    ["--cytosine_report", "--CHG", "--CHH"],
    ["--methylKit", "--CHG", "--CHH"],
    ["--CHG", "--CHH", "--minOppositeDepth", "2", "--maxVariantFrac", "0.1"],
    ["--mergeContext", "--CHG", "--minOppositeDepth", "2", "--maxVariantFrac", "0.1"],
    ["--OT", "5,140,10,130", "--OB", "3,0,0,120"],
    ["--nOT", "5,5,7,7", "--nOB", "2,3,4,5", "--nCTOT", "1,1,1,1", "--nCTOB", "9,9,9,9"],
    ["-F", "0", "--keepDupes", "--keepSingleton", "--keepDiscordant", "-q", "0", "-p", "1", "--ignoreNH"],
    ["--chunkSize", "777", "--mergeContext", "--CHG"],
    ["--chunkSize", "1000", "--cytosine_report"],
    ["-r", "chr1:5000-20000", "--CHG"],
    ["-r", "chr2", "--mergeContext"],
    ["-d", "5", "--noCpG", "--CHH"],
    ["-q", "40", "-p", "20", "-R", "2"],
    # the conversion-efficiency filter is only well defined in the reference for reads that start inside the chunk window
    # (it indexes the window with pos - offset, common.c:379); one chunk per contig keeps every read inside
    ["--minConversionEfficiency", "0.97", "--CHH"],
    ["--minConversionEfficiency", "0.9", "--mergeContext", "--OT", "3,140,3,140"],
]


def slug(opts):
    s = "_".join(o.strip("-").replace(",", ".").replace(":", ".") for o in opts) or "default"
    return s[:60]


# -l <BED> (bed.c): regions with overlaps, nesting, strands, a region past the contig end, header/comment lines.
# "@BED" in an option set stands for the path of this file (written by the bed_file fixture helper below).
BED_TEXT = """# comment
track name=foo
chr1\t1000\t3000\ta\t0\t+
chr1\t2500\t2600\tb\t0\t-
chr1\t10000\t10050\tc\t0\t-
chr1\t20000\t29999\td\t0\t+
chr1\t20500\t40000\te\t0\t.
chr1\t59990\t60050
chr2\t0\t100\tx\t1\t-
chr2\t7000\t9000\ty\t1\t+
"""
BED_OPTION_SETS = [
    ["-l", "@BED"],
    ["-l", "@BED", "--keepStrand"],
    ["-l", "@BED", "--keepStrand", "--CHG", "--CHH", "--mergeContext"],
    ["-l", "@BED", "--keepStrand", "--cytosine_report", "--CHH", "--chunkSize", "5000"],
    ["-l", "@BED", "--keepStrand", "--minOppositeDepth", "2", "--maxVariantFrac", "0.1", "--mergeContext", "--CHG"],
    ["-l", "@BED", "-r", "chr1:2000-25000", "--chunkSize", "700"],
    ["-l", "@BED", "--keepStrand", "--methylKit", "--CHH", "-F", "0", "--keepDupes"],
]
BED_MBIAS_SETS = [["--noSVG", "-l", "@BED"], ["--noSVG", "-l", "@BED", "--keepStrand", "--CHG", "--CHH", "--chunkSize", "2500"]]


def with_bed(opts, tmp_path):
    """Writes BED_TEXT under tmp_path and substitutes its path for "@BED"."""
    f = os.path.join(str(tmp_path), "regions.bed")
    if not os.path.exists(f):
        open(f, "w").write(BED_TEXT)
    return [f if o == "@BED" else o for o in opts]


# perRead (perRead.c): option sets on the synthetic "noisy" data set and on the reference's fixtures
PERREAD_SETS = [
    [],
    ["-q", "0", "-p", "20", "--chunkSize", "777"],          # many low-phred skips (perRead.c:59-63), chunk windows every 777 bp
    ["-r", "chr1:5000-20000", "-F", "1024"],
    ["-l", "@BED", "--chunkSize", "3000", "-R", "64"],      # chunk-level BED skipping (perRead.c:159-173)
    ["-p", "40", "-q", "60", "-@", "4"],
]
PERREAD_FIXTURES = [("cg100.fa", "cg_aln.bam", ["-q", "2"]), ("chgchh.fa", "chgchh_aln.bam", ["-q", "0", "-p", "1"]), ("ct100.fa", "ct_aln.bam", ["-q", "0"]),
                    ("cg100.fa", "cg_with_variants.bam", ["-q", "0", "-p", "30"]), ("cg100.fa", "NH.bam", ["-q", "0"])]


# golden cases on the reference's fixtures for the "next" rows: -l BED (file committed under tests/golden/fixtures) and perRead
FIXTURE_BED = [
    ("bed_plain", ["-q", "2", "-l", "@FXBED"], "cg100.fa", "cg_aln.bam"),
    ("bed_strand_all", ["-q", "2", "-l", "@FXBED", "--keepStrand", "--CHG", "--CHH", "--cytosine_report"], "cg100.fa", "cg_aln.bam"),
    ("bed_merge_var", ["-p", "1", "-q", "0", "-l", "@FXBED", "--keepStrand", "--mergeContext", "--minOppositeDepth", "1", "--maxVariantFrac", "0.1"], "cg100.fa", "cg_with_variants.bam"),
]
FIXTURE_PERREAD = [("pr_" + f[1].split(".")[0], f[2], f[0], f[1]) for f in PERREAD_FIXTURES]


def fx_bed(opts):
    return [fx("cg100.bed") if o == "@FXBED" else o for o in opts]
