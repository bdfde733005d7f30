"""The one reference-held vector this repository does not reproduce: tests/test.py:84-87 asserts 12 lines for
`extract --nOT 50,50,40,40 cg100.fa cg_aln.bam -q 2`.

What is checked here, without htslib (none exists in this image or on the GPU box: gpurun_out/probe_r2.txt):

1. An independent implementation (tests/pymodel.py: Python gzip + a column-major pileup written from the SAM specification, no code
   shared with oracle/htslib_shim or oracle/md_oracle.c) reproduces the other 14 asserted counts AND agrees call for call with
   oracle/_ref on all 15 commands.  So the shim's BGZF/pileup is not what makes test 8 come out at 11.
2. The derivation of 11 from the reference's own lines, executable:
   cg_aln.bam holds 4 records named read1 at position 0, 100M, MAPQ 6: flags 0x63, 0x93, 0x263, 0x293.  -F 0xF00 (default)
   drops the two QC-fail records (common.c:418).  Both remaining records carry XG:Z:CT, so getStrand gives OT for 0x63 (read 1
   forward, common.c:102) and OT for 0x93 (read 2 reverse, common.c:105).  --nOT 50,50,40,40 fills absoluteBounds[0..3]
   (extract.c:860).  trimAbsoluteAlignment (common.c:174-208): read 1 uses bounds[0],[1] = 50,50 -> query indices [0,50) and
   [50,100) become N/0: nothing is left; read 2 uses bounds[2],[3] = 40,40 -> [0,40) and 99 down to 60 become N/0: query
   indices 40..59 survive.  The overlap merge (overlaps.c:91-100) sees N/0 against C/40 or G/40 there: `b` keeps its phred.
   Reference chrCG is (cg)x49 + "cA": C columns 40,42,...,58 are CpG, read 2 shows C at each: 10 calls, 10 + header = 11.
3. What a 12th line would need: exactly one more kept CpG C column, i.e. column 60 — the right-hand trim removing 39 bases
   instead of 40 (an `i < rb - 1` / `l_qseq - i` form of the loop at common.c:199-205).  The model run with that variant gives 12.
   The loop in /root/reference/common.c does not have that form, so either upstream's assertion predates the present loop, or a
   real-htslib build differs in a way none of the three implementations here can see.  Recorded as open; the table keeps 12.
"""
import pytest

import cases
import pymodel
from util import run_ref


def _model_kwargs(args):
    kw, it = {}, iter(args)
    for a in it:
        if a == "-q":
            kw["q"] = int(next(it))
        elif a == "-p":
            kw["p"] = int(next(it))
        elif a in ("--ignoreFlags", "-F"):
            v = next(it)
            kw["F"] = int(v) if v.isdigit() else 0                 # atoi("0xD00") == 0
        elif a in ("--requireFlags", "-R"):
            v = next(it)
            kw["R"] = int(v) if v.isdigit() else 0
        elif a == "--minDepth":
            kw["min_depth"] = int(next(it))
        elif a == "--nOT":
            kw["abs_bounds"] = [int(x) for x in next(it).split(",")] + [0] * 12
        elif a == "--minOppositeDepth":
            kw["min_opp"] = int(next(it))
        elif a == "--maxVariantFrac":
            kw["max_var"] = float(next(it))
        elif a == "--minConversionEfficiency":
            import struct
            kw["min_ce"] = struct.unpack("f", struct.pack("f", float(next(it))))[0]     # Config.minConversionEfficiency is a float
        elif a == "--ignoreNH":
            kw["ignore_nh"] = True
        elif a in ("--CHG", "--CHH"):
            k = list(kw.get("keep", (1, 0, 0)))
            k[1 if a == "--CHG" else 2] = 1
            kw["keep"] = tuple(k)
        elif a == "--methylKit":
            pass
        else:
            raise AssertionError("option not modelled: " + a)
    return kw


SUFFIX_CTX = {"_CpG": 0, "_CHG": 1, "_CHH": 2}


@pytest.mark.parametrize("case", [pytest.param(c, marks=pytest.mark.xfail(strict=True, reason="see the module docstring: 11 from the sources in /root/reference, upstream asserts 12"))
                                  if c[0] in cases.DISPUTED else c for c in cases.REFERENCE_TESTS], ids=[c[0] for c in cases.REFERENCE_TESTS])
def test_python_model_reproduces_upstream_counts(case):
    name, args, fa, bam, counts = case
    got = pymodel.extract(cases.fx(bam), cases.fx(fa), **_model_kwargs(args))
    for suffix, n in counts.items():
        assert 1 + len(got[SUFFIX_CTX[suffix.split(".")[0]]]) == n


@pytest.mark.parametrize("case", cases.REFERENCE_TESTS, ids=[c[0] for c in cases.REFERENCE_TESTS])
def test_python_model_equals_reference_build_call_for_call(built, tmp_path, case):
    """position, methylated and unmethylated count of every output line: model == oracle/_ref (default bedGraph or methylKit)"""
    name, args, fa, bam, counts = case
    got = pymodel.extract(cases.fx(bam), cases.fx(fa), **_model_kwargs(args))
    refp = str(tmp_path / "ref")
    assert run_ref(built["ref_bin"], "extract", args, cases.fx(fa), cases.fx(bam), refp).returncode == 0
    for suffix in counts:
        rows = [l.rstrip("\n").split("\t") for l in open(refp + suffix)][1:]
        if suffix.endswith(".methylKit"):                              # chrBase chr base(1-based) strand coverage freqC freqT
            ref_calls = [(int(r[2]) - 1, int(r[4])) for r in rows]
            assert ref_calls == [(p, m + u) for p, m, u in got[SUFFIX_CTX[suffix.split(".")[0]]]]
        else:
            assert [(int(r[1]), int(r[4]), int(r[5])) for r in rows] == got[SUFFIX_CTX[suffix.split(".")[0]]]


def test_test8_derivation_step_by_step():
    recs = pymodel.read_bam(cases.fx("cg_aln.bam"))
    assert [(r["flag"], r["pos"], r["mapq"], r["cigar"]) for r in recs] == [(0x63, 0, 6, [(100, 0)]), (0x93, 0, 6, [(100, 0)]), (0x263, 0, 6, [(100, 0)]), (0x293, 0, 6, [(100, 0)])]
    kept = [r for r in recs if not r["flag"] & 0xF00]
    assert [pymodel.get_strand(r) for r in kept] == [1, 1]                     # both OT
    b = [50, 50, 40, 40] + [0] * 12
    for r in kept:
        pymodel.trim(r, 1, b, True)
    assert all(x == 15 for x in kept[0]["seq"]) and all(x == 0 for x in kept[0]["qual"])          # read 1: nothing left
    assert [i for i, x in enumerate(kept[1]["qual"]) if x] == list(range(40, 60))                  # read 2: query 40..59
    pymodel.merge(kept[0], kept[1])
    assert [i for i, x in enumerate(kept[1]["qual"]) if x >= 5] == list(range(40, 60))             # the merge takes nothing away
    ref = "".join(l.strip() for l in open(cases.fx("cg100.fa")) if not l.startswith(">"))
    cpg_c = [p for p in range(40, 60) if pymodel.ctx(ref, p, 0, len(ref)) == (0, 1)]
    assert cpg_c == list(range(40, 60, 2)) and len(cpg_c) == 10                                     # 10 calls + header = 11 lines


def test_what_would_make_it_twelve(monkeypatch):
    """a right-hand trim one base shorter (39 instead of 40) keeps column 60 and gives upstream's 12"""
    orig = pymodel.trim

    def trim_one_short(r, strand, bounds, absolute):
        if absolute:
            bounds = list(bounds)
            k = 4 * (strand - 1) + (2 if r["flag"] & 0x80 else 0)
            bounds[k + 1] = max(0, bounds[k + 1] - 1)
        orig(r, strand, bounds, absolute)
    monkeypatch.setattr(pymodel, "trim", trim_one_short)
    got = pymodel.extract(cases.fx("cg_aln.bam"), cases.fx("cg100.fa"), q=2, abs_bounds=[50, 50, 40, 40] + [0] * 12)
    assert 1 + len(got[0]) == 12 and got[0][-1][0] == 60
