"""GPU tier: the CUDA path (through the C ABI / the drop-in binary) against the oracle.
Byte-exact: every output file of `lib/MethylDackel` must equal the file oracle/_ref writes for the
same command, and every md_call record must equal the oracle port's."""
import ctypes as C
import os
import subprocess

import pytest

import cases
import oracle_binding as ob
from methyldackel_b200 import _abi as A
from methyldackel_b200 import api
from util import run_ref, compare_outputs

pytestmark = pytest.mark.gpu

NEW_BIN = os.path.join(cases.ROOT, "methyldackel_b200", "lib", "MethylDackel")


def _both(built, tmp_path, name, args, fa, bam):
    refp, newp = str(tmp_path / (name + "_ref")), str(tmp_path / (name + "_new"))
    r = run_ref(built["ref_bin"], "extract", args, fa, bam, refp)
    assert r.returncode == 0, r.stderr
    n = subprocess.run([NEW_BIN, "extract"] + list(args) + [fa, bam, "-o", newp], capture_output=True, text=True)
    assert n.returncode == 0, n.stderr
    assert n.stdout == r.stdout          # "N positions were excluded ..." (extract.c:1489)
    return refp, newp


@pytest.mark.parametrize("case", cases.REFERENCE_TESTS, ids=[c[0] for c in cases.REFERENCE_TESTS])
def test_cli_reference_testsuite(built, tmp_path, case):
    name, args, fa, bam, counts = case
    refp, newp = _both(built, tmp_path, name, args, cases.fx(fa), cases.fx(bam))
    for suffix, n in cases.counts_for(case).items():
        assert sum(1 for _ in open(newp + suffix)) == n
    assert compare_outputs(refp, newp) == []


@pytest.mark.parametrize("case", cases.FIXTURE_EXTRA, ids=[c[0] for c in cases.FIXTURE_EXTRA])
def test_cli_fixture_extra(built, tmp_path, case):
    name, args, fa, bam = case
    refp, newp = _both(built, tmp_path, name, args, cases.fx(fa), cases.fx(bam))
    assert compare_outputs(refp, newp) == []


@pytest.mark.parametrize("opts", cases.SYNTH_OPTION_SETS, ids=[cases.slug(o) for o in cases.SYNTH_OPTION_SETS])
def test_cli_synthetic_noisy(built, synth, tmp_path, opts):
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    refp, newp = _both(built, tmp_path, "s", opts, p + ".fa", p + ".bam")
    assert compare_outputs(refp, newp) == []


@pytest.mark.parametrize("opts", [[], ["--CHG", "--CHH", "--mergeContext"], ["--cytosine_report", "--CHG", "--CHH"],
                                  ["--minOppositeDepth", "3", "--maxVariantFrac", "0.2", "--CHG"]], ids=["default", "merge", "cyt", "variant"])
def test_cli_bismark_nondirectional(built, synth, tmp_path, opts):
    p = synth("bismark", "--contigs", "chrA:40000", "--depth", "40", "--bismark-tags", "--nondirectional", "0.3", "--single-frac", "0.1",
              "--isize-mean", "200", "--isize-sd", "30", "--read-seed", "99")
    refp, newp = _both(built, tmp_path, "b", opts, p + ".fa", p + ".bam")
    assert compare_outputs(refp, newp) == []


def test_cli_deep_overlap_panel(built, synth, tmp_path):
    p = synth("panel", "--contigs", "amp:3000", "--depth", "1500", "--isize-mean", "180", "--isize-sd", "25", "--isize-min", "150", "--isize-max", "300")
    refp, newp = _both(built, tmp_path, "p", ["--CHG", "--CHH"], p + ".fa", p + ".bam")
    assert compare_outputs(refp, newp) == []


def test_cli_multi_tile_1mbp(built, synth, tmp_path):
    """> 2 tiles of 2^19 alignments and > 1 reference chunk, with --mergeContext pairs straddling cuts"""
    p = synth("mb3", "--contigs", "chr1:3000000,chr2:400000", "--depth", "60")
    refp, newp = _both(built, tmp_path, "m", ["--CHG", "--CHH", "--mergeContext"], p + ".fa", p + ".bam")
    assert compare_outputs(refp, newp) == []


@pytest.mark.parametrize("opts", [["--noSVG"], ["--noSVG", "--CHG", "--CHH"], ["--noSVG", "--nOT", "3,3,3,3", "--chunkSize", "2500", "--CHG"],
                                  ["--noSVG", "-r", "chr1:1000-30000"]], ids=["cpg", "all", "trim_chunk", "region"])
def test_cli_mbias_txt(built, synth, opts):
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    r = subprocess.run([built["ref_bin"], "mbias"] + opts + [p + ".fa", p + ".bam"], capture_output=True, text=True)
    n = subprocess.run([NEW_BIN, "mbias"] + opts + [p + ".fa", p + ".bam"], capture_output=True, text=True)
    assert r.returncode == 0 and n.returncode == 0, (r.stderr, n.stderr)
    assert n.stdout == r.stdout and len(r.stdout.splitlines()) > 50


def test_cli_mbias_suggestions(built, synth, tmp_path):
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    r = subprocess.run([built["ref_bin"], "mbias", p + ".fa", p + ".bam", str(tmp_path / "r")], capture_output=True, text=True)
    n = subprocess.run([NEW_BIN, "mbias", p + ".fa", p + ".bam", str(tmp_path / "n")], capture_output=True, text=True)
    pick = lambda s: [l for l in s.splitlines() if l.startswith("Suggested inclusion options:")]
    assert pick(r.stderr) and pick(r.stderr) == pick(n.stderr)


# ---------------------------------------------------------------- tile level, through the C ABI
def _tile_compare(cfg, fa, bam, contig_idx=0, beg=0, end=None):
    b = api.BamFile(bam)
    name = b.names[contig_idx]
    ref = api.fetch_contig(fa, name)
    if end is None:
        end = len(ref)
    soa = b.read_region(contig_idx, beg, end)
    o = ob.lib()
    cap = (end - beg) + 16
    exp = (A.MdCall * cap)()
    est = A.MdTileStats()
    assert o.mdo_extract_tile(C.byref(cfg), ref, len(ref), beg, end, C.byref(soa), exp, cap, C.byref(est)) == 0
    with api.GpuContext(cfg) as g:
        g.load_contig(contig_idx, ref)
        got, gst = g.extract_tile(contig_idx, beg, end, soa)
    assert gst.n_calls == est.n_calls
    assert gst.n_admitted == est.n_admitted
    assert bytes(C.string_at(got, gst.n_calls * C.sizeof(A.MdCall))) == bytes(C.string_at(exp, est.n_calls * C.sizeof(A.MdCall)))
    b.close()
    return est


@pytest.mark.parametrize("kw", [dict(), dict(keepCHG=1, keepCHH=1), dict(keepCHG=1, keepCHH=1, minOppositeDepth=2, maxVariantFrac=0.1),
                                dict(minMapq=0, minPhred=1, ignoreFlags=0, keepDupes=1, keepSingleton=1, keepDiscordant=1, ignoreNH=1, keepCHH=1)],
                         ids=["cpg", "all", "variant", "nofilter"])
def test_tile_abi_vs_oracle(built, synth, kw):
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    st = _tile_compare(A.default_config(**kw), p + ".fa", p + ".bam", 0)
    assert st.n_calls > 1000 and st.n_pairs > 100


def test_tile_abi_partial_interval(built, synth):
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    _tile_compare(A.default_config(keepCHG=1, keepCHH=1), p + ".fa", p + ".bam", 0, 12345, 23456)
    _tile_compare(A.default_config(), p + ".fa", p + ".bam", 1, 0, 5)


def test_tile_abi_empty_and_tiny(built, synth):
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    cfg = A.default_config(minMapq=61)      # admits nothing
    st = _tile_compare(cfg, p + ".fa", p + ".bam", 0)
    assert st.n_calls == 0 and st.n_admitted == 0


def test_tile_abi_fixture_duplicate_names(built):
    """cg_aln.bam holds four records all named read1; with -F 0 the name occurs 4 times"""
    cfg = A.default_config(minMapq=2, ignoreFlags=0)
    _tile_compare(cfg, cases.fx("cg100.fa"), cases.fx("cg_aln.bam"))


def test_mbias_abi_vs_oracle(built, synth):
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    cfg = A.default_config(keepCHG=1, keepCHH=1, noOverlapMerge=1)
    b = api.BamFile(p + ".bam")
    ref = api.fetch_contig(p + ".fa", b.names[0])
    soa = b.read_region(0)
    h = A.load_host()
    bounds = (C.c_uint32 * 64)()
    n_chunks = h.mdh_chunk_bounds(ref, len(ref), 7000, 0, 0, bounds, 63)
    assert 8 <= n_chunks < 63
    o = ob.lib()
    ehist = (C.c_uint32 * (4 * 2 * A.MD_MBIAS_MAXLEN * 2))()
    elens = (C.c_int32 * 4)()
    assert o.mdo_mbias_tile(C.byref(cfg), ref, len(ref), 0, len(ref), bounds, n_chunks, C.byref(soa), ehist, elens, None) == 0
    with api.GpuContext(cfg) as g:
        g.load_contig(0, ref)
        g.set_mbias_chunks(0, [bounds[i] for i in range(n_chunks + 1)])
        g.mbias_tile(0, 0, len(ref), soa)
        ghist, glens = g.mbias_hist()
    assert glens == list(elens)
    assert bytes(ghist) == bytes(ehist)
    assert sum(ehist) > 10000


def test_device_resident_path_matches_host_path(built, synth):
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    cfg = A.default_config(keepCHG=1)
    b = api.BamFile(p + ".bam")
    ref = api.fetch_contig(p + ".fa", b.names[0])
    soa = b.read_region(0)
    with api.GpuContext(cfg) as g:
        g.load_contig(0, ref)
        got, st = g.extract_tile(0, 0, len(ref), soa)
        d = g.upload(soa)
        st2 = g.extract_tile_device(0, 0, len(ref), d)
        got2, n2 = g.fetch_calls(len(ref) + 16)
        g.free(d)
        assert g.launch_count() >= 8
    assert st.n_calls == st2.n_calls == n2
    assert bytes(C.string_at(got, n2 * 16)) == bytes(C.string_at(got2, n2 * 16))


# ---------------------------------------------------------------- read-length regimes of the streaming kernel
@pytest.mark.parametrize("name,args", [
    ("len250", ["--contigs", "chr1:80000", "--depth", "20", "--readlen", "250", "--isize-mean", "400", "--isize-sd", "80", "--isize-min", "250", "--isize-max", "900"]),
    ("len75", ["--contigs", "chr1:50000", "--depth", "25", "--readlen", "75", "--isize-mean", "160", "--isize-sd", "40", "--isize-min", "75", "--isize-max", "400"]),
    ("len600", ["--contigs", "chr1:90000", "--depth", "15", "--readlen", "600", "--isize-mean", "900", "--isize-sd", "150", "--isize-min", "600", "--isize-max", "2000"]),
    ("len20000", ["--contigs", "chr1:300000", "--depth", "6", "--readlen", "20000", "--isize-mean", "30000", "--isize-sd", "4000", "--isize-min", "20000", "--isize-max", "45000"]),
], ids=["250bp_batch19", "75bp", "600bp_batch7", "20kb_unstaged"])
def test_cli_read_lengths(built, synth, tmp_path, name, args):
    """250/600-mers shrink the per-warp batch (staging buffer holds fewer alignments), 20 kb reads exceed the queue's query-index
    field and take the whole-warp path from global memory"""
    p = synth(name, *args)
    for k, opts in enumerate([["--CHG", "--CHH"], ["--mergeContext", "--minOppositeDepth", "2", "--maxVariantFrac", "0.2"]]):
        refp, newp = _both(built, tmp_path, "%s_%d" % (name, k), opts, p + ".fa", p + ".bam")
        assert compare_outputs(refp, newp) == []


def test_cli_mbias_long_reads(built, synth):
    """query positions beyond the shared-memory histogram (>= 256) go to the global one; beyond MD_MBIAS_MAXLEN they are dropped by both sides... so stay below"""
    p = synth("len600", "--contigs", "chr1:90000", "--depth", "15", "--readlen", "600", "--isize-mean", "900", "--isize-sd", "150", "--isize-min", "600", "--isize-max", "2000")
    r = subprocess.run([built["ref_bin"], "mbias", "--noSVG", "--CHH", p + ".fa", p + ".bam"], capture_output=True, text=True)
    n = subprocess.run([NEW_BIN, "mbias", "--noSVG", "--CHH", p + ".fa", p + ".bam"], capture_output=True, text=True)
    assert r.returncode == 0 and n.returncode == 0, (r.stderr, n.stderr)
    assert n.stdout == r.stdout and len(r.stdout.splitlines()) > 1000


def test_cli_mbias_refuses_reads_longer_than_the_histogram(built, synth):
    """the reference grows its per-position arrays without bound (MBias.c:16-40); this build's histogram ends at
    MD_MBIAS_MAXLEN, and a longer read must be an error, not a silently shorter table"""
    p = synth("len1500", "--contigs", "chr1:120000", "--depth", "8", "--readlen", "1500", "--isize-mean", "2500", "--isize-sd", "200", "--isize-min", "1500", "--isize-max", "4000")
    for env in ({}, {"MD_DEVICE_DECODE": "0"}):
        n = subprocess.run([NEW_BIN, "mbias", "--noSVG", p + ".fa", p + ".bam"], capture_output=True, text=True, env=dict(os.environ, **env))
        assert n.returncode != 0 and "MD_MBIAS_MAXLEN" in n.stderr and n.stdout == "", (n.returncode, n.stderr[-300:])


# ---------------------------------------------------------------- phred encodings of the tile (md_reads_soa::qual_bits)
@pytest.mark.parametrize("nq,env", [(4, {}), (12, {}), (40, {}), (4, {"MD_QUAL_PACK": "0"})], ids=["2bit", "4bit", "8bit_alphabet_too_large", "packing_off"])
def test_cli_phred_encodings(built, synth, tmp_path, nq, env):
    """4 distinct phreds ship as 2-bit codes, 12 as 4-bit codes, 40 stay bytes; all three must give the reference's bytes"""
    p = synth("q%d" % nq, "--contigs", "chr1:50000", "--depth", "25", "--quals", str(nq))
    refp, newp = str(tmp_path / "ref"), str(tmp_path / "new")
    r = run_ref(built["ref_bin"], "extract", ["--CHG", "--CHH", "-p", "9"], p + ".fa", p + ".bam", refp)
    assert r.returncode == 0, r.stderr
    n = subprocess.run([NEW_BIN, "extract", "--CHG", "--CHH", "-p", "9", p + ".fa", p + ".bam", "-o", newp], capture_output=True, text=True, env=dict(os.environ, MD_DEVICE_DECODE="0", **env))   # host decoder: it is the one that packs phreds
    assert n.returncode == 0, n.stderr
    assert compare_outputs(refp, newp) == []


def test_tile_abi_reports_encoding(built, synth):
    p = synth("q12", "--contigs", "chr1:50000", "--depth", "25", "--quals", "12")
    b = api.BamFile(p + ".bam")
    soa = b.read_region(0)
    assert soa.qual_bits == 4 and sorted(soa.qual_lut[:12]) == list(range(2, 14))
    _tile_compare(A.default_config(keepCHG=1), p + ".fa", p + ".bam", 0)
    b.close()


def test_cli_ultra_deep_switches_to_wide_counters(built, synth, tmp_path):
    """> 65535 calls on one column: the packed 16-bit window counters wrap, the tile is counted again with 32-bit ones"""
    p = synth("ultradeep", "--contigs", "amp:700", "--depth", "260000", "--isize-mean", "170", "--isize-sd", "10", "--isize-min", "150", "--isize-max", "200", "--clean")
    refp, newp = _both(built, tmp_path, "ud", ["--CHG", "--CHH"], p + ".fa", p + ".bam")
    assert compare_outputs(refp, newp) == []
    top = max(int(l.split("\t")[4]) + int(l.split("\t")[5]) for l in open(refp + "_CHH.bedGraph").read().splitlines()[1:])
    assert top > 70000


@pytest.mark.parametrize("sub", ["extract", "mbias"])
@pytest.mark.parametrize("nq", [4, 12, 40], ids=["2bit", "4bit", "8bit"])
def test_cli_small_tiles_three_in_flight(built, synth, tmp_path, nq, sub):
    """tiles of 300 alignments through the asynchronous ring (3 lanes): carried reads re-encoded per tile, results of tiles
    in flight on different streams land in genome order (extract) / in one histogram (mbias)"""
    p = synth("st%d" % nq, "--contigs", "chr1:40000,chr2:9000", "--depth", "30", "--quals", str(nq), "--lower-frac", "0.02")
    env = dict(os.environ, MD_DECODE_JOB_BYTES="1", MD_TILE_READS="300", MD_DEVICE_DECODE="0")      # the host decoder's tile ring
    if sub == "extract":
        refp, newp = str(tmp_path / "ref"), str(tmp_path / "new")
        opts = ["--CHG", "--CHH", "--mergeContext"]
        r = run_ref(built["ref_bin"], "extract", opts, p + ".fa", p + ".bam", refp)
        n = subprocess.run([NEW_BIN, "extract"] + opts + ["-@", "3", p + ".fa", p + ".bam", "-o", newp], capture_output=True, text=True, env=env)
        assert r.returncode == 0 and n.returncode == 0, (r.stderr, n.stderr)
        assert compare_outputs(refp, newp) == []
    else:
        opts = ["--CHH", "--txt", "--noSVG"]
        r = subprocess.run([built["ref_bin"], "mbias"] + opts + [p + ".fa", p + ".bam"], capture_output=True, text=True)
        n = subprocess.run([NEW_BIN, "mbias"] + opts + ["-@", "3", p + ".fa", p + ".bam"], capture_output=True, text=True, env=env)
        assert r.returncode == 0 and n.returncode == 0, (r.stderr, n.stderr)
        assert n.stdout == r.stdout and len(r.stdout.splitlines()) > 10


@pytest.mark.parametrize("opts", cases.BED_OPTION_SETS, ids=[cases.slug(o) for o in cases.BED_OPTION_SETS])
@pytest.mark.parametrize("decode", ["device_decode", "host_decode"])
def test_cli_bed_regions(built, synth, tmp_path, opts, decode, monkeypatch):
    """-l / --keepStrand on the device (md_set_bed: admission in prep_kernel, column/strand masks in count_warp)"""
    monkeypatch.setenv("MD_DEVICE_DECODE", "1" if decode == "device_decode" else "0")
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    refp, newp = _both(built, tmp_path, "bed", cases.with_bed(opts, tmp_path), p + ".fa", p + ".bam")
    assert compare_outputs(refp, newp) == []


@pytest.mark.parametrize("opts", cases.BED_MBIAS_SETS, ids=["plain", "strand_allctx"])
def test_cli_bed_mbias(built, synth, tmp_path, opts):
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    o = cases.with_bed(opts, tmp_path)
    r = subprocess.run([built["ref_bin"], "mbias"] + o + [p + ".fa", p + ".bam"], capture_output=True, text=True)
    n = subprocess.run([NEW_BIN, "mbias"] + o + [p + ".fa", p + ".bam"], capture_output=True, text=True)
    assert r.returncode == 0 and n.returncode == 0, (r.stderr, n.stderr)
    assert n.stdout == r.stdout and len(r.stdout.splitlines()) > 50


def test_bed_tile_abi_vs_oracle(built, synth):
    """md_set_bed through the C ABI: one tile, strand-specific regions, variant filter on (six site bitmaps)"""
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    cfg = A.default_config(keepCHG=1, keepCHH=1, minOppositeDepth=2, maxVariantFrac=0.1)
    regs = [(1000, 3000, 1), (2500, 2600, 2), (10000, 10050, 2), (20000, 29999, 1), (20500, 40000, 0), (59990, 60001, 0)]
    arr = (A.MdBedRegion * len(regs))(*[A.MdBedRegion(*r) for r in regs])
    b = api.BamFile(p + ".bam")
    ref = api.fetch_contig(p + ".fa", "chr1")
    soa = b.read_region(0, 0, len(ref))
    o = ob.lib()
    cap = len(ref) + 16
    exp = (A.MdCall * cap)(); est = A.MdTileStats()
    o.mdo_set_bed(arr, len(regs), 1)
    try:
        assert o.mdo_extract_tile(C.byref(cfg), ref, len(ref), 0, len(ref), C.byref(soa), exp, cap, C.byref(est)) == 0
    finally:
        o.mdo_set_bed(None, 0, 0)
    with api.GpuContext(cfg) as g:
        g.load_contig(0, ref)
        assert g.g.md_set_bed(g.h, 0, arr, len(regs)) == 0
        got, gst = g.extract_tile(0, 0, len(ref), soa)
    assert gst.n_calls == est.n_calls and est.n_calls > 1000 and gst.n_admitted == est.n_admitted
    assert bytes(C.string_at(got, gst.n_calls * C.sizeof(A.MdCall))) == bytes(C.string_at(exp, est.n_calls * C.sizeof(A.MdCall)))
    b.close()


def _perread_cli(built, tmp_path, args, fa, bam):
    refp, newp = str(tmp_path / "pr_ref.txt"), str(tmp_path / "pr_new.txt")
    r = subprocess.run([built["ref_bin"], "perRead"] + list(args) + ["-o", refp, fa, bam], capture_output=True, text=True)
    n = subprocess.run([NEW_BIN, "perRead"] + list(args) + ["-o", newp, fa, bam], capture_output=True, text=True)
    assert r.returncode == 0 and n.returncode == 0, (r.stderr, n.stderr)
    return open(refp).read(), open(newp).read()


@pytest.mark.parametrize("opts", cases.PERREAD_SETS, ids=[cases.slug(o) for o in cases.PERREAD_SETS])
def test_cli_perread_synthetic(built, synth, tmp_path, opts):
    """perRead through the drop-in binary (per_read_kernel) against the reference build, byte for byte"""
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    a, b = _perread_cli(built, tmp_path, cases.with_bed(opts, tmp_path), p + ".fa", p + ".bam")
    assert a == b and len(a.splitlines()) > 500


@pytest.mark.parametrize("fx", cases.PERREAD_FIXTURES, ids=[f[1] for f in cases.PERREAD_FIXTURES])
def test_cli_perread_fixtures(built, tmp_path, fx):
    fa, bam, args = fx
    a, b = _perread_cli(built, tmp_path, args, cases.fx(fa), cases.fx(bam))
    assert a == b and a


def test_cli_perread_indels_clips_long_reads(built, synth, tmp_path):
    p = synth("len75", "--contigs", "chr1:50000", "--depth", "25", "--readlen", "75", "--isize-mean", "160", "--isize-sd", "40", "--isize-min", "75", "--isize-max", "400")
    a, b = _perread_cli(built, tmp_path, ["-p", "12", "-q", "0"], p + ".fa", p + ".bam")
    assert a == b and len(a.splitlines()) > 500


def test_perread_tile_abi_vs_oracle(built, synth):
    """md_per_read_tile through the C ABI against the port, record for record (packed phred tiles included)"""
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    cfg = A.default_config(minMapq=0, minPhred=12)
    cfg.ignoreFlags = 0
    b = api.BamFile(p + ".bam")
    ref = api.fetch_contig(p + ".fa", "chr1")
    soa = b.read_region(0, 0, len(ref))
    n = soa.n_reads
    exp = (A.MdReadMeth * n)(); got = (A.MdReadMeth * n)()
    o = ob.lib()
    for beg, end, chunk in [(0, len(ref), 1000000), (5000, 41234, 999), (0, len(ref), 1)]:
        assert o.mdo_per_read_tile(C.byref(cfg), ref, len(ref), beg, end, chunk, C.byref(soa), exp) == 0
        with api.GpuContext(cfg) as g:
            g.load_contig(0, ref)
            td = A.MdTileDesc(0, beg, end)
            assert g.g.md_per_read_tile(g.h, C.byref(td), C.byref(soa), chunk, got) == 0
        assert bytes(got) == bytes(exp)
        assert sum(1 for k in range(n) if exp[k].nmeth != 0xffffffff and exp[k].nmeth + exp[k].nunmeth > 0) > 1000
    b.close()


@pytest.mark.parametrize("opts", [["--noSVG", "--minConversionEfficiency", "0.97", "--CHH"], ["--noSVG", "--minConversionEfficiency", "0.9", "--CHG", "--nOT", "3,3,3,3"]],
                         ids=["ce097_chh", "ce09_chg_trim"])
def test_cli_mbias_conversion_efficiency(built, synth, opts):
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    r = subprocess.run([built["ref_bin"], "mbias"] + opts + [p + ".fa", p + ".bam"], capture_output=True, text=True)
    n = subprocess.run([NEW_BIN, "mbias"] + opts + [p + ".fa", p + ".bam"], capture_output=True, text=True)
    assert r.returncode == 0 and n.returncode == 0, (r.stderr, n.stderr)
    assert n.stdout == r.stdout and len(r.stdout.splitlines()) > 50
