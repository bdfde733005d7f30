"""CPU tier for the device-side BAM decode: the sub-command drivers run in device-decode mode against
tests/native/libmdemu.so, which executes the CUDA kernels' own per-thread bodies (inflate_hd.h, bamrec_hd.h, bamdev_hd.h)
in plain loops and hands every assembled tile to the oracle port.  Output must equal oracle/_ref byte for byte."""
import gzip
import os
import struct
import subprocess
import sys
import zlib

import pytest

import cases
from util import run_ref, compare_outputs


def _host_main(sub, argv, env):
    overlapped = env.pop("OVERLAPPED", "1") == "1"           # two-phase push (segment k+1 decoded while segment k's tiles are built) or the plain one
    staging = env.pop("STAGING", "0") == "1"                 # page-locked staging buffers (stand-ins that are poisoned on release) + md_bam_prefetch
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import oracle_binding as ob; "
            "b = ob.OracleBackend(device_decode=True, overlapped=%r, staging=%r); rc = ob.run_host_main(%r, %r, b); "
            "sys.stderr.write('\\nPREFETCH_USED %%d\\n' %% b.state.get('prefetch_used', 0)); sys.exit(rc)") % (
                cases.ROOT, os.path.join(cases.ROOT, "tests"), overlapped, staging, sub, argv)
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, MD_DEVICE_DECODE="1", **env))


def _prefetch_used(proc):
    return int([ln for ln in proc.stderr.splitlines() if ln.startswith("PREFETCH_USED")][-1].split()[1])


def _extract_both(built, tmp_path, opts, fa, bam, env):
    refp, newp = str(tmp_path / "ref"), str(tmp_path / "new")
    r = run_ref(built["ref_bin"], "extract", opts, fa, bam, refp)
    assert r.returncode == 0, r.stderr
    n = _host_main("extract", list(opts) + [fa, bam, "-o", newp], env)
    assert n.returncode == 0, n.stderr
    assert n.stdout == r.stdout
    return compare_outputs(refp, newp)


@pytest.mark.parametrize("seg", ["1", "70000", "100000000"], ids=["block_per_segment", "70kB_segments", "one_segment"])
@pytest.mark.parametrize("opts", [["--CHG", "--CHH", "--mergeContext"], ["--cytosine_report", "--CHH"], ["-r", "chr1:5000-20000"], ["-r", "chr2", "--methylKit"],
                                  ["--minOppositeDepth", "3", "--maxVariantFrac", "0.2", "--CHG"]], ids=["merge", "cytosine_report", "region", "contig2", "variants"])
def test_extract_noisy(built, synth, tmp_path, opts, seg):
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    assert _extract_both(built, tmp_path, opts, p + ".fa", p + ".bam", {"MD_SEGMENT_BYTES": seg}) == []


@pytest.mark.parametrize("seg", ["1", "30000", "70000"], ids=["block_per_segment", "30kB_segments", "70kB_segments"])
@pytest.mark.parametrize("opts", [["--CHG", "--CHH", "--mergeContext"], ["-r", "chr1:5000-20000"], ["-r", "chr2", "--methylKit"]], ids=["merge", "region", "contig2"])
def test_extract_through_staging_buffers_with_prefetch(built, synth, tmp_path, opts, seg):
    """the round-2 driver on the CPU: segments are read into recycled staging buffers by the background reader, three host
    segments are alive at a time, the bytes of segment k+2 are prefetched while k+1 is decoded.  The emulation reads the caller's
    buffers as late as the device may, its "page-locked" memory is poisoned on release, and it refuses a prefetched buffer whose
    bytes changed before the push that uses it ended — so a buffer recycled too early shows as an error or as different output."""
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    refp, newp = str(tmp_path / "ref"), str(tmp_path / "new")
    r = run_ref(built["ref_bin"], "extract", opts, p + ".fa", p + ".bam", refp)
    assert r.returncode == 0, r.stderr
    n = _host_main("extract", list(opts) + [p + ".fa", p + ".bam", "-o", newp], {"MD_SEGMENT_BYTES": seg, "MD_STAGE": "1", "STAGING": "1"})
    assert n.returncode == 0, n.stderr
    assert compare_outputs(refp, newp) == []
    if opts[0] != "-r":
        assert _prefetch_used(n) >= 3                        # whole file: every segment from the third on arrives by prefetch


def test_mbias_through_staging_buffers_with_prefetch(built, synth, tmp_path):
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    r = subprocess.run([built["ref_bin"], "mbias", "--noSVG", "--CHH", p + ".fa", p + ".bam"], capture_output=True, text=True)
    n = _host_main("mbias", ["--noSVG", "--CHH", p + ".fa", p + ".bam"], {"MD_SEGMENT_BYTES": "40000", "MD_STAGE": "1", "STAGING": "1"})
    assert n.returncode == 0, n.stderr
    assert n.stdout == r.stdout and len(r.stdout) > 500
    assert _prefetch_used(n) >= 3


@pytest.mark.parametrize("nth", ["1", "3"])
def test_a_failing_prefetch_ends_the_run_cleanly(built, synth, tmp_path, nth):
    """the first / a later md_bam_prefetch fails while a push is in flight: the driver reports the device error and returns
    (it must wait for the push before it lets go of the staging buffers — the emulation reads them in push_end)"""
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    n = _host_main("extract", [p + ".fa", p + ".bam", "-o", str(tmp_path / "x")], {"MD_SEGMENT_BYTES": "30000", "MD_STAGE": "1", "STAGING": "1", "MDEMU_FAIL_PREFETCH": nth})
    assert n.returncode != 0 and "device error" in n.stderr, (n.returncode, n.stderr[-500:])


def test_plain_push(built, synth, tmp_path):
    """back ends without the two-phase push are driven segment by segment"""
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    assert _extract_both(built, tmp_path, ["--CHG", "--mergeContext"], p + ".fa", p + ".bam", {"MD_SEGMENT_BYTES": "70000", "OVERLAPPED": "0"}) == []


def test_wrong_guesses_are_repaired(built, synth, tmp_path):
    """every third block's guessed record start is sabotaged: the serial repair pass must restore the true chain"""
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    assert _extract_both(built, tmp_path, ["--CHG"], p + ".fa", p + ".bam", {"MD_SEGMENT_BYTES": "300000", "MDEMU_FORCE_FIX": "1"}) == []


def test_bismark_tags_and_single_end(built, synth, tmp_path):
    p = synth("bismark", "--contigs", "chrA:40000", "--depth", "40", "--bismark-tags", "--nondirectional", "0.3", "--single-frac", "0.1",
              "--isize-mean", "200", "--isize-sd", "30", "--read-seed", "99")
    assert _extract_both(built, tmp_path, ["--CHG", "--CHH"], p + ".fa", p + ".bam", {"MD_SEGMENT_BYTES": "50000"}) == []


def test_records_larger_than_a_block(built, synth, tmp_path):
    """20 kb reads: records span several BGZF blocks, so blocks without any record start exist"""
    p = synth("len20000", "--contigs", "chr1:300000", "--depth", "6", "--readlen", "20000", "--isize-mean", "30000", "--isize-sd", "4000", "--isize-min", "20000", "--isize-max", "45000")
    assert _extract_both(built, tmp_path, ["--CHG", "--CHH"], p + ".fa", p + ".bam", {"MD_SEGMENT_BYTES": "1"}) == []
    assert _extract_both(built, tmp_path, ["--CHG", "--CHH"], p + ".fa", p + ".bam", {"MD_SEGMENT_BYTES": "150000"}) == []


@pytest.mark.parametrize("case", cases.REFERENCE_TESTS, ids=[c[0] for c in cases.REFERENCE_TESTS])
def test_reference_fixtures(built, tmp_path, case):
    name, args, fa, bam = case[0], case[1], case[2], case[3]
    if "--minConversionEfficiency" in args:
        pytest.skip("per-chunk conversion-efficiency windows use the host decode path")
    assert _extract_both(built, tmp_path, list(args), cases.fx(fa), cases.fx(bam), {"MD_SEGMENT_BYTES": "1"}) == []


@pytest.mark.parametrize("seg", ["1", "90000"])
def test_mbias(built, synth, seg):
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    opts = ["--noSVG", "--CHG", "--CHH", "--nOT", "3,3,3,3", "--chunkSize", "2500"]
    r = subprocess.run([built["ref_bin"], "mbias"] + opts + [p + ".fa", p + ".bam"], capture_output=True, text=True)
    n = _host_main("mbias", opts + [p + ".fa", p + ".bam"], {"MD_SEGMENT_BYTES": seg})
    assert r.returncode == 0 and n.returncode == 0, (r.stderr, n.stderr)
    assert n.stdout == r.stdout and len(r.stdout.splitlines()) > 50


def _rewrite_bgzf(src, dst, level, strategy, block=0xff00):
    """same BAM bytes, different deflate flavour: level 0 = stored blocks, Z_FIXED = fixed Huffman, 9 = long matches"""
    raw = gzip.open(src, "rb").read()
    with open(dst, "wb") as f:
        for off in list(range(0, len(raw), block)) + [None]:
            chunk = b"" if off is None else raw[off:off + block]
            co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
            body = co.compress(chunk) + co.flush()
            bsize = len(body) + 25
            f.write(struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize))
            f.write(body)
            f.write(struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))


@pytest.mark.parametrize("level,strategy,block", [(0, zlib.Z_DEFAULT_STRATEGY, 0xff00), (6, zlib.Z_FIXED, 0xff00), (9, zlib.Z_DEFAULT_STRATEGY, 0xff00),
                                                  (6, zlib.Z_HUFFMAN_ONLY, 3000), (1, zlib.Z_RLE, 777)], ids=["stored", "fixed", "level9", "huffman_only_small", "rle_tiny"])
def test_deflate_flavours(built, synth, tmp_path, level, strategy, block):
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    bam = str(tmp_path / "re.bam")
    _rewrite_bgzf(p + ".bam", bam, level, strategy, block)
    # the reference runs on the original file (same records; the shim cannot index the rewritten one), ours on the rewritten one, unindexed
    refp, newp = str(tmp_path / "ref"), str(tmp_path / "new")
    opts = ["--CHG", "--mergeContext"]
    r = run_ref(built["ref_bin"], "extract", opts, p + ".fa", p + ".bam", refp)
    n = _host_main("extract", opts + [p + ".fa", bam, "-o", newp], {"MD_SEGMENT_BYTES": "200000"})
    assert r.returncode == 0 and n.returncode == 0, (r.stderr, n.stderr)
    assert compare_outputs(refp, newp) == []
