"""GPU tier for the device-side BAM decode: (1) the tile the kernels assemble in HBM equals, array for array, the tile the
host decoder builds from the same records; (2) `lib/MethylDackel` in device-decode mode writes the reference's bytes."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

import cases
from methyldackel_b200 import _abi as A
from methyldackel_b200 import api
from util import run_ref, compare_outputs
from test_device_decode_emulated import _rewrite_bgzf

pytestmark = pytest.mark.gpu
NEW_BIN = os.path.join(cases.ROOT, "methyldackel_b200", "lib", "MethylDackel")


def _block_table(raw):
    """BGZF block headers -> (md_bgzf_block array, offset of the first block's payload)"""
    blocks, off = [], 0
    while off + 18 <= len(raw):
        xlen = struct.unpack_from("<H", raw, off + 10)[0]
        bsize, o = None, 0
        while o + 4 <= xlen:
            si1, si2, slen = raw[off + 12 + o], raw[off + 13 + o], struct.unpack_from("<H", raw, off + 14 + o)[0]
            if si1 == 66 and si2 == 67 and slen == 2:
                bsize = struct.unpack_from("<H", raw, off + 16 + o)[0]
            o += 4 + slen
        bs = bsize + 1
        isize = struct.unpack_from("<I", raw, off + bs - 4)[0]
        blocks.append((off + 12 + xlen, bs - 12 - xlen - 8, isize))
        off += bs
    arr = (A.MdBgzfBlock * len(blocks))()
    for k, (a, b, c) in enumerate(blocks):
        arr[k].comp_off, arr[k].comp_len, arr[k].isize = a, b, c
    return arr


def _header_bytes(raw):
    """length of the BAM header in the inflated stream (magic, text, reference table)"""
    import gzip, io
    u = gzip.GzipFile(fileobj=io.BytesIO(raw)).read(1 << 22)
    l_text = struct.unpack_from("<i", u, 4)[0]
    off = 8 + l_text
    n_ref = struct.unpack_from("<i", u, off)[0]; off += 4
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", u, off)[0]; off += 4 + l_name + 4
    return off, n_ref


def _np(ptr, n, dt):
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(dt)), shape=(max(n, 1),))[:n].copy()


@pytest.mark.parametrize("name,args", [
    ("noisy", ["--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01"]),
    ("bismark", ["--contigs", "chrA:40000", "--depth", "40", "--bismark-tags", "--nondirectional", "0.3", "--single-frac", "0.1", "--isize-mean", "200", "--isize-sd", "30", "--read-seed", "99"]),
    ("len20000", ["--contigs", "chr1:300000", "--depth", "6", "--readlen", "20000", "--isize-mean", "30000", "--isize-sd", "4000", "--isize-min", "20000", "--isize-max", "45000"]),
], ids=["noisy", "bismark_aux", "20kb_reads"])
def test_device_tile_equals_host_tile(built, synth, name, args, monkeypatch):
    monkeypatch.setenv("MD_QUAL_PACK", "0")                  # the device tile keeps plain phred bytes
    p = synth(name, *args)
    raw = open(p + ".bam", "rb").read()
    blocks = _block_table(raw)
    skip, n_ref = _header_bytes(raw)
    g = A.load_gpu()
    b = api.BamFile(p + ".bam")
    with api.GpuContext(A.default_config(keepCHG=1, keepCHH=1)) as ctx:
        s = g.md_bam_open(ctx.h, n_ref)
        assert s, g.md_last_error()
        summ = A.MdBamSummary()
        buf = C.create_string_buffer(raw, len(raw))
        assert g.md_bam_push(s, buf, len(raw), blocks, len(blocks), skip, C.byref(summ)) == 0, g.md_last_error()
        runs = (A.MdBamRun * summ.n_runs)()
        assert g.md_bam_get_runs(s, runs, summ.n_runs) == summ.n_runs
        assert summ.leftover_bytes == 0 and summ.n_records > 50
        seen = 0
        for k in range(summ.n_runs):
            tid = runs[k].tid
            if tid < 0:
                continue
            ref = api.fetch_contig(p + ".fa", b.names[tid])
            ctx.load_contig(tid, ref)
            td = A.MdTileDesc(tid, 0, len(ref), 0, 0)
            cap = len(ref) + 16
            calls = (A.MdCall * cap)(); st = A.MdTileStats()
            assert g.md_bam_extract_run(s, k, C.byref(td), 0xffffffff, calls, cap, C.byref(st)) == 0, g.md_last_error()
            shape = A.MdReadsSoa()
            assert g.md_bam_tile_shape(s, C.byref(shape)) == 0
            host = b.read_region(tid, 0, len(ref))
            assert shape.n_reads == host.n_reads == runs[k].n
            assert (shape.n_cigar_ops, shape.seq_words, shape.qual_words) == (host.n_cigar_ops, host.seq_words, host.qual_words)
            n = shape.n_reads
            arrs = {"pos": (np.int32, n), "flag": (np.uint16, n), "mapq": (np.uint8, n), "aux": (np.uint8, n), "l_qseq": (np.uint32, n), "cigar_off": (np.uint32, n + 1),
                    "seq_off": (np.uint32, n), "qual_off": (np.uint32, n), "frag_key": (np.uint64, n), "cigar": (np.uint32, shape.n_cigar_ops), "seq": (np.uint32, shape.seq_words),
                    "qual": (np.uint64, shape.qual_words)}
            keep = {}
            for f, (dt, cnt) in arrs.items():
                keep[f] = np.zeros(max(cnt, 1), dtype=dt)
                setattr(shape, f, keep[f].ctypes.data_as(type(getattr(shape, f))))
            rend = np.zeros(max(n, 1), dtype=np.int32)
            assert g.md_bam_tile_fetch(s, C.byref(shape), rend.ctypes.data_as(C.POINTER(C.c_int32))) == 0, g.md_last_error()
            ct = {np.int32: C.c_int32, np.uint16: C.c_uint16, np.uint8: C.c_uint8, np.uint32: C.c_uint32, np.uint64: C.c_uint64}
            for f, (dt, cnt) in arrs.items():
                h = _np(getattr(host, f), cnt, ct[dt])
                assert np.array_equal(keep[f][:cnt], h), "column %s differs on contig %d" % (f, tid)
            # and the calls equal the host-tile path's
            got2, st2 = ctx.extract_tile(tid, 0, len(ref), host)
            assert st2.n_calls == st.n_calls
            assert bytes(C.string_at(got2, st2.n_calls * C.sizeof(A.MdCall))) == bytes(C.string_at(calls, st.n_calls * C.sizeof(A.MdCall)))
            ctx.g.md_drop_contig(ctx.h, tid)
            seen += 1
        assert seen >= 1
        g.md_bam_close(s)
    b.close()


def _cli_both(built, tmp_path, sub_opts, fa, bam, env):
    refp, newp = str(tmp_path / "ref"), str(tmp_path / "new")
    r = run_ref(built["ref_bin"], "extract", sub_opts, fa, bam, refp)
    n = subprocess.run([NEW_BIN, "extract"] + list(sub_opts) + [fa, bam, "-o", newp], capture_output=True, text=True, env=dict(os.environ, MD_DEVICE_DECODE="1", **env))
    assert r.returncode == 0 and n.returncode == 0, (r.stderr, n.stderr)
    assert n.stdout == r.stdout
    return compare_outputs(refp, newp)


@pytest.mark.parametrize("seg", ["1", "70000", "100000000"], ids=["block_per_segment", "70kB_segments", "one_segment"])
@pytest.mark.parametrize("opts", [["--CHG", "--CHH", "--mergeContext"], ["--cytosine_report", "--CHH"], ["-r", "chr1:5000-20000"], ["-r", "chr2", "--methylKit"],
                                  ["--minOppositeDepth", "3", "--maxVariantFrac", "0.2", "--CHG"]], ids=["merge", "cytosine_report", "region", "contig2", "variants"])
def test_cli_device_decode_noisy(built, synth, tmp_path, opts, seg):
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    assert _cli_both(built, tmp_path, opts, p + ".fa", p + ".bam", {"MD_SEGMENT_BYTES": seg}) == []


@pytest.mark.parametrize("case", cases.REFERENCE_TESTS, ids=[c[0] for c in cases.REFERENCE_TESTS])
def test_cli_device_decode_reference_fixtures(built, tmp_path, case):
    name, args, fa, bam = case[0], case[1], case[2], case[3]
    assert _cli_both(built, tmp_path, list(args), cases.fx(fa), cases.fx(bam), {"MD_SEGMENT_BYTES": "1"}) == []


def test_cli_device_decode_long_reads_and_bismark(built, synth, tmp_path):
    p = synth("len20000", "--contigs", "chr1:300000", "--depth", "6", "--readlen", "20000", "--isize-mean", "30000", "--isize-sd", "4000", "--isize-min", "20000", "--isize-max", "45000")
    assert _cli_both(built, tmp_path, ["--CHG", "--CHH"], p + ".fa", p + ".bam", {"MD_SEGMENT_BYTES": "150000"}) == []
    p = synth("bismark", "--contigs", "chrA:40000", "--depth", "40", "--bismark-tags", "--nondirectional", "0.3", "--single-frac", "0.1", "--isize-mean", "200", "--isize-sd", "30", "--read-seed", "99")
    assert _cli_both(built, tmp_path, ["--CHG", "--CHH"], p + ".fa", p + ".bam", {"MD_SEGMENT_BYTES": "50000"}) == []


def test_cli_device_decode_multi_segment_3mbp(built, synth, tmp_path):
    """several 8 MB segments per contig: tiles cut at the last record of a segment, reads carried on the device"""
    p = synth("mb3", "--contigs", "chr1:3000000,chr2:400000", "--depth", "60")
    assert _cli_both(built, tmp_path, ["--CHG", "--CHH", "--mergeContext"], p + ".fa", p + ".bam", {"MD_SEGMENT_BYTES": str(8 << 20)}) == []


@pytest.mark.parametrize("level,strategy,block", [(0, 0, 0xff00), (6, 4, 0xff00), (9, 0, 0xff00), (1, 3, 777)], ids=["stored", "fixed", "level9", "rle_tiny"])
def test_cli_device_decode_deflate_flavours(built, synth, tmp_path, level, strategy, block):
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    bam = str(tmp_path / "re.bam")
    _rewrite_bgzf(p + ".bam", bam, level, strategy, block)
    refp, newp = str(tmp_path / "ref"), str(tmp_path / "new")
    opts = ["--CHG", "--mergeContext"]
    r = run_ref(built["ref_bin"], "extract", opts, p + ".fa", p + ".bam", refp)
    n = subprocess.run([NEW_BIN, "extract"] + opts + [p + ".fa", bam, "-o", newp], capture_output=True, text=True, env=dict(os.environ, MD_DEVICE_DECODE="1", MD_SEGMENT_BYTES="200000"))
    assert r.returncode == 0 and n.returncode == 0, (r.stderr, n.stderr)
    assert compare_outputs(refp, newp) == []


@pytest.mark.parametrize("seg", ["1", "90000"])
def test_cli_device_decode_mbias(built, synth, seg):
    p = synth("noisy", "--contigs", "chr1:60000,chr2:15000", "--depth", "25", "--lower-frac", "0.02", "--n-frac", "0.01")
    opts = ["--noSVG", "--CHG", "--CHH", "--nOT", "3,3,3,3", "--chunkSize", "2500"]
    r = subprocess.run([built["ref_bin"], "mbias"] + opts + [p + ".fa", p + ".bam"], capture_output=True, text=True)
    n = subprocess.run([NEW_BIN, "mbias"] + opts + [p + ".fa", p + ".bam"], capture_output=True, text=True, env=dict(os.environ, MD_DEVICE_DECODE="1", MD_SEGMENT_BYTES=seg))
    assert r.returncode == 0 and n.returncode == 0, (r.stderr, n.stderr)
    assert n.stdout == r.stdout and len(r.stdout.splitlines()) > 50


@pytest.mark.parametrize("seg", ["1", "70000", "3000000"], ids=["block_per_segment", "70kB_segments", "3MB_segments"])
def test_cli_device_decode_through_staging_buffers(built, synth, tmp_path, seg):
    """large files reach the device through page-locked staging buffers filled by a background reader (StagedSegments, forced
    here with MD_STAGE=1): same bytes out, whatever the segment size, with regions (index seek) and several contigs"""
    p = synth("stage", "--contigs", "chr1:400000,chr2:150000,chr3:30000", "--depth", "25", "--read-seed", "12")
    for opts in (["--CHG", "--CHH", "--mergeContext"], ["-r", "chr2:20000-90000", "--CHG"], ["-r", "chr3"]):
        assert _cli_both(built, tmp_path, opts, p + ".fa", p + ".bam", {"MD_SEGMENT_BYTES": seg, "MD_STAGE": "1"}) == []
    r = subprocess.run([built["ref_bin"], "mbias", "--noSVG", "--CHG", p + ".fa", p + ".bam"], capture_output=True, text=True)
    n = subprocess.run([NEW_BIN, "mbias", "--noSVG", "--CHG", p + ".fa", p + ".bam"], capture_output=True, text=True, env=dict(os.environ, MD_SEGMENT_BYTES=seg, MD_STAGE="1"))
    assert r.returncode == 0 and n.returncode == 0 and n.stdout == r.stdout and len(r.stdout) > 1000
