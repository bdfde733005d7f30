"""GPU tier, BASELINE.json configs[1] at FULL size (10 Mbp, 30x PE150, ~2.0 M alignments): size-independent properties of the
hot path, plus a bit-exact comparison with the reference build on a region of the same file.
  * tiling invariance: one tile over the whole contig, 2^17-alignment tiles in flight over three lanes, and the drop-in
    binary reading the compressed BAM through the device decoder must report byte-identical md_call records / lines;
  * conservation: the sum of nmeth + nunmeth over all calls equals the number of per-base calls an independent recount of
    a sample of columns gives (the oracle port on a 200 kbp window of the same tile);
  * the reference build's bedGraph for chr1:4,000,001-4,200,000 equals the binary's for the same -r."""
import ctypes as C
import hashlib
import os
import subprocess

import pytest

import cases
import oracle_binding as ob
from methyldackel_b200 import _abi as A
from methyldackel_b200 import api

pytestmark = pytest.mark.gpu
NEW_BIN = os.path.join(cases.ROOT, "methyldackel_b200", "lib", "MethylDackel")


@pytest.fixture(scope="module")
def c2(built, synth):
    return synth("c2_full", "--contigs", "chr1:10000000", "--depth", "30", "--read-seed", "5678")


def test_config1_tiling_invariance_and_conservation(built, c2, tmp_path):
    cfg = A.default_config()
    b = api.BamFile(c2 + ".bam")
    ref = api.fetch_contig(c2 + ".fa", "chr1")
    soa = b.read_region(0)
    assert soa.n_reads > 1900000
    cap = len(ref) + 16
    with api.GpuContext(cfg) as g:
        g.load_contig(0, ref)
        whole, st = g.extract_tile(0, 0, len(ref), soa)
        whole_bytes = bytes(C.string_at(whole, st.n_calls * 16))
        # the same region as 2^17-alignment tiles, three in flight
        tiles = b.make_tiles(0, 0, len(ref), 1 << 17)
        assert len(tiles) >= 12
        calls = (A.MdCall * cap)(); stt = A.MdTileStats()
        out_off, inflight = 0, []
        for td_k, soa_k in tiles:
            if len(inflight) == 3:
                dst = C.cast(C.addressof(calls) + out_off * 16, C.POINTER(A.MdCall))
                assert g.g.md_collect_tile(g.h, inflight.pop(0), dst, cap - out_off, C.byref(stt)) == 0, g.g.md_last_error()
                out_off += stt.n_calls
            t_id = g.g.md_submit_tile(g.h, C.byref(td_k), C.byref(soa_k))
            assert t_id >= 0, g.g.md_last_error()
            inflight.append(t_id)
        for t_id in inflight:
            dst = C.cast(C.addressof(calls) + out_off * 16, C.POINTER(A.MdCall))
            assert g.g.md_collect_tile(g.h, t_id, dst, cap - out_off, C.byref(stt)) == 0, g.g.md_last_error()
            out_off += stt.n_calls
        assert out_off == st.n_calls > 1000000
        assert hashlib.sha256(bytes(C.string_at(calls, out_off * 16))).digest() == hashlib.sha256(whole_bytes).digest()
    # conservation against an independent recount: the oracle port on a 200 kbp window of the same alignments
    beg, end = 4000000, 4200000
    win = b.read_region(0, beg, end)
    exp = (A.MdCall * (end - beg + 16))(); est = A.MdTileStats()
    assert ob.lib().mdo_extract_tile(C.byref(cfg), ref, len(ref), beg, end, C.byref(win), exp, end - beg + 16, C.byref(est)) == 0
    got = [(whole[k].pos, whole[k].nmeth, whole[k].nunmeth, whole[k].info) for k in range(st.n_calls) if beg <= whole[k].pos < end]
    want = [(exp[k].pos, exp[k].nmeth, exp[k].nunmeth, exp[k].info) for k in range(est.n_calls)]
    assert got == want and sum(x[1] + x[2] for x in want) > 100000
    # the drop-in binary (device decoder, its own tiling) writes exactly these calls
    pre = str(tmp_path / "cli")
    r = subprocess.run([NEW_BIN, "extract", c2 + ".fa", c2 + ".bam", "-o", pre], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = open(pre + "_CpG.bedGraph").read().splitlines()[1:]
    assert len(lines) == st.n_calls
    for k in (0, 1, st.n_calls // 2, st.n_calls - 1):
        f = lines[k].split("\t")
        assert (int(f[1]), int(f[4]), int(f[5])) == (whole[k].pos, whole[k].nmeth, whole[k].nunmeth)
    assert sum(int(l.rsplit("\t", 2)[1]) + int(l.rsplit("\t", 2)[2]) for l in lines) == sum(whole[k].nmeth + whole[k].nunmeth for k in range(st.n_calls))
    b.close()


def test_config1_region_equals_reference_build(built, c2, tmp_path):
    reg = "chr1:4000001-4200000"
    rp, np_ = str(tmp_path / "ref"), str(tmp_path / "new")
    r = subprocess.run([built["ref_bin"], "extract", "-r", reg, c2 + ".fa", c2 + ".bam", "-o", rp], capture_output=True, text=True)
    n = subprocess.run([NEW_BIN, "extract", "-r", reg, c2 + ".fa", c2 + ".bam", "-o", np_], capture_output=True, text=True)
    assert r.returncode == 0 and n.returncode == 0, (r.stderr, n.stderr)
    a = open(rp + "_CpG.bedGraph").read().replace(rp, "X"); bb = open(np_ + "_CpG.bedGraph").read().replace(np_, "X")
    assert a == bb and a.count("\n") > 20000


def _cmp_extract(built, tmp_path, prefix, opts, tag):
    rp, np_ = str(tmp_path / (tag + "_ref")), str(tmp_path / (tag + "_new"))
    r = subprocess.run([built["ref_bin"], "extract", "-@", str(os.cpu_count() or 1)] + opts + [prefix + ".fa", prefix + ".bam", "-o", rp], capture_output=True, text=True)
    n = subprocess.run([NEW_BIN, "extract"] + opts + [prefix + ".fa", prefix + ".bam", "-o", np_], capture_output=True, text=True)
    assert r.returncode == 0 and n.returncode == 0, (r.stderr[-500:], n.stderr[-500:])
    assert r.stdout == n.stdout
    total = 0
    for ctx in ("CpG", "CHG", "CHH"):
        fr, fn = "%s_%s.bedGraph" % (rp, ctx), "%s_%s.bedGraph" % (np_, ctx)
        assert os.path.exists(fr) == os.path.exists(fn)
        if os.path.exists(fr):
            a = open(fr, "rb").read().replace(rp.encode(), b"X"); bb = open(fn, "rb").read().replace(np_.encode(), b"X")
            assert a == bb, "%s %s differs" % (tag, ctx)
            total += len(a)
    return total


def test_config1_whole_file_equals_reference_build(built, c2, tmp_path):
    """the binary's WHOLE config[1] output against oracle/_ref's, byte for byte (not a hash against itself): the
    default CpG bedGraph, the merged all-context files of config[2]'s option set, and the mbias table + suggestion line"""
    assert _cmp_extract(built, tmp_path, c2, [], "cpg") > 25_000_000
    assert _cmp_extract(built, tmp_path, c2, ["--CHG", "--CHH", "--mergeContext"], "all") > 60_000_000
    r = subprocess.run([built["ref_bin"], "mbias", "-@", str(os.cpu_count() or 1), "--txt", c2 + ".fa", c2 + ".bam", str(tmp_path / "mb_ref")], capture_output=True, text=True)
    n = subprocess.run([NEW_BIN, "mbias", "--txt", c2 + ".fa", c2 + ".bam", str(tmp_path / "mb_new")], capture_output=True, text=True)
    assert r.returncode == 0 and n.returncode == 0, (r.stderr[-500:], n.stderr[-500:])
    assert r.stdout == n.stdout and len(r.stdout) > 5000
    assert [l for l in r.stderr.splitlines() if l.startswith("Suggested")] == [l for l in n.stderr.splitlines() if l.startswith("Suggested")]
    for s_ in ("OT", "OB"):
        assert open(str(tmp_path / ("mb_ref_%s.svg" % s_))).read() == open(str(tmp_path / ("mb_new_%s.svg" % s_))).read()


def test_config3_panel_fraction_equals_reference_build(built, synth, tmp_path):
    """BASELINE.json configs[3] (5 Mbp at 2000x, insert 180+-25: nearly every pair overlaps) at 1/20 of its length — 250 kbp,
    3.3 M alignments, ~13 000 per 4096-position window, so every window is split over several CTAs — whole file against oracle/_ref"""
    p = synth("c4_frac", "--contigs", "panel:250000", "--depth", "2000", "--isize-mean", "180", "--isize-sd", "25", "--isize-min", "150", "--isize-max", "300", "--read-seed", "4242")
    assert _cmp_extract(built, tmp_path, p, [], "panel") > 500_000
    assert _cmp_extract(built, tmp_path, p, ["--CHG", "--CHH", "--minOppositeDepth", "50", "--maxVariantFrac", "0.1"], "panel_var") > 1_000_000


SHARD_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
from methyldackel_b200 import api
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world)
def allsum(x):
    t = torch.tensor([x], dtype=torch.int64); dist.all_reduce(t); return int(t.item())
dev = rank %% max(1, torch.cuda.device_count())
mode, argv = sys.argv[1], sys.argv[2:]
if mode == "extract":
    rc = api.extract_sharded(argv, rank, world, run_main=lambda av: api.extract_main(av, device=dev)[0], barrier=dist.barrier, allreduce_sum=allsum)
else:
    rc = api.mbias_sharded(argv[1:], rank, world, argv[0], run_main=lambda av: api.mbias_main(av, device=dev)[0], barrier=dist.barrier, allreduce_sum=allsum)
dist.destroy_process_group()
sys.exit(rc)
'''


def test_sharded_product_path_on_the_device(built, synth, tmp_path):
    """the sharded drivers (api.extract_sharded / mbias_sharded: one process per shard, shards balanced by the BAM
    index) with libmdgpu as the back end — three processes sharing whatever devices the box has — against the one-process run"""
    import sys
    p = synth("shard_gpu", "--human", "6000000", "--depth", "20", "--read-seed", "31")
    opts = ["--CHG", "--CHH", "--mergeContext"]
    one, many = str(tmp_path / "one"), str(tmp_path / "many")
    assert subprocess.run([NEW_BIN, "extract"] + opts + [p + ".fa", p + ".bam", "-o", one], capture_output=True).returncode == 0
    code = SHARD_WORKER % {"root": cases.ROOT, "port": 29611}
    procs = [subprocess.Popen([sys.executable, "-c", code, "extract"] + opts + [p + ".fa", p + ".bam", "-o", many],
                              env=dict(os.environ, RANK=str(r), WORLD_SIZE="3"), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(3)]
    outs = [q.communicate(timeout=600) for q in procs]
    assert all(q.returncode == 0 for q in procs), [o[1][-1500:] for o in outs]
    for ctx in ("CpG", "CHG", "CHH"):
        a = open("%s_%s.bedGraph" % (one, ctx)).read().replace(one, "X"); bb = open("%s_%s.bedGraph" % (many, ctx)).read().replace(many, "X")
        assert a == bb and a.count("\n") > 10000
    r1 = subprocess.run([NEW_BIN, "mbias", "--noSVG", p + ".fa", p + ".bam"], capture_output=True, text=True)
    code = SHARD_WORKER % {"root": cases.ROOT, "port": 29612}
    procs = [subprocess.Popen([sys.executable, "-c", code, "mbias", str(tmp_path / "mbh"), "--noSVG", p + ".fa", p + ".bam"],
                              env=dict(os.environ, RANK=str(r), WORLD_SIZE="3"), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(3)]
    outs = [q.communicate(timeout=600) for q in procs]
    assert all(q.returncode == 0 for q in procs), [o[1][-1500:] for o in outs]
    assert outs[0][0] == r1.stdout and len(r1.stdout) > 2000
