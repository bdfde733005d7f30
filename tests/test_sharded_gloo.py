"""CPU tier, world_size 2 over gloo: the N>1 path of the drop-in (one process per GPU, genome partitioned into
contiguous runs of reference chunks, no data-path collective) gives byte-identical files to the single-process
run and to oracle/_ref.  The device back end is the oracle port here (there is no GPU in this container);
what is under test is the HOST sharding logic."""
import os
import subprocess
import sys

import cases
from util import run_ref, compare_outputs

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import torch, torch.distributed as dist
import oracle_binding as ob
from methyldackel_b200 import api
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world)
def allsum(x):
    t = torch.tensor([x], dtype=torch.int64); dist.all_reduce(t); return int(t.item())
mode = sys.argv[1]; argv = sys.argv[2:]
fail_rank = int(os.environ.get("MD_TEST_FAIL_RANK", "-1"))
def maybe_fail(sub):
    def run(av):
        if rank == fail_rank:
            if os.environ.get("MD_TEST_FAIL_HOW") == "raise":
                raise RuntimeError("injected failure")
            return -20                                     # failed before any output file was created
        return ob.run_host_main(sub, av, ob.OracleBackend())
    return run
if fail_rank >= 0 and mode == "extract":
    rc = api.extract_sharded(argv, rank, world, run_main=maybe_fail("extract"), barrier=dist.barrier, allreduce_sum=allsum)
elif fail_rank >= 0:
    rc = api.mbias_sharded(argv[1:], rank, world, argv[0], run_main=maybe_fail("mbias"), barrier=dist.barrier, allreduce_sum=allsum if os.environ.get("MD_TEST_ALLSUM") else None)
elif mode == "extract":
    rc = api.extract_sharded(argv, rank, world, run_main=lambda av: ob.run_host_main("extract", av, ob.OracleBackend()), barrier=dist.barrier, allreduce_sum=allsum)
else:
    rc = api.mbias_sharded(argv[1:], rank, world, argv[0], run_main=lambda av: ob.run_host_main("mbias", av, ob.OracleBackend()), barrier=dist.barrier)
dist.destroy_process_group()
sys.exit(rc)
'''


def _launch(mode, argv, port, world=2, extra_env=None, expect_ok=True):
    code = WORKER % {"root": cases.ROOT, "port": port}
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), **(extra_env or {}))
        procs.append(subprocess.Popen([sys.executable, "-c", code, mode] + argv, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]           # a rank left waiting in a collective would hit the timeout
    if expect_ok:
        assert all(p.returncode == 0 for p in procs), [o[1][-2000:] for o in outs]
    else:
        assert all(p.returncode != 0 for p in procs), [(p.returncode, o[1][-500:]) for p, o in zip(procs, outs)]
    return outs


def test_extract_two_ranks_equals_reference(built, synth, tmp_path):
    p = synth("shard", "--contigs", "chr1:70000,chr2:30000,chr3:9000", "--depth", "20")
    for k, opts in enumerate([["--chunkSize", "10000", "--mergeContext", "--CHG"], ["--chunkSize", "7000", "--cytosine_report", "--CHH"],
                              ["--chunkSize", "25000", "--minOppositeDepth", "2", "--maxVariantFrac", "0.1"]]):
        refp, newp = str(tmp_path / ("ref%d" % k)), str(tmp_path / ("new%d" % k))
        r = run_ref(built["ref_bin"], "extract", opts, p + ".fa", p + ".bam", refp)
        assert r.returncode == 0
        outs = _launch("extract", opts + [p + ".fa", p + ".bam", "-o", newp], 29511 + k)
        assert compare_outputs(refp, newp) == []
        assert outs[0][0] == r.stdout            # "N positions were excluded ..." summed over the shards, printed once


def test_mbias_two_ranks_equals_reference(built, synth, tmp_path):
    p = synth("shard", "--contigs", "chr1:70000,chr2:30000,chr3:9000", "--depth", "20")
    opts = ["--noSVG", "--CHG", "--chunkSize", "9000"]
    r = subprocess.run([built["ref_bin"], "mbias"] + opts + [p + ".fa", p + ".bam"], capture_output=True, text=True)
    outs = _launch("mbias", [str(tmp_path / "mb")] + opts + [p + ".fa", p + ".bam"], 29531)
    assert outs[0][0] == r.stdout and len(r.stdout) > 500


def test_a_failed_rank_neither_hangs_nor_leaks_into_the_output(built, synth, tmp_path):
    """rank 1 fails (by return code, or by raising) before it creates its shard: every rank must come back with a
    non-zero code, nothing is merged, and no partial shard or output file is left behind."""
    p = synth("shard", "--contigs", "chr1:70000,chr2:30000,chr3:9000", "--depth", "20")
    for k, how in enumerate(["rc", "raise"]):
        newp = str(tmp_path / ("fail%d" % k))
        _launch("extract", ["--chunkSize", "10000", p + ".fa", p + ".bam", "-o", newp], 29551 + k, extra_env={"MD_TEST_FAIL_RANK": "1", "MD_TEST_FAIL_HOW": how}, expect_ok=False)
        left = [f for f in os.listdir(str(tmp_path)) if f.startswith("fail%d" % k)]
        assert left == [], left
    for k, allsum in enumerate(["", "1"]):
        outs = _launch("mbias", [str(tmp_path / ("mbf%d" % k)), "--noSVG", p + ".fa", p + ".bam"], 29561 + k, extra_env={"MD_TEST_FAIL_RANK": "1", "MD_TEST_ALLSUM": allsum}, expect_ok=False)
        assert outs[0][0] == ""                                 # no report from a failed run
        assert [f for f in os.listdir(str(tmp_path)) if f.startswith("mbf%d" % k)] == []


def test_copy_into_pieces_offsets_and_fallback(tmp_path, monkeypatch):
    """the shard merge: many pieces on several threads land at the right offsets; without copy_file_range it reads and writes"""
    import os
    from methyldackel_b200 import api
    data = os.urandom(3 * 1000 * 1000 + 17)
    src, dst = tmp_path / "shard", tmp_path / "final"
    src.write_bytes(data)
    for threads, piece in ((8, 256 * 1024), (1, 1 << 20), (4, 1 << 30)):
        dst.write_bytes(b"H" * 100 + b"\0" * len(data) + b"T" * 50)
        api._copy_into(str(src), str(dst), 100, threads=threads, piece=piece)
        assert dst.read_bytes() == b"H" * 100 + data + b"T" * 50
    monkeypatch.delattr(os, "copy_file_range", raising=False)          # e.g. a file system that does not offer it
    dst.write_bytes(b"H" * 100 + b"\0" * len(data) + b"T" * 50)
    api._copy_into(str(src), str(dst), 100, threads=3, piece=700 * 1000)
    assert dst.read_bytes() == b"H" * 100 + data + b"T" * 50
    (tmp_path / "empty").write_bytes(b"")
    api._copy_into(str(tmp_path / "empty"), str(dst), 100)
    assert dst.read_bytes() == b"H" * 100 + data + b"T" * 50
