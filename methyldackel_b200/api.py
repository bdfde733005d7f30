"""Python mirror of the reference-facing surface of the B200 path.

``extract_main(argv)`` / ``mbias_main(argv)`` take the same argv as ``MethylDackel extract`` /
``MethylDackel mbias`` (extract.c:706, MBias.c:304) and run the host driver bound to the CUDA
library.  ``GpuContext`` wraps the per-tile C ABI (include/mdgpu.h).  There is no CPU fallback:
everything here raises if ``lib/libmdgpu.so`` is missing or no CUDA device is present.
"""
import ctypes as C

from . import _abi as A


class GpuContext:
    """One md_ctx (one CUDA device). Replaces the per-thread state of extractCalls (extract.c:283-323)."""

    def __init__(self, cfg=None, device=0):
        self.g = A.load_gpu()
        self.cfg = cfg if cfg is not None else A.default_config()
        self.h = self.g.md_create(C.byref(self.cfg), device)
        if not self.h:
            raise RuntimeError("md_create failed: %s" % self.g.md_last_error().decode())

    def close(self):
        if self.h:
            self.g.md_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("mdgpu error %d: %s" % (rc, self.g.md_last_error().decode()))

    def load_contig(self, tid, seq):
        self._chk(self.g.md_load_contig(self.h, tid, seq, len(seq)))

    def set_mbias_chunks(self, tid, bounds):
        arr = (C.c_uint32 * len(bounds))(*bounds)
        self._chk(self.g.md_set_mbias_chunks(self.h, tid, arr, len(bounds) - 1))

    def extract_tile(self, tid, beg, end, reads, capacity=None):
        """Host-buffer path (H2D + kernels + D2H). Returns (list of (pos, nmeth, nunmeth, info), stats)."""
        cap = capacity if capacity is not None else (end - beg) + 16
        calls = (A.MdCall * cap)()
        st = A.MdTileStats()
        td = A.MdTileDesc(tid, beg, end)
        self._chk(self.g.md_extract_tile(self.h, C.byref(td), C.byref(reads), calls, cap, C.byref(st)))
        return calls, st

    def mbias_tile(self, tid, beg, end, reads):
        st = A.MdTileStats()
        td = A.MdTileDesc(tid, beg, end)
        self._chk(self.g.md_mbias_tile(self.h, C.byref(td), C.byref(reads), C.byref(st)))
        return st

    def mbias_hist(self):
        hist = (C.c_uint32 * (4 * 2 * A.MD_MBIAS_MAXLEN * 2))()
        lens = (C.c_int32 * 4)()
        self._chk(self.g.md_mbias_hist(self.h, hist, lens))
        return hist, list(lens)

    def upload(self, reads):
        d = self.g.md_upload_reads(self.h, C.byref(reads))
        if not d:
            raise RuntimeError("md_upload_reads failed: %s" % self.g.md_last_error().decode())
        return d

    def free(self, d):
        self.g.md_free_reads(self.h, d)

    def extract_tile_device(self, tid, beg, end, dreads):
        st = A.MdTileStats()
        td = A.MdTileDesc(tid, beg, end)
        self._chk(self.g.md_extract_tile_device(self.h, C.byref(td), dreads, C.byref(st)))
        return st

    def mbias_tile_device(self, tid, beg, end, dreads):
        st = A.MdTileStats()
        td = A.MdTileDesc(tid, beg, end)
        self._chk(self.g.md_mbias_tile_device(self.h, C.byref(td), dreads, C.byref(st)))
        return st

    def fetch_calls(self, capacity):
        calls = (A.MdCall * capacity)()
        n = C.c_uint64(0)
        self._chk(self.g.md_fetch_calls(self.h, calls, capacity, C.byref(n)))
        return calls, n.value

    def last_timing(self):
        t = (C.c_float * 5)()
        self.g.md_last_timing(self.h, t)
        return list(t)

    def launch_count(self):
        return int(self.g.md_launch_count(self.h))


def _run(which, argv, device):
    """The sub-command main of libmdhost driven by the NATIVE back-end table of libMethylDackel.so (every slot bound to
    libmdgpu on `device`): no Python on the data path."""
    h = A.load_host()
    be = A.load_dropin().mdh_gpu_backend(device)
    args = [which.encode()] + [a.encode() if isinstance(a, str) else a for a in argv]
    arr = (C.c_char_p * (len(args) + 1))(*args, None)
    fn = {"extract": h.mdh_extract_main, "mbias": h.mdh_mbias_main, "perRead": h.mdh_perread_main}[which]
    rc = fn(len(args), arr, be)
    st = A.MdhRunStats()
    h.mdh_last_run_stats(C.byref(st))
    return rc, st


def last_totals():
    """md_totals of the device context the last sub-command main used (CUDA-event kernel time, tiles, alignments ...)."""
    t = A.MdTotals()
    A.load_gpu().md_last_totals(C.byref(t))
    return t


def perread_main(argv, device=0):
    """`MethylDackel perRead <argv>` on the GPU. Returns (exit code, run stats)."""
    return _run("perRead", list(argv), device)


def extract_main(argv, device=0):
    """`MethylDackel extract <argv>` on the GPU. Returns (exit code, run stats)."""
    return _run("extract", list(argv), device)


def mbias_main(argv, device=0):
    """`MethylDackel mbias <argv>` on the GPU. Returns (exit code, run stats)."""
    return _run("mbias", list(argv), device)


class BamFile:
    def __init__(self, path):
        self.h = A.load_host()
        self.p = self.h.mdh_bam_open(path.encode())
        if not self.p:
            raise IOError(self.h.mdh_last_error().decode())
        self.names = [self.h.mdh_bam_target_name(self.p, i).decode() for i in range(self.h.mdh_bam_n_targets(self.p))]
        self.lens = [self.h.mdh_bam_target_len(self.p, i) for i in range(len(self.names))]

    def read_region(self, tid, beg=0, end=None):
        """All alignments of contig tid overlapping [beg,end) as one md_reads_soa (valid until the next call)."""
        soa = A.MdReadsSoa()
        if end is None:
            end = self.lens[tid]
        if self.h.mdh_bam_read_region(self.p, tid, beg, end, C.byref(soa)) != 0:
            raise IOError(self.h.mdh_last_error().decode())
        return soa

    def make_tiles(self, tid, beg=0, end=None, target_reads=1 << 17):
        """Cuts the region into tiles as the sub-command driver does; returns [(MdTileDesc, MdReadsSoa), ...]."""
        if end is None:
            end = self.lens[tid]
        n = self.h.mdh_bam_make_tiles(self.p, tid, beg, end, target_reads)
        if n < 0:
            raise IOError(self.h.mdh_last_error().decode())
        out = []
        for k in range(n):
            td, soa = A.MdTileDesc(), A.MdReadsSoa()
            self.h.mdh_bam_get_tile(self.p, k, C.byref(td), C.byref(soa))
            out.append((td, soa))
        return out

    def close(self):
        if self.p:
            self.h.mdh_bam_close(self.p)
            self.p = None


def read_region(bam_path, tid, beg=0, end=None):
    b = BamFile(bam_path)
    return b, b.read_region(tid, beg, end)


def fetch_contig(fasta_path, name):
    h = A.load_host()
    f = h.mdh_fasta_open(fasta_path.encode())
    if not f:
        raise IOError(h.mdh_last_error().decode())
    n = C.c_uint32(0)
    p = h.mdh_fasta_fetch(f, name.encode(), C.byref(n))
    if not p:
        h.mdh_fasta_close(f)
        raise KeyError(name)
    s = C.string_at(p, n.value)
    h.mdh_fasta_close(f)
    return s


# ---------------------------------------------------------------------------------------------------
# one process per GPU: the genome is partitioned into contiguous runs of reference chunks, one run per
# rank (positions are independent, mates that straddle a cut are delivered to both sides, so
# there is no data-path collective); rank 0 concatenates the per-rank text in genome order.
def _output_names(argv):
    """Output file names `extract` will write for this argv (extract.c:1344-1439)."""
    a = list(argv)
    flags = {x for x in a if x.startswith("-")}
    pos = []
    takes = {"-q", "-p", "-r", "-l", "-o", "-D", "-d", "-F", "-R", "-@", "-M", "-t", "-b", "-N", "-B", "--opref", "--minDepth", "--OT", "--OB", "--CTOT", "--CTOB",
             "--nOT", "--nOB", "--nCTOT", "--nCTOB", "--minOppositeDepth", "--maxVariantFrac", "--chunkSize", "--minConversionEfficiency", "--ignoreFlags",
             "--requireFlags", "--shardRank", "--shardWorld"}
    opref, i = None, 0
    while i < len(a):
        if a[i] in ("-o", "--opref"):
            opref = a[i + 1]
        if a[i] in takes:
            i += 2
        elif a[i].startswith("-"):
            i += 1
        else:
            pos.append(a[i]); i += 1
    if opref is None:
        opref = pos[1].rsplit(".", 1)[0] if "." in pos[1] else pos[1]
    if "--cytosine_report" in flags:
        return [opref + ".cytosine_report.txt"]
    mid = ".meth.bedGraph" if flags & {"--fraction", "-f"} else ".counts.bedGraph" if flags & {"--counts", "-c"} else ".logit.bedGraph" if flags & {"--logit", "-m"} \
        else ".methylKit" if "--methylKit" in flags else ".bedGraph"
    out = []
    if "--noCpG" not in flags:
        out.append(opref + "_CpG" + mid)
    if "--CHG" in flags:
        out.append(opref + "_CHG" + mid)
    if "--CHH" in flags:
        out.append(opref + "_CHH" + mid)
    return out


def _default_group():
    import torch
    import torch.distributed as dist

    def allreduce_sum(x):
        t = torch.tensor([x], dtype=torch.int64, device="cuda" if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(t)
        return int(t.item())
    return dist.barrier, allreduce_sum


def extract_sharded(argv, rank, world, run_main=None, barrier=None, allreduce_sum=None):
    """`MethylDackel extract <argv>` with the genome split over `world` processes (one per GPU).

    ``run_main(argv)`` runs the sub-command main in this process (default: the CUDA library on device
    ``rank``); ``barrier()`` and ``allreduce_sum(int)`` come from the launcher's process group (default:
    torch.distributed).  Returns the exit code: this rank's own when it failed, -20 when another rank failed.

    A failed rank must neither hang the others nor leak into the result: every rank always reaches both collectives
    (an exception inside run_main counts as a failure), and rank 0 only concatenates the shard files when ALL ranks
    succeeded and every shard file exists; otherwise the partial shard files and outputs are removed."""
    import os
    import sys
    if run_main is None:
        run_main = lambda av: extract_main(av, device=rank)[0]  # noqa: E731
    if barrier is None or allreduce_sum is None:
        barrier, allreduce_sum = _default_group()
    names = []
    nvar_local = 0
    import time
    t0 = time.time()
    marks = []                                                 # MD_TIMING=1: where a sharded run spends its wall time, per rank
    try:
        names = _output_names(argv)
        rc = run_main(list(argv) + ["--shardRank", str(rank), "--shardWorld", str(world)])
        marks.append(("shard done", time.time() - t0))
        st = A.MdhRunStats()
        A.load_host().mdh_last_run_stats(C.byref(st))
        nvar_local = int(st.n_variant_positions)
    except Exception as e:  # noqa: BLE001 - the collectives below must still be reached
        print("extract_sharded: rank %d failed: %s" % (rank, e), file=sys.stderr)
        rc = -20
    nvar = allreduce_sum(nvar_local if rc == 0 else 0)
    worst = allreduce_sum(1 if rc != 0 else 0)                 # also orders every rank's shard files before anyone reads them
    # Merge: every rank copies ITS shard into the final file at the offset the lower ranks' sizes give it — world copies in
    # parallel, done by the kernel (copy_file_range), instead of one process reading and rewriting the whole output.
    parts = [["%s.shard%d" % (name, r) for r in range(world)] for name in names]
    mine_ok = 1 if (worst == 0 and all(os.path.exists(ps[rank]) for ps in parts)) else 0
    all_ok = allreduce_sum(mine_ok) == world
    marks.append(("all shards done", time.time() - t0))
    merged_ok = 1
    if all_ok:                                                 # every rank takes the same branch: the collectives below stay matched
        sizes = [[allreduce_sum(os.path.getsize(ps[r]) if r == rank else 0) for r in range(world)] for ps in parts]
        if rank == 0:
            try:
                for name, sz in zip(names, sizes):
                    with open(name, "wb") as out:
                        out.truncate(sum(sz))
            except OSError as e:
                merged_ok = 0
                print("extract_sharded: cannot create the output files: %s" % e, file=sys.stderr)
        barrier()
        try:
            for name, ps, sz in zip(names, parts, sizes):
                if os.path.exists(name):
                    _copy_into(ps[rank], name, sum(sz[:rank]))
                else:
                    merged_ok = 0
        except OSError as e:
            merged_ok = 0
            print("extract_sharded: merging the shards failed on rank %d: %s" % (rank, e), file=sys.stderr)
    else:
        merged_ok = 0
        if rank == 0 and worst == 0:
            print("extract_sharded: a shard file is missing; nothing was merged", file=sys.stderr)
    marks.append(("own shard copied", time.time() - t0))
    merged_ok = 1 if allreduce_sum(merged_ok) == world else 0   # doubles as the barrier after the copies; tells every rank the outcome
    marks.append(("all copied", time.time() - t0))
    for ps in parts:
        if os.path.exists(ps[rank]):
            os.unlink(ps[rank])
    if rank == 0:
        if merged_ok:
            if nvar:
                print("%d positions were excluded due to likely being variants." % nvar)
                sys.stdout.flush()
        else:
            for ps in parts:                                   # a rank that died before this point leaves its shard behind
                for part in ps:
                    if os.path.exists(part):
                        os.unlink(part)
            for name in names:                                 # no normal-looking output from a failed run
                if os.path.exists(name):
                    os.unlink(name)
    barrier()
    if os.environ.get("MD_TIMING"):
        marks.append(("cleaned up", time.time() - t0))
        print("[md-timing] extract_sharded rank %d: %s" % (rank, ", ".join("%s %.3f" % m for m in marks)), file=sys.stderr)
    if rc != 0:
        return rc
    return 0 if (worst == 0 and merged_ok) else -20


def _copy_into(src, dst, offset, threads=8, piece=64 << 20):
    """src -> dst[offset:], in the kernel where the file system allows it (copy_file_range), several pieces at a time: one
    thread moves ~2 GB/s through tmpfs, and a rank's shard of a genome-wide extract is gigabytes"""
    import os
    from concurrent.futures import ThreadPoolExecutor
    n = os.path.getsize(src)
    if n == 0:
        return

    def move(lo, hi):
        with open(src, "rb") as fi, open(dst, "r+b") as fo:
            done = lo
            try:
                while done < hi:
                    k = os.copy_file_range(fi.fileno(), fo.fileno(), min(hi - done, 1 << 30), done, offset + done)
                    if k <= 0:
                        raise OSError("copy_file_range made no progress")
                    done += k
            except (OSError, AttributeError):
                fi.seek(done); fo.seek(offset + done)
                while done < hi:
                    blk = fi.read(min(1 << 24, hi - done))
                    if not blk:
                        raise OSError("short read while merging %s" % src)
                    fo.write(blk); done += len(blk)

    cuts = list(range(0, n, piece)) + [n]
    if len(cuts) <= 2 or threads <= 1:
        move(0, n)
        return
    with ThreadPoolExecutor(max_workers=threads) as ex:
        for f in [ex.submit(move, cuts[k], cuts[k + 1]) for k in range(len(cuts) - 1)]:
            f.result()                                      # re-raises a worker's OSError


def mbias_sharded(argv, rank, world, tmp_prefix, run_main=None, barrier=None, allreduce_sum=None):
    """`MethylDackel mbias <argv>` over `world` processes: every rank histograms its run of chunks, rank 0 sums the
    (<= 64 KB) histograms on the host (the analogue of mergeStrandMeth, MBias.c:42-55) and prints the report.
    As in extract_sharded, the report is only produced when every rank succeeded and every histogram file is complete."""
    import os
    import sys
    import numpy as np
    if run_main is None:
        run_main = lambda av: mbias_main(av, device=rank)[0]  # noqa: E731
    if barrier is None:
        import torch.distributed as dist
        barrier = dist.barrier
    part = "%s.hist%d" % (tmp_prefix, rank)
    n = 4 * 2 * A.MD_MBIAS_MAXLEN * 2
    try:
        rc = run_main(list(argv) + ["--shardRank", str(rank), "--shardWorld", str(world), "--histOut", part])
        if rc == 0 and (not os.path.exists(part) or os.path.getsize(part) != 4 * (4 + n)):
            rc = -3
    except Exception as e:  # noqa: BLE001
        print("mbias_sharded: rank %d failed: %s" % (rank, e), file=sys.stderr)
        rc = -20
    if allreduce_sum is not None:
        worst = allreduce_sum(1 if rc != 0 else 0)
    else:                                                      # no reduction available: a failed rank leaves a marker file next to its histogram
        if rc != 0:
            open(part + ".failed", "w").close()
        barrier()
        worst = sum(os.path.exists("%s.hist%d.failed" % (tmp_prefix, r)) for r in range(world))
    barrier()
    if rank == 0:
        try:
            if worst == 0:
                hist = np.zeros(n, dtype=np.uint32); lens = np.zeros(4, dtype=np.int32)
                for r in range(world):
                    raw = np.fromfile("%s.hist%d" % (tmp_prefix, r), dtype=np.uint32)
                    lens = np.maximum(lens, raw[:4].view(np.int32))
                    hist += raw[4:4 + n]
                flags = set(argv)
                txt = 1 if ("--txt" in flags or "--noSVG" in flags) else 0
                opref = None
                if "--noSVG" not in flags:                     # the third positional argument is the plot prefix (MBias.c:445-449)
                    takes = {"-q", "-p", "-r", "-l", "-D", "-F", "-R", "-@", "--nOT", "--nOB", "--nCTOT", "--nCTOB", "--chunkSize", "--minConversionEfficiency",
                             "--ignoreFlags", "--requireFlags"}
                    pos, i = [], 0
                    while i < len(argv):
                        if argv[i] in takes:
                            i += 2
                        elif argv[i].startswith("-"):
                            i += 1
                        else:
                            pos.append(argv[i]); i += 1
                    opref = pos[2].encode() if len(pos) > 2 else None
                which = (0 if "--noCpG" in flags else 1) + (2 if "--CHG" in flags else 0) + (4 if "--CHH" in flags else 0)
                if A.load_host().mdh_mbias_report_svg(hist.ctypes.data_as(C.POINTER(C.c_uint32)), lens.ctypes.data_as(C.POINTER(C.c_int32)), opref, which, txt) != 0:
                    worst = 1
        finally:
            for r in range(world):
                for x in ("%s.hist%d" % (tmp_prefix, r), "%s.hist%d.failed" % (tmp_prefix, r)):
                    if os.path.exists(x):
                        os.unlink(x)
    barrier()
    if rc != 0:
        return rc
    return 0 if worst == 0 else -20
