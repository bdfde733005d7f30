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


class _GpuBackend:
    """mdh_backend bound to libmdgpu — the only binding the product uses."""

    def __init__(self, device=0):
        g = A.load_gpu()
        self.g = g

        def create(_u, cfg):
            return g.md_create(cfg, device)

        def last_error():
            return g.md_last_error()

        self._keep = [A.CREATE_FN(create), A.DESTROY_FN(lambda b: g.md_destroy(b)),
                      A.LOAD_CONTIG_FN(lambda b, t, s, n: g.md_load_contig(b, t, s, n)),
                      A.DROP_CONTIG_FN(lambda b, t: g.md_drop_contig(b, t)),
                      A.EXTRACT_TILE_FN(lambda b, td, r, c, cap, st: g.md_extract_tile(b, td, r, c, cap, st)),
                      A.SET_CHUNKS_FN(lambda b, t, bo, n: g.md_set_mbias_chunks(b, t, bo, n)),
                      A.MBIAS_TILE_FN(lambda b, td, r, st: g.md_mbias_tile(b, td, r, st)),
                      A.MBIAS_HIST_FN(lambda b, h, l: g.md_mbias_hist(b, h, l)),
                      A.LAST_ERROR_FN(last_error)]
        self.be = A.MdhBackend(None, *self._keep)


def _run(which, argv, device):
    h = A.load_host()
    be = _GpuBackend(device)
    args = [which.encode()] + [a.encode() if isinstance(a, str) else a for a in argv]
    arr = (C.c_char_p * (len(args) + 1))(*args, None)
    fn = h.mdh_extract_main if which == "extract" else h.mdh_mbias_main
    rc = fn(len(args), arr, C.byref(be.be))
    st = A.MdhRunStats()
    h.mdh_last_run_stats(C.byref(st))
    return rc, st


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
