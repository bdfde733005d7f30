"""TEST-ONLY glue: ctypes binding of oracle/libmdoracle.so (the CPU restatement) and an
``mdh_backend`` that routes tiles to it, so the HOST logic of the product (option parsing, BAM
decode, tiling, chunk replay, formatting) can be checked on a CPU box against oracle/_ref.
Nothing under methyldackel_b200/ imports this."""
import ctypes as C
import os

from methyldackel_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def lib():
    global _lib
    if _lib is None:
        o = C.CDLL(os.path.join(ROOT, "oracle", "libmdoracle.so"))
        o.mdo_extract_tile.argtypes = [C.POINTER(A.MdConfig), C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(A.MdReadsSoa),
                                       C.POINTER(A.MdCall), C.c_uint64, C.POINTER(A.MdTileStats)]
        o.mdo_extract_tile_ce.argtypes = [C.POINTER(A.MdConfig), C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(A.MdReadsSoa),
                                          C.POINTER(A.MdCall), C.c_uint64, C.POINTER(A.MdTileStats)]
        o.mdo_mbias_tile.argtypes = [C.POINTER(A.MdConfig), C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32,
                                     C.POINTER(A.MdReadsSoa), C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(A.MdTileStats)]
        o.mdo_per_read_tile.argtypes = [C.POINTER(A.MdConfig), C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(A.MdReadsSoa), C.POINTER(A.MdReadMeth)]
        o.mdo_set_bed.argtypes = [C.POINTER(A.MdBedRegion), C.c_uint32, C.c_int]; o.mdo_set_bed.restype = None
        o.mdo_mbias_tile_ce.argtypes = [C.POINTER(A.MdConfig), C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32,
                                        C.POINTER(A.MdReadsSoa), C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(A.MdTileStats)]
        o.mdo_strand.argtypes = [C.c_uint16, C.c_uint8]
        o.mdo_admit.argtypes = [C.POINTER(A.MdConfig), C.c_uint16, C.c_uint8, C.c_uint8]
        o.mdo_context.argtypes = [C.c_char_p, C.c_int, C.c_int]
        o.mdo_boost.argtypes = [C.c_uint8]; o.mdo_boost.restype = C.c_uint8
        _lib = o
    return _lib


_emu = None


def emu():
    """tests/native/libmdemu.so — CPU emulation of md_bam_* (built by the `built` fixture / __graft_entry__.build())."""
    global _emu
    if _emu is None:
        e = C.CDLL(os.path.join(ROOT, "tests", "native", "libmdemu.so"))
        e.emu_bam_open.restype = C.c_void_p; e.emu_bam_open.argtypes = [C.c_int32, A.EXTRACT_TILE_FN, A.MBIAS_TILE_FN, C.c_void_p]
        e.emu_bam_close.argtypes = [C.c_void_p]; e.emu_bam_close.restype = None
        e.emu_bam_reset.argtypes = [C.c_void_p]; e.emu_bam_reset.restype = None
        e.emu_bam_push.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(A.MdBgzfBlock), C.c_uint32, C.c_uint32, C.POINTER(A.MdBamSummary)]
        e.emu_bam_get_runs.argtypes = [C.c_void_p, C.POINTER(A.MdBamRun), C.c_uint32]
        e.emu_bam_push_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(A.MdBgzfBlock), C.c_uint32, C.c_uint32]
        e.emu_bam_push_end.argtypes = [C.c_void_p, C.POINTER(A.MdBamSummary)]
        e.emu_bam_extract_run.argtypes = [C.c_void_p, C.c_int, C.POINTER(A.MdTileDesc), C.c_uint32, C.POINTER(A.MdCall), C.c_uint64, C.POINTER(A.MdTileStats)]
        e.emu_bam_mbias_run.argtypes = [C.c_void_p, C.c_int, C.POINTER(A.MdTileDesc), C.c_uint32, C.POINTER(A.MdTileStats)]
        e.emu_bam_tile_view.argtypes = [C.c_void_p, C.POINTER(A.MdReadsSoa)]
        e.emu_bam_fixups.restype = C.c_uint64; e.emu_bam_fixups.argtypes = [C.c_void_p]
        e.emu_last_error.restype = C.c_char_p
        e.emu_bam_prefetch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        e.emu_bam_prefetch_used.restype = C.c_uint64; e.emu_bam_prefetch_used.argtypes = [C.c_void_p]
        e.emu_alloc_pinned.restype = C.c_void_p; e.emu_alloc_pinned.argtypes = [C.c_size_t]
        e.emu_free_pinned.restype = None; e.emu_free_pinned.argtypes = [C.c_void_p]
        e.emu_pinned_stats.restype = C.c_uint64; e.emu_pinned_stats.argtypes = [C.POINTER(C.c_uint64)]
        _emu = e
    return _emu


class OracleBackend:
    """mdh_backend whose slots call the oracle port. Keeps the callbacks alive."""

    def __init__(self, device_decode=False, overlapped=True, staging=False):
        o = lib()
        self.state = {}
        st = self.state

        def create(_u, cfg):
            st["cfg"] = A.MdConfig.from_buffer_copy(cfg.contents)
            st["contigs"] = {}
            st["chunks"] = {}
            st["hist"] = (C.c_uint32 * (4 * 2 * A.MD_MBIAS_MAXLEN * 2))()
            st["lens"] = (C.c_int32 * 4)()
            return 1

        def destroy(_b):
            o.mdo_set_bed(None, 0, 0)          # the port's BED state is global: leave it off for whoever uses the library next
            return None

        def load_contig(_b, tid, seq, n):
            st["contigs"][tid] = (C.string_at(seq, n), n)
            return 0

        def drop_contig(_b, tid):
            st["contigs"].pop(tid, None)
            return 0

        def set_bed(_b, tid, regs, n):
            st.setdefault("bed", {})[tid] = (A.MdBedRegion * max(n, 1))(*[regs[i] for i in range(n)]), n
            return 0

        def use_bed(tid):
            if "bed" in st:
                arr, n = st["bed"].get(tid, ((A.MdBedRegion * 1)(), 0))
                o.mdo_set_bed(arr, n, 1)
            else:
                o.mdo_set_bed(None, 0, 0)

        def per_read_tile(_b, td, reads, chunk, out):
            seq, n = st["contigs"][td.contents.tid]
            return o.mdo_per_read_tile(C.byref(st["cfg"]), seq, n, td.contents.beg, td.contents.end, chunk, reads, out)

        def extract_tile(_b, td, reads, calls, cap, stats):
            seq, n = st["contigs"][td.contents.tid]
            use_bed(td.contents.tid)
            return o.mdo_extract_tile_ce(C.byref(st["cfg"]), seq, n, td.contents.beg, td.contents.end, td.contents.ce_beg, td.contents.ce_end, reads, calls, cap, stats)

        def set_chunks(_b, tid, bounds, n):
            st["chunks"][tid] = (C.c_uint32 * (n + 1))(*[bounds[i] for i in range(n + 1)])
            return 0

        def mbias_tile(_b, td, reads, stats):
            seq, n = st["contigs"][td.contents.tid]
            b = st["chunks"][td.contents.tid]
            use_bed(td.contents.tid)
            return o.mdo_mbias_tile_ce(C.byref(st["cfg"]), seq, n, td.contents.beg, td.contents.end, td.contents.ce_beg, td.contents.ce_end, b, len(b) - 1, reads, st["hist"], st["lens"], stats)

        def mbias_hist(_b, hist, lens):
            C.memmove(hist, st["hist"], C.sizeof(st["hist"]))
            for i in range(4):
                lens[i] = st["lens"][i]
            return 0

        def last_error():
            return b"oracle backend"

        self._keep = [A.CREATE_FN(create), A.DESTROY_FN(destroy), A.LOAD_CONTIG_FN(load_contig), A.DROP_CONTIG_FN(drop_contig),
                      A.EXTRACT_TILE_FN(extract_tile), A.SET_CHUNKS_FN(set_chunks), A.MBIAS_TILE_FN(mbias_tile), A.MBIAS_HIST_FN(mbias_hist),
                      A.LAST_ERROR_FN(last_error)]
        self.be = A.MdhBackend(None, *self._keep)      # async slots stay NULL: the driver then runs tile by tile
        self._keep.append(A.SET_BED_FN(set_bed))
        self.be.set_bed = self._keep[-1]
        self._keep.append(A.PER_READ_FN(per_read_tile))
        self.be.per_read_tile = self._keep[-1]
        if device_decode:
            # md_bam_* emulated on the CPU (tests/native/mdemu.cpp: the kernels' own per-thread bodies in plain loops); the
            # tiles it assembles go to the oracle through the two callbacks above
            e = emu()
            ex_cb, mb_cb = self._keep[4], self._keep[6]
            extra = [A.BAM_OPEN_FN(lambda b, nt: e.emu_bam_open(nt, ex_cb, mb_cb, b)), A.BAM_CLOSE_FN(lambda s_: e.emu_bam_close(s_)), A.BAM_CLOSE_FN(lambda s_: e.emu_bam_reset(s_)),
                     A.BAM_PUSH_FN(lambda s_, c, n, bl, nb, sk, o_: e.emu_bam_push(s_, c, n, bl, nb, sk, o_)),
                     A.BAM_RUNS_FN(lambda s_, r, cap: e.emu_bam_get_runs(s_, r, cap)),
                     A.BAM_EXTRACT_FN(lambda s_, run, td, kh, c, cap, st_: e.emu_bam_extract_run(s_, run, td, kh, c, cap, st_)),
                     A.BAM_MBIAS_FN(lambda s_, run, td, kh, st_: e.emu_bam_mbias_run(s_, run, td, kh, st_))]
            self._keep += extra
            self.be.bam_open, self.be.bam_close, self.be.bam_reset, self.be.bam_push, self.be.bam_get_runs, self.be.bam_extract_run, self.be.bam_mbias_run = extra
            if overlapped:
                two = [A.BAM_PUSH_BEGIN_FN(lambda s_, c, n, bl, nb, sk: e.emu_bam_push_begin(s_, c, n, bl, nb, sk)), A.BAM_PUSH_END_FN(lambda s_, o_: e.emu_bam_push_end(s_, o_))]
                self._keep += two
                self.be.bam_push_begin, self.be.bam_push_end = two
                # md_bam_prefetch: the emulation checks the protocol (same bytes when prefetched and when the push ends)
                st["bam_streams"] = []

                def bam_open_tracked(b, nt):
                    h = e.emu_bam_open(nt, ex_cb, mb_cb, b)
                    st["bam_streams"].append(h)
                    return h

                def bam_close_tracked(s_):
                    st["prefetch_used"] = st.get("prefetch_used", 0) + int(e.emu_bam_prefetch_used(s_))
                    e.emu_bam_close(s_)
                more = [A.BAM_PREFETCH_FN(lambda s_, c, n: e.emu_bam_prefetch(s_, c, n)), A.BAM_OPEN_FN(bam_open_tracked), A.BAM_CLOSE_FN(bam_close_tracked)]
                self._keep += more
                self.be.bam_prefetch, self.be.bam_open, self.be.bam_close = more
            if staging:
                # "page-locked" memory: plain memory that is overwritten when it is released (a reader of a recycled staging
                # buffer then decodes garbage); with it the driver's StagedSegments path runs on the CPU
                pins = [A.PIN_ALLOC_FN(lambda n: e.emu_alloc_pinned(n)), A.PIN_FREE_FN(lambda p_: e.emu_free_pinned(p_))]
                self._keep += pins
                self.be.pinned_alloc, self.be.pinned_free = pins


def run_host_main(which, argv, backend):
    """Calls mdh_extract_main / mdh_mbias_main (argv[0] = sub-command name) with the given backend."""
    h = A.load_host()
    args = [which.encode()] + [a.encode() if isinstance(a, str) else a for a in argv]
    arr = (C.c_char_p * (len(args) + 1))(*args, None)
    fn = {"extract": h.mdh_extract_main, "mbias": h.mdh_mbias_main, "perRead": h.mdh_perread_main}[which]
    return fn(len(args), arr, C.byref(backend.be))
