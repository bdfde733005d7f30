"""ctypes mirror of include/mdgpu.h and include/mdhost.h.

PyTorch is not involved in the data path: the product is the C-ABI library
``lib/libmdgpu.so`` (hand-written CUDA, sm_100a) driven by ``lib/libmdhost.so``.
This module only describes the structs and loads the shared objects.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.path.join(HERE, "lib")
MD_MBIAS_MAXLEN = 1024


class MdConfig(C.Structure):
    """md_config (include/mdgpu.h) — the hot-path subset of the reference's Config (MethylDackel.h:90-126)."""
    _fields_ = [
        ("keepCpG", C.c_int32), ("keepCHG", C.c_int32), ("keepCHH", C.c_int32),
        ("minMapq", C.c_int32), ("minPhred", C.c_int32),
        ("keepDupes", C.c_int32), ("keepSingleton", C.c_int32), ("keepDiscordant", C.c_int32),
        ("ignoreFlags", C.c_int32), ("requireFlags", C.c_int32), ("ignoreNH", C.c_int32),
        ("minOppositeDepth", C.c_int32), ("maxVariantFrac", C.c_double),
        ("bounds", C.c_int32 * 16), ("absoluteBounds", C.c_int32 * 16),
        ("noOverlapMerge", C.c_int32), ("minConversionEfficiency", C.c_float), ("reserved", C.c_int32 * 6),
    ]


def default_config(**kw):
    """Defaults of extract_main (extract.c:725-746): CpG only, -q 10, -p 5, -F 0xF00."""
    c = MdConfig()
    c.keepCpG, c.minMapq, c.minPhred, c.ignoreFlags = 1, 10, 5, 0xF00
    for k, v in kw.items():
        if k in ("bounds", "absoluteBounds"):
            for i, x in enumerate(v):
                getattr(c, k)[i] = x
        else:
            setattr(c, k, v)
    return c


class MdReadsSoa(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint32), ("n_cigar_ops", C.c_uint32), ("seq_words", C.c_uint64), ("qual_words", C.c_uint64),
        ("pos", C.POINTER(C.c_int32)), ("flag", C.POINTER(C.c_uint16)), ("mapq", C.POINTER(C.c_uint8)), ("aux", C.POINTER(C.c_uint8)),
        ("l_qseq", C.POINTER(C.c_uint32)), ("cigar_off", C.POINTER(C.c_uint32)), ("seq_off", C.POINTER(C.c_uint32)), ("qual_off", C.POINTER(C.c_uint32)),
        ("frag_key", C.POINTER(C.c_uint64)), ("cigar", C.POINTER(C.c_uint32)), ("seq", C.POINTER(C.c_uint32)), ("qual", C.POINTER(C.c_uint64)),
        ("qual_bits", C.c_uint32), ("qual_lut", C.c_uint8 * 16), ("reserved_", C.c_uint8 * 12),
    ]


class MdReadMeth(C.Structure):
    _fields_ = [("nmeth", C.c_uint32), ("nunmeth", C.c_uint32)]


class MdBedRegion(C.Structure):
    _fields_ = [("start", C.c_uint32), ("end", C.c_uint32), ("strand", C.c_uint32)]


class MdTileDesc(C.Structure):
    _fields_ = [("tid", C.c_int32), ("beg", C.c_uint32), ("end", C.c_uint32), ("ce_beg", C.c_uint32), ("ce_end", C.c_uint32)]


class MdCall(C.Structure):
    _fields_ = [("pos", C.c_uint32), ("nmeth", C.c_uint32), ("nunmeth", C.c_uint32), ("info", C.c_uint32)]


class MdTileStats(C.Structure):
    _fields_ = [("n_calls", C.c_uint64), ("n_required", C.c_uint64), ("n_admitted", C.c_uint32), ("n_pairs", C.c_uint32),
                ("n_multi", C.c_uint32), ("reserved", C.c_uint32)]


class MdTotals(C.Structure):
    """md_totals (include/mdgpu.h): device time (CUDA events, ms) and work of one context over all its tiles"""
    _fields_ = [("h2d_ms", C.c_double), ("prep_ms", C.c_double), ("count_ms", C.c_double), ("d2h_ms", C.c_double), ("tile_ms", C.c_double),
                ("inflate_ms", C.c_double), ("frame_ms", C.c_double), ("push_h2d_ms", C.c_double),
                ("tiles", C.c_uint64), ("alignments", C.c_uint64), ("cigar_ops", C.c_uint64), ("calls", C.c_uint64), ("launches", C.c_uint64),
                ("comp_bytes", C.c_uint64), ("inflated_bytes", C.c_uint64)]


class MdhRunStats(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("n_tiles", C.c_uint64), ("n_calls", C.c_uint64),
                ("t_decode_s", C.c_double), ("t_device_s", C.c_double), ("t_format_s", C.c_double), ("t_total_s", C.c_double),
                ("n_variant_positions", C.c_uint64)]


CREATE_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.POINTER(MdConfig))
DESTROY_FN = C.CFUNCTYPE(None, C.c_void_p)
LOAD_CONTIG_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.c_void_p, C.c_uint32)
DROP_CONTIG_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32)
EXTRACT_TILE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(MdTileDesc), C.POINTER(MdReadsSoa), C.POINTER(MdCall), C.c_uint64, C.POINTER(MdTileStats))
SET_CHUNKS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.POINTER(C.c_uint32), C.c_uint32)
MBIAS_TILE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(MdTileDesc), C.POINTER(MdReadsSoa), C.POINTER(MdTileStats))
MBIAS_HIST_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_int32))
LAST_ERROR_FN = C.CFUNCTYPE(C.c_char_p)
SUBMIT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(MdTileDesc), C.POINTER(MdReadsSoa))
COLLECT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(MdCall), C.c_uint64, C.POINTER(MdTileStats))
PIN_ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_size_t)
PIN_FREE_FN = C.CFUNCTYPE(None, C.c_void_p)


class MdBgzfBlock(C.Structure):
    _fields_ = [("comp_off", C.c_uint64), ("comp_len", C.c_uint32), ("isize", C.c_uint32)]


class MdBamRun(C.Structure):
    _fields_ = [("tid", C.c_int32), ("start", C.c_uint32), ("n", C.c_uint32), ("first_pos", C.c_int32), ("last_pos", C.c_int32), ("prev_last_pos", C.c_int32)]


class MdBamSummary(C.Structure):
    _fields_ = [("n_records", C.c_uint32), ("n_runs", C.c_uint32), ("inflated_bytes", C.c_uint64), ("leftover_bytes", C.c_uint64)]


BAM_OPEN_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_int32)
BAM_CLOSE_FN = C.CFUNCTYPE(None, C.c_void_p)
BAM_PUSH_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(MdBgzfBlock), C.c_uint32, C.c_uint32, C.POINTER(MdBamSummary))
BAM_RUNS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(MdBamRun), C.c_uint32)
BAM_EXTRACT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(MdTileDesc), C.c_uint32, C.POINTER(MdCall), C.c_uint64, C.POINTER(MdTileStats))
BAM_MBIAS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(MdTileDesc), C.c_uint32, C.POINTER(MdTileStats))
BAM_PUSH_BEGIN_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(MdBgzfBlock), C.c_uint32, C.c_uint32)
BAM_PUSH_END_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(MdBamSummary))
BAM_PREFETCH_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64)
SET_BED_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.POINTER(MdBedRegion), C.c_uint32)
PER_READ_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(MdTileDesc), C.POINTER(MdReadsSoa), C.c_uint32, C.POINTER(MdReadMeth))


class MdhBackend(C.Structure):
    _fields_ = [("factory_user", C.c_void_p), ("create", CREATE_FN), ("destroy", DESTROY_FN), ("load_contig", LOAD_CONTIG_FN),
                ("drop_contig", DROP_CONTIG_FN), ("extract_tile", EXTRACT_TILE_FN), ("set_mbias_chunks", SET_CHUNKS_FN),
                ("mbias_tile", MBIAS_TILE_FN), ("mbias_hist", MBIAS_HIST_FN), ("last_error", LAST_ERROR_FN),
                ("submit_tile", SUBMIT_FN), ("collect_tile", COLLECT_FN), ("pinned_alloc", PIN_ALLOC_FN), ("pinned_free", PIN_FREE_FN),
                ("submit_mbias_tile", SUBMIT_FN),
                ("bam_open", BAM_OPEN_FN), ("bam_close", BAM_CLOSE_FN), ("bam_reset", BAM_CLOSE_FN), ("bam_push", BAM_PUSH_FN),
                ("bam_get_runs", BAM_RUNS_FN), ("bam_extract_run", BAM_EXTRACT_FN), ("bam_mbias_run", BAM_MBIAS_FN),
                ("bam_push_begin", BAM_PUSH_BEGIN_FN), ("bam_push_end", BAM_PUSH_END_FN), ("set_bed", SET_BED_FN), ("per_read_tile", PER_READ_FN),
                ("bam_prefetch", BAM_PREFETCH_FN)]


_host = None
_gpu = None


def load_host():
    """libmdhost.so: BAM decode -> SoA, chunk replay, formatter, sub-command mains (no CUDA dependency)."""
    global _host
    if _host is None:
        p = os.path.join(LIBDIR, "libmdhost.so")
        if not os.path.exists(p):
            raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (or make -C methyldackel_b200/csrc host)" % p)
        h = C.CDLL(p)
        h.mdh_extract_main.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(MdhBackend)]
        h.mdh_mbias_main.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(MdhBackend)]
        h.mdh_perread_main.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(MdhBackend)]
        h.mdh_last_run_stats.argtypes = [C.POINTER(MdhRunStats)]
        h.mdh_bam_open.restype = C.c_void_p; h.mdh_bam_open.argtypes = [C.c_char_p]
        h.mdh_bam_close.argtypes = [C.c_void_p]
        h.mdh_bam_n_targets.argtypes = [C.c_void_p]
        h.mdh_bam_target_name.restype = C.c_char_p; h.mdh_bam_target_name.argtypes = [C.c_void_p, C.c_int]
        h.mdh_bam_target_len.restype = C.c_uint32; h.mdh_bam_target_len.argtypes = [C.c_void_p, C.c_int]
        h.mdh_bam_read_region.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(MdReadsSoa)]
        h.mdh_bam_make_tiles.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint64]
        h.mdh_bam_get_tile.argtypes = [C.c_void_p, C.c_int, C.POINTER(MdTileDesc), C.POINTER(MdReadsSoa)]
        h.mdh_fasta_open.restype = C.c_void_p; h.mdh_fasta_open.argtypes = [C.c_char_p]
        h.mdh_fasta_close.argtypes = [C.c_void_p]
        h.mdh_fasta_fetch.restype = C.c_void_p; h.mdh_fasta_fetch.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint32)]
        h.mdh_chunk_bounds.restype = C.c_uint32
        h.mdh_chunk_bounds.argtypes = [C.c_char_p, C.c_uint32, C.c_ulong, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32]
        h.mdh_mbias_report.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.c_int, C.c_int]
        h.mdh_mbias_report_svg.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.c_char_p, C.c_int, C.c_int]
        h.mdh_last_error.restype = C.c_char_p
        _host = h
    return _host


def load_gpu():
    """libmdgpu.so: the CUDA library. Fails loudly if it has not been built — there is no CPU fallback."""
    global _gpu
    if _gpu is None:
        p = os.environ.get("MD_LIBMDGPU") or os.path.join(LIBDIR, "libmdgpu.so")      # MD_LIBMDGPU: an experimental build of the same ABI (A/B runs)
        if not os.path.exists(p):
            raise RuntimeError("%s is missing: the CUDA extension must be built (make -C methyldackel_b200/csrc gpu); there is no CPU fallback" % p)
        g = C.CDLL(p, mode=C.RTLD_GLOBAL)
        g.md_create.restype = C.c_void_p; g.md_create.argtypes = [C.POINTER(MdConfig), C.c_int]
        g.md_destroy.argtypes = [C.c_void_p]
        g.md_load_contig.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_uint32]
        g.md_drop_contig.argtypes = [C.c_void_p, C.c_int32]
        g.md_set_mbias_chunks.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_uint32), C.c_uint32]
        g.md_set_bed.argtypes = [C.c_void_p, C.c_int32, C.POINTER(MdBedRegion), C.c_uint32]
        g.md_per_read_tile.argtypes = [C.c_void_p, C.POINTER(MdTileDesc), C.POINTER(MdReadsSoa), C.c_uint32, C.POINTER(MdReadMeth)]
        g.md_extract_tile.argtypes = [C.c_void_p, C.POINTER(MdTileDesc), C.POINTER(MdReadsSoa), C.POINTER(MdCall), C.c_uint64, C.POINTER(MdTileStats)]
        g.md_submit_tile.argtypes = [C.c_void_p, C.POINTER(MdTileDesc), C.POINTER(MdReadsSoa)]
        g.md_submit_mbias_tile.argtypes = [C.c_void_p, C.POINTER(MdTileDesc), C.POINTER(MdReadsSoa)]
        g.md_collect_tile.argtypes = [C.c_void_p, C.c_int, C.POINTER(MdCall), C.c_uint64, C.POINTER(MdTileStats)]
        g.md_mbias_tile.argtypes = [C.c_void_p, C.POINTER(MdTileDesc), C.POINTER(MdReadsSoa), C.POINTER(MdTileStats)]
        g.md_mbias_hist.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]
        g.md_mbias_reset.argtypes = [C.c_void_p]
        g.md_upload_reads.restype = C.c_void_p; g.md_upload_reads.argtypes = [C.c_void_p, C.POINTER(MdReadsSoa)]
        g.md_free_reads.argtypes = [C.c_void_p, C.c_void_p]
        g.md_extract_tile_device.argtypes = [C.c_void_p, C.POINTER(MdTileDesc), C.c_void_p, C.POINTER(MdTileStats)]
        g.md_mbias_tile_device.argtypes = [C.c_void_p, C.POINTER(MdTileDesc), C.c_void_p, C.POINTER(MdTileStats)]
        g.md_fetch_calls.argtypes = [C.c_void_p, C.POINTER(MdCall), C.c_uint64, C.POINTER(C.c_uint64)]
        g.md_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        g.md_launch_count.restype = C.c_uint64; g.md_launch_count.argtypes = [C.c_void_p]
        g.md_stream.restype = C.c_void_p; g.md_stream.argtypes = [C.c_void_p]
        g.md_alloc_pinned.restype = C.c_void_p; g.md_alloc_pinned.argtypes = [C.c_size_t]
        g.md_free_pinned.argtypes = [C.c_void_p]
        g.md_host_register.argtypes = [C.c_void_p, C.c_size_t]
        g.md_host_unregister.argtypes = [C.c_void_p]
        g.md_bam_open.restype = C.c_void_p; g.md_bam_open.argtypes = [C.c_void_p, C.c_int32]
        g.md_bam_close.argtypes = [C.c_void_p]; g.md_bam_close.restype = None
        g.md_bam_reset.argtypes = [C.c_void_p]; g.md_bam_reset.restype = None
        g.md_bam_push.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(MdBgzfBlock), C.c_uint32, C.c_uint32, C.POINTER(MdBamSummary)]
        g.md_bam_get_runs.argtypes = [C.c_void_p, C.POINTER(MdBamRun), C.c_uint32]
        g.md_bam_push_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(MdBgzfBlock), C.c_uint32, C.c_uint32]
        g.md_bam_push_end.argtypes = [C.c_void_p, C.POINTER(MdBamSummary)]
        g.md_bam_prefetch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        g.md_bam_extract_run.argtypes = [C.c_void_p, C.c_int, C.POINTER(MdTileDesc), C.c_uint32, C.POINTER(MdCall), C.c_uint64, C.POINTER(MdTileStats)]
        g.md_bam_mbias_run.argtypes = [C.c_void_p, C.c_int, C.POINTER(MdTileDesc), C.c_uint32, C.POINTER(MdTileStats)]
        g.md_bam_tile_shape.argtypes = [C.c_void_p, C.POINTER(MdReadsSoa)]
        g.md_bam_tile_fetch.argtypes = [C.c_void_p, C.POINTER(MdReadsSoa), C.POINTER(C.c_int32)]
        g.md_last_error.restype = C.c_char_p
        g.md_abi_version.restype = C.c_int
        g.md_source_hash.restype = C.c_char_p
        g.md_ctx_totals.argtypes = [C.c_void_p, C.POINTER(MdTotals)]
        g.md_last_totals.argtypes = [C.POINTER(MdTotals)]
        _gpu = g
    return _gpu


_dropin = None


def load_dropin():
    """libMethylDackel.so: extract_main / mbias_main / perRead_main with the reference's signatures, and the native back-end
    table bound to libmdgpu (include/methyldackel.h).  Needs libmdgpu.so and libmdhost.so beside it."""
    global _dropin
    if _dropin is None:
        load_host(); load_gpu()
        p = os.path.join(LIBDIR, "libMethylDackel.so")
        if not os.path.exists(p):
            raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % p)
        d = C.CDLL(p, mode=C.RTLD_GLOBAL)
        d.mdh_gpu_backend.restype = C.POINTER(MdhBackend); d.mdh_gpu_backend.argtypes = [C.c_int]
        for f in (d.extract_main, d.mbias_main, d.perRead_main):
            f.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        _dropin = d
    return _dropin
