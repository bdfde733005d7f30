#!/usr/bin/env python
"""bench.py — M alignments/s of the `MethylDackel extract` / `mbias` pileup hot path on B200.

Contract (see the task brief): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.

`--config c2` (default) — BASELINE.json configs[1], the configuration the metric is quoted on: synthetic 1-contig 10 Mbp,
30x paired-end 150 bp WGBS BAM, CpG-only extract with default options (2.0 M alignments).  A "step" is one pass of the hot
path over that batch.  With N GPUs every rank owns its own 10 Mbp contig interval (weak scaling, no data-path collective).
  * value     : alignments / s with the SoA batch already resident in HBM (prep + count kernels, CUDA-event timed, max over ranks).
  * e2e       : the same work through the C ABI from HOST buffers holding what the reference starts from — the COMPRESSED BAM
                (page-locked): md_bam_push_begin/_end + md_bam_prefetch (H2D, BGZF inflate and record framing on the device) + md_bam_extract_run
                (tile assembly, prep, count) + D2H of the md_call records, every step.  This is the like-for-like counterpart of
                the reference arm (inflate + decode + pileup), minus its text formatting.
  * e2e_soa   : the round-1 `e2e`: pre-decoded SoA tiles in page-locked memory through md_submit_tile()/md_collect_tile().
  * cli_from_bam : the drop-in binary on the same file, process start to finished bedGraph (informational).
  * roofline  : count_warp (the dominant kernel), algorithmic bytes (249 B per 150M alignment + 1 B per reference
                base + 8 B per reported cytosine) / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
  * cpu_baseline : oracle/_ref/MethylDackel — the reference's own C sources compiled on the in-repo htslib shim (oracle/Makefile) —
                with -@ <all cores> and a --chunkSize that gives every core at least 4 chunks (the reference schedules one chunk
                per thread at a time, extract.c:327-350; with the default 1 Mbp chunks a 10 Mbp contig keeps at most 10 threads busy).

`--config c3|c4|c5 [--mbp M]` — the large configurations of BASELINE.json (configs[2..4]) through the drop-in entry point
extract_main()/mbias_main() (lib/libMethylDackel.so, the call a user of the reference's library makes), file in, files out:
  c3: multi-contig genome with human proportions, 30x, `extract --CHG --CHH --mergeContext` (default 300 Mbp = 1/10 scale; --mbp 3000 = full)
  c4: 5 Mbp panel at 2000x, heavy mate overlap, `extract` defaults (66.7 M alignments; --mbp scales the panel length)
  c5: `mbias` on the c3 data set
  value = alignments / (CUDA-event time of the prep + count kernels summed over all tiles of the run: inputs are in HBM when each
  starts); e2e = alignments / wall time of the call; `parity` = byte comparison against the reference build on one contig of the set;
  cpu_baseline as above on the largest contig.  With N GPUs the genome is sharded (api.extract_sharded): strong scaling.

`--impl reference` times the CPU reference instead (rank 0 only) on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "M alignments/sec on extract (CpG, 30x WGBS)"
UNIT = "M alignments/s"
DEPTH = 30
LIB = os.path.join(ROOT, "methyldackel_b200", "lib")
REFBIN = os.path.join(ROOT, "oracle", "_ref", "MethylDackel")
CPU_KIND_NOTE = "reference sources (extract.c, common.c, overlaps.c ...) compiled unmodified on the in-repo htslib shim (oracle/htslib_shim): inflate is zlib, the pileup engine is the shim's"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def ensure_built():
    need = [os.path.join(LIB, x) for x in ("libmdgpu.so", "libmdhost.so", "libMethylDackel.so", "mdsynth")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__ as g
        g.build()


def cache_dir():
    d = os.environ.get("MDBENCH_CACHE", "/tmp/mdbench")
    os.makedirs(d, exist_ok=True)
    return d


def out_dir():
    """where the sub-commands write their text output during timed runs: memory-backed when there is room"""
    d = os.environ.get("MDBENCH_OUT")
    if not d:
        d = "/dev/shm/mdbench_out" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > (24 << 30) else os.path.join(cache_dir(), "out")
    os.makedirs(d, exist_ok=True)
    return d


def synth(prefix, synth_args):
    """deterministic synthetic data set (SURVEY 8d recipe); returns the number of alignments"""
    if not (os.path.exists(prefix + ".bam.bai") and os.path.exists(prefix + ".fa.fai") and os.path.exists(prefix + ".n")):
        t0 = time.time()
        tmp = prefix + ".tmp%d" % os.getpid()
        r = subprocess.run([os.path.join(LIB, "mdsynth"), "--out", tmp] + [str(a) for a in synth_args], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        open(tmp + ".n", "w").write(r.stdout.strip().splitlines()[-1] + "\n")
        for ext in (".fa", ".fa.fai", ".bam", ".bam.bai", ".n"):
            os.replace(tmp + ext, prefix + ext)
        log("[bench] generated %s in %.1f s" % (prefix, time.time() - t0))
    return int(open(prefix + ".n").read().split()[0])


def dataset_c2(rank, mbp):
    prefix = os.path.join(cache_dir(), "c2v2_%dmbp_r%d" % (mbp, rank))
    synth(prefix, ["--contigs", "chr1:%d" % (mbp * 1_000_000), "--depth", DEPTH, "--read-seed", 5678 + rank])
    return prefix


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 6 for i in range(4) if s[2 + i].lower().startswith("active")})
        mx = max([int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()] or [0])
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm)}


def soa_bytes(soa):
    n = soa.n_reads
    return n * (4 + 2 + 1 + 1 + 4 + 4 + 4 + 8) + (n + 1) * 4 + soa.n_cigar_ops * 4 + soa.seq_words * 4 + soa.qual_words * 8


def soa_arrays(soa):
    n = soa.n_reads
    return [(soa.pos, n * 4), (soa.flag, n * 2), (soa.mapq, n), (soa.aux, n), (soa.l_qseq, n * 4), (soa.cigar_off, (n + 1) * 4), (soa.seq_off, n * 4),
            (soa.qual_off, n * 4), (soa.frag_key, n * 8), (soa.cigar, soa.n_cigar_ops * 4), (soa.seq, soa.seq_words * 4), (soa.qual, soa.qual_words * 8)]


def algorithmic_bytes(n_aln, read_len, n_cigar_ops, ref_bases, n_calls):
    """per alignment ceil(L/2) + L + 4*n_cigar + 20; + 1 B per reference base; + 8 B per reported cytosine."""
    return n_aln * ((read_len + 1) // 2 + read_len + 20) + 4 * n_cigar_ops + ref_bases + 8 * n_calls


def peak_hbm():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "fallback 6650 (of fallback)"


def traffic_of(kernel_key):
    """DRAM bytes per launch of a kernel from the committed ncu --set full capture (profiles/r2_traffic.json); None when that
    kernel / configuration has no capture"""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))[kernel_key]
        return int(tj["dram_bytes_read"] + tj["dram_bytes_write"]), tj.get("source")
    except Exception:
        return None, None


def inflate_summary(file_bytes, segs_per_step, infl_ms, inflated_total, comp_total, frame_ms, h2d_ms, hbm_peak_gbs, e2e_ms_per_step):
    """The decode kernel that bounds `e2e` (cumulative CUDA-event times over every push of the stream, warm-up included, per
    segment).  Its algorithmic bytes are compressed bytes in + inflated bytes out; `hbm_frac` only shows how far from memory
    bound it is — the kernel is bound by instruction issue / its dependency chain."""
    if not infl_ms or not segs_per_step:
        return {"kernel": "inflate_kernel", "ms_per_segment": None}
    comp_seg = file_bytes / segs_per_step
    ratio = inflated_total / max(1, comp_total)
    comp_gbs = comp_seg / (infl_ms * 1e-3) / 1e9
    out = {"kernel": "inflate_kernel", "ms_per_segment": round(infl_ms, 3), "compressed_GBps": round(comp_gbs, 2), "inflated_GBps": round(comp_gbs * ratio, 2),
           "segments_per_step": segs_per_step, "frame_ms_per_segment": round(frame_ms, 3), "h2d_ms_per_segment": round(h2d_ms, 3),
           "bound": "instruction issue / dependency chain (not hbm)", "algorithmic_GBps": round(comp_gbs * (1.0 + ratio), 1)}
    if hbm_peak_gbs:
        out["hbm_frac"] = round(comp_gbs * (1.0 + ratio) / hbm_peak_gbs, 4)
    if e2e_ms_per_step:
        out["share_of_e2e_step"] = round(infl_ms * segs_per_step / e2e_ms_per_step, 3)
    return out


# ------------------------------------------------------------------------------------------------ CPU reference arm
def chunks_for(cores, region_bp):
    """--chunkSize giving every core at least 4 chunks (None: the default 1 Mbp already does)"""
    want = 4 * cores
    if region_bp // 1_000_000 >= want:
        return None
    return max(10_000, region_bp // want)


def time_reference(sub, opts, prefix, region, cores, chunk, outp):
    """one run of the CPU reference on `region` with `cores` threads; returns seconds"""
    cmd = [REFBIN, sub, "-@", str(cores)] + list(opts) + (["--chunkSize", str(chunk)] if chunk else []) + (["-r", region] if region else [])
    cmd += [prefix + ".fa", prefix + ".bam"] if sub == "extract" else ["--noSVG", prefix + ".fa", prefix + ".bam"]
    if sub == "extract":
        cmd += ["-o", outp]
    t0 = time.perf_counter()
    with open(outp + ".stdout", "w") as so:
        subprocess.run(cmd, check=True, stdout=so, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t0


def count_region(prefix, region_name, beg, end):
    """alignments the reference's index query returns for contig[beg,end): starting before `end` and reaching beyond `beg`
    (counted with the repository's own BAM reader, outside any timed region)"""
    from methyldackel_b200 import api
    if end - beg > 40_000_000:                                          # too large to decode for a count: the synthetic coverage is uniform by construction
        total = int(open(prefix + ".n").read().split()[0])
        genome = sum(int(l.split("\t")[1]) for l in open(prefix + ".fa.fai"))
        return int(round(total * (end - beg) / float(genome)))
    b = api.BamFile(prefix + ".bam")
    tid = b.names.index(region_name)
    soa = b.read_region(tid, beg, end)
    n = soa.n_reads
    b.close()
    return n


def reference_sample(sub, opts, prefix, contig, contig_len, cores, target_s):
    """bounded sample: a prefix of `contig` sized for about target_s seconds of the reference at full thread count"""
    probe = min(2_000_000, contig_len)
    outp = os.path.join(out_dir(), "ref_probe")
    t = time_reference(sub, opts, prefix, "%s:1-%d" % (contig, probe), cores, chunks_for(cores, probe), outp)
    want = int(min(contig_len, max(probe, probe * target_s / max(t, 1e-3))))
    want = max(1_000_000, (want // 1_000_000) * 1_000_000) if contig_len >= 1_000_000 else contig_len
    return min(want, contig_len)


def reference_arm(args, sub, opts, prefix, contig, contig_len, workload, cores, target_s=8.0):
    region_bp = reference_sample(sub, opts, prefix, contig, contig_len, cores, target_s)
    chunk = chunks_for(cores, region_bp)
    region = "%s:1-%d" % (contig, region_bp)
    nreads = count_region(prefix, contig, 0, region_bp)
    outp = os.path.join(out_dir(), "ref_out")
    for _ in range(args.warmup):
        time_reference(sub, opts, prefix, region, cores, chunk, outp)
    ts = [time_reference(sub, opts, prefix, region, cores, chunk, outp) for _ in range(args.steps)]
    v = nreads * args.steps / sum(ts) / 1e6
    sample = "oracle/_ref/MethylDackel %s -@ %d %s%s -r %s of the workload BAM (%d alignments per step, %d chunks; BGZF inflate + pileup + text output included)" % (
        sub, cores, " ".join(opts) + " " if opts else "", "--chunkSize %d" % chunk if chunk else "default chunks", region, nreads,
        (region_bp + (chunk or 1_000_000) - 1) // (chunk or 1_000_000))
    return {"impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1e3 * sum(ts) / args.steps, 3), "higher_is_better": True, "scaling": "weak" if args.config == "c2" else "strong", "vs_baseline": None,
            "dtype": "u8/u32 integer", "data": "synthetic", "config": {"workload": workload, "sample": sample},
            "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "reference", "kind_note": CPU_KIND_NOTE, "sample": sample},
            "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}


def cpu_baseline(sub, opts, prefix, contig, contig_len, cores, target_s=12.0):
    region_bp = reference_sample(sub, opts, prefix, contig, contig_len, cores, target_s)
    chunk = chunks_for(cores, region_bp)
    region = "%s:1-%d" % (contig, region_bp)
    outp = os.path.join(out_dir(), "ref_out")
    tsec = time_reference(sub, opts, prefix, region, cores, chunk, outp)
    nr = count_region(prefix, contig, 0, region_bp)
    return {"value": round(nr / tsec / 1e6, 4), "unit": UNIT, "cores": cores, "kind": "reference", "kind_note": CPU_KIND_NOTE,
            "sample": "oracle/_ref/MethylDackel %s -@ %d %s%s -r %s of the workload BAM: %d alignments in %.2f s (BGZF inflate + pileup + text output included)" % (
                sub, cores, " ".join(opts) + " " if opts else "", "--chunkSize %d" % chunk if chunk else "default chunks", region, nr, tsec)}


# ------------------------------------------------------------------------------------------------ process group helpers
class Group:
    def __init__(self, world, local_rank):
        import torch
        self.torch = torch
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
        torch.cuda.set_device(local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _red(self, x, op):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def maxr(self, x):
        return self._red(x, self.dist.ReduceOp.MAX if self.dist else None)

    def sumr(self, x):
        return self._red(x, self.dist.ReduceOp.SUM if self.dist else None)

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ c2: the headline configuration
def run_c2(args, rank, local_rank, world, cores):
    from methyldackel_b200 import _abi as A
    from methyldackel_b200 import api
    contig_len = args.mbp * 1_000_000
    workload = "synthetic 1-contig %d Mbp, %dx PE150 WGBS BAM, CpG-only extract (BASELINE.json configs[1])" % (args.mbp, DEPTH)
    if args.impl == "reference":
        if rank != 0:
            return 0
        prefix = dataset_c2(0, args.mbp)
        print(json.dumps(reference_arm(args, "extract", [], prefix, "chr1", contig_len, workload, cores)))
        return 0

    grp = Group(world, local_rank)
    prefix = dataset_c2(rank, args.mbp)
    b = api.BamFile(prefix + ".bam")
    ref = api.fetch_contig(prefix + ".fa", "chr1")
    t0 = time.time()
    soa = b.read_region(0)
    log("[bench] rank %d decoded %d alignments to SoA in %.1f s (%.1f MB)" % (rank, soa.n_reads, time.time() - t0, soa_bytes(soa) / 1e6))
    n = soa.n_reads
    cfg = A.default_config()
    g = api.GpuContext(cfg, device=local_rank)
    g.load_contig(0, ref)
    reflen = len(ref)
    warm = max(args.warmup, 3)

    # ------------------------------------------------------------ kernel-only: batch resident in HBM
    d = g.upload(soa)
    for _ in range(warm):
        st = g.extract_tile_device(0, 0, reflen, d)
    sampler = ClockSampler(local_rank); sampler.start()
    grp.barrier()
    l0 = g.launch_count()
    ev_total = ev_count = ev_prep = 0.0
    w0 = time.perf_counter()
    for _ in range(args.steps):
        st = g.extract_tile_device(0, 0, reflen, d)
        t = g.last_timing()
        ev_prep += t[1]; ev_count += t[2]; ev_total += t[1] + t[2]
    grp.barrier()
    wall_dev = time.perf_counter() - w0
    launches = g.launch_count() - l0
    ms_step = grp.maxr(ev_total / args.steps)
    total_aln = grp.sumr(float(n))
    value = total_aln / (ms_step * 1e-3) / 1e6
    g.free(d)

    # ------------------------------------------------------------ e2e_soa: pre-decoded SoA tiles from page-locked host memory
    # The batch is cut into tiles exactly as the sub-command driver cuts it (Tiler, 2^17 alignments per tile) and streamed with
    # md_submit_tile / md_collect_tile over the context's lanes.  Every byte crosses PCIe every step.
    tiles = b.make_tiles(0, 0, reflen, 1 << 17)
    pinned = []
    h2d_bytes = 0
    for td_k, soa_k in tiles:
        h2d_bytes += soa_bytes(soa_k)
        for ptr, nbytes in soa_arrays(soa_k):
            addr = C.cast(ptr, C.c_void_p).value
            if addr and nbytes and g.g.md_host_register(addr, nbytes) == 0:
                pinned.append(addr)
    cap = reflen + 16
    calls = (A.MdCall * cap)()
    g.g.md_host_register(C.addressof(calls), C.sizeof(calls))
    stt = A.MdTileStats()
    NL = 3

    def soa_step():
        out_off = 0
        inflight = []
        for k, (td_k, soa_k) in enumerate(tiles):
            if len(inflight) == NL:
                t_id = inflight.pop(0)
                dst = C.cast(C.addressof(calls) + out_off * 16, C.POINTER(A.MdCall))
                assert g.g.md_collect_tile(g.h, t_id, dst, cap - out_off, C.byref(stt)) == 0, g.g.md_last_error()
                out_off += stt.n_calls
            t_id = g.g.md_submit_tile(g.h, C.byref(td_k), C.byref(soa_k))
            assert t_id >= 0, g.g.md_last_error()
            inflight.append(t_id)
        for t_id in inflight:
            dst = C.cast(C.addressof(calls) + out_off * 16, C.POINTER(A.MdCall))
            assert g.g.md_collect_tile(g.h, t_id, dst, cap - out_off, C.byref(stt)) == 0, g.g.md_last_error()
            out_off += stt.n_calls
        return out_off

    for _ in range(warm):
        n_soa_calls = soa_step()
    assert n_soa_calls == st.n_calls, (n_soa_calls, st.n_calls)     # tiled + pipelined path reports the same columns
    grp.barrier()
    l1 = g.launch_count()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        soa_step()
    grp.barrier()
    soa_s = grp.maxr((time.perf_counter() - w0) / args.steps)
    launches += g.launch_count() - l1
    soa_t = g.last_timing()

    # ------------------------------------------------------------ e2e: from the COMPRESSED BAM in page-locked host memory
    import struct
    raw = open(prefix + ".bam", "rb").read()
    hbuf = g.g.md_alloc_pinned(len(raw) + 64)
    C.memmove(hbuf, raw, len(raw))
    segs, cur_blocks, seg_start, off = [], [], 0, 0
    SEG = int(os.environ.get("MD_SEGMENT_BYTES", 96 << 20))      # the sub-command driver's default segment size (host/cli.cpp: device_segment_bytes)
    while off + 18 <= len(raw):
        xlen = struct.unpack_from("<H", raw, off + 10)[0]
        bs = struct.unpack_from("<H", raw, off + 16)[0] + 1          # BC subfield first, as every BGZF writer lays it out (checked below)
        assert raw[off + 12] == 66 and raw[off + 13] == 67
        cur_blocks.append((off - seg_start + 12 + xlen, bs - 12 - xlen - 8, struct.unpack_from("<I", raw, off + bs - 4)[0]))
        off += bs
        if off - seg_start >= SEG:
            segs.append((seg_start, off - seg_start, cur_blocks)); cur_blocks, seg_start = [], off
    if cur_blocks:
        segs.append((seg_start, off - seg_start, cur_blocks))
    seg_arr = []
    for (so, sl, bl) in segs:
        arr = (A.MdBgzfBlock * len(bl))()
        for k, (a_, b_, c_) in enumerate(bl):
            arr[k].comp_off, arr[k].comp_len, arr[k].isize = a_, b_, c_
        seg_arr.append((so, sl, arr, len(bl)))
    import gzip as _gz, io as _io
    u0 = _gz.GzipFile(fileobj=_io.BytesIO(raw[:1 << 20])).read(1 << 16)
    l_text = struct.unpack_from("<i", u0, 4)[0]; hoff = 8 + l_text
    n_ref = struct.unpack_from("<i", u0, hoff)[0]; hoff += 4
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", u0, hoff)[0]; hoff += 4 + l_name + 4
    bs_ = g.g.md_bam_open(g.h, n_ref)
    assert bs_, g.g.md_last_error()
    summ = A.MdBamSummary(); runs1 = (A.MdBamRun * 8)()

    PREFETCH = not os.environ.get("MD_NO_PREFETCH")

    def bam_step():
        g.g.md_bam_reset(bs_)
        so, sl, arr, nb = seg_arr[0]
        assert g.g.md_bam_push_begin(bs_, hbuf + so, sl, arr, nb, hoff) == 0, g.g.md_last_error()
        if len(seg_arr) > 1 and PREFETCH:                               # the next segment's bytes travel while this one is inflated
            assert g.g.md_bam_prefetch(bs_, hbuf + seg_arr[1][0], seg_arr[1][1]) == 0, g.g.md_last_error()
        out_off, open_beg = 0, 0
        for k in range(len(seg_arr)):
            assert g.g.md_bam_push_end(bs_, C.byref(summ)) == 0, g.g.md_last_error()
            if k + 1 < len(seg_arr):
                so, sl, arr, nb = seg_arr[k + 1]
                assert g.g.md_bam_push_begin(bs_, hbuf + so, sl, arr, nb, 0) == 0, g.g.md_last_error()
                if k + 2 < len(seg_arr) and PREFETCH:
                    assert g.g.md_bam_prefetch(bs_, hbuf + seg_arr[k + 2][0], seg_arr[k + 2][1]) == 0, g.g.md_last_error()
            nr = g.g.md_bam_get_runs(bs_, runs1, 8)
            assert nr == 1 and runs1[0].tid == 0
            cut = reflen if k + 1 == len(seg_arr) else max(runs1[0].last_pos, open_beg)
            td = A.MdTileDesc(0, open_beg, cut, 0, 0)
            dst = C.cast(C.addressof(calls) + out_off * 16, C.POINTER(A.MdCall))
            assert g.g.md_bam_extract_run(bs_, 0, C.byref(td), 0xffffffff, dst, cap - out_off, C.byref(stt)) == 0, g.g.md_last_error()
            out_off += stt.n_calls; open_beg = cut
        return out_off

    for _ in range(warm):
        n_bam_calls = bam_step()
    assert n_bam_calls == st.n_calls, (n_bam_calls, st.n_calls)       # same columns as the SoA paths
    tot0 = A.MdTotals(); g.g.md_ctx_totals(g.h, C.byref(tot0))
    grp.barrier()
    l2 = g.launch_count()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        bam_step()
    grp.barrier()
    bam_s = grp.maxr((time.perf_counter() - w0) / args.steps)
    bam_launches = g.launch_count() - l2
    launches += bam_launches
    tot1 = A.MdTotals(); g.g.md_ctx_totals(g.h, C.byref(tot1))
    g.g.md_bam_close(bs_)
    tot2 = A.MdTotals(); g.g.md_ctx_totals(g.h, C.byref(tot2))
    g.g.md_free_pinned(hbuf)
    sampler.stop_flag = True; sampler.join(timeout=2)
    g.g.md_host_unregister(C.addressof(calls))
    for addr in pinned:
        g.g.md_host_unregister(addr)

    if rank != 0:
        g.close()
        grp.close()
        return 0

    # ------------------------------------------------------------ roofline of the dominant kernel
    peak, peak_src = peak_hbm()
    alg = algorithmic_bytes(n, 150, soa.n_cigar_ops, reflen, st.n_calls)
    count_ms = ev_count / args.steps
    achieved = alg / (count_ms * 1e-3) / 1e9
    traffic, traffic_src = traffic_of("count_warp")
    roofline = {"bound": "hbm", "kernel": "count_warp<0, 2>", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg), "kernel_ms": round(count_ms, 4), "prep_pair_window_ms": round(ev_prep / args.steps, 4),
                "traffic": traffic, "traffic_source": traffic_src}
    # the decode kernel that bounds `e2e`: compressed bytes in + inflated bytes out per segment, against the same peak
    nseg = max(1, len(seg_arr) * (args.steps + warm))
    inflate = inflate_summary(len(raw), len(seg_arr), tot2.inflate_ms / nseg, tot2.inflated_bytes, tot2.comp_bytes, tot2.frame_ms / nseg, tot2.push_h2d_ms / nseg, peak, bam_s * 1e3)

    # ------------------------------------------------------------ the drop-in binary from the BAM file (informational)
    cli = None
    if world == 1:
        binp = os.path.join(LIB, "MethylDackel")
        outp = os.path.join(out_dir(), "cli_out")
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            subprocess.run([binp, "extract", prefix + ".fa", prefix + ".bam", "-o", outp], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            ts.append(time.perf_counter() - t0)
        cli = {"value": round(n / min(ts) / 1e6, 3), "unit": UNIT, "seconds": round(min(ts), 3),
               "what": "lib/MethylDackel extract <fa> <bam> (process start, CUDA context, device-side BGZF inflate + decode, kernels, D2H, text output), best of 3"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline and os.path.exists(REFBIN):
        cpu = cpu_baseline("extract", [], prefix, "chr1", contig_len, cores)

    out = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
           "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32 integer", "data": "synthetic",
           "config": {"workload": workload, "alignments_per_gpu": n, "contig_bp": reflen, "soa_phred_bits": int(soa.qual_bits) or 8, "options": "extract defaults (-q 10 -p 5 -F 0xF00, CpG only)",
                      "parallelism": "contig interval per GPU, no collective", "l2": "inputs (%.0f MB SoA / %.0f MB compressed BAM per GPU) larger than the 126 MB L2" % (soa_bytes(soa) / 1e6, len(raw) / 1e6)},
           "e2e": {"value": round(total_aln / bam_s / 1e6, 3), "unit": UNIT, "ms_per_step": round(bam_s * 1e3, 3), "h2d_bytes_per_step": len(raw), "d2h_bytes_per_step": int(n_bam_calls * 16),
                   "segments_per_step": len(seg_arr), "launches_per_step": int(bam_launches // max(args.steps, 1)),
                   "path": "compressed BAM bytes (page-locked) -> md_bam_push_begin/_end + md_bam_prefetch (H2D of segment k+2 under the inflate of k+1; BGZF inflate, record framing on the device; segment k+1 overlaps the tiles of segment k) -> md_bam_extract_run (tile assembly in HBM + prep + count) -> D2H md_call records"},
           "e2e_soa": {"value": round(total_aln / soa_s / 1e6, 3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(n_soa_calls * 16 + 32 * len(tiles)),
                       "ms_per_step": round(soa_s * 1e3, 3), "tiles_per_step": len(tiles), "lanes": NL,
                       "last_tile_ms": {"h2d": round(soa_t[0], 3), "prep": round(soa_t[1], 3), "count": round(soa_t[2], 3), "d2h": round(soa_t[3], 3)},
                       "path": "md_submit_tile()/md_collect_tile(): page-locked PRE-DECODED SoA tiles -> H2D -> kernels -> D2H md_call records, 3 lanes in flight (skips inflate + record decode: not like-for-like with the reference arm)"},
           "gpu_launches": int(launches), "wall_ms_per_step_device_resident": round(1e3 * wall_dev / args.steps, 3),
           "roofline": roofline, "inflate": inflate, "clocks": sampler.summary(), "calls_per_step": int(st.n_calls), "pairs_per_step": int(st.n_pairs)}
    if cli is not None:
        out["cli_from_bam"] = cli
    if cpu is not None:
        out["cpu_baseline"] = cpu
    print(json.dumps(out))
    g.close()
    grp.close()
    return 0


# ------------------------------------------------------------------------------------------------ c3 / c4 / c5: the large configurations
def big_spec(config, mbp):
    if config in ("c3", "c5"):
        mbp = mbp or 300
        name = "h%d" % mbp
        synth_args = ["--human", mbp * 1_000_000, "--depth", DEPTH, "--read-seed", 77]
        what = "synthetic %d Mbp genome (24 contigs, human proportions), %dx PE150 WGBS" % (mbp, DEPTH)
        if config == "c3":
            return name, synth_args, "extract", ["--CHG", "--CHH", "--mergeContext"], what + ", extract --CHG --CHH --mergeContext (BASELINE.json configs[2]%s)" % ("" if mbp >= 3000 else " at %d/3000 scale" % mbp)
        return name, synth_args, "mbias", [], what + ", mbias (BASELINE.json configs[4]%s)" % ("" if mbp >= 3000 else " at %d/3000 scale" % mbp)
    kbp = int((mbp or 5) * 1000)
    name = "panel%dk" % kbp
    synth_args = ["--contigs", "panel:%d" % (kbp * 1000), "--depth", 2000, "--isize-mean", 180, "--isize-sd", 25, "--isize-min", 150, "--isize-max", 300, "--read-seed", 4242]
    return name, synth_args, "extract", [], "synthetic targeted panel %.1f Mbp at 2000x, PE150, insert 180+-25 (heavy mate overlap), extract defaults (BASELINE.json configs[3]%s)" % (
        kbp / 1000.0, "" if kbp >= 5000 else " at %d/5000 scale" % kbp)


def fai(prefix):
    return [(l.split("\t")[0], int(l.split("\t")[1])) for l in open(prefix + ".fa.fai")]


def run_big(args, rank, local_rank, world, cores):
    from methyldackel_b200 import api
    name, synth_args, sub, opts, workload = big_spec(args.config, args.mbp)
    prefix = os.path.join(cache_dir(), name)
    if rank == 0:
        n_aln = synth(prefix, synth_args)
    contigs = None
    if args.impl == "reference":
        if rank != 0:
            return 0
        contigs = fai(prefix)
        big = max(contigs, key=lambda c: c[1])
        print(json.dumps(reference_arm(args, sub, opts, prefix, big[0], big[1], workload, cores, target_s=20.0)))
        return 0

    grp = Group(world, local_rank)
    grp.barrier()                                                       # rank 0 has generated the set
    n_aln = int(open(prefix + ".n").read().split()[0])
    contigs = fai(prefix)
    genome_bp = sum(c[1] for c in contigs)
    outp = os.path.join(out_dir(), "%s_%s_w%d" % (name, args.config, world))
    argv = list(opts) + [prefix + ".fa", prefix + ".bam"] + (["-o", outp] if sub == "extract" else ["--noSVG"])
    steps, warm = args.steps, args.warmup

    def one_step():
        """the public call: extract_main / mbias_main (lib/libMethylDackel.so) — sharded over the ranks when there are several"""
        so = os.dup(1)
        fd = os.open(outp + ".stdout.%d" % rank, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
        sys.stdout.flush(); os.dup2(fd, 1)                              # mbias prints its table to stdout
        try:
            if world == 1:
                rc = (api.extract_main if sub == "extract" else api.mbias_main)(argv, device=local_rank)[0]
            elif sub == "extract":
                rc = api.extract_sharded(argv, rank, world, run_main=lambda av: api.extract_main(av, device=local_rank)[0])
            else:
                rc = api.mbias_sharded(argv, rank, world, outp + ".hist", run_main=lambda av: api.mbias_main(av, device=local_rank)[0])
        finally:
            sys.stdout.flush(); os.dup2(so, 1); os.close(fd); os.close(so)
        assert rc == 0, "sub-command failed with %d" % rc
        return api.last_totals()

    for _ in range(warm):
        one_step()
    sampler = ClockSampler(local_rank); sampler.start()
    grp.barrier()
    w0 = time.perf_counter()
    k_ms = c_ms = p_ms = infl_ms = 0.0
    launches = 0
    tot = None
    for _ in range(steps):
        tot = one_step()
        k_ms += tot.prep_ms + tot.count_ms; c_ms += tot.count_ms; p_ms += tot.prep_ms; infl_ms += tot.inflate_ms; launches += tot.launches
    grp.barrier()
    wall = grp.maxr((time.perf_counter() - w0) / steps)
    sampler.stop_flag = True; sampler.join(timeout=2)
    k_ms_step = grp.maxr(k_ms / steps)                                  # ranks run concurrently: the slowest rank's kernel time
    launches = int(grp.sumr(float(launches)))
    tiles_aln = grp.sumr(float(tot.alignments))                        # alignments handed to the kernels (straddling reads once per tile)
    calls = grp.sumr(float(tot.calls)); cig = grp.sumr(float(tot.cigar_ops))
    comp = grp.sumr(float(tot.comp_bytes)); infl_b = grp.sumr(float(tot.inflated_bytes))
    c_ms_max = grp.maxr(c_ms / steps); infl_ms_max = grp.maxr(infl_ms / steps); p_ms_max = grp.maxr(p_ms / steps)
    h2d_rate = tot.comp_bytes / (tot.push_h2d_ms * 1e-3) / 1e9 if tot.push_h2d_ms else 0.0      # this rank's segment copies (page-locked staging -> HBM)
    h2d_min = -grp.maxr(-h2d_rate); h2d_max = grp.maxr(h2d_rate)
    if rank != 0:
        grp.close()
        return 0

    value = n_aln / (k_ms_step * 1e-3) / 1e6
    peak, peak_src = peak_hbm()
    alg = algorithmic_bytes(tiles_aln, 150, cig, genome_bp, calls)
    achieved = alg / (c_ms_max * 1e-3) / 1e9 if c_ms_max else 0.0
    kern = "count_warp<2, 2>" if sub == "mbias" else "count_warp<0, 2>"
    traffic, traffic_src = traffic_of("count_warp_" + args.config)
    roofline = {"bound": "hbm", "kernel": kern, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "peak_source": peak_src,
                "algorithmic_bytes_per_step": int(alg), "kernel_ms_per_step": round(c_ms_max, 3), "launches_of_kernel_per_step": int(tot.tiles), "prep_ms_per_step": round(p_ms_max, 3),
                "traffic": traffic, "traffic_source": traffic_src,
                "note": "summed over the %d tiles of one pass (CUDA events around every tile's kernels); traffic is per launch of a typical tile" % tot.tiles}
    h2d = {"GBps_slowest_rank": round(h2d_min, 1), "GBps_fastest_rank": round(h2d_max, 1), "what": "compressed segments, page-locked staging buffers -> HBM, CUDA events on the decode stream"}
    inflate = {"kernel": "inflate_kernel", "ms_per_step": round(infl_ms_max, 2), "compressed_GBps": round(comp / world / (infl_ms_max * 1e-3) / 1e9, 2) if infl_ms_max else None,
               "inflated_GBps": round(infl_b / world / (infl_ms_max * 1e-3) / 1e9, 2) if infl_ms_max else None}

    # ------------------------------------------------------------ parity on one contig against the reference build, and the CPU baseline
    parity, cpu = None, None
    if os.path.exists(REFBIN) and not args.no_cpu_baseline:
        small = min(contigs, key=lambda c: c[1])
        pr, pn = os.path.join(out_dir(), "par_ref"), os.path.join(out_dir(), "par_new")
        region = small[0]
        if sub == "extract":
            subprocess.run([REFBIN, "extract", "-@", str(cores)] + opts + ["-r", region, "-o", pr, prefix + ".fa", prefix + ".bam"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            assert api.extract_main(opts + ["-r", region, "-o", pn, prefix + ".fa", prefix + ".bam"], device=local_rank)[0] == 0
            same, nbytes = True, 0
            for ctx in ("CpG", "CHG", "CHH"):
                fr, fn = "%s_%s.bedGraph" % (pr, ctx), "%s_%s.bedGraph" % (pn, ctx)
                if os.path.exists(fr) or os.path.exists(fn):
                    a = open(fr, "rb").read().replace(pr.encode(), b"P"); bb = open(fn, "rb").read().replace(pn.encode(), b"P")
                    same = same and a == bb; nbytes += len(a)
            parity = {"against": "oracle/_ref/MethylDackel extract %s -r %s" % (" ".join(opts), region), "identical": bool(same), "bytes_compared": nbytes}
        else:
            r = subprocess.run([REFBIN, "mbias", "-@", str(cores), "--txt", "-r", region, prefix + ".fa", prefix + ".bam", pr], check=True, capture_output=True, text=True)
            n = subprocess.run([os.path.join(LIB, "MethylDackel"), "mbias", "--txt", "-r", region, prefix + ".fa", prefix + ".bam", pn], check=True, capture_output=True, text=True)
            sug = lambda e: [l for l in e.splitlines() if l.startswith("Suggested")]  # noqa: E731
            svgs = all(open("%s_%s.svg" % (pr, s_)).read() == open("%s_%s.svg" % (pn, s_)).read() for s_ in ("OT", "OB"))
            parity = {"against": "oracle/_ref/MethylDackel mbias --txt -r %s <prefix>" % region, "identical": bool(r.stdout == n.stdout and sug(r.stderr) == sug(n.stderr) and svgs),
                      "bytes_compared": len(r.stdout), "svg_identical": bool(svgs)}
        big = max(contigs, key=lambda c: c[1])
        cpu = cpu_baseline(sub, opts, prefix, big[0], big[1], cores, target_s=20.0)

    # the drop-in BINARY from a cold process (CUDA context creation, first-touch of every buffer included), best of 2
    cli = None
    if world == 1:
        ts = []
        for _ in range(2):
            t0 = time.perf_counter()
            with open(outp + ".cli.stdout", "w") as so:
                subprocess.run([os.path.join(LIB, "MethylDackel"), sub] + list(opts) + [prefix + ".fa", prefix + ".bam"] + (["-o", outp] if sub == "extract" else ["--noSVG"]),   # same files as the library call: one copy of the output at a time
                               check=True, stdout=so, stderr=subprocess.DEVNULL)
            ts.append(time.perf_counter() - t0)
        cli = {"value": round(n_aln / min(ts) / 1e6, 3), "unit": UNIT, "seconds": round(min(ts), 3), "what": "lib/MethylDackel %s ... as a fresh process, best of 2" % sub}

    bam_bytes = os.path.getsize(prefix + ".bam")
    out = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": round(k_ms_step, 3),
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8/u32 integer", "data": "synthetic",
           "config": {"workload": workload, "alignments": n_aln, "genome_bp": genome_bp, "contigs": len(contigs), "options": " ".join([sub] + opts),
                      "parallelism": "contiguous runs of reference chunks per GPU, no collective (api.%s_sharded)" % sub if world > 1 else "single GPU",
                      "l2": "inputs (%.1f GB compressed BAM) far larger than the 126 MB L2" % (bam_bytes / 1e9)},
           "e2e": {"value": round(n_aln / wall / 1e6, 3), "unit": UNIT, "seconds_per_step": round(wall, 3), "h2d_bytes_per_step": int(bam_bytes + genome_bp),
                   "d2h_bytes_per_step": int(calls * 16) if sub == "extract" else 4 * 2 * 1024 * 2 * 4,
                   "path": "%s_main(argv) of lib/libMethylDackel.so: BAM + FASTA files in -> device-side BGZF inflate + decode -> prep/count kernels -> %s" % (
                       sub, "md_call records -> host formatter -> bedGraph files (in %s)" % out_dir() if sub == "extract" else "histogram -> --txt table")},
           "gpu_launches": launches, "roofline": roofline, "inflate": inflate, "segment_h2d": h2d, "clocks": sampler.summary(), "calls_per_step": int(calls), "tiles_per_step": int(tot.tiles)}
    if cli is not None:
        out["cli_from_bam"] = cli
    if parity is not None:
        out["parity"] = parity
    if cpu is not None:
        out["cpu_baseline"] = cpu
    print(json.dumps(out))
    grp.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"], help="c2 = BASELINE.json configs[1] (default, the headline); c3/c4/c5 = configs[2]/[3]/[4]")
    ap.add_argument("--mbp", type=float, default=None, help="genome size in Mbp (c2: 10; c3/c5: 300, 3000 = full size; c4: 5)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    ensure_built()
    cores = os.cpu_count() or 1
    if args.config == "c2":
        args.steps = 10 if args.steps is None else args.steps
        args.warmup = 3 if args.warmup is None else args.warmup
        args.mbp = int(args.mbp or 10)
        return run_c2(args, rank, local_rank, world, cores)
    args.steps = 2 if args.steps is None else args.steps
    args.warmup = 3 if args.warmup is None else args.warmup
    if args.mbp is not None and args.config != "c4":
        args.mbp = int(args.mbp)
    return run_big(args, rank, local_rank, world, cores)


if __name__ == "__main__":
    sys.exit(main())
