#!/usr/bin/env python
"""bench.py — M alignments/s of the `MethylDackel extract` pileup hot path on B200.

Contract (see the task brief): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.
  * workload  : BASELINE.json configs[1] — synthetic 1-contig 10 Mbp, 30x paired-end 150 bp WGBS BAM,
                CpG-only extract with default options (2.0 M alignments).  A "step" is one pass of the hot
                path (prep + pair + window + count kernels) over that batch.  With N GPUs every rank owns
                its own 10 Mbp contig interval (weak scaling, no collective on the data path).
  * value     : alignments / s with the SoA batch already resident in HBM (CUDA-event timed, max over ranks).
  * e2e       : the same through md_extract_tile() with HOST (page-locked) SoA buffers: H2D copy of the batch,
                kernels, D2H of the compact call records, every step.
  * roofline  : count_warp (the dominant kernel), algorithmic bytes (249 B per 150M alignment + 1 B per reference
                base + 8 B per reported cytosine) / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
  * cpu_baseline : oracle/_ref/MethylDackel (the reference's own C sources, see oracle/Makefile) with -@ <all cores>
                on a bounded region of the same BAM.
`--impl reference` times that CPU reference instead (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "M alignments/sec on extract (CpG, 30x WGBS)"
UNIT = "M alignments/s"
CONTIG_LEN = 10_000_000
DEPTH = 30


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def ensure_built():
    lib = os.path.join(ROOT, "methyldackel_b200", "lib")
    need = [os.path.join(lib, x) for x in ("libmdgpu.so", "libmdhost.so", "mdsynth")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__ as g
        g.build()


def dataset(rank, mbp):
    """Deterministic synthetic data set of SURVEY 8d (genome seed 1234, read seed 5678 + rank)."""
    cache = os.environ.get("MDBENCH_CACHE", "/tmp/mdbench")
    os.makedirs(cache, exist_ok=True)
    prefix = os.path.join(cache, "c2_%dmbp_r%d" % (mbp, rank))
    if not (os.path.exists(prefix + ".bam.bai") and os.path.exists(prefix + ".fa.fai")):
        t0 = time.time()
        tmp = prefix + ".tmp%d" % os.getpid()
        subprocess.run([os.path.join(ROOT, "methyldackel_b200", "lib", "mdsynth"), "--out", tmp, "--contigs", "chr1:%d" % (mbp * 1_000_000),
                        "--depth", str(DEPTH), "--read-seed", str(5678 + rank)], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for ext in (".fa", ".fa.fai", ".bam", ".bam.bai"):
            os.replace(tmp + ext, prefix + ext)
        log("[bench] generated %s in %.1f s" % (prefix, time.time() - t0))
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


def algorithmic_bytes(soa, reflen, n_calls):
    """per alignment ceil(L/2) + L + 4*n_cigar + 20; + 1 B per reference base; + 8 B per reported cytosine."""
    n = soa.n_reads
    lq = sum(soa.l_qseq[i] for i in range(0, n, max(1, n // 4096)))  # sampled mean read length (all reads are 150 bp here)
    cnt = len(range(0, n, max(1, n // 4096)))
    mean_l = lq / cnt
    per_aln = (mean_l + 1) // 2 + mean_l + 20
    return n * per_aln + 4 * soa.n_cigar_ops + reflen + 8 * n_calls


def time_reference(prefix, region_bp, cores, outdir):
    """One run of the CPU reference on contig[0:region_bp] with `cores` threads; returns seconds."""
    refbin = os.path.join(ROOT, "oracle", "_ref", "MethylDackel")
    out = os.path.join(outdir, "ref_out")
    t0 = time.perf_counter()
    subprocess.run([refbin, "extract", "-@", str(cores), "-r", "chr1:1-%d" % region_bp, "-o", out, prefix + ".fa", prefix + ".bam"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t0


def reads_in_region(soa, region_bp):
    lo, hi = 0, soa.n_reads
    while lo < hi:
        mid = (lo + hi) // 2
        if soa.pos[mid] < region_bp:
            lo = mid + 1
        else:
            hi = mid
    return lo


def pick_reference_sample(prefix, soa, cores, outdir, target_s, contig_len=CONTIG_LEN):
    """Bounded CPU sample: probe 1 Mbp, then size the region for ~target_s seconds (at most the whole contig)."""
    probe = min(1_000_000, contig_len)
    t = time_reference(prefix, probe, cores, outdir)
    rate = reads_in_region(soa, probe) / max(t, 1e-3)
    want = int(min(contig_len, max(probe, probe * target_s / max(t, 1e-3))))
    want = max(1_000_000, (want // 1_000_000) * 1_000_000)
    return want, rate


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mbp", type=int, default=CONTIG_LEN // 1_000_000, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    ensure_built()
    from methyldackel_b200 import _abi as A
    from methyldackel_b200 import api
    cores = os.cpu_count() or 1
    tmpdir = os.environ.get("MDBENCH_CACHE", "/tmp/mdbench")
    os.makedirs(tmpdir, exist_ok=True)
    workload = "synthetic 1-contig %d Mbp, %dx PE150 WGBS BAM, CpG-only extract (BASELINE.json configs[1])" % (args.mbp, DEPTH)

    if args.impl == "reference":
        if rank != 0:
            return 0
        prefix = dataset(0, args.mbp)
        b = api.BamFile(prefix + ".bam")
        soa = b.read_region(0)
        region, _ = pick_reference_sample(prefix, soa, cores, tmpdir, 6.0, args.mbp * 1_000_000)
        nreads = reads_in_region(soa, region)
        for _ in range(args.warmup):
            time_reference(prefix, region, cores, tmpdir)
        ts = [time_reference(prefix, region, cores, tmpdir) for _ in range(args.steps)]
        v = nreads * args.steps / sum(ts) / 1e6
        sample = "extract -@ %d -r chr1:1-%d of the workload BAM (%d alignments per step)" % (cores, region, nreads)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": round(1e3 * sum(ts) / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32 integer",
                          "data": "synthetic", "config": {"workload": workload, "sample": sample},
                          "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
                          "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumr(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    prefix = dataset(rank, args.mbp)
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

    # ------------------------------------------------------------ kernel-only: batch resident in HBM
    d = g.upload(soa)
    for _ in range(max(args.warmup, 3)):
        st = g.extract_tile_device(0, 0, reflen, d)
    sampler = ClockSampler(local_rank); sampler.start()
    barrier()
    l0 = g.launch_count()
    ev_total = ev_count = ev_prep = 0.0
    w0 = time.perf_counter()
    for _ in range(args.steps):
        st = g.extract_tile_device(0, 0, reflen, d)
        t = g.last_timing()
        ev_prep += t[1]; ev_count += t[2]; ev_total += t[1] + t[2]
    barrier()
    wall_dev = time.perf_counter() - w0
    launches = g.launch_count() - l0
    ms_step = maxr(ev_total / args.steps)
    total_aln = sumr(float(n))
    value = total_aln / (ms_step * 1e-3) / 1e6
    g.free(d)

    # ------------------------------------------------------------ end to end through the C ABI, host buffers
    # The batch is cut into tiles exactly as the sub-command driver cuts it (Tiler, 2^17 alignments per tile) and
    # streamed with md_submit_tile / md_collect_tile over the context's lanes: H2D of tile k+1 overlaps the kernels
    # of tile k and the read-back of tile k-1.  Host buffers are page-locked.  Every byte crosses PCIe every step.
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

    def e2e_step():
        """one pass over all tiles; returns (calls written, last stats)"""
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

    for _ in range(max(args.warmup, 3)):
        n_e2e_calls = e2e_step()
    assert n_e2e_calls == st.n_calls, (n_e2e_calls, st.n_calls)     # tiled + pipelined path reports the same columns
    barrier()
    l1 = g.launch_count()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = maxr((time.perf_counter() - w0) / args.steps)
    launches += g.launch_count() - l1
    e2e_value = total_aln / e2e_s / 1e6
    e2e_t = g.last_timing()

    # ------------------------------------------------------------ end to end from the COMPRESSED BAM (device-side BGZF inflate + decode)
    # What the drop-in binary does by default: the file's bytes (page-locked here) go to the device in segments of whole BGZF
    # blocks, md_bam_push_begin/_end inflates and frames them (segment k+1 overlapping the tiles of segment k), md_bam_extract_run
    # assembles one tile per segment in HBM and counts it; md_call records come back.  This is the path whose work matches what
    # the reference arm does with the same file (inflate + record decode + pileup), minus the text output.
    import struct
    raw = open(prefix + ".bam", "rb").read()
    hbuf = g.g.md_alloc_pinned(len(raw) + 64)
    C.memmove(hbuf, raw, len(raw))
    segs, cur_blocks, seg_start, off = [], [], 0, 0
    SEG = 128 << 20                                               # the sub-command driver's default segment size (host/cli.cpp: device_segment_bytes)
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

    def bam_step():
        g.g.md_bam_reset(bs_)
        so, sl, arr, nb = seg_arr[0]
        assert g.g.md_bam_push_begin(bs_, hbuf + so, sl, arr, nb, hoff) == 0, g.g.md_last_error()
        out_off, open_beg = 0, 0
        for k in range(len(seg_arr)):
            assert g.g.md_bam_push_end(bs_, C.byref(summ)) == 0, g.g.md_last_error()
            if k + 1 < len(seg_arr):
                so, sl, arr, nb = seg_arr[k + 1]
                assert g.g.md_bam_push_begin(bs_, hbuf + so, sl, arr, nb, 0) == 0, g.g.md_last_error()
            nr = g.g.md_bam_get_runs(bs_, runs1, 8)
            assert nr == 1 and runs1[0].tid == 0
            cut = reflen if k + 1 == len(seg_arr) else max(runs1[0].last_pos, open_beg)
            td = A.MdTileDesc(0, open_beg, cut, 0, 0)
            dst = C.cast(C.addressof(calls) + out_off * 16, C.POINTER(A.MdCall))
            assert g.g.md_bam_extract_run(bs_, 0, C.byref(td), 0xffffffff, dst, cap - out_off, C.byref(stt)) == 0, g.g.md_last_error()
            out_off += stt.n_calls; open_beg = cut
        return out_off

    for _ in range(max(args.warmup, 3)):
        n_bam_calls = bam_step()
    assert n_bam_calls == st.n_calls, (n_bam_calls, st.n_calls)       # same columns as the SoA paths
    barrier()
    l2 = g.launch_count()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        bam_step()
    barrier()
    bam_s = maxr((time.perf_counter() - w0) / args.steps)
    bam_launches = g.launch_count() - l2
    launches += bam_launches
    g.g.md_bam_close(bs_)
    g.g.md_free_pinned(hbuf)
    sampler.stop_flag = True; sampler.join(timeout=2)
    g.g.md_host_unregister(C.addressof(calls))
    for addr in pinned:
        g.g.md_host_unregister(addr)

    if rank != 0:
        g.close()
        if dist is not None:
            dist.barrier(); dist.destroy_process_group()
        return 0

    # ------------------------------------------------------------ roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg = algorithmic_bytes(soa, reflen, st.n_calls)
    count_ms = ev_count / args.steps
    achieved = alg / (count_ms * 1e-3) / 1e9
    # DRAM traffic of the same kernel on the same workload from the committed ncu --set full capture (profiles/r1_traffic.json,
    # written by tools/gpu_final.sh); a number measured under the profiler is only ever used for this field
    traffic = None; traffic_src = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        traffic = int(tj["dram_bytes_read"] + tj["dram_bytes_write"]); traffic_src = tj.get("source")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "count_warp<0, 1>", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 (of fallback)",
                "algorithmic_bytes_per_launch": int(alg), "kernel_ms": round(count_ms, 4), "prep_pair_window_ms": round(ev_prep / args.steps, 4), "traffic": traffic, "traffic_source": traffic_src}

    # ------------------------------------------------------------ the drop-in binary from the BAM file (informational)
    cli = None
    if world == 1:
        binp = os.path.join(ROOT, "methyldackel_b200", "lib", "MethylDackel")
        outp = os.path.join(tmpdir, "cli_out")
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            subprocess.run([binp, "extract", prefix + ".fa", prefix + ".bam", "-o", outp], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            ts.append(time.perf_counter() - t0)
        cli = {"value": round(n / min(ts) / 1e6, 3), "unit": UNIT, "seconds": round(min(ts), 3),
               "what": "lib/MethylDackel extract <fa> <bam> (process start, CUDA context, device-side BGZF inflate + decode, kernels, D2H, text output), best of 3"}

    # ------------------------------------------------------------ CPU baseline (reference build, all host threads, bounded sample)
    cpu = None
    if world == 1 and not args.no_cpu_baseline and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "MethylDackel")):
        region, _ = pick_reference_sample(prefix, soa, cores, tmpdir, 12.0, args.mbp * 1_000_000)
        tsec = time_reference(prefix, region, cores, tmpdir)
        nr = reads_in_region(soa, region)
        cpu = {"value": round(nr / tsec / 1e6, 4), "unit": UNIT, "cores": cores, "kind": "reference",
               "sample": "oracle/_ref/MethylDackel extract -@ %d -r chr1:1-%d of the workload BAM: %d alignments in %.2f s (BGZF inflate + pileup + text output included)" % (cores, region, nr, tsec)}

    out = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32 integer", "data": "synthetic",
           "config": {"workload": workload, "alignments_per_gpu": n, "contig_bp": reflen, "soa_phred_bits": int(soa.qual_bits) or 8, "options": "extract defaults (-q 10 -p 5 -F 0xF00, CpG only)",
                      "parallelism": "contig interval per GPU, no collective", "l2": "inputs (%.0f MB SoA per GPU) larger than the 126 MB L2" % (soa_bytes(soa) / 1e6)},
           "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(n_e2e_calls * 16 + 32 * len(tiles)),
                   "ms_per_step": round(e2e_s * 1e3, 3), "tiles_per_step": len(tiles), "lanes": NL,
                   "last_tile_ms": {"h2d": round(e2e_t[0], 3), "prep": round(e2e_t[1], 3), "count": round(e2e_t[2], 3), "d2h": round(e2e_t[3], 3)},
                   "path": "md_submit_tile()/md_collect_tile(): page-locked host SoA tiles -> H2D -> kernels -> D2H md_call records, 3 lanes in flight, every step"},
           "e2e_bam": {"value": round(total_aln / bam_s / 1e6, 3), "unit": UNIT, "ms_per_step": round(bam_s * 1e3, 3), "h2d_bytes_per_step": len(raw), "d2h_bytes_per_step": int(n_bam_calls * 16),
                       "segments_per_step": len(seg_arr), "launches_per_step": int(bam_launches // max(args.steps, 1)),
                       "path": "compressed BAM bytes (page-locked) -> md_bam_push_begin/_end (H2D, BGZF inflate, record framing on the device; segment k+1 overlaps the tiles of segment k) -> md_bam_extract_run (tile assembly in HBM + prep + count) -> D2H md_call records"},
           "gpu_launches": int(launches), "wall_ms_per_step_device_resident": round(1e3 * wall_dev / args.steps, 3),
           "roofline": roofline, "clocks": sampler.summary(), "calls_per_step": int(st.n_calls), "pairs_per_step": int(st.n_pairs)}
    if cli is not None:
        out["cli_from_bam"] = cli
    if cpu is not None:
        out["cpu_baseline"] = cpu
    print(json.dumps(out))
    g.close()
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
