#!/usr/bin/env python
"""Runs every kernel of the path a few times on the config[1] data set so that ONE ncu invocation can capture them all
(tools/gpu_r2_profile.sh).  Order of the launches after the warm-up (what --launch-skip / -k select):
  extract defaults (prep_kernel, count_warp<0,1>), extract with the variant filter (count_warp<1,1>), mbias (count_warp<2,1>),
  perRead (per_read_kernel), device decode of the compressed BAM (inflate_kernel, scan_blocks ..., tile_gather).
Refuses to run when the library was not built from the sources in the tree (md_source_hash)."""
import ctypes as C
import os
import struct
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from methyldackel_b200 import _abi as A  # noqa: E402
from methyldackel_b200 import api  # noqa: E402

tree = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "methyldackel_b200", "csrc"), "srchash"], capture_output=True, text=True).stdout.strip()
lib = A.load_gpu().md_source_hash().decode()
print("source hash: tree %s, libmdgpu.so %s" % (tree, lib), flush=True)
if tree != lib and not os.environ.get("KPROF_NOHASH"):
    sys.exit("kprof: lib/libmdgpu.so was not built from the sources in this tree — rebuild (make -C methyldackel_b200/csrc gpu)")

cache = os.environ.get("MDBENCH_CACHE", "/tmp/mdbench"); os.makedirs(cache, exist_ok=True)
p = os.path.join(cache, "kprof_c2")
if not os.path.exists(p + ".bam.bai"):
    subprocess.run([os.path.join(ROOT, "methyldackel_b200", "lib", "mdsynth"), "--out", p, "--contigs", "chr1:10000000", "--depth", "30"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
b = api.BamFile(p + ".bam")
ref = api.fetch_contig(p + ".fa", "chr1")
soa = b.read_region(0)
reps = int(os.environ.get("KPROF_REPS", "2"))

only_decode = os.environ.get("KPROF_ONLY") == "decode"       # variant sweeps of the inflate kernel (tools/gpu_r2_groups.sh)
for name, cfg, mb in () if only_decode else (("cpg", A.default_config(), False), ("var", A.default_config(minOppositeDepth=5, maxVariantFrac=0.25), False), ("mbias", A.default_config(noOverlapMerge=1), True)):
    g = api.GpuContext(cfg)
    g.load_contig(0, ref)
    if mb:
        g.set_mbias_chunks(0, list(range(0, len(ref), 1000000)) + [len(ref)])
    d = g.upload(soa)
    for _ in range(reps):
        st = (g.mbias_tile_device if mb else g.extract_tile_device)(0, 0, len(ref), d)
    t = g.last_timing()
    print("%s: prep %.3f ms, count %.3f ms, %d calls" % (name, t[1], t[2], st.n_calls), flush=True)
    g.free(d); g.close()

# perRead: host tiles of 2^19 alignments as the sub-command driver cuts them
g = api.GpuContext(A.default_config())
g.load_contig(0, ref)
tiles = b.make_tiles(0, 0, len(ref), 1 << 19)
td, s0 = tiles[0]
out = (A.MdReadMeth * s0.n_reads)()
for _ in range(0 if only_decode else reps):
    assert g.g.md_per_read_tile(g.h, C.byref(A.MdTileDesc(0, 0, len(ref), 0, 0)), C.byref(s0), 1000000, out) == 0, g.g.md_last_error()
print("perRead: %d alignments" % s0.n_reads, flush=True)

# device decode: the compressed file in one segment
raw = open(p + ".bam", "rb").read()
hbuf = g.g.md_alloc_pinned(len(raw) + 64)
C.memmove(hbuf, raw, len(raw))
blocks, off = [], 0
while off + 18 <= len(raw):
    xlen = struct.unpack_from("<H", raw, off + 10)[0]
    bs = struct.unpack_from("<H", raw, off + 16)[0] + 1
    blocks.append((off + 12 + xlen, bs - 12 - xlen - 8, struct.unpack_from("<I", raw, off + bs - 4)[0]))
    off += bs
nb = int(os.environ.get("KPROF_BLOCKS", str(len(blocks))))
blocks = blocks[:nb]
arr = (A.MdBgzfBlock * len(blocks))()
for k, (a_, b_, c_) in enumerate(blocks):
    arr[k].comp_off, arr[k].comp_len, arr[k].isize = a_, b_, c_
import gzip, io
u0 = gzip.GzipFile(fileobj=io.BytesIO(raw[:1 << 20])).read(1 << 16)
l_text = struct.unpack_from("<i", u0, 4)[0]; hoff = 8 + l_text
n_ref = struct.unpack_from("<i", u0, hoff)[0]; hoff += 4
for _ in range(n_ref):
    l_name = struct.unpack_from("<i", u0, hoff)[0]; hoff += 4 + l_name + 4
bs_ = g.g.md_bam_open(g.h, n_ref)
summ = A.MdBamSummary()
comp_bytes = blocks[-1][0] + blocks[-1][1] + 8
for _ in range(reps):
    g.g.md_bam_reset(bs_)
    assert g.g.md_bam_push(bs_, hbuf, comp_bytes, arr, len(blocks), hoff, C.byref(summ)) == 0, g.g.md_last_error()
tot = A.MdTotals(); g.g.md_ctx_totals(g.h, C.byref(tot))
print("device decode: %d blocks, %.1f MB compressed -> %.1f MB, %d records; inflate %.3f ms per push (%.1f GB/s compressed in, %.1f GB/s out), framing %.3f ms, H2D %.3f ms" % (
    len(blocks), comp_bytes / 1e6, summ.inflated_bytes / 1e6, summ.n_records, tot.inflate_ms / reps, comp_bytes / (tot.inflate_ms / reps) / 1e6, summ.inflated_bytes / (tot.inflate_ms / reps) / 1e6,
    tot.frame_ms / reps, tot.push_h2d_ms / reps), flush=True)
g.g.md_bam_close(bs_)
g.g.md_free_pinned(hbuf)
g.close()
