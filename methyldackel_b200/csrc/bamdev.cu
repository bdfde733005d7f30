// Device-side BGZF inflate + BAM record -> SoA decode.  Included at the end of mdgpu.cu (one
// translation unit: it drives the same lanes, scratch and kernels as the host-tile entry points).
//
// What it replaces: the work inside sam_itr_next (common.c:413) — bgzf block inflate, record framing, field extraction —
// which the CLI otherwise does on host cores (host/pardecode.hpp).  The caller hands over SEGMENTS of the compressed file
// (whole BGZF blocks, cut by scanning the 18-byte block headers); everything downstream happens in HBM:
//
//   inflate_kernel     one warp per BGZF block; the leader lane runs inflate_hd.h:inflate_block with its tables in shared memory
//   scan_blocks_kernel one thread per block: first plausible record start inside the block + walk of the record chain from it
//   check/fix_chain    adopt the per-block chains whose start equals the predecessor's exit (a parallel check); a serial
//                      walk repairs the blocks where the guess was wrong, so results never depend on the guess
//   fill_offsets       record start offsets, in file order (block bases from a prefix sum)
//   head_kernel        per record: contig id, position, reference end, well-formedness
//   runs_kernel        runs of records on the same contig (what the caller iterates over)
//   tile_sizes/gather  per contig run: carried reads of the previous tile (those reaching beyond its cut) + the run's own
//                      records -> one md_reads_soa tile in HBM, bit-identical to what host/tiles.hpp builds from the same records
//
// after which the tile goes through prep_kernel / count_warp / gather exactly like a host-supplied one.
#include <cub/device/device_scan.cuh>
#include "bamdev_hd.h"

namespace {
using mdbam::BlockScan; using mdbam::TileSrc; using mdbam::TileDst; using mdbam::Sz4;

constexpr int INF_WARPS = 8;
__global__ void __launch_bounds__(INF_WARPS * 32) inflate_kernel(const uint8_t *comp, const md_bgzf_block *blk, const unsigned long long *uoff, uint8_t *ubuf, uint32_t n_blocks, int *err) {
    __shared__ mdinflate::Tables T[INF_WARPS];
    const uint32_t b = blockIdx.x * INF_WARPS + (threadIdx.x >> 5);
    if (b >= n_blocks || (threadIdx.x & 31)) return;
    const md_bgzf_block d = blk[b];
    if (d.isize == 0) return;
    const int rc = mdinflate::inflate_block(comp, d.comp_off, d.comp_len, ubuf + uoff[b], d.isize, T[threadIdx.x >> 5]);
    if (rc) atomicCAS(err, 0, (int)((b << 4) | (uint32_t)(-rc)));
}
__global__ void scan_blocks_kernel(const uint8_t *u, const unsigned long long *uoff, uint32_t n_blocks, unsigned long long first, unsigned long long U, int32_t n_targets, BlockScan *out) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n_blocks) mdbam::scan_block_body(b, u, uoff, first, U, n_targets, out);
}
__global__ void check_chain_kernel(const unsigned long long *uoff, uint32_t n_blocks, unsigned long long first, const BlockScan *sc, int *bad) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n_blocks && !mdbam::check_block_body(b, uoff, first, sc)) atomicExch(bad, 1);
}
__global__ void fix_chain_kernel(const uint8_t *u, const unsigned long long *uoff, uint32_t n_blocks, unsigned long long first, unsigned long long U, BlockScan *sc, const int *bad, unsigned long long *final_exit) {
    if (threadIdx.x || blockIdx.x) return;
    mdbam::fix_chain_body(u, uoff, n_blocks, first, U, sc, *bad, final_exit);
}
__global__ void counts_kernel(const BlockScan *sc, uint32_t n_blocks, uint32_t *cnt) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n_blocks) cnt[b] = sc[b].guess != mdbam::NONE ? sc[b].count : 0u;
}
__global__ void fill_offsets_kernel(const uint8_t *u, const unsigned long long *uoff, uint32_t n_blocks, unsigned long long U, const BlockScan *sc, const uint32_t *base, unsigned long long *rec_off) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n_blocks) mdbam::fill_offsets_body(b, u, uoff, U, sc, base, rec_off);
}
__global__ void head_kernel(const uint8_t *u, const unsigned long long *rec_off, uint32_t n, int32_t *tid, int32_t *pos, int32_t *rend, int *err) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !mdbam::head_body(i, u, rec_off, tid, pos, rend)) atomicCAS(err, 0, -77);
}

// run r: records [start, start+n) share a contig.  Entries are written unordered; the host sorts them by start.
__global__ void runs_kernel(const int32_t *tid, const int32_t *pos, uint32_t n, md_bam_run *runs, uint32_t cap, uint32_t *n_runs, int32_t *last_pos) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == n - 1) *last_pos = pos[i];
    if (i == 0 || tid[i] != tid[i - 1]) {
        const uint32_t k = atomicAdd(n_runs, 1u);
        if (k < cap) { md_bam_run r; r.tid = tid[i]; r.start = i; r.n = 0; r.first_pos = pos[i]; r.last_pos = 0; r.prev_last_pos = i ? pos[i - 1] : 0; runs[k] = r; }
    }
}

struct U4Sum { __host__ __device__ __forceinline__ Sz4 operator()(const Sz4 &a, const Sz4 &b) const { Sz4 r; r.x = a.x + b.x; r.y = a.y + b.y; r.z = a.z + b.z; r.w = a.w + b.w; return r; } };
__global__ void tile_sizes_kernel(TileSrc S, Sz4 *sz) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < S.n_prev + S.n_own) sz[e] = mdbam::tile_sizes_body(e, S);
}
__global__ void tile_gather_kernel(TileSrc S, const Sz4 *sz, const Sz4 *off, TileDst D) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < S.n_prev + S.n_own) mdbam::tile_gather_body(e, S, sz[e], off[e], D);
}

}  // namespace

// ---- host side ------------------------------------------------------------------------------------------------------
static const size_t BAM_HEADROOM = (size_t) 64 << 20;      // room in front of a segment's bytes for the record that straddles in

struct TileArena { DevBuf buf; DevReads view; int32_t *rend = nullptr; uint32_t n = 0; int32_t tid = -1; uint32_t cut = 0; bool valid = false; };

struct md_bam_stream {
    md_ctx *c = nullptr; int32_t n_targets = 0;
    DevBuf comp, blk, uoff, ubuf, scan, cnt, base, rec_off, tid, pos, rend, runs, small, cub_tmp, sz, off;
    uint64_t U = 0, D0 = 0;             // current segment: data is ubuf[D0, U)
    uint64_t leftover_from = 0, leftover = 0;
    uint32_t n_records = 0; bool have_segment = false;
    std::vector<md_bam_run> runs_host;
    TileArena tile[2]; int cur = 0;
    // page-locked mailbox: [0] n_runs, [1] inflate/parse error, [2] chain-check flag, [3] last_pos, [4..5] final exit, [6..7] n_records(total)
    uint32_t *h_small = nullptr;
    Sz4 *h_tot = nullptr;
    double t_push[8] = {0}, t_tile[8] = {0}; uint64_t n_push = 0, n_tiles = 0, comp_total = 0, infl_total = 0, rec_total = 0;
};

extern "C" md_bam_stream *md_bam_open(md_ctx *c, int32_t n_targets) {
    if (!c) { g_err = "md_bam_open: no context"; return nullptr; }
    CKN(cudaSetDevice(c->device));
    md_bam_stream *s = new md_bam_stream();
    s->c = c; s->n_targets = n_targets;
    if (cudaMallocHost((void **) &s->h_small, 64) != cudaSuccess || cudaMallocHost((void **) &s->h_tot, 2 * sizeof(Sz4)) != cudaSuccess) { g_err = "cudaMallocHost failed"; delete s; return nullptr; }
    if (s->small.reserve(256)) { delete s; return nullptr; }
    return s;
}
extern "C" void md_bam_close(md_bam_stream *s) {
    if (!s) return;
    cudaSetDevice(s->c->device); sync_all(s->c);
    if (getenv("MD_TIMING"))
        fprintf(stderr, "[md-timing] device decode: %llu segments, %.1f MB compressed -> %.1f MB, %llu records; H2D %.2f ms, inflate %.2f ms, record chains %.2f ms, offsets+heads+runs %.2f ms; %llu tiles: sizes+scan %.2f ms, gather %.2f ms\n",
                (unsigned long long) s->n_push, s->comp_total / 1e6, s->infl_total / 1e6, (unsigned long long) s->rec_total, s->t_push[0], s->t_push[1], s->t_push[2], s->t_push[3], (unsigned long long) s->n_tiles, s->t_tile[0], s->t_tile[1]);
    DevBuf *bufs[] = {&s->comp, &s->blk, &s->uoff, &s->ubuf, &s->scan, &s->cnt, &s->base, &s->rec_off, &s->tid, &s->pos, &s->rend, &s->runs, &s->small, &s->cub_tmp, &s->sz, &s->off, &s->tile[0].buf, &s->tile[1].buf};
    for (DevBuf *b : bufs) b->release();
    if (s->h_small) cudaFreeHost(s->h_small);
    if (s->h_tot) cudaFreeHost(s->h_tot);
    delete s;
}
extern "C" void md_bam_reset(md_bam_stream *s) {       // after a seek: forget the straddling record and the carried reads
    s->leftover = 0; s->have_segment = false; s->tile[0].valid = s->tile[1].valid = false; s->n_records = 0; s->runs_host.clear();
}

static const uint32_t BAM_MAX_RUNS = 1u << 16;

// MD_TIMING=1: per-stage device times of the decode (CUDA events on the lane's stream), accumulated per stream object
struct BamTimer {
    bool on; cudaStream_t st; cudaEvent_t ev[8]; int n = 0;
    explicit BamTimer(cudaStream_t s) : on(getenv("MD_TIMING") != nullptr), st(s) { if (on) for (auto &e : ev) cudaEventCreate(&e); }
    ~BamTimer() { if (on) for (auto &e : ev) cudaEventDestroy(e); }
    void tick() { if (on && n < 8) cudaEventRecord(ev[n++], st); }
    void add(double *acc) { if (!on) return; cudaEventSynchronize(ev[n - 1]); for (int k = 0; k + 1 < n; ++k) { float ms = 0; cudaEventElapsedTime(&ms, ev[k], ev[k + 1]); acc[k] += ms; } }
};

extern "C" int md_bam_push(md_bam_stream *s, const void *comp, uint64_t comp_bytes, const md_bgzf_block *blocks, uint32_t n_blocks, uint32_t skip, md_bam_summary *out) {
    md_ctx *c = s->c;
    CK(cudaSetDevice(c->device));
    Lane *L = &c->lanes[0];
    cudaStream_t st = L->stream;
    sync_all(c);                                              // v1: one segment at a time; tiles of the previous segment are complete
    memset(out, 0, sizeof *out);
    // stream layout of this segment
    std::vector<unsigned long long> uoff(n_blocks + 1);
    uint64_t tot = 0;
    for (uint32_t b = 0; b < n_blocks; ++b) {
        if (blocks[b].comp_off + blocks[b].comp_len > comp_bytes) { g_err = "md_bam_push: block outside the buffer"; return -2; }
        uoff[b] = BAM_HEADROOM + tot; tot += blocks[b].isize;
    }
    uoff[n_blocks] = BAM_HEADROOM + tot;
    if (s->leftover > BAM_HEADROOM) { g_err = "md_bam_push: a record larger than 64 MB straddles two segments"; return -2; }
    const uint64_t D0 = BAM_HEADROOM - s->leftover, U = BAM_HEADROOM + tot;
    if (tot + s->leftover >= ((uint64_t) 1 << 32) - BAM_HEADROOM) { g_err = "md_bam_push: segment inflates to more than 4 GB"; return -2; }
    // the straddling record's first bytes move in front of the new data (before ubuf may be re-allocated)
    DevBuf keep;
    if (s->leftover) {
        if (keep.reserve(s->leftover)) return -100;
        CK(cudaMemcpyAsync(keep.p, (uint8_t *) s->ubuf.p + s->leftover_from, s->leftover, cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
    }
    if (s->comp.reserve(comp_bytes + 64) || s->blk.reserve((size_t) n_blocks * sizeof(md_bgzf_block) + 16) || s->uoff.reserve((size_t)(n_blocks + 1) * 8) ||
        s->ubuf.reserve(U + 64) || s->scan.reserve((size_t) n_blocks * sizeof(BlockScan) + 16) || s->cnt.reserve((size_t) n_blocks * 4 + 16) || s->base.reserve((size_t) n_blocks * 4 + 16) ||
        s->runs.reserve((size_t) BAM_MAX_RUNS * sizeof(md_bam_run))) { keep.release(); return -100; }
    BamTimer tm(st); tm.tick();
    if (s->leftover) { CK(cudaMemcpyAsync((uint8_t *) s->ubuf.p + D0, keep.p, s->leftover, cudaMemcpyDeviceToDevice, st)); }
    CK(cudaMemcpyAsync(s->comp.p, comp, comp_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync((uint8_t *) s->comp.p + comp_bytes, 0, 64, st));
    CK(cudaMemcpyAsync(s->blk.p, blocks, (size_t) n_blocks * sizeof(md_bgzf_block), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->uoff.p, uoff.data(), (size_t)(n_blocks + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(s->small.p, 0, 256, st));
    uint32_t *d_small = (uint32_t *) s->small.p;       // [0] n_runs [1] err [2] bad [3] last_pos [4,5] final_exit
    const uint8_t *u = (const uint8_t *) s->ubuf.p;
    const unsigned long long first = D0 + (s->leftover ? 0 : skip);
    tm.tick();
    if (n_blocks) {
        inflate_kernel<<<(n_blocks + INF_WARPS - 1) / INF_WARPS, INF_WARPS * 32, 0, st>>>((const uint8_t *) s->comp.p, (const md_bgzf_block *) s->blk.p, (const unsigned long long *) s->uoff.p, (uint8_t *) s->ubuf.p, n_blocks, (int *)(d_small + 1));
        tm.tick();
        const uint32_t g = (n_blocks + 127) / 128;
        // block 0's slice starts at D0 so that the straddling record is part of its chain
        unsigned long long d0 = D0;
        CK(cudaMemcpyAsync(s->uoff.p, &d0, 8, cudaMemcpyHostToDevice, st));
        scan_blocks_kernel<<<g, 128, 0, st>>>(u, (const unsigned long long *) s->uoff.p, n_blocks, first, U, s->n_targets, (BlockScan *) s->scan.p);
        check_chain_kernel<<<g, 128, 0, st>>>((const unsigned long long *) s->uoff.p, n_blocks, first, (const BlockScan *) s->scan.p, (int *)(d_small + 2));
        fix_chain_kernel<<<1, 32, 0, st>>>(u, (const unsigned long long *) s->uoff.p, n_blocks, first, U, (BlockScan *) s->scan.p, (const int *)(d_small + 2), (unsigned long long *)(d_small + 4));
        counts_kernel<<<g, 128, 0, st>>>((const BlockScan *) s->scan.p, n_blocks, (uint32_t *) s->cnt.p);
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const uint32_t *) s->cnt.p, (uint32_t *) s->base.p, (int) n_blocks, st);
        if (s->cub_tmp.reserve(tmp + 256)) { keep.release(); return -100; }
        cub::DeviceScan::ExclusiveSum(s->cub_tmp.p, tmp, (const uint32_t *) s->cnt.p, (uint32_t *) s->base.p, (int) n_blocks, st);
        c->launches += 6;
    }
    tm.tick();
    // number of records = base[last] + cnt[last]
    uint32_t last2[2] = {0, 0};
    if (n_blocks) {
        CK(cudaMemcpyAsync(&last2[0], (uint32_t *) s->base.p + (n_blocks - 1), 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&last2[1], (uint32_t *) s->cnt.p + (n_blocks - 1), 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaMemcpyAsync(s->h_small, d_small, 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    keep.release();
    if (s->h_small[1]) { char msg[128]; snprintf(msg, sizeof msg, "BGZF inflate failed on the device (block %u, code -%u)", s->h_small[1] >> 4, s->h_small[1] & 15u); g_err = msg; return -5; }
    const uint32_t n = last2[0] + last2[1];
    unsigned long long final_exit; memcpy(&final_exit, s->h_small + 4, 8);
    if (!n_blocks) final_exit = first;
    s->U = U; s->D0 = D0; s->n_records = n; s->have_segment = true;
    s->leftover_from = final_exit; s->leftover = U - final_exit;
    s->runs_host.clear();
    if (n) {
        if (s->rec_off.reserve((size_t) n * 8 + 16) || s->tid.reserve((size_t) n * 4 + 16) || s->pos.reserve((size_t) n * 4 + 16) || s->rend.reserve((size_t) n * 4 + 16)) return -100;
        const uint32_t g = (n_blocks + 127) / 128, gr = (n + 255) / 256;
        fill_offsets_kernel<<<g, 128, 0, st>>>(u, (const unsigned long long *) s->uoff.p, n_blocks, U, (const BlockScan *) s->scan.p, (const uint32_t *) s->base.p, (unsigned long long *) s->rec_off.p);
        head_kernel<<<gr, 256, 0, st>>>(u, (const unsigned long long *) s->rec_off.p, n, (int32_t *) s->tid.p, (int32_t *) s->pos.p, (int32_t *) s->rend.p, (int *)(d_small + 1));
        runs_kernel<<<gr, 256, 0, st>>>((const int32_t *) s->tid.p, (const int32_t *) s->pos.p, n, (md_bam_run *) s->runs.p, BAM_MAX_RUNS, d_small, (int32_t *)(d_small + 3));
        c->launches += 3;
        tm.tick();
        CK(cudaMemcpyAsync(s->h_small, d_small, 32, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (s->h_small[1]) { g_err = "malformed BAM record"; return -5; }
        const uint32_t nr = s->h_small[0];
        if (nr > BAM_MAX_RUNS) { g_err = "md_bam_push: more than 65536 contig runs in one segment"; return -2; }
        s->runs_host.resize(nr);
        CK(cudaMemcpy(s->runs_host.data(), s->runs.p, (size_t) nr * sizeof(md_bam_run), cudaMemcpyDeviceToHost));
        std::sort(s->runs_host.begin(), s->runs_host.end(), [](const md_bam_run &a, const md_bam_run &b) { return a.start < b.start; });
        for (uint32_t k = 0; k < nr; ++k) {
            md_bam_run &r = s->runs_host[k];
            const uint32_t nxt = k + 1 < nr ? s->runs_host[k + 1].start : n;
            r.n = nxt - r.start;
            r.last_pos = k + 1 < nr ? s->runs_host[k + 1].prev_last_pos : (int32_t) s->h_small[3];
        }
    }
    tm.add(s->t_push); s->n_push++; s->comp_total += comp_bytes; s->infl_total += tot; s->rec_total += n;
    CK(cudaGetLastError());
    out->n_records = n; out->n_runs = (uint32_t) s->runs_host.size(); out->inflated_bytes = tot; out->leftover_bytes = s->leftover;
    return 0;
}
extern "C" int md_bam_get_runs(md_bam_stream *s, md_bam_run *runs, uint32_t cap) {
    const uint32_t n = (uint32_t) s->runs_host.size();
    for (uint32_t k = 0; k < n && k < cap; ++k) runs[k] = s->runs_host[k];
    return (int) n;
}

// Build the tile (carried reads of the previous tile of this contig + records [run.start, run.start+run.n) of the last segment)
// and leave it device-resident in s->tile[s->cur].
static int bam_build_tile(md_bam_stream *s, int run, const md_tile_desc *t, uint32_t keep_hi) {
    md_ctx *c = s->c; Lane *L = &c->lanes[0]; cudaStream_t st = L->stream;
    TileArena &P = s->tile[s->cur], &N = s->tile[s->cur ^ 1];
    TileSrc S; memset(&S, 0, sizeof S);
    S.u = (const uint8_t *) s->ubuf.p; S.rec_off = (const unsigned long long *) s->rec_off.p; S.pos = (const int32_t *) s->pos.p; S.rend = (const int32_t *) s->rend.p;
    if (run >= 0) {
        if (!s->have_segment || (size_t) run >= s->runs_host.size()) { g_err = "md_bam: no such run"; return -2; }
        const md_bam_run &r = s->runs_host[(size_t) run];
        if (r.tid != t->tid) { g_err = "md_bam: tile and run are on different contigs"; return -2; }
        S.r0 = r.start; S.n_own = r.n;
    }
    const bool carry = P.valid && P.tid == t->tid && P.cut == t->beg;
    if (carry) {
        const DevReads &V = P.view;
        S.prev.pos = V.pos; S.prev.flag = V.flag; S.prev.mapq = V.mapq; S.prev.aux = V.aux; S.prev.l_qseq = V.l_qseq; S.prev.cigar_off = V.cigar_off; S.prev.seq_off = V.seq_off;
        S.prev.qual_off = V.qual_off; S.prev.frag_key = V.frag_key; S.prev.cigar = V.cigar; S.prev.seq = V.seq; S.prev.qual = V.qual;
        S.prev_rend = P.rend; S.n_prev = P.n;
    }
    S.keep_lo = t->beg; S.keep_hi = keep_hi;
    const uint32_t m = S.n_prev + S.n_own;
    N.valid = false; N.n = 0;
    BamTimer tm(st); tm.tick();
    Sz4 tot; tot.x = tot.y = tot.z = tot.w = 0;
    const Sz4 zero4 = tot;
    if (m) {
        if (s->sz.reserve((size_t) m * 16 + 16) || s->off.reserve((size_t) m * 16 + 16)) return -100;
        tile_sizes_kernel<<<(m + 255) / 256, 256, 0, st>>>(S, (Sz4 *) s->sz.p);
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveScan(nullptr, tmp, (const Sz4 *) s->sz.p, (Sz4 *) s->off.p, U4Sum(), zero4, (int) m, st);
        if (s->cub_tmp.reserve(tmp + 256)) return -100;
        cub::DeviceScan::ExclusiveScan(s->cub_tmp.p, tmp, (const Sz4 *) s->sz.p, (Sz4 *) s->off.p, U4Sum(), zero4, (int) m, st);
        CK(cudaMemcpyAsync(&s->h_tot[0], (Sz4 *) s->off.p + (m - 1), 16, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&s->h_tot[1], (Sz4 *) s->sz.p + (m - 1), 16, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        tot = U4Sum()(s->h_tot[0], s->h_tot[1]);
        c->launches += 2;
    }
    tm.tick();
    const size_t n = tot.x;
    size_t szs[13] = {n * 4, n * 2, n, n, n * 4, (n + 1) * 4, n * 4, n * 4, n * 8, n * 4, (size_t) tot.y * 4, (size_t) tot.z * 4, (size_t) tot.w * 8};
    size_t offs[13], total = 0;
    for (int k = 0; k < 13; ++k) { offs[k] = total; total += al256(szs[k] + 16); }
    if (N.buf.reserve(total)) return -100;
    unsigned char *base = (unsigned char *) N.buf.p;
    TileDst D;
    D.pos = (int32_t *)(base + offs[0]); D.flag = (uint16_t *)(base + offs[1]); D.mapq = base + offs[2]; D.aux = base + offs[3]; D.l_qseq = (uint32_t *)(base + offs[4]);
    D.cigar_off = (uint32_t *)(base + offs[5]); D.seq_off = (uint32_t *)(base + offs[6]); D.qual_off = (uint32_t *)(base + offs[7]); D.frag_key = (uint64_t *)(base + offs[8]);
    D.rend = (int32_t *)(base + offs[9]); D.cigar = (uint32_t *)(base + offs[10]); D.seq = (uint32_t *)(base + offs[11]); D.qual = (uint64_t *)(base + offs[12]);
    if (m) { tile_gather_kernel<<<(m + 255) / 256, 256, 0, st>>>(S, (const Sz4 *) s->sz.p, (const Sz4 *) s->off.p, D); c->launches += 1; }
    else CK(cudaMemsetAsync(D.cigar_off, 0, 4, st));
    DevReads &v = N.view; memset(&v, 0, sizeof v);
    v.n = (uint32_t) n; v.seq_words = tot.z; v.qual_words = tot.w; v.qbits = 8;
    v.pos = D.pos; v.flag = D.flag; v.mapq = D.mapq; v.aux = D.aux; v.l_qseq = D.l_qseq; v.cigar_off = D.cigar_off; v.seq_off = D.seq_off; v.qual_off = D.qual_off;
    v.frag_key = D.frag_key; v.cigar = D.cigar; v.seq = D.seq; v.qual = D.qual;
    tm.tick(); tm.add(s->t_tile); s->n_tiles++;
    N.rend = D.rend; N.n = (uint32_t) n; N.tid = t->tid; N.cut = t->end; N.valid = true;
    s->cur ^= 1;
    CK(cudaGetLastError());
    return 0;
}

static int bam_run_tile(md_bam_stream *s, int run, const md_tile_desc *t, uint32_t keep_hi, bool mbias, md_call *calls, uint64_t cap, md_tile_stats *stats) {
    md_ctx *c = s->c;
    CK(cudaSetDevice(c->device));
    Lane *L = &c->lanes[0];
    CK(cudaEventRecord(L->ev[0], L->stream));
    int rc = bam_build_tile(s, run, t, keep_hi);
    if (rc) return rc;
    TileArena &T = s->tile[s->cur];
    rc = run_pipeline(c, L, t, T.view, mbias);
    if (rc) return rc;
    rc = finish_counters(c, L, stats);
    if (rc) return rc;
    if (!mbias) rc = fetch_sorted(c, L, calls, cap, nullptr);
    CK(cudaEventRecord(L->ev[4], L->stream));
    CK(cudaStreamSynchronize(L->stream));
    collect_timing(L);
    c->last = L;
    return rc;
}
extern "C" int md_bam_extract_run(md_bam_stream *s, int run, const md_tile_desc *tile, uint32_t keep_hi, md_call *calls, uint64_t capacity, md_tile_stats *stats) {
    return bam_run_tile(s, run, tile, keep_hi, false, calls, capacity, stats);
}
extern "C" int md_bam_mbias_run(md_bam_stream *s, int run, const md_tile_desc *tile, uint32_t keep_hi, md_tile_stats *stats) {
    return bam_run_tile(s, run, tile, keep_hi, true, nullptr, 0, stats);
}
// the tile just built, as host arrays (tests compare it with the host decoder's tile); arrays sized by the caller from md_bam_tile_shape
extern "C" int md_bam_tile_shape(md_bam_stream *s, md_reads_soa *shape) {
    TileArena &T = s->tile[s->cur];
    memset(shape, 0, sizeof *shape);
    if (!T.valid) return -1;
    shape->n_reads = T.n; shape->seq_words = T.view.seq_words; shape->qual_words = T.view.qual_words; shape->qual_bits = 8;
    uint32_t nc = 0;
    CK(cudaMemcpy(&nc, T.view.cigar_off + T.n, 4, cudaMemcpyDeviceToHost));
    shape->n_cigar_ops = nc;
    return 0;
}
extern "C" int md_bam_tile_fetch(md_bam_stream *s, md_reads_soa *dst, int32_t *rend) {
    TileArena &T = s->tile[s->cur];
    if (!T.valid) return -1;
    const size_t n = T.n;
    CK(cudaMemcpy((void *) dst->pos, T.view.pos, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy((void *) dst->flag, T.view.flag, n * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy((void *) dst->mapq, T.view.mapq, n, cudaMemcpyDeviceToHost)); CK(cudaMemcpy((void *) dst->aux, T.view.aux, n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy((void *) dst->l_qseq, T.view.l_qseq, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy((void *) dst->cigar_off, T.view.cigar_off, (n + 1) * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy((void *) dst->seq_off, T.view.seq_off, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy((void *) dst->qual_off, T.view.qual_off, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy((void *) dst->frag_key, T.view.frag_key, n * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy((void *) dst->cigar, T.view.cigar, (size_t) dst->n_cigar_ops * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy((void *) dst->seq, T.view.seq, (size_t) dst->seq_words * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy((void *) dst->qual, T.view.qual, (size_t) dst->qual_words * 8, cudaMemcpyDeviceToHost));
    if (rend) CK(cudaMemcpy(rend, T.rend, n * 4, cudaMemcpyDeviceToHost));
    return 0;
}
