// Device-side BGZF inflate + BAM record -> SoA decode.  Included at the end of mdgpu.cu (one
// translation unit: it drives the same lanes, scratch and kernels as the host-tile entry points).
//
// What it replaces: the work inside sam_itr_next (common.c:413) — bgzf block inflate, record framing, field extraction —
// which the CLI otherwise does on host cores (host/pardecode.hpp).  The caller hands over SEGMENTS of the compressed file
// (whole BGZF blocks, cut by scanning the 18-byte block headers); everything downstream happens in HBM:
//
//   inflate_kernel     one warp per BGZF block: inflate_hd.h:inflate_block, tables + input ring + output ring in shared memory
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
#include <thread>
#include "bamdev_hd.h"

namespace {
using mdbam::BlockScan; using mdbam::TileSrc; using mdbam::TileDst; using mdbam::Sz4;

// One warp per BGZF block, INF_WARPS warps per CTA, INF_CTAS_PER_SM CTAs per SM; every warp owns a mdinflate::Decoder (tables + the
// ring of its most recent output) in shared memory.  Deflate decoding is a serial bit-stream walk; all 32 lanes run it
// redundantly (see inflate_hd.h) and share out the match copies and the flushes of the ring to global memory.
#ifndef MD_INFLATE_WARPS
#define MD_INFLATE_WARPS 8
#endif
constexpr int INF_WARPS = MD_INFLATE_WARPS, INF_CTAS_PER_SM = 4;  // 32 warps per SM: 5.25 KB of shared memory each (2 KB ring), 64 registers per thread
constexpr size_t INF_SMEM = INF_WARPS * sizeof(mdinflate::Decoder);
static_assert(INF_CTAS_PER_SM * (INF_SMEM + 1024) <= 228 * 1024, "inflate_kernel: decoders do not fit the SM's shared memory");
__global__ void __launch_bounds__(INF_WARPS * 32, INF_CTAS_PER_SM) inflate_kernel(const uint8_t *comp, const md_bgzf_block *blk, const unsigned long long *uoff, uint8_t *ubuf, uint32_t b0, uint32_t n_blocks, int *err) {
    extern __shared__ __align__(16) unsigned char inf_smem[];
    const int lane = (int)(threadIdx.x & 31);
    mdinflate::Decoder &D = ((mdinflate::Decoder *) inf_smem)[threadIdx.x >> 5];
    // one block per warp and the CTA retires: the decode stream has the lowest priority, so the slots that free up go to the
    // count / prep kernels of the tile in flight first (persistent decoder CTAs would hold every SM until the segment is done)
    const uint32_t b = b0 + blockIdx.x * INF_WARPS + (threadIdx.x >> 5);      // this launch covers blocks [b0, n_blocks)
    if (b >= n_blocks) return;
    const md_bgzf_block d = blk[b];
    if (d.isize == 0) return;
    const int rc = mdinflate::inflate_block(comp, d.comp_off, d.comp_len, ubuf + uoff[b], d.isize, D, lane, 32);
    if (rc && lane == 0) atomicCAS(err, 0, (int)((b << 4) | (uint32_t)(-rc)));
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

struct TileArena { DevBuf buf; DevReads view; int32_t *rend = nullptr; uint32_t n = 0, n_cigar = 0; int32_t tid = -1; uint32_t cut = 0; bool valid = false; };

// One pushed segment: its compressed bytes, inflated stream, record table and the per-contig runs.  Two slots alternate so
// that segment k+1 is copied and decoded (on its own stream, driven by a helper thread) while the tiles of segment k are
// assembled and counted on the lane's stream.
struct BamSlot {
    DevBuf comp, blk, uoff, ubuf, scan, cnt, base, rec_off, tid, pos, rend, runs, small, cub_tmp;
    uint32_t *h_small = nullptr;        // page-locked mailbox: [0] n_runs, [1] inflate/parse error, [2] chain-check flag, [3] last_pos, [4..5] final exit
    uint64_t U = 0, D0 = 0, leftover_from = 0, leftover = 0;
    uint32_t n_records = 0;
    std::vector<md_bam_run> runs_host;
    md_bam_summary sum; int rc = 0; std::string err; uint64_t launches = 0;
    // md_bam_prefetch: the compressed bytes of the segment this slot will decode next are already on their way (copy stream)
    const void *pref_ptr = nullptr; uint64_t pref_bytes = 0; cudaEvent_t ev_p0 = nullptr, ev_p1 = nullptr;
};

struct md_bam_stream {
    md_ctx *c = nullptr; int32_t n_targets = 0;
    BamSlot slot[2]; int cur_slot = 0; bool have_segment = false;
    cudaStream_t sd = nullptr;          // decode stream
    // a large segment is copied in PUSH_PARTS pieces on `sc`, and every piece is inflated (its own launch, on its own stream)
    // as soon as it has arrived, so that only the first piece's copy is exposed instead of the whole segment's
    enum { PUSH_PARTS = 4 };
    cudaStream_t sc = nullptr, si[PUSH_PARTS] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_pre = nullptr, ev_c0 = nullptr, ev_i0 = nullptr, ev_copy[PUSH_PARTS] = {nullptr, nullptr, nullptr, nullptr}, ev_inf[PUSH_PARTS] = {nullptr, nullptr, nullptr, nullptr};
    bool parts_ok = false, use_parts = false, prefetch_ok = false;
    std::thread worker; bool inflight = false; int target = 0;
    DevBuf cub_tmp, sz, off;
    TileArena tile[2]; int cur = 0;
    Sz4 *h_tot = nullptr;
    double t_push[8] = {0}, t_tile[8] = {0}; uint64_t n_push = 0, n_tiles = 0, comp_total = 0, infl_total = 0, rec_total = 0;
};

extern "C" md_bam_stream *md_bam_open(md_ctx *c, int32_t n_targets) {
    if (!c) { g_err = "md_bam_open: no context"; return nullptr; }
    CKN(cudaSetDevice(c->device));
    md_bam_stream *s = new md_bam_stream();
    s->c = c; s->n_targets = n_targets;
    cudaFuncSetAttribute(inflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) INF_SMEM);
    int prio_lo = 0, prio_hi = 0; cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);      // numerically: lowest priority, greatest priority
    bool ok = cudaStreamCreateWithPriority(&s->sd, cudaStreamNonBlocking, prio_lo) == cudaSuccess && cudaMallocHost((void **) &s->h_tot, 2 * sizeof(Sz4)) == cudaSuccess;
    for (int k = 0; k < 2 && ok; ++k) ok = cudaMallocHost((void **) &s->slot[k].h_small, 64) == cudaSuccess && s->slot[k].small.reserve(256) == 0;
    if (!ok) { g_err = "md_bam_open: allocation failed"; delete s; return nullptr; }
    {   // optional: without these the segment is copied and inflated in one piece
        // (measured: for 96 MB segments — one wave of decoder warps — the staggered quarter waves cost more than the hidden copy
        //  saves, 9.8 against 9.1 ms per segment; for a 183 MB segment they gain 1 ms of 17.6.  Off unless MD_PUSH_PARTS=4; what
        //  hides the copy by default is md_bam_prefetch.)
        const char *e = getenv("MD_PUSH_PARTS");
        s->use_parts = e && atoi(e) == md_bam_stream::PUSH_PARTS;
        bool po = cudaStreamCreateWithFlags(&s->sc, cudaStreamNonBlocking) == cudaSuccess;
        s->prefetch_ok = po && cudaEventCreate(&s->slot[0].ev_p0) == cudaSuccess && cudaEventCreate(&s->slot[0].ev_p1) == cudaSuccess &&
                         cudaEventCreate(&s->slot[1].ev_p0) == cudaSuccess && cudaEventCreate(&s->slot[1].ev_p1) == cudaSuccess;
        po = po && cudaEventCreate(&s->ev_pre) == cudaSuccess && cudaEventCreate(&s->ev_c0) == cudaSuccess && cudaEventCreate(&s->ev_i0) == cudaSuccess;
        for (int k = 0; k < md_bam_stream::PUSH_PARTS && po; ++k)
            po = cudaStreamCreateWithPriority(&s->si[k], cudaStreamNonBlocking, prio_lo) == cudaSuccess && cudaEventCreate(&s->ev_copy[k]) == cudaSuccess && cudaEventCreate(&s->ev_inf[k]) == cudaSuccess;
        s->parts_ok = po;
    }
    return s;
}
static void bam_join(md_bam_stream *s) { if (s->worker.joinable()) s->worker.join(); }
extern "C" void md_bam_close(md_bam_stream *s) {
    if (!s) return;
    bam_join(s);
    cudaSetDevice(s->c->device); sync_all(s->c); if (s->sd) cudaStreamSynchronize(s->sd);
    if (getenv("MD_TIMING"))
        fprintf(stderr, "[md-timing] device decode: %llu segments, %.1f MB compressed -> %.1f MB, %llu records; H2D %.2f ms, inflate %.2f ms, record chains %.2f ms, offsets+heads+runs %.2f ms; %llu tiles: sizes+scan %.2f ms, gather %.2f ms\n",
                (unsigned long long) s->n_push, s->comp_total / 1e6, s->infl_total / 1e6, (unsigned long long) s->rec_total, s->t_push[0], s->t_push[1], s->t_push[2], s->t_push[3], (unsigned long long) s->n_tiles, s->t_tile[0], s->t_tile[1]);
    for (int k = 0; k < 2; ++k) {
        BamSlot &S = s->slot[k];
        DevBuf *bufs[] = {&S.comp, &S.blk, &S.uoff, &S.ubuf, &S.scan, &S.cnt, &S.base, &S.rec_off, &S.tid, &S.pos, &S.rend, &S.runs, &S.small, &S.cub_tmp};
        for (DevBuf *b : bufs) b->release();
        if (S.h_small) cudaFreeHost(S.h_small);
    }
    DevBuf *bufs[] = {&s->cub_tmp, &s->sz, &s->off, &s->tile[0].buf, &s->tile[1].buf};
    for (DevBuf *b : bufs) b->release();
    if (s->h_tot) cudaFreeHost(s->h_tot);
    if (s->sd) cudaStreamDestroy(s->sd);
    if (s->sc) { cudaStreamSynchronize(s->sc); cudaStreamDestroy(s->sc); }
    for (int k = 0; k < 2; ++k) { if (s->slot[k].ev_p0) cudaEventDestroy(s->slot[k].ev_p0); if (s->slot[k].ev_p1) cudaEventDestroy(s->slot[k].ev_p1); }
    for (int k = 0; k < md_bam_stream::PUSH_PARTS; ++k) { if (s->si[k]) cudaStreamDestroy(s->si[k]); if (s->ev_copy[k]) cudaEventDestroy(s->ev_copy[k]); if (s->ev_inf[k]) cudaEventDestroy(s->ev_inf[k]); }
    if (s->ev_pre) cudaEventDestroy(s->ev_pre);
    if (s->ev_c0) cudaEventDestroy(s->ev_c0);
    if (s->ev_i0) cudaEventDestroy(s->ev_i0);
    delete s;
}
extern "C" void md_bam_reset(md_bam_stream *s) {       // after a seek: forget the straddling record and the carried reads
    bam_join(s); s->inflight = false;
    if (s->sc) { cudaSetDevice(s->c->device); cudaStreamSynchronize(s->sc); }
    s->slot[0].pref_ptr = s->slot[1].pref_ptr = nullptr;
    s->have_segment = false; s->slot[0].leftover = s->slot[1].leftover = 0; s->tile[0].valid = s->tile[1].valid = false;
}

static const uint32_t BAM_MAX_RUNS = 1u << 16;

// MD_TIMING=1: per-stage device times of the decode (CUDA events on the stream the work is queued on)
struct BamTimer {
    bool on; cudaStream_t st; cudaEvent_t ev[8]; int n = 0;
    explicit BamTimer(cudaStream_t s, bool always = false) : on(always || getenv("MD_TIMING") != nullptr), st(s) { if (on) for (auto &e : ev) cudaEventCreate(&e); }
    ~BamTimer() { if (on) for (auto &e : ev) cudaEventDestroy(e); }
    void tick() { if (on && n < 8) cudaEventRecord(ev[n++], st); }
    void add(double *acc) { if (!on || n < 2) return; cudaEventSynchronize(ev[n - 1]); for (int k = 0; k + 1 < n; ++k) { float ms = 0; cudaEventElapsedTime(&ms, ev[k], ev[k + 1]); acc[k] += ms; } }
};

#define PCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { S.err = std::string(#call) + ": " + cudaGetErrorString(e_); return -100; } } while (0)
// Everything one segment needs, queued on the decode stream; runs on the helper thread.  `P` (may be null) is the slot of
// the previous segment: the record that straddles in is copied from its inflated bytes.
static int bam_push_impl(md_bam_stream *s, BamSlot &S, const BamSlot *P, const void *comp, uint64_t comp_bytes, const md_bgzf_block *blocks, uint32_t n_blocks, uint32_t skip) {
    md_ctx *c = s->c;
    PCK(cudaSetDevice(c->device));
    cudaStream_t st = s->sd;
    memset(&S.sum, 0, sizeof S.sum); S.runs_host.clear(); S.n_records = 0; S.launches = 0;
    const uint64_t carry_in = P ? P->leftover : 0;
    std::vector<unsigned long long> uoff(n_blocks + 1);
    uint64_t tot = 0;
    for (uint32_t b = 0; b < n_blocks; ++b) {
        if (blocks[b].comp_off + blocks[b].comp_len > comp_bytes) { S.err = "md_bam_push: block outside the buffer"; return -2; }
        uoff[b] = BAM_HEADROOM + tot; tot += blocks[b].isize;
    }
    uoff[n_blocks] = BAM_HEADROOM + tot;
    if (carry_in > BAM_HEADROOM) { S.err = "md_bam_push: a record larger than 64 MB straddles two segments"; return -2; }
    const uint64_t D0 = BAM_HEADROOM - carry_in, U = BAM_HEADROOM + tot;
    if (tot + carry_in >= ((uint64_t) 1 << 32) - BAM_HEADROOM) { S.err = "md_bam_push: segment inflates to more than 4 GB"; return -2; }
    if (S.comp.reserve(comp_bytes + 1024) || S.blk.reserve((size_t) n_blocks * sizeof(md_bgzf_block) + 16) || S.uoff.reserve((size_t)(n_blocks + 1) * 8) ||
        S.ubuf.reserve(U + 64) || S.scan.reserve((size_t) n_blocks * sizeof(BlockScan) + 16) || S.cnt.reserve((size_t) n_blocks * 4 + 16) || S.base.reserve((size_t) n_blocks * 4 + 16) ||
        S.runs.reserve((size_t) BAM_MAX_RUNS * sizeof(md_bam_run))) { S.err = "md_bam_push: out of device memory"; return -100; }
    BamTimer tm(st, true); tm.tick();
    uint32_t *d_small = (uint32_t *) S.small.p;       // [0] n_runs [1] err [2] bad [3] last_pos [4,5] final_exit
    const uint8_t *u = (const uint8_t *) S.ubuf.p;
    const unsigned long long first = D0 + (carry_in ? 0 : skip);
    // the straddling record's first bytes go in front of the new data
    if (carry_in) PCK(cudaMemcpyAsync((uint8_t *) S.ubuf.p + D0, (const uint8_t *) P->ubuf.p + P->leftover_from, carry_in, cudaMemcpyDeviceToDevice, st));
    PCK(cudaMemsetAsync(S.small.p, 0, 256, st));
    const int NP = md_bam_stream::PUSH_PARTS;
    // the compressed bytes may already be here (md_bam_prefetch with the same buffer); a prefetch of something else is waited out
    const bool prefetched = S.pref_ptr && S.pref_ptr == comp && S.pref_bytes == comp_bytes;
    if (S.pref_ptr) PCK(cudaStreamWaitEvent(st, S.ev_p1, 0));
    const bool had_prefetch = S.pref_ptr != nullptr;
    S.pref_ptr = nullptr; S.pref_bytes = 0;
    const bool in_parts = s->parts_ok && s->use_parts && !prefetched && n_blocks >= 1024;
    if (!in_parts) {
        if (!prefetched) {
            PCK(cudaMemcpyAsync(S.comp.p, comp, comp_bytes, cudaMemcpyHostToDevice, st));
            PCK(cudaMemsetAsync((uint8_t *) S.comp.p + comp_bytes, 0, 1024, st));     // the decoder stages the stream in 256-byte chunks: it reads up to two chunks past a stream's end
        }
        PCK(cudaMemcpyAsync(S.blk.p, blocks, (size_t) n_blocks * sizeof(md_bgzf_block), cudaMemcpyHostToDevice, st));
        PCK(cudaMemcpyAsync(S.uoff.p, uoff.data(), (size_t)(n_blocks + 1) * 8, cudaMemcpyHostToDevice, st));
        tm.tick();
        if (n_blocks) inflate_kernel<<<(n_blocks + INF_WARPS - 1) / INF_WARPS, INF_WARPS * 32, INF_SMEM, st>>>((const uint8_t *) S.comp.p, (const md_bgzf_block *) S.blk.p, (const unsigned long long *) S.uoff.p, (uint8_t *) S.ubuf.p, 0u, n_blocks, (int *)(d_small + 1));
        tm.tick();
    } else {
        // piece p = blocks [bp[p], bp[p+1]) and the bytes [lo[p], lo[p+1]) of the buffer, cut where a block's stream begins.  A decoder
        // reads up to 3 chunks of 256 bytes beyond its stream's end (never using them), so every copy runs 1 KB into the next
        // piece: the bytes a launch touches have all arrived when its event fires (copies are in order on `sc`).
        uint32_t bp[md_bam_stream::PUSH_PARTS + 1]; uint64_t lo[md_bam_stream::PUSH_PARTS + 1];
        bp[0] = 0; lo[0] = 0; bp[NP] = n_blocks; lo[NP] = comp_bytes;
        for (int q = 1; q < NP; ++q) {
            const uint64_t want = comp_bytes * (uint64_t) q / NP;
            uint32_t a = bp[q - 1], z = n_blocks;                  // first block whose stream starts at or beyond `want`
            while (a < z) { const uint32_t m = (a + z) / 2; if (blocks[m].comp_off < want) a = m + 1; else z = m; }
            bp[q] = a; lo[q] = a < n_blocks ? blocks[a].comp_off : comp_bytes;
        }
        PCK(cudaEventRecord(s->ev_pre, st));                       // the error flag is cleared, the carried bytes are queued
        PCK(cudaEventRecord(s->ev_c0, s->sc));
        PCK(cudaMemsetAsync((uint8_t *) S.comp.p + comp_bytes, 0, 1024, s->sc));
        PCK(cudaMemcpyAsync(S.blk.p, blocks, (size_t) n_blocks * sizeof(md_bgzf_block), cudaMemcpyHostToDevice, s->sc));
        PCK(cudaMemcpyAsync(S.uoff.p, uoff.data(), (size_t)(n_blocks + 1) * 8, cudaMemcpyHostToDevice, s->sc));
        for (int q = 0; q < NP; ++q) {
            const uint64_t hi = std::min<uint64_t>(comp_bytes, lo[q + 1] + 1024);
            if (hi > lo[q]) PCK(cudaMemcpyAsync((uint8_t *) S.comp.p + lo[q], (const uint8_t *) comp + lo[q], hi - lo[q], cudaMemcpyHostToDevice, s->sc));
            PCK(cudaEventRecord(s->ev_copy[q], s->sc));
        }
        for (int q = 0; q < NP; ++q) {
            PCK(cudaStreamWaitEvent(s->si[q], s->ev_pre, 0));
            PCK(cudaStreamWaitEvent(s->si[q], s->ev_copy[q], 0));
            if (q == 0) PCK(cudaEventRecord(s->ev_i0, s->si[0]));
            if (bp[q + 1] > bp[q])
                inflate_kernel<<<(bp[q + 1] - bp[q] + INF_WARPS - 1) / INF_WARPS, INF_WARPS * 32, INF_SMEM, s->si[q]>>>((const uint8_t *) S.comp.p, (const md_bgzf_block *) S.blk.p, (const unsigned long long *) S.uoff.p, (uint8_t *) S.ubuf.p, bp[q], bp[q + 1], (int *)(d_small + 1));
            PCK(cudaEventRecord(s->ev_inf[q], s->si[q]));
            S.launches += 1;
        }
        for (int q = 0; q < NP; ++q) PCK(cudaStreamWaitEvent(st, s->ev_inf[q], 0));
        tm.tick(); tm.tick();                                      // (copy and inflate are timed by their own events below)
    }
    if (n_blocks) {
        const uint32_t g = (n_blocks + 127) / 128;
        // block 0's slice starts at D0 so that the straddling record is part of its chain
        unsigned long long d0 = D0;
        PCK(cudaMemcpyAsync(S.uoff.p, &d0, 8, cudaMemcpyHostToDevice, st));
        scan_blocks_kernel<<<g, 128, 0, st>>>(u, (const unsigned long long *) S.uoff.p, n_blocks, first, U, s->n_targets, (BlockScan *) S.scan.p);
        check_chain_kernel<<<g, 128, 0, st>>>((const unsigned long long *) S.uoff.p, n_blocks, first, (const BlockScan *) S.scan.p, (int *)(d_small + 2));
        fix_chain_kernel<<<1, 32, 0, st>>>(u, (const unsigned long long *) S.uoff.p, n_blocks, first, U, (BlockScan *) S.scan.p, (const int *)(d_small + 2), (unsigned long long *)(d_small + 4));
        counts_kernel<<<g, 128, 0, st>>>((const BlockScan *) S.scan.p, n_blocks, (uint32_t *) S.cnt.p);
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const uint32_t *) S.cnt.p, (uint32_t *) S.base.p, (int) n_blocks, st);
        if (S.cub_tmp.reserve(tmp + 256)) { S.err = "md_bam_push: out of device memory"; return -100; }
        cub::DeviceScan::ExclusiveSum(S.cub_tmp.p, tmp, (const uint32_t *) S.cnt.p, (uint32_t *) S.base.p, (int) n_blocks, st);
        S.launches += in_parts ? 5 : 6;
    }
    tm.tick();
    // number of records = base[last] + cnt[last]
    uint32_t last2[2] = {0, 0};
    if (n_blocks) {
        PCK(cudaMemcpyAsync(&last2[0], (uint32_t *) S.base.p + (n_blocks - 1), 4, cudaMemcpyDeviceToHost, st));
        PCK(cudaMemcpyAsync(&last2[1], (uint32_t *) S.cnt.p + (n_blocks - 1), 4, cudaMemcpyDeviceToHost, st));
    }
    PCK(cudaMemcpyAsync(S.h_small, d_small, 32, cudaMemcpyDeviceToHost, st));
    PCK(cudaStreamSynchronize(st));
    if (S.h_small[1]) { char msg[128]; snprintf(msg, sizeof msg, "BGZF inflate failed on the device (block %u, code -%u)", S.h_small[1] >> 4, S.h_small[1] & 15u); S.err = msg; return -5; }
    const uint32_t n = last2[0] + last2[1];
    unsigned long long final_exit; memcpy(&final_exit, S.h_small + 4, 8);
    if (!n_blocks) final_exit = first;
    S.U = U; S.D0 = D0; S.n_records = n;
    S.leftover_from = final_exit; S.leftover = U - final_exit;
    if (n) {
        if (S.rec_off.reserve((size_t) n * 8 + 16) || S.tid.reserve((size_t) n * 4 + 16) || S.pos.reserve((size_t) n * 4 + 16) || S.rend.reserve((size_t) n * 4 + 16)) { S.err = "md_bam_push: out of device memory"; return -100; }
        const uint32_t g = (n_blocks + 127) / 128, gr = (n + 255) / 256;
        fill_offsets_kernel<<<g, 128, 0, st>>>(u, (const unsigned long long *) S.uoff.p, n_blocks, U, (const BlockScan *) S.scan.p, (const uint32_t *) S.base.p, (unsigned long long *) S.rec_off.p);
        head_kernel<<<gr, 256, 0, st>>>(u, (const unsigned long long *) S.rec_off.p, n, (int32_t *) S.tid.p, (int32_t *) S.pos.p, (int32_t *) S.rend.p, (int *)(d_small + 1));
        runs_kernel<<<gr, 256, 0, st>>>((const int32_t *) S.tid.p, (const int32_t *) S.pos.p, n, (md_bam_run *) S.runs.p, BAM_MAX_RUNS, d_small, (int32_t *)(d_small + 3));
        S.launches += 3;
        tm.tick();
        PCK(cudaMemcpyAsync(S.h_small, d_small, 32, cudaMemcpyDeviceToHost, st));
        PCK(cudaStreamSynchronize(st));
        if (S.h_small[1]) { S.err = "malformed BAM record"; return -5; }
        const uint32_t nr = S.h_small[0];
        if (nr > BAM_MAX_RUNS) { S.err = "md_bam_push: more than 65536 contig runs in one segment"; return -2; }
        S.runs_host.resize(nr);
        PCK(cudaMemcpyAsync(S.runs_host.data(), S.runs.p, (size_t) nr * sizeof(md_bam_run), cudaMemcpyDeviceToHost, st));
        PCK(cudaStreamSynchronize(st));
        std::sort(S.runs_host.begin(), S.runs_host.end(), [](const md_bam_run &a, const md_bam_run &b) { return a.start < b.start; });
        for (uint32_t k = 0; k < nr; ++k) {
            md_bam_run &r = S.runs_host[k];
            const uint32_t nxt = k + 1 < nr ? S.runs_host[k + 1].start : n;
            r.n = nxt - r.start;
            r.last_pos = k + 1 < nr ? S.runs_host[k + 1].prev_last_pos : (int32_t) S.h_small[3];
        }
    }
    {
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        tm.add(acc);
        if (prefetched) { float h = 0; cudaEventElapsedTime(&h, S.ev_p0, S.ev_p1); acc[0] = h; }      // the copy ran on the copy stream, earlier
        (void) had_prefetch;
        if (in_parts) {                                            // everything has completed: the stream was synchronised above
            float h = 0, f = 0;
            cudaEventElapsedTime(&h, s->ev_c0, s->ev_copy[NP - 1]);
            cudaEventElapsedTime(&f, s->ev_i0, tm.ev[1]);          // first launch's start -> all pieces inflated
            acc[0] = h; acc[1] = f;
        }
        for (int k = 0; k < 8; ++k) s->t_push[k] += acc[k];
    }
    s->n_push++; s->comp_total += comp_bytes; s->infl_total += tot; s->rec_total += n;
    PCK(cudaGetLastError());
    S.sum.n_records = n; S.sum.n_runs = (uint32_t) S.runs_host.size(); S.sum.inflated_bytes = tot; S.sum.leftover_bytes = S.leftover;
    return 0;
}

// While a push is in flight: start copying the compressed bytes of the segment that will be pushed AFTER it (on the copy stream,
// into the slot that push will use — its previous contents were inflated two pushes ago).  md_bam_push_begin() with the same
// buffer and size then skips its copy, so the host-to-device transfer of segment k+2 hides behind the inflate of segment k+1.
// comp == NULL: wait for outstanding prefetches and forget them (before the caller releases its buffers).
extern "C" int md_bam_prefetch(md_bam_stream *s, const void *comp, uint64_t comp_bytes) {
    if (!s || !s->prefetch_ok) return 0;
    cudaSetDevice(s->c->device);
    if (!comp) {
        cudaStreamSynchronize(s->sc);
        if (!s->inflight) { s->slot[0].pref_ptr = s->slot[1].pref_ptr = nullptr; }
        else s->slot[s->target ^ 1].pref_ptr = nullptr;
        return 0;
    }
    if (!s->inflight) return 0;                                    // only defined relative to a push in flight; otherwise a no-op
    BamSlot &S = s->slot[s->target ^ 1];
    if (S.pref_ptr) return 0;
    if (S.comp.reserve(comp_bytes + 1024)) { g_err = "md_bam_prefetch: out of device memory"; return -100; }
    if (cudaEventRecord(S.ev_p0, s->sc) != cudaSuccess ||
        cudaMemcpyAsync(S.comp.p, comp, comp_bytes, cudaMemcpyHostToDevice, s->sc) != cudaSuccess ||
        cudaMemsetAsync((uint8_t *) S.comp.p + comp_bytes, 0, 1024, s->sc) != cudaSuccess ||
        cudaEventRecord(S.ev_p1, s->sc) != cudaSuccess) { g_err = "md_bam_prefetch: copy failed"; return -100; }
    S.pref_ptr = comp; S.pref_bytes = comp_bytes;
    return 0;
}

// Start copying + decoding a segment in the background; the caller's buffers must stay valid until md_bam_push_end().
extern "C" int md_bam_push_begin(md_bam_stream *s, const void *comp, uint64_t comp_bytes, const md_bgzf_block *blocks, uint32_t n_blocks, uint32_t skip) {
    if (s->inflight) { g_err = "md_bam_push_begin: a push is already in flight"; return -4; }
    bam_join(s);
    const int tgt = s->have_segment ? (s->cur_slot ^ 1) : 0;
    BamSlot *S = &s->slot[tgt]; const BamSlot *P = s->have_segment ? &s->slot[s->cur_slot] : nullptr;
    s->target = tgt; s->inflight = true;
    S->rc = 0; S->err.clear();
    s->worker = std::thread([=] { S->rc = bam_push_impl(s, *S, P, comp, comp_bytes, blocks, n_blocks, skip); });
    return 0;
}
// Wait for it; from here on the runs / tiles refer to that segment.
extern "C" int md_bam_push_end(md_bam_stream *s, md_bam_summary *out) {
    if (!s->inflight) { g_err = "md_bam_push_end: nothing in flight"; return -4; }
    bam_join(s); s->inflight = false;
    BamSlot &S = s->slot[s->target];
    s->c->launches += S.launches;
    {   // totals for md_ctx_totals (the helper thread has finished: t_push is quiescent)
        md_totals &T = s->c->tot;
        T.push_h2d_ms = s->t_push[0]; T.inflate_ms = s->t_push[1]; T.frame_ms = s->t_push[2] + s->t_push[3];
        T.comp_bytes = s->comp_total; T.inflated_bytes = s->infl_total;
    }
    if (S.rc) { g_err = S.err.empty() ? "md_bam_push failed" : S.err; return S.rc; }
    s->cur_slot = s->target; s->have_segment = true;
    if (out) *out = S.sum;
    return 0;
}
extern "C" int md_bam_push(md_bam_stream *s, const void *comp, uint64_t comp_bytes, const md_bgzf_block *blocks, uint32_t n_blocks, uint32_t skip, md_bam_summary *out) {
    int rc = md_bam_push_begin(s, comp, comp_bytes, blocks, n_blocks, skip);
    if (rc) return rc;
    return md_bam_push_end(s, out);
}
extern "C" int md_bam_get_runs(md_bam_stream *s, md_bam_run *runs, uint32_t cap) {
    if (!s->have_segment) return 0;
    const std::vector<md_bam_run> &R = s->slot[s->cur_slot].runs_host;
    for (uint32_t k = 0; k < R.size() && k < cap; ++k) runs[k] = R[k];
    return (int) R.size();
}

// Build the tile (carried reads of the previous tile of this contig + records [run.start, run.start+run.n) of the last segment)
// and leave it device-resident in s->tile[s->cur].
static int bam_build_tile(md_bam_stream *s, int run, const md_tile_desc *t, uint32_t keep_hi) {
    md_ctx *c = s->c; Lane *L = &c->lanes[0]; cudaStream_t st = L->stream;
    TileArena &P = s->tile[s->cur], &N = s->tile[s->cur ^ 1];
    TileSrc S; memset(&S, 0, sizeof S);
    const BamSlot &B = s->slot[s->cur_slot];
    S.u = (const uint8_t *) B.ubuf.p; S.rec_off = (const unsigned long long *) B.rec_off.p; S.pos = (const int32_t *) B.pos.p; S.rend = (const int32_t *) B.rend.p;
    if (run >= 0) {
        if (!s->have_segment || (size_t) run >= B.runs_host.size()) { g_err = "md_bam: no such run"; return -2; }
        const md_bam_run &r = B.runs_host[(size_t) run];
        if (r.tid != t->tid) { g_err = "md_bam: tile and run are on different contigs"; return -2; }
        S.r0 = r.start; S.n_own = r.n;
    }
    const bool carry = P.valid && P.tid == t->tid && P.cut == t->beg;
    if (carry) {
        const DevReads &V = P.view;
        S.prev.pos = V.pos; S.prev.flag = V.flag; S.prev.mapq = V.mapq; S.prev.aux = V.aux; S.prev.l_qseq = V.l_qseq; S.prev.cigar_off = V.cigar_off; S.prev.seq_off = V.seq_off;
        S.prev.qual_off = V.qual_off; S.prev.frag_key = V.frag_key; S.prev.cigar = V.cigar; S.prev.seq = V.seq; S.prev.qual = V.qual; S.prev.name_chk = V.name_chk;
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
    size_t szs[14] = {n * 4, n * 2, n, n, n * 4, (n + 1) * 4, n * 4, n * 4, n * 8, n * 4, (size_t) tot.y * 4, (size_t) tot.z * 4, (size_t) tot.w * 8, n * 4};
    size_t offs[14], total = 0;
    for (int k = 0; k < 14; ++k) { offs[k] = total; total += al256(szs[k] + 16); }
    if (N.buf.reserve(total)) return -100;
    unsigned char *base = (unsigned char *) N.buf.p;
    TileDst D;
    D.pos = (int32_t *)(base + offs[0]); D.flag = (uint16_t *)(base + offs[1]); D.mapq = base + offs[2]; D.aux = base + offs[3]; D.l_qseq = (uint32_t *)(base + offs[4]);
    D.cigar_off = (uint32_t *)(base + offs[5]); D.seq_off = (uint32_t *)(base + offs[6]); D.qual_off = (uint32_t *)(base + offs[7]); D.frag_key = (uint64_t *)(base + offs[8]);
    D.rend = (int32_t *)(base + offs[9]); D.cigar = (uint32_t *)(base + offs[10]); D.seq = (uint32_t *)(base + offs[11]); D.qual = (uint64_t *)(base + offs[12]); D.name_chk = (uint32_t *)(base + offs[13]);
    if (m) { tile_gather_kernel<<<(m + 255) / 256, 256, 0, st>>>(S, (const Sz4 *) s->sz.p, (const Sz4 *) s->off.p, D); c->launches += 1; }
    else CK(cudaMemsetAsync(D.cigar_off, 0, 4, st));
    DevReads &v = N.view; memset(&v, 0, sizeof v);
    v.n = (uint32_t) n; v.seq_words = tot.z; v.qual_words = tot.w; v.qbits = 8;
    v.pos = D.pos; v.flag = D.flag; v.mapq = D.mapq; v.aux = D.aux; v.l_qseq = D.l_qseq; v.cigar_off = D.cigar_off; v.seq_off = D.seq_off; v.qual_off = D.qual_off;
    v.frag_key = D.frag_key; v.cigar = D.cigar; v.seq = D.seq; v.qual = D.qual; v.name_chk = D.name_chk;
    tm.tick(); tm.add(s->t_tile); s->n_tiles++;
    N.rend = D.rend; N.n = (uint32_t) n; N.n_cigar = tot.y; N.tid = t->tid; N.cut = t->end; N.valid = true;
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
    L->last_ncigar = T.n_cigar;
    rc = run_pipeline(c, L, t, T.view, mbias);
    if (rc) return rc;
    rc = finish_counters(c, L, stats);
    if (rc) return rc;
    if (!mbias) rc = fetch_sorted(c, L, calls, cap, nullptr);
    CK(cudaEventRecord(L->ev[4], L->stream));
    CK(cudaStreamSynchronize(L->stream));
    collect_timing(c, L);
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
