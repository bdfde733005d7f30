// TEST INFRASTRUCTURE — CPU emulation of the device-side BAM decode entry points (include/mdgpu.h: md_bam_*).
// It runs the very same per-thread bodies as the CUDA kernels (methyldackel_b200/csrc/{inflate_hd,bamrec_hd,bamdev_hd}.h)
// in plain loops, so that the CPU test tier covers the deflate decoder, the record-chain logic, the tile assembly with
// carried reads and the CLI's segment orchestration.  The tile it builds is handed to a callback (the oracle port in the
// tests).  Never linked into the product.
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <string>
#include <map>
#include <mutex>
#include <algorithm>
#include "../../include/mdgpu.h"
#include "../../methyldackel_b200/csrc/bamdev_hd.h"

typedef int (*emu_extract_cb)(void *be, const md_tile_desc *, const md_reads_soa *, md_call *, uint64_t, md_tile_stats *);
typedef int (*emu_mbias_cb)(void *be, const md_tile_desc *, const md_reads_soa *, md_tile_stats *);

namespace {
const size_t HEADROOM = 1 << 20;
struct Tile {
    std::vector<int32_t> pos, rend; std::vector<uint16_t> flag; std::vector<uint8_t> mapq, aux; std::vector<uint32_t> l_qseq, cigar_off, seq_off, qual_off, cigar, seq;
    std::vector<uint64_t> frag_key, qual; std::vector<uint32_t> name_chk; uint32_t n = 0; int32_t tid = -1; uint32_t cut = 0; bool valid = false;
};
struct Seg {
    std::vector<uint8_t> ubuf; uint64_t U = 0, D0 = 0, leftover_from = 0, leftover = 0;
    std::vector<unsigned long long> rec_off; std::vector<int32_t> tid, pos, rend; std::vector<md_bam_run> runs;
};
struct Emu {
    int32_t n_targets; emu_extract_cb ex; emu_mbias_cb mb; void *be;
    Seg seg[2]; int cur_seg = 0; bool have = false;          // two slots, as the device library: tiles read seg[cur_seg]
    bool pending = false; int pending_rc = 0, target = 0; md_bam_summary pending_sum;
    // The device reads the caller's buffers until md_bam_push_end() returns, so the emulation reads them as LATE as it may: the
    // decode of a two-phase push happens in emu_bam_push_end().  A caller that recycles a buffer too early is then seen.
    const void *p_comp = nullptr; uint64_t p_bytes = 0; const md_bgzf_block *p_blocks = nullptr; uint32_t p_nblocks = 0, p_skip = 0;
    // md_bam_prefetch: the bytes may be copied at any time between the prefetch and the end of the push that uses them, so they
    // must read the same at both ends
    const void *pf_ptr = nullptr; uint64_t pf_bytes = 0, pf_sum = 0, n_prefetch_used = 0;
    Tile tile[2]; int cur = 0; std::string err;
    uint64_t n_fix = 0;
};
std::string g_emu_err;
}

extern "C" const char *emu_last_error() { return g_emu_err.c_str(); }
extern "C" void *emu_bam_open(int32_t n_targets, emu_extract_cb ex, emu_mbias_cb mb, void *be) { Emu *e = new Emu(); e->n_targets = n_targets; e->ex = ex; e->mb = mb; e->be = be; return e; }
extern "C" void emu_bam_close(void *s) { delete (Emu *) s; }
extern "C" void emu_bam_reset(void *s) { Emu *e = (Emu *) s; e->seg[0].leftover = e->seg[1].leftover = 0; e->have = false; e->pending = false; e->tile[0].valid = e->tile[1].valid = false; }
extern "C" uint64_t emu_bam_fixups(void *s) { return ((Emu *) s)->n_fix; }

static int push_impl(Emu *e, Seg *s, const Seg *P, const void *comp, uint64_t comp_bytes, const md_bgzf_block *blocks, uint32_t n_blocks, uint32_t skip, md_bam_summary *out) {
    memset(out, 0, sizeof *out);
    const uint64_t carry_in = P ? P->leftover : 0;
    std::vector<unsigned long long> uoff(n_blocks + 1);
    uint64_t tot = 0;
    for (uint32_t b = 0; b < n_blocks; ++b) { if (blocks[b].comp_off + blocks[b].comp_len > comp_bytes) { g_emu_err = "block outside the buffer"; return -2; } uoff[b] = HEADROOM + tot; tot += blocks[b].isize; }
    uoff[n_blocks] = HEADROOM + tot;
    if (carry_in > HEADROOM) { g_emu_err = "straddling record too large for the emulation"; return -2; }
    const uint64_t D0 = HEADROOM - carry_in, U = HEADROOM + tot;
    s->ubuf.assign(U + 64, 0);
    if (carry_in) memcpy(s->ubuf.data() + D0, P->ubuf.data() + P->leftover_from, carry_in);
    // 4-byte aligned, padded copy of the compressed bytes (the device buffer is)
    std::vector<uint32_t> cal((comp_bytes + 64 + 3) / 4, 0); memcpy(cal.data(), comp, comp_bytes);
    static thread_local mdinflate::Decoder T;
    for (uint32_t b = 0; b < n_blocks; ++b) {
        if (!blocks[b].isize) continue;
        int rc = mdinflate::inflate_block(cal.data(), blocks[b].comp_off, blocks[b].comp_len, s->ubuf.data() + uoff[b], blocks[b].isize, T);
        if (rc) { g_emu_err = "inflate failed in block " + std::to_string(b) + " code " + std::to_string(rc); return -5; }
    }
    const uint8_t *u = s->ubuf.data();
    const unsigned long long first = D0 + (carry_in ? 0 : skip);
    std::vector<mdbam::BlockScan> sc(n_blocks);
    unsigned long long final_exit = first;
    std::vector<uint32_t> base(n_blocks, 0);
    uint32_t n = 0;
    if (n_blocks) {
        uoff[0] = D0;
        for (uint32_t b = 0; b < n_blocks; ++b) mdbam::scan_block_body(b, u, uoff.data(), first, U, e->n_targets, sc.data());
        int bad = 0;
        for (uint32_t b = 0; b < n_blocks; ++b) if (!mdbam::check_block_body(b, uoff.data(), first, sc.data())) bad = 1;
        if (getenv("MDEMU_FORCE_FIX")) { bad = 1; for (uint32_t b = 1; b < n_blocks; b += 3) if (sc[b].guess != mdbam::NONE) { sc[b].guess += 1; } }   // sabotage guesses: the repair must restore them
        e->n_fix += (uint64_t) bad;
        mdbam::fix_chain_body(u, uoff.data(), n_blocks, first, U, sc.data(), bad, &final_exit);
        for (uint32_t b = 0; b < n_blocks; ++b) { base[b] = n; n += sc[b].guess != mdbam::NONE ? sc[b].count : 0u; }
    }
    s->U = U; s->D0 = D0; s->leftover_from = final_exit; s->leftover = U - final_exit;
    s->rec_off.assign(n, 0); s->tid.assign(n, 0); s->pos.assign(n, 0); s->rend.assign(n, 0); s->runs.clear();
    for (uint32_t b = 0; b < n_blocks; ++b) mdbam::fill_offsets_body(b, u, uoff.data(), U, sc.data(), base.data(), s->rec_off.data());
    for (uint32_t i = 0; i < n; ++i) if (!mdbam::head_body(i, u, s->rec_off.data(), s->tid.data(), s->pos.data(), s->rend.data())) { g_emu_err = "malformed BAM record"; return -5; }
    for (uint32_t i = 0; i < n; ++i) if (i == 0 || s->tid[i] != s->tid[i - 1]) { md_bam_run r; r.tid = s->tid[i]; r.start = i; r.n = 0; r.first_pos = s->pos[i]; r.last_pos = 0; r.prev_last_pos = i ? s->pos[i - 1] : 0; s->runs.push_back(r); }
    for (size_t k = 0; k < s->runs.size(); ++k) { md_bam_run &r = s->runs[k]; uint32_t nxt = k + 1 < s->runs.size() ? s->runs[k + 1].start : n; r.n = nxt - r.start; r.last_pos = s->pos[nxt - 1]; }
    out->n_records = n; out->n_runs = (uint32_t) s->runs.size(); out->inflated_bytes = tot; out->leftover_bytes = s->leftover;
    return 0;
}
static uint64_t fnv64(const void *p, uint64_t n) { const uint8_t *b = (const uint8_t *) p; uint64_t h = 1469598103934665603ull; for (uint64_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull; return h; }
extern "C" int emu_bam_push_begin(void *sv, const void *comp, uint64_t comp_bytes, const md_bgzf_block *blocks, uint32_t n_blocks, uint32_t skip) {
    Emu *e = (Emu *) sv;
    if (e->pending) { g_emu_err = "a push is already in flight"; return -4; }
    e->target = e->have ? (e->cur_seg ^ 1) : 0;
    e->p_comp = comp; e->p_bytes = comp_bytes; e->p_blocks = blocks; e->p_nblocks = n_blocks; e->p_skip = skip;
    if (e->pf_ptr && (e->pf_ptr != comp || e->pf_bytes != comp_bytes)) e->pf_ptr = nullptr;       // a prefetch is for the very next push
    e->pending = true;
    return 0;
}
extern "C" int emu_bam_push_end(void *sv, md_bam_summary *out) {
    Emu *e = (Emu *) sv;
    if (!e->pending) { g_emu_err = "nothing in flight"; return -4; }
    e->pending = false;
    if (e->pf_ptr && e->pf_ptr == e->p_comp && e->pf_bytes == e->p_bytes) {
        if (fnv64(e->p_comp, e->p_bytes) != e->pf_sum) { g_emu_err = "a prefetched buffer changed before the push that uses it ended"; return -9; }
        ++e->n_prefetch_used;
        e->pf_ptr = nullptr;
    }
    e->pending_rc = push_impl(e, &e->seg[e->target], e->have ? &e->seg[e->cur_seg] : nullptr, e->p_comp, e->p_bytes, e->p_blocks, e->p_nblocks, e->p_skip, &e->pending_sum);
    if (e->pending_rc) return e->pending_rc;
    e->cur_seg = e->target; e->have = true;
    if (out) *out = e->pending_sum;
    return 0;
}
extern "C" int emu_bam_push(void *sv, const void *comp, uint64_t comp_bytes, const md_bgzf_block *blocks, uint32_t n_blocks, uint32_t skip, md_bam_summary *out) {
    int rc = emu_bam_push_begin(sv, comp, comp_bytes, blocks, n_blocks, skip);
    return rc ? rc : emu_bam_push_end(sv, out);
}
// md_bam_prefetch: only between push_begin and push_end; comp == NULL drops an outstanding prefetch
extern "C" int emu_bam_prefetch(void *sv, const void *comp, uint64_t comp_bytes) {
    Emu *e = (Emu *) sv;
    if (!comp) { e->pf_ptr = nullptr; return 0; }
    if (const char *f = getenv("MDEMU_FAIL_PREFETCH")) { static int calls = 0; if (++calls == atoi(f)) { g_emu_err = "injected prefetch failure"; return -100; } }
    if (!e->pending || e->pf_ptr) return 0;
    e->pf_ptr = comp; e->pf_bytes = comp_bytes; e->pf_sum = fnv64(comp, comp_bytes);
    return 0;
}
extern "C" uint64_t emu_bam_prefetch_used(void *sv) { return ((Emu *) sv)->n_prefetch_used; }
// stand-ins for the page-locked allocator: freed memory is overwritten first, so that a reader of a released buffer sees garbage
static std::map<void *, size_t> g_pins; static std::mutex g_pins_m; static uint64_t g_pins_total = 0;
extern "C" void *emu_alloc_pinned(size_t n) { void *p = malloc(n ? n : 1); if (p) { memset(p, 0xA5, n); std::lock_guard<std::mutex> g(g_pins_m); g_pins[p] = n; ++g_pins_total; } return p; }
extern "C" void emu_free_pinned(void *p) { if (!p) return; size_t n = 0; { std::lock_guard<std::mutex> g(g_pins_m); auto it = g_pins.find(p); if (it != g_pins.end()) { n = it->second; g_pins.erase(it); } } memset(p, 0xDD, n); free(p); }
extern "C" uint64_t emu_pinned_stats(uint64_t *live) { std::lock_guard<std::mutex> g(g_pins_m); if (live) *live = g_pins.size(); return g_pins_total; }
extern "C" int emu_bam_get_runs(void *sv, md_bam_run *runs, uint32_t cap) { Emu *e = (Emu *) sv; if (!e->have) return 0; const Seg &s = e->seg[e->cur_seg]; for (size_t k = 0; k < s.runs.size() && k < cap; ++k) runs[k] = s.runs[k]; return (int) s.runs.size(); }

static int build(Emu *s, int run, const md_tile_desc *t, uint32_t keep_hi) {
    Tile &P = s->tile[s->cur], &N = s->tile[s->cur ^ 1];
    mdbam::TileSrc S; memset(&S, 0, sizeof S);
    const Seg &G = s->seg[s->cur_seg];
    S.u = G.ubuf.data(); S.rec_off = G.rec_off.data(); S.pos = G.pos.data(); S.rend = G.rend.data();
    if (run >= 0) {
        if (!s->have || (size_t) run >= G.runs.size()) { g_emu_err = "no such run"; return -2; }
        const md_bam_run &r = G.runs[(size_t) run];
        if (r.tid != t->tid) { g_emu_err = "tile and run are on different contigs"; return -2; }
        S.r0 = r.start; S.n_own = r.n;
    }
    if (P.valid && P.tid == t->tid && P.cut == t->beg) {
        S.prev.pos = P.pos.data(); S.prev.flag = P.flag.data(); S.prev.mapq = P.mapq.data(); S.prev.aux = P.aux.data(); S.prev.l_qseq = P.l_qseq.data(); S.prev.cigar_off = P.cigar_off.data();
        S.prev.seq_off = P.seq_off.data(); S.prev.qual_off = P.qual_off.data(); S.prev.frag_key = P.frag_key.data(); S.prev.cigar = P.cigar.data(); S.prev.seq = P.seq.data(); S.prev.qual = P.qual.data(); S.prev.name_chk = P.name_chk.data();
        S.prev_rend = P.rend.data(); S.n_prev = P.n;
    }
    S.keep_lo = t->beg; S.keep_hi = keep_hi;
    const uint32_t m = S.n_prev + S.n_own;
    std::vector<mdbam::Sz4> sz(m), off(m);
    mdbam::Sz4 acc; acc.x = acc.y = acc.z = acc.w = 0;
    for (uint32_t e = 0; e < m; ++e) { sz[e] = mdbam::tile_sizes_body(e, S); off[e] = acc; acc.x += sz[e].x; acc.y += sz[e].y; acc.z += sz[e].z; acc.w += sz[e].w; }
    const size_t n = acc.x;
    N.pos.assign(n, 0); N.rend.assign(n, 0); N.flag.assign(n, 0); N.mapq.assign(n, 0); N.aux.assign(n, 0); N.l_qseq.assign(n, 0); N.cigar_off.assign(n + 1, 0); N.seq_off.assign(n, 0); N.qual_off.assign(n, 0);
    N.frag_key.assign(n, 0); N.name_chk.assign(n + 1, 0); N.cigar.assign(acc.y + 4, 0); N.seq.assign(acc.z + 4, 0); N.qual.assign(acc.w + 4, 0);
    mdbam::TileDst D; D.pos = N.pos.data(); D.flag = N.flag.data(); D.mapq = N.mapq.data(); D.aux = N.aux.data(); D.l_qseq = N.l_qseq.data(); D.cigar_off = N.cigar_off.data(); D.seq_off = N.seq_off.data();
    D.qual_off = N.qual_off.data(); D.frag_key = N.frag_key.data(); D.name_chk = N.name_chk.data(); D.rend = N.rend.data(); D.cigar = N.cigar.data(); D.seq = N.seq.data(); D.qual = N.qual.data();
    for (uint32_t e = 0; e < m; ++e) mdbam::tile_gather_body(e, S, sz[e], off[e], D);
    N.cigar.resize(acc.y); N.seq.resize(acc.z); N.qual.resize(acc.w);
    N.n = (uint32_t) n; N.tid = t->tid; N.cut = t->end; N.valid = true;
    s->cur ^= 1;
    return 0;
}
static md_reads_soa view_of(Tile &T) {
    md_reads_soa v; memset(&v, 0, sizeof v);
    v.n_reads = T.n; v.n_cigar_ops = (uint32_t) T.cigar.size(); v.seq_words = T.seq.size(); v.qual_words = T.qual.size();
    v.pos = T.pos.data(); v.flag = T.flag.data(); v.mapq = T.mapq.data(); v.aux = T.aux.data(); v.l_qseq = T.l_qseq.data(); v.cigar_off = T.cigar_off.data(); v.seq_off = T.seq_off.data();
    v.qual_off = T.qual_off.data(); v.frag_key = T.frag_key.data(); v.cigar = T.cigar.data(); v.seq = T.seq.data(); v.qual = T.qual.data(); v.qual_bits = 8;
    return v;
}
extern "C" int emu_bam_extract_run(void *sv, int run, const md_tile_desc *t, uint32_t keep_hi, md_call *calls, uint64_t cap, md_tile_stats *st) {
    Emu *s = (Emu *) sv; int rc = build(s, run, t, keep_hi); if (rc) return rc;
    md_reads_soa v = view_of(s->tile[s->cur]);
    return s->ex(s->be, t, &v, calls, cap, st);
}
extern "C" int emu_bam_mbias_run(void *sv, int run, const md_tile_desc *t, uint32_t keep_hi, md_tile_stats *st) {
    Emu *s = (Emu *) sv; int rc = build(s, run, t, keep_hi); if (rc) return rc;
    md_reads_soa v = view_of(s->tile[s->cur]);
    return s->mb(s->be, t, &v, st);
}
// the tile built last (for tests that compare it with the host decoder's)
extern "C" int emu_bam_tile_view(void *sv, md_reads_soa *out) { Emu *s = (Emu *) sv; if (!s->tile[s->cur].valid) return -1; *out = view_of(s->tile[s->cur]); return 0; }

// the deflate decoder on its own (tests/test_inflate_vs_zlib.py): raw-deflate stream -> out[out_off .. out_off + out_len)
extern "C" int emu_inflate_raw(const uint8_t *comp, uint64_t comp_len, uint64_t in_off, uint8_t *out, uint32_t out_off, uint32_t out_len) {
    std::vector<uint32_t> cal((in_off + comp_len + 64 + 3) / 4, 0);
    memcpy((uint8_t *) cal.data() + in_off, comp, comp_len);
    static thread_local mdinflate::Decoder T;
    return mdinflate::inflate_block(cal.data(), in_off, comp_len, out + out_off, out_len, T);
}
