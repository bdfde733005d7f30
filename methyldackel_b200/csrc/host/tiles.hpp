// BAM records -> structure-of-arrays tiles (md_reads_soa, include/mdgpu.h).
// Replaces, for the B200 path, the record hand-off htslib does inside sam_itr_next /
// bam_plp_push (common.c:413): the fields filter_func,
// getStrand, the trims and the overlap code read are laid out column-wise so the
// device can stream them with coalesced loads.
#pragma once
#include "hostio.hpp"
#include "../../../include/mdgpu.h"
#include <memory>
#include <functional>
#include <array>

namespace mdhost {

// Allocation hooks so the CLI can back tiles with pinned (page-locked) memory obtained
// from the CUDA library while tests use plain malloc.
struct TileAlloc {
    void *(*alloc)(size_t) = nullptr;
    void (*release)(void *) = nullptr;
};

template <class T> class PodVec {   // minimal growable POD array on a TileAlloc
public:
    explicit PodVec(const TileAlloc *a = nullptr) : a_(a) {}
    ~PodVec() { free_(p_); }
    PodVec(const PodVec &) = delete;
    PodVec &operator=(const PodVec &) = delete;
    size_t size() const { return n_; }
    T *data() { return p_; }
    const T *data() const { return p_; }
    void clear() { n_ = 0; }
    T &operator[](size_t i) { return p_[i]; }
    const T &operator[](size_t i) const { return p_[i]; }
    void reserve(size_t c) {
        if (c <= cap_) return;
        size_t nc = std::max(c, cap_ * 2 + 1024);
        T *q = (T *) alloc_(nc * sizeof(T));
        if (!q) throw std::bad_alloc();
        if (n_) memcpy(q, p_, n_ * sizeof(T));
        free_(p_); p_ = q; cap_ = nc;
    }
    void push_back(const T &v) { if (n_ == cap_) reserve(n_ + 1); p_[n_++] = v; }
    T *grow(size_t k) { reserve(n_ + k); T *r = p_ + n_; n_ += k; return r; }
    void swap(PodVec &o) { std::swap(a_, o.a_); std::swap(p_, o.p_); std::swap(n_, o.n_); std::swap(cap_, o.cap_); }
private:
    void *alloc_(size_t b) { return (a_ && a_->alloc) ? a_->alloc(b) : malloc(b); }
    void free_(void *p) { if (!p) return; if (a_ && a_->release) a_->release(p); else free(p); }
    const TileAlloc *a_; T *p_ = nullptr; size_t n_ = 0, cap_ = 0;
};

struct SoaTile {
    explicit SoaTile(const TileAlloc *a = nullptr)
        : pos(a), flag(a), mapq(a), aux(a), l_qseq(a), cigar_off(a), seq_off(a), qual_off(a), frag_key(a), cigar(a), seq(a), qual(a), rend(nullptr) {}
    int32_t tid = -1;
    uint32_t beg = 0, end = 0;
    uint32_t qual_bits = 8; uint8_t qual_lut[16] = {0};      // phred encoding of qual[] (include/mdgpu.h)
    uint8_t qual_code[256] = {0};                             // inverse of qual_lut while qual_bits < 8
    uint8_t qseen[256] = {0};                                 // phred values add() has stored (alphabet of an 8-bit tile)
    PodVec<int32_t> pos; PodVec<uint16_t> flag; PodVec<uint8_t> mapq, aux; PodVec<uint32_t> l_qseq, cigar_off, seq_off, qual_off;
    PodVec<uint64_t> frag_key; PodVec<uint32_t> cigar, seq; PodVec<uint64_t> qual;
    PodVec<int32_t> rend;   // host-only: reference end of each read (for carry-over between tiles)
    size_t n() const { return pos.size(); }
    // one allocation per column up front (page-locked allocations are expensive; growth stays possible)
    void reserve_for(size_t reads, size_t read_len) {
        pos.reserve(reads); flag.reserve(reads); mapq.reserve(reads); aux.reserve(reads); l_qseq.reserve(reads); cigar_off.reserve(reads + 1); seq_off.reserve(reads); qual_off.reserve(reads);
        frag_key.reserve(reads); rend.reserve(reads); cigar.reserve(reads * 2); seq.reserve(reads * ((read_len / 2 + 3) / 4 + 1)); qual.reserve(reads * ((read_len + 7) / 8 + 1));
    }
    void clear() { qual_bits = 8; memset(qseen, 0, sizeof qseen); pos.clear(); flag.clear(); mapq.clear(); aux.clear(); l_qseq.clear(); cigar_off.clear(); seq_off.clear(); qual_off.clear(); frag_key.clear(); cigar.clear(); seq.clear(); qual.clear(); rend.clear(); }
    size_t bytes() const { return n() * (4 + 2 + 1 + 1 + 4 + 4 + 4 + 4 + 8) + 4 + cigar.size() * 4 + seq.size() * 4 + qual.size() * 8; }

    void add(const BamRec &r) {
        pos.push_back(r.pos); flag.push_back(r.flag); mapq.push_back(r.mapq);
        uint8_t a = 0;
        // getStrand (common.c:85-87): only the first value byte of XG matters
        if (const uint8_t *xg = aux_find(r.aux, r.l_aux, 'X', 'G')) { if (xg[1] == 'C') a |= 1; else if (xg[1] == 'G') a |= 2; }
        // filter_func (common.c:421-427): NH present and > 1
        if (const uint8_t *nh = aux_find(r.aux, r.l_aux, 'N', 'H')) { if (aux_to_int(nh) > 1) a |= 4; }
        aux.push_back(a);
        l_qseq.push_back((uint32_t) r.l_qseq);
        cigar_off.push_back((uint32_t) cigar.size());
        int rl = 0;
        uint32_t *c = cigar.grow(r.n_cigar);
        for (uint32_t k = 0; k < r.n_cigar; ++k) { c[k] = le32(r.cigar + 4 * k); uint32_t op = c[k] & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rl += (int)(c[k] >> 4); }
        rend.push_back(r.pos + rl);
        size_t sb = ((size_t) r.l_qseq + 1) / 2, sw = (sb + 3) / 4, qw = ((size_t) r.l_qseq + 7) / 8;
        seq_off.push_back((uint32_t) seq.size()); qual_off.push_back((uint32_t) qual.size());
        uint32_t *s = seq.grow(sw); if (sw) { s[sw - 1] = 0; memcpy(s, r.seq, sb); }
        uint64_t *q = qual.grow(qw); if (qw) { q[qw - 1] = 0; uint8_t *qb = (uint8_t *) q; for (int32_t j = 0; j < r.l_qseq; ++j) { const uint8_t v = r.qual[j]; qb[j] = v; qseen[v] = 1; } }
        frag_key.push_back(qname_key(r.qname, strnlen(r.qname, r.l_qname)));
    }
    // phred encoding of this tile: 8 = plain bytes; 2/4 = codes into `lut` (at most 4/16 distinct values, ascending)
    void set_encoding(uint32_t bits, const uint8_t *lut, int n_lut) {
        qual_bits = bits; memset(qual_lut, 0, 16); memset(qual_code, 0, 256);
        if (bits < 8) for (int k = 0; k < n_lut; ++k) { qual_lut[k] = lut[k]; qual_code[lut[k]] = (uint8_t) k; }
    }
    size_t qual_words_for(uint32_t l) const { return ((size_t) l * qual_bits + 63) / 64; }
    uint8_t qual_value(size_t i, uint32_t j) const {
        const uint64_t *w = qual.data() + qual_off[i];
        if (qual_bits == 8) return ((const uint8_t *) w)[j];
        if (qual_bits == 4) return qual_lut[(w[j >> 4] >> ((j & 15) * 4)) & 15];
        return qual_lut[(w[j >> 5] >> ((j & 31) * 2)) & 3];
    }
    // write l phred values (plain bytes) at w in this tile's encoding; w has qual_words_for(l) words
    void encode_quals(uint64_t *w, const uint8_t *q, uint32_t l) const {
        const size_t words = qual_words_for(l);
        if (!words) return;
        if (qual_bits == 8) { w[words - 1] = 0; memcpy(w, q, l); return; }
        if (qual_bits == 2) {
            uint32_t j = 0;
            for (size_t x = 0; x < words; ++x) { uint64_t v = 0; const uint32_t e = std::min<uint32_t>(l, j + 32); for (int sh = 0; j < e; ++j, sh += 2) v |= (uint64_t) qual_code[q[j]] << sh; w[x] = v; }
        } else {
            uint32_t j = 0;
            for (size_t x = 0; x < words; ++x) { uint64_t v = 0; const uint32_t e = std::min<uint32_t>(l, j + 16); for (int sh = 0; j < e; ++j, sh += 4) v |= (uint64_t) qual_code[q[j]] << sh; w[x] = v; }
        }
    }
    // copy read i of another tile (carry-over of reads that straddle a tile boundary); the phreds are re-encoded if the two
    // tiles differ in encoding (every value must exist in this tile's alphabet)
    void add_from(const SoaTile &o, size_t i) {
        pos.push_back(o.pos[i]); flag.push_back(o.flag[i]); mapq.push_back(o.mapq[i]); aux.push_back(o.aux[i]); l_qseq.push_back(o.l_qseq[i]);
        uint32_t c0 = o.cigar_off[i], c1 = (i + 1 < o.n()) ? o.cigar_off[i + 1] : (uint32_t) o.cigar.size();
        cigar_off.push_back((uint32_t) cigar.size());
        uint32_t *c = cigar.grow(c1 - c0); memcpy(c, o.cigar.data() + c0, (c1 - c0) * 4);
        const uint32_t l = o.l_qseq[i];
        size_t sw = (((size_t) l + 1) / 2 + 3) / 4, qw = qual_words_for(l);
        seq_off.push_back((uint32_t) seq.size()); qual_off.push_back((uint32_t) qual.size());
        uint32_t *s = seq.grow(sw); memcpy(s, o.seq.data() + o.seq_off[i], sw * 4);
        uint64_t *q = qual.grow(qw);
        if (o.qual_bits == 8) { const uint8_t *src = (const uint8_t *)(o.qual.data() + o.qual_off[i]); if (qual_bits == 8) for (uint32_t j = 0; j < l; ++j) qseen[src[j]] = 1; encode_quals(q, src, l); }
        else if (o.qual_bits == qual_bits && !memcmp(o.qual_lut, qual_lut, 16)) memcpy(q, o.qual.data() + o.qual_off[i], qw * 8);
        else { std::vector<uint8_t> tmp(l); for (uint32_t j = 0; j < l; ++j) { tmp[j] = o.qual_value(i, j); if (qual_bits == 8) qseen[tmp[j]] = 1; } encode_quals(q, tmp.data(), l); }
        frag_key.push_back(o.frag_key[i]); rend.push_back(o.rend[i]);
    }
    // bulk copy of reads [a,b) of another tile whose blobs are laid out in read order (as add() produces them)
    void add_range_from(const SoaTile &o, size_t a, size_t b) {
        if (b <= a) return;
        const size_t k = b - a, on = o.n();
        memcpy(pos.grow(k), o.pos.data() + a, k * 4); memcpy(flag.grow(k), o.flag.data() + a, k * 2);
        memcpy(mapq.grow(k), o.mapq.data() + a, k); memcpy(aux.grow(k), o.aux.data() + a, k);
        memcpy(l_qseq.grow(k), o.l_qseq.data() + a, k * 4); memcpy(frag_key.grow(k), o.frag_key.data() + a, k * 8);
        memcpy(rend.grow(k), o.rend.data() + a, k * 4);
        const uint32_t c0 = o.cigar_off[a], c1 = b < on ? o.cigar_off[b] : (uint32_t) o.cigar.size();
        const uint32_t s0 = o.seq_off[a], s1 = b < on ? o.seq_off[b] : (uint32_t) o.seq.size();
        const uint32_t q0 = o.qual_off[a], q1 = b < on ? o.qual_off[b] : (uint32_t) o.qual.size();
        const uint32_t cb = (uint32_t) cigar.size(), sb = (uint32_t) seq.size(), qb = (uint32_t) qual.size();
        uint32_t *co = cigar_off.grow(k), *so = seq_off.grow(k), *qo = qual_off.grow(k);
        for (size_t i = 0; i < k; ++i) { co[i] = cb + (o.cigar_off[a + i] - c0); so[i] = sb + (o.seq_off[a + i] - s0); qo[i] = qb + (o.qual_off[a + i] - q0); }
        memcpy(cigar.grow(c1 - c0), o.cigar.data() + c0, (size_t)(c1 - c0) * 4);
        memcpy(seq.grow(s1 - s0), o.seq.data() + s0, (size_t)(s1 - s0) * 4);
        memcpy(qual.grow(q1 - q0), o.qual.data() + q0, (size_t)(q1 - q0) * 8);
        for (int v = 0; v < 256; ++v) qseen[v] |= o.qseen[v];
    }
    // Planned assembly: the caller sums the sizes of the ranges it wants, grows every column once with extend(), and then the
    // ranges are copied into place independently (copy_range_at), e.g. on a thread pool.
    struct Extent { size_t n = 0, c = 0, s = 0, q = 0; };
    // sizes of reads [a,b) of an 8-bit tile `o` once stored in this tile's encoding
    Extent extent_of(const SoaTile &o, size_t a, size_t b) const {
        Extent e; const size_t on = o.n();
        e.n = b - a;
        e.c = (b < on ? o.cigar_off[b] : (uint32_t) o.cigar.size()) - o.cigar_off[a];
        e.s = (b < on ? o.seq_off[b] : (uint32_t) o.seq.size()) - o.seq_off[a];
        if (qual_bits == 8) e.q = (b < on ? o.qual_off[b] : (uint32_t) o.qual.size()) - o.qual_off[a];
        else for (size_t i = a; i < b; ++i) e.q += qual_words_for(o.l_qseq[i]);
        return e;
    }
    Extent extend(const Extent &by) {
        Extent at; at.n = n(); at.c = cigar.size(); at.s = seq.size(); at.q = qual.size();
        pos.grow(by.n); flag.grow(by.n); mapq.grow(by.n); aux.grow(by.n); l_qseq.grow(by.n); frag_key.grow(by.n); rend.grow(by.n);
        cigar_off.grow(by.n); seq_off.grow(by.n); qual_off.grow(by.n); cigar.grow(by.c); seq.grow(by.s); qual.grow(by.q);
        return at;
    }
    // `o` holds plain 8-bit phreds (a decoded fragment); they are written in this tile's encoding
    void copy_range_at(const SoaTile &o, size_t a, size_t b, const Extent &at) {
        const size_t k = b - a; if (!k) return;
        memcpy(pos.data() + at.n, o.pos.data() + a, k * 4); memcpy(flag.data() + at.n, o.flag.data() + a, k * 2);
        memcpy(mapq.data() + at.n, o.mapq.data() + a, k); memcpy(aux.data() + at.n, o.aux.data() + a, k);
        memcpy(l_qseq.data() + at.n, o.l_qseq.data() + a, k * 4); memcpy(frag_key.data() + at.n, o.frag_key.data() + a, k * 8);
        memcpy(rend.data() + at.n, o.rend.data() + a, k * 4);
        const size_t on = o.n();
        const uint32_t c0 = o.cigar_off[a], s0 = o.seq_off[a], q0 = o.qual_off[a];
        const uint32_t c1 = b < on ? o.cigar_off[b] : (uint32_t) o.cigar.size(), s1 = b < on ? o.seq_off[b] : (uint32_t) o.seq.size(), q1 = b < on ? o.qual_off[b] : (uint32_t) o.qual.size();
        uint32_t *co = cigar_off.data() + at.n, *so = seq_off.data() + at.n, *qo = qual_off.data() + at.n;
        for (size_t i = 0; i < k; ++i) { co[i] = (uint32_t) at.c + (o.cigar_off[a + i] - c0); so[i] = (uint32_t) at.s + (o.seq_off[a + i] - s0); }
        memcpy(cigar.data() + at.c, o.cigar.data() + c0, (size_t)(c1 - c0) * 4);
        memcpy(seq.data() + at.s, o.seq.data() + s0, (size_t)(s1 - s0) * 4);
        if (qual_bits == 8) {
            for (size_t i = 0; i < k; ++i) qo[i] = (uint32_t) at.q + (o.qual_off[a + i] - q0);
            memcpy(qual.data() + at.q, o.qual.data() + q0, (size_t)(q1 - q0) * 8);
        } else {
            size_t w = at.q;
            for (size_t i = 0; i < k; ++i) {
                const uint32_t l = o.l_qseq[a + i];
                qo[i] = (uint32_t) w;
                encode_quals(qual.data() + w, (const uint8_t *)(o.qual.data() + o.qual_off[a + i]), l);
                w += qual_words_for(l);
            }
        }
    }
    // Re-encode the phred column as 2- or 4-bit codes when the tile's alphabet allows it (lossless; see md_reads_soa).
    // `scratch` receives the packed words and is swapped in, so a ring of tiles re-uses its allocations.  `par(n, fn)` runs
    // fn(k) for k in [0,n) — the driver passes its decode pool, so a 2^17-alignment tile is packed in about a millisecond.
    template <class Par>
    void pack_quals(PodVec<uint64_t> &scratch, PodVec<uint32_t> &scratch_off, Par &&par) {
        if (qual_bits != 8 || n() == 0) return;
        const size_t N = n(), parts = std::min<size_t>(16, (N + 4095) / 4096);
        std::vector<std::array<bool, 256>> seen(parts);
        for (auto &a : seen) a.fill(false);
        par(parts, [&](size_t k) {
            const size_t i0 = N * k / parts, i1 = N * (k + 1) / parts;
            auto &sk = seen[k];
            for (size_t i = i0; i < i1; ++i) { const uint8_t *q = (const uint8_t *)(qual.data() + qual_off[i]); for (uint32_t j = 0, l = l_qseq[i]; j < l; ++j) sk[q[j]] = true; }
        });
        uint8_t code[256]; int na = 0;
        for (int v = 0; v < 256; ++v) { bool any = false; for (auto &a : seen) any = any || a[(size_t) v]; if (any) { if (na < 16) { qual_lut[na] = (uint8_t) v; code[v] = (uint8_t) na; } ++na; } }
        if (na > 16) return;
        const uint32_t bits = na <= 4 ? 2 : 4;
        for (int k = na; k < 16; ++k) qual_lut[k] = 0;
        // offsets first (serial, trivial), then the ranges pack independently
        scratch_off.clear(); scratch.clear();
        uint32_t *off = scratch_off.grow(N);
        size_t tot = 0;
        for (size_t i = 0; i < N; ++i) { off[i] = (uint32_t) tot; tot += ((size_t) l_qseq[i] * bits + 63) / 64; }
        uint64_t *dst = scratch.grow(tot);
        par(parts, [&](size_t k) {
            const size_t i0 = N * k / parts, i1 = N * (k + 1) / parts;
            for (size_t i = i0; i < i1; ++i) {
                const uint32_t l = l_qseq[i]; const size_t words = ((size_t) l * bits + 63) / 64;
                uint64_t *w = dst + off[i];
                const uint8_t *q = (const uint8_t *)(qual.data() + qual_off[i]);
                for (size_t x = 0; x < words; ++x) w[x] = 0;
                if (bits == 2) for (uint32_t j = 0; j < l; ++j) w[j >> 5] |= (uint64_t) code[q[j]] << ((j & 31) * 2);
                else for (uint32_t j = 0; j < l; ++j) w[j >> 4] |= (uint64_t) code[q[j]] << ((j & 15) * 4);
            }
        });
        qual.swap(scratch); qual_off.swap(scratch_off);
        qual_bits = bits; memset(qual_code, 0, 256); for (int k = 0; k < na; ++k) qual_code[qual_lut[k]] = (uint8_t) k;
    }
    void pack_quals(PodVec<uint64_t> &scratch, PodVec<uint32_t> &scratch_off) {
        pack_quals(scratch, scratch_off, [](size_t n, const std::function<void(size_t)> &fn) { for (size_t k = 0; k < n; ++k) fn(k); });
    }
    // Finalise (cigar_off gets its n+1'th entry) and expose as the C-ABI view.
    md_reads_soa view() {
        if (cigar_off.size() == n()) cigar_off.push_back((uint32_t) cigar.size());
        else cigar_off[n()] = (uint32_t) cigar.size();
        md_reads_soa v; memset(&v, 0, sizeof v);
        v.n_reads = (uint32_t) n(); v.n_cigar_ops = (uint32_t) cigar.size(); v.seq_words = seq.size(); v.qual_words = qual.size();
        v.pos = pos.data(); v.flag = flag.data(); v.mapq = mapq.data(); v.aux = aux.data(); v.l_qseq = l_qseq.data();
        v.cigar_off = cigar_off.data(); v.seq_off = seq_off.data(); v.qual_off = qual_off.data(); v.frag_key = frag_key.data();
        v.cigar = cigar.data(); v.seq = seq.data(); v.qual = qual.data();
        v.qual_bits = qual_bits; memcpy(v.qual_lut, qual_lut, 16);
        return v;
    }
};

// Sequential BAM record source with one-record look-ahead; records come back in file order.
class BamStream {
public:
    explicit BamStream(const std::string &path) : rd_(path) { hdr_ = read_bam_header(rd_); buf_.resize(1 << 16); }
    const BamHeader &header() const { return hdr_; }
    void seek(uint64_t voff) { rd_.seek(voff); have_ = false; eof_ = false; }
    // Look at the next record without consuming it; false at EOF. The view stays valid until pop().
    bool peek(BamRec &r) {
        if (!have_) { if (eof_ || !fetch()) { eof_ = true; return false; } have_ = true; }
        r = cur_;
        return true;
    }
    void pop() { have_ = false; }
private:
    bool fetch() {
        uint8_t b[4];
        size_t got = rd_.read(b, 4);
        if (got == 0) return false;
        if (got != 4) throw std::runtime_error("truncated BAM record");
        uint32_t bs = le32(b);
        if (bs > buf_.size()) buf_.resize(bs);
        if (rd_.read(buf_.data(), bs) != bs) throw std::runtime_error("truncated BAM record");
        if (!parse_bam_record(buf_.data(), bs, cur_)) throw std::runtime_error("malformed BAM record");
        return true;
    }
    BgzfReader rd_;
    BamHeader hdr_;
    std::vector<uint8_t> buf_;
    BamRec cur_;
    bool have_ = false, eof_ = false;
};

// Cuts the coordinate-sorted record stream of ONE contig interval into tiles.
//   - tile k owns [beg_k, end_k); beg_0 = region start, end_last = region end;
//   - every record overlapping the owned interval is in the tile, so records that straddle a cut
//     appear in both neighbours (the reference re-fetches them per chunk the same way, extract.c:379);
//   - a cut is only placed at the start coordinate of a record, once the tile holds >= target reads.
class Tiler {
public:
    Tiler(BamStream &bs, int tid, uint32_t reg_beg, uint32_t reg_end, size_t target_reads)
        : bs_(bs), tid_(tid), reg_beg_(reg_beg), reg_end_(reg_end), target_(target_reads), cur_beg_(reg_beg) {}
    // Fills `t` (cleared first) using `carry` (reads straddling the previous cut; updated for the next call).
    // Returns false when the interval is exhausted.  Records of earlier contigs are skipped; the first record
    // of a later contig / beyond the region stays un-consumed in the stream.
    bool next(SoaTile &t, SoaTile &carry) {
        if (done_) return false;
        t.clear(); t.tid = tid_; t.beg = cur_beg_;
        for (size_t i = 0; i < carry.n(); ++i) if (span_end(carry, i) > cur_beg_) t.add_from(carry, i);
        carry.clear();
        uint32_t cut = reg_end_;
        bool stream_end = false;
        BamRec r;
        for (;;) {
            if (!bs_.peek(r)) { stream_end = true; break; }
            if (r.tid != tid_) { if (r.tid > tid_ || r.tid < 0) { stream_end = true; break; } bs_.pop(); continue; }
            if ((uint32_t) r.pos >= reg_end_) { stream_end = true; break; }
            if (t.n() >= target_ && (uint32_t) r.pos > cur_beg_ && r.pos > last_pos_) { cut = (uint32_t) r.pos; break; }
            int rl = 0;
            for (uint32_t k = 0; k < r.n_cigar; ++k) { uint32_t c = le32(r.cigar + 4 * k), op = c & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rl += (int)(c >> 4); }
            if ((int64_t) r.pos + (rl ? rl : 1) > (int64_t) reg_beg_) { t.add(r); last_pos_ = r.pos; }   // index semantics: endpos > beg
            bs_.pop();
        }
        t.end = cut;
        // reads reaching beyond the cut (or beyond the region end) are handed on: to the next tile, or to the tiler of the adjacent region
        for (size_t i = 0; i < t.n(); ++i) if (span_end(t, i) > cut) carry.add_from(t, i);
        if (stream_end) done_ = true; else cur_beg_ = cut;
        return true;
    }
private:
    static uint32_t span_end(const SoaTile &t, size_t i) { return (uint32_t) std::max(t.rend[i], t.pos[i] + 1); }
    BamStream &bs_; int tid_; uint32_t reg_beg_, reg_end_; size_t target_;
    uint32_t cur_beg_; bool done_ = false; int32_t last_pos_ = -1;
};

}  // namespace mdhost
