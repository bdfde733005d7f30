// Host-side file formats for the B200 MethylDackel hot path: BGZF, BAM records,
// BAI, FASTA/.fai.  The reference does all of this through htslib
// (hts_open/sam_hdr_read/sam_itr_queryi/sam_itr_next/faidx_fetch_seq, call sites
// extract.c:283-295,379-381, common.c:413); htslib is not available here, so this is
// an independent implementation written against the SAM/BAM specification.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>
#include <algorithm>
#include <zlib.h>
#include <unistd.h>

namespace mdhost {

// ------------------------------------------------------------------ BGZF
static const int kBgzfBlock = 0xff00;  // uncompressed payload per block

class BgzfWriter {
public:
    BgzfWriter(const std::string &path, int level = 1) : level_(level) {
        fp_ = fopen(path.c_str(), "wb");
        if (!fp_) throw std::runtime_error("cannot open " + path + " for writing");
        buf_.reserve(kBgzfBlock);
    }
    ~BgzfWriter() { if (fp_) close(); }
    uint64_t tell() const { return (file_off_ << 16) | (uint64_t) buf_.size(); }
    void write(const void *p, size_t n) {
        const uint8_t *s = (const uint8_t *) p;
        while (n) {
            size_t take = std::min(n, (size_t) kBgzfBlock - buf_.size());
            buf_.insert(buf_.end(), s, s + take);
            s += take; n -= take;
            if (buf_.size() == (size_t) kBgzfBlock) flush_block();
        }
    }
    void flush_block() {
        if (buf_.empty()) return;
        emit(buf_.data(), buf_.size());
        buf_.clear();
    }
    void close() {
        flush_block();
        emit(nullptr, 0);  // EOF marker block
        fclose(fp_); fp_ = nullptr;
    }
    void close_no_eof() { flush_block(); fclose(fp_); fp_ = nullptr; }   // the caller appends more blocks (mdsynth)
private:
    void emit(const uint8_t *data, size_t n) {
        uint8_t out[65536 + 64];
        z_stream zs; memset(&zs, 0, sizeof zs);
        if (deflateInit2(&zs, level_, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw std::runtime_error("deflateInit2");
        zs.next_in = (Bytef *) data; zs.avail_in = (uInt) n;
        zs.next_out = out + 18; zs.avail_out = sizeof(out) - 18 - 8;
        if (deflate(&zs, Z_FINISH) != Z_STREAM_END) throw std::runtime_error("deflate: block did not fit");
        size_t clen = zs.total_out;
        deflateEnd(&zs);
        size_t bsize = 18 + clen + 8;
        static const uint8_t hdr[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
        memcpy(out, hdr, 16);
        out[16] = (uint8_t)((bsize - 1) & 0xff); out[17] = (uint8_t)((bsize - 1) >> 8);
        uint32_t crc = (uint32_t) crc32(crc32(0L, Z_NULL, 0), data, (uInt) n);
        uint8_t *t = out + 18 + clen;
        for (int i = 0; i < 4; ++i) t[i] = (uint8_t)(crc >> (8 * i));
        for (int i = 0; i < 4; ++i) t[4 + i] = (uint8_t)(((uint32_t) n) >> (8 * i));
        if (fwrite(out, 1, bsize, fp_) != bsize) throw std::runtime_error("short write");
        file_off_ += bsize;
    }
    FILE *fp_ = nullptr;
    int level_;
    std::vector<uint8_t> buf_;
    uint64_t file_off_ = 0;
};

// Sequential / seekable single-threaded BGZF reader.
class BgzfReader {
public:
    explicit BgzfReader(const std::string &path) {
        fp_ = fopen(path.c_str(), "rb");
        if (!fp_) throw std::runtime_error("Couldn't open " + path + " for reading!");
        setvbuf(fp_, nullptr, _IOFBF, 4 << 20);
        memset(&zs_, 0, sizeof zs_);
        if (inflateInit2(&zs_, -15) != Z_OK) throw std::runtime_error("inflateInit2");
    }
    ~BgzfReader() { inflateEnd(&zs_); if (fp_) fclose(fp_); }
    BgzfReader(const BgzfReader &) = delete;
    uint64_t tell() const { return ((uint64_t) block_addr_ << 16) | (uint64_t) uoff_; }
    void seek(uint64_t voff) {
        int64_t addr = (int64_t)(voff >> 16);
        eof_ = false;
        if (addr != block_addr_ || ulen_ == 0) { if (!load(addr)) { ulen_ = 0; uoff_ = 0; return; } }
        uoff_ = (int)(voff & 0xffff);
        if (uoff_ > ulen_) throw std::runtime_error("BGZF seek past block end");
    }
    // reads exactly n bytes unless EOF; returns bytes read
    size_t read(void *dst_, size_t n) {
        uint8_t *dst = (uint8_t *) dst_;
        size_t done = 0;
        while (done < n) {
            if (uoff_ >= ulen_) {
                if (eof_) break;
                if (!load(next_addr_)) break;
                if (ulen_ == 0) continue;
            }
            size_t take = std::min(n - done, (size_t)(ulen_ - uoff_));
            memcpy(dst + done, ubuf_ + uoff_, take);
            uoff_ += (int) take; done += take;
            if (uoff_ == ulen_) { block_addr_ = next_addr_; ulen_ = 0; uoff_ = 0; }
        }
        return done;
    }
private:
    bool load(int64_t addr) {
        uint8_t h[18];
        if (fseeko(fp_, addr, SEEK_SET) != 0) throw std::runtime_error("seek failed");
        size_t got = fread(h, 1, 12, fp_);
        if (got == 0) { eof_ = true; block_addr_ = addr; next_addr_ = addr; ulen_ = uoff_ = 0; return false; }
        if (got != 12 || h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) throw std::runtime_error("not a BGZF block");
        int xlen = h[10] | (h[11] << 8);
        std::vector<uint8_t> extra((size_t) xlen);
        if (fread(extra.data(), 1, (size_t) xlen, fp_) != (size_t) xlen) throw std::runtime_error("truncated BGZF header");
        int bsize = -1;
        for (int off = 0; off + 4 <= xlen;) {
            int slen = extra[off + 2] | (extra[off + 3] << 8);
            if (extra[off] == 'B' && extra[off + 1] == 'C' && slen == 2) bsize = extra[off + 4] | (extra[off + 5] << 8);
            off += 4 + slen;
        }
        if (bsize < 0) throw std::runtime_error("BGZF block without BC field");
        int clen = bsize + 1 - 12 - xlen;
        if (clen < 8 || clen > 65536) throw std::runtime_error("bad BGZF block size");
        if (fread(cbuf_, 1, (size_t) clen, fp_) != (size_t) clen) throw std::runtime_error("truncated BGZF block");
        uint32_t isize = (uint32_t) cbuf_[clen - 4] | ((uint32_t) cbuf_[clen - 3] << 8) | ((uint32_t) cbuf_[clen - 2] << 16) | ((uint32_t) cbuf_[clen - 1] << 24);
        if (isize > 65536) throw std::runtime_error("bad BGZF isize");
        inflateReset(&zs_);
        zs_.next_in = cbuf_; zs_.avail_in = (uInt)(clen - 8);
        zs_.next_out = ubuf_; zs_.avail_out = 65536;
        if (inflate(&zs_, Z_FINISH) != Z_STREAM_END || zs_.total_out != isize) throw std::runtime_error("BGZF inflate failed");
        ulen_ = (int) isize; uoff_ = 0; block_addr_ = addr; next_addr_ = addr + bsize + 1;
        return true;
    }
    FILE *fp_ = nullptr;
    z_stream zs_;
    uint8_t cbuf_[65536], ubuf_[65536];
    int ulen_ = 0, uoff_ = 0;
    int64_t block_addr_ = 0, next_addr_ = 0;
    bool eof_ = false;
};

// ------------------------------------------------------------------ BAM
struct BamHeader {
    std::string text;
    std::vector<std::string> names;
    std::vector<uint32_t> lens;
    int name2tid(const std::string &n) const {
        for (size_t i = 0; i < names.size(); ++i) if (names[i] == n) return (int) i;
        return -1;
    }
};

inline uint32_t le32(const uint8_t *p) { return (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24); }
inline void put32(std::vector<uint8_t> &v, uint32_t x) { for (int i = 0; i < 4; ++i) v.push_back((uint8_t)(x >> (8 * i))); }

inline BamHeader read_bam_header(BgzfReader &r) {
    BamHeader h;
    uint8_t b[8];
    r.seek(0);
    if (r.read(b, 8) != 8 || memcmp(b, "BAM\1", 4) != 0) throw std::runtime_error("not a BAM file");
    uint32_t l_text = le32(b + 4);
    h.text.resize(l_text);
    if (r.read(&h.text[0], l_text) != l_text) throw std::runtime_error("truncated BAM header");
    if (r.read(b, 4) != 4) throw std::runtime_error("truncated BAM header");
    uint32_t n_ref = le32(b);
    for (uint32_t i = 0; i < n_ref; ++i) {
        if (r.read(b, 4) != 4) throw std::runtime_error("truncated BAM header");
        uint32_t l_name = le32(b);
        std::string nm(l_name, 0);
        if (r.read(&nm[0], l_name) != l_name || r.read(b, 4) != 4) throw std::runtime_error("truncated BAM header");
        if (!nm.empty() && nm.back() == 0) nm.pop_back();
        h.names.push_back(nm);
        h.lens.push_back(le32(b));
    }
    return h;
}

inline void write_bam_header(BgzfWriter &w, const BamHeader &h) {
    std::vector<uint8_t> v;
    v.insert(v.end(), {'B', 'A', 'M', 1});
    put32(v, (uint32_t) h.text.size());
    v.insert(v.end(), h.text.begin(), h.text.end());
    put32(v, (uint32_t) h.names.size());
    for (size_t i = 0; i < h.names.size(); ++i) {
        put32(v, (uint32_t) h.names[i].size() + 1);
        v.insert(v.end(), h.names[i].begin(), h.names[i].end());
        v.push_back(0);
        put32(v, h.lens[i]);
    }
    w.write(v.data(), v.size());
}

// BAI binning (SAM spec section 5.3)
inline int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

inline int cigar_ref_len(const uint32_t *cig, uint32_t n) {
    int l = 0;
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t op = cig[k] & 15;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) l += (int)(cig[k] >> 4);
    }
    return l;
}

// A decoded view of one BAM record (pointers into the caller's buffer).
struct BamRec {
    int32_t tid, pos, mtid, mpos, isize;
    uint16_t flag, n_cigar, bin;
    uint8_t mapq, l_qname;
    int32_t l_qseq;
    const char *qname;
    const uint8_t *cigar;   // unaligned little-endian u32[n_cigar]
    const uint8_t *seq, *qual, *aux;
    size_t l_aux;
};

inline bool parse_bam_record(const uint8_t *p, size_t block_size, BamRec &r) {
    if (block_size < 32) return false;
    r.tid = (int32_t) le32(p); r.pos = (int32_t) le32(p + 4);
    uint32_t bmq = le32(p + 8), fnc = le32(p + 12);
    r.bin = (uint16_t)(bmq >> 16); r.mapq = (uint8_t)((bmq >> 8) & 0xff); r.l_qname = (uint8_t)(bmq & 0xff);
    r.flag = (uint16_t)(fnc >> 16); r.n_cigar = (uint16_t)(fnc & 0xffff);
    r.l_qseq = (int32_t) le32(p + 16); r.mtid = (int32_t) le32(p + 20); r.mpos = (int32_t) le32(p + 24); r.isize = (int32_t) le32(p + 28);
    size_t need = 32 + (size_t) r.l_qname + 4 * (size_t) r.n_cigar + ((size_t) r.l_qseq + 1) / 2 + (size_t) r.l_qseq;
    if (r.l_qseq < 0 || need > block_size) return false;
    r.qname = (const char *)(p + 32);
    r.cigar = p + 32 + r.l_qname;
    r.seq = r.cigar + 4 * (size_t) r.n_cigar;
    r.qual = r.seq + ((size_t) r.l_qseq + 1) / 2;
    r.aux = r.qual + r.l_qseq;
    r.l_aux = block_size - need;
    return true;
}

// Scan aux fields; returns pointer to the type byte of `tag` or nullptr (bam_aux_get contract).
inline const uint8_t *aux_find(const uint8_t *s, size_t l, char t0, char t1) {
    const uint8_t *end = s + l;
    auto tsize = [](int t) { switch (t) { case 'A': case 'c': case 'C': return 1; case 's': case 'S': return 2; case 'i': case 'I': case 'f': return 4; case 'd': return 8; default: return 0; } };
    while (s + 3 <= end) {
        bool match = (s[0] == (uint8_t) t0 && s[1] == (uint8_t) t1);
        const uint8_t *tp = s + 2; int t = *tp; const uint8_t *v = tp + 1;
        if (t == 'Z' || t == 'H') { const uint8_t *q = v; while (q < end && *q) ++q; if (q >= end) return nullptr; if (match) return tp; s = q + 1; }
        else if (t == 'B') { if (v + 5 > end) return nullptr; int sz = tsize(v[0]); if (!sz) return nullptr; uint32_t n = le32(v + 1); if (match) return tp; s = v + 5 + (size_t) sz * n; }
        else { int sz = tsize(t); if (!sz || v + sz > end) return nullptr; if (match) return tp; s = v + sz; }
    }
    return nullptr;
}

inline int64_t aux_to_int(const uint8_t *tp) {
    const uint8_t *s = tp + 1;
    switch (*tp) {
        case 'c': return (int8_t) s[0]; case 'C': return s[0];
        case 's': return (int16_t)(s[0] | (s[1] << 8)); case 'S': return (uint16_t)(s[0] | (s[1] << 8));
        case 'i': return (int32_t) le32(s); case 'I': return (uint32_t) le32(s);
        default: return 0;
    }
}

// BAI writer fed with (tid, beg, end, voffset range) per record, in file order.
class BaiBuilder {
public:
    explicit BaiBuilder(size_t n_ref) : refs_(n_ref) {}
    void add(int tid, int64_t beg, int64_t end, uint64_t v0, uint64_t v1) {
        if (tid < 0) return;
        if (end <= beg) end = beg + 1;
        Ref &R = refs_[(size_t) tid];
        int bin = reg2bin(beg, end);
        auto &ch = R.bins[(uint32_t) bin];
        if (!ch.empty() && ch.back().second == v0) ch.back().second = v1; else ch.emplace_back(v0, v1);
        size_t w0 = (size_t)(beg >> 14), w1 = (size_t)((end - 1) >> 14);
        if (R.ioff.size() <= w1) R.ioff.resize(w1 + 1, 0);
        for (size_t w = w0; w <= w1; ++w) if (R.ioff[w] == 0) R.ioff[w] = v0;
    }
    void write(const std::string &path) {
        FILE *f = fopen(path.c_str(), "wb");
        if (!f) throw std::runtime_error("cannot write " + path);
        auto w32 = [&](uint32_t x) { uint8_t b[4]; for (int i = 0; i < 4; ++i) b[i] = (uint8_t)(x >> (8 * i)); fwrite(b, 1, 4, f); };
        auto w64 = [&](uint64_t x) { w32((uint32_t) x); w32((uint32_t)(x >> 32)); };
        fwrite("BAI\1", 1, 4, f);
        w32((uint32_t) refs_.size());
        for (Ref &R : refs_) {
            w32((uint32_t) R.bins.v.size());
            for (auto &kv : R.bins.v) { w32(kv.first); w32((uint32_t) kv.second.size()); for (auto &c : kv.second) { w64(c.first); w64(c.second); } }
            // empty 16 kb windows take the next window's offset (a valid lower bound)
            for (size_t k = R.ioff.size(); k-- > 1;) if (R.ioff[k - 1] == 0) R.ioff[k - 1] = R.ioff[k];
            w32((uint32_t) R.ioff.size());
            for (uint64_t o : R.ioff) w64(o);
        }
        fclose(f);
    }
private:
    typedef std::vector<std::pair<uint64_t, uint64_t>> Chunks;
    struct BinMap {
        std::vector<std::pair<uint32_t, Chunks>> v;
        Chunks &operator[](uint32_t bin) {
            // bins are visited with strong locality (sorted input), so look at the most recent first
            for (size_t k = v.size(); k-- > 0;) if (v[k].first == bin) return v[k].second;
            v.emplace_back(bin, Chunks());
            return v.back().second;
        }
    };
    struct Ref { BinMap bins; std::vector<uint64_t> ioff; };
    std::vector<Ref> refs_;
};

// BAI reader: only what a region seek needs (bins + linear index).
struct BaiIndex {
    struct Ref { std::vector<std::pair<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>>> bins; std::vector<uint64_t> ioff; };
    std::vector<Ref> refs;
    static bool load(const std::string &path, BaiIndex &idx) {
        FILE *f = fopen(path.c_str(), "rb");
        if (!f) return false;
        auto r32 = [&](uint32_t &x) { uint8_t b[4]; if (fread(b, 1, 4, f) != 4) return false; x = le32(b); return true; };
        auto r64 = [&](uint64_t &x) { uint32_t lo, hi; if (!r32(lo) || !r32(hi)) return false; x = ((uint64_t) hi << 32) | lo; return true; };
        char m[4]; uint32_t n_ref;
        bool ok = fread(m, 1, 4, f) == 4 && memcmp(m, "BAI\1", 4) == 0 && r32(n_ref);
        if (ok) {
            idx.refs.resize(n_ref);
            for (uint32_t r = 0; ok && r < n_ref; ++r) {
                uint32_t n_bin, n_intv;
                ok = r32(n_bin);
                for (uint32_t k = 0; ok && k < n_bin; ++k) {
                    uint32_t bin, n_chunk; ok = r32(bin) && r32(n_chunk);
                    std::vector<std::pair<uint64_t, uint64_t>> ch(ok ? n_chunk : 0);
                    for (uint32_t c = 0; ok && c < n_chunk; ++c) ok = r64(ch[c].first) && r64(ch[c].second);
                    idx.refs[r].bins.emplace_back(bin, std::move(ch));
                }
                ok = ok && r32(n_intv);
                if (ok) { idx.refs[r].ioff.resize(n_intv); for (uint32_t k = 0; ok && k < n_intv; ++k) ok = r64(idx.refs[r].ioff[k]); }
            }
        }
        fclose(f);
        return ok;
    }
    // Compressed file position (bytes) near which the alignments at `pos` of contig `tid` are stored, from the linear index
    // (one entry per 16 kb); used to weigh genome intervals by the amount of alignment data they hold.  0 when unknown.
    uint64_t file_pos(int tid, int64_t pos) const {
        if (tid < 0 || (size_t) tid >= refs.size()) return 0;
        const Ref &R = refs[(size_t) tid];
        if (R.ioff.empty()) return 0;
        const size_t w = (size_t)(pos >> 14);
        return (w < R.ioff.size() ? R.ioff[w] : R.ioff.back()) >> 16;
    }
    // Smallest virtual offset at which an alignment overlapping [beg, ...) on tid can start; 0 if unknown.
    // found=false when the index proves there is no alignment on tid at all.
    uint64_t start_offset(int tid, int64_t beg, bool &found) const {
        found = false;
        if (tid < 0 || (size_t) tid >= refs.size()) return 0;
        const Ref &R = refs[(size_t) tid];
        uint64_t min_chunk = UINT64_MAX;
        for (auto &kv : R.bins) if (kv.first < 37449) for (auto &c : kv.second) min_chunk = std::min(min_chunk, c.first);
        if (min_chunk == UINT64_MAX) return 0;
        found = true;
        uint64_t lin = 0;
        if (!R.ioff.empty()) { size_t w = (size_t)(beg >> 14); lin = w < R.ioff.size() ? R.ioff[w] : R.ioff.back(); }
        return std::max(lin, min_chunk);
    }
};

// ------------------------------------------------------------------ FASTA (+ .fai)
struct FaiEntry { std::string name; int64_t len = 0, offset = 0; int line_blen = 0, line_len = 0; };

class Fasta {
public:
    // Loads fn.fai, or scans the FASTA to build the index in memory (and writes fn.fai when the
    // directory is writable, as fai_load does — extract.c:283).
    explicit Fasta(const std::string &fn) : fn_(fn) {
        FILE *fi = fopen((fn + ".fai").c_str(), "r");
        if (fi) {
            char line[8192];
            while (fgets(line, sizeof line, fi)) {
                char nm[4096]; long long len, off; int lb, ll;
                if (sscanf(line, "%4095[^\t]\t%lld\t%lld\t%d\t%d", nm, &len, &off, &lb, &ll) == 5) {
                    FaiEntry e; e.name = nm; e.len = len; e.offset = off; e.line_blen = lb; e.line_len = ll; entries_.push_back(e);
                }
            }
            fclose(fi);
        } else {
            build();
            FILE *fo = fopen((fn + ".fai").c_str(), "w");
            if (fo) { for (auto &e : entries_) fprintf(fo, "%s\t%lld\t%lld\t%d\t%d\n", e.name.c_str(), (long long) e.len, (long long) e.offset, e.line_blen, e.line_len); fclose(fo); }
        }
        fp_ = fopen(fn.c_str(), "rb");
        if (!fp_) throw std::runtime_error("Couldn't open the index for " + fn + "!");
    }
    ~Fasta() { if (fp_) fclose(fp_); }
    Fasta(const Fasta &) = delete;
    const std::vector<FaiEntry> &entries() const { return entries_; }
    const FaiEntry *find(const std::string &name) const { for (auto &e : entries_) if (e.name == name) return &e; return nullptr; }
    // whole contig, case preserved, newlines stripped
    bool fetch(const std::string &name, std::string &out) const {
        const FaiEntry *e = find(name);
        if (!e) return false;
        return fetch_range(*e, 0, e->len, out) && (int64_t) out.size() == e->len;
    }
    // bases [beg, beg+n) of a contig (clamped to its end), as faidx_fetch_seq returns them.  Reads through its own descriptor
    // position (pread), so concurrent calls are fine.
    bool fetch_range(const FaiEntry &e, int64_t beg, int64_t n, std::string &out) const {
        out.clear();
        if (beg < 0) beg = 0;
        if (beg >= e.len || n <= 0) return true;
        if (beg + n > e.len) n = e.len - beg;
        const int64_t lb = e.line_blen > 0 ? e.line_blen : e.len, ll = e.line_len >= lb ? e.line_len : lb;
        const int64_t first_line = beg / lb, last_line = (beg + n - 1) / lb;
        const int64_t off0 = e.offset + first_line * ll + beg % lb;
        const int64_t off1 = e.offset + last_line * ll + (beg + n - 1) % lb + 1;
        std::vector<char> raw((size_t)(off1 - off0));
        size_t got = 0;
        while (got < raw.size()) { ssize_t r = pread(fileno(fp_), raw.data() + got, raw.size() - got, (off_t)(off0 + (int64_t) got)); if (r <= 0) break; got += (size_t) r; }
        out.resize((size_t) n);
        // fast path: uniform lines — copy line by line, then make sure nothing but sequence characters came along
        bool clean = got == raw.size();
        if (clean) {
            int64_t src = 0, dst = 0, col = beg % lb;
            while (dst < n) { const int64_t take = std::min<int64_t>(lb - col, n - dst); memcpy(&out[(size_t) dst], raw.data() + src, (size_t) take); dst += take; src += take + (ll - lb); col = 0; }
            unsigned char bad = 0;
            for (size_t i = 0; i < out.size(); ++i) { const unsigned char c = (unsigned char) out[i]; bad |= (unsigned char)((c <= 32) | (c >= 127)); }
            clean = !bad;
        }
        if (!clean) {          // irregular file: drop everything that is not a printable character, as the reference's fetch does
            out.clear();
            for (size_t i = 0; i < got && (int64_t) out.size() < n; ++i) { unsigned char c = (unsigned char) raw[i]; if (c > 32 && c < 127) out.push_back((char) c); }
        }
        return true;
    }
private:
    void build() {
        FILE *f = fopen(fn_.c_str(), "rb");
        if (!f) throw std::runtime_error("Couldn't open the index for " + fn_ + "!");
        std::vector<char> buf(1 << 20);
        int64_t off = 0; FaiEntry *cur = nullptr;
        bool in_name = false, name_done = false, first = true; int64_t lb = 0, ll = 0; std::string nm;
        size_t n;
        while ((n = fread(buf.data(), 1, buf.size(), f)) > 0) {
            for (size_t i = 0; i < n; ++i) {
                int c = (unsigned char) buf[i]; ++off;
                if (in_name) {
                    if (c == '\n') { in_name = false; entries_.emplace_back(); cur = &entries_.back(); cur->name = nm; cur->offset = off; first = true; lb = ll = 0; }
                    else if (!name_done) { if (c == ' ' || c == '\t' || c == '\r') name_done = true; else nm.push_back((char) c); }
                    continue;
                }
                if (c == '>' && ll == 0) { in_name = true; name_done = false; nm.clear(); continue; }
                if (!cur) continue;
                ++ll;
                if (c == '\n') { if (first && lb > 0) { cur->line_blen = (int) lb; cur->line_len = (int) ll; first = false; } lb = ll = 0; }
                else if (c != '\r') { ++lb; ++cur->len; }
            }
        }
        if (cur && first && lb > 0) { cur->line_blen = (int) lb; cur->line_len = (int) lb + 1; }
        fclose(f);
    }
    std::string fn_;
    FILE *fp_ = nullptr;
    std::vector<FaiEntry> entries_;
};

// 64-bit FNV-1a style fingerprint of a query name, finished with a splitmix avalanche.
// This is the pairing key that replaces the khash string key of overlaps.c:125.
inline uint64_t qname_key(const char *s, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull ^ (n * 0x9e3779b97f4a7c15ull);
    for (size_t i = 0; i < n; ++i) { h ^= (uint8_t) s[i]; h *= 0x100000001b3ull; }
    h ^= h >> 30; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 27; h *= 0x94d049bb133111ebull; h ^= h >> 31;
    if (h == 0) h = 1;                 // 0 is reserved as "empty slot" on the device
    return h;
}

}  // namespace mdhost
