// mdsynth — deterministic synthetic WGBS data set generator (FASTA + .fai, coordinate-sorted
// BAM + .bai) following the recipe in SURVEY.md section 8d.  Used by the tests, by bench.py
// and by the CPU reference arm (which needs real files).  Not part of the extract/mbias path.
//
//   mdsynth --out PREFIX [--contigs name:len,...] [--depth 30] [--readlen 150]
//           [--isize-mean 300 --isize-sd 60 --isize-min 150 --isize-max 800]
//           [--genome-seed 1234] [--read-seed 5678] [--level 1] [--threads N]
//           [--bismark-tags] [--nondirectional F] [--single-frac F] [--lower-frac F] [--n-frac F]
//           [--quals N]  (N distinct phred values instead of the 4 binned ones)
//           [--clean]   (no flag noise / indels / clips: every record is a plain properly paired 150M)
//           [--human N] (contigs chr1..chr22,chrX,chrY with the proportions of GRCh38, N bases in total; replaces --contigs)
//
// The genome is cut into slabs (about 50 k fragments each).  Every slab draws its fragments from its own random stream
// (seeded by read seed, contig and slab number), so slabs are generated, and their BGZF blocks compressed, by a pool of
// threads while the file that comes out does not depend on the number of threads.  A 3 Gbp x 30x set (600 M records)
// takes minutes instead of hours.
#include "hostio.hpp"
#include <cmath>
#include <cstdlib>
#include <memory>
#include <thread>
#include <atomic>

using namespace mdhost;

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() { uint64_t z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t) n) >> 32); }
    double expo() { double u; do u = uni(); while (u <= 0.0); return -std::log(u); }
    double gamma_int(int k) { double x = 0; for (int i = 0; i < k; ++i) x += expo(); return x; }
    double beta_int(int a, int b) { double x = gamma_int(a), y = gamma_int(b); return x / (x + y); }
    double normal() { double u1; do u1 = uni(); while (u1 <= 0.0); double u2 = uni(); return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2); }
};
static uint64_t mix(uint64_t a, uint64_t b, uint64_t c) { Rng r(a ^ (b * 0xd6e8feb86659fd93ull) ^ (c * 0xca5a826395121157ull)); r.next(); return r.next(); }

struct Opts {
    std::string out;
    std::vector<std::pair<std::string, uint32_t>> contigs{{"chr1", 100000}};
    double depth = 30; int readlen = 150;
    double isize_mean = 300, isize_sd = 60; int isize_min = 150, isize_max = 800;
    uint64_t genome_seed = 1234, read_seed = 5678;
    int level = 1, threads = 0;
    bool bismark_tags = false, clean = false;
    double nondirectional = 0.0, single_frac = 0.0, lower_frac = 0.0, n_frac = 0.0;
    int quals = 4;          // distinct phred values: 4 = binned Illumina-like; more = uniform over [2, 2+quals)
};

struct RecMeta { int32_t pos, end; uint32_t ord; uint32_t off, len; };
struct Slab { std::vector<uint8_t> bytes; std::vector<RecMeta> recs; uint64_t n_frag = 0; };

static uint8_t nib(char c) { switch (c) { case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': return 8; default: return 15; } }
static void put32v(std::vector<uint8_t> &v, uint32_t x) { size_t n = v.size(); v.resize(n + 4); for (int i = 0; i < 4; ++i) v[n + (size_t) i] = (uint8_t)(x >> (8 * i)); }

// methylation probability (x255) of reference position i: CpG from the bimodal mixture (shared by the C and the G of the site), else 1 %
static uint8_t beta_at(const std::string &ref, int64_t i, uint64_t seed, size_t tid) {
    const int64_t G = (int64_t) ref.size();
    const char c = (char) toupper(ref[(size_t) i]);
    int64_t site = -1;
    if (c == 'C' && i + 1 < G && toupper(ref[(size_t) i + 1]) == 'G') site = i;
    else if (c == 'G' && i > 0 && toupper(ref[(size_t) i - 1]) == 'C') site = i - 1;
    if (site < 0) return 3;                                      // lround(0.01 * 255)
    Rng m(mix(seed, 0x5151ull + tid, (uint64_t) site));
    const double b = (m.uni() < 0.8) ? m.beta_int(8, 2) : m.beta_int(1, 8);
    return (uint8_t) std::lround(b * 255.0);
}

// all fragments that START in [beg, end) of contig tid
static void gen_slab(const Opts &o, size_t tid, const std::string &ref, uint64_t slab_global, int64_t beg, int64_t end, Slab &S) {
    const int L = o.readlen;
    const int64_t G = (int64_t) ref.size();
    const double n_frag = o.depth * (double) G / (2.0 * L);
    const double gap = (double) G / n_frag;
    Rng r(mix(o.read_seed, tid + 1, slab_global + 1));
    // phred alphabet and the chance of a wrong base for each value, as thresholds on a 32-bit draw
    uint32_t err_thr[256];
    for (int q = 0; q < 256; ++q) err_thr[q] = (uint32_t) std::min(4294967295.0, std::pow(10.0, -q / 10.0) * 4294967296.0);
    std::vector<uint8_t> beta_cache((size_t)(end - beg) + (size_t) o.isize_max + 64, 0), beta_have(beta_cache.size(), 0);
    auto beta = [&](int64_t p) -> uint8_t {
        const size_t k = (size_t)(p - beg);
        if (k < beta_cache.size()) { if (!beta_have[k]) { beta_cache[k] = beta_at(ref, p, o.genome_seed, tid); beta_have[k] = 1; } return beta_cache[k]; }
        return beta_at(ref, p, o.genome_seed, tid);
    };
    S.bytes.clear(); S.recs.clear(); S.n_frag = 0;
    S.bytes.reserve((size_t)((double)(end - beg) / gap * 2.0 * (L * 1.5 + 60)) + 4096);
    uint32_t ord = 0;
    std::vector<std::pair<int, int>> cig;
    std::vector<uint8_t> bases((size_t) L), quals((size_t) L);
    double cursor = (double) beg + r.expo() * gap;
    while (cursor < (double) end && cursor < (double)(G - 1)) {
        const int64_t s = (int64_t) cursor;
        cursor += r.expo() * gap;
        int isize = (int) std::lround(o.isize_mean + o.isize_sd * r.normal());
        isize = std::max(o.isize_min, std::min(o.isize_max, isize));
        if (isize < L) isize = L;
        if (s + isize + 8 > G) continue;
        const uint64_t frag_id = (slab_global << 21) + (++S.n_frag);
        char qname[32]; const int lq = snprintf(qname, sizeof qname, "f%011llu", (unsigned long long) frag_id) + 1;
        // library strand: OT / OB, optionally CTOT / CTOB (needs the XG tag to be told apart)
        const bool ob = r.uni() < 0.5;
        const bool compl_strand = o.nondirectional > 0 && r.uni() < o.nondirectional;
        const bool single = o.single_frac > 0 && r.uni() < o.single_frac;
        // record-level noise shared by the pair
        uint8_t mapq; { double u = r.uni(); mapq = u < 0.88 ? 60 : u < 0.93 ? 30 : u < 0.97 ? 9 : 0; }
        const bool dup = !o.clean && r.uni() < 0.02, qcfail = !o.clean && r.uni() < 0.005, secondary = !o.clean && r.uni() < 0.005;
        const bool singleton = !o.clean && !single && r.uni() < 0.01, improper = !o.clean && !single && r.uni() < 0.02, nh2 = !o.clean && r.uni() < 0.005;
        if (o.clean) mapq = 60;
        const int64_t left_pos = s, right_pos = s + isize - L;
        for (int mate = 0; mate < 2; ++mate) {   // mate 0 = leftmost record, 1 = rightmost
            if (single && mate == 1) break;
            if (singleton && mate == 1) break;
            const bool is_left = mate == 0;
            // OT: read1 forward on the left, read2 reverse on the right.  OB: read2 forward on the left, read1 reverse on the right.
            // CTOT (XG=CT): read1 reverse (right), read2 forward (left).  CTOB (XG=GA): read1 forward (left), read2 reverse (right).
            const bool conv_ct = !ob;                           // which conversion the bases carry, in reference orientation
            const bool read1_left = compl_strand ? ob : !ob;    // OT,CTOB: read1 is the left (forward) record
            const bool is_read1 = (is_left == read1_left);
            uint16_t flag;
            if (single) flag = (uint16_t)(ob ? 16 : 0);
            else {
                flag = 1;
                if (!improper && !singleton) flag |= 2;
                if (singleton) flag |= 8;
                flag |= is_left ? 0x20 : 0x10;            // left record is forward (mate reverse), right record is reverse
                flag |= is_read1 ? 0x40 : 0x80;
            }
            if (dup) flag |= 0x400;
            if (qcfail) flag |= 0x200;
            if (secondary) flag |= 0x100;
            const int64_t pos = is_left ? left_pos : right_pos;
            // CIGAR shape
            cig.clear();   // (op, len) ; ops: 0 M, 1 I, 2 D, 4 S
            {
                const double u = o.clean ? 0.0 : r.uni();
                int clip5 = 0, clip3 = 0, indel_at = -1, indel_len = 0; bool ins = false;
                if (u >= 0.94 && u < 0.97) { if (r.uni() < 0.5) clip5 = 1 + (int) r.below(20); else clip3 = 1 + (int) r.below(20); }
                else if (u >= 0.97 && u < 0.99) { indel_len = 1 + (int) r.below(3); ins = r.uni() < 0.5; }
                else if (u >= 0.99) { if (r.uni() < 0.5) clip5 = 1 + (int) r.below(20); else clip3 = 1 + (int) r.below(20); indel_len = 1 + (int) r.below(3); ins = r.uni() < 0.5; }
                const int body = L - clip5 - clip3;
                if (body < 45) indel_len = 0;                       // short reads: no room for an indel away from the ends
                if (indel_len) indel_at = 10 + (int) r.below((uint32_t)(body - 30));
                if (clip5) cig.emplace_back(4, clip5);
                if (indel_len) {
                    cig.emplace_back(0, indel_at);
                    if (ins) { cig.emplace_back(1, indel_len); cig.emplace_back(0, body - indel_at - indel_len); }
                    else { cig.emplace_back(2, indel_len); cig.emplace_back(0, body - indel_at); }
                } else cig.emplace_back(0, body);
                if (clip3) cig.emplace_back(4, clip3);
            }
            // bases + quals, walking the CIGAR over the reference
            int q = 0; int64_t p = pos;
            for (auto &oplen : cig) {
                const int op = oplen.first, len = oplen.second;
                if (op == 0) {
                    if (p + len > G) { p += len; q += len; continue; }
                    for (int j = 0; j < len; ++j, ++p) {
                        char c = (char) toupper(ref[(size_t) p]);
                        if (conv_ct && c == 'C') { if ((r.next() >> 56) >= beta(p)) c = 'T'; }
                        else if (!conv_ct && c == 'G') { if ((r.next() >> 56) >= beta(p)) c = 'A'; }
                        bases[(size_t) q++] = (uint8_t) c;
                    }
                } else if (op == 1 || op == 4) { for (int j = 0; j < len; ++j) bases[(size_t) q++] = (uint8_t) "ACGT"[r.next() >> 62]; }
                else if (op == 2) p += len;
            }
            const int32_t rend = (int32_t) p;
            if (rend > G) continue;
            for (int j = 0; j < L; ++j) {
                const uint64_t z = r.next();
                const uint32_t uq = (uint32_t)(z >> 48);                       // 16 bits pick the phred bin: 75 % / 15 % / 8 % / 2 %
                uint8_t ql = uq < 49152u ? 37 : uq < 58982u ? 25 : uq < 64225u ? 11 : 2;
                if (o.quals != 4) ql = (uint8_t)(2 + (uint32_t)((((z >> 48) & 0xffffu) * (uint64_t) o.quals) >> 16));
                quals[(size_t) j] = ql;
                if ((uint32_t) z < err_thr[ql]) { char c; do c = "ACGT"[r.next() >> 62]; while (c == (char) bases[(size_t) j]); bases[(size_t) j] = (uint8_t) c; }
            }
            // encode
            std::vector<uint8_t> &d = S.bytes;
            const size_t start = d.size();
            const int64_t mpos = single ? -1 : (is_left ? right_pos : left_pos);
            const int32_t tlen = single ? 0 : (is_left ? isize : -isize);
            put32v(d, 0);  // block_size placeholder
            put32v(d, (uint32_t) tid); put32v(d, (uint32_t) pos);
            put32v(d, ((uint32_t) reg2bin(pos, rend > pos ? rend : pos + 1) << 16) | ((uint32_t) mapq << 8) | (uint32_t) lq);
            put32v(d, ((uint32_t) flag << 16) | (uint32_t) cig.size());
            put32v(d, (uint32_t) L);
            put32v(d, single ? 0xffffffffu : (uint32_t) tid); put32v(d, (uint32_t) mpos); put32v(d, (uint32_t) tlen);
            d.insert(d.end(), qname, qname + lq);
            for (auto &oplen : cig) put32v(d, ((uint32_t) oplen.second << 4) | (uint32_t) oplen.first);
            for (int j = 0; j < L; j += 2) d.push_back((uint8_t)((nib((char) bases[(size_t) j]) << 4) | (j + 1 < L ? nib((char) bases[(size_t) j + 1]) : 0)));
            d.insert(d.end(), quals.begin(), quals.end());
            if (o.bismark_tags) {
                d.insert(d.end(), {'N', 'M', 'C', 3});
                d.insert(d.end(), {'X', 'M', 'Z'}); for (int j = 0; j < L; ++j) d.push_back('.'); d.push_back(0);
                const char *xr = is_read1 ? "CT" : "GA"; d.insert(d.end(), {'X', 'R', 'Z', (uint8_t) xr[0], (uint8_t) xr[1], 0});
                const char *xg = conv_ct ? "CT" : "GA"; d.insert(d.end(), {'X', 'G', 'Z', (uint8_t) xg[0], (uint8_t) xg[1], 0});
            }
            if (nh2) { d.insert(d.end(), {'N', 'H', 'C', 2}); }
            else if (o.bismark_tags && (frag_id & 7) == 0) { d.insert(d.end(), {'N', 'H', 'i', 1, 0, 0, 0}); }
            const uint32_t bs = (uint32_t)(d.size() - start) - 4;
            for (int k = 0; k < 4; ++k) d[start + (size_t) k] = (uint8_t)(bs >> (8 * k));
            S.recs.push_back(RecMeta{(int32_t) pos, rend, ord++, (uint32_t) start, (uint32_t)(d.size() - start)});
        }
    }
    std::sort(S.recs.begin(), S.recs.end(), [](const RecMeta &a, const RecMeta &b) { return a.pos != b.pos ? a.pos < b.pos : a.ord < b.ord; });
}

// one BGZF block (header + raw deflate + crc/isize trailer) into out; returns its size
static size_t bgzf_compress(const uint8_t *data, size_t n, int level, uint8_t *out, size_t out_cap) {
    z_stream zs; memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw std::runtime_error("deflateInit2");
    zs.next_in = (Bytef *) data; zs.avail_in = (uInt) n;
    zs.next_out = out + 18; zs.avail_out = (uInt)(out_cap - 18 - 8);
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) throw std::runtime_error("deflate: block did not fit");
    const size_t clen = zs.total_out;
    deflateEnd(&zs);
    const size_t bsize = 18 + clen + 8;
    static const uint8_t hdr[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    memcpy(out, hdr, 16);
    out[16] = (uint8_t)((bsize - 1) & 0xff); out[17] = (uint8_t)((bsize - 1) >> 8);
    const uint32_t crc = (uint32_t) crc32(crc32(0L, Z_NULL, 0), data, (uInt) n);
    uint8_t *t = out + 18 + clen;
    for (int i = 0; i < 4; ++i) t[i] = (uint8_t)(crc >> (8 * i));
    for (int i = 0; i < 4; ++i) t[4 + i] = (uint8_t)(((uint32_t) n) >> (8 * i));
    return bsize;
}

template <class F> static void parallel_for(size_t n, int threads, F &&f) {
    if (n == 0) return;
    std::atomic<size_t> next{0};
    auto work = [&] { for (;;) { size_t k = next.fetch_add(1); if (k >= n) break; f(k); } };
    std::vector<std::thread> th;
    const int nt = (int) std::min<size_t>((size_t) std::max(1, threads), n);
    for (int t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
}

int main(int argc, char **argv) {
    Opts o;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto val = [&]() -> std::string { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(1); } return argv[++i]; };
        if (a == "--out") o.out = val();
        else if (a == "--contigs") {
            o.contigs.clear();
            std::string s = val(); size_t p = 0;
            while (p < s.size()) { size_t c = s.find(',', p); if (c == std::string::npos) c = s.size(); std::string t = s.substr(p, c - p); size_t k = t.find(':'); o.contigs.emplace_back(t.substr(0, k), (uint32_t) strtoul(t.c_str() + k + 1, nullptr, 10)); p = c + 1; }
        }
        else if (a == "--human") {
            // chr1..22, X, Y in the proportions of GRCh38 (Mbp), scaled to the requested total
            static const double mbp[24] = {248.96, 242.19, 198.30, 190.21, 181.54, 170.81, 159.35, 145.14, 138.39, 133.80, 135.09, 133.28, 114.36, 107.04, 101.99, 90.34, 83.26, 80.37, 58.62, 64.44, 46.71, 50.82, 156.04, 57.23};
            const double total = atof(val().c_str()); double sum = 0; for (double m : mbp) sum += m;
            o.contigs.clear();
            for (int k = 0; k < 24; ++k) { std::string nm = k < 22 ? "chr" + std::to_string(k + 1) : (k == 22 ? "chrX" : "chrY"); o.contigs.emplace_back(nm, (uint32_t) std::max(1000.0, std::floor(total * mbp[k] / sum))); }
        }
        else if (a == "--depth") o.depth = atof(val().c_str());
        else if (a == "--readlen") o.readlen = atoi(val().c_str());
        else if (a == "--isize-mean") o.isize_mean = atof(val().c_str());
        else if (a == "--isize-sd") o.isize_sd = atof(val().c_str());
        else if (a == "--isize-min") o.isize_min = atoi(val().c_str());
        else if (a == "--isize-max") o.isize_max = atoi(val().c_str());
        else if (a == "--genome-seed") o.genome_seed = strtoull(val().c_str(), nullptr, 10);
        else if (a == "--read-seed") o.read_seed = strtoull(val().c_str(), nullptr, 10);
        else if (a == "--level") o.level = atoi(val().c_str());
        else if (a == "--threads") o.threads = atoi(val().c_str());
        else if (a == "--bismark-tags") o.bismark_tags = true;
        else if (a == "--clean") o.clean = true;
        else if (a == "--nondirectional") o.nondirectional = atof(val().c_str());
        else if (a == "--single-frac") o.single_frac = atof(val().c_str());
        else if (a == "--lower-frac") o.lower_frac = atof(val().c_str());
        else if (a == "--n-frac") o.n_frac = atof(val().c_str());
        else if (a == "--quals") o.quals = atoi(val().c_str());
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 1; }
    }
    if (o.out.empty()) { fprintf(stderr, "usage: mdsynth --out PREFIX [options]\n"); return 1; }
    const int L = o.readlen;
    if (o.isize_max < L) o.isize_max = L;
    const int threads = o.threads > 0 ? o.threads : (int) std::max(1u, std::min(64u, std::thread::hardware_concurrency()));

    // ---------------- genome: every contig from its own stream, in parallel
    std::vector<std::string> genome(o.contigs.size());
    parallel_for(o.contigs.size(), threads, [&](size_t c) {
        Rng g(mix(o.genome_seed, c + 1, 0x67656e6full));
        std::string &s = genome[c]; const uint32_t len = o.contigs[c].second;
        s.resize(len);
        for (uint32_t i = 0; i < len;) { uint64_t z = g.next(); for (int k = 0; k < 32 && i < len; ++k, ++i, z >>= 2) s[i] = "ACGT"[z & 3]; }
        // optional soft-masked (lowercase) and N runs, to exercise the case/N rules of isCpG & co
        if (o.lower_frac > 0) for (uint32_t i = 0; i < len;) { if (g.uni() < o.lower_frac / 50) { uint32_t e = std::min(len, i + 20 + g.below(60)); for (; i < e; ++i) s[i] = (char) tolower(s[i]); } else ++i; }
        if (o.n_frac > 0) for (uint32_t i = 0; i < len;) { if (g.uni() < o.n_frac / 30) { uint32_t e = std::min(len, i + 5 + g.below(50)); for (; i < e; ++i) s[i] = 'N'; } else ++i; }
    });
    {
        FILE *fa = fopen((o.out + ".fa").c_str(), "w"), *fai = fopen((o.out + ".fa.fai").c_str(), "w");
        if (!fa || !fai) { fprintf(stderr, "cannot write %s.fa\n", o.out.c_str()); return 1; }
        setvbuf(fa, nullptr, _IOFBF, 8 << 20);
        int64_t off = 0;
        std::string line;
        for (size_t c = 0; c < o.contigs.size(); ++c) {
            const std::string &s = genome[c]; const uint32_t len = o.contigs[c].second;
            off += fprintf(fa, ">%s\n", o.contigs[c].first.c_str());
            fprintf(fai, "%s\t%u\t%lld\t60\t61\n", o.contigs[c].first.c_str(), len, (long long) off);
            for (uint32_t i = 0; i < len; i += 60 * 4096) {             // 4096 lines at a time
                const uint32_t e = std::min<uint64_t>(len, (uint64_t) i + 60 * 4096);
                line.clear();
                for (uint32_t j = i; j < e; j += 60) { const uint32_t n = std::min<uint32_t>(60, e - j); line.append(s.data() + j, n); line.push_back('\n'); }
                fwrite(line.data(), 1, line.size(), fa); off += (int64_t) line.size();
            }
        }
        fclose(fa); fclose(fai);
    }

    // ---------------- alignments
    BamHeader hdr;
    hdr.text = "@HD\tVN:1.6\tSO:coordinate\n";
    for (auto &c : o.contigs) { hdr.names.push_back(c.first); hdr.lens.push_back(c.second); hdr.text += "@SQ\tSN:" + c.first + "\tLN:" + std::to_string(c.second) + "\n"; }
    hdr.text += "@PG\tID:mdsynth\tPN:mdsynth\n";
    uint64_t file_off = 0;
    {
        BgzfWriter bw(o.out + ".bam", o.level);
        write_bam_header(bw, hdr);
        bw.flush_block();
        file_off = bw.tell() >> 16;
        bw.close_no_eof();
    }
    FILE *fp = fopen((o.out + ".bam").c_str(), "ab");
    if (!fp) { fprintf(stderr, "cannot append to %s.bam\n", o.out.c_str()); return 1; }
    setvbuf(fp, nullptr, _IOFBF, 8 << 20);
    BaiBuilder bai(o.contigs.size());
    uint64_t n_records = 0, n_frag = 0, slab_global = 0;

    struct OutRec { int32_t pos, end; uint64_t uoff; uint32_t len; };
    std::vector<uint8_t> ublob, spill_bytes, spill_next;           // merged records of the wave; records starting beyond their slab
    std::vector<RecMeta> spill, spill_n;
    std::vector<OutRec> outrecs;
    std::vector<std::vector<uint8_t>> cblocks;
    std::vector<size_t> csize;
    for (size_t tid = 0; tid < o.contigs.size(); ++tid) {
        const std::string &ref = genome[tid];
        const int64_t G = (int64_t) ref.size();
        if (G < 2 * L + 10) continue;
        const double gap = 2.0 * L / o.depth;
        int64_t slab_bp = (int64_t) std::max(4096.0, std::ceil(50000.0 * gap));
        slab_bp = std::max<int64_t>(slab_bp, 4 * (int64_t)(o.isize_max + L));
        const size_t n_slabs = (size_t)((G + slab_bp - 1) / slab_bp);
        const size_t wave = 64;                                  // constant: where a wave ends a BGZF block ends, and the file must not depend on the thread count
        spill.clear(); spill_bytes.clear();
        for (size_t s0 = 0; s0 < n_slabs; s0 += wave) {
            const size_t ns = std::min(wave, n_slabs - s0);
            std::vector<Slab> slabs(ns);
            parallel_for(ns, threads, [&](size_t k) { const int64_t b = (int64_t)(s0 + k) * slab_bp; gen_slab(o, tid, ref, slab_global + s0 + k, b, std::min(G, b + slab_bp), slabs[k]); });
            // merge in slab order: the previous slab's spill (records starting at or beyond this slab's start) with this slab's
            // records up to its end; what starts beyond the end waits for the next slab
            ublob.clear(); outrecs.clear();
            const bool last_wave = s0 + ns >= n_slabs;
            for (size_t k = 0; k < ns; ++k) {
                Slab &S = slabs[k]; n_frag += S.n_frag;
                const int64_t send = (last_wave && k + 1 == ns) ? INT64_MAX : (int64_t)(s0 + k + 1) * slab_bp;
                spill_n.clear(); spill_next.clear();
                size_t a = 0, b = 0;
                auto take = [&](const RecMeta &m, const std::vector<uint8_t> &src) {
                    if ((int64_t) m.pos < send) { outrecs.push_back(OutRec{m.pos, m.end, (uint64_t) ublob.size(), m.len}); ublob.insert(ublob.end(), src.begin() + m.off, src.begin() + m.off + m.len); }
                    else { RecMeta c = m; c.off = (uint32_t) spill_next.size(); spill_next.insert(spill_next.end(), src.begin() + m.off, src.begin() + m.off + m.len); spill_n.push_back(c); }
                };
                while (a < spill.size() || b < S.recs.size()) {
                    const bool from_spill = b >= S.recs.size() || (a < spill.size() && spill[a].pos <= S.recs[b].pos);   // ties: the earlier slab's record first
                    if (from_spill) take(spill[a++], spill_bytes); else take(S.recs[b++], S.bytes);
                }
                spill.swap(spill_n); spill_bytes.swap(spill_next);
                std::vector<uint8_t>().swap(S.bytes);
            }
            // BGZF blocks of the wave, compressed in parallel, written in order
            const size_t nb = (ublob.size() + kBgzfBlock - 1) / kBgzfBlock;
            if (cblocks.size() < nb) { cblocks.resize(nb); csize.resize(nb); }
            parallel_for(nb, threads, [&](size_t k) {
                if (cblocks[k].size() < 65536 + 64) cblocks[k].resize(65536 + 64);
                const size_t off = k * (size_t) kBgzfBlock, n = std::min((size_t) kBgzfBlock, ublob.size() - off);
                csize[k] = bgzf_compress(ublob.data() + off, n, o.level, cblocks[k].data(), cblocks[k].size());
            });
            std::vector<uint64_t> boff(nb + 1);
            for (size_t k = 0; k < nb; ++k) { boff[k] = file_off; if (fwrite(cblocks[k].data(), 1, csize[k], fp) != csize[k]) { fprintf(stderr, "short write\n"); return 1; } file_off += csize[k]; }
            boff[nb] = file_off;
            auto voff = [&](uint64_t u) { const size_t k = (size_t)(u / kBgzfBlock); return k >= nb ? (boff[nb] << 16) : ((boff[k] << 16) | (u % kBgzfBlock)); };
            for (const OutRec &r : outrecs) { bai.add((int) tid, r.pos, r.end, voff(r.uoff), voff(r.uoff + r.len)); ++n_records; }
        }
        slab_global += n_slabs;
    }
    {
        uint8_t eof[64]; const size_t n = bgzf_compress(nullptr, 0, o.level, eof, sizeof eof);
        fwrite(eof, 1, n, fp);
    }
    fclose(fp);
    bai.write(o.out + ".bam.bai");
    fprintf(stderr, "mdsynth: %llu records, %llu fragments -> %s.{fa,bam}\n", (unsigned long long) n_records, (unsigned long long) n_frag, o.out.c_str());
    printf("%llu\n", (unsigned long long) n_records);
    return 0;
}
