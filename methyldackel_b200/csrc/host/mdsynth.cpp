// mdsynth — deterministic synthetic WGBS data set generator (FASTA + .fai, coordinate-sorted
// BAM + .bai) following the recipe in SURVEY.md section 8d.  Used by the tests, by bench.py
// and by the CPU reference arm (which needs real files).  Not part of the extract/mbias path.
//
//   mdsynth --out PREFIX [--contigs name:len,...] [--depth 30] [--readlen 150]
//           [--isize-mean 300 --isize-sd 60 --isize-min 150 --isize-max 800]
//           [--genome-seed 1234] [--read-seed 5678] [--level 1]
//           [--bismark-tags] [--nondirectional F] [--single-frac F] [--lower-frac F] [--n-frac F]
//           [--quals N]  (N distinct phred values instead of the 4 binned ones)
//           [--clean]   (no flag noise / indels / clips: every record is a plain properly paired 150M)
#include "hostio.hpp"
#include <cmath>
#include <queue>
#include <cstdlib>
#include <memory>

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

struct Opts {
    std::string out;
    std::vector<std::pair<std::string, uint32_t>> contigs{{"chr1", 100000}};
    double depth = 30; int readlen = 150;
    double isize_mean = 300, isize_sd = 60; int isize_min = 150, isize_max = 800;
    uint64_t genome_seed = 1234, read_seed = 5678;
    int level = 1;
    bool bismark_tags = false, clean = false;
    double nondirectional = 0.0, single_frac = 0.0, lower_frac = 0.0, n_frac = 0.0;
    int quals = 4;          // distinct phred values: 4 = binned Illumina-like; more = uniform over [2, 2+quals)
};

struct Rec { int32_t pos; uint64_t order; std::vector<uint8_t> data; int32_t end; };
struct RecCmp { bool operator()(const std::shared_ptr<Rec> &a, const std::shared_ptr<Rec> &b) const { return a->pos != b->pos ? a->pos > b->pos : a->order > b->order; } };

static uint8_t nib(char c) { switch (c) { case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': return 8; default: return 15; } }

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
        else if (a == "--depth") o.depth = atof(val().c_str());
        else if (a == "--readlen") o.readlen = atoi(val().c_str());
        else if (a == "--isize-mean") o.isize_mean = atof(val().c_str());
        else if (a == "--isize-sd") o.isize_sd = atof(val().c_str());
        else if (a == "--isize-min") o.isize_min = atoi(val().c_str());
        else if (a == "--isize-max") o.isize_max = atoi(val().c_str());
        else if (a == "--genome-seed") o.genome_seed = strtoull(val().c_str(), nullptr, 10);
        else if (a == "--read-seed") o.read_seed = strtoull(val().c_str(), nullptr, 10);
        else if (a == "--level") o.level = atoi(val().c_str());
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

    // ---------------- genome
    std::vector<std::string> genome(o.contigs.size());
    {
        Rng g(o.genome_seed);
        FILE *fa = fopen((o.out + ".fa").c_str(), "w"), *fai = fopen((o.out + ".fa.fai").c_str(), "w");
        if (!fa || !fai) { fprintf(stderr, "cannot write %s.fa\n", o.out.c_str()); return 1; }
        int64_t off = 0;
        for (size_t c = 0; c < o.contigs.size(); ++c) {
            std::string &s = genome[c]; uint32_t len = o.contigs[c].second;
            s.resize(len);
            for (uint32_t i = 0; i < len; ++i) s[i] = "ACGT"[g.next() >> 62];
            // optional soft-masked (lowercase) and N runs, to exercise the case/N rules of isCpG & co
            if (o.lower_frac > 0) for (uint32_t i = 0; i < len;) { if (g.uni() < o.lower_frac / 50) { uint32_t e = std::min(len, i + 20 + g.below(60)); for (; i < e; ++i) s[i] = (char) tolower(s[i]); } else ++i; }
            if (o.n_frac > 0) for (uint32_t i = 0; i < len;) { if (g.uni() < o.n_frac / 30) { uint32_t e = std::min(len, i + 5 + g.below(50)); for (; i < e; ++i) s[i] = 'N'; } else ++i; }
            off += fprintf(fa, ">%s\n", o.contigs[c].first.c_str());
            fprintf(fai, "%s\t%u\t%lld\t60\t61\n", o.contigs[c].first.c_str(), len, (long long) off);
            for (uint32_t i = 0; i < len; i += 60) { uint32_t n = std::min<uint32_t>(60, len - i); fwrite(s.data() + i, 1, n, fa); fputc('\n', fa); off += n + 1; }
        }
        fclose(fa); fclose(fai);
    }

    // ---------------- alignments
    BamHeader hdr;
    hdr.text = "@HD\tVN:1.6\tSO:coordinate\n";
    for (auto &c : o.contigs) { hdr.names.push_back(c.first); hdr.lens.push_back(c.second); hdr.text += "@SQ\tSN:" + c.first + "\tLN:" + std::to_string(c.second) + "\n"; }
    hdr.text += "@PG\tID:mdsynth\tPN:mdsynth\n";
    BgzfWriter bw(o.out + ".bam", o.level);
    write_bam_header(bw, hdr);
    bw.flush_block();
    BaiBuilder bai(o.contigs.size());
    Rng r(o.read_seed);
    uint64_t n_records = 0, frag_id = 0, order = 0;

    for (size_t tid = 0; tid < o.contigs.size(); ++tid) {
        const std::string &ref = genome[tid];
        const int64_t G = (int64_t) ref.size();
        if (G < 2 * L + 10) continue;
        // per-position methylation probability (truth); CpG beta from the bimodal mixture, else 1 %
        std::vector<uint8_t> beta(ref.size());
        {
            Rng m(o.genome_seed ^ (0x5151ull + tid));
            for (int64_t i = 0; i < G; ++i) {
                char c = (char) toupper(ref[i]); bool cpg = false;
                if (c == 'C' && i + 1 < G && toupper(ref[i + 1]) == 'G') cpg = true;
                if (c == 'G' && i > 0 && toupper(ref[i - 1]) == 'C') cpg = true;
                double b = 0.01;
                if (cpg) { if (c == 'G') { beta[i] = beta[i - 1]; continue; } b = (m.uni() < 0.8) ? m.beta_int(8, 2) : m.beta_int(1, 8); }
                beta[i] = (uint8_t) std::lround(b * 255.0);
            }
        }
        const double n_frag = o.depth * (double) G / (2.0 * L);
        const double gap = (double) G / n_frag;
        std::priority_queue<std::shared_ptr<Rec>, std::vector<std::shared_ptr<Rec>>, RecCmp> heap;
        auto flush_upto = [&](int64_t pos) {
            while (!heap.empty() && heap.top()->pos <= pos) {
                auto rec = heap.top(); heap.pop();
                uint64_t v0 = bw.tell();
                bw.write(rec->data.data(), rec->data.size());
                uint64_t v1 = bw.tell();
                bai.add((int) tid, rec->pos, rec->end, v0, v1);
                ++n_records;
            }
        };
        double cursor = r.expo() * gap;
        while (cursor < (double)(G - 1)) {
            int64_t s = (int64_t) cursor;
            cursor += r.expo() * gap;
            flush_upto(s - 1);
            int isize = (int) std::lround(o.isize_mean + o.isize_sd * r.normal());
            isize = std::max(o.isize_min, std::min(o.isize_max, isize));
            if (isize < L) isize = L;
            if (s + isize + 8 > G) continue;
            ++frag_id;
            char qname[32]; int lq = snprintf(qname, sizeof qname, "f%09llu", (unsigned long long) frag_id) + 1;
            // library strand: OT / OB, optionally CTOT / CTOB (needs the XG tag to be told apart)
            bool ob = r.uni() < 0.5;
            bool compl_strand = o.nondirectional > 0 && r.uni() < o.nondirectional;
            bool single = o.single_frac > 0 && r.uni() < o.single_frac;
            // record-level noise shared by the pair
            uint8_t mapq; { double u = r.uni(); mapq = u < 0.88 ? 60 : u < 0.93 ? 30 : u < 0.97 ? 9 : 0; }
            bool dup = !o.clean && r.uni() < 0.02, qcfail = !o.clean && r.uni() < 0.005, secondary = !o.clean && r.uni() < 0.005;
            bool singleton = !o.clean && !single && r.uni() < 0.01, improper = !o.clean && !single && r.uni() < 0.02, nh2 = !o.clean && r.uni() < 0.005;
            if (o.clean) mapq = 60;
            int64_t left_pos = s, right_pos = s + isize - L;
            for (int mate = 0; mate < 2; ++mate) {   // mate 0 = leftmost record, 1 = rightmost
                if (single && mate == 1) break;
                if (singleton && mate == 1) break;
                bool is_left = mate == 0;
                // OT: read1 forward on the left, read2 reverse on the right.  OB: read2 forward on the left, read1 reverse on the right.
                // CTOT (XG=CT): read1 reverse (right), read2 forward (left).  CTOB (XG=GA): read1 forward (left), read2 reverse (right).
                bool conv_ct = !ob;                           // which conversion the bases carry, in reference orientation
                bool read1_left = compl_strand ? ob : !ob;    // OT,CTOB: read1 is the left (forward) record
                bool is_read1 = (is_left == read1_left);
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
                int64_t pos = is_left ? left_pos : right_pos;
                // CIGAR shape
                std::vector<std::pair<int, int>> cig;   // (op, len) ; ops: 0 M, 1 I, 2 D, 4 S
                {
                    double u = o.clean ? 0.0 : r.uni();
                    int clip5 = 0, clip3 = 0, indel_at = -1, indel_len = 0; bool ins = false;
                    if (u >= 0.94 && u < 0.97) { if (r.uni() < 0.5) clip5 = 1 + (int) r.below(20); else clip3 = 1 + (int) r.below(20); }
                    else if (u >= 0.97 && u < 0.99) { indel_len = 1 + (int) r.below(3); ins = r.uni() < 0.5; }
                    else if (u >= 0.99) { if (r.uni() < 0.5) clip5 = 1 + (int) r.below(20); else clip3 = 1 + (int) r.below(20); indel_len = 1 + (int) r.below(3); ins = r.uni() < 0.5; }
                    int body = L - clip5 - clip3;
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
                std::vector<uint8_t> bases((size_t) L), quals((size_t) L);
                int q = 0; int64_t p = pos;
                auto emit_ref = [&](int64_t rp) {
                    char c = (char) toupper(ref[(size_t) rp]);
                    if (conv_ct && c == 'C') { if (r.uni() * 255.0 >= beta[(size_t) rp]) c = 'T'; }
                    else if (!conv_ct && c == 'G') { if (r.uni() * 255.0 >= beta[(size_t) rp]) c = 'A'; }
                    return c;
                };
                for (auto &oplen : cig) {
                    int op = oplen.first, len = oplen.second;
                    for (int j = 0; j < len; ++j) {
                        if (op == 0) { bases[(size_t) q++] = (uint8_t) emit_ref(p++); }
                        else if (op == 1 || op == 4) { bases[(size_t) q++] = (uint8_t) "ACGT"[r.next() >> 62]; }
                        else if (op == 2) ++p;
                    }
                }
                int32_t end = (int32_t) p;
                if (end > G) continue;
                for (int j = 0; j < L; ++j) {
                    double u = r.uni(); uint8_t ql = u < 0.75 ? 37 : u < 0.90 ? 25 : u < 0.98 ? 11 : 2;
                    if (o.quals != 4) ql = (uint8_t)(2 + r.below((uint32_t) o.quals));
                    quals[(size_t) j] = ql;
                    if (r.uni() < std::pow(10.0, -ql / 10.0)) { char c; do c = "ACGT"[r.next() >> 62]; while (c == (char) bases[(size_t) j]); bases[(size_t) j] = (uint8_t) c; }
                }
                // encode
                auto rec = std::make_shared<Rec>();
                std::vector<uint8_t> &d = rec->data;
                int64_t mpos = single ? -1 : (is_left ? right_pos : left_pos);
                int32_t tlen = single ? 0 : (is_left ? isize : -isize);
                put32(d, 0);  // block_size placeholder
                put32(d, (uint32_t) tid); put32(d, (uint32_t) pos);
                put32(d, ((uint32_t) reg2bin(pos, end > pos ? end : pos + 1) << 16) | ((uint32_t) mapq << 8) | (uint32_t) lq);
                put32(d, ((uint32_t) flag << 16) | (uint32_t) cig.size());
                put32(d, (uint32_t) L);
                put32(d, single ? 0xffffffffu : (uint32_t) tid); put32(d, (uint32_t) mpos); put32(d, (uint32_t) tlen);
                d.insert(d.end(), qname, qname + lq);
                for (auto &oplen : cig) put32(d, ((uint32_t) oplen.second << 4) | (uint32_t) oplen.first);
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
                uint32_t bs = (uint32_t) d.size() - 4;
                for (int k = 0; k < 4; ++k) d[(size_t) k] = (uint8_t)(bs >> (8 * k));
                rec->pos = (int32_t) pos; rec->end = end; rec->order = order++;
                heap.push(rec);
            }
        }
        flush_upto(INT32_MAX);
    }
    bw.close();
    bai.write(o.out + ".bam.bai");
    fprintf(stderr, "mdsynth: %llu records, %llu fragments -> %s.{fa,bam}\n", (unsigned long long) n_records, (unsigned long long) frag_id, o.out.c_str());
    printf("%llu\n", (unsigned long long) n_records);
    return 0;
}
