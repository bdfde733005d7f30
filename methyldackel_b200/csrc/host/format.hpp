// Host-side output stage of `extract`: turns the per-column counts that come back from the
// device (md_call records) into the reference's text formats, byte for byte.
//
// What is emulated, and where it lives in the reference:
//   * the chunk cursor + adjustBounds (extract.c:325-350,376-378; common.c:466-493): the GPU tiles
//     the genome however it likes, but --mergeContext pairing and the cytosine_report blanks are
//     flushed at the REFERENCE's chunk ends (extract.c:496-510), so the writer replays those chunks;
//   * writeCall's six line formats (extract.c:39-99), processLast (extract.c:207-222),
//     writeBlank + getTriNucContext (extract.c:120-205), printHeader (extract.c:562-569);
//   * the variant-exclusion side effects on a pending merged call (extract.c:444-459).
#pragma once
#include <cstdio>
#include <cstdarg>
#include <cstdint>
#include <cmath>
#include <string>
#include <vector>
#include <cinttypes>
#include <cstring>
#include <algorithm>
#include "../../../include/mdgpu.h"

namespace mdhost {

struct ExtractOptions {
    md_config core;
    int minDepth = 1;
    int merge = 0, methylKit = 0, fraction = 0, counts = 0, logit = 0, cytosine_report = 0;
    unsigned long chunkSize = 1000000;
    ExtractOptions() {
        memset(&core, 0, sizeof core);
        core.keepCpG = 1; core.minMapq = 10; core.minPhred = 5; core.ignoreFlags = 0xF00;   // extract.c:725-746
    }
};

// ---- context helpers on a whole contig (equivalent to the reference's window-relative calls, because
// the extract window is contig[localPos-2 .. localEnd+10], extract.c:369-381, so only contig ends clip) ----
inline bool isC(char b) { return b == 'C' || b == 'c'; }
inline bool isG(char b) { return b == 'G' || b == 'g'; }
// 0 none, +-1 CpG, +-2 CHG, +-3 CHH (sign: + for C, - for G); common.c:49-82 chained as extract.c:407-418
inline int context_at(const char *seq, int64_t pos, int64_t seqlen) {
    if (pos < 0 || pos >= seqlen) return 0;
    if (isC(seq[pos])) {
        if (pos + 1 != seqlen && isG(seq[pos + 1])) return 1;
        if (pos + 2 < seqlen && isG(seq[pos + 2])) return 2;
        return 3;
    }
    if (isG(seq[pos])) {
        if (pos != 0 && isC(seq[pos - 1])) return -1;
        if (pos > 1 && isC(seq[pos - 2])) return -2;
        return -3;
    }
    return 0;
}

struct Chunk { uint32_t tid, beg, end; };

// Replays the (globalTid, globalPos, globalEnd) cursor of extract.c:325-350 / MBias.c:111-135.
// `fetch(tid)` must return the contig's bases (or nullptr if the FASTA lacks it).
class ChunkCursor {
public:
    ChunkCursor(const std::vector<uint32_t> &target_len, unsigned long chunkSize, uint32_t gTid, uint32_t gPos, uint32_t gEnd)
        : len_(target_len), chunk_(chunkSize), gTid_(gTid), gPos_(gPos), gEnd_(gEnd) {}
    template <class Fetch> bool next(Chunk &c, Fetch &&fetch) {
        uint32_t n_targets = (uint32_t) len_.size();
        uint32_t localTid = gTid_, localPos = gPos_, localEnd = (uint32_t)(localPos + chunk_);
        if (localTid >= n_targets) return false;
        if (gEnd_ && localEnd > gEnd_) localEnd = gEnd_;
        // adjustBounds, common.c:466-493: look at contig[localEnd-1 .. localEnd+1]
        {
            // fetch(tid, start, n, out) -> length of the contig in the FASTA (0 if absent), out = bases [start, start+n) clamped
            int64_t start = localEnd > 0 ? (int64_t) localEnd - 1 : 0, end = (int64_t) localEnd + 1;
            std::string w;
            const int64_t L = fetch(localTid, start, (int64_t) 3, w);
            // faidx_fetch_seq clamping (end inclusive)
            if (start >= L) start = L;
            if (end >= L) end = L - 1;
            int64_t seqlen = end + 1 - start; if (seqlen < 0) seqlen = 0;
            if (L > 0 && seqlen > 1 && (int64_t) w.size() >= seqlen) {
                const char *q = w.data();
                if (seqlen > 2 && (q[0] & 0x5F) == 'C' && (q[2] & 0x5F) == 'G') localEnd += 2;
                else if ((q[1] & 0x5F) == 'G') localEnd += 1;
            }
            if (localPos > localEnd) std::swap(localPos, localEnd);
        }
        gPos_ = localEnd;
        if (gEnd_ > 0 && gPos_ >= gEnd_) gTid_ = (uint32_t) -1;
        if (localTid < n_targets && gTid_ != (uint32_t) -1) {
            if (gPos_ >= len_[localTid]) { localEnd = len_[localTid]; gTid_++; gPos_ = 0; }
        }
        if (gEnd_ && localPos >= gEnd_) return false;
        c.tid = localTid; c.beg = localPos; c.end = localEnd;
        return true;
    }
private:
    const std::vector<uint32_t> &len_;
    unsigned long chunk_;
    uint32_t gTid_, gPos_, gEnd_;
};

// Output text of one chunk for one file: a plain growable byte buffer (a FILE* costs a locked call per line, and this
// stage writes ~10^9 lines on a genome).
struct TextBuf {
    std::vector<char> v; size_t n = 0;
    char *room(size_t k) { if (n + k > v.size()) v.resize(std::max(v.size() * 2, n + k + (1u << 16))); return v.data() + n; }
    void put(const char *p, size_t k) { memcpy(room(k), p, k); n += k; }
    void printf_(const char *fmt, ...) __attribute__((format(printf, 2, 3))) {
        va_list ap; va_start(ap, fmt);
        char *w = room(1024);
        int k = vsnprintf(w, 1024, fmt, ap);
        va_end(ap);
        if (k >= 1024) { va_start(ap, fmt); w = room((size_t) k + 1); vsnprintf(w, (size_t) k + 1, fmt, ap); va_end(ap); }
        if (k > 0) n += (size_t) k;
    }
};

class ExtractWriter {
public:
    // out[k]: where the lines of context k (CpG, CHG, CHH) go; the three may be the same buffer (cytosine_report), or null
    ExtractWriter(const ExtractOptions &o, TextBuf *out[3]) : o_(o) { fp_[0] = out[0]; fp_[1] = out[1]; fp_[2] = out[2]; }
    uint64_t n_variant_positions() const { return nVariant_; }

    // printHeader, extract.c:562-569
    static void print_header(FILE *of, const char *context, const char *opref, const ExtractOptions &o) {
        fprintf(of, "track type=\"bedGraph\" description=\"%s %s", opref, context);
        if (o.merge) fprintf(of, " merged");
        if (o.fraction) fprintf(of, " methylation fractions\"\n");
        else if (o.counts) fprintf(of, " methylation counts\"\n");
        else if (o.logit) fprintf(of, " logit transformed methylation fractions\"\n");
        else fprintf(of, " methylation levels\"\n");
    }

    // One reference chunk [beg,end) of contig `chrom`; `calls` are that chunk's records, ascending.
    void process_chunk(const char *chrom, const std::string &ref, uint32_t beg, uint32_t end, const md_call *calls, size_t n) {
        Last lastCpG, lastCHG;
        uint32_t lastPos = beg;
        const char *seq = ref.data(); int64_t L = (int64_t) ref.size();
        for (size_t i = 0; i < n; ++i) {
            const md_call &c = calls[i];
            int type = (int) MD_CALL_CTX(c.info);
            int32_t pos = (int32_t) c.pos;
            char base = seq[c.pos];
            bool g = isG(base);
            if (MD_CALL_EXCLUDED(c.info)) {                       // extract.c:444-459
                ++nVariant_;
                if (o_.merge) {
                    if (type == 0 && lastCpG.live && lastCpG.pos == pos - 1 && g) { lastCpG.nm = 0; lastCpG.nu = 0; }
                    else if (type == 1 && lastCHG.live && lastCHG.pos == pos - 2 && g) { lastCHG.nm = 0; lastCHG.nu = 0; }
                }
                continue;
            }
            if (c.nmeth + c.nunmeth == 0 && !o_.cytosine_report) continue;   // extract.c:461
            if (!o_.merge || type == 2) {
                if (o_.cytosine_report) {
                    write_blank(chrom, pos, &lastPos, seq, L);
                    const char *context = type == 0 ? "G" : type == 1 ? "HG" : "HH";
                    int tnc = trinuc(seq, c.pos, L, g ? -1 : 1);
                    write_call(fp_[0], chrom, pos, 1, c.nmeth, c.nunmeth, base, context, kTri[tnc]);
                } else write_call(fp_[type], chrom, pos, 1, c.nmeth, c.nunmeth, base, nullptr, nullptr);
            } else if (type == 0) {
                if (g) pos--;
                process_last(fp_[0], chrom, lastCpG, pos, 2, c.nmeth, c.nunmeth, base);
            } else {
                if (g) pos -= 2;
                process_last(fp_[1], chrom, lastCHG, pos, 3, c.nmeth, c.nunmeth, base);
            }
            lastPos = (uint32_t)(pos + 1);
        }
        if (o_.merge) {                                          // extract.c:499-507
            if (o_.core.keepCpG && lastCpG.live) write_call(fp_[0], chrom, lastCpG.pos, 2, lastCpG.nm, lastCpG.nu, 'C', nullptr, nullptr);
            if (o_.core.keepCHG && lastCHG.live) write_call(fp_[1], chrom, lastCHG.pos, 3, lastCHG.nm, lastCHG.nu, 'C', nullptr, nullptr);
        } else if (o_.cytosine_report) write_blank(chrom, (int32_t) end, &lastPos, seq, L);
    }

private:
    struct Last { bool live = false; int32_t pos = 0; uint32_t nm = 0, nu = 0; };
    static constexpr const char *kTri[25] = {"CAA", "CAC", "CAG", "CAT", "CAN", "CCA", "CCC", "CCG", "CCT", "CCN", "CGA", "CGC", "CGG", "CGT", "CGN",
                                             "CTA", "CTC", "CTG", "CTT", "CTN", "CNA", "CNC", "CNG", "CNT", "CNN"};   // extract.c:33-37

    static char revcomp(char b) {
        switch (b) { case 'A': case 'a': return 'T'; case 'C': case 'c': return 'G'; case 'G': case 'g': return 'C'; case 'T': case 't': return 'A'; default: return 'N'; }
    }
    static int base_code(char b) { switch (b) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; } }
    // getTriNucContext, extract.c:120-180, on whole-contig coordinates
    static int trinuc(const char *seq, int64_t off, int64_t seqlen, int dir) {
        int rv;
        if ((dir > 0 && off + 2 >= seqlen) || (dir < 0 && off <= 1)) rv = 4;
        else { char b = seq[off + 2 * dir]; if (dir < 0) b = revcomp(b); rv = base_code(b); }
        if ((dir > 0 && off + 1 >= seqlen) || (dir < 0 && off == 0)) rv += 20;
        else { char b = seq[off + dir]; if (dir < 0) b = revcomp(b); rv += 5 * base_code(b); }
        return rv;
    }
    static double logit(double p) { return log(p) - log(1 - p); }
    // decimal digits of v; two at a time from a table, written back to front into their final place
    static char *put_uint(char *w, uint32_t v) {
        static const char D2[201] = "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";
        const int n = v < 10u ? 1 : v < 100u ? 2 : v < 1000u ? 3 : v < 10000u ? 4 : v < 100000u ? 5 : v < 1000000u ? 6 : v < 10000000u ? 7 : v < 100000000u ? 8 : v < 1000000000u ? 9 : 10;
        char *e = w + n, *q = e;
        while (v >= 100u) { const uint32_t r = v % 100u; v /= 100u; q -= 2; memcpy(q, D2 + 2 * r, 2); }
        if (v >= 10u) { q -= 2; memcpy(q, D2 + 2 * v, 2); } else *--q = (char)('0' + v);
        return e;
    }
    static char *put_int(char *w, int32_t v) { if (v < 0) { *w++ = '-'; return put_uint(w, (uint32_t)(-(int64_t) v)); } return put_uint(w, (uint32_t) v); }

    // writeCall, extract.c:39-99
    void write_call(TextBuf *f, const char *chrom, int32_t pos, int32_t width, uint32_t nm, uint32_t nu, char base, const char *context, const char *tnc) {
        char strand = (base == 'C' || base == 'c') ? 'F' : 'R';
        if ((nm + nu) < (uint32_t) o_.minDepth && !o_.cytosine_report) return;   // unsigned compare, as in C
        if (!o_.fraction && !o_.logit && !o_.counts && !o_.methylKit && !o_.cytosine_report) {
            // "%s\t%i\t%i\t%i\t%u\t%u\n" (extract.c:45-52) assembled by hand: this line is written ~10^9 times on a genome
            if (chrom != chrom_) { chrom_ = chrom; chrom_len_ = strlen(chrom); }
            char *w0 = f->room(chrom_len_ + 80), *w = w0;
            memcpy(w, chrom, chrom_len_); w += chrom_len_; *w++ = '\t';
            w = put_int(w, pos); *w++ = '\t'; w = put_int(w, pos + width); *w++ = '\t';
            // (int)(100.0 * nm / (nm + nu)) of extract.c:50 in integers: the double quotient of two 32-bit counts is never close
            // enough to an integer it does not equal for the truncation to differ (gap >= 2^-32, rounding error <= 2^-46)
            const uint32_t tot = nm + nu;                                   // 32-bit sum, as in the reference
            w = put_int(w, tot ? (int)(100ull * nm / tot) : (int)(100.0 * ((double) nm) / tot)); *w++ = '\t';
            w = put_uint(w, nm); *w++ = '\t'; w = put_uint(w, nu); *w++ = '\n';
            f->n += (size_t)(w - w0);
        }
        else if (o_.fraction) f->printf_("%s\t%i\t%i\t%f\n", chrom, pos, pos + width, ((double) nm) / (nm + nu));
        else if (o_.counts) f->printf_("%s\t%i\t%i\t%i\n", chrom, pos, pos + width, nm + nu);
        else if (o_.logit) f->printf_("%s\t%i\t%i\t%f\n", chrom, pos, pos + width, logit(((double) nm) / (nm + nu)));
        else if (o_.methylKit)
            f->printf_("%s.%i\t%s\t%i\t%c\t%i\t%6.2f\t%6.2f\n", chrom, pos + 1, chrom, pos + 1, strand, nm + nu, 100.0 * ((double) nm) / (nm + nu), 100.0 * ((double) nu) / (nm + nu));
        else if (o_.cytosine_report) {
            strand = (base == 'C' || base == 'c') ? '+' : '-';
            f->printf_("%s\t%i\t%c\t%" PRIu32 "\t%" PRIu32 "\tC%s\t%s\n", chrom, pos + 1, strand, nm, nu, context, tnc);
        }
    }
    // processLast, extract.c:207-222
    void process_last(TextBuf *f, const char *chrom, Last &last, int32_t pos, int width, uint32_t nm, uint32_t nu, char base) {
        if (last.live && last.pos == pos) {
            write_call(f, chrom, pos, width, nm + last.nm, nu + last.nu, base, nullptr, nullptr);
            last.live = false;
        } else {
            if (last.live) write_call(f, chrom, last.pos, width, last.nm, last.nu, base, nullptr, nullptr);
            last.live = true; last.pos = pos; last.nm = nm; last.nu = nu;
        }
    }
    // writeBlank, extract.c:182-205
    void write_blank(const char *chrom, int32_t pos, uint32_t *lastPos, const char *seq, int64_t L) {
        if (pos == -1) return;
        for (; (int64_t) *lastPos < (int64_t) pos; (*lastPos)++) {
            int ctx = context_at(seq, *lastPos, L);
            if (ctx == 0) continue;
            int type = (ctx < 0 ? -ctx : ctx) - 1;
            if ((type == 0 && !o_.core.keepCpG) || (type == 1 && !o_.core.keepCHG) || (type == 2 && !o_.core.keepCHH)) continue;
            int tnc = trinuc(seq, *lastPos, L, ctx > 0 ? 1 : -1);
            write_call(fp_[0], chrom, (int32_t) *lastPos, 1, 0, 0, ctx > 0 ? 'C' : 'G', type == 0 ? "G" : type == 1 ? "HG" : "HH", kTri[tnc]);
        }
    }
    ExtractOptions o_;
    TextBuf *fp_[3];
    const char *chrom_ = nullptr; size_t chrom_len_ = 0;
    uint64_t nVariant_ = 0;
};

}  // namespace mdhost
