// -l <BED> (+ --keepStrand) on the host side: file parsing with the reference's leniencies (parseBED, bed.c:91-236), the
// reference's region order (sortBED, bed.c:64-85) and the chunk-level test that makes a worker skip a whole chunk
// (spanOverlapsBED on the chunk, extract.c:353-367 / MBias.c:139 / perRead.c:159-173).  The per-read and per-column tests run
// on the device (md_set_bed, include/mdgpu.h).
#pragma once
#include <zlib.h>
#include <cctype>
#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>
#include "../../../include/mdgpu.h"

struct BedFile {
    std::vector<std::vector<md_bed_region>> by_tid;     // sorted per contig
    size_t n = 0;

    // false (after the reference's message on stderr) when the file cannot be used
    bool load(const char *fn, const std::vector<std::string> &names, const std::vector<uint32_t> &lens, bool keepStrand) {
        gzFile fp = gzopen(fn, "r");
        if (!fp) { fprintf(stderr, "Couldn't open %s for reading.\n", fn); return false; }
        std::string text; char buf[1 << 16]; int got;
        while ((got = gzread(fp, buf, sizeof buf)) > 0) text.append(buf, (size_t) got);
        gzclose(fp);
        by_tid.assign(names.size(), {});
        int32_t lnum = 0; size_t at = 0;
        while (at < text.size()) {
            size_t nl = text.find('\n', at);
            std::string line = text.substr(at, nl == std::string::npos ? std::string::npos : nl - at);
            at = nl == std::string::npos ? text.size() : nl + 1;
            ++lnum;
            if (line.empty()) break;                           // ks_getuntil() returns the line length and parseBED loops while it is > 0 (bed.c:118): an empty line ends the file
            if (line[0] == '#') continue;
            const char *s = line.c_str(), *p = s;
            while (*p && !isspace((unsigned char) *p)) ++p;
            const std::string name(s, (size_t)(p - s));
            int tid = -1;
            for (size_t i = 0; i < names.size(); ++i) if (names[i] == name) { tid = (int) i; break; }
            if (tid < 0) {
                if (name == "track" || name == "browser") continue;
                fprintf(stderr, "Couldn't properly parse line number %i in %s.\n", lnum, fn); return false;
            }
            int32_t start = -1, end = -1;
            if (!*p || sscanf(p + 1, "%" SCNd32, &start) != 1 || start == -1) { fprintf(stderr, "Line %" PRId32 " of %s is malformed.\n", lnum, fn); return false; }
            ++p;
            while (*p && !isspace((unsigned char) *p)) ++p;
            if (!*p || sscanf(p + 1, "%" SCNd32, &end) != 1 || end == -1) { fprintf(stderr, "Line %" PRId32 " of %s is malformed.\n", lnum, fn); return false; }
            if (start >= end) { fprintf(stderr, "The position on line %" PRId32 " of %s is incorrect (%" PRId32 " >= %" PRId32 ".\n", lnum, fn, start, end); return false; }
            if (start < 0) start = 0;
            if ((int64_t) end > (int64_t) lens[(size_t) tid] + 1) end = (int32_t)(lens[(size_t) tid] + 1);
            md_bed_region r; r.start = (uint32_t) start; r.end = (uint32_t) end; r.strand = 0;
            if (keepStrand) {                                  // name, score, then the strand column (bed.c:207-229)
                ++p;                                           // the separator after the start column
                while (*p && !isspace((unsigned char) *p)) ++p;                              // rest of the end column
                for (int col = 0; col < 2 && *p; ++col) { while (*p && isspace((unsigned char) *p)) ++p; while (*p && !isspace((unsigned char) *p)) ++p; }
                while (*p && isspace((unsigned char) *p)) ++p;
                if (*p == '+') r.strand = 1; else if (*p == '-') r.strand = 2;
            }
            by_tid[(size_t) tid].push_back(r); ++n;
        }
        for (auto &v : by_tid) std::sort(v.begin(), v.end(), [](const md_bed_region &a, const md_bed_region &b) {
            if (a.start != b.start) return a.start < b.start;
            if (a.end != b.end) return a.end < b.end;
            return a.strand < b.strand; });
        fprintf(stderr, "Parsed %zu regions in %s\n", n, fn);
        return true;
    }
    // the chunk [beg,end) of contig tid overlaps a region (compareRegions == 0 on region [start, end-1], bed.c:10-16)
    bool chunk_overlaps(uint32_t tid, uint32_t beg, uint32_t end) const {
        if (tid >= by_tid.size()) return false;
        for (const md_bed_region &r : by_tid[tid]) {
            if (r.start >= end) break;
            if ((r.start < beg && r.end - 1 >= beg) || (r.start >= beg && r.start < end)) return true;
        }
        return false;
    }
    const std::vector<md_bed_region> &regions(uint32_t tid) const { static const std::vector<md_bed_region> none; return tid < by_tid.size() ? by_tid[tid] : none; }
};
