// BAM record helpers shared by the device decode kernels (bamdev.cu) and the host self-test: record-start discovery in an
// inflated byte stream and record -> structure-of-arrays conversion.  Layout per the SAM specification section 4.2; the
// conversion follows host/tiles.hpp:SoaTile::add (which the CLI's host decode path uses) field for field, so that a tile
// decoded on the device is identical to the one the host would have built.
#pragma once
#include <stdint.h>
#include "inflate_hd.h"   // MD_HD

namespace mdbam {

static const uint64_t NONE = ~0ull;

MD_HD uint32_t ld16(const uint8_t *p) { return (uint32_t) p[0] | ((uint32_t) p[1] << 8); }
MD_HD uint32_t ld32(const uint8_t *p) { return (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24); }

// Does a chain of up to `depth` well-formed record headers start at byte o of u[0..U)?  (block_size sane, contig ids in
// range, name NUL-terminated, fixed part + name + cigar + seq + qual fit the block.)  A chain that runs off the end of the
// data after at least one full header counts.  Used only to GUESS; the caller verifies against the true chain.
MD_HD bool plausible(const uint8_t *u, uint64_t o, uint64_t U, int32_t n_targets, int depth) {
    for (int d = 0; d < depth; ++d) {
        if (o + 36 > U) return d > 0;
        const uint32_t bs = ld32(u + o);
        if (bs < 32 || bs > (1u << 28)) return false;
        const int32_t tid = (int32_t) ld32(u + o + 4), pos = (int32_t) ld32(u + o + 8);
        const uint32_t lq = u[o + 12], ncig = ld16(u + o + 16); const int32_t lseq = (int32_t) ld32(u + o + 20);
        const int32_t mtid = (int32_t) ld32(u + o + 24);
        if (tid < -1 || tid >= n_targets || mtid < -1 || mtid >= n_targets || pos < -1 || lq < 1 || lseq < 0) return false;
        const uint64_t need = 32ull + lq + 4ull * ncig + ((uint64_t) lseq + 1) / 2 + (uint64_t) lseq;
        if (need > bs) return false;
        if (o + 36 + lq <= U && u[o + 36 + lq - 1] != 0) return false;
        o += 4ull + bs;
        if (o == U) return true;
    }
    return true;
}

// first plausible record start in [lo, hi), or NONE
MD_HD uint64_t guess_start(const uint8_t *u, uint64_t lo, uint64_t hi, uint64_t U, int32_t n_targets) {
    for (uint64_t o = lo; o < hi && o + 36 <= U; ++o) if (plausible(u, o, U, n_targets, 4)) return o;
    return NONE;
}

// Walk the record chain from `s` while records START below `hi`; only whole records (within U) count.  Returns the
// number of records and, in *exit, the offset where the chain left off (start of the first record not counted).
// If `offs` is non-null the record starts are written to offs[0..count).
MD_HD uint32_t walk_chain(const uint8_t *u, uint64_t s, uint64_t hi, uint64_t U, uint64_t *exit, uint64_t *offs) {
    uint32_t n = 0; uint64_t o = s;
    while (o < hi && o + 4 <= U) {
        const uint64_t len = 4ull + ld32(u + o);
        if (o + len > U) break;
        if (offs) offs[n] = o;
        ++n; o += len;
    }
    *exit = o;
    return n;
}

struct Head { int32_t tid, pos; uint32_t l_qname, mapq, n_cigar, flag, l_seq, block_size; };
MD_HD Head head_of(const uint8_t *r) {          // r points at the 4-byte block_size
    Head h; h.block_size = ld32(r); h.tid = (int32_t) ld32(r + 4); h.pos = (int32_t) ld32(r + 8);
    h.l_qname = r[12]; h.mapq = r[13]; h.n_cigar = ld16(r + 16); h.flag = ld16(r + 18); h.l_seq = ld32(r + 20);
    return h;
}
// well-formedness as host/hostio.hpp:parse_bam_record demands
MD_HD bool head_ok(const Head &h) {
    if (h.block_size < 32 || (int32_t) h.l_seq < 0) return false;
    const uint64_t need = 32ull + h.l_qname + 4ull * h.n_cigar + ((uint64_t) h.l_seq + 1) / 2 + (uint64_t) h.l_seq;
    return need <= h.block_size;
}
MD_HD const uint8_t *qname_of(const uint8_t *r) { return r + 36; }
MD_HD const uint8_t *cigar_of(const uint8_t *r, const Head &h) { return r + 36 + h.l_qname; }
MD_HD const uint8_t *seq_of(const uint8_t *r, const Head &h) { return cigar_of(r, h) + 4ull * h.n_cigar; }
MD_HD const uint8_t *qual_of(const uint8_t *r, const Head &h) { return seq_of(r, h) + ((uint64_t) h.l_seq + 1) / 2; }
MD_HD const uint8_t *aux_of(const uint8_t *r, const Head &h) { return qual_of(r, h) + h.l_seq; }
MD_HD const uint8_t *end_of(const uint8_t *r, const Head &h) { return r + 4ull + h.block_size; }

// reference end = pos + sum of M/D/N/=/X lengths (tiles.hpp: rend)
MD_HD int32_t ref_end(const uint8_t *r, const Head &h) {
    const uint8_t *c = cigar_of(r, h); int32_t rl = 0;
    for (uint32_t k = 0; k < h.n_cigar; ++k) { const uint32_t w = ld32(c + 4 * k), op = w & 15u; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rl += (int32_t)(w >> 4); }
    return h.pos + rl;
}

// aux scan with the bam_aux_get contract (hostio.hpp:aux_find): pointer to the type byte of tag t0t1, or null
MD_HD const uint8_t *aux_find(const uint8_t *s, const uint8_t *end, uint8_t t0, uint8_t t1) {
    while (s + 3 <= end) {
        const bool match = s[0] == t0 && s[1] == t1;
        const uint8_t *tp = s + 2; const int t = *tp; const uint8_t *v = tp + 1;
        int sz;
        switch (t) { case 'A': case 'c': case 'C': sz = 1; break; case 's': case 'S': sz = 2; break; case 'i': case 'I': case 'f': sz = 4; break; case 'd': sz = 8; break; default: sz = 0; }
        if (t == 'Z' || t == 'H') { const uint8_t *q = v; while (q < end && *q) ++q; if (q >= end) return nullptr; if (match) return tp; s = q + 1; }
        else if (t == 'B') {
            if (v + 5 > end) return nullptr;
            int es; switch (v[0]) { case 'c': case 'C': es = 1; break; case 's': case 'S': es = 2; break; case 'i': case 'I': case 'f': es = 4; break; default: es = 0; }
            if (!es) return nullptr;
            const uint32_t n = ld32(v + 1);
            if (match) return tp;
            s = v + 5 + (uint64_t) es * n;
        } else { if (!sz || v + sz > end) return nullptr; if (match) return tp; s = v + sz; }
    }
    return nullptr;
}
MD_HD int64_t aux_to_int(const uint8_t *tp) {
    const uint8_t *s = tp + 1;
    switch (*tp) {
        case 'c': return (int8_t) s[0]; case 'C': return s[0];
        case 's': return (int16_t) ld16(s); case 'S': return (uint16_t) ld16(s);
        case 'i': return (int32_t) ld32(s); case 'I': return (uint32_t) ld32(s);
        default: return 0;
    }
}
// md_reads_soa::aux: bit0 XG[0]=='C', bit1 XG[0]=='G' (getStrand, common.c:85-87), bit2 NH > 1 (filter_func, common.c:421-427)
MD_HD uint8_t aux_bits(const uint8_t *r, const Head &h) {
    const uint8_t *a = aux_of(r, h), *e = end_of(r, h); uint8_t v = 0;
    if (const uint8_t *xg = aux_find(a, e, 'X', 'G')) { if (xg[1] == 'C') v |= 1; else if (xg[1] == 'G') v |= 2; }
    if (const uint8_t *nh = aux_find(a, e, 'N', 'H')) { if (aux_to_int(nh) > 1) v |= 4; }
    return v;
}
// pairing key: hostio.hpp:qname_key over the name without its terminator (strnlen semantics)
MD_HD uint64_t name_key(const uint8_t *r, const Head &h) {
    const uint8_t *s = qname_of(r); uint64_t n = 0;
    while (n < h.l_qname && s[n]) ++n;
    uint64_t k = 0xcbf29ce484222325ull ^ (n * 0x9e3779b97f4a7c15ull);
    for (uint64_t i = 0; i < n; ++i) { k ^= s[i]; k *= 0x100000001b3ull; }
    k ^= k >> 30; k *= 0xbf58476d1ce4e5b9ull; k ^= k >> 27; k *= 0x94d049bb133111ebull; k ^= k >> 31;
    if (k == 0) k = 1;
    return k;
}
// A second, independent 32-bit hash of the same name (different multiplier, bytes taken back to front).  Two names are taken
// for the same fragment only when both the 64-bit key and this check agree: 96 bits instead of 64.
MD_HD uint32_t name_check(const uint8_t *r, const Head &h) {
    const uint8_t *s = qname_of(r); uint32_t n = 0;
    while (n < h.l_qname && s[n]) ++n;
    uint32_t k = 0x811c9dc5u ^ (n * 0x85ebca6bu);
    for (uint32_t i = n; i-- > 0;) { k = (k ^ s[i]) * 0x01000193u; k ^= k >> 15; }
    k ^= k >> 16; k *= 0x7feb352du; k ^= k >> 15; k *= 0x846ca68bu; k ^= k >> 16;
    return k;
}
MD_HD uint32_t seq_words(uint32_t l) { return (((l + 1u) >> 1) + 3u) >> 2; }
MD_HD uint32_t qual_words8(uint32_t l) { return (l + 7u) >> 3; }

// copy n bytes into 32-bit / 64-bit words (little-endian byte order within the word, zero padded) — what memcpy into a
// zeroed word array does on the host
MD_HD void bytes_to_words32(uint32_t *dst, const uint8_t *src, uint32_t n) {
    const uint32_t w = (n + 3u) >> 2;
    for (uint32_t k = 0; k < w; ++k) {
        uint32_t v = 0; const uint32_t b0 = 4u * k;
        for (uint32_t j = 0; j < 4u && b0 + j < n; ++j) v |= (uint32_t) src[b0 + j] << (8u * j);
        dst[k] = v;
    }
}
MD_HD void bytes_to_words64(uint64_t *dst, const uint8_t *src, uint32_t n) {
    const uint32_t w = (n + 7u) >> 3;
    for (uint32_t k = 0; k < w; ++k) {
        uint64_t v = 0; const uint32_t b0 = 8u * k;
        for (uint32_t j = 0; j < 8u && b0 + j < n; ++j) v |= (uint64_t) src[b0 + j] << (8u * j);
        dst[k] = v;
    }
}

}  // namespace mdbam
