// DEFLATE (RFC 1951) decoder for one BGZF block, written once for host and device.
//
// Replaces, on the device, the zlib inflate that runs inside htslib's bgzf_read under sam_itr_next
// (reference call site common.c:413; BGZF framing per the SAM specification, section 4.1).  A BGZF block is an
// independent raw-deflate stream of at most 64 KB of output, so blocks decode in parallel: the CUDA kernel gives every
// warp one block and lets its leader lane run inflate_block() below; the very same function is compiled for the host
// by tests (checked against zlib on real files), which is what keeps a decoder that cannot be debugged interactively
// on the GPU honest.
//
// Decoding tables (per block decoder, ~3.6 KB; shared memory on the device): a 10-bit primary table for literal/length
// codes and an 8-bit one for distance codes resolve every code up to that length in one look-up; longer codes (rare by
// construction: they are the improbable symbols) fall back to the canonical count/offset walk.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MD_HD __host__ __device__ __forceinline__
#else
#define MD_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define MD_SYNCWARP() __syncwarp()
#else
#define MD_SYNCWARP() ((void) 0)
#endif

namespace mdinflate {

enum { LIT_ROOT = 10, DIST_ROOT = 8 };

struct Tables {
    uint16_t lit[1 << LIT_ROOT];      // (symbol << 4) | code length, 0 = code longer than LIT_ROOT
    uint16_t dist[1 << DIST_ROOT];    // same for distance codes
    uint16_t lit_sorted[288], dist_sorted[32];   // symbols ordered by (length, symbol): canonical walk
    uint16_t lit_count[16], dist_count[16];
};

struct BitReader {
    const uint32_t *words;   // 4-byte aligned base of the buffer the stream lives in (padded by >= 8 readable bytes)
    uint64_t next;           // index of the next word to go into the bit buffer
    uint64_t buf; int cnt;   // LSB-first bit buffer
    uint64_t end_word;       // first word index wholly beyond the stream (reads past it yield zeros)
    uint32_t pre0, pre1;     // words[next], words[next + 1], loaded ahead of their use: on the device the compressed stream comes
                             // from L2/HBM, and a load issued only when the buffer runs dry would sit on the decoder's critical path
};

MD_HD uint32_t br_word(const BitReader &b, uint64_t i) { return i < b.end_word ? b.words[i] : 0u; }
MD_HD void br_init(BitReader &b, const void *base_aligned, uint64_t byte_off, uint64_t byte_len) {
    b.words = (const uint32_t *) base_aligned;
    b.next = byte_off >> 2;
    b.end_word = (byte_off + byte_len + 3) >> 2;
    b.buf = 0; b.cnt = 0;
    const int mis = (int)(byte_off & 3);
    if (byte_len) { b.buf = (uint64_t)(b.words[b.next++] >> (8 * mis)); b.cnt = 32 - 8 * mis; }
    b.pre0 = br_word(b, b.next); b.pre1 = br_word(b, b.next + 1);
}
// at least 33 bits available afterwards (zeros once the stream is exhausted)
MD_HD void br_fill(BitReader &b) {
    if (b.cnt <= 32) {
        b.buf |= (uint64_t) b.pre0 << b.cnt; b.cnt += 32;
        ++b.next;
        b.pre0 = b.pre1; b.pre1 = br_word(b, b.next + 1);
    }
}
MD_HD uint32_t br_peek(const BitReader &b, int n) { return (uint32_t)(b.buf & ((1ull << n) - 1ull)); }
MD_HD void br_drop(BitReader &b, int n) { b.buf >>= n; b.cnt -= n; }
MD_HD uint32_t br_take(BitReader &b, int n) { uint32_t v = br_peek(b, n); br_drop(b, n); return v; }
// bytes of the stream consumed so far (for the stored-block path and for diagnostics)
MD_HD void br_align_byte(BitReader &b) { br_drop(b, b.cnt & 7); }

MD_HD uint32_t bit_reverse(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}

// Canonical Huffman set from code lengths (RFC 1951 3.2.2).  Returns false for an over-subscribed set.
// All lanes of a warp run this with the same arguments; only lane 0 stores to the (shared) tables, with a warp barrier
// between the phases, so no lane reads a table entry another lane is still writing.
MD_HD bool build_table(const uint8_t *lens, int n, uint16_t *primary, int root, uint16_t *sorted, uint16_t *count, int lane) {
    uint16_t cnt[16];
    for (int i = 0; i < 16; ++i) cnt[i] = 0;
    for (int i = 0; i < n; ++i) cnt[lens[i]]++;
    cnt[0] = 0;
    int left = 1;
    for (int l = 1; l < 16; ++l) { left <<= 1; left -= cnt[l]; if (left < 0) return false; }
    uint16_t offs[16]; offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + cnt[l]);
    MD_SYNCWARP();                                       // nobody is still decoding with the previous tables
    if (lane == 0) {
        for (int i = 0; i < 16; ++i) count[i] = cnt[i];
        for (int i = 0; i < n; ++i) if (lens[i]) sorted[offs[lens[i]]++] = (uint16_t) i;
        if (root) for (int i = 0; i < (1 << root); ++i) primary[i] = 0;
    }
    MD_SYNCWARP();
    // assign codes in (length, symbol) order and spread the short ones over the primary table
    if (root && lane == 0) {
        uint32_t code = 0; int idx = 0;
        for (int l = 1; l <= root; ++l) {
            for (int k = 0; k < cnt[l]; ++k, ++idx, ++code) {
                const uint32_t rev = bit_reverse(code, l);
                const uint16_t e = (uint16_t)((sorted[idx] << 4) | l);
                for (uint32_t x = rev; x < (1u << root); x += (1u << l)) primary[x] = e;
            }
            code <<= 1;
        }
    }
    MD_SYNCWARP();
    return true;
}

// canonical walk, one bit at a time (codes longer than the primary table, and the code-length alphabet)
MD_HD int decode_slow(BitReader &b, const uint16_t *sorted, const uint16_t *count) {
    int code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; ++l) {
        code |= (int) br_take(b, 1);
        const int c = count[l];
        if (code - c < first) return sorted[index + (code - first)];
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return -1;
}

MD_HD int decode_sym(BitReader &b, const uint16_t *primary, int root, const uint16_t *sorted, const uint16_t *count) {
    const uint16_t e = primary[br_peek(b, root)];
    if (e) { br_drop(b, e & 15); return e >> 4; }
    return decode_slow(b, sorted, count);
}

// Inflate one raw-deflate stream of `in_len` bytes at base+in_off into out[0..out_len).  The stream must produce exactly
// out_len bytes.  Returns 0, or a negative code: -1 bad block type / stored length, -2 bad code lengths, -3 bad symbol,
// -4 output overrun, -5 distance before start, -6 output short.
//
// `lane` / `nl`: on the device all 32 lanes of a warp run this function REDUNDANTLY on the same block (identical control
// flow and register state, so the warp never diverges and the cost is that of one lane); what they share out is the
// output: lane 0 stores literals, and a match of `len` bytes is copied by all lanes at once (one round of memory latency
// per 32 bytes instead of one per byte — the byte-serial copy is what bounds a single-lane decoder).  Host: lane 0 of 1.
MD_HD int inflate_block(const void *base_aligned, uint64_t in_off, uint64_t in_len, uint8_t *out, uint32_t out_len, Tables &T, int lane = 0, int nl = 1) {
    BitReader b; br_init(b, base_aligned, in_off, in_len);
    uint32_t op = 0;
    const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    const uint8_t cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    for (;;) {
        br_fill(b);
        const uint32_t last = br_take(b, 1), type = br_take(b, 2);
        if (type == 0) {
            br_align_byte(b); br_fill(b);
            const uint32_t len = br_take(b, 16); br_fill(b);
            const uint32_t nlen = br_take(b, 16);
            if ((len ^ 0xffffu) != nlen) return -1;
            if (op + len > out_len) return -4;
            for (uint32_t i = 0; i < len; ++i) { br_fill(b); const uint8_t v = (uint8_t) br_take(b, 8); if (lane == 0) out[op] = v; ++op; }
        } else if (type == 1 || type == 2) {
            uint8_t lens[320];
            int nlit, ndist;
            if (type == 1) {
                nlit = 288; ndist = 30;
                for (int i = 0; i < 144; ++i) lens[i] = 8;
                for (int i = 144; i < 256; ++i) lens[i] = 9;
                for (int i = 256; i < 280; ++i) lens[i] = 7;
                for (int i = 280; i < 288; ++i) lens[i] = 8;
                for (int i = 0; i < 30; ++i) lens[288 + i] = 5;
            } else {
                nlit = (int) br_take(b, 5) + 257; ndist = (int) br_take(b, 5) + 1;
                const int ncl = (int) br_take(b, 4) + 4;
                if (nlit > 286 || ndist > 30) return -2;
                uint8_t cl[19];
                for (int i = 0; i < 19; ++i) cl[i] = 0;
                for (int i = 0; i < ncl; ++i) { br_fill(b); cl[cl_order[i]] = (uint8_t) br_take(b, 3); }
                // the code-length alphabet is decoded with the canonical walk only (19 symbols, <= 7 bits; root 0 = no primary
                // table); T.dist_sorted / T.dist_count serve as its scratch until the real distance set is built
                if (!build_table(cl, 19, T.dist, 0, T.dist_sorted, T.dist_count, lane)) return -2;
                int i = 0;
                while (i < nlit + ndist) {
                    br_fill(b);
                    const int sym = decode_slow(b, T.dist_sorted, T.dist_count);
                    if (sym < 0) return -2;
                    if (sym < 16) lens[i++] = (uint8_t) sym;
                    else {
                        int rep; uint8_t v = 0;
                        br_fill(b);
                        if (sym == 16) { if (i == 0) return -2; v = lens[i - 1]; rep = 3 + (int) br_take(b, 2); }
                        else if (sym == 17) rep = 3 + (int) br_take(b, 3);
                        else rep = 11 + (int) br_take(b, 7);
                        if (i + rep > nlit + ndist) return -2;
                        while (rep--) lens[i++] = v;
                    }
                }
                if (lens[256] == 0) return -2;
                // distance lengths follow the literal/length ones; move them to a fixed offset
                for (int k = ndist - 1; k >= 0; --k) lens[288 + k] = lens[nlit + k];
            }
            if (!build_table(lens, nlit, T.lit, LIT_ROOT, T.lit_sorted, T.lit_count, lane)) return -2;
            if (!build_table(lens + 288, ndist, T.dist, DIST_ROOT, T.dist_sorted, T.dist_count, lane)) return -2;
            for (;;) {
                br_fill(b);
                int sym = decode_sym(b, T.lit, LIT_ROOT, T.lit_sorted, T.lit_count);
                if (sym < 0) return -3;
                if (sym < 256) { if (op >= out_len) return -4; if (lane == 0) out[op] = (uint8_t) sym; ++op; continue; }
                if (sym == 256) break;
                sym -= 257;
                if (sym >= 29) return -3;
                uint32_t len = len_base[sym] + br_take(b, len_extra[sym]);
                br_fill(b);
                const int ds = decode_sym(b, T.dist, DIST_ROOT, T.dist_sorted, T.dist_count);
                if (ds < 0 || ds >= 30) return -3;
                const uint32_t dist = dist_base[ds] + br_take(b, dist_extra[ds]);
                if (dist > op) return -5;
                if (op + len > out_len) return -4;
                const uint8_t *src = out + op - dist; uint8_t *dst = out + op;
                MD_SYNCWARP();                                   // the bytes being copied were stored by other lanes
                if (dist >= len) { for (uint32_t k = (uint32_t) lane; k < len; k += (uint32_t) nl) dst[k] = src[k]; }
                else if (nl == 1) { for (uint32_t k = 0; k < len; ++k) dst[k] = src[k]; }
                else { for (uint32_t k = (uint32_t) lane; k < len; k += (uint32_t) nl) dst[k] = src[k % dist]; }   // overlapping match = the last `dist` bytes repeated
                op += len;
            }
        } else return -1;
        if (last) break;
    }
    return op == out_len ? 0 : -6;
}

}  // namespace mdinflate
