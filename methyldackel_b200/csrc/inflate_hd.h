// DEFLATE (RFC 1951) decoder for one BGZF block, written once for host and device.
//
// Replaces, on the device, the zlib inflate that runs inside htslib's bgzf_read under sam_itr_next
// (reference call site common.c:413; BGZF framing per the SAM specification, section 4.1).  A BGZF block is an
// independent raw-deflate stream of at most 64 KB of output, so blocks decode in parallel: the CUDA kernel gives every
// warp one block at a time.  The very same function is compiled for the host by tests (lane 0 of 1, checked against zlib
// on real files), which is what keeps a decoder that cannot be debugged interactively on the GPU honest.
//
// How a warp decodes (round 2 design; the round-1 decoder copied every match through global memory, one L2 round trip per
// match, and BAM data is 80 % matches of ~6 bytes):
//   * The bit-stream walk is serial, so all 32 lanes run it REDUNDANTLY — identical control flow and register state, the warp
//     never diverges and the walk costs what one lane would cost.  What the lanes share out is the data movement.
//   * Output goes into a ring of the last WIN bytes in SHARED memory.  A match whose source lies inside the ring (distance
//     <= WIN - 264) is a shared-memory load + store by `len`
//     lanes at once; only the far ones read global memory, from bytes that were flushed long before.
//   * Every FLUSH bytes the finished part of the ring is written to global memory by all lanes with 16-byte stores.
//   * Decoding tables hold 32-bit entries with everything a symbol needs (code length, extra-bit count, base value, kind),
//     so a match costs two table look-ups and no further memory access; codes longer than the primary table (rare by
//     construction: they are the improbable symbols) fall back to the canonical count/offset walk.
//   * The tables are built by the warp in parallel: lane l owns the codes of length l + 1.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MD_HD __host__ __device__ __forceinline__
#define MD_HD_COLD __host__ __device__ __noinline__      /* rare paths: kept out of the hot loop's instruction stream */
#else
#define MD_HD inline
#define MD_HD_COLD inline
#endif
#if defined(__CUDA_ARCH__)
#define MD_SYNCWARP() __syncwarp()
#define MD_ATOMIC_INC(p) atomicAdd((p), 1u)
#else
#define MD_SYNCWARP() ((void) 0)
#define MD_ATOMIC_INC(p) (++*(p))
#endif

namespace mdinflate {

#ifndef MD_INFLATE_WIN
#define MD_INFLATE_WIN 2048
#endif
// Sizes are chosen for occupancy: the walk is one long dependency chain (measured: ~7 cycles from one issued instruction of
// a warp to its next), so throughput grows with the number of resident warps until the schedulers saturate at ~7 warps each.
// 8-bit / 6-bit primary tables cover 94 % / 97 % of the codes of a BAM stream; a 2 KB ring holds the source of 3 matches in 4
// (measured on the config[1] file: 32 warps per SM with a 2 KB ring beat 28 with 4 KB by 6 %, and 15 with 8 KB by 13 %).
enum { LIT_ROOT = 8, DIST_ROOT = 6, CL_ROOT = 7, WIN = MD_INFLATE_WIN, WMASK = WIN - 1, FLUSH = WIN / 4, NEAR = WIN - 264,
       IN_CHUNK = 64, IN_WORDS = 2 * IN_CHUNK };   // device: the compressed stream is staged through a ring of two 256-byte chunks

// entry layout (lit/len and distance tables): bits 0-3 code length (0: not in the primary table), bits 4-7 number of extra
// bits, bits 8-9 kind (0 literal / distance, 1 length, 2 end of block, 3 invalid symbol: never stored in a primary table, so the
// hot loop does not test for it), bits 16-31 literal byte or base value
struct Decoder {                         // one per warp; shared memory on the device (5.25 KB with a 2 KB ring)
    uint32_t lit[1 << LIT_ROOT];
    uint32_t dist[1 << CL_ROOT];         // distance table (1 << DIST_ROOT entries used); holds the code-length alphabet's table while a dynamic header is read
    uint16_t lit_sorted[288], dist_sorted[32];   // symbols ordered by (length, symbol): canonical walk for codes beyond the primary table
    uint32_t lit_limit[16], dist_limit[16];      // per code length l: (first code of length l + number of codes of length l) << (15 - l)
    int32_t lit_off[16], dist_off[16];           // per code length l: index into sorted[] of its first symbol - its first code
    uint32_t scratch[16];                // per-length counters while a table is built
    uint8_t lens[320];                   // code lengths of the block being set up (lit/len at 0, distance at 288)
    alignas(16) uint32_t in[IN_WORDS];   // device only: word i of the compressed buffer sits at in[i % IN_WORDS] while the reader is near it
    alignas(16) uint8_t win[WIN];        // ring of the most recent output, indexed by (output address & WMASK)
};

// The compressed stream is read 32 bits at a time.  On the device a load issued when the word is needed — or kept in a
// register queue that is shifted at every refill, which makes each refill wait for the load of the previous one — puts an
// L2/HBM round trip on the critical path of a walk that is nothing but a dependency chain.  So the warp stages the stream
// through shared memory: 256-byte chunks copied with cp.async (LDGSTS, no registers involved) one chunk ahead of the reader,
// and a refill is a shared-memory load of a word that arrived hundreds of symbols earlier.
struct BitReader {
    const uint32_t *words;   // 4-byte aligned base of the buffer the stream lives in (256-byte aligned on the device)
    uint32_t w0, w1;         // the 64-bit window the next bits come from: bit p of w0 is the next one
    uint32_t pre0;           // the word after w1, fetched one refill ahead of its use
    uint32_t p;              // < 32
    uint32_t idx;            // index of the word in w0 (a pushed segment is far below 16 GB)
    uint32_t end_word;       // first word index wholly beyond the stream (reads past it yield zeros)
    uint32_t *ring;          // device: Decoder::in
    uint32_t nci;            // device: next chunk to be copied into the ring
    int lane;
};

MD_HD uint32_t br_word(const BitReader &b, uint32_t i) {
#if defined(__CUDA_ARCH__)
    return i < b.end_word ? __ldg(b.words + i) : 0u;
#else
    return i < b.end_word ? b.words[i] : 0u;
#endif
}
#if defined(__CUDA_ARCH__)
// chunk c of the buffer -> its half of the ring (lanes 0-15, 16 bytes each); chunks wholly beyond the stream are zeros
__device__ __forceinline__ void br_issue_chunk(BitReader &b, uint32_t c) {
    if (b.lane < 16) {
        uint32_t *dst = b.ring + (c & 1u) * IN_CHUNK + (uint32_t) b.lane * 4u;
        const uint32_t w = c * IN_CHUNK + (uint32_t) b.lane * 4u;
        if (c * IN_CHUNK < b.end_word + IN_CHUNK) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t) __cvta_generic_to_shared(dst)), "l"(b.words + w) : "memory");
        } else { dst[0] = 0u; dst[1] = 0u; dst[2] = 0u; dst[3] = 0u; }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void br_wait_chunks() { asm volatile("cp.async.wait_group 0;" ::: "memory"); __syncwarp(); }
#endif
MD_HD void br_init(BitReader &b, const void *base_aligned, uint64_t byte_off, uint64_t end_byte, uint32_t *ring, int lane) {
    b.words = (const uint32_t *) base_aligned;
    b.end_word = (uint32_t)((end_byte + 3) >> 2);
    b.idx = (uint32_t)(byte_off >> 2); b.p = (uint32_t)(byte_off & 3) * 8u;
    b.w0 = br_word(b, b.idx); b.w1 = br_word(b, b.idx + 1); b.pre0 = br_word(b, b.idx + 2);
    b.ring = ring; b.lane = lane; b.nci = 0;
#if defined(__CUDA_ARCH__)
    const uint32_t c0 = (b.idx + 3u) / IN_CHUNK;         // the chunk the next fetch falls into
    __syncwarp();                                        // nobody still reads the ring on behalf of an earlier stream
    br_issue_chunk(b, c0); br_issue_chunk(b, c0 + 1u);
    b.nci = c0 + 2u;
    br_wait_chunks();
#endif
}
#if defined(__CUDA_ARCH__)
// entering a new chunk: it was requested a whole chunk ago; the chunk before it is used up, its half of the ring takes the next one
__device__ __noinline__ uint32_t br_next_chunk(const uint32_t *words, uint32_t *ring, uint32_t end_word, uint32_t nci, int lane) {
    BitReader t; t.words = words; t.ring = ring; t.end_word = end_word; t.lane = lane;
    br_wait_chunks();
    br_issue_chunk(t, nci);
    return nci + 1u;
}
#endif
// word i of the stream for the refill
MD_HD uint32_t br_fetch(BitReader &b, uint32_t i) {
#if defined(__CUDA_ARCH__)
    if ((i & (IN_CHUNK - 1u)) == 0u && i / IN_CHUNK + 1u == b.nci) b.nci = br_next_chunk(b.words, b.ring, b.end_word, b.nci, b.lane);
    return b.ring[i & (IN_WORDS - 1u)];
#else
    return br_word(b, i);
#endif
}
// the next 32 bits of the stream
MD_HD uint32_t br_peek(const BitReader &b) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(b.w0, b.w1, b.p);
#else
    return (uint32_t)((((uint64_t) b.w1 << 32) | (uint64_t) b.w0) >> b.p);
#endif
}
MD_HD void br_drop(BitReader &b, uint32_t n) {      // n <= 32
    b.p += n;
    if (b.p >= 32u) {
        b.p -= 32u; ++b.idx;
        b.w0 = b.w1; b.w1 = b.pre0; b.pre0 = br_fetch(b, b.idx + 2u);
    }
}
MD_HD uint32_t br_take(BitReader &b, uint32_t n) { const uint32_t v = br_peek(b) & ((1u << n) - 1u); br_drop(b, n); return v; }   // n < 32

MD_HD uint32_t bit_reverse(uint32_t v, int n) {
#if defined(__CUDA_ARCH__)
    return __brev(v) >> (32 - n);
#else
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
#endif
}

// table entries from (symbol, code length); base values and extra-bit counts of RFC 1951 3.2.5 in closed form
MD_HD uint32_t lit_entry(uint32_t sym, uint32_t len) {
    if (sym < 256u) return (sym << 16) | len;
    if (sym == 256u) return (2u << 8) | len;
    if (sym > 285u) return (3u << 8) | len;
    const uint32_t s = sym - 257u;
    uint32_t base, xb;
    if (s < 8u) { base = 3u + s; xb = 0u; }
    else if (s == 28u) { base = 258u; xb = 0u; }
    else { xb = (s >> 2) - 1u; base = ((4u + (s & 3u)) << xb) + 3u; }
    return (base << 16) | (1u << 8) | (xb << 4) | len;
}
MD_HD uint32_t dist_entry(uint32_t sym, uint32_t len) {
    if (sym >= 30u) return (3u << 8) | len;
    uint32_t base, xb;
    if (sym < 4u) { base = 1u + sym; xb = 0u; }
    else { xb = (sym >> 1) - 1u; base = ((2u + (sym & 1u)) << xb) + 1u; }
    return (base << 16) | (xb << 4) | len;
}

// Canonical Huffman set from code lengths (RFC 1951 3.2.2).  Returns false for an over-subscribed set.
// kind: 0 lit/len, 1 distance, 2 code-length alphabet.  The warp builds it together: lane l handles the codes of length
// l + 1 (on the host the single lane loops over the lengths).
MD_HD bool build_table(Decoder &D, const uint8_t *lens, int n, uint32_t *primary, int root, uint16_t *sorted, uint32_t *limit, int32_t *offv, int kind, int lane, int nl) {
    MD_SYNCWARP();                                       // nobody is still decoding with the previous tables
    for (int i = lane; i < 16; i += nl) D.scratch[i] = 0;
    for (int i = lane; i < (1 << root); i += nl) primary[i] = 0;
    MD_SYNCWARP();
    for (int i = lane; i < n; i += nl) { const uint32_t l = lens[i]; if (l) MD_ATOMIC_INC(&D.scratch[l]); }
    MD_SYNCWARP();
    int left = 1;
    for (int l = 1; l < 16; ++l) { left <<= 1; left -= (int) D.scratch[l]; if (left < 0) return false; }
    for (int l = 1 + lane; l < 16; l += nl) {
        uint32_t off = 0, code = 0;
        for (int k = 1; k < l; ++k) { const uint32_t c = D.scratch[k]; off += c; code = (code + c) << 1; }
        const uint32_t c = D.scratch[l];
        limit[l] = (code + c) << (15 - l);
        offv[l] = (int32_t) off - (int32_t) code;
        if (!c) continue;
        for (int i = 0; i < n; ++i) {
            if (lens[i] != (uint8_t) l) continue;
            sorted[off++] = (uint16_t) i;
            if (l <= root) {
                const uint32_t e = kind == 0 ? lit_entry((uint32_t) i, (uint32_t) l) : kind == 1 ? dist_entry((uint32_t) i, (uint32_t) l) : (((uint32_t) i << 16) | (uint32_t) l);
                if ((e & 0x300u) != 0x300u)              // an invalid symbol is left to the slow path, which reports it
                    for (uint32_t x = bit_reverse(code, l); x < (1u << root); x += (1u << l)) primary[x] = e;
            }
            ++code;
        }
    }
    MD_SYNCWARP();
    return true;
}

// A code longer than the primary table: canonical decode by limits.  With the next 15 bits taken MSB first (x), the code's
// length is the first l with x < limit[l]; its symbol is sorted[off[l] + (x >> (15 - l))].  `bits` = the next bits of the stream;
// returns symbol | code length << 16 (the caller drops the bits), or -1.
MD_HD_COLD int decode_slow(uint32_t bits, const uint16_t *sorted, const uint32_t *limit, const int32_t *offv, int root) {
    const uint32_t x = bit_reverse(bits & 0x7fffu, 15);
#if defined(__CUDA_ARCH__)
    #pragma unroll 1
#endif
    for (int l = root + 1; l < 16; ++l) {
        if (x < limit[l]) return (int) sorted[offv[l] + (int32_t)(x >> (15 - l))] | (l << 16);
    }
    return -1;
}

// ring -> global: bytes [lo, hi) of the output (addresses relative to the 16-byte aligned `outb`)
MD_HD_COLD void flush_ring(const Decoder &D, uint8_t *outb, uint32_t lo, uint32_t hi, int lane, int nl) {
    MD_SYNCWARP();                                       // everything below `hi` has been stored
    if (hi <= lo) return;
    const uint32_t lo16 = (lo + 15u) & ~15u, hi16 = hi & ~15u;
    if (lo16 >= hi16) { for (uint32_t x = lo + (uint32_t) lane; x < hi; x += (uint32_t) nl) outb[x] = D.win[x & WMASK]; }
    else {
        for (uint32_t x = lo + (uint32_t) lane; x < lo16; x += (uint32_t) nl) outb[x] = D.win[x & WMASK];
        for (uint32_t x = lo16 + 16u * (uint32_t) lane; x < hi16; x += 16u * (uint32_t) nl) {
#if defined(__CUDA_ARCH__)
            *(uint4 *)(outb + x) = *(const uint4 *)(D.win + (x & WMASK));
#else
            for (int k = 0; k < 16; ++k) outb[x + (uint32_t) k] = D.win[(x + (uint32_t) k) & WMASK];
#endif
        }
        for (uint32_t x = hi16 + (uint32_t) lane; x < hi; x += (uint32_t) nl) outb[x] = D.win[x & WMASK];
    }
    MD_SYNCWARP();                                       // the ring bytes may be overwritten from here on; far matches may read the flushed bytes
}

// Inflate one raw-deflate stream of `in_len` bytes at base+in_off into out[0..out_len).  The stream must produce exactly
// out_len bytes.  Returns 0, or a negative code: -1 bad block type / stored length, -2 bad code lengths, -3 bad symbol,
// -4 output overrun, -5 distance before start, -6 output short.
// `lane` / `nl`: on the device all 32 lanes of a warp call this with the same arguments (see the header comment).  Host: lane 0 of 1.
MD_HD int inflate_block(const void *base_aligned, uint64_t in_off, uint64_t in_len, uint8_t *out, uint32_t out_len, Decoder &D, int lane = 0, int nl = 1) {
    BitReader b; br_init(b, base_aligned, in_off, in_off + in_len, D.in, lane);
    // output addresses are kept relative to the 16-byte aligned address at or below `out`, so that ring index, global
    // address and the 16-byte flush stores agree in alignment
    const uint32_t a0 = (uint32_t)((uintptr_t) out & 15u);
    uint8_t *outb = out - a0;
    uint32_t op = a0, flushed = a0, next_flush = FLUSH;
    bool bad = false;
    const uint32_t oend = a0 + out_len;
    const uint8_t cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    for (;;) {
        const uint32_t hdr = br_peek(b);
        const uint32_t last = hdr & 1u, type = (hdr >> 1) & 3u;
        br_drop(b, 3);
        if (type == 0) {
            br_drop(b, (8u - (b.p & 7u)) & 7u);
            const uint32_t ln = br_peek(b);
            br_drop(b, 32);
            const uint32_t len = ln & 0xffffu;
            if ((len ^ 0xffffu) != (ln >> 16)) return -1;
            if (op + len > oend) return -4;
            uint64_t bp = (uint64_t) b.idx * 4u + (b.p >> 3);               // byte position of the stored data in the buffer
            const uint8_t *src = (const uint8_t *) b.words;
            uint32_t rem = len;
            while (rem) {
                const uint32_t nmax = next_flush - op, c = rem < nmax ? rem : nmax;
                for (uint32_t k = (uint32_t) lane; k < c; k += (uint32_t) nl) D.win[(op + k) & WMASK] = src[bp + k];
                op += c; bp += c; rem -= c;
                if (op >= next_flush) { flush_ring(D, outb, flushed, next_flush, lane, nl); flushed = next_flush; next_flush += FLUSH; }
            }
            br_init(b, base_aligned, bp, in_off + in_len, D.in, lane);
        } else if (type == 1 || type == 2) {
            int nlit, ndist;
            MD_SYNCWARP();
            if (type == 1) {
                nlit = 288; ndist = 30;
                for (int i = lane; i < 320; i += nl) D.lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : i < 288 ? 8 : 5);
            } else {
                const uint32_t h = br_peek(b);
                nlit = (int)(h & 31u) + 257; ndist = (int)((h >> 5) & 31u) + 1;
                const int ncl = (int)((h >> 10) & 15u) + 4;
                br_drop(b, 14);
                if (nlit > 286 || ndist > 30) return -2;
                // the code-length alphabet (19 symbols, <= 7 bits) borrows the distance table and its canonical-walk arrays
                uint8_t *cl = D.lens + 288;
                for (int i = lane; i < 19; i += nl) cl[i] = 0;
                MD_SYNCWARP();
                for (int i = 0; i < ncl; ++i) { const uint32_t v = br_take(b, 3); if (lane == 0) cl[cl_order[i]] = (uint8_t) v; }
                MD_SYNCWARP();
                if (!build_table(D, cl, 19, D.dist, CL_ROOT, D.dist_sorted, D.dist_limit, D.dist_off, 2, lane, nl)) return -2;
                int i = 0; uint32_t prev = 0;
                while (i < nlit + ndist) {
                    const uint32_t e = D.dist[br_peek(b) & ((1u << CL_ROOT) - 1u)];
                    if (!(e & 15u)) return -2;                   // code-length codes are at most 7 bits: all of them are in the table
                    br_drop(b, e & 15u);
                    const uint32_t sym = e >> 16;
                    if (sym < 16u) { if (lane == 0) D.lens[i] = (uint8_t) sym; prev = sym; ++i; }
                    else {
                        int rep; uint32_t v = 0;
                        if (sym == 16u) { if (i == 0) return -2; v = prev; rep = 3 + (int) br_take(b, 2); }
                        else if (sym == 17u) rep = 3 + (int) br_take(b, 3);
                        else rep = 11 + (int) br_take(b, 7);
                        if (i + rep > nlit + ndist) return -2;
                        for (int k = lane; k < rep; k += nl) D.lens[i + k] = (uint8_t) v;
                        i += rep; prev = v;
                    }
                }
                MD_SYNCWARP();
                if (D.lens[256] == 0) return -2;
                // distance lengths follow the literal/length ones; move them to a fixed offset
                uint8_t dl = 0;
                if (nl == 1) { for (int k = ndist - 1; k >= 0; --k) D.lens[288 + k] = D.lens[nlit + k]; }
                else { if (lane < ndist) dl = D.lens[nlit + lane]; MD_SYNCWARP(); if (lane < ndist) D.lens[288 + lane] = dl; }
                MD_SYNCWARP();
            }
            MD_SYNCWARP();
            if (!build_table(D, D.lens, nlit, D.lit, LIT_ROOT, D.lit_sorted, D.lit_limit, D.lit_off, 0, lane, nl)) return -2;
            if (!build_table(D, D.lens + 288, ndist, D.dist, DIST_ROOT, D.dist_sorted, D.dist_limit, D.dist_off, 1, lane, nl)) return -2;
            // The hot loop.  Range violations (a distance reaching before the block, output beyond its announced size) only set
            // `bad`: the ring absorbs the stray bytes, flushes are clamped to the block, and the flag is looked at when the ring is
            // flushed and at the end — two compares per symbol instead of two branches.
            for (;;) {
                uint32_t bits = br_peek(b);
                uint32_t e = D.lit[bits & ((1u << LIT_ROOT) - 1u)];
                if (!(e & 15u)) {                                 // a code longer than the primary table
                    const int sym = decode_slow(bits, D.lit_sorted, D.lit_limit, D.lit_off, LIT_ROOT);
                    if (sym < 0) return -3;
                    br_drop(b, (uint32_t) sym >> 16);
                    e = lit_entry((uint32_t) sym & 0xffffu, 0u);
                    if ((e & 0x300u) == 0x300u) return -3;
                    bits = br_peek(b);
                }
                const uint32_t cl = e & 15u;
                if (e & 0x100u) {                                 // length symbol: a match
                    const uint32_t xb = (e >> 4) & 15u;
                    const uint32_t len = (e >> 16) + ((bits >> cl) & ~(0xffffffffu << xb));
                    uint32_t used = cl + xb;                      // <= 20: the distance code's first DIST_ROOT bits are still in `bits`
                    uint32_t d = D.dist[(bits >> used) & ((1u << DIST_ROOT) - 1u)];
                    if (!(d & 15u)) {
                        br_drop(b, used); used = 0;
                        const int ds = decode_slow(br_peek(b), D.dist_sorted, D.dist_limit, D.dist_off, DIST_ROOT);
                        if (ds < 0) return -3;
                        br_drop(b, (uint32_t) ds >> 16);
                        d = dist_entry((uint32_t) ds & 0xffffu, 0u);
                        if (d & 0x300u) return -3;
                        bits = br_peek(b);
                    }
                    const uint32_t dcl = d & 15u, dxb = (d >> 4) & 15u;
                    if (used + dcl + dxb > 32u) { br_drop(b, used); used = 0; bits = br_peek(b); }   // rare: long codes with many extra bits (dcl + dxb <= 28)
                    const uint32_t dist = (d >> 16) + ((bits >> (used + dcl)) & ~(0xffffffffu << dxb));
                    br_drop(b, used + dcl + dxb);
                    const bool viol = dist > op - a0 || op + len > oend;
                    bad |= viol;
                    MD_SYNCWARP();                               // the bytes being copied were stored by other lanes
                    const uint32_t s0 = op - dist;
                    if (dist <= (uint32_t) NEAR) {               // source inside the ring
                        if (dist >= len) {
                            if (len <= 32u && nl == 32) { if ((uint32_t) lane < len) D.win[(op + (uint32_t) lane) & WMASK] = D.win[(s0 + (uint32_t) lane) & WMASK]; }
                            else for (uint32_t k = (uint32_t) lane; k < len; k += (uint32_t) nl) D.win[(op + k) & WMASK] = D.win[(s0 + k) & WMASK];
                        }
                        else if (nl == 1) { for (uint32_t k = 0; k < len; ++k) D.win[(op + k) & WMASK] = D.win[(s0 + k) & WMASK]; }
                        else { for (uint32_t k = (uint32_t) lane; k < len; k += (uint32_t) nl) D.win[(op + k) & WMASK] = D.win[(s0 + k % dist) & WMASK]; }   // overlapping match = the last `dist` bytes repeated
                    } else {                                     // far match: those bytes left the ring, but were flushed long ago (dist - len >= FLUSH)
                        const uint32_t lim = viol ? 0u : len;    // never form an address from a distance that reaches before the block
                        for (uint32_t k = (uint32_t) lane; k < lim; k += (uint32_t) nl) {
#if defined(__CUDA_ARCH__)
                            D.win[(op + k) & WMASK] = __ldcg(outb + s0 + k);
#else
                            D.win[(op + k) & WMASK] = outb[s0 + k];
#endif
                        }
                    }
                    op += len;
                } else if (!(e & 0x200u)) {                       // literal
                    br_drop(b, cl);
                    bad |= op >= oend;
                    if (lane == 0) D.win[op & WMASK] = (uint8_t)(e >> 16);
                    ++op;
                } else { br_drop(b, cl); break; }                 // end of block
                if (op >= next_flush) {
                    if (bad) return -4;
                    flush_ring(D, outb, flushed, next_flush, lane, nl); flushed = next_flush; next_flush += FLUSH;
                }
            }
            if (bad) return -4;
        } else return -1;
        if (last) break;
    }
    if (op != oend) return -6;
    if (op > flushed) flush_ring(D, outb, flushed, op, lane, nl);
    return 0;
}

}  // namespace mdinflate
