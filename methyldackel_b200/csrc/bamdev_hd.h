// Per-thread bodies of the device decode kernels (bamdev.cu), written host/device so that the host self-test can run the
// very same logic over a BAM file in plain loops and compare it with the host decoder.
#pragma once
#include "bamrec_hd.h"

namespace mdbam {

struct BlockScan { unsigned long long guess, exit; uint32_t count, pad; };

// uoff[b]..uoff[b+1] is block b's slice of the stream; `first` is the known start of the first record of the segment
MD_HD void scan_block_body(uint32_t b, const uint8_t *u, const unsigned long long *uoff, unsigned long long first, unsigned long long U, int32_t n_targets, BlockScan *out) {
    const unsigned long long lo = uoff[b], hi = uoff[b + 1];
    BlockScan s; s.guess = NONE; s.exit = NONE; s.count = 0; s.pad = 0;
    if (hi > lo) {
        unsigned long long g;
        if (b == 0) g = first < hi ? first : NONE;
        else g = guess_start(u, lo, hi, U, n_targets);
        if (g != NONE) { uint64_t ex; s.count = walk_chain(u, g, hi, U, &ex, nullptr); s.guess = g; s.exit = ex; }
    }
    out[b] = s;
}
// Is block b's guessed chain the true one, given that all earlier ones are?  The chain enters the block at the exit of
// the nearest earlier block that has a chain; if it does not enter it at all (a record from before covers the block, or
// the chain ended in a partial record earlier) the block must not contribute a chain of its own.
MD_HD bool check_block_body(uint32_t b, const unsigned long long *uoff, unsigned long long first, const BlockScan *sc) {
    const unsigned long long lo = uoff[b], hi = uoff[b + 1];
    if (hi == lo) return sc[b].guess == NONE;
    unsigned long long expect = first;
    for (uint32_t p = b; p-- > 0;) if (sc[p].guess != NONE) { expect = sc[p].exit; break; }
    return (expect >= lo && expect < hi) ? sc[b].guess == expect : sc[b].guess == NONE;
}
// Serial pass (one thread): with bad == 0 it only reports where the stream's chain ends; otherwise it re-walks every block
// whose guess is not the true entry point and cancels chains the stream's chain never reaches.
MD_HD void fix_chain_body(const uint8_t *u, const unsigned long long *uoff, uint32_t n_blocks, unsigned long long first, unsigned long long U, BlockScan *sc, int bad, unsigned long long *final_exit) {
    if (!bad) {
        unsigned long long e = first;
        for (uint32_t p = n_blocks; p-- > 0;) if (sc[p].guess != NONE) { e = sc[p].exit; break; }
        *final_exit = e;
        return;
    }
    unsigned long long s = first;
    for (uint32_t b = 0; b < n_blocks; ++b) {
        const unsigned long long lo = uoff[b], hi = uoff[b + 1];
        BlockScan r; r.guess = NONE; r.exit = NONE; r.count = 0; r.pad = 0;
        if (hi > lo && s >= lo && s < hi) {
            if (sc[b].guess == s) { s = sc[b].exit; continue; }
            uint64_t ex; r.count = walk_chain(u, s, hi, U, &ex, nullptr); r.guess = s; r.exit = ex;
            if (ex == s) { r.guess = NONE; r.exit = NONE; r.count = 0; }      // partial record: the chain ends here
            else s = ex;
        }
        sc[b] = r;
    }
    *final_exit = s;
}
MD_HD void fill_offsets_body(uint32_t b, const uint8_t *u, const unsigned long long *uoff, unsigned long long U, const BlockScan *sc, const uint32_t *base, unsigned long long *rec_off) {
    if (sc[b].guess == NONE || !sc[b].count) return;
    uint64_t ex;
    walk_chain(u, sc[b].guess, uoff[b + 1], U, &ex, (uint64_t *)(rec_off + base[b]));
}
// returns false for a malformed record
MD_HD bool head_body(uint32_t i, const uint8_t *u, const unsigned long long *rec_off, int32_t *tid, int32_t *pos, int32_t *rend) {
    const uint8_t *r = u + rec_off[i];
    const Head h = head_of(r);
    if (!head_ok(h)) { tid[i] = -1; pos[i] = 0; rend[i] = 0; return false; }
    tid[i] = h.tid; pos[i] = h.pos; rend[i] = ref_end(r, h);
    return true;
}

// ---- tile assembly ------------------------------------------------------------------------------------------------
struct SoaView {            // device (or host) arrays of a tile, as md_reads_soa
    const int32_t *pos; const uint16_t *flag; const uint8_t *mapq; const uint8_t *aux; const uint32_t *l_qseq;
    const uint32_t *cigar_off, *seq_off, *qual_off; const uint64_t *frag_key; const uint32_t *cigar, *seq; const uint64_t *qual;
    const uint32_t *name_chk;   // second hash of the query name (device-built tiles)
};
struct TileSrc {
    const uint8_t *u; const unsigned long long *rec_off; const int32_t *pos, *rend;   // own records [r0, r0+n_own) of the segment
    uint32_t r0, n_own;
    SoaView prev; const int32_t *prev_rend; uint32_t n_prev;                           // previous tile on this contig (carry source)
    uint32_t keep_lo, keep_hi;      // a read stays if max(rend,pos+1) > keep_lo and (own records only) pos < keep_hi
};
struct TileDst {
    int32_t *pos; uint16_t *flag; uint8_t *mapq, *aux; uint32_t *l_qseq, *cigar_off, *seq_off, *qual_off; uint64_t *frag_key; int32_t *rend;
    uint32_t *cigar, *seq; uint64_t *qual;
    uint32_t *name_chk;
};
struct Sz4 { uint32_t x, y, z, w; };     // reads, cigar words, seq words, qual words

MD_HD Sz4 tile_sizes_body(uint32_t e, const TileSrc &S) {
    Sz4 v; v.x = v.y = v.z = v.w = 0;
    if (e < S.n_prev) {
        const int32_t p = S.prev.pos[e], re = S.prev_rend[e];
        const int32_t se = re > p + 1 ? re : p + 1;
        if ((long long) se > (long long) S.keep_lo) {
            const uint32_t l = S.prev.l_qseq[e];
            v.x = 1u; v.y = S.prev.cigar_off[e + 1] - S.prev.cigar_off[e]; v.z = seq_words(l); v.w = qual_words8(l);
        }
    } else {
        const uint32_t i = S.r0 + (e - S.n_prev);
        const int32_t p = S.pos[i], re = S.rend[i];
        const int32_t se = re > p + 1 ? re : p + 1;
        if ((long long) se > (long long) S.keep_lo && (long long) p < (long long) S.keep_hi) {
            const Head h = head_of(S.u + S.rec_off[i]);
            v.x = 1u; v.y = h.n_cigar; v.z = seq_words(h.l_seq); v.w = qual_words8(h.l_seq);
        }
    }
    return v;
}
MD_HD void tile_gather_body(uint32_t e, const TileSrc &S, const Sz4 &v, const Sz4 &o, TileDst &D) {
    if (e == S.n_prev + S.n_own - 1) D.cigar_off[o.x + v.x] = o.y + v.y;      // the n+1'th entry
    if (!v.x) return;
    const uint32_t k = o.x;
    D.cigar_off[k] = o.y; D.seq_off[k] = o.z; D.qual_off[k] = o.w;
    if (e < S.n_prev) {
        const SoaView &P = S.prev;
        D.pos[k] = P.pos[e]; D.flag[k] = P.flag[e]; D.mapq[k] = P.mapq[e]; D.aux[k] = P.aux[e]; D.l_qseq[k] = P.l_qseq[e]; D.frag_key[k] = P.frag_key[e]; D.name_chk[k] = P.name_chk[e]; D.rend[k] = S.prev_rend[e];
        const uint32_t c0 = P.cigar_off[e], s0 = P.seq_off[e], q0 = P.qual_off[e];
        for (uint32_t j = 0; j < v.y; ++j) D.cigar[o.y + j] = P.cigar[c0 + j];
        for (uint32_t j = 0; j < v.z; ++j) D.seq[o.z + j] = P.seq[s0 + j];
        for (uint32_t j = 0; j < v.w; ++j) D.qual[o.w + j] = P.qual[q0 + j];
    } else {
        const uint32_t i = S.r0 + (e - S.n_prev);
        const uint8_t *r = S.u + S.rec_off[i];
        const Head h = head_of(r);
        D.pos[k] = h.pos; D.flag[k] = (uint16_t) h.flag; D.mapq[k] = (uint8_t) h.mapq; D.aux[k] = aux_bits(r, h); D.l_qseq[k] = h.l_seq;
        D.frag_key[k] = name_key(r, h); D.name_chk[k] = name_check(r, h); D.rend[k] = S.rend[i];
        const uint8_t *c = cigar_of(r, h);
        for (uint32_t j = 0; j < h.n_cigar; ++j) D.cigar[o.y + j] = ld32(c + 4 * j);
        bytes_to_words32(D.seq + o.z, seq_of(r, h), (h.l_seq + 1u) >> 1);
        bytes_to_words64(D.qual + o.w, qual_of(r, h), h.l_seq);
    }
}

}  // namespace mdbam
