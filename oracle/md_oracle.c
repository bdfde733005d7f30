/* TEST INFRASTRUCTURE — see md_oracle.h.  Plain-C restatement of the reference's
 * extract / mbias counting path on SoA tiles.  Every function cites the reference
 * lines it follows.  The shape is deliberately simple (whole-tile dense arrays,
 * sequential loops); it is a checker, not a fast implementation. */
#include "md_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* ---- getStrand, common.c:84-116.  aux bits0-1: XG first value byte ('C' -> 1, 'G' -> 2) ---- */
int mdo_strand(uint16_t f, uint8_t aux) {
    int xg = aux & 3;
    if (xg == 0) {
        if (f & 1) {
            if ((f & 0x50) == 0x50) return 2;
            else if (f & 0x40) return 1;
            else if ((f & 0x90) == 0x90) return 1;
            else if (f & 0x80) return 2;
            return 0;
        }
        return (f & 0x10) ? 2 : 1;
    } else if (xg == 1) {
        if ((f & 0x51) == 0x41) return 1;
        else if ((f & 0x51) == 0x51) return 3;
        else if ((f & 0x91) == 0x81) return 3;
        else if ((f & 0x91) == 0x91) return 1;
        else if (f & 0x10) return 3;
        return 1;
    } else {
        if ((f & 0x51) == 0x41) return 4;
        else if ((f & 0x51) == 0x51) return 2;
        else if ((f & 0x91) == 0x81) return 2;
        else if ((f & 0x91) == 0x91) return 4;
        else if (f & 0x10) return 2;
        return 4;
    }
}

/* ---- filter_func, common.c:416-430 (the tests that only need flag / MAPQ / NH) ---- */
int mdo_admit(const md_config *c, uint16_t f, uint8_t mapq, uint8_t aux) {
    if (f & 0x4) return 0;                                             /* :416 */
    if ((int) mapq < c->minMapq) return 0;                             /* :417 */
    if (f & c->ignoreFlags) return 0;                                  /* :418 */
    if (c->requireFlags && (f & c->requireFlags) != c->requireFlags) return 0; /* :419 */
    if (!c->keepDupes && (f & 0x400)) return 0;                        /* :420 */
    if (!c->ignoreNH && (aux & 4)) return 0;                           /* :421-427 */
    if (!c->keepSingleton && (f & 0x9) == 0x9) return 0;               /* :429 */
    if (!c->keepDiscordant && (f & 0x3) == 0x1) return 0;              /* :430 */
    return 1;
}

/* ---- isCpG / isCHG / isCHH, common.c:49-82, chained as extract.c:407-418 does ---- */
static int is_c(char b) { return b == 'C' || b == 'c'; }
static int is_g(char b) { return b == 'G' || b == 'g'; }
int mdo_context(const char *seq, int pos, int seqlen) {
    if (pos < 0 || pos >= seqlen) return 0;
    if (is_c(seq[pos])) {
        if (pos + 1 != seqlen && is_g(seq[pos + 1])) return 1;         /* isCpG :51-54 */
        if (!(pos + 2 >= seqlen) && is_g(seq[pos + 2])) return 2;      /* isCHG :65-68 */
        return 3;                                                      /* isCHH :79 */
    } else if (is_g(seq[pos])) {
        if (pos != 0 && is_c(seq[pos - 1])) return -1;                 /* isCpG :55-58 */
        if (!(pos <= 1) && is_c(seq[pos - 2])) return -2;              /* isCHG :69-72 */
        return -3;                                                     /* isCHH :80 */
    }
    return 0;
}

/* ---- overlaps.c:103,106: `q += 0.2*q` on a uint8_t, i.e. (uint8_t)(q + 0.2*q); the conversion
 * of an out-of-range double is what x86-64 gcc emits (truncate to int, keep the low byte) ---- */
uint8_t mdo_boost(uint8_t q) { double v = (double) q + 0.2 * (double) q; return (uint8_t)(int) v; }

static inline uint8_t seq_nib(const md_reads_soa *r, uint32_t i, uint32_t q) {
    const uint8_t *s = (const uint8_t *)(r->seq + r->seq_off[i]);
    return (uint8_t)((s[q >> 1] >> ((~q & 1) << 2)) & 0xf);           /* bam_seqi */
}
static inline uint8_t qual_at(const md_reads_soa *r, uint32_t i, uint32_t q) {
    const uint8_t *b = (const uint8_t *)(r->qual + r->qual_off[i]);
    if (r->qual_bits == 0 || r->qual_bits == 8) return b[q];
    /* packed tile (include/mdgpu.h): qual_bits-wide codes, base q at bit q*qual_bits; the table gives back the phred byte */
    const uint32_t bit = q * r->qual_bits;
    return r->qual_lut[(b[bit >> 3] >> (bit & 7)) & ((1u << r->qual_bits) - 1u)];
}

typedef struct {
    uint8_t *b, *q;     /* per-base effective base (nibble) / phred of every read, after trims + overlap */
    uint64_t *off;      /* start of read i in b/q */
    uint8_t *admit, *strand;
    int32_t *rend;
} work_t;

static void free_work(work_t *w) { free(w->b); free(w->q); free(w->off); free(w->admit); free(w->strand); free(w->rend); }

/* computeConversionEfficiency, common.c:361-404, on the window contig[ce_beg, ce_end) of the chunk.
 * Restated with its quirks: the reference position is NOT advanced after a match op (common.c:373-391 has no
 * `pos += opLen`), CpG positions are skipped, the walk stops at the window end (:378).  Window indices below 0
 * (a read starting before the window) read out of bounds in the reference; they are treated as "no context" here. */
static float conversion_efficiency(const md_config *c, const md_reads_soa *r, uint32_t i, int strand, const char *ref, uint32_t ce_beg, uint32_t ce_end) {
    unsigned nMethyl = 0, nUMethyl = 0;
    int64_t pos = r->pos[i]; uint32_t seqPos = 0;
    const int lseq = (int)(ce_end - ce_beg);
    for (uint32_t k = r->cigar_off[i]; k < r->cigar_off[i + 1]; ++k) {
        uint32_t op = r->cigar[k] & 15, len = r->cigar[k] >> 4;
        if (op == 0 || op == 7 || op == 8) {
            for (uint32_t j = 0; j < len; ++j, ++seqPos) {
                if (pos + j >= (int64_t) ce_end) goto done;                                     /* :378 */
                int64_t idx = pos + j - (int64_t) ce_beg;
                if (idx < 0) continue;
                int ctx = mdo_context(ref + ce_beg, (int) idx, lseq);
                if (ctx == 0 || ctx == 1 || ctx == -1) continue;                               /* :379-380: CpG skipped */
                /* getMethylState, common.c:338-354: raw base and phred (trims come later, :458) */
                int state = 0; uint8_t b = seq_nib(r, i, seqPos);
                if ((int) qual_at(r, i, seqPos) >= c->minPhred) {
                    if (b == 2 && (strand & 1)) state = 1; else if (b == 8 && (strand & 1)) state = -1;
                    else if (b == 4 && !(strand & 1)) state = 1; else if (b == 1 && !(strand & 1)) state = -1;
                }
                if (state > 0) nMethyl++; else if (state < 0) nUMethyl++;
            }
        } else if (op == 1 || op == 4) seqPos += len;
        else if (op == 2 || op == 3) pos += len;
    }
done:
    if (nMethyl + nUMethyl == 0) return 1.0f;                                                   /* :357 */
    return nUMethyl / ((float)(nMethyl + nUMethyl));                                            /* :358 */
}

/* filter + strand + trimAlignment (common.c:137-172) + trimAbsoluteAlignment (common.c:174-208) */
/* ---- -l <BED> (bed.c).  The regions of the contig being processed, in sortBED order; on = a BED file is in play ---- */
static const md_bed_region *g_bed = 0; static uint32_t g_nbed = 0; static int g_bed_on = 0;
void mdo_set_bed(const md_bed_region *regs, uint32_t n, int on) { g_bed = regs; g_nbed = n; g_bed_on = on; }
/* compareRegions (bed.c:10-16) of region k, given as [start, end-1] by the callers (bed.c:29,32), against [start1, end1) */
static int64_t bed_cmp(uint32_t k, int64_t start1, int64_t end1) {
    int64_t start0 = g_bed[k].start, end0 = (int64_t) g_bed[k].end - 1;
    if (start0 < start1 && end0 >= start1) return 0;
    if (start0 >= start1 && start0 < end1) return 0;
    return start0 - start1;
}
/* spanOverlapsBED on a read (common.c:432-439, bed.c:22-41) from a fresh cursor: the first region that is not before the read decides */
static int bed_read_ok(int64_t pos, int64_t endpos) {
    for (uint32_t k = 0; k < g_nbed; ++k) { int64_t rv = bed_cmp(k, pos, endpos); if (rv >= 0) return rv == 0; }
    return 0;
}
/* posOverlapsBED (bed.c:46-54) driven as in extract.c:403: regions whose end is <= pos are stepped over, the next one decides.
 * Returns 1 when pos is inside it and sets *strand to that region's strand */
static int bed_pos_ok(int64_t pos, int *strand) {
    for (uint32_t k = 0; k < g_nbed; ++k) {
        if (pos >= (int64_t) g_bed[k].end) continue;
        if (pos < (int64_t) g_bed[k].start) return 0;
        *strand = (int) g_bed[k].strand; return 1;
    }
    return 0;
}
/* readStrandOverlapsBED (bed.c:57-63) */
static int bed_strand_ok(int region_strand, int s) {
    if (region_strand == 1) return s == 1 || s == 3;
    if (region_strand == 2) return s == 2 || s == 4;
    return 1;
}

static int build_work(const md_config *c, const md_reads_soa *r, work_t *w, uint32_t *n_adm, const char *ref, uint32_t ce_beg, uint32_t ce_end) {
    uint32_t n = r->n_reads;
    memset(w, 0, sizeof *w);
    w->off = (uint64_t *) malloc(((size_t) n + 1) * sizeof(uint64_t));
    w->admit = (uint8_t *) calloc((size_t) n + 1, 1);
    w->strand = (uint8_t *) calloc((size_t) n + 1, 1);
    w->rend = (int32_t *) calloc((size_t) n + 1, sizeof(int32_t));
    uint64_t tot = 0;
    for (uint32_t i = 0; i < n; ++i) { w->off[i] = tot; tot += r->l_qseq[i]; }
    w->off[n] = tot;
    w->b = (uint8_t *) malloc(tot + 1); w->q = (uint8_t *) malloc(tot + 1);
    if (!w->off || !w->admit || !w->strand || !w->rend || !w->b || !w->q) return -1;
    *n_adm = 0;
    for (uint32_t i = 0; i < n; ++i) {
        int rl = 0; uint32_t qlen = 0;
        for (uint32_t k = r->cigar_off[i]; k < r->cigar_off[i + 1]; ++k) {
            uint32_t op = r->cigar[k] & 15, len = r->cigar[k] >> 4;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rl += (int) len;
            if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) qlen += len;
        }
        w->rend[i] = r->pos[i] + rl;
        int s = mdo_strand(r->flag[i], r->aux[i]);
        w->strand[i] = (uint8_t) s;
        /* a record whose CIGAR does not describe its SEQ, with no reference span, or with an
         * undeterminable strand (the reference asserts, common.c:122-125) cannot be piled up */
        int ok = mdo_admit(c, r->flag[i], r->mapq[i], r->aux[i]) && s != 0 && rl > 0 && qlen == r->l_qseq[i] && r->l_qseq[i] > 0;
        if (ok && g_bed_on && !bed_read_ok(r->pos[i], w->rend[i])) ok = 0;                     /* common.c:432-439 */
        if (ok && c->minConversionEfficiency > 0.0f && conversion_efficiency(c, r, i, s, ref, ce_beg, ce_end) < c->minConversionEfficiency) ok = 0;   /* common.c:442-444 */
        w->admit[i] = (uint8_t) ok;
        if (!ok) continue;
        ++*n_adm;
        int l = (int) r->l_qseq[i];
        uint8_t *b = w->b + w->off[i], *q = w->q + w->off[i];
        for (int j = 0; j < l; ++j) { b[j] = seq_nib(r, i, (uint32_t) j); q[j] = qual_at(r, i, (uint32_t) j); }
        int base = 4 * (s - 1) + ((r->flag[i] & 0x80) ? 2 : 0);
        int lb = c->bounds[base], rb = c->bounds[base + 1];
        if (lb > l) lb = l;                                            /* common.c:151 */
        for (int j = 0; j < lb; ++j) { q[j] = 0; b[j] = 15; }          /* :154-160 */
        if (rb) for (int j = rb; j < l; ++j) { q[j] = 0; b[j] = 15; }  /* :163-169 */
        lb = c->absoluteBounds[base]; rb = c->absoluteBounds[base + 1];
        if (lb > l) lb = l;
        if (rb > l) rb = l;                                            /* :188-189 */
        for (int j = 0; j < lb; ++j) { q[j] = 0; b[j] = 15; }
        for (int j = 0; j < rb; ++j) { q[l - 1 - j] = 0; b[l - 1 - j] = 15; } /* :199-205 */
    }
    return 0;
}

/* calculate_positions, overlaps.c:27-52 */
static int32_t *calc_positions(const md_reads_soa *r, uint32_t i) {
    int32_t *p = (int32_t *) malloc(sizeof(int32_t) * ((size_t) r->l_qseq[i] + 1));
    int off = 0; int32_t prev = r->pos[i];
    for (uint32_t k = r->cigar_off[i]; k < r->cigar_off[i + 1]; ++k) {
        uint32_t op = r->cigar[k] & 15, len = r->cigar[k] >> 4;
        for (uint32_t j = 0; j < len; ++j) {
            if (op == 0 || op == 7 || op == 8) p[off++] = prev++;
            else if (op == 1 || op == 4) p[off++] = -1;
            else if (op == 2 || op == 3) prev++;
        }
    }
    return p;
}

/* cust_tweak_overlap_quality, overlaps.c:54-119 (a = first in file order) */
static void tweak(const md_reads_soa *r, work_t *w, uint32_t a, uint32_t b) {
    int sa = w->strand[a], sb = w->strand[b];
    if (((sa - sb) & 1) == 1) return;                                  /* :65 */
    int na = (int) r->l_qseq[a], nb = (int) r->l_qseq[b], ia = 0, ib = 0;
    int32_t *pa = calc_positions(r, a), *pb = calc_positions(r, b);
    uint8_t *aq = w->q + w->off[a], *bq = w->q + w->off[b], *as = w->b + w->off[a], *bs = w->b + w->off[b];
    while (ia < na && pa[ia] < 0) ia++;
    while (ib < nb && pb[ib] < 0) ib++;
    if (ia == na || ib == nb) goto quit;
    if (pa[ia] < pb[ib]) { while (ia < na && pa[ia] < pb[ib]) ia++; }
    else { while (ib < nb && pb[ib] < pa[ia]) ib++; }
    if (ia == na || ib == nb) goto quit;
    while (ia < na && ib < nb) {
        if (pa[ia] < pb[ib] || pa[ia] < 0) { ia++; continue; }
        if (pb[ib] < pa[ia] || pb[ib] < 0) { ib++; continue; }
        if (as[ia] != bs[ib]) {
            if (aq[ia] > bq[ib] && as[ia] != 15) { aq[ia] = (uint8_t)(aq[ia] - bq[ib]); bq[ib] = 0; }
            else if (bq[ib] > aq[ia] && bs[ib] != 15) { bq[ib] = (uint8_t)(bq[ib] - aq[ia]); aq[ia] = 0; }
            else { aq[ia] = 0; bq[ib] = 0; }
        } else {
            if (aq[ia] > bq[ib]) { aq[ia] = mdo_boost(aq[ia]); bq[ib] = 0; }
            else { bq[ib] = mdo_boost(bq[ib]); aq[ia] = 0; }
        }
        ia++; ib++;
    }
quit:
    free(pa); free(pb);
}

/* Emulation of the qname hash traffic generated by the pileup engine around
 * custom_overlap_constructor / custom_overlap_destructor (overlaps.c:121-147):
 *  - records are pushed in file order; an eligible record (paired, !(flag&12), :128) either
 *    stores itself under its name or, if the name is present, is merged with the stored
 *    record and the name is removed (:129-136);
 *  - a buffered record is dropped — destructor, which removes whatever is stored under that
 *    record's name (:141-147) — once a record starting beyond its end has been pushed.
 * Keys are the 64-bit name fingerprints of the tile. */
typedef struct { uint64_t key; uint32_t idx; uint8_t live; } slot_t;
typedef struct { slot_t *s; uint32_t cap; } map_t;
static uint32_t map_find(const map_t *m, uint64_t key) {
    uint32_t mask = m->cap - 1, i = (uint32_t)(key * 0x9e3779b97f4a7c15ull >> 40) & mask;
    while (m->s[i].live) { if (m->s[i].live == 1 && m->s[i].key == key) return i; i = (i + 1) & mask; }
    return m->cap;
}
static void map_put(map_t *m, uint64_t key, uint32_t idx) {
    uint32_t mask = m->cap - 1, i = (uint32_t)(key * 0x9e3779b97f4a7c15ull >> 40) & mask;
    while (m->s[i].live == 1) i = (i + 1) & mask;   /* reuse tombstones (live==2) */
    m->s[i].key = key; m->s[i].idx = idx; m->s[i].live = 1;
}

static int cmp_end(const void *a, const void *b) {
    const int64_t x = *(const int64_t *) a, y = *(const int64_t *) b;
    return (x > y) - (x < y);
}

static int pair_and_merge(const md_reads_soa *r, work_t *w, uint32_t *n_pairs, uint32_t *n_multi) {
    uint32_t n = r->n_reads, cap = 16;
    while (cap < 4 * (n + 1)) cap <<= 1;
    map_t m; m.cap = cap; m.s = (slot_t *) calloc(cap, sizeof(slot_t));
    /* eviction order: by reference end; (end<<32 | idx) sorted ascending */
    int64_t *ev = (int64_t *) malloc(sizeof(int64_t) * ((size_t) n + 1));
    uint32_t *occ = (uint32_t *) calloc(cap, sizeof(uint32_t));   /* name multiplicity, for stats only */
    if (!m.s || !ev || !occ) return -1;
    uint32_t nev = 0;
    for (uint32_t i = 0; i < n; ++i) if (w->admit[i]) ev[nev++] = ((int64_t) w->rend[i] << 32) | i;
    qsort(ev, nev, sizeof(int64_t), cmp_end);
    uint32_t evp = 0;
    *n_pairs = 0; *n_multi = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (!w->admit[i]) continue;
        /* constructor */
        uint16_t f = r->flag[i];
        if ((f & 1) && !(f & 12)) {
            uint32_t k = map_find(&m, r->frag_key[i]);
            if (k == m.cap) map_put(&m, r->frag_key[i], i);
            else { tweak(r, w, m.s[k].idx, i); m.s[k].live = 2; ++*n_pairs; }
        }
        /* this push lets the engine build every column < pos[i]; records ending at or before such a
         * column are dropped: end <= pos[i]-1  <=>  end < pos[i] */
        while (evp < nev && (int32_t)(ev[evp] >> 32) < r->pos[i]) {
            uint32_t x = (uint32_t)(ev[evp] & 0xffffffff);
            if (x < i) { uint32_t k = map_find(&m, r->frag_key[x]); if (k != m.cap) m.s[k].live = 2; }
            ++evp;
        }
    }
    /* stats: names seen more than twice among eligible admitted records */
    {
        uint32_t mask = cap - 1;
        uint64_t *keys = (uint64_t *) calloc(cap, sizeof(uint64_t));
        if (keys) {
            for (uint32_t i = 0; i < n; ++i) {
                if (!w->admit[i] || !(r->flag[i] & 1) || (r->flag[i] & 12)) continue;
                uint64_t key = r->frag_key[i]; uint32_t j = (uint32_t)(key * 0x9e3779b97f4a7c15ull >> 40) & mask;
                while (occ[j] && keys[j] != key) j = (j + 1) & mask;
                keys[j] = key; occ[j]++;
            }
            for (uint32_t j = 0; j < cap; ++j) if (occ[j] > 2) *n_multi += occ[j];
            free(keys);
        }
    }
    free(m.s); free(ev); free(occ);
    return 0;
}

int mdo_extract_tile(const md_config *c, const char *ref, uint32_t reflen, uint32_t beg, uint32_t end,
                     const md_reads_soa *r, md_call *out, uint64_t cap, md_tile_stats *st) {
    return mdo_extract_tile_ce(c, ref, reflen, beg, end, 0, 0, r, out, cap, st);
}

int mdo_extract_tile_ce(const md_config *c, const char *ref, uint32_t reflen, uint32_t beg, uint32_t end, uint32_t ce_beg, uint32_t ce_end,
                        const md_reads_soa *r, md_call *out, uint64_t cap, md_tile_stats *st) {
    work_t w; uint32_t n_adm = 0, n_pairs = 0, n_multi = 0;
    if (end > reflen) end = reflen;
    if (beg > end) beg = end;
    if (ce_end == 0 || ce_end > reflen) ce_end = reflen;
    if (ce_beg > ce_end) ce_beg = ce_end;
    if (build_work(c, r, &w, &n_adm, ref, ce_beg, ce_end) < 0) { free_work(&w); return -2; }
    if (!c->noOverlapMerge && pair_and_merge(r, &w, &n_pairs, &n_multi) < 0) { free_work(&w); return -2; }
    size_t span = (size_t)(end - beg);
    uint32_t *nm = (uint32_t *) calloc(span + 1, 4), *nu = (uint32_t *) calloc(span + 1, 4), *noff = (uint32_t *) calloc(span + 1, 4), *nvar = (uint32_t *) calloc(span + 1, 4);
    if (!nm || !nu || !noff || !nvar) { free(nm); free(nu); free(noff); free(nvar); free_work(&w); return -2; }
    /* the per-column loop of extract.c:420-441, regrouped per alignment */
    for (uint32_t i = 0; i < r->n_reads; ++i) {
        if (!w.admit[i]) continue;
        int s = w.strand[i];
        const uint8_t *b = w.b + w.off[i], *q = w.q + w.off[i];
        int64_t p = r->pos[i]; uint32_t qi = 0;
        for (uint32_t k = r->cigar_off[i]; k < r->cigar_off[i + 1]; ++k) {
            uint32_t op = r->cigar[k] & 15, len = r->cigar[k] >> 4;
            if (op == 0 || op == 7 || op == 8) {
                for (uint32_t j = 0; j < len; ++j, ++p, ++qi) {
                    if (p < (int64_t) beg || p >= (int64_t) end) continue;             /* extract.c:400 */
                    if (g_bed_on) { int sd = 0; if (!bed_pos_ok(p, &sd) || !bed_strand_ok(sd, s)) continue; }   /* extract.c:402-405, 425 */
                    char rb = ref[p];
                    int ctx = mdo_context(ref, (int) p, (int) reflen);               /* extract.c:407-418 */
                    if (ctx == 0) continue;
                    int type = (ctx < 0 ? -ctx : ctx) - 1;
                    if ((type == 0 && !c->keepCpG) || (type == 1 && !c->keepCHG) || (type == 2 && !c->keepCHH)) continue;
                    size_t o = (size_t)(p - beg);
                    if ((s & 1) ? !is_c(rb) : !is_g(rb)) {                              /* extract.c:427-437 */
                        if ((int) q[qi] < c->minPhred) continue;                       /* isVariant, extract.c:229 */
                        noff[o]++;
                        if (s & 1) { if (b[qi] != 4 && b[qi] != 15) nvar[o]++; }        /* :232-234 */
                        else { if (b[qi] != 2 && b[qi] != 15) nvar[o]++; }            /* :235-237 */
                        continue;
                    }
                    if ((int) q[qi] < c->minPhred) continue;                           /* updateMetrics, common.c:127 */
                    if (s & 1) { if (b[qi] == 2) nm[o]++; else if (b[qi] == 8) nu[o]++; } /* :129-130 */
                    else { if (b[qi] == 4) nm[o]++; else if (b[qi] == 1) nu[o]++; }       /* :131-132 */
                }
            } else if (op == 1 || op == 4) qi += len;
            else if (op == 2 || op == 3) p += len;                                     /* is_del / is_refskip columns: extract.c:423-424 */
        }
    }
    uint64_t k = 0, need = 0;
    for (size_t o = 0; o < span; ++o) {
        int excluded = 0;
        if (c->minOppositeDepth > 0 && noff[o] >= (uint32_t) c->minOppositeDepth &&
            ((double) nvar[o]) / ((double) noff[o]) >= c->maxVariantFrac) excluded = 1;      /* extract.c:444-446 */
        if (!excluded && nm[o] + nu[o] == 0) continue;                                      /* extract.c:461 */
        if (excluded) {
            /* a column only reaches the variant test if it is a kept context (extract.c:407-418) */
            int ctx0 = mdo_context(ref, (int)(beg + o), (int) reflen);
            int t0 = (ctx0 < 0 ? -ctx0 : ctx0) - 1;
            if (ctx0 == 0 || (t0 == 0 && !c->keepCpG) || (t0 == 1 && !c->keepCHG) || (t0 == 2 && !c->keepCHH)) continue;
        }
        int ctx = mdo_context(ref, (int)(beg + o), (int) reflen);
        ++need;
        if (k < cap) {
            out[k].pos = beg + (uint32_t) o; out[k].nmeth = nm[o]; out[k].nunmeth = nu[o];
            out[k].info = (uint32_t)((ctx < 0 ? -ctx : ctx) - 1) | (ctx < 0 ? 4u : 0u) | (excluded ? 8u : 0u);
            ++k;
        }
    }
    if (st) { memset(st, 0, sizeof *st); st->n_calls = k; st->n_required = need; st->n_admitted = n_adm; st->n_pairs = n_pairs; st->n_multi = n_multi; }
    free(nm); free(nu); free(noff); free(nvar); free_work(&w);
    return need > cap ? -1 : 0;
}

int mdo_mbias_tile(const md_config *c, const char *ref, uint32_t reflen, uint32_t beg, uint32_t end,
                   const uint32_t *bounds, uint32_t n_chunks,
                   const md_reads_soa *r, uint32_t *hist, int32_t lens[4], md_tile_stats *st) {
    return mdo_mbias_tile_ce(c, ref, reflen, beg, end, 0, 0, bounds, n_chunks, r, hist, lens, st);
}

int mdo_mbias_tile_ce(const md_config *c, const char *ref, uint32_t reflen, uint32_t beg, uint32_t end, uint32_t ce_beg, uint32_t ce_end,
                      const uint32_t *bounds, uint32_t n_chunks,
                      const md_reads_soa *r, uint32_t *hist, int32_t lens[4], md_tile_stats *st) {
    work_t w; uint32_t n_adm = 0;
    if (end > reflen) end = reflen;
    if (ce_end == 0 || ce_end > reflen) ce_end = reflen;
    if (ce_beg > ce_end) ce_beg = ce_end;
    if (build_work(c, r, &w, &n_adm, ref, ce_beg, ce_end) < 0) { free_work(&w); return -2; }
    for (uint32_t i = 0; i < r->n_reads; ++i) {
        if (!w.admit[i]) continue;
        int s = w.strand[i];
        int rd2 = (r->flag[i] & 0x80) ? 1 : 0;
        const uint8_t *b = w.b + w.off[i], *q = w.q + w.off[i];
        int64_t p = r->pos[i]; uint32_t qi = 0;
        for (uint32_t k = r->cigar_off[i]; k < r->cigar_off[i + 1]; ++k) {
            uint32_t op = r->cigar[k] & 15, len = r->cigar[k] >> 4;
            if (op == 0 || op == 7 || op == 8) {
                for (uint32_t j = 0; j < len; ++j, ++p, ++qi) {
                    if (p < (int64_t) beg || p >= (int64_t) end) continue;             /* MBias.c:163 */
                    if (g_bed_on) { int sd = 0; if (!bed_pos_ok(p, &sd) || !bed_strand_ok(sd, s)) continue; }   /* MBias.c:165-168, 184 */
                    /* chunk window contig[localPos..localEnd] (MBias.c:147): find the chunk that owns p */
                    uint32_t lo = 0, hi = n_chunks;
                    while (lo + 1 < hi) { uint32_t mid = (lo + hi) >> 1; if (bounds[mid] <= (uint32_t) p) lo = mid; else hi = mid; }
                    if (n_chunks == 0 || (uint32_t) p < bounds[lo] || (uint32_t) p >= bounds[lo + 1]) continue;
                    uint32_t cs = bounds[lo], ce = bounds[lo + 1];
                    uint32_t last = ce < reflen ? ce : reflen - 1;                      /* end-inclusive fetch, clamped */
                    int seqlen = (int)(last - cs + 1);
                    int ctx = mdo_context(ref + cs, (int)((uint32_t) p - cs), seqlen);  /* MBias.c:170-178 */
                    if (ctx == 0) continue;
                    int type = (ctx < 0 ? -ctx : ctx) - 1;
                    if ((type == 0 && !c->keepCpG) || (type == 1 && !c->keepCHG) || (type == 2 && !c->keepCHH)) continue;
                    char rb = ref[p];
                    if ((s & 1) ? !is_c(rb) : !is_g(rb)) continue;                     /* MBias.c:186-190 */
                    if ((int) q[qi] < c->minPhred) continue;
                    int rv = 0;
                    if (s & 1) { if (b[qi] == 2) rv = 1; else if (b[qi] == 8) rv = -1; }
                    else { if (b[qi] == 4) rv = 1; else if (b[qi] == 1) rv = -1; }
                    if (rv == 0) continue;
                    if (qi >= MD_MBIAS_MAXLEN) continue;                               /* beyond the fixed histogram; counted by neither side */
                    hist[(((size_t)(s - 1) * 2 + rd2) * MD_MBIAS_MAXLEN + qi) * 2 + (rv < 0 ? 1 : 0)]++; /* MBias.c:195-211 */
                    if ((int32_t) qi + 1 > lens[s - 1]) lens[s - 1] = (int32_t) qi + 1;  /* MBias.c:212 */
                }
            } else if (op == 1 || op == 4) qi += len;
            else if (op == 2 || op == 3) p += len;
        }
    }
    if (st) { memset(st, 0, sizeof *st); st->n_admitted = n_adm; }
    free_work(&w);
    return 0;
}


/* ---- perRead (perRead.c:37-94, 183-196) -------------------------------------------------------------------------------- */
static int cpg_dir(const char *seq, int64_t pos, int64_t seqlen) {          /* isCpG, common.c:49-62 */
    if (pos >= seqlen) return 0;
    if (is_c(seq[pos])) { if (pos + 1 == seqlen) return 0; return is_g(seq[pos + 1]) ? 1 : 0; }
    if (is_g(seq[pos])) { if (pos == 0) return 0; return is_c(seq[pos - 1]) ? -1 : 0; }
    return 0;
}
int mdo_per_read_tile(const md_config *c, const char *ref, uint32_t reflen, uint32_t beg, uint32_t end, uint32_t chunk,
                      const md_reads_soa *r, md_read_meth *out) {
    static const int ctype[16] = {3, 1, 2, 2, 1, 0, 0, 3, 3, 0, 0, 0, 0, 0, 0, 0};   /* bam_cigar_type: MIDNSHP=XB */
    if (end > reflen) end = reflen;
    if (beg > end) beg = end;
    for (uint32_t i = 0; i < r->n_reads; ++i) {
        out[i].nmeth = 0xffffffffu; out[i].nunmeth = 0;
        int64_t pos = r->pos[i]; uint16_t f = r->flag[i];
        if (pos < (int64_t) beg || pos >= (int64_t) end) continue;                                   /* perRead.c:186-187 */
        if (c->requireFlags && (c->requireFlags & f) != c->requireFlags) continue;                   /* :189 */
        if (c->ignoreFlags && (c->ignoreFlags & f) != 0) continue;                                   /* :190 */
        if ((int) r->mapq[i] < c->minMapq) continue;                                                 /* :191 */
        /* chunk of the alignment's start and its reference window contig[localPos2 .. localEnd+10000] (perRead.c:118-137, 176-181) */
        uint64_t localPos = (uint64_t) beg + ((uint64_t)(pos - beg) / chunk) * chunk, localEnd = localPos + chunk;
        if (localEnd > end) localEnd = end;
        int64_t localPos2 = localPos > 1 ? (int64_t) localPos - 2 : 0;
        int64_t wend = (int64_t) localEnd + 10000; if (wend > (int64_t) reflen - 1) wend = (int64_t) reflen - 1;
        const char *seq = ref + localPos2; int64_t seqlen = wend - localPos2 + 1;
        int strand = mdo_strand(f, r->aux[i]);
        uint32_t lq = r->l_qseq[i], ncig = r->cigar_off[i + 1] - r->cigar_off[i];
        const uint32_t *cig = r->cigar + r->cigar_off[i];
        uint32_t readPosition = 0, op = 0, opOffset = 0, nmethyl = 0, nunmethyl = 0;
        uint64_t mappedPosition = (uint64_t) pos;
        while (readPosition < lq && op < ncig) {                                                    /* perRead.c:51 */
            if (opOffset >= (cig[op] >> 4)) { opOffset = 0; op++; if (op >= ncig) break; }            /* :52-55 (the reference then reads cig[ncig]) */
            int t = ctype[cig[op] & 15];
            if (t & 2) {
                if (t & 1) {
                    if ((int) qual_at(r, i, readPosition) < c->minPhred) { mappedPosition++; readPosition++; opOffset++; }   /* :59-63 */
                    int direction = cpg_dir(seq, (int64_t) mappedPosition - localPos2, seqlen);     /* :65 */
                    if (direction) {
                        /* bam_seqi(readSeq, readPosition); index l_qseq reads the nibble behind the sequence in the BAM record */
                        uint8_t base;
                        if (readPosition < lq) base = seq_nib(r, i, readPosition);
                        else if (lq & 1) base = (uint8_t)(((const uint8_t *)(r->seq + r->seq_off[i]))[lq >> 1] & 0xf);
                        else base = (uint8_t)(qual_at(r, i, 0) >> 4);
                        if (direction == 1 && (strand & 1) == 1) { if (base == 2) nmethyl++; else if (base == 8) nunmethyl++; }        /* :68-70 */
                        else if (direction == -1 && (strand & 1) == 0) { if (base == 4) nmethyl++; else if (base == 1) nunmethyl++; }  /* :71-74 */
                    }
                    mappedPosition++; readPosition++; opOffset++;
                } else { mappedPosition += cig[op++] >> 4; opOffset = 0; }                           /* :80-83 */
            } else if (t & 1) { readPosition += cig[op++] >> 4; opOffset = 0; }                      /* :85-88 */
            else { opOffset = 0; op++; }                                                             /* :89-93 */
        }
        out[i].nmeth = nmethyl; out[i].nunmeth = nunmethyl;
    }
    return 0;
}
