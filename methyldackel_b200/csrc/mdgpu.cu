// libmdgpu — hand-written CUDA (sm_100a) implementation of the MethylDackel extract / mbias
// pileup hot path behind the C ABI of include/mdgpu.h.
//
// The reference walks a pileup: for every covered reference column it resolves the CIGAR of every
// buffered read (htslib bam_plp, driven from extract.c:399) and then classifies read bases
// (extract.c:420-441).  That is column-major and serial.  Here the work is read-major inside
// position-owned windows:
//
//   K1 prep_kernel    grid-stride, one thread per alignment: filter_func's admission tests (common.c:416-430),
//                     getStrand (common.c:84-116), reference span, the optional conversion-efficiency filter
//                     (common.c:361-404) and mate pairing through a 4-byte-per-slot open-addressing table that stays
//                     in L2 (replaces the khash of overlaps.c:121-139).
//   K2 count_warp     one CTA per window of 4096 reference positions: stages the reference window in shared memory and
//                     classifies every position bit-parallel (isCpG/isCHG/isCHH, common.c:49-82); then every warp
//                     streams batches of 32 alignments through TMA bulk copies into its own shared-memory staging
//                     buffer, walks the CIGARs against bitmaps of the kept C-/G-sites, queues candidate bases with
//                     warp ballots and evaluates them 32 at a time: trims (common.c:137-208), the overlap phred merge
//                     (overlaps.c:54-119), updateMetrics / isVariant (common.c:118-134, extract.c:225-239), shared-memory
//                     atomicAdd; the epilogue applies the variant-site test (extract.c:444-446), drops empty columns
//                     (extract.c:461) and appends compact md_call records to HBM.  count_warp<2> is the mbias flavour
//                     (MBias.c:181-214: a per-(strand,read,qpos) histogram).
//   K3 dir_scan + gather   put the per-window segments into position order.
//
// Integer / byte work, HBM-bound by design: no tensor cores, no GEMM reshaping.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>
#include <algorithm>
#include "../../include/mdgpu.h"

// ------------------------------------------------------------------------------------------------
// error plumbing
// The last error of the calling thread; a thread that has none of its own (the CLI creates the context on a helper thread
// and reports from the calling one) sees the most recent error of the process.
#include <mutex>
static std::mutex g_err_mu; static std::string g_err_any;
struct ErrSlot {
    std::string s;
    ErrSlot &operator=(const std::string &v) { s = v; std::lock_guard<std::mutex> l(g_err_mu); g_err_any = v; return *this; }
    ErrSlot &operator=(const char *v) { return *this = std::string(v); }
    bool empty() const { return s.empty(); }
    const char *c_str() { if (!s.empty()) return s.c_str(); std::lock_guard<std::mutex> l(g_err_mu); s = g_err_any; return s.c_str(); }
};
static thread_local ErrSlot g_err;
static void set_err(const char *what, cudaError_t e) { g_err = std::string(what) + ": " + cudaGetErrorString(e); }
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err(#call, e_); return -100; } } while (0)
#define CKN(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err(#call, e_); return nullptr; } } while (0)
extern "C" const char *md_last_error(void) { return g_err.c_str(); }
extern "C" int md_abi_version(void) { return MDGPU_ABI_VERSION; }
#ifndef MD_SRC_HASH
#define MD_SRC_HASH "unknown"
#endif
extern "C" const char *md_source_hash(void) { return MD_SRC_HASH; }

// ------------------------------------------------------------------------------------------------
// device-side views
struct DevReads {
    uint32_t n; uint32_t seq_words, qual_words;   // totals of seq[] (32-bit words) / qual[] (64-bit words)
    uint32_t qbits; unsigned char qlut[16];       // phred encoding of qual[]: 8 = bytes, 2/4 = codes through qlut (md_reads_soa::qual_bits)
    const int32_t *pos; const uint16_t *flag; const uint8_t *mapq; const uint8_t *aux; const uint32_t *l_qseq;
    const uint32_t *cigar_off, *seq_off, *qual_off; const uint64_t *frag_key;
    const uint32_t *cigar, *seq; const uint64_t *qual;
    const uint32_t *name_chk;                       // second hash of the query name, or null (host-supplied tiles carry the 64-bit key only)
};

struct KParams {
    int minMapq, minPhred, keepDupes, keepSingleton, keepDiscordant, ignoreFlags, requireFlags, ignoreNH;
    int keepMask;              // bit0 CpG, bit1 CHG, bit2 CHH
    int minOppositeDepth; double maxVariantFrac;
    int bounds[16], abounds[16];
    int anyTrim;               // some --OT/.../--nOT/... bound is set (none: every read keeps [0, l_qseq), and the table look-ups are skipped)
    int noOverlap;
    float minCE;               // --minConversionEfficiency (common.c:442-444); 0 = off
    unsigned char boost[256];  // (uint8_t)(q + 0.2*q), overlaps.c:103,106, tabulated on the host in double
};

// info byte per read: bits0-2 strand (1..4), bit3 admitted, bit4 pair-eligible
#define INFO_STRAND(x) ((x) & 7)
#define INFO_ADMIT 8
#define INFO_ELIG 16

// Pairing table: one 32-bit slot per name = index of the first record seen with that name (0xffffffff = empty);
// the name itself is checked against frag_key[] of that record, so the table stays small enough to live in L2.
struct HashTab { uint32_t *tab; uint32_t cap; };

// -l BED regions of one contig (bed.c), sorted as sortBED does: start[], pmax[] = running maximum of end[] (so that the first
// region whose end lies beyond a position is a binary search), strand[] (0 any, 1 '+', 2 '-').  on = 0: no BED file in play.
struct BedView { const uint32_t *start, *pmax, *strand; uint32_t n, on; };
// index of the first region with end > pos — the region posOverlapsBED (bed.c:46-54) settles on for pos, and the first one
// spanOverlapsBED (bed.c:22-41) does not skip as "before" a read starting at pos
__device__ __forceinline__ uint32_t bed_first_beyond(const BedView &B, uint32_t pos) {
    uint32_t lo = 0, hi = B.n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(B.pmax + mid) > pos) hi = mid; else lo = mid + 1; }
    return lo;
}

// counters[]: 0 n_admitted, 1 max reference span, 2 n_pairs, 3 n_multi, 4-5 n_calls (64-bit append cursor), 6 overflow flag, 7 max l_qseq
enum { C_ADMIT = 0, C_MAXSPAN = 1, C_PAIRED = 2, C_MULTI = 3, C_NCALLS = 4 /* 64-bit: slots 4,5 */, C_OVERFLOW = 6, C_MAXLQ = 7, C_N = 8 };

// ------------------------------------------------------------------------------------------------
// per-read primitives
__device__ __forceinline__ int dev_strand(unsigned f, unsigned aux) {   // getStrand, common.c:84-116
    unsigned xg = aux & 3u;
    if (xg == 0) {
        if (f & 1u) {
            if ((f & 0x50u) == 0x50u) return 2;
            if (f & 0x40u) return 1;
            if ((f & 0x90u) == 0x90u) return 1;
            if (f & 0x80u) return 2;
            return 0;
        }
        return (f & 0x10u) ? 2 : 1;
    }
    int fwd1 = (f & 0x51u) == 0x41u, rev1 = (f & 0x51u) == 0x51u, fwd2 = (f & 0x91u) == 0x81u, rev2 = (f & 0x91u) == 0x91u;
    if (xg == 1) { if (fwd1) return 1; if (rev1) return 3; if (fwd2) return 3; if (rev2) return 1; return (f & 0x10u) ? 3 : 1; }
    if (fwd1) return 4; if (rev1) return 2; if (fwd2) return 2; if (rev2) return 4; return (f & 0x10u) ? 2 : 4;
}

__device__ __forceinline__ bool dev_admit(const KParams &P, unsigned f, unsigned mapq, unsigned aux) {   // common.c:416-430
    if (f & 0x4u) return false;
    if ((int) mapq < P.minMapq) return false;
    if (f & (unsigned) P.ignoreFlags) return false;
    if (P.requireFlags && (f & (unsigned) P.requireFlags) != (unsigned) P.requireFlags) return false;
    if (!P.keepDupes && (f & 0x400u)) return false;
    if (!P.ignoreNH && (aux & 4u)) return false;
    if (!P.keepSingleton && (f & 0x9u) == 0x9u) return false;
    if (!P.keepDiscordant && (f & 0x3u) == 0x1u) return false;
    return true;
}

// kept query range [lo,hi) after trimAlignment + trimAbsoluteAlignment (common.c:137-208);
// bases outside read as N with phred 0.
__device__ __forceinline__ void dev_trim(const KParams &P, int strand, unsigned f, int l, int &lo, int &hi) {
    if (!P.anyTrim) { lo = 0; hi = l; return; }
    int b = 4 * (strand - 1) + ((f & 0x80u) ? 2 : 0);
    int lb = min(P.bounds[b], l), rb = P.bounds[b + 1];
    lo = lb; hi = rb ? min(rb, l) : l;
    int alb = min(P.abounds[b], l), arb = min(P.abounds[b + 1], l);
    lo = max(lo, alb); hi = min(hi, l - arb);
}

__device__ __forceinline__ unsigned dev_base(const uint32_t *seq, uint32_t off, int q) {   // bam_seqi on word-packed bytes
    uint32_t w = __ldg(seq + off + (q >> 3));
    unsigned byte = (w >> (((q >> 1) & 3) << 3)) & 0xffu;
    return (q & 1) ? (byte & 0xfu) : (byte >> 4);
}
__device__ __forceinline__ unsigned dev_qual(const DevReads &R, uint32_t off, int q) {
    const unsigned char *p = (const unsigned char *) (R.qual + off);
    if (R.qbits == 8u) return __ldg(p + q);
    const unsigned bit = (unsigned) q * R.qbits;                        // 2- or 4-bit codes never straddle a byte
    return R.qlut[(__ldg(p + (bit >> 3)) >> (bit & 7u)) & ((1u << R.qbits) - 1u)];
}
// same, from the staged copy in shared memory
__device__ __forceinline__ unsigned staged_qual(const DevReads &R, const unsigned char *p, int q) {
    if (R.qbits == 8u) return p[q];
    const unsigned bit = (unsigned) q * R.qbits;
    return R.qlut[(p[bit >> 3] >> (bit & 7u)) & ((1u << R.qbits) - 1u)];
}
__device__ __forceinline__ uint32_t qual_words_of(const DevReads &R, uint32_t lq) { return (lq * R.qbits + 63u) >> 6; }

__device__ __forceinline__ uint32_t hash_slot(unsigned long long key, uint32_t cap) {
    const uint32_t h = (uint32_t)((key * 0x9e3779b97f4a7c15ull) >> 32);
    return (uint32_t)(((unsigned long long) h * cap) >> 32);               // multiply-shift range reduction, any cap
}

// ------------------------------------------------------------------------------------------------
// K1.  Besides admission / strand / span, resolves mate pairs: the first record of a name claims a table slot with its
// index (atomicCAS); a later record with the same name finds it, and claims it as its mate with one atomicCAS on mate[]
// (overlaps.c:129-135: the stored record is `a`).  A third record of the same name cannot claim anything: the tile is
// then flagged (C_MULTI) and replayed exactly on the host.  mate[] must be preset to -1 and tab[] to 0xffffffff.
// context classification on an absolute window [lo,hi) of the contig (common.c:49-82 chained as
// extract.c:407-418).  Returns 0, or (type+1) | isG<<2 with type 0 CpG, 1 CHG, 2 CHH.
__device__ __forceinline__ bool d_isC(unsigned char b) { return b == 'C' || b == 'c'; }
__device__ __forceinline__ bool d_isG(unsigned char b) { return b == 'G' || b == 'g'; }

template <class Get>
__device__ __forceinline__ unsigned dev_context(Get get, long long p, long long lo, long long hi) {
    if (p < lo || p >= hi) return 0;
    unsigned char c = get(p);
    if (d_isC(c)) {
        if (p + 1 != hi && d_isG(get(p + 1))) return 1;
        if (p + 2 < hi && d_isG(get(p + 2))) return 2;
        return 3;
    }
    if (d_isG(c)) {
        if (p != lo && d_isC(get(p - 1))) return 1 | 4;
        if (p - lo > 1 && d_isC(get(p - 2))) return 2 | 4;
        return 3 | 4;
    }
    return 0;
}

// computeConversionEfficiency (common.c:361-404) of one alignment inside the chunk window contig[ce_beg, ce_end).
// Quirks kept: the reference position is not advanced after a match op (no `pos += opLen` at common.c:373-391); CpG
// positions are skipped; the walk ends at the window end (:378).  Window indices below 0 — a read that starts before
// the window, where the reference reads out of bounds — count as "no context".
__device__ float dev_conversion_efficiency(const DevReads &R, const KParams &P, uint32_t i, int strand, const unsigned char *ref, uint32_t ce_beg, uint32_t ce_end) {
    unsigned nM = 0, nU = 0;
    long long pos = R.pos[i]; int seqPos = 0;
    const uint32_t soff = R.seq_off[i], qoff = R.qual_off[i];
    auto get = [&](long long x) -> unsigned char { return __ldg(ref + x); };
    for (uint32_t k = R.cigar_off[i], ke = R.cigar_off[i + 1]; k < ke; ++k) {
        const uint32_t c = __ldg(R.cigar + k), op = c & 15u; const int len = (int)(c >> 4);
        if (op == 0 || op == 7 || op == 8) {
            for (int j = 0; j < len; ++j, ++seqPos) {
                if (pos + j >= (long long) ce_end) goto done;
                if (pos + j < (long long) ce_beg) continue;
                const unsigned cx = dev_context(get, pos + j, (long long) ce_beg, (long long) ce_end);
                if (cx == 0 || (cx & 3u) == 1u) continue;                        // not a cytosine column, or CpG
                if ((int) dev_qual(R, qoff, seqPos) < P.minPhred) continue;      // getMethylState, common.c:347
                const unsigned b = dev_base(R.seq, soff, seqPos);
                if (strand & 1) { if (b == 2u) ++nM; else if (b == 8u) ++nU; }
                else { if (b == 4u) ++nM; else if (b == 1u) ++nU; }
            }
        } else if (op == 1 || op == 4) seqPos += len;
        else if (op == 2 || op == 3) pos += len;
    }
done:
    if (nM + nU == 0) return 1.0f;
    return __fdiv_rn((float) nU, (float)(nM + nU));
}

__global__ void __launch_bounds__(256, 8) prep_kernel(DevReads R, KParams P, int32_t *rend, uint8_t *info, HashTab T, int32_t *mate, uint32_t *counters,
                                                   const unsigned char *ref, uint32_t ce_beg, uint32_t ce_end, BedView B) {
    // grid-stride over the alignments; the five tile-wide statistics are reduced per thread, then per CTA, so the
    // global counters see one atomic per CTA instead of one per warp (same-address L2 atomics serialise)
    uint32_t n_ok = 0, n_paired = 0, n_multi = 0, span_max = 0, lq_max = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < R.n; i += gridDim.x * blockDim.x) {
        const unsigned f = R.flag[i], a = R.aux[i];
        const int strand = dev_strand(f, a);
        uint32_t ql = 0; int rl = 0;
        for (uint32_t k = R.cigar_off[i], ke = R.cigar_off[i + 1]; k < ke; ++k) {
            const uint32_t c = __ldg(R.cigar + k), op = c & 15u, len = c >> 4;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rl += (int) len;
            if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) ql += len;
        }
        const uint32_t lq = R.l_qseq[i];
        bool ok = dev_admit(P, f, R.mapq[i], a) && strand != 0 && rl > 0 && ql == lq && lq > 0;
        if (ok && B.on) {                                                       // common.c:432-439: the read must overlap a BED region (strand independent)
            const uint32_t j = bed_first_beyond(B, (uint32_t) R.pos[i]);
            ok = j < B.n && (long long) __ldg(B.start + j) < (long long) R.pos[i] + rl;
        }
        if (ok && P.minCE > 0.0f && dev_conversion_efficiency(R, P, i, strand, ref, ce_beg, ce_end) < P.minCE) ok = false;   // common.c:442-444
        const bool elig = ok && (f & 1u) && !(f & 12u) && !P.noOverlap;      // overlaps.c:128
        rend[i] = R.pos[i] + rl;
        info[i] = (uint8_t)(strand | (ok ? INFO_ADMIT : 0) | (elig ? INFO_ELIG : 0));
        lq_max = max(lq_max, lq);
        if (ok) { ++n_ok; span_max = max(span_max, (uint32_t) rl); }
        if (elig) {
            const unsigned long long key = R.frag_key[i];
            uint32_t h = hash_slot(key, T.cap);
            for (uint32_t probe = 0; probe < T.cap; ++probe) {
                const uint32_t old = atomicCAS(T.tab + h, 0xffffffffu, i);
                if (old == 0xffffffffu) break;                                 // first record of this name
                if (R.frag_key[old] == key && (!R.name_chk || R.name_chk[old] == R.name_chk[i])) {   // same name (64-bit key, and the 32-bit check where the tile carries it): pair up, or detect a third record
                    if (atomicCAS((int *) mate + old, -1, (int) i) == -1) { mate[i] = (int32_t) old; ++n_paired; }
                    else ++n_multi;
                    break;
                }
                if (++h == T.cap) h = 0;
            }
        }
    }
    __shared__ uint32_t red[5][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 16; o; o >>= 1) {
        n_ok += __shfl_xor_sync(0xffffffffu, n_ok, o); n_paired += __shfl_xor_sync(0xffffffffu, n_paired, o); n_multi += __shfl_xor_sync(0xffffffffu, n_multi, o);
        span_max = max(span_max, __shfl_xor_sync(0xffffffffu, span_max, o)); lq_max = max(lq_max, __shfl_xor_sync(0xffffffffu, lq_max, o));
    }
    if (lane == 0) { red[0][warp] = n_ok; red[1][warp] = n_paired; red[2][warp] = n_multi; red[3][warp] = span_max; red[4][warp] = lq_max; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
        for (int k = 0; k < 8; ++k) { a0 += red[0][k]; a1 += red[1][k]; a2 += red[2][k]; a3 = max(a3, red[3][k]); a4 = max(a4, red[4][k]); }
        if (a0) atomicAdd(counters + C_ADMIT, a0);
        if (a1) atomicAdd(counters + C_PAIRED, a1);
        if (a2) atomicAdd(counters + C_MULTI, a2);
        if (a3) atomicMax(counters + C_MAXSPAN, a3);
        if (a4) atomicMax(counters + C_MAXLQ, a4);
    }
}

// First index in [0,n) whose pos is > key (pos ascending), found by the whole warp: 32 probes per step, ~log32(n) dependent loads.
__device__ __forceinline__ uint32_t warp_upper_bound(const int32_t *pos, uint32_t n, long long key, int lane) {
    uint32_t lo = 0, hi = n;
    while (hi > lo) {
        const uint32_t span = hi - lo, step = (span + 31u) / 32u;
        const uint32_t idx = lo + (uint32_t) lane * step;
        const bool le = idx < hi && (long long) __ldg(pos + idx) <= key;        // monotone: true for every index below the answer
        const unsigned k = (unsigned) __popc(__ballot_sync(0xffffffffu, le));
        if (k == 0) { hi = lo; break; }
        const uint32_t nlo = lo + (k - 1u) * step + 1u, nhi = min(lo + k * step, hi);
        lo = nlo; hi = nhi;
    }
    return lo;
}

// Locate the query index of reference position rp in a read (M/=/X only); -1 if rp falls in D/N or outside.
__device__ __forceinline__ int dev_qpos_at(const uint32_t *cigar, uint32_t k, uint32_t ke, int pos, int rp) {
    int p = pos, q = 0;
    for (; k < ke; ++k) {
        uint32_t c = __ldg(cigar + k), op = c & 15u; int len = (int)(c >> 4);
        if (op == 0 || op == 7 || op == 8) { if (rp < p + len) return rp >= p ? q + (rp - p) : -1; p += len; q += len; }
        else if (op == 1 || op == 4) q += len;
        else if (op == 2 || op == 3) { if (rp < p + len) return -1; p += len; }
    }
    return -1;
}

struct CountArgs {
    DevReads R; KParams P;
    const int32_t *rend; const uint8_t *info; const int32_t *mate; const uint2 *win;
    const unsigned char *ref; uint32_t reflen;
    uint32_t beg, end, W;
    const uint32_t *chunk_bounds; uint32_t n_chunks;       // mbias only
    md_call *calls; unsigned long long cap; uint2 *dir;    // extract: output records + per-window directory (offset,count)
    uint32_t *counters;
    uint32_t *hist; int32_t *lens;                         // mbias
    uint32_t st_seq, st_qual;                              // per-warp staging capacities in bytes (count_warp)
    uint32_t wide;                                         // 1: 32-bit window counters; 0: 16-bit pairs packed in one word (overflow -> rerun wide)
    BedView bed;
    uint32_t split; uint32_t *acc; uint32_t *done;         // deep tiles: `split` CTAs share a window's alignments; their counters meet in acc[] (32-bit layout per window), the last one to arrive (done[]) writes the calls
    uint32_t ablate;                                       // timing experiments only (MD_ABLATE): 1 skip evaluation, 2 skip generation, 4 skip staging copies
    uint32_t okmask;                                       // 2-/4-bit phred tiles: bit c set when code c decodes to a phred >= minPhred
    uint16_t *sites;                                       // GEN 2: per CTA 2 * W entries of scratch in HBM (L1-resident): the window's kept C-sites, then its kept G-sites, ascending
};

#define MB_SM_Q 256   // mbias: query positions < this are histogrammed in shared memory

// phred gate (common.c:127) on a staged phred column without decoding the value: packed codes are looked up in okmask
__device__ __forceinline__ bool staged_pass(const CountArgs &A, const unsigned char *p, int q) {
    if (A.R.qbits == 8u) return (int) p[q] >= A.P.minPhred;
    if (A.R.qbits == 2u) return (A.okmask >> ((p[q >> 2] >> ((q & 3) << 1)) & 3u)) & 1u;
    return (A.okmask >> ((p[q >> 1] >> ((q & 1) << 2)) & 15u)) & 1u;
}

// K4.  MODE 0: extract, 1: extract with variant filter, 2: mbias.
//
// Read-level state shared by the two per-read code paths
struct ReadCtx {
    int strand, rd2; bool wantG;
    int lo, hi;                      // kept query range after trims
    uint32_t soff, qoff;
    // mate (overlap merge); mi < 0 when there is nothing to merge
    int mi, mpos, mend, mlo, mhi; uint32_t mk0, mk1, msoff, mqoff; bool is_a, mate_simple; int mq0;
};

__device__ __forceinline__ void load_mate(const CountArgs &A, uint32_t i, int mi, ReadCtx &rc) {
    const DevReads &R = A.R;
    rc.mi = mi;
    if (mi < 0) return;
    rc.mpos = R.pos[mi]; rc.mend = A.rend[mi];
    const int ms = INFO_STRAND(A.info[mi]);
    // cust_tweak_overlap_quality bails out when the strands differ in parity (overlaps.c:65); disjoint spans have nothing to merge
    if (((rc.strand - ms) & 1) || !(R.pos[i] < rc.mend && rc.mpos < A.rend[i])) { rc.mi = -1; return; }
    rc.mk0 = R.cigar_off[mi]; rc.mk1 = R.cigar_off[mi + 1];
    rc.msoff = R.seq_off[mi]; rc.mqoff = R.qual_off[mi];
    dev_trim(A.P, ms, R.flag[mi], (int) R.l_qseq[mi], rc.mlo, rc.mhi);
    rc.is_a = (uint32_t) mi > i;                               // first in file order is `a` (overlaps.c:129-135)
    // a mate whose CIGAR is a single match op maps reference -> query by subtraction
    rc.mate_simple = false; rc.mq0 = 0;
    if (rc.mk1 - rc.mk0 == 1) { uint32_t op = __ldg(R.cigar + rc.mk0) & 15u; rc.mate_simple = (op == 0 || op == 7 || op == 8); }
}

// A base that passed the phred gate on a kept-context column: call (common.c:118-134) or variant evidence (extract.c:225-239).
// ovf: the window holds enough alignments for a packed 16-bit counter to wrap, so the wrap test is needed.
template <int MODE>
__device__ __forceinline__ void eval_plain(const CountArgs &A, int strand, int rd2, bool wantG, uint32_t *cnt, uint32_t W, uint32_t o, int qi, unsigned b, bool siteG, bool ovf) {
    if (siteG == wantG) {
        // OT/CTOT: C (2) methylated, T (8) unmethylated; OB/CTOB: G (4) methylated, A (1) unmethylated (common.c:129-132)
        const int rv = (b == (wantG ? 4u : 2u)) ? 1 : ((b == (wantG ? 1u : 8u)) ? -1 : 0);
        if (rv) {
            if (MODE == 2) {
                const int s1 = strand - 1;
                // the per-strand length `l` (MBias.c:212: largest query position with a call, plus one) is read off the finished
                // histogram by md_mbias_hist; a global atomicMax per call here serialised the whole kernel on four addresses
                if (qi < MB_SM_Q) atomicAdd(cnt + (((s1 * 2 + rd2) * MB_SM_Q + qi) * 2 + (rv < 0 ? 1 : 0)), 1u);
                else if (qi < MD_MBIAS_MAXLEN) atomicAdd(A.hist + ((((size_t) s1 * 2 + rd2) * MD_MBIAS_MAXLEN + qi) * 2 + (rv < 0 ? 1 : 0)), 1u);
            } else if (A.wide) atomicAdd(cnt + (rv > 0 ? o : W + o), 1u);
            else {                                                // meth in the low half, unmeth in the high half of one word
                const uint32_t old = atomicAdd(cnt + o, rv > 0 ? 1u : 0x10000u);
                if (ovf && (rv > 0 ? (old & 0xffffu) : (old >> 16)) == 0xffffu) atomicExch(A.counters + C_OVERFLOW, 3u);   // a 16-bit field wrapped: the host reruns the tile wide
            }
        }
    } else if (MODE == 1) {                                     // isVariant, extract.c:225-239
        const bool var = wantG ? (b != 2u && b != 15u) : (b != 4u && b != 15u);
        if (A.wide) { atomicAdd(cnt + 2 * W + o, 1u); if (var) atomicAdd(cnt + 3 * W + o, 1u); }
        else {                                                    // nOff low half, nVariant high half
            const uint32_t old = atomicAdd(cnt + W + o, var ? 0x10001u : 1u);
            if (ovf && (old & 0xffffu) == 0xffffu) atomicExch(A.counters + C_OVERFLOW, 3u);
        }
    }
}

// One base that sits on a kept-context column: overlap merge, phred gate, call / variant evidence.
// b / ql are the read's own base and phred AFTER trimming.  (overlaps.c:81-114, common.c:118-134, extract.c:225-239)
template <int MODE, bool MATE = true>
__device__ __forceinline__ void eval_hit(const CountArgs &A, const ReadCtx &rc, uint32_t *cnt, uint32_t W, int w0i, int rp, int qi, unsigned b, unsigned ql, bool siteG) {
    const DevReads &R = A.R;
    if (MATE && rc.mi >= 0 && rp >= rc.mpos && rp < rc.mend) {
        const int mq = rc.mate_simple ? (rp - rc.mpos) : dev_qpos_at(R.cigar, rc.mk0, rc.mk1, rc.mpos, rp);
        if (mq >= 0) {                                          // aligned in both mates
            unsigned mb = 15u, mql = 0u;
            if (mq >= rc.mlo && mq < rc.mhi) { mb = dev_base(R.seq, rc.msoff, mq); mql = dev_qual(R, rc.mqoff, mq); }
            const unsigned qa = rc.is_a ? ql : mql, qb = rc.is_a ? mql : ql;
            unsigned na, nb;
            if (b != mb) {                                      // a is tested first (overlaps.c:91-100)
                const unsigned ba = rc.is_a ? b : mb, bb = rc.is_a ? mb : b;
                if (qa > qb && ba != 15u) { na = qa - qb; nb = 0; }
                else if (qb > qa && bb != 15u) { nb = qb - qa; na = 0; }
                else { na = 0; nb = 0; }
            } else if (qa > qb) { na = A.P.boost[qa]; nb = 0; }
            else { nb = A.P.boost[qb]; na = 0; }                // ties favour b (overlaps.c:102-108)
            ql = rc.is_a ? na : nb;
        }
    }
    if ((int) ql < A.P.minPhred) return;                       // common.c:127 / extract.c:229
    eval_plain<MODE>(A, rc.strand, rc.rd2, rc.wantG, cnt, W, (uint32_t)(rp - w0i), qi, b, siteG, true);
}

// General path: any CIGAR.  The whole warp walks one alignment, lanes stride over the bases of each match op.
template <int MODE, class Ctx>
__device__ __forceinline__ void slow_read(const CountArgs &A, uint32_t i, unsigned inf, long long w0, long long own1, Ctx ctx, uint32_t *cnt, int lane) {
    const DevReads &R = A.R;
    ReadCtx rc;
    rc.strand = INFO_STRAND(inf);
    const unsigned f = R.flag[i];
    rc.rd2 = (f & 0x80u) ? 1 : 0; rc.wantG = !(rc.strand & 1);
    dev_trim(A.P, rc.strand, f, (int) R.l_qseq[i], rc.lo, rc.hi);
    rc.soff = R.seq_off[i]; rc.qoff = R.qual_off[i];
    load_mate(A, i, (MODE == 2) ? -1 : A.mate[i], rc);
    int p = R.pos[i], q = 0;
    for (uint32_t k = R.cigar_off[i], k1 = R.cigar_off[i + 1]; k < k1; ++k) {
        const uint32_t c = __ldg(R.cigar + k), op = c & 15u; const int len = (int)(c >> 4);
        if (op == 0 || op == 7 || op == 8) {
            const int j0 = (int) max(0ll, w0 - (long long) p), j1 = (int) min((long long) len, own1 - (long long) p);
            for (int j = j0 + lane; j < j1; j += 32) {
                const int rp = p + j, qi = q + j;
                const unsigned cx = ctx(rp - (int) w0, rc.wantG);
                if (!cx) continue;
                const bool siteG = (cx & 4u) != 0;
                if (MODE != 1 && siteG != rc.wantG) continue;   // wrong-strand columns only matter to the variant filter
                unsigned b = 15u, ql = 0u;
                if (qi >= rc.lo && qi < rc.hi) { b = dev_base(R.seq, rc.soff, qi); ql = dev_qual(R, rc.qoff, qi); }
                eval_hit<MODE>(A, rc, cnt, A.W, (int) w0, rp, qi, b, ql, siteG);
            }
            p += len; q += len;
        } else if (op == 1 || op == 4) q += len;
        else if (op == 2 || op == 3) p += len;
    }
}

// ------------------------------------------------------------------------------------------------
// async-copy primitives (sm_90+ / sm_100a): mbarrier + cp.async.bulk, i.e. the TMA engine in its 1-D "bulk" form
// (SASS: UBLKCP for the copy, SYNCS for the barrier)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}


// ------------------------------------------------------------------------------------------------
// K4w — warp-streaming count kernel (default).  Same contract as count_kernel / count_stream.
//
//   * one CTA (8 warps) per window of 4096 reference positions, two CTAs resident per SM;
//   * the window's alignments are consumed in batches of 32, handed out dynamically to the CTA's warps
//     (shared-memory ticket), so there is no CTA-wide barrier in the streaming phase and no tail imbalance;
//   * each warp owns a staging buffer and an mbarrier: lane 0 streams the batch's packed bases and phreds into
//     shared memory with two cp.async.bulk copies (TMA) while all lanes fetch the scalars of "their" alignment;
//   * GENERATE: every lane walks its alignment's CIGAR against the C-/G-site bitmaps of the window and emits one
//     candidate base per iteration; a warp ballot compacts the candidates into a shared-memory queue;
//   * EVALUATE: as soon as 32 candidates are queued, every lane takes one — all 32 lanes busy — and runs trims,
//     overlap merge, phred gate, call / variant evidence and the shared-memory atomicAdd.
//
// This is the "warp-ballot branch flattening" of the north star: the divergent part (how many cytosines an
// alignment covers) is reduced to a cheap iterator, the expensive part runs converged.
#define WS_SEQ_BYTES 2560          // default per-warp staging: 32 alignments x 76 B (150-mers) + alignment slack
#define WS_QUAL_BYTES 5120         // upper bound for the phred staging (32 x 152 B + slack for plain bytes); 2-/4-bit tiles use less
#define WS_CTX_WORDS 10
#define WS_QUEUE 448               // candidate queue entries per warp: plain candidates grow from the bottom, in-mate-span ones from the top
#define WS_EV 2                    // candidates per lane per evaluate round
#define WS_WARPS 8

struct WarpLayout { uint32_t off_bm, off_cnt, off_warp, warp_stride, off_seq, off_qual, off_ctx, off_queue, total; };
__host__ __device__ inline WarpLayout warp_layout(uint32_t W, int mode, uint32_t wide, uint32_t st_seq, uint32_t st_qual) {
    WarpLayout L; const uint32_t NW = W >> 5;
    auto up = [](uint32_t x) { return (x + 127u) & ~127u; };
    L.off_bm = 128;                                              // [0,128): 8 mbarriers + ticket counter
    L.off_cnt = up(L.off_bm + (mode == 1 ? 6u : 4u) * (NW + 2u) * 4u);   // bmC, bmG, bmT0, bmT1 (+ bmCo, bmGo with the variant filter)
    const uint32_t ncnt = mode == 2 ? 4u * 2u * MB_SM_Q * 2u : ((mode == 1 ? 4u * W : 2u * W) >> (wide ? 0 : 1));
    L.off_warp = up(L.off_cnt + ncnt * 4u);
    L.off_seq = 0; L.off_qual = up(st_seq + 32u); L.off_ctx = L.off_qual + up(st_qual + 32u);
    L.off_queue = L.off_ctx + WS_CTX_WORDS * 32u * 4u;
    L.warp_stride = up(L.off_queue + WS_QUEUE * 4u);
    L.total = L.off_warp + WS_WARPS * L.warp_stride;
    return L;
}

// per-alignment context words in shared memory
//  0: staged seq byte offset | staged qual byte offset << 16
//  1: flags: bit0 wantG, bit1 rd2, bit2 has mate, bit3 is_a, bit4 mate_simple, bits8-10 strand
//  2: mpos   3: mend   4: msoff   5: mqoff   6: mlo | mhi << 16   7: mk0   8: mk1
//  9: all a plain candidate needs: staged seq word offset | staged qual 8-byte offset << 10 | wantG << 20 | rd2 << 21 | strand << 22
template <int MODE, int GEN>
__global__ void __launch_bounds__(WS_WARPS * 32, 3) count_warp(CountArgs A) {
    constexpr int EV = WS_EV;
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t W = A.W, NW = W >> 5;
    const WarpLayout SL = warp_layout(W, MODE, A.wide, A.st_seq, A.st_qual);
    uint64_t *bars = (uint64_t *) smem;                                   // [8]
    uint32_t *ticket = (uint32_t *)(smem + 64);
    // site bitmaps of the window (bit = position; word w of the window lives at index w + 1, words 0 and NW + 1 are zero pads):
    //   bmC / bmG   kept-context C / G columns where a call may be made (a '-' / '+' BED region with --keepStrand blanks them)
    //   bmT0 / bmT1 column is CpG / CHG (neither: CHH)
    //   bmCo / bmGo (variant filter only) C / G columns where opposite-strand evidence is collected; they differ from bmC / bmG
    //               only inside strand-specific BED regions (extract.c:425 skips the read before either use)
    uint32_t *bmC = (uint32_t *)(smem + SL.off_bm), *bmG = bmC + NW + 2, *bmT0 = bmG + NW + 2, *bmT1 = bmT0 + NW + 2;
    uint32_t *bmCo = (MODE == 1) ? bmT1 + NW + 2 : bmC, *bmGo = (MODE == 1) ? bmCo + NW + 2 : bmG;
    uint32_t *cnt = (uint32_t *)(smem + SL.off_cnt);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char *wbase = smem + SL.off_warp + (size_t) warp * SL.warp_stride;
    unsigned char *sseq = wbase + SL.off_seq, *squal = wbase + SL.off_qual;
    uint32_t *rctx = (uint32_t *)(wbase + SL.off_ctx), *queue = (uint32_t *)(wbase + SL.off_queue);
    unsigned char *refw = smem + SL.off_warp;                             // prologue only: overlays the warps' staging area
    const uint32_t S = A.split, w = blockIdx.x / S, part = blockIdx.x - w * S;
    const long long w0 = (long long) A.beg + (long long) w * W;
    const long long own1 = min((long long) A.end, w0 + (long long) W);
    const int own = (int)(own1 - w0), w0i = (int) w0;
    const DevReads &R = A.R;
    __shared__ uint32_t s_rr[2];

    // ---- prologue: barriers, reference window, site bitmaps --------------------------
    if (tid == 0) { for (int kk = 0; kk < WS_WARPS; ++kk) mbar_init(bars + kk, 1); *ticket = 0; mbar_fence_init(); }
    // the window's alignment range: warp 0 finds the first alignment that can reach w0 (a read starting at or before w0 - maxspan
    // cannot), warp 1 the first one starting at or beyond the window end — replaces a separate launch with two 32-ary searches
    if (warp == 0) { const uint32_t r0 = warp_upper_bound(R.pos, R.n, w0 - (long long) A.counters[C_MAXSPAN], lane); if (lane == 0) s_rr[0] = r0; }
    else if (warp == 1) { const uint32_t r1 = warp_upper_bound(R.pos, R.n, w0 + (long long) W - 1, lane); if (lane == 0) s_rr[1] = r1; }
    for (uint32_t t = tid; t < W + 4; t += WS_WARPS * 32) {
        long long p = w0 - 2 + t;
        refw[t] = (p >= 0 && p < (long long) A.reflen) ? __ldg(A.ref + p) : (unsigned char) 'N';
    }
    const uint32_t ncnt = (MODE == 2) ? 4u * 2u * MB_SM_Q * 2u : ((MODE == 1 ? 4u * W : 2u * W) >> (A.wide ? 0 : 1));
    for (uint32_t t = tid; t < ncnt; t += WS_WARPS * 32) cnt[t] = 0;
    if (tid < (MODE == 1 ? 12 : 8)) { uint32_t *g = bmC + (tid >> 1) * (NW + 2); g[(tid & 1) ? NW + 1 : 0] = 0; }
    __syncthreads();
    unsigned site16C = 0, site16G = 0;                                     // this thread's 16 positions: kept C / G sites (GEN 2 builds the site lists from them)
    {
        const uint32_t t16 = 16u * tid;
        unsigned sC = 0, sG = 0, t0 = 0, t1 = 0;
        if (MODE == 2) {
            for (int j = 0; j < 16; ++j) {                                // MBias.c:147,170-178: context inside the chunk's own window
                long long p = w0 + t16 + j; unsigned c = 0;
                if (p < own1 && A.n_chunks) {
                    uint32_t lo = 0, hi = A.n_chunks;
                    while (lo + 1 < hi) { uint32_t mid = (lo + hi) >> 1; if ((long long) A.chunk_bounds[mid] <= p) lo = mid; else hi = mid; }
                    if (p >= (long long) A.chunk_bounds[lo] && p < (long long) A.chunk_bounds[lo + 1]) {
                        long long cs = A.chunk_bounds[lo], ce = A.chunk_bounds[lo + 1];
                        long long last = ce < (long long) A.reflen ? ce : (long long) A.reflen - 1;
                        auto get = [&](long long x) -> unsigned char { long long o = x - (w0 - 2); return (o >= 0 && o < (long long) W + 4) ? refw[o] : __ldg(A.ref + x); };
                        c = dev_context(get, p, cs, last + 1);
                    }
                }
                if (c && !((A.P.keepMask >> ((c & 3) - 1)) & 1)) c = 0;
                if (c) { if (c & 4u) sG |= 1u << j; else sC |= 1u << j; if ((c & 3u) == 1u) t0 |= 1u << j; else if ((c & 3u) == 2u) t1 |= 1u << j; }
            }
        } else {
            unsigned C = 0, G = 0;
            #pragma unroll
            for (int j = 0; j < 20; ++j) { unsigned ch = refw[t16 + j] | 0x20u; C |= (ch == 'c' ? 1u : 0u) << j; G |= (ch == 'g' ? 1u : 0u) << j; }
            const unsigned cpgC = C & (G >> 1), cpgG = G & (C << 1);
            const unsigned chgC = C & ~cpgC & (G >> 2), chgG = G & ~cpgG & (C << 2);
            const unsigned chhC = C & ~cpgC & ~chgC, chhG = G & ~cpgG & ~chgG;
            const unsigned k0 = (A.P.keepMask & 1) ? ~0u : 0u, k1 = (A.P.keepMask & 2) ? ~0u : 0u, k2 = (A.P.keepMask & 4) ? ~0u : 0u;
            int nown = own - (int) t16; nown = nown < 0 ? 0 : (nown > 16 ? 16 : nown);
            const unsigned ownm = nown >= 16 ? 0xffffu : ((1u << nown) - 1u);
            sC = (((cpgC & k0) | (chgC & k1) | (chhC & k2)) >> 2) & ownm;
            sG = (((cpgG & k0) | (chgG & k1) | (chhG & k2)) >> 2) & ownm;
            t0 = (((cpgC | cpgG) & k0) >> 2) & ownm;
            t1 = (((chgC | chgG) & k1) >> 2) & ownm;
        }
        // -l: positions outside every BED region drop out; with --keepStrand a '+' region keeps only OT/CTOT reads (which call at C
        // columns and give evidence at G columns), a '-' region only OB/CTOB reads (extract.c:402-405,425; bed.c:46-63)
        unsigned oC = sC, oG = sG;                                          // evidence columns
        if (A.bed.on) {
            unsigned aO = 0, aE = 0;
            const long long p0 = w0 + t16;
            if (p0 < own1) {
                uint32_t j = bed_first_beyond(A.bed, (uint32_t) p0);
                for (int jj = 0; jj < 16; ++jj) {
                    const uint32_t pp = (uint32_t)(p0 + jj);
                    while (j < A.bed.n && __ldg(A.bed.pmax + j) <= pp) ++j;
                    if (j < A.bed.n && __ldg(A.bed.start + j) <= pp) { const uint32_t sd = __ldg(A.bed.strand + j); if (sd != 2u) aO |= 1u << jj; if (sd != 1u) aE |= 1u << jj; }
                }
            }
            oC = sC & aE; oG = sG & aO; sC &= aO; sG &= aE; t0 &= aO | aE; t1 &= aO | aE;
        }
        site16C = sC; site16G = sG;
        const unsigned pC = __shfl_down_sync(0xffffffffu, sC, 1), pG = __shfl_down_sync(0xffffffffu, sG, 1), p0 = __shfl_down_sync(0xffffffffu, t0, 1), p1 = __shfl_down_sync(0xffffffffu, t1, 1);
        const unsigned pCo = __shfl_down_sync(0xffffffffu, oC, 1), pGo = __shfl_down_sync(0xffffffffu, oG, 1);
        if (!(tid & 1)) {
            const uint32_t wi = (t16 >> 5) + 1;
            bmC[wi] = sC | (pC << 16); bmG[wi] = sG | (pG << 16); bmT0[wi] = t0 | (p0 << 16); bmT1[wi] = t1 | (p1 << 16);
            if (MODE == 1) { bmCo[wi] = oC | (pCo << 16); bmGo[wi] = oG | (pGo << 16); }
        }
    }
    // GEN 2: the window's kept sites as sorted lists (C-sites at sites[0..nC), G-sites at sites[W..W+nG), window-relative) and, per
    // bitmap word, the number of sites below it: rank(x) = prefix[x >> 5] + popc(word & below(x)) turns "the kept sites of a
    // reference interval" into an index range of the list, and the list turns an index back into a position.
    __shared__ uint16_t s_prefC[(4096 >> 5) + 1], s_prefG[(4096 >> 5) + 1];
    __shared__ uint32_t s_scan[WS_WARPS];
    uint16_t *siteC = A.sites ? A.sites + (size_t) blockIdx.x * 2u * W : nullptr, *siteG = siteC ? siteC + W : nullptr;
    if (GEN == 2) {
        const uint32_t v = (uint32_t) __popc(site16C) | ((uint32_t) __popc(site16G) << 16);
        uint32_t inc = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_scan[warp] = inc;
        __syncthreads();
        uint32_t base = 0;
        for (int k = 0; k < warp; ++k) base += s_scan[k];
        const uint32_t ex = base + inc - v;
        if (!(tid & 1)) { s_prefC[tid >> 1] = (uint16_t)(ex & 0xffffu); s_prefG[tid >> 1] = (uint16_t)(ex >> 16); }
        if (tid == WS_WARPS * 32 - 1) { s_prefC[NW] = (uint16_t)((ex + v) & 0xffffu); s_prefG[NW] = (uint16_t)((ex + v) >> 16); }
        uint32_t oc = ex & 0xffffu, og = ex >> 16;
        for (unsigned m = site16C; m; m &= m - 1u) siteC[oc++] = (uint16_t)(16u * tid + (uint32_t)(__ffs(m) - 1));
        for (unsigned m = site16G; m; m &= m - 1u) siteG[og++] = (uint16_t)(16u * tid + (uint32_t)(__ffs(m) - 1));
    }
    __syncthreads();                                                      // bitmaps (and site lists) complete; refw (overlay) no longer needed
    uint2 rr = make_uint2(s_rr[0], max(s_rr[0], s_rr[1]));
    if (S > 1) {                                                          // this CTA's share of the window's alignments
        const unsigned long long len = rr.y - rr.x;
        const uint32_t lo = rr.x + (uint32_t)(len * part / S), hi = rr.x + (uint32_t)(len * (part + 1) / S);
        rr = make_uint2(lo, hi);
    }

    // ---- streaming phase: every warp on its own -----------------------------------------------------------
    const uint32_t maxlq = A.counters[C_MAXLQ];
    const uint32_t seqb_max = ((((maxlq + 1u) >> 1) + 3u) >> 2) * 4u, qualb_max = qual_words_of(R, maxlq) * 8u;
    uint32_t nb = 32;
    if (seqb_max) nb = min(nb, (A.st_seq - 32u) / seqb_max);
    if (qualb_max) nb = min(nb, (A.st_qual - 32u) / qualb_max);
    if (maxlq >= (1u << 14)) nb = 0;                                      // query index must fit the queue entry
    // context code of a window position for a read of the given orientation: 0 = nothing to do there, else type (1 CpG, 2 CHG, 3 CHH) | 4 if it is a G column
    auto ctx_code = [&](int rel, bool wantG) -> unsigned {
        const uint32_t wi = ((uint32_t) rel >> 5) + 1, b = 1u << (rel & 31);
        const bool own = (wantG ? bmG[wi] : bmC[wi]) & b, opp = (MODE == 1) && ((wantG ? bmCo[wi] : bmGo[wi]) & b);
        if (!own && !opp) return 0u;
        return ((bmT0[wi] & b) ? 1u : (bmT1[wi] & b) ? 2u : 3u) | ((own ? wantG : !wantG) ? 4u : 0u);
    };
    if (nb == 0) {
        for (uint32_t i = rr.x + warp; i < rr.y; i += WS_WARPS) {         // reads too long to stage: whole warp per alignment, from global memory
            const unsigned inf = A.info[i];
            if ((inf & INFO_ADMIT) && (long long) A.rend[i] > w0) slow_read<MODE>(A, i, inf, w0, own1, ctx_code, cnt, lane);
        }
    } else {
        uint64_t *bar = bars + warp;
        uint32_t phase = 0;
        int maxbits = 256;                                                // GEN 1: longest segment a lane takes per round (adapts to the site density)
        for (;;) {
            uint32_t b = 0;
            if (lane == 0) b = atomicAdd(ticket, 1u);
            b = __shfl_sync(0xffffffffu, b, 0);
            const uint64_t s64 = (uint64_t) rr.x + (uint64_t) b * nb;
            if (s64 >= rr.y) break;
            const uint32_t s = (uint32_t) s64, e = min(s + nb, rr.y), i = s + lane;
            // -- every lane: scalars of its alignment (the first and last lanes' offsets also delimit the batch's byte ranges)
            // Loads are issued level by level (everything addressed by i first, then what those values address), unconditionally
            // where a condition would only serialise two round trips to L2/HBM; the values are used further down.
            unsigned inf = 0; bool live = false;
            int pos = 0, lq = 0, mate = -1, rend_i = 0; unsigned f = 0; uint32_t soff = 0, qoff = 0, k0 = 0, k1 = 0, c0 = 0;
            int m_pos = 0, m_end = 0; unsigned m_inf = 0, m_flag = 0; uint32_t m_k0 = 0, m_k1 = 0, m_soff = 0, m_qoff = 0, m_lq = 0, m_c0 = 0;
            if (i < e) {
                soff = R.seq_off[i]; qoff = R.qual_off[i]; lq = (int) R.l_qseq[i];
                inf = A.info[i]; rend_i = A.rend[i];
                pos = R.pos[i]; f = R.flag[i];
                k0 = R.cigar_off[i]; k1 = R.cigar_off[i + 1];
                mate = (MODE == 2) ? -1 : A.mate[i];
                live = (inf & INFO_ADMIT) && (long long) rend_i > w0;
                if (live) {
                    c0 = __ldg(R.cigar + k0);
                    if (mate >= 0) {
                        m_pos = R.pos[mate]; m_end = A.rend[mate]; m_inf = A.info[mate];
                        m_k0 = R.cigar_off[mate]; m_k1 = R.cigar_off[mate + 1];
                        m_soff = R.seq_off[mate]; m_qoff = R.qual_off[mate]; m_flag = R.flag[mate]; m_lq = R.l_qseq[mate];
                    }
                }
            }
            // -- lane 0 streams the batch's bases and phreds into the warp's staging buffer (TMA, completion on the warp's mbarrier)
            const int last = (int)(e - 1 - s);
            uint32_t sw0 = __shfl_sync(0xffffffffu, soff, 0) & ~3u, qw0 = __shfl_sync(0xffffffffu, qoff, 0) & ~1u;
            const uint32_t lql = (uint32_t) __shfl_sync(0xffffffffu, lq, last);
            uint32_t sw_end = min(__shfl_sync(0xffffffffu, soff, last) + ((((lql + 1u) >> 1) + 3u) >> 2), R.seq_words);
            uint32_t qw_end = min(__shfl_sync(0xffffffffu, qoff, last) + qual_words_of(R, lql), R.qual_words);
            uint32_t sbytes = sw_end > sw0 ? (sw_end - sw0) * 4u : 0u, qbytes = qw_end > qw0 ? (qw_end - qw0) * 8u : 0u;
            sbytes = min((sbytes + 15u) & ~15u, A.st_seq); qbytes = min((qbytes + 15u) & ~15u, A.st_qual);
            if (A.ablate & 4u) { sbytes = 0; qbytes = 0; }
            const uint32_t sw1 = sw0 + sbytes / 4u, qw1 = qw0 + qbytes / 8u;
            if (lane == 0) {
                mbar_arrive_expect_tx(bar, sbytes + qbytes);
                if (sbytes) bulk_copy_g2s(sseq, R.seq + sw0, sbytes, bar);
                if (qbytes) bulk_copy_g2s(squal, R.qual + qw0, qbytes, bar);
            }
            ReadCtx rc; rc.mi = -1; rc.lo = 0; rc.hi = 0; rc.strand = 1; rc.wantG = false; rc.rd2 = 0;
            bool staged = false;
            if (live) {
                rc.strand = INFO_STRAND(inf); rc.rd2 = (f & 0x80u) ? 1 : 0; rc.wantG = !(rc.strand & 1);
                dev_trim(A.P, rc.strand, f, lq, rc.lo, rc.hi);
                rc.soff = soff; rc.qoff = qoff;
                // mate descriptor (overlap merge): cust_tweak_overlap_quality bails out when the strands differ in parity (overlaps.c:65);
                // disjoint spans have nothing to merge
                rc.mi = mate;
                if (mate >= 0) {
                    const int ms = INFO_STRAND(m_inf);
                    rc.mpos = m_pos; rc.mend = m_end;
                    if (((rc.strand - ms) & 1) || !(pos < m_end && m_pos < rend_i)) rc.mi = -1;
                    else {
                        rc.mk0 = m_k0; rc.mk1 = m_k1; rc.msoff = m_soff; rc.mqoff = m_qoff;
                        dev_trim(A.P, ms, m_flag, (int) m_lq, rc.mlo, rc.mhi);
                        rc.is_a = (uint32_t) mate > i;                    // first in file order is `a` (overlaps.c:129-135)
                        rc.mate_simple = false; rc.mq0 = 0;                // a mate whose CIGAR is a single match op maps reference -> query by subtraction
                        if (m_k1 - m_k0 == 1) { m_c0 = __ldg(R.cigar + m_k0) & 15u; rc.mate_simple = (m_c0 == 0 || m_c0 == 7 || m_c0 == 8); }
                    }
                }
                const uint32_t sw_need = ((((uint32_t) lq + 1u) >> 1) + 3u) >> 2, qw_need = qual_words_of(R, (uint32_t) lq);
                staged = soff >= sw0 && soff + sw_need <= sw1 && qoff >= qw0 && qoff + qw_need <= qw1;
                uint32_t *cx = rctx + WS_CTX_WORDS * lane;
                cx[0] = ((soff - sw0) * 4u) | (((qoff - qw0) * 8u) << 16);
                cx[9] = ((soff - sw0) & 1023u) | (((qoff - qw0) & 1023u) << 10) | (rc.wantG ? 1u << 20 : 0u) | (rc.rd2 ? 1u << 21 : 0u) | ((unsigned) rc.strand << 22);
                cx[1] = (rc.wantG ? 1u : 0u) | (rc.rd2 ? 2u : 0u) | (rc.mi >= 0 ? 4u : 0u) | (rc.is_a ? 8u : 0u) | (rc.mate_simple ? 16u : 0u) | ((unsigned) rc.strand << 8);
                if (rc.mi >= 0) { cx[2] = (uint32_t) rc.mpos; cx[3] = (uint32_t) rc.mend; cx[4] = rc.msoff; cx[5] = rc.mqoff; cx[6] = (uint32_t) rc.mlo | ((uint32_t) rc.mhi << 16); cx[7] = rc.mk0; cx[8] = rc.mk1; }
            }
            // alignments that cannot use the queue (not inside the staged range, trims beyond 16 bits) go the direct way below
            const bool direct = live && (!staged || (rc.mi >= 0 && rc.mhi > 0xffff));
            { uint32_t spins = 0; while (!mbar_try_wait(bar, phase)) { if (++spins > (1u << 26)) { atomicExch(A.counters + C_OVERFLOW, 2u); break; } } }
            phase ^= 1u;
            __syncwarp();                                              // context words visible to the whole warp

            // -- GENERATE / EVALUATE
            // one queued candidate -> call / evidence.  MATE candidates carry the mate descriptor of their alignment.
            const bool ovf = rr.y - rr.x >= 0xffffu;                   // only a window this deep can wrap a packed 16-bit counter
            auto eval_entry = [&](uint32_t en, bool with_mate) {
                const int src = en & 31u, rel = (en >> 5) & 0xfffu, qi = (en >> 17) & 0x3fffu; const bool is_opp = en >> 31;
                const uint32_t *cx = rctx + WS_CTX_WORDS * src;
                if (!with_mate) {
                    const uint32_t d = cx[9];
                    const bool wg = (d >> 20) & 1u;
                    if (!staged_pass(A, squal + ((d >> 10) & 1023u) * 8u, qi)) return;
                    const unsigned byte = sseq[(d & 1023u) * 4u + (qi >> 1)];
                    eval_plain<MODE>(A, (int)(d >> 22), (int)((d >> 21) & 1u), wg, cnt, W, (uint32_t) rel, qi, (qi & 1) ? (byte & 0xfu) : (byte >> 4), is_opp ? !wg : wg, ovf);
                    return;
                }
                const uint32_t c0w = cx[0], c1w = cx[1];
                ReadCtx hc;
                hc.wantG = c1w & 1u; hc.rd2 = (c1w >> 1) & 1u; hc.strand = (c1w >> 8) & 7u; hc.mi = 0;
                hc.is_a = (c1w & 8u) != 0; hc.mate_simple = (c1w & 16u) != 0;
                hc.mpos = (int) cx[2]; hc.mend = (int) cx[3]; hc.msoff = cx[4]; hc.mqoff = cx[5]; hc.mlo = (int)(cx[6] & 0xffffu); hc.mhi = (int)(cx[6] >> 16); hc.mk0 = cx[7]; hc.mk1 = cx[8];
                const unsigned byte = sseq[(c0w & 0xffffu) + (qi >> 1)];
                const unsigned bb = (qi & 1) ? (byte & 0xfu) : (byte >> 4), ql = staged_qual(R, squal + (c0w >> 16), qi);
                eval_hit<MODE, true>(A, hc, cnt, W, w0i, w0i + rel, qi, bb, ql, is_opp ? !hc.wantG : hc.wantG);
            };
            // iterator over this lane's candidate bases: current match op [ra,rb) (clipped, window-relative), bitmap word wi, remaining bits cur
            const uint32_t *own_bm = rc.wantG ? bmG : bmC, *opp_bm = rc.wantG ? bmCo : bmGo;
            uint32_t k = k0; int p = pos, q = 0, ra = 0, rb = 0, wi = 0; unsigned cur = 0, curo = 0; bool in_op = false;
            bool done = !(live && !direct);
            if constexpr (GEN == 1) {
                // Segment-at-a-time generator.  A round: every lane holds ONE segment of its current match op — a run of reference
                // positions that lies entirely outside or entirely inside the mate's span, at most `maxbits` long.  The lanes count the
                // kept sites of their segment in the window bitmaps (popc per word), the counts (plain | in-mate-span, packed in one
                // register) are prefix-summed across the warp, and every lane then writes its candidates to its own slice of the queue:
                // no ballot per candidate, and the divergent CIGAR walk runs once per op instead of once per base.
                // Plain candidates are stacked from the bottom of the queue, in-mate-span ones (global look-ups) from the top.
                int qrel = 0;                                              // query index = window-relative position + qrel inside the current op
                uint32_t qa = 0, qb2 = 0;
                const bool has_mate = rc.mi >= 0;
                const int mrel0 = has_mate ? rc.mpos - w0i : 0, mrel1 = has_mate ? rc.mend - w0i : 0;
                if (A.ablate & 2u) done = true;
                for (;;) {
                    while (!done && !in_op) {
                        if (k >= k1) { done = true; break; }
                        const uint32_t c = (k == k0) ? c0 : __ldg(R.cigar + k), op = c & 15u; const int len = (int)(c >> 4);
                        ++k;
                        if (op == 0 || op == 7 || op == 8) {
                            const int a = max(max(p, w0i), p + (rc.lo - q)), bnd = min(min(p + len, w0i + own), p + (rc.hi - q));
                            if (bnd > a) { ra = a - w0i; rb = bnd - w0i; qrel = q - (p - w0i); in_op = true; }
                            p += len; q += len;
                        } else if (op == 1 || op == 4) q += len;
                        else if (op == 2 || op == 3) p += len;
                    }
                    const bool all_done = !__any_sync(0xffffffffu, in_op);     // every lane ran out of match ops: the batch is exhausted
                    // this round's segment [ra, se)
                    int sb = ra; bool tB = false;
                    if (in_op) {
                        sb = rb;
                        if (has_mate) { if (ra < mrel0) sb = min(rb, mrel0); else if (ra < mrel1) { sb = min(rb, mrel1); tB = true; } }
                    }
                    const uint32_t room = WS_QUEUE - qa - qb2;
                    int se; uint32_t v, inc, tot;
                    for (;;) {
                        se = min(sb, ra + maxbits);
                        uint32_t n = 0;
                        if (se > ra) {
                            const int wa = ra >> 5, wb = (se - 1) >> 5;
                            for (int w = wa; w <= wb; ++w) {
                                unsigned m = own_bm[w + 1]; if (MODE == 1) m |= opp_bm[w + 1];
                                if (w == wa) m &= 0xffffffffu << (ra & 31);
                                if (w == wb) m &= 0xffffffffu >> (31 - ((se - 1) & 31));
                                n += (uint32_t) __popc(m);
                            }
                        }
                        v = tB ? (n << 16) : n;
                        inc = v;
                        #pragma unroll
                        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                        tot = __shfl_sync(0xffffffffu, inc, 31);
                        if ((tot & 0xffffu) + (tot >> 16) <= room || maxbits <= 8) break;
                        maxbits >>= 1;                                        // unusually site-dense: shorter segments (8 positions per lane always fit)
                    }
                    if (((tot & 0xffffu) + (tot >> 16)) * 4u < room && maxbits < 256) maxbits <<= 1;
                    if (se > ra) {
                        const uint32_t ex = inc - v;
                        uint32_t *dst = tB ? queue + (WS_QUEUE - 1u - (qb2 + (ex >> 16))) : queue + (qa + (ex & 0xffffu));
                        const int dir = tB ? -1 : 1;
                        const uint32_t base = (uint32_t) lane + ((uint32_t) qrel << 17);
                        const int wa = ra >> 5, wb = (se - 1) >> 5;
                        for (int w = wa; w <= wb; ++w) {
                            const unsigned mo = own_bm[w + 1]; unsigned m = mo; if (MODE == 1) m |= opp_bm[w + 1];
                            if (w == wa) m &= 0xffffffffu << (ra & 31);
                            if (w == wb) m &= 0xffffffffu >> (31 - ((se - 1) & 31));
                            const uint32_t entw = base + (uint32_t)(w << 5) * 0x20020u;
                            while (m) {
                                const int bit = __ffs(m) - 1; m &= m - 1u;
                                uint32_t ent = entw + (uint32_t) bit * 0x20020u;
                                if (MODE == 1) ent |= ((mo >> bit) & 1u) ? 0u : 0x80000000u;
                                *dst = ent; dst += dir;
                            }
                        }
                        ra = se; if (ra >= rb) in_op = false;
                    }
                    qa += tot & 0xffffu; qb2 += tot >> 16;
                    __syncwarp();
                    while (qa >= 32u * EV || (all_done && qa > 0u)) {
                        const uint32_t take = min(qa, 32u * EV);
                        #pragma unroll
                        for (int r = 0; r < EV; ++r) if ((uint32_t)(lane + 32 * r) < take && !(A.ablate & 1u)) eval_entry(queue[qa - take + lane + 32 * r], false);
                        qa -= take;
                        __syncwarp();
                    }
                    while (qb2 >= 32u * EV || (all_done && qb2 > 0u)) {
                        const uint32_t take = min(qb2, 32u * EV);
                        #pragma unroll
                        for (int r = 0; r < EV; ++r) if ((uint32_t)(lane + 32 * r) < take && !(A.ablate & 1u)) eval_entry(queue[WS_QUEUE - qb2 + lane + 32 * r], true);
                        qb2 -= take;
                        __syncwarp();
                    }
                    if (all_done) break;
                }
            } else {
                // GEN 2: list generator (MODE 0 and 2).  The kept sites of a lane's match op are an index range of its strand's site
                // list: two rank look-ups instead of a walk over bitmap words, and writing the candidates out is a counted loop
                // over list entries — every lane runs the same few iterations — instead of a find-first-set loop whose trip count
                // differs from lane to lane (that loop and the per-word counting were 42 % of the kernel's instructions at 17 of 32
                // lanes in round 1's generator, which stays as GEN 1 for the variant filter).  A match op is cut where the mate's
                // span begins and ends, in list-index space: phase 0 = sites before the mate's span, 1 = inside it, 2 = beyond it.
                const uint16_t *lst = rc.wantG ? siteG : siteC, *pref = rc.wantG ? s_prefG : s_prefC;
                auto rank = [&](int x) -> uint32_t {                    // kept sites of this lane's strand below window-relative position x, 0 <= x <= W
                    const uint32_t w = (uint32_t) x >> 5;
                    return (uint32_t) pref[w] + (uint32_t) __popc(own_bm[w + 1] & ((1u << (x & 31)) - 1u));
                };
                int qrel = 0;
                uint32_t qa = 0, qb2 = 0;
                const bool has_mate = rc.mi >= 0;
                const int mrel0 = has_mate ? rc.mpos - w0i : 0, mrel1 = has_mate ? rc.mend - w0i : 0;
                // cursors into the site list for the current match op: plain candidates [a1, b1) (before the mate's span) and [a2, b2)
                // (beyond it), in-mate-span candidates [m1, m2); a round takes from all three, so one round per op is the rule
                uint32_t a1 = 0, b1 = 0, m1 = 0, m2 = 0, a2 = 0, b2 = 0;
                if (A.ablate & 2u) done = true;
                for (;;) {
                    while (!done && a1 >= b1 && m1 >= m2 && a2 >= b2) {
                        if (k >= k1) { done = true; break; }
                        const uint32_t c = (k == k0) ? c0 : __ldg(R.cigar + k), op = c & 15u; const int len = (int)(c >> 4);
                        ++k;
                        if (op == 0 || op == 7 || op == 8) {
                            const int a = max(max(p, w0i), p + (rc.lo - q)), bnd = min(min(p + len, w0i + own), p + (rc.hi - q));
                            if (bnd > a) {
                                const int ra = a - w0i, rb = bnd - w0i;
                                qrel = q - (p - w0i);
                                a1 = rank(ra); b2 = rank(rb);
                                if (has_mate) { b1 = rank(min(max(mrel0, ra), rb)); a2 = rank(min(max(mrel1, ra), rb)); }
                                else { b1 = b2; a2 = b2; }
                                m1 = b1; m2 = a2;
                            }
                            p += len; q += len;
                        } else if (op == 1 || op == 4) q += len;
                        else if (op == 2 || op == 3) p += len;
                    }
                    const uint32_t lA1 = b1 - a1, lA = lA1 + (b2 - a2), lB = m2 - m1;
                    const bool busy = (lA | lB) != 0u;
                    const bool all_done = !__any_sync(0xffffffffu, busy);     // every lane ran out of match ops: the batch is exhausted
                    const uint32_t room = WS_QUEUE - qa - qb2, cap = room >> 5;   // the queues are drained below 64 entries each: cap >= 10
                    const uint32_t nB = min(lB, cap >> 1), nA = min(lA, cap - nB);
                    const uint32_t v = nA | (nB << 16);
                    uint32_t inc = v;
                    #pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                    const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
                    const uint32_t ex = inc - v;
                    const uint32_t base = (uint32_t) lane + ((uint32_t) qrel << 17);
                    if (nA) {
                        uint32_t *dst = queue + (qa + (ex & 0xffffu));
                        for (uint32_t r = 0; r < nA; ++r) { const uint32_t jj = r < lA1 ? a1 + r : a2 + (r - lA1); dst[r] = base + (uint32_t) lst[jj] * 0x20020u; }
                        const uint32_t t1 = min(nA, lA1); a1 += t1; a2 += nA - t1;
                    }
                    if (nB) {
                        uint32_t *dst = queue + (WS_QUEUE - 1u - (qb2 + (ex >> 16)));
                        for (uint32_t r = 0; r < nB; ++r) *(dst - r) = base + (uint32_t) lst[m1 + r] * 0x20020u;
                        m1 += nB;
                    }
                    qa += tot & 0xffffu; qb2 += tot >> 16;
                    __syncwarp();
                    while (qa >= 32u * EV || (all_done && qa > 0u)) {
                        const uint32_t take = min(qa, 32u * EV);
                        #pragma unroll
                        for (int r = 0; r < EV; ++r) if ((uint32_t)(lane + 32 * r) < take && !(A.ablate & 1u)) eval_entry(queue[qa - take + lane + 32 * r], false);
                        qa -= take;
                        __syncwarp();
                    }
                    while (qb2 >= 32u * EV || (all_done && qb2 > 0u)) {
                        const uint32_t take = min(qb2, 32u * EV);
                        #pragma unroll
                        for (int r = 0; r < EV; ++r) if ((uint32_t)(lane + 32 * r) < take && !(A.ablate & 1u)) eval_entry(queue[WS_QUEUE - qb2 + lane + 32 * r], true);
                        qb2 -= take;
                        __syncwarp();
                    }
                    if (all_done) break;
                }
            }
            // -- alignments outside the staged range: the lane walks it alone, from global memory
            if (direct) {
                int p2 = pos, q2 = 0;
                for (uint32_t kk = k0; kk < k1; ++kk) {
                    const uint32_t c = __ldg(R.cigar + kk), op = c & 15u; const int len = (int)(c >> 4);
                    if (op == 0 || op == 7 || op == 8) {
                        const int a = max(max(p2, w0i), p2 + (rc.lo - q2)), bnd = min(min(p2 + len, w0i + own), p2 + (rc.hi - q2));
                        for (int x = a; x < bnd; ++x) {
                            const unsigned cxc = ctx_code(x - w0i, rc.wantG);
                            if (!cxc) continue;
                            const bool siteG = (cxc & 4u) != 0;
                            if (MODE != 1 && siteG != rc.wantG) continue;
                            const int qi = q2 + (x - p2);
                            eval_hit<MODE>(A, rc, cnt, W, w0i, x, qi, dev_base(R.seq, soff, qi), dev_qual(R, qoff, qi), siteG);
                        }
                        p2 += len; q2 += len;
                    } else if (op == 1 || op == 4) q2 += len;
                    else if (op == 2 || op == 3) p2 += len;
                }
            }
            __syncwarp();                                              // staging buffer and context free for the next batch
        }
    }
    __syncthreads();

    // ---- epilogue (as count_stream) ---------------------------------------------------------------------
    if (MODE == 2) {
        for (uint32_t t = tid; t < 4u * 2u * MB_SM_Q * 2u; t += WS_WARPS * 32) {
            uint32_t v = cnt[t];
            if (v) { uint32_t sr = t / (MB_SM_Q * 2), rest = t % (MB_SM_Q * 2); atomicAdd(A.hist + (size_t) sr * MD_MBIAS_MAXLEN * 2 + rest, v); }
        }
        return;
    }
    __shared__ uint32_t s_warp_tot[WS_WARPS], s_base, s_last;
    // deep tiles: add this CTA's counters to the window's accumulators in HBM; whoever arrives last reads the sums back
    const uint32_t *acc = A.acc + (size_t) w * (MODE == 1 ? 4u : 2u) * W;
    if (S > 1) {
        uint32_t *accw = A.acc + (size_t) w * (MODE == 1 ? 4u : 2u) * W;
        const uint32_t nwords = (MODE == 1 ? 4u * W : 2u * W) >> (A.wide ? 0 : 1);
        for (uint32_t t = tid; t < nwords; t += WS_WARPS * 32) {
            const uint32_t v = cnt[t];
            if (!v) continue;
            if (A.wide) atomicAdd(accw + t, v);
            else {                                                          // word t < W: meth | unmeth << 16 of column t; word W + t: nOff | nVariant << 16
                const uint32_t col = t < W ? t : t - W, base0 = t < W ? 0u : 2u * W;
                if (v & 0xffffu) atomicAdd(accw + base0 + col, v & 0xffffu);
                if (v >> 16) atomicAdd(accw + base0 + W + col, v >> 16);
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(A.done + w, 1u) == S - 1u) ? 1u : 0u;
        __syncthreads();
        if (!s_last) return;
        __threadfence();
    }
    auto c_meth = [&](uint32_t t) -> uint32_t { return S > 1 ? __ldcg(acc + t) : (A.wide ? cnt[t] : (cnt[t] & 0xffffu)); };
    auto c_unmeth = [&](uint32_t t) -> uint32_t { return S > 1 ? __ldcg(acc + W + t) : (A.wide ? cnt[W + t] : (cnt[t] >> 16)); };
    const uint32_t t16 = 16u * tid;
    unsigned rep = 0;
    {
        const uint32_t wi = (t16 >> 5) + 1, sh = t16 & 31u;
        unsigned sites = ((bmC[wi] | bmG[wi] | bmCo[wi] | bmGo[wi]) >> sh) & 0xffffu;
        while (sites) {
            const int kbit = __ffs(sites) - 1; sites &= sites - 1;
            const uint32_t t = t16 + kbit;
            bool excl = false;
            if (MODE == 1) {
                const uint32_t noff = S > 1 ? __ldcg(acc + 2 * W + t) : (A.wide ? cnt[2 * W + t] : (cnt[W + t] & 0xffffu));
                const uint32_t nvar = S > 1 ? __ldcg(acc + 3 * W + t) : (A.wide ? cnt[3 * W + t] : (cnt[W + t] >> 16));
                excl = A.P.minOppositeDepth > 0 && noff >= (uint32_t) A.P.minOppositeDepth && ((double) nvar) / ((double) noff) >= A.P.maxVariantFrac;
            }
            if (excl) rep |= 0x10000u << kbit;
            if (excl || (c_meth(t) + c_unmeth(t))) rep |= 1u << kbit;
        }
    }
    const uint32_t mine = __popc(rep & 0xffffu);
    uint32_t incl = mine;
    for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    if (tid == 0) {
        uint32_t tot = 0;
        for (int kk = 0; kk < WS_WARPS; ++kk) { uint32_t v = s_warp_tot[kk]; s_warp_tot[kk] = tot; tot += v; }
        unsigned long long base = tot ? atomicAdd((unsigned long long *)(A.counters + C_NCALLS), (unsigned long long) tot) : 0ull;
        if (base + tot > A.cap) { atomicExch(A.counters + C_OVERFLOW, 1u); s_base = 0xffffffffu; }
        else s_base = (uint32_t) base;
        A.dir[w] = make_uint2((uint32_t) base, tot);
    }
    __syncthreads();
    if (s_base == 0xffffffffu) return;
    uint32_t o = s_base + s_warp_tot[warp] + (incl - mine);
    {
        const uint32_t wi = (t16 >> 5) + 1, sh = t16 & 31u;
        const unsigned g16 = ((bmG[wi] | bmGo[wi]) >> sh) & 0xffffu, a16 = (bmT0[wi] >> sh) & 0xffffu, b16 = (bmT1[wi] >> sh) & 0xffffu;
        unsigned r16 = rep & 0xffffu;
        while (r16) {
            const int kbit = __ffs(r16) - 1; r16 &= r16 - 1;
            const uint32_t t = t16 + kbit;
            md_call c; c.pos = (uint32_t)(w0 + t); c.nmeth = c_meth(t); c.nunmeth = c_unmeth(t);
            c.info = (((a16 >> kbit) & 1u) ? 0u : ((b16 >> kbit) & 1u) ? 1u : 2u) | (((g16 >> kbit) & 1u) ? 4u : 0u) | (((rep >> (16 + kbit)) & 1u) ? 8u : 0u);
            A.calls[o++] = c;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// perRead (perRead.c:37-94, 183-196): one thread per alignment replays processRead() — a sequential walk that cannot be
// turned into a pileup because of its quirks, all of which are observable and therefore kept:
//   * a base below minPhred is stepped over and the NEXT base is then looked at without a phred test of its own, under the
//     CIGAR op of the skipped base even when that op has just ended (perRead.c:59-63); when the skipped base was the read's
//     last one, the "next base" is the nibble behind the sequence: the pad nibble for odd lengths, the high nibble of the
//     first phred for even ones (BAM record layout, bam_seqi on index l_qseq);
//   * CpG context is looked up in the reference window of the chunk the alignment STARTS in: contig[localPos-2, localEnd+10000]
//     (perRead.c:176-181), so a C on the window's last base is no CpG and everything beyond it is no context at all;
//   * the only filters are -R, -F, -q (perRead.c:189-191) — no duplicate / NH / singleton logic; an alignment is reported by
//     the chunk holding its start (perRead.c:186-187);
//   * a strand of 0 (paired without read1/read2) is treated as a bottom strand: (strand & 1) == 0 (perRead.c:72).
__global__ void __launch_bounds__(128) per_read_kernel(DevReads R, KParams P, const unsigned char *ref, uint32_t reflen, uint32_t beg, uint32_t end, uint32_t chunk, md_read_meth *out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < R.n; i += gridDim.x * blockDim.x) {
        const unsigned f = R.flag[i];
        const long long pos = R.pos[i];
        md_read_meth res; res.nmeth = 0xffffffffu; res.nunmeth = 0;
        const bool rep = pos >= (long long) beg && pos < (long long) end && !(P.requireFlags && ((unsigned) P.requireFlags & f) != (unsigned) P.requireFlags) &&
                         !(P.ignoreFlags && ((unsigned) P.ignoreFlags & f) != 0u) && (int) R.mapq[i] >= P.minMapq;
        if (!rep) { out[i] = res; continue; }
        // the chunk this alignment starts in, and its reference window (perRead.c:118-137, 176-181)
        const unsigned long long lp = (unsigned long long) beg + ((unsigned long long)(pos - beg) / chunk) * chunk;
        const unsigned long long le = min(lp + (unsigned long long) chunk, (unsigned long long) end);
        const long long w0 = lp > 1 ? (long long) lp - 2 : 0, w1 = (long long) min(le + 10000ull, (unsigned long long) reflen - 1ull);   // inclusive
        const long long seqlen = w1 - w0 + 1;
        const int strand = dev_strand(f, R.aux[i]);
        const bool top = (strand & 1) == 1;
        const uint32_t lq = R.l_qseq[i], soff = R.seq_off[i], qoff = R.qual_off[i];
        const uint32_t k0 = R.cigar_off[i], k1 = R.cigar_off[i + 1];
        uint32_t rp = 0, k = k0, off = 0, nm = 0, nu = 0;
        unsigned long long mp = (unsigned long long) pos;
        while (rp < lq && k < k1) {
            uint32_t c = __ldg(R.cigar + k);
            if (off >= (c >> 4)) { off = 0; ++k; if (k >= k1) break; c = __ldg(R.cigar + k); }   // (the reference reads one op past the array here)
            const unsigned type = (0x3C1A7u >> ((c & 15u) << 1)) & 3u;       // bam_cigar_type: bit0 consumes query, bit1 consumes reference
            if (type & 2u) {
                if (type & 1u) {
                    if ((int) dev_qual(R, qoff, (int) rp) < P.minPhred) { ++mp; ++rp; ++off; }
                    const long long rel = (long long) mp - w0;
                    int dir = 0;
                    if (rel < seqlen) {
                        const unsigned char b0 = __ldg(ref + mp);
                        if (d_isC(b0)) { if (rel + 1 != seqlen && d_isG(__ldg(ref + mp + 1))) dir = 1; }
                        else if (d_isG(b0)) { if (rel != 0 && d_isC(__ldg(ref + mp - 1))) dir = -1; }
                    }
                    if (dir) {
                        unsigned base;
                        if (rp < lq) base = dev_base(R.seq, soff, (int) rp);
                        else if (lq & 1u) base = (__ldg(R.seq + soff + (lq >> 3)) >> (((lq >> 1) & 3u) << 3)) & 0xfu;   // pad nibble of the last sequence byte
                        else base = dev_qual(R, qoff, 0) >> 4;                                                          // first phred byte follows the sequence
                        if (dir == 1 && top) { if (base == 2u) ++nm; else if (base == 8u) ++nu; }
                        else if (dir == -1 && !top) { if (base == 4u) ++nm; else if (base == 1u) ++nu; }
                    }
                    ++mp; ++rp; ++off;
                } else { mp += c >> 4; ++k; off = 0; }
            } else if (type & 1u) { rp += c >> 4; ++k; off = 0; }
            else { ++k; off = 0; }
        }
        res.nmeth = nm; res.nunmeth = nu;
        out[i] = res;
    }
}

// K5/K6: put the per-window segments into position order on the device (exclusive scan of the directory
// counts by one CTA, then a segment copy), so the D2H transfer lands in the caller's buffer already sorted.
__global__ void __launch_bounds__(1024) dir_scan_kernel(uint2 *dir, uint32_t n_win, uint32_t *sorted_off) {
    __shared__ uint32_t s_part[32]; __shared__ uint32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_win; base += 1024) {
        const uint32_t i = base + tid;
        const uint32_t v = i < n_win ? dir[i].y : 0u;
        uint32_t incl = v;
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_part[warp] = incl;
        __syncthreads();
        if (warp == 0) { uint32_t p = s_part[lane], q = p; for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, q, o); if (lane >= o) q += t; } s_part[lane] = q - p; }
        __syncthreads();
        const uint32_t excl = s_carry + s_part[warp] + incl - v;
        if (i < n_win) sorted_off[i] = excl;
        __syncthreads();
        if (tid == 1023) s_carry = excl + v;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(128) gather_kernel(const md_call *raw, const uint2 *dir, const uint32_t *sorted_off, md_call *out) {
    const uint2 d = dir[blockIdx.x];
    const uint32_t o = sorted_off[blockIdx.x];
    const uint4 *src = (const uint4 *)(raw + d.x); uint4 *dst = (uint4 *)(out + o);
    for (uint32_t k = threadIdx.x; k < d.y; k += blockDim.x) dst[k] = src[k];
}

// ------------------------------------------------------------------------------------------------
// host side of the library
struct Contig { unsigned char *d_seq = nullptr; uint32_t len = 0; uint32_t *d_bounds = nullptr; uint32_t n_chunks = 0; uint32_t *d_bed = nullptr; uint32_t n_bed = 0;
                size_t cap_seq = 0, cap_bounds = 0, cap_bed = 0; };   // capacities of the pool blocks behind the three pointers

// cudaFree waits for EVERYTHING in flight on the device — during a run that is the next segment's inflate (10-15 ms), and a
// genome's worth of growing buffers and contig swaps added up to 0.4 - 1.6 s of a 2 - 3 s run (MD_TIMING, round 2).  Buffers
// that are outgrown or dropped while a context is alive therefore go to a graveyard that is emptied when a context is destroyed
// (their total stays below ~3x the final sizes: buffers grow by 1.5x), and contig-sized blocks are recycled (DevPool).
static std::mutex g_grave_m;
static std::vector<void *> g_grave;
static void dev_free_later(void *p) { if (!p) return; std::lock_guard<std::mutex> g(g_grave_m); g_grave.push_back(p); }
static void dev_free_graveyard() {                        // the caller has synchronised its streams
    std::vector<void *> dead;
    { std::lock_guard<std::mutex> g(g_grave_m); dead.swap(g_grave); }
    for (void *p : dead) cudaFree(p);
}

struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return 0;
        dev_free_later(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 2 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { set_err("cudaMalloc", e); return -100; }
        cap = want; return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Blocks handed back by md_drop_contig, reused by the next md_load_contig / md_set_* (contigs come largest first in the usual
// genome order, so after the first few every request is served from here); freed with the context.
struct DevPool {
    std::vector<std::pair<void *, size_t>> free_;
    void *get(size_t n, size_t *cap_out) {
        size_t best = free_.size();
        for (size_t k = 0; k < free_.size(); ++k) if (free_[k].second >= n && (best == free_.size() || free_[k].second < free_[best].second)) best = k;
        if (best != free_.size() && free_[best].second <= 4 * n + (1u << 20)) { void *p = free_[best].first; *cap_out = free_[best].second; free_.erase(free_.begin() + (ptrdiff_t) best); return p; }
        void *p = nullptr; const size_t want = n + n / 8 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) return nullptr;
        *cap_out = want; return p;
    }
    void put(void *p, size_t cap) { if (p) free_.emplace_back(p, cap); }
    void release() { for (auto &b : free_) cudaFree(b.first); free_.clear(); }
};

struct md_dev_reads {
    DevBuf arena; DevReads view; uint32_t n = 0, n_cigar = 0;
};

#define MD_NLANES 3
// One lane = one stream with its own staging arena, scratch and output buffers, so that the copy of tile k+1,
// the kernels of tile k and the read-back of tile k-1 overlap (md_submit_tile / md_collect_tile).
struct Lane {
    cudaStream_t stream = nullptr;
    md_dev_reads staged;                 // device copy of the host tile
    DevBuf rend, info, mate, tab, win, dir, calls, counters, sorted, sorted_off, acc, done, sites;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float timing[5] = {0, 0, 0, 0, 0};
    uint32_t last_nwin = 0; uint64_t last_ncalls = 0; bool pending = false; md_tile_stats last_stats;
    uint32_t *h_counters = nullptr;      // page-locked
    md_tile_desc last_tile; DevReads last_reads; bool last_mbias = false; uint32_t last_ncigar = 0;
};

struct md_ctx {
    int device = 0; md_config cfg; KParams kp;
    std::map<int32_t, Contig> contigs;
    DevPool pool;
    Lane lanes[MD_NLANES]; int rr = 0; Lane *last = nullptr;
    uint32_t *d_hist = nullptr; int32_t *d_lens = nullptr;
    uint64_t launches = 0;
    uint32_t W = 4096;
    uint32_t split_above = 4096, force_split = 0;   // windows averaging more alignments than this are split across CTAs (MD_SPLIT_ABOVE, MD_FORCE_SPLIT for tests)
    bool bed_mode = false;               // md_set_bed was called: -l semantics for every tile
    uint32_t ablate = 0;                 // MD_ABLATE: timing experiments (results are wrong when set)
    md_totals tot;                       // device time and work over all tiles (md_ctx_totals)
    int gen = 2;                         // count_warp<MODE, GEN>: 2 = list generator (site lists + rank look-ups), 1 = bitmap generator (always used with the variant filter; MD_GEN=1 forces it)
};

static void sync_all(md_ctx *c) { for (int k = 0; k < MD_NLANES; ++k) if (c->lanes[k].stream) cudaStreamSynchronize(c->lanes[k].stream); }

static void fill_kparams(const md_config *c, KParams &k) {
    memset(&k, 0, sizeof k);
    k.minMapq = c->minMapq; k.minPhred = c->minPhred < 1 ? 1 : c->minPhred;   /* extract.c:997-1000: -p below 1 is reset to 1 */ k.keepDupes = c->keepDupes; k.keepSingleton = c->keepSingleton;
    k.keepDiscordant = c->keepDiscordant; k.ignoreFlags = c->ignoreFlags; k.requireFlags = c->requireFlags; k.ignoreNH = c->ignoreNH;
    k.keepMask = (c->keepCpG ? 1 : 0) | (c->keepCHG ? 2 : 0) | (c->keepCHH ? 4 : 0);
    k.minOppositeDepth = c->minOppositeDepth; k.maxVariantFrac = c->maxVariantFrac;
    for (int i = 0; i < 16; ++i) { k.bounds[i] = c->bounds[i] < 0 ? 0 : c->bounds[i]; k.abounds[i] = c->absoluteBounds[i] < 0 ? 0 : c->absoluteBounds[i]; if (k.bounds[i] | k.abounds[i]) k.anyTrim = 1; }
    k.noOverlap = c->noOverlapMerge; k.minCE = c->minConversionEfficiency;
    for (int q = 0; q < 256; ++q) { volatile double v = (double) q; volatile double t = 0.2 * v; volatile double s = v + t; k.boost[q] = (unsigned char)(int) s; }
}

extern "C" md_ctx *md_create(const md_config *cfg, int device) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { g_err = std::string("no CUDA device available: ") + cudaGetErrorString(e); return nullptr; }
    if (device < 0 || device >= ndev) { g_err = "bad device index"; return nullptr; }
    CKN(cudaSetDevice(device));
    md_ctx *c = new md_ctx();
    c->device = device; c->cfg = *cfg; fill_kparams(cfg, c->kp); memset(&c->tot, 0, sizeof c->tot);
    for (int k = 0; k < MD_NLANES; ++k) {
        Lane *L = &c->lanes[k];
        int prio_lo = 0, prio_hi = 0; cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if (cudaStreamCreateWithPriority(&L->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { g_err = "cudaStreamCreate failed"; delete c; return nullptr; }   // the tile kernels go ahead of the device decoder's
        for (int i = 0; i < 5; ++i) cudaEventCreate(&L->ev[i]);
        if (cudaMallocHost((void **) &L->h_counters, C_N * 4) != cudaSuccess) { g_err = "cudaMallocHost failed"; delete c; return nullptr; }
    }
    Lane *L = &c->lanes[0]; c->last = L;
    size_t hb = (size_t) 4 * 2 * MD_MBIAS_MAXLEN * 2 * sizeof(uint32_t);
    if (cudaMalloc(&c->d_hist, hb) != cudaSuccess || cudaMalloc(&c->d_lens, 4 * sizeof(int32_t)) != cudaSuccess) { g_err = "cudaMalloc(hist) failed"; delete c; return nullptr; }
    cudaMemsetAsync(c->d_hist, 0, hb, L->stream); cudaMemsetAsync(c->d_lens, 0, 4 * sizeof(int32_t), L->stream);
    cudaFuncSetAttribute(count_warp<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) warp_layout(4096, 0, 1, WS_SEQ_BYTES, WS_QUAL_BYTES).total);
    cudaFuncSetAttribute(count_warp<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) warp_layout(4096, 1, 1, WS_SEQ_BYTES, WS_QUAL_BYTES).total);
    cudaFuncSetAttribute(count_warp<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) warp_layout(4096, 2, 1, WS_SEQ_BYTES, WS_QUAL_BYTES).total);
    cudaFuncSetAttribute(count_warp<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) warp_layout(4096, 0, 1, WS_SEQ_BYTES, WS_QUAL_BYTES).total);
    cudaFuncSetAttribute(count_warp<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) warp_layout(4096, 2, 1, WS_SEQ_BYTES, WS_QUAL_BYTES).total);
    if (const char *v = getenv("MD_ABLATE")) c->ablate = (uint32_t) atoi(v);
    if (const char *v = getenv("MD_SPLIT_ABOVE")) c->split_above = (uint32_t) atol(v);
    if (const char *v = getenv("MD_FORCE_SPLIT")) { int e = atoi(v); c->force_split = e > 0 && e <= 64 ? (uint32_t) e : 0u; }
    if (const char *v = getenv("MD_GEN")) { int e = atoi(v); if (e == 1 || e == 2) c->gen = e; }
    cudaStreamSynchronize(L->stream);
    return c;
}

static md_totals g_last_totals;
extern "C" int md_ctx_totals(md_ctx *c, md_totals *out) { if (!c || !out) return -2; *out = c->tot; out->launches = c->launches; return 0; }
extern "C" int md_last_totals(md_totals *out) { if (!out) return -2; *out = g_last_totals; return 0; }

extern "C" void md_destroy(md_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    sync_all(c);
    g_last_totals = c->tot; g_last_totals.launches = c->launches;
    for (auto &kv : c->contigs) { cudaFree(kv.second.d_seq); if (kv.second.d_bounds) cudaFree(kv.second.d_bounds); if (kv.second.d_bed) cudaFree(kv.second.d_bed); }
    c->pool.release();
    dev_free_graveyard();
    for (int k = 0; k < MD_NLANES; ++k) {
        Lane *L = &c->lanes[k];
        L->staged.arena.release();
        DevBuf *bufs[] = {&L->rend, &L->info, &L->mate, &L->tab, &L->win, &L->dir, &L->calls, &L->counters, &L->sorted, &L->sorted_off, &L->acc, &L->done, &L->sites};
        for (DevBuf *b : bufs) b->release();
        for (int i = 0; i < 5; ++i) if (L->ev[i]) cudaEventDestroy(L->ev[i]);
        if (L->h_counters) cudaFreeHost(L->h_counters);
        if (L->stream) cudaStreamDestroy(L->stream);
    }
    if (c->d_hist) cudaFree(c->d_hist);
    if (c->d_lens) cudaFree(c->d_lens);
    delete c;
}

extern "C" int md_load_contig(md_ctx *c, int32_t tid, const char *seq, uint32_t len) {
    CK(cudaSetDevice(c->device));
    md_drop_contig(c, tid);
    Lane *L = &c->lanes[0];
    Contig g; g.len = len;
    g.d_seq = (unsigned char *) c->pool.get((size_t) len + 16, &g.cap_seq);
    if (!g.d_seq) { g_err = "md_load_contig: out of device memory"; return -100; }
    CK(cudaMemcpyAsync(g.d_seq, seq, len, cudaMemcpyHostToDevice, L->stream));
    CK(cudaStreamSynchronize(L->stream));
    c->contigs[tid] = g;
    return 0;
}

extern "C" int md_drop_contig(md_ctx *c, int32_t tid) {
    auto it = c->contigs.find(tid);
    if (it == c->contigs.end()) return 0;
    cudaSetDevice(c->device);
    sync_all(c);                                              // the lanes' kernels are done with it (the decode stream never reads contigs)
    c->pool.put(it->second.d_seq, it->second.cap_seq);
    c->pool.put(it->second.d_bounds, it->second.cap_bounds);
    c->pool.put(it->second.d_bed, it->second.cap_bed);
    c->contigs.erase(it);
    return 0;
}

extern "C" int md_set_mbias_chunks(md_ctx *c, int32_t tid, const uint32_t *bounds, uint32_t n_chunks) {
    auto it = c->contigs.find(tid);
    if (it == c->contigs.end()) { g_err = "md_set_mbias_chunks: contig not loaded"; return -2; }
    CK(cudaSetDevice(c->device));
    Lane *L = &c->lanes[0];
    if (it->second.d_bounds) { sync_all(c); c->pool.put(it->second.d_bounds, it->second.cap_bounds); it->second.d_bounds = nullptr; }
    it->second.d_bounds = (uint32_t *) c->pool.get(((size_t) n_chunks + 1) * sizeof(uint32_t), &it->second.cap_bounds);
    if (!it->second.d_bounds) { g_err = "md_set_mbias_chunks: out of device memory"; return -100; }
    CK(cudaMemcpyAsync(it->second.d_bounds, bounds, ((size_t) n_chunks + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, L->stream));
    CK(cudaStreamSynchronize(L->stream));
    it->second.n_chunks = n_chunks;
    return 0;
}

// -l: the BED regions of one contig.  From the first call on, the context is in BED mode: contigs without regions yield nothing.
extern "C" int md_set_bed(md_ctx *c, int32_t tid, const md_bed_region *regs, uint32_t n) {
    c->bed_mode = true;
    auto it = c->contigs.find(tid);
    if (it == c->contigs.end()) { g_err = "md_set_bed: contig not loaded"; return -2; }
    CK(cudaSetDevice(c->device));
    Lane *L = &c->lanes[0];
    if (it->second.d_bed) { sync_all(c); c->pool.put(it->second.d_bed, it->second.cap_bed); it->second.d_bed = nullptr; it->second.n_bed = 0; }
    if (!n) return 0;
    for (uint32_t k = 1; k < n; ++k) {
        const md_bed_region &a = regs[k - 1], &b = regs[k];
        if (a.start > b.start || (a.start == b.start && (a.end > b.end || (a.end == b.end && a.strand > b.strand)))) { g_err = "md_set_bed: regions must be sorted by start, end, strand"; return -2; }
    }
    std::vector<uint32_t> h((size_t) n * 3);
    uint32_t run = 0;
    for (uint32_t k = 0; k < n; ++k) { run = std::max(run, regs[k].end); h[k] = regs[k].start; h[(size_t) n + k] = run; h[(size_t) 2 * n + k] = regs[k].strand; }
    it->second.d_bed = (uint32_t *) c->pool.get(h.size() * sizeof(uint32_t), &it->second.cap_bed);
    if (!it->second.d_bed) { g_err = "md_set_bed: out of device memory"; return -100; }
    CK(cudaMemcpyAsync(it->second.d_bed, h.data(), h.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, L->stream));
    CK(cudaStreamSynchronize(L->stream));
    it->second.n_bed = n;
    return 0;
}

// ---- staging a host tile into one device arena ---------------------------------------------------
static size_t al256(size_t x) { return (x + 255) & ~(size_t) 255; }

static int stage_reads(md_ctx *c, Lane *L, md_dev_reads &d, const md_reads_soa *r) {
    (void) c;
    const size_t n = r->n_reads;
    size_t sz[12] = {n * 4, n * 2, n, n, n * 4, (n + 1) * 4, n * 4, n * 4, n * 8, (size_t) r->n_cigar_ops * 4, (size_t) r->seq_words * 4, (size_t) r->qual_words * 8};
    const void *src[12] = {r->pos, r->flag, r->mapq, r->aux, r->l_qseq, r->cigar_off, r->seq_off, r->qual_off, r->frag_key, r->cigar, r->seq, r->qual};
    size_t off[12], tot = 0;
    for (int k = 0; k < 12; ++k) { off[k] = tot; tot += al256(sz[k] + 16); }
    if (d.arena.reserve(tot) != 0) return -100;
    unsigned char *base = (unsigned char *) d.arena.p;
    for (int k = 0; k < 12; ++k) if (sz[k]) CK(cudaMemcpyAsync(base + off[k], src[k], sz[k], cudaMemcpyHostToDevice, L->stream));
    DevReads &v = d.view;
    v.n = (uint32_t) n; v.seq_words = (uint32_t) r->seq_words; v.qual_words = (uint32_t) r->qual_words;
    v.qbits = (r->qual_bits == 2 || r->qual_bits == 4) ? r->qual_bits : 8u; memcpy(v.qlut, r->qual_lut, 16);
    if (r->qual_bits != 0 && r->qual_bits != 2 && r->qual_bits != 4 && r->qual_bits != 8) { g_err = "md_reads_soa.qual_bits must be 0, 2, 4 or 8"; return -2; }
    v.pos = (const int32_t *)(base + off[0]); v.flag = (const uint16_t *)(base + off[1]); v.mapq = base + off[2]; v.aux = base + off[3];
    v.l_qseq = (const uint32_t *)(base + off[4]); v.cigar_off = (const uint32_t *)(base + off[5]); v.seq_off = (const uint32_t *)(base + off[6]);
    v.qual_off = (const uint32_t *)(base + off[7]); v.frag_key = (const uint64_t *)(base + off[8]); v.cigar = (const uint32_t *)(base + off[9]);
    v.seq = (const uint32_t *)(base + off[10]); v.qual = (const uint64_t *)(base + off[11]); v.name_chk = nullptr;
    d.n = (uint32_t) n; d.n_cigar = r->n_cigar_ops;
    return 0;
}

extern "C" md_dev_reads *md_upload_reads(md_ctx *c, const md_reads_soa *reads) {
    CKN(cudaSetDevice(c->device));
    md_dev_reads *d = new md_dev_reads();
    Lane *L = &c->lanes[0];
    if (stage_reads(c, L, *d, reads) != 0) { delete d; return nullptr; }
    if (cudaStreamSynchronize(L->stream) != cudaSuccess) { g_err = "upload sync failed"; d->arena.release(); delete d; return nullptr; }
    return d;
}
extern "C" void md_free_reads(md_ctx *c, md_dev_reads *d) { if (!d) return; cudaSetDevice(c->device); sync_all(c); d->arena.release(); delete d; }


// ---- K4 launch (also used to re-run a tile after the host resolved duplicate query names) -------
static int launch_count(md_ctx *c, Lane *L, const Contig &g, const DevReads &R, const KParams &kp, uint32_t beg, uint32_t end, uint32_t n_win, unsigned long long cap_calls, bool mbias, bool wide) {
    if (!n_win) return 0;
    const uint32_t W = c->W;
    cudaStream_t s = L->stream;
    CountArgs A; memset(&A, 0, sizeof A);
    A.R = R; A.P = kp; A.rend = (const int32_t *) L->rend.p; A.info = (const uint8_t *) L->info.p; A.mate = (const int32_t *) L->mate.p; A.win = (const uint2 *) L->win.p;
    A.ref = g.d_seq; A.reflen = g.len; A.beg = beg; A.end = end; A.W = W; A.chunk_bounds = g.d_bounds; A.n_chunks = g.n_chunks;
    A.calls = (md_call *) L->calls.p; A.cap = cap_calls; A.dir = (uint2 *) L->dir.p; A.counters = (uint32_t *) L->counters.p; A.hist = c->d_hist; A.lens = c->d_lens;
    if (W != 16 * WS_WARPS * 32) { g_err = "internal: window size must be 4096"; return -3; }
    const int mode = mbias ? 2 : (kp.minOppositeDepth > 0 ? 1 : 0);
    // staging sized for 32 alignments of up to 160 bases in the tile's phred encoding (longer reads shrink the batch in the kernel);
    // 2-bit tiles with packed counters fit three CTAs per SM, everything else two
    A.st_seq = WS_SEQ_BYTES; A.st_qual = std::min<uint32_t>(WS_QUAL_BYTES, 32u * (((160u * R.qbits + 63u) >> 6) * 8u) + 32u);
    A.wide = wide ? 1u : 0u;
    // deep, narrow tiles (targeted panels: thousands of alignments per position window) would run on a handful of CTAs;
    // give every window several CTAs, each with ~2048 of its alignments
    A.split = 1; A.acc = nullptr; A.done = nullptr;
    {
        const uint64_t avg = (uint64_t) R.n / n_win;
        uint32_t S = avg > c->split_above ? (uint32_t) std::min<uint64_t>(48, (avg + 2047) / 2048) : 1u;
        if (c->force_split) S = c->force_split;
        if (S > 1) {
            const size_t acc_bytes = (size_t) n_win * (mode == 1 ? 4u : 2u) * W * 4u;
            if (mode != 2) {
                if (L->acc.reserve(acc_bytes) || L->done.reserve((size_t) n_win * 4)) return -100;
                CK(cudaMemsetAsync(L->acc.p, 0, acc_bytes, s)); CK(cudaMemsetAsync(L->done.p, 0, (size_t) n_win * 4, s));
                A.acc = (uint32_t *) L->acc.p; A.done = (uint32_t *) L->done.p;
            }
            A.split = S;
        }
    }
    const uint32_t n_cta = n_win * A.split;
    A.ablate = c->ablate;
    A.bed.start = g.d_bed; A.bed.pmax = g.d_bed ? g.d_bed + g.n_bed : nullptr; A.bed.strand = g.d_bed ? g.d_bed + 2 * (size_t) g.n_bed : nullptr; A.bed.n = g.n_bed; A.bed.on = c->bed_mode ? 1u : 0u;
    A.okmask = 0; for (int cde = 0; cde < 16; ++cde) if ((int) R.qlut[cde] >= kp.minPhred) A.okmask |= 1u << cde;
    const size_t sm = warp_layout(W, mode, A.wide, A.st_seq, A.st_qual).total;
    A.sites = nullptr;
    if (c->gen == 2 && mode != 1) {                              // list generator: 2 * W site slots per CTA (the variant filter keeps the bitmap generator)
        if (L->sites.reserve((size_t) n_cta * 2u * W * sizeof(uint16_t))) return -100;
        A.sites = (uint16_t *) L->sites.p;
        if (mode == 2) count_warp<2, 2><<<n_cta, WS_WARPS * 32, sm, s>>>(A);
        else count_warp<0, 2><<<n_cta, WS_WARPS * 32, sm, s>>>(A);
    } else {
        if (mode == 2) count_warp<2, 1><<<n_cta, WS_WARPS * 32, sm, s>>>(A);
        else if (mode == 1) count_warp<1, 1><<<n_cta, WS_WARPS * 32, sm, s>>>(A);
        else count_warp<0, 1><<<n_cta, WS_WARPS * 32, sm, s>>>(A);
    }
    c->launches += 1;
    if (!mbias) {
        dir_scan_kernel<<<1, 1024, 0, s>>>((uint2 *) L->dir.p, n_win, (uint32_t *) L->sorted_off.p);
        gather_kernel<<<n_win, 128, 0, s>>>((const md_call *) L->calls.p, (const uint2 *) L->dir.p, (const uint32_t *) L->sorted_off.p, (md_call *) L->sorted.p);
        c->launches += 2;
    }
    CK(cudaGetLastError());
    return 0;
}

// ---- kernel pipeline on a device-resident tile ---------------------------------------------------
static int run_pipeline(md_ctx *c, Lane *L, const md_tile_desc *t, const DevReads &R, bool mbias) {
    auto it = c->contigs.find(t->tid);
    if (it == c->contigs.end()) { g_err = "tile refers to a contig that was not loaded (md_load_contig)"; return -2; }
    const Contig &g = it->second;
    uint32_t beg = t->beg, end = std::min(t->end, g.len);
    if (beg > end) beg = end;
    if (mbias && !g.d_bounds) { g_err = "mbias tile without md_set_mbias_chunks"; return -2; }
    const uint32_t n = R.n, W = c->W;
    const uint32_t n_win = (end - beg + W - 1) / W;
    const uint32_t tab_cap = n + n / 4 + 1024;                         // names <= eligible records <= n  ->  load factor <= 0.8, ~0.4 for paired data
    const bool need_hash = !mbias && !c->kp.noOverlap;
    if (L->rend.reserve((size_t) n * 4 + 4) || L->info.reserve((size_t) n + 4) || L->mate.reserve((size_t) n * 4 + 4) ||
        L->win.reserve((size_t) n_win * 8 + 8) || L->dir.reserve((size_t) n_win * 8 + 8) || L->counters.reserve(C_N * 4)) return -100;
    if (need_hash && L->tab.reserve((size_t) tab_cap * 4)) return -100;
    const unsigned long long cap_calls = (unsigned long long)(end - beg) + 16;
    if (!mbias && (L->calls.reserve((size_t) cap_calls * sizeof(md_call)) || L->sorted.reserve((size_t) cap_calls * sizeof(md_call)) || L->sorted_off.reserve((size_t) n_win * 4 + 4))) return -100;
    cudaStream_t s = L->stream;
    CK(cudaEventRecord(L->ev[1], s));
    CK(cudaMemsetAsync(L->counters.p, 0, C_N * 4, s));
    HashTab T; T.tab = (uint32_t *) L->tab.p; T.cap = tab_cap;
    if (need_hash) CK(cudaMemsetAsync(L->tab.p, 0xff, (size_t) tab_cap * 4, s));
    if (n) CK(cudaMemsetAsync(L->mate.p, 0xff, (size_t) n * 4, s));
    KParams kp = c->kp; if (mbias) kp.noOverlap = 1;
    if (n) {
        const uint32_t gb = std::min<uint32_t>((n + 255) / 256, 148u * 8u);     // 8 CTAs of 256 threads per SM, grid-stride
        uint32_t ce_beg = t->ce_beg, ce_end = (t->ce_end == 0 || t->ce_end > g.len) ? g.len : t->ce_end;
        if (ce_beg > ce_end) ce_beg = ce_end;
        BedView bv; bv.start = g.d_bed; bv.pmax = g.d_bed ? g.d_bed + g.n_bed : nullptr; bv.strand = g.d_bed ? g.d_bed + 2 * (size_t) g.n_bed : nullptr; bv.n = g.n_bed; bv.on = c->bed_mode ? 1u : 0u;
        prep_kernel<<<gb, 256, 0, s>>>(R, kp, (int32_t *) L->rend.p, (uint8_t *) L->info.p, T, (int32_t *) L->mate.p, (uint32_t *) L->counters.p, g.d_seq, ce_beg, ce_end, bv);
        c->launches += 1;
    }
    CK(cudaEventRecord(L->ev[2], s));
    L->last_tile = *t; L->last_reads = R; L->last_mbias = mbias;
    int rc = launch_count(c, L, g, R, kp, beg, end, n_win, cap_calls, mbias, /*wide=*/false);
    if (rc) return rc;
    CK(cudaEventRecord(L->ev[3], s));
    CK(cudaGetLastError());
    L->last_nwin = mbias ? 0 : n_win;
    return 0;
}


// ---- duplicate query names ---------------------------------------------------------------------
// A name seen more than twice among the eligible records of a tile (secondary/supplementary records
// admitted with -F 0, or hand-made files such as the reference's own tests/cg_aln.bam) cannot be
// resolved by the two-slot table.  It is rare, so it is replayed on the host exactly as the pileup
// engine drives custom_overlap_constructor / _destructor (overlaps.c:121-147): records are pushed in
// file order; an eligible record either stores itself under its name or is merged with the stored one,
// which removes the name; a buffered record is dropped (and whatever sits under its name with it) once
// a record starting beyond its end has been pushed.  The count kernel is then run again.
#include <unordered_map>
static int resolve_duplicates_on_host(md_ctx *c, Lane *L) {
    const DevReads &R = L->last_reads;
    const uint32_t n = R.n;
    std::vector<int32_t> pos(n), rend(n), mate(n, -1); std::vector<uint16_t> flag(n); std::vector<uint8_t> info(n); std::vector<uint64_t> key(n); std::vector<uint32_t> chk(R.name_chk ? n : 0);
    cudaStream_t s = L->stream;
    CK(cudaMemcpyAsync(pos.data(), R.pos, (size_t) n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(rend.data(), L->rend.p, (size_t) n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(flag.data(), R.flag, (size_t) n * 2, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(info.data(), L->info.p, (size_t) n, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(key.data(), R.frag_key, (size_t) n * 8, cudaMemcpyDeviceToHost, s));
    if (R.name_chk) CK(cudaMemcpyAsync(chk.data(), R.name_chk, (size_t) n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (R.name_chk) for (uint32_t i = 0; i < n; ++i) key[i] ^= (uint64_t) chk[i] * 0x9e3779b97f4a7c15ull;      // one map key from both hashes
    std::vector<std::pair<int32_t, uint32_t>> by_end;
    for (uint32_t i = 0; i < n; ++i) if (info[i] & INFO_ADMIT) by_end.emplace_back(rend[i], i);
    std::sort(by_end.begin(), by_end.end());
    std::unordered_map<uint64_t, uint32_t> stored;
    size_t ev = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (!(info[i] & INFO_ADMIT)) continue;
        if (info[i] & INFO_ELIG) {
            auto it = stored.find(key[i]);
            if (it == stored.end()) stored.emplace(key[i], i);
            else {
                uint32_t a = it->second;
                stored.erase(it);
                int sa = INFO_STRAND(info[a]), sb = INFO_STRAND(info[i]);
                if (!((sa - sb) & 1) && pos[a] < rend[i] && pos[i] < rend[a]) { mate[a] = (int32_t) i; mate[i] = (int32_t) a; }
            }
        }
        while (ev < by_end.size() && by_end[ev].first < pos[i]) { uint32_t x = by_end[ev].second; if (x < i) stored.erase(key[x]); ++ev; }
    }
    CK(cudaMemcpyAsync(L->mate.p, mate.data(), (size_t) n * 4, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

// Run the count stage of the lane's last tile again from a clean output cursor (after the host fixed mate[], or with
// 32-bit window counters after a 16-bit one wrapped).
static int rerun_count(md_ctx *c, Lane *L, bool wide) {
    cudaStream_t s = L->stream;
    CK(cudaMemsetAsync((uint32_t *) L->counters.p + C_NCALLS, 0, 8, s));
    CK(cudaMemsetAsync((uint32_t *) L->counters.p + C_OVERFLOW, 0, 4, s));
    auto it = c->contigs.find(L->last_tile.tid);
    if (it == c->contigs.end()) { g_err = "internal: contig vanished"; return -3; }
    const Contig &g = it->second;
    uint32_t beg = L->last_tile.beg, end = std::min(L->last_tile.end, g.len);
    if (beg > end) beg = end;
    const uint32_t n_win = (end - beg + c->W - 1) / c->W;
    KParams kp = c->kp;
    int rc = launch_count(c, L, g, L->last_reads, kp, beg, end, n_win, (unsigned long long)(end - beg) + 16, false, wide);
    if (rc) return rc;
    CK(cudaMemcpyAsync(L->h_counters, L->counters.p, C_N * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

static int finish_counters(md_ctx *c, Lane *L, md_tile_stats *st) {
    CK(cudaMemcpyAsync(L->h_counters, L->counters.p, C_N * 4, cudaMemcpyDeviceToHost, L->stream));
    CK(cudaStreamSynchronize(L->stream));
    if (L->h_counters[C_MULTI] && !L->last_mbias && L->h_counters[C_OVERFLOW] != 1u && L->h_counters[C_OVERFLOW] != 2u) {
        int rc = resolve_duplicates_on_host(c, L);
        if (!rc) rc = rerun_count(c, L, false);
        if (rc) return rc;
    }
    if (L->h_counters[C_OVERFLOW] == 3u && !L->last_mbias) {      // a packed 16-bit window counter wrapped (depth >= 65536): count again, wide
        int rc = rerun_count(c, L, true);
        if (rc) return rc;
    }
    if (L->last_mbias && L->h_counters[C_MAXLQ] > (uint32_t) MD_MBIAS_MAXLEN) {
        // the reference grows its per-position arrays without bound (MBias.c:16-40); this build's histogram is fixed: refuse
        // rather than hand back a table that silently lacks the positions beyond it
        char msg[160]; snprintf(msg, sizeof msg, "mbias: the tile holds a read of %u bases; this build histograms query positions below %d (MD_MBIAS_MAXLEN)", L->h_counters[C_MAXLQ], MD_MBIAS_MAXLEN);
        g_err = msg; return -6;
    }
    unsigned long long ncalls; memcpy(&ncalls, &L->h_counters[C_NCALLS], 8);   // C_NCALLS is 8-byte aligned (index 4)
    if (L->h_counters[C_OVERFLOW]) { g_err = L->h_counters[C_OVERFLOW] == 2u ? "internal: staging copy did not complete" : "internal: call buffer overflow"; return -3; }
    L->last_ncalls = ncalls;
    md_tile_stats s; memset(&s, 0, sizeof s);
    s.n_calls = ncalls; s.n_required = ncalls; s.n_admitted = L->h_counters[C_ADMIT]; s.n_pairs = L->h_counters[C_PAIRED]; s.n_multi = L->h_counters[C_MULTI];
    L->last_stats = s;
    if (st) *st = s;
    return 0;
}

// the records are already in position order on the device (dir_scan_kernel + gather_kernel)
static int fetch_sorted(md_ctx *c, Lane *L, md_call *out, uint64_t capacity, uint64_t *n_out) {
    (void) c;
    const uint64_t n = L->last_ncalls;
    if (n_out) *n_out = n;
    if (n > capacity) { g_err = "md_call capacity too small"; return -1; }
    if (n == 0) return 0;
    CK(cudaMemcpyAsync(out, L->sorted.p, n * sizeof(md_call), cudaMemcpyDeviceToHost, L->stream));
    CK(cudaStreamSynchronize(L->stream));
    return 0;
}

static void collect_timing(md_ctx *c, Lane *L) {
    float t;
    for (int k = 0; k < 4; ++k) { t = 0; if (cudaEventElapsedTime(&t, L->ev[k], L->ev[k + 1]) == cudaSuccess) L->timing[k] = t; else L->timing[k] = 0; }
    t = 0; if (cudaEventElapsedTime(&t, L->ev[0], L->ev[4]) == cudaSuccess) L->timing[4] = t;
    md_totals &T = c->tot;
    T.h2d_ms += L->timing[0]; T.prep_ms += L->timing[1]; T.count_ms += L->timing[2]; T.d2h_ms += L->timing[3]; T.tile_ms += L->timing[4];
    T.tiles += 1; T.alignments += L->last_reads.n; T.cigar_ops += L->last_ncigar; T.calls += L->last_mbias ? 0 : L->last_ncalls;
}

// Lane selection: round-robin over the lanes that have nothing in flight.
static Lane *free_lane(md_ctx *c, int *ticket) {
    for (int k = 0; k < MD_NLANES; ++k) {
        int idx = (c->rr + k) % MD_NLANES;
        if (!c->lanes[idx].pending) { c->rr = (idx + 1) % MD_NLANES; *ticket = idx; return &c->lanes[idx]; }
    }
    return nullptr;
}

static int submit_common(md_ctx *c, const md_tile_desc *tile, const md_reads_soa *reads, bool mbias) {
    CK(cudaSetDevice(c->device));
    int ticket = -1;
    Lane *L = free_lane(c, &ticket);
    if (!L) { g_err = "md_submit_tile: all lanes are busy (collect a ticket first)"; return -4; }
    CK(cudaEventRecord(L->ev[0], L->stream));
    int rc = stage_reads(c, L, L->staged, reads);
    if (rc) return rc;
    L->last_ncigar = L->staged.n_cigar;
    rc = run_pipeline(c, L, tile, L->staged.view, mbias);
    if (rc) return rc;
    // queue the small read-backs now so that collect only has to wait
    CK(cudaMemcpyAsync(L->h_counters, L->counters.p, C_N * 4, cudaMemcpyDeviceToHost, L->stream));
    L->pending = true;
    c->last = L;
    return ticket;
}
extern "C" int md_submit_tile(md_ctx *c, const md_tile_desc *tile, const md_reads_soa *reads) { return submit_common(c, tile, reads, false); }
extern "C" int md_submit_mbias_tile(md_ctx *c, const md_tile_desc *tile, const md_reads_soa *reads) { return submit_common(c, tile, reads, true); }

extern "C" int md_collect_tile(md_ctx *c, int ticket, md_call *calls, uint64_t capacity, md_tile_stats *stats) {
    CK(cudaSetDevice(c->device));
    if (ticket < 0 || ticket >= MD_NLANES || !c->lanes[ticket].pending) { g_err = "md_collect_tile: bad ticket / nothing in flight"; return -4; }
    Lane *L = &c->lanes[ticket];
    L->pending = false;
    int rc = finish_counters(c, L, stats);
    if (rc) return rc;
    if (!L->last_mbias) rc = fetch_sorted(c, L, calls, capacity, nullptr);
    CK(cudaEventRecord(L->ev[4], L->stream));
    CK(cudaStreamSynchronize(L->stream));
    collect_timing(c, L);
    c->last = L;
    return rc;
}

extern "C" int md_extract_tile(md_ctx *c, const md_tile_desc *tile, const md_reads_soa *reads, md_call *calls, uint64_t capacity, md_tile_stats *stats) {
    int t = md_submit_tile(c, tile, reads);
    if (t < 0) return t;
    return md_collect_tile(c, t, calls, capacity, stats);
}

// perRead: one kernel over the tile's alignments, results copied back in input order (out[i].nmeth == 0xffffffff: not reported)
extern "C" int md_per_read_tile(md_ctx *c, const md_tile_desc *tile, const md_reads_soa *reads, uint32_t chunk_size, md_read_meth *out) {
    CK(cudaSetDevice(c->device));
    auto it = c->contigs.find(tile->tid);
    if (it == c->contigs.end()) { g_err = "tile refers to a contig that was not loaded (md_load_contig)"; return -2; }
    if (!chunk_size) { g_err = "md_per_read_tile: chunk_size must be at least 1"; return -2; }
    const Contig &g = it->second;
    int ticket = -1;
    Lane *L = free_lane(c, &ticket);
    if (!L) { g_err = "md_per_read_tile: all lanes are busy (collect a ticket first)"; return -4; }
    int rc = stage_reads(c, L, L->staged, reads);
    if (rc) return rc;
    const uint32_t n = L->staged.view.n;
    if (!n || !g.len) return 0;
    if (L->calls.reserve((size_t) n * sizeof(md_read_meth))) return -100;
    const uint32_t end = std::min(tile->end, g.len), beg = std::min(tile->beg, end);
    const uint32_t gb = std::min<uint32_t>((n + 127) / 128, 148u * 16u);
    per_read_kernel<<<gb, 128, 0, L->stream>>>(L->staged.view, c->kp, g.d_seq, g.len, beg, end, chunk_size, (md_read_meth *) L->calls.p);
    c->launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, L->calls.p, (size_t) n * sizeof(md_read_meth), cudaMemcpyDeviceToHost, L->stream));
    CK(cudaStreamSynchronize(L->stream));
    c->last = L;
    return 0;
}

extern "C" int md_extract_tile_device(md_ctx *c, const md_tile_desc *tile, const md_dev_reads *reads, md_tile_stats *stats) {
    CK(cudaSetDevice(c->device));
    Lane *L = &c->lanes[0];
    if (L->pending) { g_err = "md_extract_tile_device: lane 0 has a tile in flight"; return -4; }
    CK(cudaEventRecord(L->ev[0], L->stream));
    L->last_ncigar = reads->n_cigar;
    int rc = run_pipeline(c, L, tile, reads->view, false);
    if (rc) return rc;
    rc = finish_counters(c, L, stats);
    CK(cudaEventRecord(L->ev[4], L->stream));
    CK(cudaStreamSynchronize(L->stream));
    collect_timing(c, L);
    c->last = L;
    return rc;
}

extern "C" int md_fetch_calls(md_ctx *c, md_call *calls, uint64_t capacity, uint64_t *n_calls) {
    CK(cudaSetDevice(c->device));
    return fetch_sorted(c, c->last, calls, capacity, n_calls);
}

extern "C" int md_mbias_tile(md_ctx *c, const md_tile_desc *tile, const md_reads_soa *reads, md_tile_stats *stats) {
    int t = md_submit_mbias_tile(c, tile, reads);
    if (t < 0) return t;
    return md_collect_tile(c, t, nullptr, 0, stats);
}

extern "C" int md_mbias_tile_device(md_ctx *c, const md_tile_desc *tile, const md_dev_reads *reads, md_tile_stats *stats) {
    CK(cudaSetDevice(c->device));
    Lane *L = &c->lanes[0];
    CK(cudaEventRecord(L->ev[0], L->stream));
    L->last_ncigar = reads->n_cigar;
    int rc = run_pipeline(c, L, tile, reads->view, true);
    if (rc) return rc;
    rc = finish_counters(c, L, stats);
    CK(cudaEventRecord(L->ev[4], L->stream));
    CK(cudaStreamSynchronize(L->stream));
    collect_timing(c, L);
    c->last = L;
    return rc;
}

extern "C" int md_mbias_hist(md_ctx *c, uint32_t *hist, int32_t lens[4]) {
    CK(cudaSetDevice(c->device));
    sync_all(c);
    Lane *L = &c->lanes[0];
    CK(cudaMemcpyAsync(hist, c->d_hist, (size_t) 4 * 2 * MD_MBIAS_MAXLEN * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, L->stream));
    CK(cudaStreamSynchronize(L->stream));
    // strandMeth::l (MBias.c:212): one past the largest query position that received a call, over both reads of the strand
    for (int s = 0; s < 4; ++s) {
        int32_t l = 0;
        for (int rd = 0; rd < 2; ++rd) {
            const uint32_t *h = hist + ((size_t) s * 2 + rd) * MD_MBIAS_MAXLEN * 2;
            for (int q = MD_MBIAS_MAXLEN - 1; q >= l; --q) if (h[2 * q] | h[2 * q + 1]) { l = q + 1; break; }
        }
        lens[s] = l;
    }
    return 0;
}

extern "C" int md_mbias_reset(md_ctx *c) {
    CK(cudaSetDevice(c->device));
    sync_all(c);
    Lane *L = &c->lanes[0];
    CK(cudaMemsetAsync(c->d_hist, 0, (size_t) 4 * 2 * MD_MBIAS_MAXLEN * 2 * sizeof(uint32_t), L->stream));
    CK(cudaMemsetAsync(c->d_lens, 0, 4 * sizeof(int32_t), L->stream));
    CK(cudaStreamSynchronize(L->stream));
    return 0;
}

extern "C" void *md_alloc_pinned(size_t bytes) { void *p = nullptr; if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { g_err = "cudaMallocHost failed"; return nullptr; } return p; }
extern "C" void md_free_pinned(void *p) { if (p) cudaFreeHost(p); }
extern "C" int md_host_register(void *p, size_t bytes) { if (!p || !bytes) return 0; CK(cudaHostRegister(p, bytes, cudaHostRegisterDefault)); return 0; }
extern "C" int md_host_unregister(void *p) { if (!p) return 0; CK(cudaHostUnregister(p)); return 0; }
extern "C" int md_last_timing(md_ctx *c, float out[5]) { for (int k = 0; k < 5; ++k) out[k] = c->last->timing[k]; return 0; }
extern "C" uint64_t md_launch_count(md_ctx *c) { return c->launches; }
extern "C" void *md_stream(md_ctx *c) { return (void *) c->lanes[0].stream; }

#include "bamdev.cu"
