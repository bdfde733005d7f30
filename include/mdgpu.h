/*
 * mdgpu.h — C ABI of the B200-native `MethylDackel extract` / `mbias` pileup hot path.
 *
 * The reference (dpryan79/MethylDackel v0.6.1) has no FFI: its seam for this path is
 * the per-chunk body of extractCalls (extract.c:379-511) and extractMBias
 * (MBias.c:145-218), fed by htslib callbacks (common.c:407 filter_func,
 * overlaps.c:121 custom_overlap_constructor).  This header is the boundary a
 * maintainer would bind instead of that body: plain pointers and sizes, no C++
 * or torch types.  Each entry point cites the reference code it replaces.
 *
 * Units: a *tile* is a batch of coordinate-sorted alignments from ONE contig plus
 * the half-open interval [beg,end) of reference positions the tile OWNS.  The
 * caller must put into the tile every alignment that overlaps [beg,end)
 * (exactly what sam_itr_queryi(bai,tid,beg,end) returns, extract.c:379); each
 * position is owned by exactly one tile (extract.c:400).
 */
#ifndef MDGPU_H
#define MDGPU_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDGPU_ABI_VERSION 6

/* ---- options that reach the hot path: the subset of `Config`
 *      (MethylDackel.h:90-126) read by filter_func / the per-column loop ---- */
typedef struct md_config {
    int32_t keepCpG, keepCHG, keepCHH;      /* extract.c:725, --noCpG/--CHG/--CHH          */
    int32_t minMapq, minPhred;              /* extract.c:726 (-q 10, -p 5)                  */
    int32_t keepDupes, keepSingleton, keepDiscordant; /* common.c:420,429,430             */
    int32_t ignoreFlags, requireFlags;      /* common.c:418-419 (0xF00, 0)                  */
    int32_t ignoreNH;                       /* common.c:421                                 */
    int32_t minOppositeDepth;               /* extract.c:444                                */
    double  maxVariantFrac;                 /* extract.c:446                                */
    int32_t bounds[16];                     /* --OT/--OB/--CTOT/--CTOB, common.c:137-172    */
    int32_t absoluteBounds[16];             /* --nOT/.., common.c:174-208                   */
    int32_t noOverlapMerge;                 /* 1 = mbias semantics (MBias.c:160: no ctor)   */
    float   minConversionEfficiency;        /* common.c:442-444; 0 = off                     */
    int32_t reserved[6];
} md_config;

/* ---- one tile of decoded alignments, structure-of-arrays.
 * All arrays are host pointers (pinned memory makes the copy asynchronous) for
 * md_extract_tile()/md_mbias_tile(), or device pointers for the *_device variants.
 * Field provenance: the BAM record fields the reference reads through htslib
 * macros (bam1_core_t + bam_get_cigar/seq/qual, common.c:119-127,
 * overlaps.c:27-60) and the two aux tags it scans (XG common.c:85, NH common.c:422). */
typedef struct md_reads_soa {
    uint32_t n_reads;
    uint32_t n_cigar_ops;        /* total entries of cigar[]                                   */
    uint64_t seq_words;          /* length of seq[] in 32-bit words                             */
    uint64_t qual_words;         /* length of qual[] in 64-bit words                            */
    const int32_t  *pos;         /* [n] 0-based leftmost reference coordinate (core.pos)        */
    const uint16_t *flag;        /* [n] core.flag                                               */
    const uint8_t  *mapq;        /* [n] core.qual                                               */
    const uint8_t  *aux;         /* [n] bits0-1: XG tag (0 absent/other, 1 'C', 2 'G');
                                        bit2: NH tag present with value > 1                     */
    const uint32_t *l_qseq;      /* [n] query length                                            */
    const uint32_t *cigar_off;   /* [n+1] index of the read's first op in cigar[]               */
    const uint32_t *seq_off;     /* [n] index of the read's first 32-bit word in seq[]          */
    const uint32_t *qual_off;    /* [n] index of the read's first 64-bit word in qual[]         */
    const uint64_t *frag_key;    /* [n] 64-bit fingerprint of the query name (pairing key,
                                        replaces the khash string key of overlaps.c:125)        */
    const uint32_t *cigar;       /* BAM encoding: len<<4 | op                                   */
    const uint32_t *seq;         /* 4-bit bases, BAM nibble order (high nibble first in each
                                    byte), each read padded to a 32-bit word                    */
    const uint64_t *qual;        /* phreds, each read padded to a 64-bit word: plain bytes when
                                    qual_bits is 8 (or 0), else qual_bits-wide codes, base j of a
                                    read at bit j*qual_bits of its words (little endian)          */
    uint32_t qual_bits;          /* 8, 4 or 2 (0 means 8).  Sequencers emit few distinct phreds
                                    (4 on current Illumina instruments), so a tile whose alphabet
                                    fits is shipped as codes + table: the PCIe and HBM bytes of the
                                    largest column shrink 2-4x.  Lossless: qual_lut[code] is the
                                    phred byte the reference would read with bam_get_qual()        */
    uint8_t  qual_lut[16];
    uint8_t  reserved_[12];
} md_reads_soa;

typedef struct md_tile_desc {
    int32_t  tid;                /* contig id previously given to md_load_contig()              */
    uint32_t beg, end;           /* owned reference interval [beg,end)                          */
    uint32_t ce_beg, ce_end;     /* only read when minConversionEfficiency > 0: the reference window
                                    contig[ce_beg, ce_end) of the chunk the tile lies in, i.e.
                                    [localPos-2, localEnd+11) clipped to the contig (extract.c:369-381,
                                    common.c:363): computeConversionEfficiency only sees this window,
                                    so the verdict on a read depends on the chunk; ce_end = 0 means
                                    the whole contig                                               */
} md_tile_desc;

/* One reported reference column (the values extract.c:420-461 hands to writeCall) */
typedef struct md_call {
    uint32_t pos;                /* 0-based reference position                                  */
    uint32_t nmeth, nunmeth;     /* extract.c:438-440                                           */
    uint32_t info;               /* bits0-1 context (0 CpG,1 CHG,2 CHH; extract.c:407-415),
                                    bit2: reference base is G (direction<0),
                                    bit3: column excluded as a likely variant (extract.c:444-459;
                                          counts are then the pre-exclusion values)             */
} md_call;
#define MD_CALL_CTX(info)      ((info) & 3u)
#define MD_CALL_IS_G(info)     (((info) >> 2) & 1u)
#define MD_CALL_EXCLUDED(info) (((info) >> 3) & 1u)

typedef struct md_tile_stats {
    uint64_t n_calls;            /* records written to calls[]                                  */
    uint64_t n_required;         /* records the tile produced (== n_calls unless capacity hit)  */
    uint32_t n_admitted;         /* alignments that passed the filter_func tests                */
    uint32_t n_pairs;            /* mate pairs resolved for the overlap merge                   */
    uint32_t n_multi;            /* alignments whose query name occurred > 2 times in the tile  */
    uint32_t reserved;
} md_tile_stats;

/* mbias histogram layout: hist[((strand*2 + read2)*lmax + qpos)*2 + {0:meth,1:unmeth}],
 * strand 0..3 = OT,OB,CTOT,CTOB (MethylDackel.h:172-176 strandMeth, MBias.c:193-212) */
#define MD_MBIAS_MAXLEN 1024   /* the reference grows without bound (MBias.c:16-40); an mbias tile holding a longer read is
                                  refused with an error (-6) instead of being histogrammed incompletely */

typedef struct md_ctx md_ctx;

/* Create a context on CUDA device `device` (replaces the per-thread state set up at
 * extract.c:283-323).  Returns NULL on failure; see md_last_error(). */
md_ctx *md_create(const md_config *cfg, int device);
void    md_destroy(md_ctx *ctx);

/* Upload one contig's bases (ASCII, case preserved, as faidx_fetch_seq returns them,
 * extract.c:381).  The library keeps it resident in HBM until md_drop_contig(). */
int md_load_contig(md_ctx *ctx, int32_t tid, const char *seq, uint32_t len);
int md_drop_contig(md_ctx *ctx, int32_t tid);

/* mbias only: the reference classifies context inside the per-chunk window
 * contig[localPos..localEnd] (MBias.c:147,170-178), so context at chunk edges depends on
 * the chunk layout.  bounds[0..n] are the n chunks' starts followed by the last end. */
int md_set_mbias_chunks(md_ctx *ctx, int32_t tid, const uint32_t *bounds, uint32_t n_chunks);

/* -l <BED> (+ --keepStrand): the regions of one loaded contig, sorted the way the reference sorts them (sortBED, bed.c:64-85:
 * start, end, strand).  Replaces, on the device, the three BED tests of the pileup path:
 *   - read admission (filter_func -> spanOverlapsBED, common.c:432-439, bed.c:22-41): a read is dropped unless it overlaps a region;
 *   - column selection (posOverlapsBED, extract.c:402-405 / MBias.c:166, bed.c:46-54): a column counts when the first region
 *     whose end lies beyond it starts at or before it;
 *   - strand (readStrandOverlapsBED, extract.c:425 / MBias.c:184, bed.c:57-63): in a '+' region only OT/CTOT reads are looked at,
 *     in a '-' region only OB/CTOB reads (strand = 0 unless --keepStrand was given).
 * From the first call on the context is in BED mode: tiles of contigs without regions produce nothing.  Call after md_load_contig. */
/* perRead (perRead.c): the CpG calls processRead() makes on one alignment.  tile->[beg,end) is the region of the contig being
 * processed (the whole contig, or the -r interval); chunk_size is --chunkSize (chunks start at tile->beg, perRead.c:118-137) and
 * matters because the CpG context is evaluated in the reference window of the chunk an alignment starts in (perRead.c:176-181).
 * out[i] belongs to alignment i of `reads`; nmeth == 0xffffffff marks an alignment the sub-command does not report (it starts
 * outside [beg,end) or fails -q / -F / -R, perRead.c:186-191).  Uses md_config::minMapq, minPhred, ignoreFlags, requireFlags. */
typedef struct md_read_meth { uint32_t nmeth, nunmeth; } md_read_meth;
int md_per_read_tile(md_ctx *ctx, const md_tile_desc *tile, const md_reads_soa *reads, uint32_t chunk_size, md_read_meth *out /* n_reads entries */);

typedef struct md_bed_region { uint32_t start, end; uint32_t strand; /* 0 any, 1 '+', 2 '-' */ } md_bed_region;
int md_set_bed(md_ctx *ctx, int32_t tid, const md_bed_region *regs, uint32_t n);

/* extract: replaces the chunk body extract.c:379-494 (pileup + per-column counting +
 * variant test).  Writes up to `capacity` md_call records sorted by position into
 * calls[] (host memory) and fills *stats.  Synchronous.  0 on success. */
int md_extract_tile(md_ctx *ctx, const md_tile_desc *tile, const md_reads_soa *reads,
                    md_call *calls, uint64_t capacity, md_tile_stats *stats);

/* Asynchronous pair: md_submit_tile() enqueues copy + kernels on one of the context's
 * streams and returns a ticket >= 0; md_collect_tile() waits for it.  The caller's input
 * buffers must stay valid until the matching collect returns (ownership as in SURVEY 8b). */
int md_submit_tile(md_ctx *ctx, const md_tile_desc *tile, const md_reads_soa *reads);
int md_collect_tile(md_ctx *ctx, int ticket, md_call *calls, uint64_t capacity, md_tile_stats *stats);

/* mbias: replaces MBias.c:145-218; accumulates into the context's histogram. */
int md_mbias_tile(md_ctx *ctx, const md_tile_desc *tile, const md_reads_soa *reads, md_tile_stats *stats);
/* Asynchronous form: returns a ticket for md_collect_tile() (calls = NULL, capacity = 0); tiles in flight on different
 * streams add into the same histogram. */
int md_submit_mbias_tile(md_ctx *ctx, const md_tile_desc *tile, const md_reads_soa *reads);
/* Copies the accumulated histogram (uint32[4*2*MD_MBIAS_MAXLEN*2]) and the per-strand
 * lengths `l` (MBias.c:212) to the host. */
int md_mbias_hist(md_ctx *ctx, uint32_t *hist, int32_t lens[4]);
int md_mbias_reset(md_ctx *ctx);

/* Device-resident variants used by the benchmark's kernel-only timing: `reads` holds
 * DEVICE pointers (e.g. from md_upload_reads) and results stay on the device. */
typedef struct md_dev_reads md_dev_reads;
md_dev_reads *md_upload_reads(md_ctx *ctx, const md_reads_soa *reads);
void          md_free_reads(md_ctx *ctx, md_dev_reads *d);
int md_extract_tile_device(md_ctx *ctx, const md_tile_desc *tile, const md_dev_reads *reads, md_tile_stats *stats);
int md_mbias_tile_device(md_ctx *ctx, const md_tile_desc *tile, const md_dev_reads *reads, md_tile_stats *stats);
/* Fetch the calls of the last *_device extract (for checking). */
int md_fetch_calls(md_ctx *ctx, md_call *calls, uint64_t capacity, uint64_t *n_calls);

/* CUDA-event timing of the last tile on this context, in milliseconds:
 * out[0]=h2d, out[1]=prep+pair kernels, out[2]=count kernel, out[3]=d2h, out[4]=total. */
int md_last_timing(md_ctx *ctx, float out[5]);
/* Device time (CUDA events) and work totals of a context over all the tiles it ran: what a caller that only sees the
 * sub-command mains (extract_main ...) needs to tell kernel time from end-to-end time.  md_last_totals() returns the
 * totals of the context destroyed last in this process. */
typedef struct md_totals {
    double h2d_ms, prep_ms, count_ms, d2h_ms, tile_ms;   /* summed over tiles (tile_ms: first copy to last read-back of each tile) */
    double inflate_ms, frame_ms, push_h2d_ms;            /* device decode: BGZF inflate kernel, record framing kernels, segment copies */
    uint64_t tiles, alignments, cigar_ops, calls, launches;
    uint64_t comp_bytes, inflated_bytes;                 /* device decode: compressed bytes pushed / bytes they inflated to */
} md_totals;
int md_ctx_totals(md_ctx *ctx, md_totals *out);
int md_last_totals(md_totals *out);
/* Number of kernel launches issued by this context so far (for bench.py gpu_launches). */
uint64_t md_launch_count(md_ctx *ctx);
/* CUDA stream handle (cudaStream_t) the kernels are launched on. */
void *md_stream(md_ctx *ctx);

/* Page-lock / unlock a host buffer so md_extract_tile's copies run asynchronously at full PCIe rate
 * (cudaHostRegister); optional. */
void *md_alloc_pinned(size_t bytes);   /* cudaMallocHost; NULL on failure */
void md_free_pinned(void *p);
int md_host_register(void *p, size_t bytes);
int md_host_unregister(void *p);

/* ---- Device-side BAM decode -------------------------------------------------------------------
 * Replaces the work inside sam_itr_next (common.c:413; htslib bgzf_read + bam_read1): BGZF block inflate, record framing
 * and field extraction happen in HBM, and the resulting tiles never exist on the host.  The caller cuts the compressed
 * file into SEGMENTS of whole BGZF blocks (scanning the 18-byte block headers) and pushes them in file order; each push
 * reports the runs of records per contig, and the caller asks for one tile per run.  Reads that reach beyond a tile's end
 * are carried into the next tile of the same contig on the device (the reference re-fetches them per chunk, extract.c:379). */
typedef struct md_bgzf_block {
    uint64_t comp_off;     /* offset of the raw-deflate payload (after the 12+XLEN byte header) in the pushed buffer */
    uint32_t comp_len;     /* payload bytes (block size - header - 8 byte trailer) */
    uint32_t isize;        /* uncompressed size (BGZF trailer ISIZE) */
} md_bgzf_block;
typedef struct md_bam_run {  /* records [start, start+n) of the pushed segment lie on contig `tid`, positions first_pos..last_pos */
    int32_t tid; uint32_t start, n; int32_t first_pos, last_pos, prev_last_pos;
} md_bam_run;
typedef struct md_bam_summary { uint32_t n_records, n_runs; uint64_t inflated_bytes, leftover_bytes; } md_bam_summary;
typedef struct md_bam_stream md_bam_stream;
md_bam_stream *md_bam_open(md_ctx *ctx, int32_t n_targets);
void md_bam_close(md_bam_stream *s);
void md_bam_reset(md_bam_stream *s);      /* after a seek in the file */
/* `skip`: offset of the first record in the segment's inflated bytes (only after open/reset: the in-block part of a BAI
 * virtual offset, or the end of the BAM header); later segments continue the record that straddled in. */
int md_bam_push(md_bam_stream *s, const void *comp, uint64_t comp_bytes, const md_bgzf_block *blocks, uint32_t n_blocks, uint32_t skip, md_bam_summary *out);
/* The same in two halves: _begin() starts the copy + decode of the segment on a helper thread and its own stream and returns
 * at once; _end() waits for it, after which runs and tiles refer to that segment.  Between the two the caller may request
 * tiles of the PREVIOUS segment, so the transfer and inflate of segment k+1 overlap the counting of segment k.  The
 * compressed bytes and the block table must stay valid until _end(). */
int md_bam_push_begin(md_bam_stream *s, const void *comp, uint64_t comp_bytes, const md_bgzf_block *blocks, uint32_t n_blocks, uint32_t skip);
int md_bam_push_end(md_bam_stream *s, md_bam_summary *out);
/* Optional, between md_bam_push_begin() and md_bam_push_end(): start the host-to-device copy of the segment that will be pushed
 * next (same `comp` / `comp_bytes` as that later md_bam_push_begin(), which then skips its copy), so that it overlaps the decode in
 * flight.  The buffer must stay valid until that later push has ended.  comp == NULL waits for outstanding prefetches and drops
 * them.  A no-op (0) when no push is in flight. */
int md_bam_prefetch(md_bam_stream *s, const void *comp, uint64_t comp_bytes);
int md_bam_get_runs(md_bam_stream *s, md_bam_run *runs, uint32_t cap);
/* One tile = reads carried from the previous tile of this contig (if that tile ended where this one begins) + run `run`
 * of the last segment (run < 0: carried reads only, to close a contig), restricted to records with pos < keep_hi; then
 * the extract / mbias pipeline over the owned interval [tile->beg, tile->end) as md_extract_tile / md_mbias_tile. */
int md_bam_extract_run(md_bam_stream *s, int run, const md_tile_desc *tile, uint32_t keep_hi, md_call *calls, uint64_t capacity, md_tile_stats *stats);
int md_bam_mbias_run(md_bam_stream *s, int run, const md_tile_desc *tile, uint32_t keep_hi, md_tile_stats *stats);
/* the tile built last, copied back (for tests: must equal the tile the host decoder builds from the same records) */
int md_bam_tile_shape(md_bam_stream *s, md_reads_soa *shape);
int md_bam_tile_fetch(md_bam_stream *s, md_reads_soa *dst, int32_t *rend);

const char *md_last_error(void);
int md_abi_version(void);
/* first 16 hex digits of the SHA-256 over the CUDA sources this library was built from (`make -C methyldackel_b200/csrc srchash`
 * prints the tree's): lets a profile be tied to the source it maps to */
const char *md_source_hash(void);

#ifdef __cplusplus
}
#endif
#endif
