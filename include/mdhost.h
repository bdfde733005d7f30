/*
 * mdhost.h — C ABI of the host side of the B200 MethylDackel path: the drop-in sub-command
 * entry points and the helpers the tests and the benchmark use to decode BAM files into
 * md_reads_soa tiles.
 *
 * mdh_extract_main / mdh_mbias_main / mdh_perread_main take the reference's argv (extract.c:706, MBias.c:304,
 * perRead.c:305) plus the device back end to drive; the entry points with EXACTLY the reference's signatures
 * (`int extract_main(int, char **)`, main.c:17-20) are in methyldackel.h / lib/libMethylDackel.so, bound to libmdgpu.
 */
#ifndef MDHOST_H
#define MDHOST_H
#include "mdgpu.h"
#ifdef __cplusplus
extern "C" {
#endif

/* The device back end the sub-commands drive.  The shipped binary binds these slots to
 * libmdgpu (md_create / md_load_contig / md_extract_tile / ...).  The seam exists so the host
 * logic (option parsing, tiling, chunk replay, formatting) can be exercised by tests with a
 * checker bound instead; the product never binds anything but libmdgpu. */
typedef struct mdh_backend {
    void *factory_user;
    void *(*create)(void *factory_user, const md_config *cfg);
    void  (*destroy)(void *be);
    int   (*load_contig)(void *be, int32_t tid, const char *seq, uint32_t len);
    int   (*drop_contig)(void *be, int32_t tid);
    int   (*extract_tile)(void *be, const md_tile_desc *tile, const md_reads_soa *reads, md_call *calls, uint64_t cap, md_tile_stats *st);
    int   (*set_mbias_chunks)(void *be, int32_t tid, const uint32_t *bounds, uint32_t n_chunks);
    int   (*mbias_tile)(void *be, const md_tile_desc *tile, const md_reads_soa *reads, md_tile_stats *st);
    int   (*mbias_hist)(void *be, uint32_t *hist, int32_t lens[4]);
    const char *(*last_error)(void);
    /* optional (may be NULL): asynchronous tiles + page-locked tile memory; when present the extract driver keeps
     * several tiles in flight so that BAM decode, H2D, kernels and D2H overlap */
    int   (*submit_tile)(void *be, const md_tile_desc *tile, const md_reads_soa *reads);
    int   (*collect_tile)(void *be, int ticket, md_call *calls, uint64_t cap, md_tile_stats *st);
    void *(*pinned_alloc)(size_t bytes);
    void  (*pinned_free)(void *p);
    int   (*submit_mbias_tile)(void *be, const md_tile_desc *tile, const md_reads_soa *reads);   /* optional; collected with collect_tile(calls = NULL) */
    /* optional: device-side BGZF inflate + BAM decode (md_bam_* in mdgpu.h).  When present (and MD_DEVICE_DECODE is not 0)
     * the sub-commands ship the compressed file in segments of whole BGZF blocks instead of decoding it on host cores. */
    void *(*bam_open)(void *be, int32_t n_targets);
    void  (*bam_close)(void *s);
    void  (*bam_reset)(void *s);
    int   (*bam_push)(void *s, const void *comp, uint64_t bytes, const md_bgzf_block *blocks, uint32_t n_blocks, uint32_t skip, md_bam_summary *out);
    int   (*bam_get_runs)(void *s, md_bam_run *runs, uint32_t cap);
    int   (*bam_extract_run)(void *s, int run, const md_tile_desc *tile, uint32_t keep_hi, md_call *calls, uint64_t cap, md_tile_stats *st);
    int   (*bam_mbias_run)(void *s, int run, const md_tile_desc *tile, uint32_t keep_hi, md_tile_stats *st);
    int   (*bam_push_begin)(void *s, const void *comp, uint64_t bytes, const md_bgzf_block *blocks, uint32_t n_blocks, uint32_t skip);   /* optional pair: overlapped push */
    int   (*bam_push_end)(void *s, md_bam_summary *out);
    /* optional: -l <BED> (md_set_bed); without it the sub-commands refuse -l */
    int   (*set_bed)(void *be, int32_t tid, const md_bed_region *regs, uint32_t n);
    /* optional: perRead (md_per_read_tile); without it the perRead sub-command refuses to run */
    int   (*per_read_tile)(void *be, const md_tile_desc *tile, const md_reads_soa *reads, uint32_t chunk_size, md_read_meth *out);
    /* optional: md_bam_prefetch — copy of the next segment overlapped with the decode in flight */
    int   (*bam_prefetch)(void *s, const void *comp, uint64_t bytes);
} mdh_backend;

/* Same argv conventions as the reference: argv[0] is the sub-command name. */
int mdh_extract_main(int argc, char *argv[], const mdh_backend *be);
int mdh_mbias_main(int argc, char *argv[], const mdh_backend *be);
int mdh_perread_main(int argc, char *argv[], const mdh_backend *be);   /* perRead_main, perRead.c:305, main.c:20 */

/* Run statistics of the last mdh_extract_main / mdh_mbias_main call in this process */
typedef struct mdh_run_stats {
    uint64_t n_records;        /* alignments handed to the device (straddling reads counted once per tile) */
    uint64_t n_tiles;
    uint64_t n_calls;
    double   t_decode_s, t_device_s, t_format_s, t_total_s;
    uint64_t n_variant_positions;   /* extract.c:1489 counter (summed over shards by the caller) */
} mdh_run_stats;
void mdh_last_run_stats(mdh_run_stats *out);

/* ---- BAM -> SoA helpers (tests, bench) ---- */
typedef struct mdh_bam mdh_bam;
mdh_bam *mdh_bam_open(const char *path);
void     mdh_bam_close(mdh_bam *b);
int      mdh_bam_n_targets(const mdh_bam *b);
const char *mdh_bam_target_name(const mdh_bam *b, int tid);
uint32_t mdh_bam_target_len(const mdh_bam *b, int tid);
/* Every alignment of contig `tid` overlapping [beg,end), as one tile (arrays owned by the handle and
 * valid until the next call on it).  Scans from the start of the file (no index needed). */
int mdh_bam_read_region(mdh_bam *b, int tid, uint32_t beg, uint32_t end, md_reads_soa *out);

/* The same region cut into tiles of ~target_reads alignments, exactly as the sub-command driver cuts it
 * (alignments straddling a cut are present in both neighbours). Returns the number of tiles; they stay
 * alive (and at fixed addresses) until the next mdh_bam_make_tiles()/close. */
int mdh_bam_make_tiles(mdh_bam *b, int tid, uint32_t beg, uint32_t end, uint64_t target_reads);
int mdh_bam_get_tile(mdh_bam *b, int k, md_tile_desc *td, md_reads_soa *out);

typedef struct mdh_fasta mdh_fasta;
mdh_fasta *mdh_fasta_open(const char *path);
void       mdh_fasta_close(mdh_fasta *f);
/* whole contig; pointer owned by the handle, valid until the next fetch */
const char *mdh_fasta_fetch(mdh_fasta *f, const char *name, uint32_t *len);

/* Reference chunk layout (extract.c:325-350 + adjustBounds) of one contig, as md_set_mbias_chunks wants it:
 * writes up to cap+1 bounds, returns the number of chunks. */
uint32_t mdh_chunk_bounds(const char *seq, uint32_t len, unsigned long chunk_size, uint32_t reg_beg, uint32_t reg_end,
                          uint32_t *bounds, uint32_t cap);

/* Report stage of mbias on a summed histogram (layout of md_mbias_hist): suggestion line (svg!=0) and/or --txt table. */
void mdh_mbias_report(const uint32_t *hist, const int32_t lens[4], int svg, int txt);
/* The same with the M-bias plots (makeSVGs, svg.c:302-437): opref = the SVG prefix (NULL: neither plots nor suggestion line),
 * which = keepCpG + 2*keepCHG + 4*keepCHH (MBias.c:558).  Returns 0, or -3 when a file could not be written. */
int mdh_mbias_report_svg(const uint32_t *hist, const int32_t lens[4], const char *opref, int which, int txt);

const char *mdh_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
