/* TEST INFRASTRUCTURE — CPU restatement ("port") of the reference's extract / mbias
 * hot path (plus the -l BED tests and perRead's processRead), consuming the same SoA tiles as the CUDA library.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this; the product
 * (methyldackel_b200/) never links or loads it.
 *
 * Parity status: PINNED — this restatement is checked (tests/test_oracle_vs_ref.py)
 * against oracle/_ref/MethylDackel, i.e. the reference's own C sources compiled against
 * oracle/htslib_shim, on the reference's fixture BAMs and on synthetic BAMs; the shim-built
 * reference itself reproduces 14 of the 15 assertions of the reference's tests/test.py
 * (test 8 yields 11 lines where the script asserts 12). */
#ifndef MD_ORACLE_H
#define MD_ORACLE_H
#include "../include/mdgpu.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Counts for the owned interval [beg,end) of one contig.  `ref` is the WHOLE contig
 * (ASCII, case preserved), `reflen` its length.  Emits md_call records sorted by position:
 * every column of a kept context with nmeth+nunmeth > 0, plus every excluded-variant column
 * (info bit3).  Returns 0, or -1 if `cap` was too small (stats->n_required tells how many). */
int mdo_extract_tile(const md_config *cfg, const char *ref, uint32_t reflen, uint32_t beg, uint32_t end,
                     const md_reads_soa *reads, md_call *out, uint64_t cap, md_tile_stats *stats);
/* same with the conversion-efficiency window of the tile's chunk (md_tile_desc::ce_beg/ce_end; 0,0 = whole contig) */
int mdo_extract_tile_ce(const md_config *cfg, const char *ref, uint32_t reflen, uint32_t beg, uint32_t end, uint32_t ce_beg, uint32_t ce_end,
                        const md_reads_soa *reads, md_call *out, uint64_t cap, md_tile_stats *stats);

/* mbias accumulation (MBias.c:145-218). `bounds`/`n_chunks` as md_set_mbias_chunks();
 * hist is uint32[4*2*MD_MBIAS_MAXLEN*2] and is ADDED to; lens[] is max-updated. */
int mdo_mbias_tile(const md_config *cfg, const char *ref, uint32_t reflen, uint32_t beg, uint32_t end,
                   const uint32_t *bounds, uint32_t n_chunks,
                   const md_reads_soa *reads, uint32_t *hist, int32_t lens[4], md_tile_stats *stats);

/* perRead: processRead() on every alignment of the tile that the sub-command reports (contract of md_per_read_tile) */
int mdo_per_read_tile(const md_config *cfg, const char *ref, uint32_t reflen, uint32_t beg, uint32_t end, uint32_t chunk_size,
                      const md_reads_soa *reads, md_read_meth *out);

/* -l <BED>: the regions (sortBED order, bed.c:64-85) of the contig the following mdo_*_tile calls work on; on = 0 switches
 * the BED tests off.  The array must stay alive while it is set. */
void mdo_set_bed(const md_bed_region *regs, uint32_t n, int on);

/* same with the conversion-efficiency window of the tile's chunk (MBias.c:154-156: contig[localPos, localEnd]; 0,0 = whole contig) */
int mdo_mbias_tile_ce(const md_config *cfg, const char *ref, uint32_t reflen, uint32_t beg, uint32_t end, uint32_t ce_beg, uint32_t ce_end,
                      const uint32_t *bounds, uint32_t n_chunks,
                      const md_reads_soa *reads, uint32_t *hist, int32_t lens[4], md_tile_stats *stats);

/* Per-read helpers exposed for unit tests */
int mdo_strand(uint16_t flag, uint8_t aux);                       /* getStrand, common.c:84-116 */
int mdo_admit(const md_config *cfg, uint16_t flag, uint8_t mapq, uint8_t aux); /* filter_func, common.c:416-430 */
int mdo_context(const char *seq, int pos, int seqlen);           /* isCpG/isCHG/isCHH chain: 0 none, +-1 CpG, +-2 CHG, +-3 CHH */
uint8_t mdo_boost(uint8_t q);                                     /* (uint8_t)(q + 0.2*q), overlaps.c:103,106 */
#ifdef __cplusplus
}
#endif
#endif
