// `MethylDackel` drop-in binary for the B200 path: same dispatcher surface as main.c:39-62 for the
// two sub-commands this build accelerates.  The device back end is libmdgpu and nothing else;
// if no CUDA device is usable the program fails with an error instead of computing on the CPU.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <chrono>
#include "../../../include/mdhost.h"

static void *be_create(void *, const md_config *cfg) { return md_create(cfg, 0); }
static void be_destroy(void *b) { md_destroy((md_ctx *) b); }
static int be_load(void *b, int32_t tid, const char *s, uint32_t n) { return md_load_contig((md_ctx *) b, tid, s, n); }
static int be_drop(void *b, int32_t tid) { return md_drop_contig((md_ctx *) b, tid); }
static int be_extract(void *b, const md_tile_desc *t, const md_reads_soa *r, md_call *c, uint64_t cap, md_tile_stats *st) { return md_extract_tile((md_ctx *) b, t, r, c, cap, st); }
static int be_chunks(void *b, int32_t tid, const uint32_t *bo, uint32_t n) { return md_set_mbias_chunks((md_ctx *) b, tid, bo, n); }
static int be_mbias(void *b, const md_tile_desc *t, const md_reads_soa *r, md_tile_stats *st) { return md_mbias_tile((md_ctx *) b, t, r, st); }
static int be_hist(void *b, uint32_t *h, int32_t l[4]) { return md_mbias_hist((md_ctx *) b, h, l); }
static int be_submit(void *b, const md_tile_desc *t, const md_reads_soa *r) { return md_submit_tile((md_ctx *) b, t, r); }
static int be_collect(void *b, int ticket, md_call *c, uint64_t cap, md_tile_stats *st) { return md_collect_tile((md_ctx *) b, ticket, c, cap, st); }
static int be_submit_mbias(void *b, const md_tile_desc *t, const md_reads_soa *r) { return md_submit_mbias_tile((md_ctx *) b, t, r); }

static void *be_bam_open(void *b, int32_t nt) { return md_bam_open((md_ctx *) b, nt); }
static void be_bam_close(void *s) { md_bam_close((md_bam_stream *) s); }
static void be_bam_reset(void *s) { md_bam_reset((md_bam_stream *) s); }
static int be_bam_push(void *s, const void *comp, uint64_t bytes, const md_bgzf_block *blocks, uint32_t n, uint32_t skip, md_bam_summary *out) { return md_bam_push((md_bam_stream *) s, comp, bytes, blocks, n, skip, out); }
static int be_bam_runs(void *s, md_bam_run *runs, uint32_t cap) { return md_bam_get_runs((md_bam_stream *) s, runs, cap); }
static int be_bam_extract(void *s, int run, const md_tile_desc *t, uint32_t keep_hi, md_call *c, uint64_t cap, md_tile_stats *st) { return md_bam_extract_run((md_bam_stream *) s, run, t, keep_hi, c, cap, st); }
static int be_bam_mbias(void *s, int run, const md_tile_desc *t, uint32_t keep_hi, md_tile_stats *st) { return md_bam_mbias_run((md_bam_stream *) s, run, t, keep_hi, st); }

static int be_bam_push_begin(void *s, const void *comp, uint64_t bytes, const md_bgzf_block *blocks, uint32_t n, uint32_t skip) { return md_bam_push_begin((md_bam_stream *) s, comp, bytes, blocks, n, skip); }
static int be_bam_push_end(void *s, md_bam_summary *out) { return md_bam_push_end((md_bam_stream *) s, out); }

static int be_set_bed(void *b, int32_t tid, const md_bed_region *r, uint32_t n) { return md_set_bed((md_ctx *) b, tid, r, n); }

static int be_per_read(void *b, const md_tile_desc *t, const md_reads_soa *r, uint32_t chunk, md_read_meth *out) { return md_per_read_tile((md_ctx *) b, t, r, chunk, out); }

static void usage_main() {
    fprintf(stderr, "MethylDackel (B200 build of the extract/mbias hot path): A tool for processing bisulfite sequencing alignments.\n"
                    "Usage: MethylDackel <command> [options]\n\nCommands:\n"
                    "    mbias    Determine the position-dependent methylation bias in a dataset.\n"
                    "    extract  Extract methylation metrics from an alignment file in BAM format.\n"
                    "    perRead  Generate a per-read methylation summary.\n"
                    "(mergeContext is not part of this build.)\n");
}

int main(int argc, char *argv[]) {
    mdh_backend be = {nullptr, be_create, be_destroy, be_load, be_drop, be_extract, be_chunks, be_mbias, be_hist, md_last_error, be_submit, be_collect, md_alloc_pinned, md_free_pinned, be_submit_mbias,
                      be_bam_open, be_bam_close, be_bam_reset, be_bam_push, be_bam_runs, be_bam_extract, be_bam_mbias, be_bam_push_begin, be_bam_push_end, be_set_bed, be_per_read};
    if (argc == 1) { usage_main(); return 0; }
    if (!strcmp(argv[1], "-h") || !strcmp(argv[1], "--help")) { usage_main(); return 0; }
    if (!strcmp(argv[1], "-v") || !strcmp(argv[1], "--version")) { printf("0.6.1-b200 (B200 build; no HTSlib)\n"); return 0; }
    if (!strcmp(argv[1], "extract") || !strcmp(argv[1], "mbias")) {
        auto t0 = std::chrono::steady_clock::now();
        int rc = !strcmp(argv[1], "extract") ? mdh_extract_main(argc - 1, argv + 1, &be) : mdh_mbias_main(argc - 1, argv + 1, &be);
        if (getenv("MD_TIMING")) {      // where the wall clock went (stderr), for tuning
            mdh_run_stats st; mdh_last_run_stats(&st);
            double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            fprintf(stderr, "[md-timing] wall %.3f s: tiler (decode wait + assembly) %.3f, device calls %.3f, other %.3f on the calling thread; writer thread busy %.3f; %llu alignments in %llu tiles, %llu calls\n",
                    wall, st.t_decode_s, st.t_device_s, st.t_total_s - st.t_decode_s - st.t_device_s, st.t_format_s,
                    (unsigned long long) st.n_records, (unsigned long long) st.n_tiles, (unsigned long long) st.n_calls);
        }
        return rc;
    }
    if (!strcmp(argv[1], "perRead")) return mdh_perread_main(argc - 1, argv + 1, &be);
    if (!strcmp(argv[1], "mergeContext")) { fprintf(stderr, "The %s sub-command is not part of the B200 build.\n", argv[1]); return -1; }
    fprintf(stderr, "Unknown command!\n"); usage_main();
    return -1;
}
