// libMethylDackel.so — the reference's sub-command entry points (main.c:17-20) bound to libmdgpu; see include/methyldackel.h.
// The device back end is libmdgpu and nothing else; if no CUDA device is usable the calls fail with an error instead of
// computing on the CPU.
#include <cstdlib>
#include "../../../include/methyldackel.h"

namespace {
int g_device = 0;
void *be_create(void *, const md_config *cfg) { return md_create(cfg, g_device); }
void be_destroy(void *b) { md_destroy((md_ctx *) b); }
int be_load(void *b, int32_t tid, const char *s, uint32_t n) { return md_load_contig((md_ctx *) b, tid, s, n); }
int be_drop(void *b, int32_t tid) { return md_drop_contig((md_ctx *) b, tid); }
int be_extract(void *b, const md_tile_desc *t, const md_reads_soa *r, md_call *c, uint64_t cap, md_tile_stats *st) { return md_extract_tile((md_ctx *) b, t, r, c, cap, st); }
int be_chunks(void *b, int32_t tid, const uint32_t *bo, uint32_t n) { return md_set_mbias_chunks((md_ctx *) b, tid, bo, n); }
int be_mbias(void *b, const md_tile_desc *t, const md_reads_soa *r, md_tile_stats *st) { return md_mbias_tile((md_ctx *) b, t, r, st); }
int be_hist(void *b, uint32_t *h, int32_t l[4]) { return md_mbias_hist((md_ctx *) b, h, l); }
int be_submit(void *b, const md_tile_desc *t, const md_reads_soa *r) { return md_submit_tile((md_ctx *) b, t, r); }
int be_collect(void *b, int ticket, md_call *c, uint64_t cap, md_tile_stats *st) { return md_collect_tile((md_ctx *) b, ticket, c, cap, st); }
int be_submit_mbias(void *b, const md_tile_desc *t, const md_reads_soa *r) { return md_submit_mbias_tile((md_ctx *) b, t, r); }
void *be_bam_open(void *b, int32_t nt) { return md_bam_open((md_ctx *) b, nt); }
void be_bam_close(void *s) { md_bam_close((md_bam_stream *) s); }
void be_bam_reset(void *s) { md_bam_reset((md_bam_stream *) s); }
int be_bam_push(void *s, const void *comp, uint64_t bytes, const md_bgzf_block *blocks, uint32_t n, uint32_t skip, md_bam_summary *out) { return md_bam_push((md_bam_stream *) s, comp, bytes, blocks, n, skip, out); }
int be_bam_runs(void *s, md_bam_run *runs, uint32_t cap) { return md_bam_get_runs((md_bam_stream *) s, runs, cap); }
int be_bam_extract(void *s, int run, const md_tile_desc *t, uint32_t keep_hi, md_call *c, uint64_t cap, md_tile_stats *st) { return md_bam_extract_run((md_bam_stream *) s, run, t, keep_hi, c, cap, st); }
int be_bam_mbias(void *s, int run, const md_tile_desc *t, uint32_t keep_hi, md_tile_stats *st) { return md_bam_mbias_run((md_bam_stream *) s, run, t, keep_hi, st); }
int be_bam_push_begin(void *s, const void *comp, uint64_t bytes, const md_bgzf_block *blocks, uint32_t n, uint32_t skip) { return md_bam_push_begin((md_bam_stream *) s, comp, bytes, blocks, n, skip); }
int be_bam_push_end(void *s, md_bam_summary *out) { return md_bam_push_end((md_bam_stream *) s, out); }
int be_bam_prefetch(void *s, const void *comp, uint64_t bytes) { return md_bam_prefetch((md_bam_stream *) s, comp, bytes); }
int be_set_bed(void *b, int32_t tid, const md_bed_region *r, uint32_t n) { return md_set_bed((md_ctx *) b, tid, r, n); }
int be_per_read(void *b, const md_tile_desc *t, const md_reads_soa *r, uint32_t chunk, md_read_meth *out) { return md_per_read_tile((md_ctx *) b, t, r, chunk, out); }

const mdh_backend g_backend = {nullptr, be_create, be_destroy, be_load, be_drop, be_extract, be_chunks, be_mbias, be_hist, md_last_error, be_submit, be_collect, md_alloc_pinned, md_free_pinned, be_submit_mbias,
                               be_bam_open, be_bam_close, be_bam_reset, be_bam_push, be_bam_runs, be_bam_extract, be_bam_mbias, be_bam_push_begin, be_bam_push_end, be_set_bed, be_per_read, be_bam_prefetch};
int env_device() { const char *e = getenv("MD_DEVICE"); return e ? atoi(e) : 0; }
}  // namespace

extern "C" const mdh_backend *mdh_gpu_backend(int device) { g_device = device; return &g_backend; }
extern "C" int extract_main(int argc, char *argv[]) { return mdh_extract_main(argc, argv, mdh_gpu_backend(env_device())); }
extern "C" int mbias_main(int argc, char *argv[]) { return mdh_mbias_main(argc, argv, mdh_gpu_backend(env_device())); }
extern "C" int perRead_main(int argc, char *argv[]) { return mdh_perread_main(argc, argv, mdh_gpu_backend(env_device())); }
