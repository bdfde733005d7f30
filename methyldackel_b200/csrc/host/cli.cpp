// Host driver of the B200 `MethylDackel extract` / `mbias` sub-commands (libmdhost).
//
// Mirrors the reference's drop-in surface for this path: the option tables and defaults of
// extract_main (extract.c:715-946) and mbias_main (MBias.c:312-443), their validation and exit
// codes (extract.c:979-1034, MBias.c:445-469), output file naming and headers
// (extract.c:1344-1439), `-r` handling (extract.c:1441-1468) and the end-of-run messages
// (extract.c:1489).  The per-chunk worker body (extract.c:379-511 / MBias.c:145-218) is NOT
// here: alignments are decoded into SoA tiles and handed to the device back end.
#include <getopt.h>
#include <cstdlib>
#include <cerrno>
#include <climits>
#include <chrono>
#include <map>
#include <future>
#include <malloc.h>
#include "tiles.hpp"
#include "pardecode.hpp"
#include "format.hpp"
#include "mbias_report.hpp"
#include "mbias_svg.hpp"
#include "bed.hpp"
#include "../../../include/mdhost.h"

using namespace mdhost;

#define MD_VERSION "0.6.1-b200"

static thread_local std::string g_err;
static mdh_run_stats g_stats;
extern "C" const char *mdh_last_error(void) { return g_err.c_str(); }
extern "C" void mdh_last_run_stats(mdh_run_stats *out) { if (out) *out = g_stats; }

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
// MD_TIMING=1: phase marks on stderr (seconds since the sub-command started)
static double g_t0 = 0; static bool g_marks = false;
static double g_acc[16];  // 0-6: see the [md-timing] calling-thread line; 7-12: device-decode driver (push_end, push_begin, runs, waits for the hand-over thread, read-back buffer growth, tile call)
struct Acc { int k; double t0; explicit Acc(int k_) : k(k_), t0(now_s()) {} ~Acc() { g_acc[k] += now_s() - t0; } };
static void mark(const char *what) { if (g_marks) fprintf(stderr, "[md-timing] %8.3f  %s\n", now_s() - g_t0, what); }

// parseBounds, common.c:11-43: four comma-separated non-negative ints into vals[4*mult..]
static void parse_bounds(const char *s2, int *vals, int mult) {
    std::string s(s2);
    char *save = nullptr, *p = strtok_r(&s[0], ",", &save);
    for (int i = 0; i < 4; ++i) {
        char *end = nullptr; long v = -1;
        if (p) { errno = 0; long t = strtol(p, &end, 10); if (!(errno != 0 || end == p || t > INT_MAX || t < 0)) v = t; }
        if (v < 0) { fprintf(stderr, "Invalid bounds string, %s\n", s2); return; }
        vals[4 * mult + i] = (int) v;
        p = strtok_r(nullptr, ",", &save);
    }
}

// hts_parse_reg as the reference uses it (extract.c:1446): returns length of the name part, or -1
static int parse_region(const char *s, int *beg, int *end) {
    const char *colon = strrchr(s, ':');
    if (!colon) { *beg = 0; *end = INT_MAX; return (int) strlen(s); }
    auto dec = [](const char *p, const char **e) { long long v = 0; bool neg = false; if (*p == '-') { neg = true; ++p; } while (isdigit((unsigned char) *p) || *p == ',') { if (*p != ',') v = v * 10 + (*p - '0'); ++p; } *e = p; return neg ? -v : v; };
    const char *p; long long b = dec(colon + 1, &p) - 1, e;
    if (b < 0) {
        if (b != -1 && *p == '-' && colon[1] != '\0') return -1;
        if (isdigit((unsigned char) *p) || *p == '\0' || *p == ',') { e = (b == -1) ? INT_MAX : -(b + 1); *beg = 0; *end = (int) e; return (int)(colon - s); }
        if (b < -1) return -1;
    }
    if (*p == '\0' || *p == ',') e = INT_MAX;
    else if (*p == '-') { const char *q; e = dec(p + 1, &q); if (*q != '\0' && *q != ',') return -1; }
    else return -1;
    if (b >= e || b > INT_MAX) return -1;
    if (e > INT_MAX) e = INT_MAX;
    *beg = (int) b; *end = (int) e;
    return (int)(colon - s);
}

static void extract_usage() {
    fprintf(stderr,
"\nUsage: MethylDackel extract [OPTIONS] <ref.fa> <sorted_alignments.bam>\n\n"
"B200 build of the extract hot path. Options (same names, defaults and meaning as MethylDackel 0.6.1):\n"
"  -q INT  -p INT  -d INT  -r STR  -l FILE (BED)  --keepStrand  -o/--opref STR  -@ INT  --chunkSize INT  -D INT (ignored)\n"
"  --noCpG  --CHG  --CHH  --mergeContext  --fraction  --counts  --logit  --methylKit  --cytosine_report\n"
"  --keepDupes  --keepSingleton  --keepDiscordant  -F/--ignoreFlags INT  -R/--requireFlags INT  --ignoreNH\n"
"  --minOppositeDepth INT  --maxVariantFrac FLOAT\n"
"  --OT/--OB/--CTOT/--CTOB INT,INT,INT,INT   --nOT/--nOB/--nCTOT/--nCTOB INT,INT,INT,INT\n"
"  -h/--help  -v/--version\n"
"  --minConversionEfficiency FLOAT\n"
"Not available in this build: -M/-t/-b/-O/-N/-B (mappability).\n"
"Note that --fraction, --counts, and --logit are mutually exclusive!\n");
}

static void mbias_usage() {
    fprintf(stderr,
"\nUsage: MethylDackel mbias [OPTIONS] <ref.fa> <sorted_alignments.bam> <output.prefix>\n\n"
"B200 build of the mbias hot path. Options (as MethylDackel 0.6.1):\n"
"  -q INT  -p INT  -r STR  -l FILE (BED)  --keepStrand  -@ INT  --chunkSize INT  -D INT (ignored)  --noCpG  --CHG  --CHH\n"
"  --keepDupes  --keepSingleton  --keepDiscordant  -F INT  -R INT  --ignoreNH  --txt  --noSVG\n"
"  --nOT/--nOB/--nCTOT/--nCTOB INT,INT,INT,INT  -h/--help  -v/--version\n"
"  --minConversionEfficiency FLOAT\n");
}

static bool load_bai(const std::string &bam, BaiIndex &idx) {
    if (BaiIndex::load(bam + ".bai", idx)) return true;
    size_t n = bam.size();
    if (n > 4 && bam.compare(n - 4, 4, ".bam") == 0) return BaiIndex::load(bam.substr(0, n - 4) + ".bai", idx);
    return false;
}

extern "C" uint32_t mdh_chunk_bounds(const char *seq, uint32_t len, unsigned long chunk_size, uint32_t reg_beg, uint32_t reg_end, uint32_t *bounds, uint32_t cap) {
    std::vector<uint32_t> lens{len};
    ChunkCursor cur(lens, chunk_size, 0, reg_beg, reg_end);
    Chunk c; uint32_t n = 0;
    auto fetch = [&](uint32_t, int64_t start, int64_t cnt, std::string &w) { w.clear(); if (start < (int64_t) len) w.assign(seq + start, (size_t) std::min<int64_t>(cnt, (int64_t) len - start)); return (int64_t) len; };
    while (cur.next(c, fetch)) {
        if (c.tid != 0) break;
        if (n < cap) { bounds[n] = c.beg; bounds[n + 1] = c.end; }
        ++n;
    }
    return n;
}

// -@ N is the number of host decode threads here (the reference uses it for its pileup workers, extract.c:607).
// Without -@ the decode pool uses the machine (capped), because the GPU otherwise starves behind one inflate thread.
static int decode_threads(int n, bool given) {
    if (given) return n < 1 ? 1 : n;
    unsigned hw = std::thread::hardware_concurrency();
    if (const char *e = getenv("MD_DECODE_THREADS")) { int v = atoi(e); if (v > 0) return v; }
    return (int) std::min<unsigned>(hw ? hw : 1, 64);
}

// While the CUDA context is being created, full-speed decoding on every core makes that creation several times slower (both
// sides fight over the process's address-space lock); a handful of decode threads run ahead until the device is up.
static int warm_threads(int nthreads) { const char *e = getenv("MD_WARM_THREADS"); int v = e ? atoi(e) : 8; return v < 0 ? 0 : (v > nthreads ? nthreads : v); }

// alignments per device tile (testing knob: small values exercise tile cuts and carried reads on small inputs)
static size_t tile_reads_default(bool async) { if (const char *e = getenv("MD_TILE_READS")) { long v = atol(e); if (v > 0) return (size_t) v; } return async ? ((size_t) 1 << 17) : ((size_t) 1 << 19); }

// Device-side BGZF inflate + BAM decode (md_bam_*) is the default when the back end offers it; MD_DEVICE_DECODE=0 selects the
// multi-threaded host decoder (also used for --minConversionEfficiency, whose tiles follow the reference's chunks).
static bool device_decode_enabled(const mdh_backend *be) {
    if (!(be->bam_open && be->bam_close && be->bam_reset && be->bam_push && be->bam_get_runs && be->bam_extract_run && be->bam_mbias_run)) return false;
    const char *e = getenv("MD_DEVICE_DECODE");
    return !(e && e[0] == '0');
}
// Segment size of the device decoder.  One warp decodes one BGZF block and 148 SMs x 32 warps are resident at a time, so a
// segment of about one such wave (4 736 blocks of ~20 KB) keeps every decoder slot busy without a nearly empty second round.
static size_t device_segment_bytes() { if (const char *e = getenv("MD_SEGMENT_BYTES")) { long v = atol(e); if (v > 0) return (size_t) v; } return (size_t) 96 << 20; }

static bool pack_quals_enabled() { const char *e = getenv("MD_QUAL_PACK"); return !(e && e[0] == '0'); }

// The decode workers allocate and free multi-megabyte buffers at a high rate; glibc's default turns each of those into an
// mmap/munmap pair plus a page fault per 4 KB, all serialised on the process's address-space lock (which also slows the CUDA
// context creation running beside them).  Keep such blocks in the arenas instead.
static void tune_allocator() {
    static bool done = false; if (done) return; done = true;
    if (getenv("MD_NO_MALLOPT")) return;
    mallopt(M_MMAP_THRESHOLD, 32 << 20); mallopt(M_TRIM_THRESHOLD, 1 << 30); mallopt(M_TOP_PAD, 16 << 20);
}

namespace {
// Text formatting: reference chunks are independent (extract.c:496-507 flushes the merge state at every chunk end), so a few
// threads format chunks concurrently, each into memory, and the finished texts are appended to the output files strictly
// in chunk order (the analogue of the reference's ordered output bins, extract.c:514-535).  post() applies back-pressure.
// Buffers that cycle between the calling thread and the text stage.  A whole-genome run pushes tens of gigabytes through these
// (md_call records in, text out); allocating each one afresh means faulting in that many zeroed pages — on a cold process
// that was most of the wall clock.  A pool keeps the working set at what is in flight.
template <class T> class BufPool {
public:
    std::unique_ptr<T> get() { std::lock_guard<std::mutex> g(m_); if (free_.empty()) return std::unique_ptr<T>(new T()); std::unique_ptr<T> p = std::move(free_.back()); free_.pop_back(); return p; }
    void put(std::unique_ptr<T> p) { if (!p) return; std::lock_guard<std::mutex> g(m_); if (free_.size() < 256) free_.push_back(std::move(p)); }
private:
    std::mutex m_; std::vector<std::unique_ptr<T>> free_;
};
typedef std::vector<md_call> CallVec;

class OrderedFormatter {
public:
    // fp[k]: the output files (headers already written through them); lines are appended with pwrite on their descriptors
    OrderedFormatter(int nthreads, const ExtractOptions &o, FILE *fp[3]) : o_(o) {
        one_file_ = fp[0] && fp[1] == fp[0];                      // cytosine_report: the contexts interleave in one file
        for (int k = 0; k < 3; ++k) {
            fd_[k] = -1; off_[k] = 0;
            if (fp[k] && !(one_file_ && k)) { fflush(fp[k]); fd_[k] = fileno(fp[k]); off_[k] = (uint64_t) ftello(fp[k]); }
        }
        max_inflight_ = (uint64_t) std::max(8, 3 * std::max(1, nthreads));
        for (int i = 0; i < std::max(1, nthreads); ++i) th_.emplace_back([this] { run(); });
    }
    ~OrderedFormatter() { { std::lock_guard<std::mutex> g(m_); stop_ = true; } cv_.notify_all(); for (auto &t : th_) t.join(); }
    // a vector for the next chunk's records (recycled)
    std::unique_ptr<CallVec> call_buffer() { std::unique_ptr<CallVec> p = calls_.get(); p->clear(); return p; }
    void post(const char *chrom, std::shared_ptr<const std::string> ref, Chunk k, std::unique_ptr<CallVec> part) {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return next_seq_ - written_ < max_inflight_; });
        q_.push_back(Job{next_seq_++, chrom, std::move(ref), k, std::move(part)}); cv_.notify_all();
    }
    void drain() { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return written_ == next_seq_; }); }
    double busy_seconds() { std::lock_guard<std::mutex> g(m_); return busy_s_; }
    double write_seconds() { std::lock_guard<std::mutex> g(m_); return write_s_; }
    uint64_t n_variant_positions() { std::lock_guard<std::mutex> g(m_); return n_variant_; }
    bool failed() { std::lock_guard<std::mutex> g(m_); return failed_; }
private:
    struct Job { uint64_t seq; const char *chrom; std::shared_ptr<const std::string> ref; Chunk k; std::unique_ptr<CallVec> part; };
    struct Done { std::unique_ptr<TextBuf> buf[3]; uint64_t at[3] = {0, 0, 0}; };
    void run() {
        for (;;) {
            Job j;
            { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return stop_ || !q_.empty(); }); if (q_.empty()) return; j = std::move(q_.front()); q_.pop_front(); }
            const double t0 = now_s();
            Done d; TextBuf *m[3] = {nullptr, nullptr, nullptr};
            for (int k = 0; k < 3; ++k) {
                if (one_file_) { if (k == 0) { d.buf[0] = text_.get(); d.buf[0]->n = 0; } m[k] = d.buf[0].get(); }
                else if (fd_[k] >= 0) { d.buf[k] = text_.get(); d.buf[k]->n = 0; m[k] = d.buf[k].get(); }
            }
            uint64_t nvar;
            { ExtractWriter w(o_, m); w.process_chunk(j.chrom, *j.ref, j.k.beg, j.k.end, j.part->data(), j.part->size()); nvar = w.n_variant_positions(); }
            calls_.put(std::move(j.part)); j.ref.reset();
            // Chunks finish out of order; file offsets are handed out strictly in chunk order (the analogue of the reference's
            // ordered output bins, extract.c:514-535), after which the bytes can land in any order: every thread writes the
            // chunks IT released, with pwrite, outside the lock.
            std::vector<Done> mine;
            {
                std::lock_guard<std::mutex> g(m_);
                busy_s_ += now_s() - t0; n_variant_ += nvar;
                ready_.emplace(j.seq, std::move(d));
                for (auto it = ready_.find(next_assign_); it != ready_.end(); it = ready_.find(next_assign_)) {
                    for (int k = 0; k < 3; ++k) if (it->second.buf[k]) { it->second.at[k] = off_[k]; off_[k] += it->second.buf[k]->n; }
                    mine.push_back(std::move(it->second));
                    ready_.erase(it); ++next_assign_;
                }
            }
            bool ok = true;
            const double t1 = now_s();
            for (Done &w : mine)
                for (int k = 0; k < 3; ++k) if (w.buf[k]) {
                    const char *p = w.buf[k]->v.data(); size_t left = w.buf[k]->n; uint64_t at = w.at[k];
                    while (left) { ssize_t r = pwrite(fd_[k], p, left, (off_t) at); if (r <= 0) { ok = false; break; } p += r; left -= (size_t) r; at += (uint64_t) r; }
                    text_.put(std::move(w.buf[k]));
                }
            {
                std::lock_guard<std::mutex> g(m_);
                written_ += mine.size(); write_s_ += now_s() - t1;
                if (!ok) failed_ = true;
            }
            cv_.notify_all();
        }
    }
    const ExtractOptions &o_; int fd_[3]; uint64_t off_[3]; bool one_file_ = false; uint64_t max_inflight_ = 32;
    BufPool<TextBuf> text_; BufPool<CallVec> calls_;
    std::deque<Job> q_; std::map<uint64_t, Done> ready_; uint64_t next_seq_ = 0, next_assign_ = 0, written_ = 0, n_variant_ = 0;
    std::mutex m_; std::condition_variable cv_; bool stop_ = false, failed_ = false; double busy_s_ = 0, write_s_ = 0;
    std::vector<std::thread> th_;
};
// threads for the text stage: with the file decoded on the device the host cores have nothing else to do
static int format_threads(bool device_decode) {
    if (const char *e = getenv("MD_FORMAT_THREADS")) { int v = atoi(e); if (v > 0) return v; }
    unsigned hw = std::thread::hardware_concurrency();
    return (int) (device_decode ? std::max(2u, std::min(24u, hw > 4 ? hw - 3 : 2u)) : std::max(1u, std::min(8u, hw / 4)));
}

struct Driver {
    const mdh_backend *be; void *dev = nullptr;
    std::unique_ptr<ParallelBam> bam; std::unique_ptr<Fasta> fa; BaiIndex bai; bool have_bai = false;
    BamHeader own_hdr; uint64_t start_voff = 0;              // device-decode mode: no host decoder, just the header
    std::shared_ptr<Fragment> frag; size_t frag_i = 0;      // decode cursor shared by consecutive FragTilers
    const BamHeader *hdr = nullptr;
    BedFile bed; bool have_bed = false;                     // -l
    // a chunk the reference's workers skip because no BED region overlaps it (extract.c:353-367): nothing is written for it
    bool chunk_skipped(const Chunk &k) const { return have_bed && !bed.chunk_overlaps(k.tid, k.beg, k.end); }
    // hand the contig's regions to the device (after load_contig)
    int push_bed(uint32_t tid) {
        if (!have_bed) return 0;
        if (!be->set_bed) { fprintf(stderr, "The device back end does not implement -l.\n"); return -20; }
        const std::vector<md_bed_region> &v = bed.regions(tid);
        if (be->set_bed(dev, (int32_t) tid, v.data(), (uint32_t) v.size()) != 0) { fprintf(stderr, "device error: %s\n", be->last_error ? be->last_error() : "?"); return -20; }
        return 0;
    }
    // contig sequences are shared with the text stage, which may still be formatting a contig's last chunks when the next
    // contig is loaded (no drain between contigs)
    std::shared_ptr<std::string> cur_seq = std::make_shared<std::string>(); int cur_seq_tid = -1; bool cur_seq_ok = false;
    std::future<bool> next_ready; std::shared_ptr<std::string> next_seq; int next_tid = -1;
    // whole contig for the device and the writer; the one announced with prefetch() was read in the background meanwhile
    const std::string *fetch(uint32_t tid) {
        if ((int) tid != cur_seq_tid) {
            cur_seq_tid = (int) tid;
            if (next_ready.valid() && next_tid == (int) tid) { cur_seq_ok = next_ready.get(); cur_seq = next_seq; next_seq.reset(); }
            else { if (next_ready.valid()) next_ready.get(); cur_seq = std::make_shared<std::string>(); cur_seq_ok = tid < hdr->names.size() && fa->fetch(hdr->names[tid], *cur_seq); }
        }
        return cur_seq_ok ? cur_seq.get() : nullptr;
    }
    std::shared_ptr<const std::string> shared_ref() const { return cur_seq; }
    void prefetch(uint32_t tid) {
        if (next_ready.valid() || (int) tid == cur_seq_tid || tid >= hdr->names.size()) return;
        next_tid = (int) tid;
        next_seq = std::make_shared<std::string>();
        std::shared_ptr<std::string> dst = next_seq;
        next_ready = std::async(std::launch::async, [this, tid, dst] { return fa->fetch(hdr->names[tid], *dst); });
    }
    ~Driver() { if (next_ready.valid()) next_ready.wait(); }
    // a few bases around a chunk end (adjustBounds, common.c:477) without loading the contig; returns the contig's length
    int64_t window(uint32_t tid, int64_t start, int64_t n, std::string &w) {
        w.clear();
        if (tid >= hdr->names.size()) return 0;
        const FaiEntry *e = fa->find(hdr->names[tid]);
        if (!e) return 0;
        fa->fetch_range(*e, start, n, w);
        return e->len;
    }
    bool sought = false;
    void seek_to(int tid, uint32_t beg) {
        // One index jump to where this run starts; after that contigs and regions are visited in file order, so the decode
        // stream simply continues (its read-ahead into the next contig is kept) and the tilers skip what lies between.
        if (!have_bai || sought) return;           // no index: sequential scan, the stream only moves forward and skips earlier contigs
        bool found; uint64_t off = bai.start_offset(tid, beg, found);
        if (found && off) { bam->seek(off); frag.reset(); frag_i = 0; sought = true; }
    }
};
}  // namespace


// ---- device-decode drivers --------------------------------------------------------------------------------------
// The compressed file goes to the device in segments of whole BGZF blocks (md_bam_push); every push reports the runs of
// records per contig and one tile is requested per run.  A contig that may continue in the next segment is cut at the
// position of the run's last record (those reads, and every read reaching beyond the cut, are carried into the next tile on
// the device); a contig that ends inside the segment gets its final tile.  The reference's chunks are replayed over the
// returned md_call records exactly as in the host-decode path.
namespace {
struct ContigJob { uint32_t tid, rbeg, rend; std::vector<Chunk> chunks; };

template <class OpenContig, class Tile, class CloseContig>
int drive_segments(Driver &d, const mdh_backend *be, void *bs, const char *bamName, const std::vector<ContigJob> &jobs, OpenContig &&open_contig, Tile &&tile, CloseContig &&close_contig) {
    if (jobs.empty()) return 0;
    BgzfSegmenter seg(bamName);
    uint64_t voff = d.start_voff;
    if (d.have_bai) { bool found; uint64_t o = d.bai.start_offset((int) jobs[0].tid, jobs[0].rbeg, found); if (found && o) voff = o; }
    seg.seek(voff >> 16);
    uint32_t skip = (uint32_t)(voff & 0xffff);
    be->bam_reset(bs);
    const size_t target = device_segment_bytes();
    size_t cj = 0; bool open = false; uint32_t open_beg = 0; int rc = 0;
    std::vector<md_bam_run> runs;
    // finish job cj: its last tile comes from carried reads only (run -1); a contig that never had a record just gets its chunks written
    auto finish = [&]() -> int {
        const ContigJob &J = jobs[cj];
        int r = 0;
        if (!open) { r = open_contig(J); if (r) return r; open_beg = J.rbeg; }
        md_tile_desc td{(int32_t) J.tid, open_beg, J.rend, 0, 0};
        r = tile(J, -1, td, open);                       // nothing to compute if the contig was never opened on the device
        if (r) return r;
        r = close_contig(J);
        open = false; ++cj;
        return r;
    };
    // Segment k+1 is handed to the device (copy + inflate + record tables, on a helper thread and its own stream) before
    // the tiles of segment k are requested, when the back end offers the two-phase push; and when it offers the prefetch, the
    // compressed bytes of segment k+2 start travelling as soon as segment k+1 is being decoded.
    const bool overlapped = be->bam_push_begin && be->bam_push_end;
    const bool prefetching = overlapped && be->bam_prefetch && !getenv("MD_NO_PREFETCH");
    // large files go through page-locked staging buffers filled by a background reader (StagedSegments); small ones are
    // pushed straight from the file mapping (allocating the buffers would cost more than it saves)
    std::unique_ptr<StagedSegments> staged;
    {
        const char *e = getenv("MD_STAGE");
        const bool want = e ? e[0] != '0' : seg.remaining() >= ((size_t) 512 << 20);
        if (want && be->pinned_alloc && be->pinned_free) { staged.reset(new StagedSegments(bamName, seg.tell(), target, be->pinned_alloc, be->pinned_free)); if (!staged->staged()) staged.reset(); }
    }
    // three segments are alive at a time: k (its tiles), k+1 (being decoded), k+2 (being copied)
    struct HostSeg { const uint8_t *base = nullptr; size_t bytes = 0; std::vector<md_bgzf_block> blk; bool have = false; } hs[3];
    auto next_segment = [&](HostSeg &h) {
        Acc a_(5);
        if (!staged) { h.have = seg.next(target, h.base, h.bytes, h.blk); return; }
        StagedSegments::Seg sg;
        h.have = staged->next(sg);
        if (h.have) { h.base = sg.base; h.bytes = sg.bytes; h.blk.swap(sg.blocks); }
    };
    auto dev_fail = [&]() { fprintf(stderr, "device error: %s\n", be->last_error ? be->last_error() : "?"); return -20; };
    size_t n_seg = 0;                                      // index of the segment whose tiles are being requested: lives in hs[n_seg % 3]
    next_segment(hs[0]);
    bool pushed = false, fetched_ahead = false;            // fetched_ahead: hs[(n_seg + 1) % 3] already holds (or was asked for) segment n_seg + 1
    if (hs[0].have && overlapped) {
        if (be->bam_push_begin(bs, hs[0].base, hs[0].bytes, hs[0].blk.data(), (uint32_t) hs[0].blk.size(), skip) != 0) return dev_fail();
        pushed = true;
        if (prefetching) { next_segment(hs[1]); fetched_ahead = true; if (hs[1].have && be->bam_prefetch(bs, hs[1].base, hs[1].bytes) != 0) rc = dev_fail(); }   // (the push in flight is waited for below)
    }
    bool done = false;
    while (hs[n_seg % 3].have && !done && rc == 0) {
        HostSeg &cur = hs[n_seg % 3], &nxt = hs[(n_seg + 1) % 3], &nxt2 = hs[(n_seg + 2) % 3];
        double t0 = now_s();
        md_bam_summary sum;
        int r; { Acc a_(7); r = overlapped ? be->bam_push_end(bs, &sum) : be->bam_push(bs, cur.base, cur.bytes, cur.blk.data(), (uint32_t) cur.blk.size(), skip, &sum); }
        pushed = false; skip = 0;
        g_stats.t_decode_s += now_s() - t0;
        if (g_marks && (n_seg < 4 || n_seg % 10 == 0)) { char m[64]; snprintf(m, sizeof m, "segment %zu decoded", n_seg); mark(m); }
        if (r != 0) { rc = dev_fail(); break; }
        // cut the next segment now: whether the file ends here decides how this segment's last run is closed
        if (!fetched_ahead) next_segment(nxt);
        fetched_ahead = false;
        const bool have_next = nxt.have;
        const bool file_end = !have_next;
        if (have_next && overlapped) {
            { Acc a_(8); if (be->bam_push_begin(bs, nxt.base, nxt.bytes, nxt.blk.data(), (uint32_t) nxt.blk.size(), 0) != 0) { rc = dev_fail(); break; } }
            pushed = true;
            if (prefetching) { next_segment(nxt2); fetched_ahead = true; if (nxt2.have && be->bam_prefetch(bs, nxt2.base, nxt2.bytes) != 0) { rc = dev_fail(); break; } }
        }
        g_stats.n_records += sum.n_records;
        runs.resize(sum.n_runs);
        if (sum.n_runs) { Acc a_(9); be->bam_get_runs(bs, runs.data(), sum.n_runs); }
        for (size_t k = 0; k < runs.size() && !done && rc == 0; ++k) {
            const md_bam_run &run = runs[k];
            if (run.tid < 0) { done = true; break; }                                   // unmapped records close a sorted file
            while (cj < jobs.size() && (int32_t) jobs[cj].tid < run.tid && rc == 0) rc = finish();
            if (rc || cj >= jobs.size()) { done = true; break; }
            const ContigJob &J = jobs[cj];
            if ((int32_t) J.tid > run.tid) continue;                                   // a contig this run does not cover
            if ((int64_t) run.first_pos >= (int64_t) J.rend) { rc = finish(); if (cj >= jobs.size()) done = true; continue; }
            if (!open) { rc = open_contig(J); if (rc) break; open = true; open_beg = J.rbeg; }
            const bool may_continue = (k + 1 == runs.size()) && !file_end && (int64_t) run.last_pos < (int64_t) J.rend;
            if (may_continue) {
                uint32_t cut = (uint32_t) std::max<int64_t>(run.last_pos, (int64_t) open_beg);
                md_tile_desc td{(int32_t) J.tid, open_beg, cut, 0, 0};
                rc = tile(J, (int) k, td, true);
                open_beg = cut;
            } else {
                md_tile_desc td{(int32_t) J.tid, open_beg, J.rend, 0, 0};
                rc = tile(J, (int) k, td, true);
                if (!rc) rc = close_contig(J);
                open = false; ++cj;
                if (cj >= jobs.size()) done = true;
            }
        }
        ++n_seg;
    }
    if (pushed) { md_bam_summary dummy; be->bam_push_end(bs, &dummy); }            // a segment beyond the region was already on its way
    if (prefetching) be->bam_prefetch(bs, nullptr, 0);                             // ... and so may be a copy: it must not outlive the staging buffers
    while (rc == 0 && cj < jobs.size()) rc = finish();
    return rc;
}

std::vector<ContigJob> contig_jobs(const std::vector<Chunk> &all, size_t c0, size_t c1) {
    std::vector<ContigJob> jobs;
    for (size_t ci = c0; ci < c1; ++ci) {
        if (jobs.empty() || jobs.back().tid != all[ci].tid) { ContigJob j; j.tid = all[ci].tid; j.rbeg = all[ci].beg; j.rend = all[ci].end; jobs.push_back(j); }
        jobs.back().chunks.push_back(all[ci]); jobs.back().rend = all[ci].end;
    }
    return jobs;
}
}  // namespace

static int extract_device_decode(Driver &d, const mdh_backend *be, const char *bamName, const std::vector<Chunk> &all, size_t c0, size_t c1, OrderedFormatter &out_thread) {
    void *bs = be->bam_open(d.dev, (int32_t) d.hdr->names.size());
    if (!bs) { fprintf(stderr, "device error: %s\n", be->last_error ? be->last_error() : "?"); return -20; }
    std::vector<ContigJob> jobs = contig_jobs(all, c0, c1);
    const std::string *ref = nullptr; bool loaded = false;
    size_t next_chunk = 0;
    PodVec<md_call> tile_calls; md_call *pin_calls[2] = {nullptr, nullptr}; uint64_t pin_cap[2] = {0, 0}; uint64_t n_dev_tiles = 0; std::vector<md_call *> old_pins;
    // The hand-over below (a copy of every record + the hand-off to the text stage) runs on a helper thread, one tile behind the
    // calling thread, which meanwhile has the device build and count the next tile; two read-back buffers alternate.
    struct FeedJob { const ContigJob *J; const md_call *rec; size_t n; uint32_t upto; bool last; };
    std::mutex fm; std::condition_variable fcv; std::deque<FeedJob> fq; bool fstop = false; uint64_t f_submitted = 0, f_done = 0;
    // The records of a tile go straight from the read-back buffer into the vectors of the reference chunks they belong to (one
    // copy, located by binary search; a chunk that straddles a tile cut keeps collecting in `pending`), and every chunk that is
    // complete — its end lies at or before `done_upto`, or the contig is being closed — is handed to the text stage.
    std::unique_ptr<CallVec> pending;
    auto feed = [&](const ContigJob &J, const md_call *rec, size_t n, uint32_t done_upto, bool last) {
        const char *cname = d.hdr->names[J.tid].c_str();
        size_t i = 0;
        while (next_chunk < J.chunks.size()) {
            const Chunk k = J.chunks[next_chunk];
            const md_call *e = std::lower_bound(rec + i, rec + n, k.end, [](const md_call &c, uint32_t v) { return c.pos < v; });
            if (e != rec + i) { Acc a_(2); if (!pending) pending = out_thread.call_buffer(); pending->insert(pending->end(), rec + i, e); i = (size_t)(e - rec); }
            if (!(k.end <= done_upto || last)) break;                       // the chunk continues in the next tile
            if (!pending) pending = out_thread.call_buffer();
            if (!d.chunk_skipped(k)) { Acc a_(3); out_thread.post(cname, d.shared_ref(), k, std::move(pending)); }
            pending.reset(); ++next_chunk;
        }
    };
    std::thread feeder([&] {
        for (;;) {
            FeedJob j;
            { std::unique_lock<std::mutex> l(fm); fcv.wait(l, [&] { return fstop || !fq.empty(); }); if (fq.empty()) return; j = fq.front(); fq.pop_front(); }
            feed(*j.J, j.rec, j.n, j.upto, j.last);
            { std::lock_guard<std::mutex> g(fm); ++f_done; } fcv.notify_all();
        }
    });
    auto feed_async = [&](const ContigJob &J, const md_call *rec, size_t n, uint32_t upto, bool last) { { std::lock_guard<std::mutex> g(fm); fq.push_back(FeedJob{&J, rec, n, upto, last}); ++f_submitted; } fcv.notify_all(); };
    auto feed_wait = [&](uint64_t outstanding) { std::unique_lock<std::mutex> l(fm); fcv.wait(l, [&] { return f_submitted - f_done <= outstanding; }); };
    size_t job_i = 0;
    auto open_contig = [&](const ContigJob &J) -> int {
        { Acc a_(10); feed_wait(0); }                           // the hand-over thread is done with the previous contig's state
        { Acc a_(1); ref = d.fetch(J.tid); }
        while (job_i < jobs.size() && jobs[job_i].tid != J.tid) ++job_i;
        if (job_i + 1 < jobs.size()) d.prefetch(jobs[job_i + 1].tid);
        pending.reset(); next_chunk = 0; loaded = false;
        if (!ref) {
            fprintf(stderr, "faidx_fetch_seq returned %i while trying to fetch the sequence for tid %s:%" PRIu32 "-%" PRIu32 "!\n", -2, d.hdr->names[J.tid].c_str(), J.chunks.front().beg, J.chunks.front().end);
            fprintf(stderr, "Note that the output will be truncated!\n");
            return 0;
        }
        return 0;
    };
    auto tile = [&](const ContigJob &J, int run, const md_tile_desc &td, bool on_device) -> int {
        if (!ref) return 0;
        const md_call *got = nullptr; size_t n_got = 0;
        if (on_device) {
            if (!loaded) { Acc a_(6); if (be->load_contig(d.dev, (int32_t) J.tid, ref->data(), (uint32_t) ref->size()) != 0) { fprintf(stderr, "device error: %s\n", be->last_error ? be->last_error() : "?"); return -20; } if (d.push_bed(J.tid)) return -20; loaded = true; }
            md_tile_desc t = td; if (t.end > ref->size()) t.end = (uint32_t) ref->size(); if (t.beg > t.end) t.beg = t.end;
            const uint64_t cap = (uint64_t)(t.end - t.beg) + 16;
            // the records come back into page-locked memory when the back end offers it (a pageable target halves the D2H rate and
            // makes the copy synchronous); the buffer only ever grows
            md_call *dst = nullptr;
            const int pb = (int)(n_dev_tiles++ & 1);
            { Acc a_(10); feed_wait(1); }                       // the tile before last has left this buffer
            if (be->pinned_alloc && be->pinned_free) {
                Acc a_(11);
                // (an outgrown buffer is released at the end of the run: freeing page-locked memory waits for the device, i.e. for the segment being decoded)
                if (cap > pin_cap[pb]) { if (pin_calls[pb]) old_pins.push_back(pin_calls[pb]); pin_cap[pb] = 2 * cap; pin_calls[pb] = (md_call *) be->pinned_alloc(pin_cap[pb] * sizeof(md_call)); if (!pin_calls[pb]) pin_cap[pb] = 0; }
                dst = pin_calls[pb];
            }
            if (!dst) { feed_wait(0); tile_calls.clear(); tile_calls.grow(cap); dst = tile_calls.data(); }
            md_tile_stats st;
            double t0 = now_s();
            int r = be->bam_extract_run(bs, run, &t, J.rend, dst, cap, &st);
            g_stats.t_device_s += now_s() - t0;
            if (r != 0) { fprintf(stderr, "device error: %s\n", be->last_error ? be->last_error() : "?"); return -20; }
            got = dst; n_got = (size_t) st.n_calls;
            g_stats.n_calls += st.n_calls; g_stats.n_tiles++;
        }
        feed_async(J, got, n_got, td.end, false);
        return 0;
    };
    auto close_contig = [&](const ContigJob &J) -> int {
        if (ref) feed_async(J, nullptr, 0, J.rend, true);
        { Acc a_(10); feed_wait(0); }
        if (loaded) { Acc a_(6); be->drop_contig(d.dev, (int32_t) J.tid); }
        loaded = false;
        return 0;
    };
    int rc = drive_segments(d, be, bs, bamName, jobs, open_contig, tile, close_contig);
    feed_wait(0);
    { std::lock_guard<std::mutex> g(fm); fstop = true; } fcv.notify_all();
    feeder.join();
    be->bam_close(bs);
    for (int k = 0; k < 2; ++k) if (pin_calls[k]) be->pinned_free(pin_calls[k]);
    for (md_call *p : old_pins) be->pinned_free(p);
    return rc;
}


static int mbias_device_decode(Driver &d, const mdh_backend *be, const char *bamName, const std::vector<Chunk> &all, size_t c0, size_t c1) {
    void *bs = be->bam_open(d.dev, (int32_t) d.hdr->names.size());
    if (!bs) { fprintf(stderr, "device error: %s\n", be->last_error ? be->last_error() : "?"); return -20; }
    std::vector<ContigJob> jobs = contig_jobs(all, c0, c1);
    const std::string *ref = nullptr; bool loaded = false; bool truncated = false;
    size_t job_i = 0;
    auto open_contig = [&](const ContigJob &J) -> int {
        if (truncated) { ref = nullptr; return 0; }
        ref = d.fetch(J.tid);
        while (job_i < jobs.size() && jobs[job_i].tid != J.tid) ++job_i;
        if (job_i + 1 < jobs.size()) d.prefetch(jobs[job_i + 1].tid);
        loaded = false;
        if (!ref) {
            fprintf(stderr, "faidx_fetch_seq returned %i while trying to fetch the sequence for tid %s:%" PRIu32 "-%" PRIu32 "!\n", -2, d.hdr->names[J.tid].c_str(), J.chunks.front().beg, J.chunks.front().end);
            fprintf(stderr, "Note that the output will be truncated!\n");
            truncated = true;                                  // MBias.c:152 returns from the worker
        }
        return 0;
    };
    auto tile = [&](const ContigJob &J, int run, const md_tile_desc &td, bool on_device) -> int {
        if (!ref || !on_device) return 0;
        if (!loaded) {
            std::vector<uint32_t> bounds; bounds.push_back(J.chunks.front().beg); for (auto &k : J.chunks) bounds.push_back(k.end);
            if (be->load_contig(d.dev, (int32_t) J.tid, ref->data(), (uint32_t) ref->size()) != 0 ||
                be->set_mbias_chunks(d.dev, (int32_t) J.tid, bounds.data(), (uint32_t) bounds.size() - 1) != 0) { fprintf(stderr, "device error: %s\n", be->last_error ? be->last_error() : "?"); return -20; }
            if (d.push_bed(J.tid)) return -20;
            loaded = true;
        }
        md_tile_desc t = td; if (t.end > ref->size()) t.end = (uint32_t) ref->size(); if (t.beg > t.end) t.beg = t.end;
        md_tile_stats st;
        double t0 = now_s();
        int r = be->bam_mbias_run(bs, run, &t, J.rend, &st);
        g_stats.t_device_s += now_s() - t0;
        if (r != 0) { fprintf(stderr, "device error: %s\n", be->last_error ? be->last_error() : "?"); return -20; }
        g_stats.n_tiles++;
        return 0;
    };
    auto close_contig = [&](const ContigJob &J) -> int { if (loaded) be->drop_contig(d.dev, (int32_t) J.tid); loaded = false; return 0; };
    int rc = drive_segments(d, be, bs, bamName, jobs, open_contig, tile, close_contig);
    be->bam_close(bs);
    return rc;
}

// One process per GPU: shard `rank` of `world` takes a contiguous run [c0, c1) of the reference's chunks.  The runs are
// balanced by the amount of ALIGNMENT DATA they hold — compressed file bytes between the chunks' positions in the BAM
// index's linear index — not by base pairs: a targeted panel or an unevenly covered genome otherwise leaves most GPUs
// idle (the reference's dynamic chunk cursor, extract.c:327-350, balances itself the same way).  Without an
// index, or when it shows no data at all, the split falls back to base pairs.
static void shard_chunks(const std::vector<Chunk> &all, const BaiIndex *bai, int rank, int world, size_t &c0, size_t &c1) {
    std::vector<uint64_t> w(all.size());
    uint64_t total = 0;
    if (bai) {
        for (size_t k = 0; k < all.size(); ++k) {
            const uint64_t a = bai->file_pos((int) all[k].tid, all[k].beg);
            uint64_t b = bai->file_pos((int) all[k].tid, all[k].end);
            if (k + 1 < all.size() && all[k + 1].tid != all[k].tid) { const uint64_t nb = bai->file_pos((int) all[k + 1].tid, all[k + 1].beg); if (nb > a) b = std::max(b, nb); }   // a contig's last chunk reaches to the next contig's data
            w[k] = (a && b > a) ? b - a : 0; total += w[k];
        }
    }
    if (total == 0) { for (size_t k = 0; k < all.size(); ++k) { w[k] = all[k].end - all[k].beg; total += w[k]; } }
    else for (size_t k = 0; k < all.size(); ++k) { w[k] += 1; total += 1; }     // empty chunks still cost something, and every prefix sum moves
    const uint64_t lo = total * (uint64_t) rank / (uint64_t) world, hi = total * (uint64_t)(rank + 1) / (uint64_t) world;
    uint64_t acc = 0;
    c0 = c1 = all.size();
    for (size_t k = 0; k < all.size(); ++k) { if (acc >= lo && c0 == all.size()) c0 = k; if (acc >= hi) { c1 = k; break; } acc += w[k]; }
    if (c0 > c1) c0 = c1;
}

extern "C" int mdh_extract_main(int argc, char *argv[], const mdh_backend *be) {
    ExtractOptions o;
    char *opref = nullptr; const char *reg = nullptr, *bedName = nullptr, *bwName = nullptr, *bbmName = nullptr;
    int c, nThreads = 1, keepStrand = 0; double minConvEff = 0.0; bool threads_given = false;
    double t_start = now_s(); g_t0 = t_start; g_marks = getenv("MD_TIMING") != nullptr;
    tune_allocator();
    memset(&g_stats, 0, sizeof g_stats); memset(g_acc, 0, sizeof g_acc);

    static struct option lopts[] = {
        {"opref", 1, NULL, 'o'}, {"fraction", 0, NULL, 'f'}, {"counts", 0, NULL, 'c'}, {"logit", 0, NULL, 'm'}, {"minDepth", 1, NULL, 'd'},
        {"noCpG", 0, NULL, 1}, {"CHG", 0, NULL, 2}, {"CHH", 0, NULL, 3}, {"keepDupes", 0, NULL, 4}, {"keepSingleton", 0, NULL, 5}, {"keepDiscordant", 0, NULL, 6},
        {"OT", 1, NULL, 7}, {"OB", 1, NULL, 8}, {"CTOT", 1, NULL, 9}, {"CTOB", 1, NULL, 10}, {"mergeContext", 0, NULL, 11}, {"methylKit", 0, NULL, 12},
        {"nOT", 1, NULL, 13}, {"nOB", 1, NULL, 14}, {"nCTOT", 1, NULL, 15}, {"nCTOB", 1, NULL, 16}, {"minOppositeDepth", 1, NULL, 17}, {"maxVariantFrac", 1, NULL, 18},
        {"chunkSize", 1, NULL, 19}, {"keepStrand", 0, NULL, 20}, {"cytosine_report", 0, NULL, 21}, {"minConversionEfficiency", 1, NULL, 22}, {"ignoreNH", 0, NULL, 23},
        {"ignoreFlags", 1, NULL, 'F'}, {"requireFlags", 1, NULL, 'R'}, {"help", 0, NULL, 'h'}, {"version", 0, NULL, 'v'},
        {"mappability", 1, NULL, 'M'}, {"mappabilityThreshold", 1, NULL, 't'}, {"minMappableBases", 1, NULL, 'b'}, {"outputBBMFile", 1, NULL, 'O'},
        {"outputBBMFileName", 1, NULL, 'N'}, {"mappabilityBBM", 1, NULL, 'B'},
        {"shardRank", 1, NULL, 200}, {"shardWorld", 1, NULL, 201}, {0, 0, NULL, 0}};   // 200/201: one process per GPU, see api.extract_sharded
    int shardRank = 0, shardWorld = 1;
    optind = 0;   // re-entrant use from one process
    while ((c = getopt_long(argc, argv, "hvq:p:r:l:o:D:f:c:m:d:F:R:@:M:t:b:ON:B:", lopts, NULL)) >= 0) {
        switch (c) {
        case 'h': extract_usage(); return 0;
        case 'v': printf("%s (B200 build; no HTSlib)\n", MD_VERSION); return 0;
        case 'o': free(opref); opref = strdup(optarg); break;
        case 'D': break;
        case 'd': o.minDepth = atoi(optarg); if (o.minDepth < 1) { fprintf(stderr, "Error, the minimum depth must be at least 1!\n"); return 1; } break;
        case 'r': reg = optarg; break;
        case 'l': bedName = optarg; break;
        case 1: o.core.keepCpG = 0; break;
        case 2: o.core.keepCHG = 1; break;
        case 3: o.core.keepCHH = 1; break;
        case 4: o.core.keepDupes = 1; break;
        case 5: o.core.keepSingleton = 1; break;
        case 6: o.core.keepDiscordant = 1; break;
        case 7: case 8: case 9: case 10: parse_bounds(optarg, o.core.bounds, c - 7); break;
        case 11: o.merge = 1; break;
        case 12: o.methylKit = 1; break;
        case 13: case 14: case 15: case 16: parse_bounds(optarg, o.core.absoluteBounds, c - 13); break;
        case 17: o.core.minOppositeDepth = atoi(optarg); break;
        case 18: o.core.maxVariantFrac = atof(optarg); break;
        case 19: o.chunkSize = strtoul(optarg, NULL, 10); if (o.chunkSize < 1) { fprintf(stderr, "Error: The chunk size must be at least 1!\n"); return 1; } break;
        case 20: keepStrand = 1; break;
        case 21: o.cytosine_report = 1; break;
        case 22: minConvEff = atof(optarg); break;
        case 23: o.core.ignoreNH = 1; break;
        case 200: shardRank = atoi(optarg); break;
        case 201: shardWorld = atoi(optarg); if (shardWorld < 1) shardWorld = 1; break;
        case 'M': bwName = optarg; break;
        case 't': case 'b': case 'O': case 'N': break;
        case 'B': bbmName = optarg; break;
        case 'F': o.core.ignoreFlags = atoi(optarg); break;       // atoi: "0xD00" parses as 0 (tests/test.py:68)
        case 'R': o.core.requireFlags = atoi(optarg); break;
        case 'q': o.core.minMapq = atoi(optarg); break;
        case 'p': o.core.minPhred = atoi(optarg); break;
        case 'm': o.logit = 1; break;
        case 'f': o.fraction = 1; break;
        case 'c': o.counts = 1; break;
        case '@': nThreads = atoi(optarg); threads_given = true; break;
        case '?': default: fprintf(stderr, "Invalid option '%c'\n", c); extract_usage(); return 1;
        }
    }
    if (argc == 1) { extract_usage(); return 0; }
    if (argc - optind < 2) { fprintf(stderr, "You must supply a reference genome in fasta format and an input BAM file!!!\n"); extract_usage(); return -1; }
    if (o.core.minPhred < 1) { fprintf(stderr, "-p %i is invalid. resetting to 1, which is the lowest possible value.\n", o.core.minPhred); o.core.minPhred = 1; }
    if (o.core.minMapq < 0) { fprintf(stderr, "-q %i is invalid. Resetting to 0, which is the lowest possible value.\n", o.core.minMapq); o.core.minMapq = 0; }
    if (o.core.keepDupes > 0 && (o.core.ignoreFlags & 0x400)) o.core.ignoreFlags -= 0x400;
    if (o.fraction + o.counts + o.logit + o.methylKit + o.cytosine_report > 1) {
        fprintf(stderr, "More than one of --fraction, --counts, --methylKit, --cytosine_report and --logit were specified. These are mutually exclusive.\n"); extract_usage(); return 1; }
    if (o.methylKit + o.merge == 2) { fprintf(stderr, "--mergeContext and --methylKit are mutually exclusive.\n"); extract_usage(); return 1; }
    if (o.cytosine_report + o.merge == 2) { fprintf(stderr, "--mergeContext and --cytosine_report are mutually exclusive.\n"); extract_usage(); return 1; }
    if (!(o.core.keepCpG + o.core.keepCHG + o.core.keepCHH)) {
        fprintf(stderr, "You haven't specified any metrics to output!\nEither don't use the --noCpG option or specify --CHG and/or --CHH.\n"); return -1; }
    if (bwName || bbmName) {
        fprintf(stderr, "This B200 build of the extract path does not implement -M/-B.\n"); return 1; }
    o.core.minConversionEfficiency = (float) minConvEff;            // Config.minConversionEfficiency is a float (MethylDackel.h:110)

    Driver d; d.be = be;
    const char *fastaName = argv[optind], *bamName = argv[optind + 1];
    // A query name that occurs more than twice among the admitted records (secondary / supplementary alignments let in by -F)
    // pairs up alternately in file order WITHIN each of the reference's chunks (a fresh hash per chunk, extract.c:393,536), so
    // which records merge depends on the chunk grid.  Tiles then follow the chunks exactly (one tile per chunk, host decoder);
    // with the default -F 0xF00 names occur at most twice and the tiling is free.
    const bool chunk_exact = (o.core.ignoreFlags & 0x900) != 0x900;
    const bool dev_decode = device_decode_enabled(be) && !(minConvEff > 0.0) && !chunk_exact;
    try {
        if (dev_decode) { BgzfReader rd(bamName); d.own_hdr = read_bam_header(rd); d.start_voff = rd.tell(); d.hdr = &d.own_hdr; }
        else { const int nt = decode_threads(nThreads, threads_given); d.bam.reset(new ParallelBam(bamName, nt, warm_threads(nt))); d.hdr = &d.bam->header(); }
    } catch (std::exception &e) { fprintf(stderr, "Couldn't open %s for reading!\n", bamName); return -4; }
    d.have_bai = load_bai(bamName, d.bai);
    try { d.fa.reset(new Fasta(fastaName)); }
    catch (std::exception &e) { fprintf(stderr, "Couldn't open the index for %s!\n", fastaName); return -4; }
    // the device context takes a noticeable fraction of a second to come up: create it while the host opens files,
    // loads the first contig and starts decoding
    mark("files open");
    std::future<void *> dev_future = std::async(std::launch::async, [be, &o] { void *p = be->create(be->factory_user, &o.core); mark("device context ready"); return p; });
    struct DevJoin { std::future<void *> &f; const mdh_backend *be; bool taken = false; ~DevJoin() { if (!taken && f.valid()) { void *p = f.get(); if (p) be->destroy(p); } } } dev_join{dev_future, be};

    // output files, extract.c:1344-1439
    if (opref == NULL) {
        opref = strdup(bamName);
        char *p = strrchr(opref, '.');
        if (p != NULL) *p = '\0';
        fprintf(stderr, "writing to prefix:'%s'\n", opref);
    }
    FILE *fp[3] = {nullptr, nullptr, nullptr};
    std::string pre(opref);
    const char *mid = o.fraction ? ".meth.bedGraph" : o.counts ? ".counts.bedGraph" : o.logit ? ".logit.bedGraph" : o.methylKit ? ".methylKit" : ".bedGraph";
    const std::string shardSuffix = shardWorld > 1 ? ".shard" + std::to_string(shardRank) : std::string();
    if (o.cytosine_report) {
        fp[0] = fopen((pre + ".cytosine_report.txt" + shardSuffix).c_str(), "w"); fp[1] = fp[2] = fp[0];
        if (!fp[0]) { fprintf(stderr, "Couldn't open the output CpG metrics file for writing! Insufficient permissions?\n"); free(opref); return -3; }
    } else {
        static const char *ctxName[3] = {"CpG", "CHG", "CHH"};
        int keep[3] = {o.core.keepCpG, o.core.keepCHG, o.core.keepCHH};
        for (int k = 0; k < 3; ++k) if (keep[k]) {
            fp[k] = fopen((pre + "_" + ctxName[k] + mid + shardSuffix).c_str(), "w");
            if (!fp[k]) { fprintf(stderr, "Couldn't open the output %s metrics file for writing! Insufficient permissions?\n", ctxName[k]); free(opref); return -3; }
            if (shardRank == 0) {
                if (o.methylKit) fprintf(fp[k], "chrBase\tchr\tbase\tstrand\tcoverage\tfreqC\tfreqT\n");
                else ExtractWriter::print_header(fp[k], ctxName[k], opref, o);
            }
        }
    }
    // -r, extract.c:1441-1468
    uint32_t gTid = 0, gPos = 0, gEnd = 0;
    if (reg) {
        int s, e, nl = parse_region(reg, &s, &e);
        if (nl < 0) { fprintf(stderr, "Could not parse the specified region!\n"); return -4; }
        int tid = d.hdr->name2tid(std::string(reg, (size_t) nl));
        if (tid < 0) { fprintf(stderr, "%s did not match a known chromosome/contig name!\n", reg); return -6; }
        gTid = (uint32_t) tid;
        if (s > 0) gPos = (uint32_t) s;
        if (e > 0) gEnd = (uint32_t) e;
        if (gEnd > d.hdr->lens[gTid]) gEnd = d.hdr->lens[gTid];
    }

    if (bedName) {                                                  // extract.c:1469-1477
        if (!d.bed.load(bedName, d.hdr->names, d.hdr->lens, keepStrand != 0)) { fprintf(stderr, "There was an error while reading in your BED file!\n"); return 1; }
        d.have_bed = true;
    }
    for (int k = 0; k < 3; ++k) if (fp[k] && (k == 0 || fp[k] != fp[0])) setvbuf(fp[k], nullptr, _IOFBF, 4 << 20);
    OrderedFormatter out_thread(std::max(2, format_threads(dev_decode) / shardWorld), o, fp);     // shards of one job share the host's cores
    int rc = 0;
    {
        // the reference's chunk list (extract.c:325-350), enumerated up front; a shard takes a contiguous,
        // length-balanced run of chunks (one process per GPU, no exchange between shards — SURVEY 8e)
        std::vector<Chunk> all;
        {
            ChunkCursor cursor(d.hdr->lens, o.chunkSize, gTid, gPos, gEnd);
            auto fetch = [&](uint32_t t, int64_t start, int64_t cnt, std::string &w) { return d.window(t, start, cnt, w); };
            Chunk ch;
            while (cursor.next(ch, fetch)) all.push_back(ch);
        }
        size_t c0 = 0, c1 = all.size();
        if (shardWorld > 1) shard_chunks(all, d.have_bai ? &d.bai : nullptr, shardRank, shardWorld, c0, c1);
        if (dev_decode) {
            d.dev = dev_future.get(); dev_join.taken = true;
            mark("device joined");
            if (!d.dev) { fprintf(stderr, "Could not initialise the device back end: %s\n", be->last_error ? be->last_error() : "?"); return -20; }
            rc = extract_device_decode(d, be, bamName, all, c0, c1, out_thread);
        } else {
        // tile ring: with an asynchronous back end, tiles live in page-locked memory and up to two are in flight
        // while the next one is being decoded (decode || H2D || kernels || D2H)
        const bool use_async = be->submit_tile && be->collect_tile;
        TileAlloc pin; if (use_async && be->pinned_alloc && be->pinned_free) { pin.alloc = be->pinned_alloc; pin.release = be->pinned_free; }
        std::vector<std::unique_ptr<SoaTile>> ring;
        for (int k = 0; k < (use_async ? 3 : 1); ++k) ring.emplace_back(new SoaTile(use_async ? &pin : nullptr));
        const size_t tile_reads = chunk_exact ? (size_t) 1 << 40 : tile_reads_default(use_async);
        if (use_async && !chunk_exact) for (auto &t : ring) t->reserve_for(tile_reads + tile_reads / 8, 160);
        // phred column re-encoded as 2/4-bit codes when the tile's alphabet allows (md_reads_soa::qual_bits)
        const bool pack_q = pack_quals_enabled();
        SoaTile carry;
        std::vector<md_call> calls; size_t calls_head = 0;
        PodVec<md_call> tile_calls;                              // collect target (plain capacity, never value-initialised)
        mark("tile ring reserved");
        d.dev = dev_future.get(); dev_join.taken = true;
        if (d.bam) d.bam->set_active_threads(d.bam->threads());
        mark("device joined");
        if (!d.dev) { fprintf(stderr, "Could not initialise the device back end: %s\n", be->last_error ? be->last_error() : "?"); return -20; }
        size_t ci = c0;
        while (ci < c1 && rc == 0) {
            // all chunks of this shard that lie on one contig
            uint32_t tid = all[ci].tid;
            std::vector<Chunk> chunks;
            while (ci < c1 && all[ci].tid == tid) chunks.push_back(all[ci++]);
            double t_ld = now_s();
            const std::string *ref = d.fetch(tid);
            if (ci < c1) d.prefetch(all[ci].tid);                  // the next contig is read from the FASTA while this one is processed
            if (!ref) {
                fprintf(stderr, "faidx_fetch_seq returned %i while trying to fetch the sequence for tid %s:%" PRIu32 "-%" PRIu32 "!\n", -2, d.hdr->names[tid].c_str(), chunks.front().beg, chunks.front().end);
                fprintf(stderr, "Note that the output will be truncated!\n");
                continue;
            }
            uint32_t rbeg = chunks.front().beg, rend = chunks.back().end;
            if (rend > ref->size()) rend = (uint32_t) ref->size();
            if (be->load_contig(d.dev, (int32_t) tid, ref->data(), (uint32_t) ref->size()) != 0) { fprintf(stderr, "device error: %s\n", be->last_error ? be->last_error() : "?"); rc = -20; break; }
            if ((rc = d.push_bed(tid)) != 0) break;
            g_acc[1] += now_s() - t_ld;
            mark("contig loaded");
            d.seek_to((int) tid, rbeg);
            calls.clear(); calls_head = 0; carry.clear();
            size_t next_chunk = 0;
            // With --minConversionEfficiency the verdict on an alignment depends on the reference chunk it is looked at in
            // (computeConversionEfficiency only sees that chunk's window, common.c:363,378), so tiles are cut at chunk ends
            // and carry the chunk window; otherwise the whole run of chunks is one region.
            const bool per_chunk = o.core.minConversionEfficiency > 0.0f || chunk_exact;
            std::vector<Chunk> regions;
            if (per_chunk) regions = chunks; else regions.push_back(Chunk{tid, rbeg, rend});
            // completed tiles arrive in order; hand every finished reference chunk to the writer
            auto absorb = [&](uint32_t done_upto, bool last) {
                while (next_chunk < chunks.size() && (chunks[next_chunk].end <= done_upto || last)) {
                    const Chunk k = chunks[next_chunk];
                    size_t a = calls_head; while (a < calls.size() && calls[a].pos < k.beg) ++a;
                    size_t b = a; while (b < calls.size() && calls[b].pos < k.end) ++b;
                    const char *cname = d.hdr->names[tid].c_str();
                    if (!d.chunk_skipped(k)) { std::unique_ptr<CallVec> part = out_thread.call_buffer(); part->assign(calls.begin() + (ptrdiff_t) a, calls.begin() + (ptrdiff_t) b); out_thread.post(cname, d.shared_ref(), k, std::move(part)); }
                    calls_head = b; ++next_chunk;
                }
                if (calls_head > (1u << 20)) { calls.erase(calls.begin(), calls.begin() + (ptrdiff_t) calls_head); calls_head = 0; }
            };
            struct Flight { int ticket; int slot; uint32_t end; uint64_t cap; };
            std::vector<Flight> flight;
            auto collect_one = [&]() -> int {
                Flight f = flight.front(); flight.erase(flight.begin());
                { Acc a_(4); tile_calls.clear(); tile_calls.grow(f.cap); }
                md_tile_stats st;
                double t0 = now_s();
                int r = be->collect_tile(d.dev, f.ticket, tile_calls.data(), f.cap, &st);
                g_stats.t_device_s += now_s() - t0;
                if (r != 0) { fprintf(stderr, "device error: %s\n", be->last_error ? be->last_error() : "?"); return -20; }
                { Acc a_(2); calls.insert(calls.end(), tile_calls.data(), tile_calls.data() + (ptrdiff_t) st.n_calls);
                g_stats.n_calls += st.n_calls;
                absorb(f.end, false); }
                return 0;
            };
            size_t rk = 0;
            for (size_t ri = 0; ri < regions.size() && rc == 0; ++ri) {
            const uint32_t gbeg = regions[ri].beg, gend = std::min<uint32_t>(regions[ri].end, (uint32_t) ref->size());
            const uint32_t ce_beg = per_chunk ? (gbeg > 1 ? gbeg - 2 : 0) : 0, ce_end = per_chunk ? (uint32_t) std::min<uint64_t>((uint64_t) regions[ri].end + 11, ref->size()) : 0;
            FragTiler tiler(*d.bam, d.frag, d.frag_i, (int) tid, gbeg, gend, tile_reads); tiler.set_pack_quals(pack_q);
            for (;;) {
                SoaTile &tile = *ring[rk % ring.size()];
                // the ring slot we are about to refill must have been collected
                if (use_async) while (!flight.empty() && flight.size() >= ring.size() - 0 && rc == 0) rc = collect_one();
                if (rc) break;
                double t0 = now_s();
                bool got = tiler.next(tile, carry);
                g_stats.t_decode_s += now_s() - t0;
                if (!got) break;
                if (tile.n() == 0) { if (!use_async) absorb(tile.end, false); continue; }
                md_reads_soa v = tile.view(); md_tile_desc td{(int32_t) tid, tile.beg, tile.end, ce_beg, ce_end};
                uint64_t cap = (uint64_t)(tile.end - tile.beg) + 16;
                g_stats.n_records += tile.n(); g_stats.n_tiles++;
                if (use_async) {
                    t0 = now_s();
                    int ticket = be->submit_tile(d.dev, &td, &v);
                    g_stats.t_device_s += now_s() - t0;
                    if (ticket < 0) { fprintf(stderr, "device error: %s\n", be->last_error ? be->last_error() : "?"); rc = -20; break; }
                    flight.push_back(Flight{ticket, (int)(rk % ring.size()), tile.end, cap});
                    ++rk;
                } else {
                    md_tile_stats st;
                    tile_calls.clear(); tile_calls.grow(cap);
                    t0 = now_s();
                    int r = be->extract_tile(d.dev, &td, &v, tile_calls.data(), cap, &st);
                    g_stats.t_device_s += now_s() - t0;
                    if (r != 0) { fprintf(stderr, "device error: %s\n", be->last_error ? be->last_error() : "?"); rc = -20; break; }
                    calls.insert(calls.end(), tile_calls.data(), tile_calls.data() + (ptrdiff_t) st.n_calls);
                    g_stats.n_calls += st.n_calls;
                    absorb(tile.end, false);
                }
            }
            }   // regions
            mark("last tile submitted");
            while (rc == 0 && !flight.empty()) rc = collect_one();
            if (rc == 0) absorb(rend, true);
            mark("contig written");
            be->drop_contig(d.dev, (int32_t) tid);
        }
        }   // host decode
    }
    out_thread.drain();
    if (out_thread.failed()) { fprintf(stderr, "Couldn't write the output file(s)! Disk full?\n"); if (rc == 0) rc = -3; }
    g_stats.t_format_s = out_thread.busy_seconds();
    if (g_marks) fprintf(stderr, "[md-timing] text stage: formatting %.3f s, pwrite %.3f s (summed over its threads)\n", out_thread.busy_seconds(), out_thread.write_seconds());
    if (g_marks) fprintf(stderr, "[md-timing] calling thread: record append / phred packing %.3f, contig load %.3f, chunk copy %.3f, waiting for the text stage %.3f, buffer upkeep %.3f, waiting for the next file segment %.3f, contig to / from the device %.3f\n", g_acc[0], g_acc[1], g_acc[2], g_acc[3], g_acc[4], g_acc[5], g_acc[6]);
    if (g_marks && dev_decode) fprintf(stderr, "[md-timing] device-decode driver (calling thread): waiting for a segment's decode %.3f, handing over the next segment %.3f, run table %.3f, waiting for the hand-over thread %.3f, growing the read-back buffers %.3f\n", g_acc[7], g_acc[8], g_acc[9], g_acc[10], g_acc[11]);
    if (g_marks && d.bam) fprintf(stderr, "[md-timing] record chains: %zu jobs adopted from the inflating worker, %zu walked by the stitcher\n", d.bam->jobs_adopted(), d.bam->jobs_walked());
    be->destroy(d.dev);
    mark("device destroyed");
    g_stats.n_variant_positions = out_thread.n_variant_positions();
    if (g_stats.n_variant_positions && shardWorld == 1) printf("%" PRIu64 " positions were excluded due to likely being variants.\n", (uint64_t) g_stats.n_variant_positions);
    if (o.cytosine_report) { if (fp[0]) fclose(fp[0]); }
    else for (int k = 0; k < 3; ++k) if (fp[k]) fclose(fp[k]);
    free(opref);
    g_stats.t_total_s = now_s() - t_start;
    return rc;
}

// ---- perRead ------------------------------------------------------------------------------------------------------------
// `MethylDackel perRead [opts] <ref.fa> <aln.bam>` (perRead.c:205-464): per-alignment CpG methylation.  The host walks the
// coordinate-sorted file once, hands batches of alignments (SoA, plain phreds) to md_per_read_tile and prints one line per
// reported alignment in file order, which is the order the reference's chunk-ordered output amounts to.
static void perread_usage() {
    fprintf(stderr,
"\nUsage: MethylDackel perRead [OPTIONS] <ref.fa> <input>\n\n"
"B200 build of the perRead path: the average CpG methylation level of every read.\n"
"Output columns: read name, chromosome, position, CpG methylation (%%), number of informative bases.\n"
"Options (as MethylDackel 0.6.1): -q INT (10)  -p INT (5)  -r STR  -l FILE  --keepStrand  -o STR  -F/--ignoreFlags INT (0)\n"
"  -R/--requireFlags INT (0)  -@ INT  --chunkSize INT (1000000)  -h/--help  -v/--version\n"
"Note that this program will produce incorrect values for alignments spanning more than 10kb.\n");
}

extern "C" int mdh_perread_main(int argc, char *argv[], const mdh_backend *be) {
    md_config cfg; memset(&cfg, 0, sizeof cfg);
    cfg.keepCpG = 1; cfg.minMapq = 10; cfg.minPhred = 5;                    // perRead.c:215-230
    unsigned long chunkSize = 1000000;
    const char *reg = nullptr, *bedName = nullptr, *oname = nullptr; int c, keepStrand = 0;
    static struct option lopts[] = {{"help", 0, NULL, 'h'}, {"version", 0, NULL, 'v'}, {"chunkSize", 1, NULL, 19}, {"keepStrand", 0, NULL, 20},
                                    {"ignoreFlags", 1, NULL, 'F'}, {"requireFlags", 1, NULL, 'R'}, {0, 0, NULL, 0}};
    FILE *ofile = stdout;
    optind = 0;
    while ((c = getopt_long(argc, argv, "hvq:p:o:@:r:l:F:R:", lopts, NULL)) >= 0) {
        switch (c) {
        case 'h': perread_usage(); return 0;
        case 'v': printf("%s (B200 build; no HTSlib)\n", MD_VERSION); return 0;
        case 'o': oname = optarg; if ((ofile = fopen(optarg, "w")) == NULL) { fprintf(stderr, "Couldn't open %s for writing\n", optarg); return 2; } break;
        case 'q': cfg.minMapq = atoi(optarg); break;
        case 'p': cfg.minPhred = atoi(optarg); break;
        case '@': break;
        case 'r': reg = optarg; break;
        case 'l': bedName = optarg; break;
        case 'F': cfg.ignoreFlags = atoi(optarg); break;
        case 'R': cfg.requireFlags = atoi(optarg); break;
        case 19: chunkSize = strtoul(optarg, NULL, 10); if (chunkSize < 1) { fprintf(stderr, "Error: The chunk size must be at least 1!\n"); return 1; } break;
        case 20: keepStrand = 1; break;
        default: fprintf(stderr, "Invalid option '%c'\n", c); perread_usage(); return 1;
        }
    }
    (void) oname;
    if (argc == 1) { perread_usage(); return 0; }
    if (argc - optind != 2) { fprintf(stderr, "You must supply a reference genome in fasta format and a BAM or CRAM file\n"); perread_usage(); return -1; }
    if (cfg.minPhred < 1) { fprintf(stderr, "-p %i is invalid. resetting to 1, which is the lowest possible value.\n", cfg.minPhred); cfg.minPhred = 1; }
    if (cfg.minMapq < 0) { fprintf(stderr, "-q %i is invalid. Resetting to 0, which is the lowest possible value.\n", cfg.minMapq); cfg.minMapq = 0; }
    if (!be->per_read_tile) { fprintf(stderr, "The device back end does not implement perRead.\n"); return -20; }
    const char *fastaName = argv[optind], *bamName = argv[optind + 1];
    std::unique_ptr<Fasta> fa;
    try { fa.reset(new Fasta(fastaName)); } catch (std::exception &e) { fprintf(stderr, "Couldn't open the index for %s!\n", fastaName); perread_usage(); return -2; }
    std::unique_ptr<BamStream> bs;
    try { bs.reset(new BamStream(bamName)); } catch (std::exception &e) { fprintf(stderr, "Couldn't open %s for reading!\n", bamName); return -4; }
    const BamHeader &hdr = bs->header();
    BaiIndex bai; const bool have_bai = load_bai(bamName, bai);
    uint32_t gTid = 0, gPos = 0, gEnd = 0;
    if (reg) {                                                              // perRead.c:395-420
        int s, e, nl = parse_region(reg, &s, &e);
        if (nl < 0) { fprintf(stderr, "Could not parse the specified region!\n"); return -4; }
        int tid = hdr.name2tid(std::string(reg, (size_t) nl));
        if (tid < 0) { fprintf(stderr, "%s did not match a known chromosome/contig name!\n", reg); return -6; }
        gTid = (uint32_t) tid;
        if (s > 0) gPos = (uint32_t) s;
        if (e > 0) gEnd = (uint32_t) e;
        if (gEnd > hdr.lens[gTid]) gEnd = hdr.lens[gTid];
    }
    BedFile bed; bool have_bed = false;
    if (bedName) {
        if (!bed.load(bedName, hdr.names, hdr.lens, keepStrand != 0)) { fprintf(stderr, "There was an error while reading in your BED file!\n"); return 1; }
        have_bed = true;
    }
    void *dev = be->create(be->factory_user, &cfg);
    if (!dev) { fprintf(stderr, "Could not initialise the device back end: %s\n", be->last_error ? be->last_error() : "?"); return -20; }
    int rc = 0;
    const size_t batch = tile_reads_default(false);
    const uint32_t chunk32 = (uint32_t) std::min<unsigned long>(chunkSize, 0x7fffffffUL);
    SoaTile tile; std::vector<char> names; std::vector<uint32_t> name_off; std::vector<md_read_meth> res; std::string line, seq;
    if (have_bai) { bool found; uint64_t o = bai.start_offset((int) gTid, gPos, found); if (found && o) bs->seek(o); }
    BamRec r; bool more = bs->peek(r);
    // contigs in file order from the region's contig on (all of them without -r; with -r only that one, perRead.c:128-131)
    for (uint32_t tid = gTid; tid < hdr.names.size() && rc == 0; ++tid) {
        const uint32_t beg = tid == gTid ? gPos : 0, end = (reg && gEnd) ? gEnd : hdr.lens[tid];
        while (more && r.tid >= 0 && (uint32_t) r.tid < tid) { bs->pop(); more = bs->peek(r); }
        bool loaded = false;
        while (more && r.tid == (int32_t) tid && rc == 0) {
            tile.clear(); names.clear(); name_off.clear();
            while (more && r.tid == (int32_t) tid && tile.n() < batch) {
                if (r.pos >= (int32_t) beg && r.pos < (int64_t) end) {
                    tile.add(r);
                    name_off.push_back((uint32_t) names.size());
                    const size_t nl = strnlen(r.qname, r.l_qname);
                    names.insert(names.end(), r.qname, r.qname + nl); names.push_back('\0');
                } else if (r.pos >= (int64_t) end) { while (more && r.tid == (int32_t) tid) { bs->pop(); more = bs->peek(r); } break; }
                bs->pop(); more = bs->peek(r);
            }
            if (!tile.n()) continue;
            if (!loaded) {
                if (!fa->fetch(hdr.names[tid], seq)) { fprintf(stderr, "Couldn't fetch the sequence of %s!\n", hdr.names[tid].c_str()); rc = -2; break; }
                if (be->load_contig(dev, (int32_t) tid, seq.data(), (uint32_t) seq.size()) != 0) { fprintf(stderr, "device error: %s\n", be->last_error ? be->last_error() : "?"); rc = -20; break; }
                loaded = true;
            }
            md_reads_soa v = tile.view(); md_tile_desc td{(int32_t) tid, beg, end, 0, 0};
            res.resize(tile.n());
            if (be->per_read_tile(dev, &td, &v, chunk32, res.data()) != 0) { fprintf(stderr, "device error: %s\n", be->last_error ? be->last_error() : "?"); rc = -20; break; }
            g_stats.n_records += tile.n(); g_stats.n_tiles++;
            line.clear();
            char buf[64];
            for (size_t i = 0; i < tile.n(); ++i) {
                if (res[i].nmeth == 0xffffffffu) continue;
                if (have_bed) {                                            // a chunk no BED region overlaps is skipped as a whole (perRead.c:159-173)
                    const uint64_t lp = (uint64_t) beg + ((uint64_t)((uint32_t) tile.pos[i] - beg) / chunk32) * chunk32, le = std::min<uint64_t>(lp + chunk32, end);
                    if (!bed.chunk_overlaps(tid, (uint32_t) lp, (uint32_t) le)) continue;
                }
                const uint32_t nm = res[i].nmeth, nu = res[i].nunmeth;
                line += names.data() + name_off[i]; line += '\t'; line += hdr.names[tid]; line += '\t';
                snprintf(buf, sizeof buf, "%" PRId64 "\t", (int64_t) tile.pos[i]); line += buf;
                if (nm + nu > 0) { snprintf(buf, sizeof buf, "%f\t%" PRIu32 "\n", 100. * ((double) nm) / (nm + nu), nm + nu); line += buf; }   // perRead.c:19-25
                else { snprintf(buf, sizeof buf, "0.0\t%" PRIu32 "\n", nm + nu); line += buf; }                                              // perRead.c:27-31
            }
            if (!line.empty()) fputs(line.c_str(), ofile);
        }
        if (loaded) be->drop_contig(dev, (int32_t) tid);
        if (reg) break;
    }
    be->destroy(dev);
    if (ofile != stdout) fclose(ofile);
    return rc;
}

extern "C" int mdh_mbias_main(int argc, char *argv[], const mdh_backend *be) {
    md_config cfg; memset(&cfg, 0, sizeof cfg);
    cfg.keepCpG = 1; cfg.minMapq = 10; cfg.minPhred = 5; cfg.ignoreFlags = 0xF00; cfg.noOverlapMerge = 1;   // MBias.c:312-328, :160
    unsigned long chunkSize = 1000000;
    const char *reg = nullptr, *bedName = nullptr; char *opref = nullptr;
    int c, SVG = 1, txt = 0, nThreads = 1, keepStrand = 0; double minConvEff = 0.0; bool threads_given = false;
    double t_start = now_s(); g_t0 = t_start; g_marks = getenv("MD_TIMING") != nullptr;
    tune_allocator();
    memset(&g_stats, 0, sizeof g_stats);
    static struct option lopts[] = {
        {"noCpG", 0, NULL, 1}, {"CHG", 0, NULL, 2}, {"CHH", 0, NULL, 3}, {"keepDupes", 0, NULL, 4}, {"keepSingleton", 0, NULL, 5}, {"keepDiscordant", 0, NULL, 6},
        {"txt", 0, NULL, 7}, {"noSVG", 0, NULL, 8}, {"nOT", 1, NULL, 9}, {"nOB", 1, NULL, 10}, {"nCTOT", 1, NULL, 11}, {"nCTOB", 1, NULL, 12},
        {"chunkSize", 1, NULL, 13}, {"keepStrand", 0, NULL, 14}, {"minConversionEfficiency", 1, NULL, 15}, {"ignoreNH", 0, NULL, 16},
        {"ignoreFlags", 1, NULL, 'F'}, {"requireFlags", 1, NULL, 'R'}, {"help", 0, NULL, 'h'}, {"version", 0, NULL, 'v'},
        {"shardRank", 1, NULL, 200}, {"shardWorld", 1, NULL, 201}, {"histOut", 1, NULL, 202}, {0, 0, NULL, 0}};   // one process per GPU, see api.mbias_sharded
    int shardRank = 0, shardWorld = 1; const char *histOut = nullptr;
    optind = 0;
    while ((c = getopt_long(argc, argv, "hvq:p:r:l:D:F:@:", lopts, NULL)) >= 0) {
        switch (c) {
        case 'h': mbias_usage(); return 0;
        case 'v': printf("%s (B200 build; no HTSlib)\n", MD_VERSION); return 0;
        case 'D': break;
        case 'r': reg = optarg; break;
        case 'l': bedName = optarg; break;
        case 1: cfg.keepCpG = 0; break;
        case 2: cfg.keepCHG = 1; break;
        case 3: cfg.keepCHH = 1; break;
        case 4: cfg.keepDupes = 1; break;
        case 5: cfg.keepSingleton = 1; break;
        case 6: cfg.keepDiscordant = 1; break;
        case 7: txt = 1; break;
        case 8: SVG = 0; txt = 1; break;
        case 9: case 10: case 11: case 12: parse_bounds(optarg, cfg.absoluteBounds, c - 9); break;
        case 13: chunkSize = strtoul(optarg, NULL, 10); if (chunkSize < 1) { fprintf(stderr, "Error: The chunk size must be at least 1!\n"); return 1; } break;
        case 14: keepStrand = 1; break;
        case 15: minConvEff = atof(optarg); break;
        case 16: cfg.ignoreNH = 1; break;
        case 200: shardRank = atoi(optarg); break;
        case 201: shardWorld = atoi(optarg); if (shardWorld < 1) shardWorld = 1; break;
        case 202: histOut = optarg; break;
        case 'F': cfg.ignoreFlags = atoi(optarg); break;
        case 'R': cfg.requireFlags = atoi(optarg); break;
        case 'q': cfg.minMapq = atoi(optarg); break;
        case 'p': cfg.minPhred = atoi(optarg); break;
        case '@': nThreads = atoi(optarg); threads_given = true; break;
        default: fprintf(stderr, "Invalid option '%c'\n", c); mbias_usage(); return 1;
        }
    }
    if (argc == 1) { mbias_usage(); return 0; }
    if (histOut) SVG = 0;
    if ((SVG && argc - optind != 3) || (!SVG && argc - optind < 2)) {
        fprintf(stderr, "You must supply a reference genome in fasta format, an input BAM file, and an output prefix!!!\n"); mbias_usage(); return -1; }
    if (cfg.minPhred < 1) { fprintf(stderr, "-p %i is invalid. resetting to 1, which is the lowest possible value.\n", cfg.minPhred); cfg.minPhred = 1; }
    if (cfg.minMapq < 0) { fprintf(stderr, "-q %i is invalid. Resetting to 0, which is the lowest possible value.\n", cfg.minMapq); cfg.minMapq = 0; }
    if (!(cfg.keepCpG + cfg.keepCHG + cfg.keepCHH)) {
        fprintf(stderr, "You haven't specified any metrics to output!\nEither don't use the --noCpG option or specify --CHG and/or --CHH.\n"); return -1; }
    cfg.minConversionEfficiency = (float) minConvEff;               // Config.minConversionEfficiency is a float (MethylDackel.h:110)
    // NB: mbias never applies the 0x400 adjustment of extract.c:1005-1007 (MBias.c has no such line)

    Driver d; d.be = be;
    const char *fastaName = argv[optind], *bamName = argv[optind + 1];
    const bool dev_decode = device_decode_enabled(be) && !(minConvEff > 0.0);
    try {
        if (dev_decode) { BgzfReader rd(bamName); d.own_hdr = read_bam_header(rd); d.start_voff = rd.tell(); d.hdr = &d.own_hdr; }
        else { const int nt = decode_threads(nThreads, threads_given); d.bam.reset(new ParallelBam(bamName, nt, warm_threads(nt))); d.hdr = &d.bam->header(); }
    } catch (std::exception &e) { fprintf(stderr, "Couldn't open %s for reading!\n", bamName); return -4; }
    d.have_bai = load_bai(bamName, d.bai);
    try { d.fa.reset(new Fasta(fastaName)); }
    catch (std::exception &e) { fprintf(stderr, "Couldn't open the index for %s!\n", fastaName); return -4; }
    if (SVG) opref = argv[optind + 2];
    uint32_t gTid = 0, gPos = 0, gEnd = 0;
    if (reg) {
        int s, e, nl = parse_region(reg, &s, &e);
        if (nl < 0) { fprintf(stderr, "Could not parse the specified region!\n"); return -4; }
        int tid = d.hdr->name2tid(std::string(reg, (size_t) nl));
        if (tid < 0) { fprintf(stderr, "%s did not match a known chromosome/contig name!\n", reg); return -6; }
        gTid = (uint32_t) tid;
        if (s > 0) gPos = (uint32_t) s;
        if (e > 0) gEnd = (uint32_t) e;
        if (gEnd > d.hdr->lens[gTid]) gEnd = d.hdr->lens[gTid];
    }
    if (bedName) {                                                  // MBias.c:522-530
        if (!d.bed.load(bedName, d.hdr->names, d.hdr->lens, keepStrand != 0)) { fprintf(stderr, "There was an error while reading in your BED file!\n"); return 1; }
        d.have_bed = true;
    }
    // as in extract: the device context comes up while the host opens files and starts decoding
    std::future<void *> dev_future = std::async(std::launch::async, [be, &cfg] { void *p = be->create(be->factory_user, &cfg); mark("device context ready"); return p; });
    struct DevJoin { std::future<void *> &f; const mdh_backend *be; bool taken = false; ~DevJoin() { if (!taken && f.valid()) { void *p = f.get(); if (p) be->destroy(p); } } } dev_join{dev_future, be};
    int rc = 0;
    {
        std::vector<Chunk> all;
        {
            ChunkCursor cursor(d.hdr->lens, chunkSize, gTid, gPos, gEnd);
            auto fetch = [&](uint32_t t, int64_t start, int64_t cnt, std::string &w) { return d.window(t, start, cnt, w); };
            Chunk ch;
            while (cursor.next(ch, fetch)) all.push_back(ch);
        }
        size_t c0 = 0, c1 = all.size();
        if (shardWorld > 1) shard_chunks(all, d.have_bai ? &d.bai : nullptr, shardRank, shardWorld, c0, c1);
        if (dev_decode) {
            d.dev = dev_future.get(); dev_join.taken = true;
            if (!d.dev) { fprintf(stderr, "Could not initialise the device back end: %s\n", be->last_error ? be->last_error() : "?"); return -20; }
            rc = mbias_device_decode(d, be, bamName, all, c0, c1);
        } else {
        // tile ring as in extract: page-locked tiles, up to three in flight when the back end is asynchronous
        const bool use_async = be->submit_mbias_tile && be->collect_tile;
        TileAlloc pin; if (use_async && be->pinned_alloc && be->pinned_free) { pin.alloc = be->pinned_alloc; pin.release = be->pinned_free; }
        std::vector<std::unique_ptr<SoaTile>> ring;
        for (int k = 0; k < (use_async ? 3 : 1); ++k) ring.emplace_back(new SoaTile(use_async ? &pin : nullptr));
        const size_t tile_reads = tile_reads_default(use_async);
        if (use_async) for (auto &t : ring) t->reserve_for(tile_reads + tile_reads / 8, 160);
        SoaTile carry;
        mark("tile ring reserved");
        d.dev = dev_future.get(); dev_join.taken = true;
        if (d.bam) d.bam->set_active_threads(d.bam->threads());
        if (!d.dev) { fprintf(stderr, "Could not initialise the device back end: %s\n", be->last_error ? be->last_error() : "?"); return -20; }
        std::vector<int> flight;
        auto collect_one = [&]() -> int {
            int ticket = flight.front(); flight.erase(flight.begin());
            md_tile_stats st;
            double t0 = now_s();
            int r = be->collect_tile(d.dev, ticket, nullptr, 0, &st);
            g_stats.t_device_s += now_s() - t0;
            if (r != 0) { fprintf(stderr, "device error: %s\n", be->last_error ? be->last_error() : "?"); return -20; }
            return 0;
        };
        size_t rk = 0;
        size_t ci = c0;
        while (ci < c1 && rc == 0) {
            uint32_t tid = all[ci].tid;
            std::vector<uint32_t> bounds;
            while (ci < c1 && all[ci].tid == tid) { if (bounds.empty()) bounds.push_back(all[ci].beg); bounds.push_back(all[ci].end); ++ci; }
            const std::string *ref = d.fetch(tid);
            if (ci < c1) d.prefetch(all[ci].tid);                  // the next contig is read from the FASTA while this one is processed
            if (!ref) {
                fprintf(stderr, "faidx_fetch_seq returned %i while trying to fetch the sequence for tid %s:%" PRIu32 "-%" PRIu32 "!\n", -2, d.hdr->names[tid].c_str(), bounds.front(), bounds.back());
                fprintf(stderr, "Note that the output will be truncated!\n");
                break;                                             // MBias.c:152 returns from the worker
            }
            uint32_t rbeg = bounds.front(), rend = bounds.back();
            if (rend > ref->size()) rend = (uint32_t) ref->size();
            if (be->load_contig(d.dev, (int32_t) tid, ref->data(), (uint32_t) ref->size()) != 0 ||
                be->set_mbias_chunks(d.dev, (int32_t) tid, bounds.data(), (uint32_t) bounds.size() - 1) != 0) { fprintf(stderr, "device error: %s\n", be->last_error ? be->last_error() : "?"); rc = -20; break; }
            if ((rc = d.push_bed(tid)) != 0) break;
            d.seek_to((int) tid, rbeg);
            // With --minConversionEfficiency the verdict on an alignment depends on the chunk it is looked at in: the filter only sees
            // the chunk's own window contig[localPos, localEnd] (MBias.c:147,154-156; common.c:363,378), so tiles are cut at chunk ends
            // and carry that window; otherwise the whole run of chunks is one region.
            const bool per_chunk = cfg.minConversionEfficiency > 0.0f;
            std::vector<std::pair<uint32_t, uint32_t>> regions;
            if (per_chunk) for (size_t k = 0; k + 1 < bounds.size(); ++k) regions.emplace_back(bounds[k], std::min<uint32_t>(bounds[k + 1], rend));
            else regions.emplace_back(rbeg, rend);
            for (size_t ri = 0; ri < regions.size() && rc == 0; ++ri) {
            const uint32_t gbeg = regions[ri].first, gend = regions[ri].second;
            const uint32_t ce_beg = per_chunk ? gbeg : 0, ce_end = per_chunk ? (uint32_t) std::min<uint64_t>((uint64_t) bounds[ri + 1] + 1, ref->size()) : 0;
            FragTiler tiler(*d.bam, d.frag, d.frag_i, (int) tid, gbeg, gend, tile_reads); tiler.set_pack_quals(pack_quals_enabled());
            carry.clear();
            for (;;) {
                SoaTile &tile = *ring[rk % ring.size()];
                if (use_async) while (flight.size() >= ring.size() && rc == 0) rc = collect_one();
                if (rc) break;
                double t0 = now_s();
                bool got = tiler.next(tile, carry);
                g_stats.t_decode_s += now_s() - t0;
                if (!got) break;
                if (tile.n() == 0) continue;
                md_reads_soa v = tile.view(); md_tile_desc td{(int32_t) tid, tile.beg, tile.end, ce_beg, ce_end}; md_tile_stats st;
                g_stats.n_records += tile.n(); g_stats.n_tiles++;
                t0 = now_s();
                if (use_async) {
                    int ticket = be->submit_mbias_tile(d.dev, &td, &v);
                    g_stats.t_device_s += now_s() - t0;
                    if (ticket < 0) { fprintf(stderr, "device error: %s\n", be->last_error ? be->last_error() : "?"); rc = -20; break; }
                    flight.push_back(ticket); ++rk;
                } else {
                    int r = be->mbias_tile(d.dev, &td, &v, &st);
                    g_stats.t_device_s += now_s() - t0;
                    if (r != 0) { fprintf(stderr, "device error: %s\n", be->last_error ? be->last_error() : "?"); rc = -20; break; }
                }
            }
            }   // regions
            while (rc == 0 && !flight.empty()) rc = collect_one();       // the contig (and its chunk table) is dropped next
            be->drop_contig(d.dev, (int32_t) tid);
        }
        }   // host decode
    }
    std::vector<uint32_t> hist((size_t) 4 * 2 * MD_MBIAS_MAXLEN * 2, 0); int32_t lens[4] = {0, 0, 0, 0};
    if (rc == 0 && be->mbias_hist(d.dev, hist.data(), lens) != 0) rc = -20;
    if (g_marks && dev_decode) fprintf(stderr, "[md-timing] device-decode driver (calling thread): waiting for a segment's decode %.3f, handing over the next segment %.3f, run table %.3f, contig load %.3f, waiting for the next file segment %.3f\n", g_acc[7], g_acc[8], g_acc[9], g_acc[1], g_acc[5]);
    be->destroy(d.dev);
    if (rc == 0 && histOut) {                                  // shard: hand the raw histogram to the caller, who sums the shards
        FILE *f = fopen(histOut, "wb");
        if (!f || fwrite(lens, sizeof(int32_t), 4, f) != 4 || fwrite(hist.data(), sizeof(uint32_t), hist.size(), f) != hist.size()) rc = -3;
        if (f) fclose(f);
    } else if (rc == 0) {
        // makeSVGs (svg.c:302-437) draws <prefix>_<strand>.svg and prints the suggestion line; then makeTXT (MBias.c:558-559)
        if (SVG) mbias_write_svgs(opref, hist.data(), lens, cfg.keepCpG + 2 * cfg.keepCHG + 4 * cfg.keepCHH, stderr);
        if (txt) mbias_print_txt(stdout, hist.data(), lens);
    }
    g_stats.t_total_s = now_s() - t_start;
    return rc;
}

// ------------------------------------------------------------------ BAM/FASTA helpers for tests + bench
struct mdh_bam { std::string path; BamHeader hdr; SoaTile tile; std::vector<std::unique_ptr<SoaTile>> tiles; };
extern "C" mdh_bam *mdh_bam_open(const char *path) {
    try { BamStream s(path); mdh_bam *b = new mdh_bam(); b->path = path; b->hdr = s.header(); return b; }
    catch (std::exception &e) { g_err = e.what(); return nullptr; }
}
extern "C" void mdh_bam_close(mdh_bam *b) { delete b; }
extern "C" int mdh_bam_n_targets(const mdh_bam *b) { return (int) b->hdr.names.size(); }
extern "C" const char *mdh_bam_target_name(const mdh_bam *b, int tid) { return b->hdr.names[(size_t) tid].c_str(); }
extern "C" uint32_t mdh_bam_target_len(const mdh_bam *b, int tid) { return b->hdr.lens[(size_t) tid]; }
extern "C" int mdh_bam_read_region(mdh_bam *b, int tid, uint32_t beg, uint32_t end, md_reads_soa *out) {
    try {
        BamStream s(b->path);
        Tiler t(s, tid, beg, end, (size_t) -1);
        SoaTile carry;
        b->tile.clear();
        t.next(b->tile, carry);
        if (pack_quals_enabled()) { PodVec<uint64_t> sq; PodVec<uint32_t> so; b->tile.pack_quals(sq, so); }
        *out = b->tile.view();
        return 0;
    } catch (std::exception &e) { g_err = e.what(); return -1; }
}
// Cut [beg,end) of contig tid into tiles of ~target_reads alignments exactly as the sub-command driver does
// (Tiler: straddling alignments are present in both neighbours). Returns the number of tiles, kept alive by the handle.
extern "C" int mdh_bam_make_tiles(mdh_bam *b, int tid, uint32_t beg, uint32_t end, uint64_t target_reads) {
    try {
        BamStream s(b->path);
        Tiler t(s, tid, beg, end, (size_t) target_reads);
        SoaTile carry;
        b->tiles.clear();
        for (;;) {
            std::unique_ptr<SoaTile> tile(new SoaTile());
            if (!t.next(*tile, carry)) break;
            if (tile->n() == 0) continue;
            if (pack_quals_enabled()) { PodVec<uint64_t> sq; PodVec<uint32_t> so; tile->pack_quals(sq, so); }
            b->tiles.push_back(std::move(tile));
        }
        return (int) b->tiles.size();
    } catch (std::exception &e) { g_err = e.what(); return -1; }
}
extern "C" int mdh_bam_get_tile(mdh_bam *b, int k, md_tile_desc *td, md_reads_soa *out) {
    if (k < 0 || (size_t) k >= b->tiles.size()) return -1;
    SoaTile &t = *b->tiles[(size_t) k];
    td->tid = t.tid; td->beg = t.beg; td->end = t.end; td->ce_beg = 0; td->ce_end = 0;
    *out = t.view();
    return 0;
}
struct mdh_fasta { std::unique_ptr<Fasta> fa; std::string seq; };
extern "C" mdh_fasta *mdh_fasta_open(const char *path) {
    try { mdh_fasta *f = new mdh_fasta(); f->fa.reset(new Fasta(path)); return f; }
    catch (std::exception &e) { g_err = e.what(); return nullptr; }
}
extern "C" void mdh_fasta_close(mdh_fasta *f) { delete f; }
extern "C" const char *mdh_fasta_fetch(mdh_fasta *f, const char *name, uint32_t *len) {
    if (!f->fa->fetch(name, f->seq)) { g_err = std::string("sequence not found: ") + name; return nullptr; }
    *len = (uint32_t) f->seq.size();
    return f->seq.data();
}

// Report stage of mbias on an (already summed) histogram: what mbias_main does after joining its threads
// (MBias.c:557-559): the suggestion line of makeSVGs (svg.c:423-425) and/or makeTXT (svg.c:439-454).
extern "C" void mdh_mbias_report(const uint32_t *hist, const int32_t lens[4], int svg, int txt) {
    if (svg) mbias_print_suggestions(stderr, hist, lens);
    if (txt) mbias_print_txt(stdout, hist, lens);
    fflush(stdout); fflush(stderr);
}
// The same with the plots: opref = the SVG prefix (NULL: no plots, no suggestion line), which = keepCpG + 2 keepCHG + 4 keepCHH
extern "C" int mdh_mbias_report_svg(const uint32_t *hist, const int32_t lens[4], const char *opref, int which, int txt) {
    bool ok = true;
    if (opref) ok = mbias_write_svgs(opref, hist, lens, which, stderr);
    if (txt) mbias_print_txt(stdout, hist, lens);
    fflush(stdout); fflush(stderr);
    return ok ? 0 : -3;
}
