// Multi-threaded BAM -> SoA decode for the host driver.
//
// The reference decodes inside each worker thread's sam_itr_next (common.c:413): BGZF inflate + record
// parse are part of its per-chunk pthread work.  Here decode is a pipeline feeding the GPU:
//
//   reader (caller's thread)  cuts the compressed file into jobs of whole BGZF blocks (header scan only)
//   pool:   inflate           every job's blocks are inflated independently into one contiguous buffer
//   caller: stitch            walks the 4-byte record length prefixes across job boundaries (cheap, ordered)
//   pool:   parse             each job's records -> a SoA fragment (SoaTile::add: aux scan, name hash, copies)
//   caller: FragTiler         concatenates fragments into device tiles with bulk array copies
//
// Only the stitch and the tile assembly are serial, and both are memcpy-class work.
#pragma once
#include "tiles.hpp"
#include <thread>
#include <mutex>
#include <condition_variable>
#include <deque>
#include <future>
#include <atomic>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>

namespace mdhost {

class ThreadPool {
public:
    explicit ThreadPool(int n) {
        if (n < 1) n = 1;
        for (int i = 0; i < n; ++i) th_.emplace_back([this, i] { run(i); });
    }
    ~ThreadPool() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    // fire-and-forget; `urgent` tasks overtake the queued decode work (tile-level helpers the caller waits on)
    void post(std::function<void()> fn, bool urgent = false) {
        bool all; { std::lock_guard<std::mutex> g(m_); if (urgent) q_.emplace_front(std::move(fn)); else q_.emplace_back(std::move(fn)); all = active_ < (int) th_.size(); }
        if (all) cv_.notify_all(); else cv_.notify_one();
    }
    int size() const { return (int) th_.size(); }
    // only the first n workers take tasks (the rest sleep) until the limit is raised again
    void set_active(int n) { { std::lock_guard<std::mutex> g(m_); active_ = n; } cv_.notify_all(); }
private:
    void run(int me) {
        for (;;) {
            std::function<void()> fn;
            { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return stop_ || (!q_.empty() && me < active_); }); if (stop_ && (q_.empty() || me >= active_)) return; fn = std::move(q_.front()); q_.pop_front(); }
            fn();
        }
    }
    std::vector<std::thread> th_; std::deque<std::function<void()>> q_; std::mutex m_; std::condition_variable cv_; bool stop_ = false; int active_ = 1 << 30;
};

// A decoded run of consecutive records: SoA columns + the contig id of each record.
struct Fragment {
    SoaTile soa;
    std::vector<int32_t> tid;
};

// The compressed file is mapped, not read: jobs point into the mapping, so the only bytes the caller's thread touches are
// the 18-byte block headers.  Inflate -> stitch -> parse advance by continuation on the pool (the worker that completes the
// inflate of the oldest unstitched job walks the record-length chain for it and every inflated successor, then queues
// their parses), so decoding runs ahead of the consumer — in particular while the CUDA context is still coming up.
class ParallelBam {
public:
    // `warm_threads` >= 0 keeps all but that many decode workers asleep until set_active_threads() raises the limit
    ParallelBam(const std::string &path, int nthreads, int warm_threads = -1) : pool_(nthreads), aux_(std::max(1, std::min(16, nthreads))) {
        if (warm_threads >= 0) pool_.set_active(warm_threads);
        // header through the sequential reader; remember where the records start
        BgzfReader rd(path);
        hdr_ = read_bam_header(rd);
        start_voff_ = rd.tell();
        fd_ = ::open(path.c_str(), O_RDONLY);
        if (fd_ < 0) throw std::runtime_error("Couldn't open " + path + " for reading!");
        struct stat st;
        if (fstat(fd_, &st) != 0 || st.st_size <= 0) { ::close(fd_); throw std::runtime_error("Couldn't stat " + path); }
        size_ = (size_t) st.st_size;
        void *m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (m == MAP_FAILED) { ::close(fd_); throw std::runtime_error("Couldn't map " + path); }
        map_ = (const uint8_t *) m;
        madvise((void *) map_, size_, MADV_SEQUENTIAL);
        seek(start_voff_);
    }
    ~ParallelBam() {
        quiesce();
        if (map_) munmap((void *) map_, size_);
        if (fd_ >= 0) ::close(fd_);
    }
    const BamHeader &header() const { return hdr_; }
    uint64_t first_record_voffset() const { return start_voff_; }
    int threads() const { return pool_.size(); }
    size_t jobs_adopted() const { return n_adopted_; }
    size_t jobs_walked() const { return n_walked_; }
    void set_active_threads(int n) { pool_.set_active(n); }
    // run fn(0..n-1) and wait (tile-level work the caller is blocked on: column copies, phred packing).  These go to a small
    // pool of their own: the decode pool's workers sit in ~10 ms inflate/parse jobs, which would be the latency of every call.
    void parallel_for(size_t n, const std::function<void(size_t)> &fn) {
        if (n <= 1) { if (n) fn(0); return; }
        std::mutex m; std::condition_variable cv; size_t left = n - 1;
        for (size_t k = 1; k < n; ++k) aux_.post([&, k] { fn(k); std::lock_guard<std::mutex> g(m); if (--left == 0) cv.notify_one(); });
        fn(0);
        std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return left == 0; });
    }

    // Restart decoding at a BGZF virtual offset that points at a record start (from a BAI, or first_record_voffset()).
    void seek(uint64_t voff) {
        quiesce();
        std::lock_guard<std::mutex> g(m_);
        jobs_.clear(); pending_.clear(); stitch_next_ = 0; cancel_ = false;
        file_off_ = (size_t)(voff >> 16); skip_ = (size_t)(voff & 0xffff); eof_ = false;
        if (file_off_ > size_) throw std::runtime_error("seek failed");
        top_up_locked();
    }

    // Next fragment in file order; nullptr at end of file.
    std::unique_ptr<Fragment> next() {
        std::unique_lock<std::mutex> l(m_);
        for (;;) {
            top_up_locked();
            if (jobs_.empty()) return nullptr;
            Job &head = *jobs_.front();
            cv_.wait(l, [&] { return head.state == PARSED || head.state == FAILED; });
            if (head.state == FAILED) throw std::runtime_error(head.err);
            std::unique_ptr<Fragment> out = std::move(head.frag);
            jobs_.pop_front(); if (stitch_next_) --stitch_next_;
            if (out->soa.n() == 0 && !(jobs_.empty() && eof_)) continue;   // a job that only completed a straddling record elsewhere
            top_up_locked();
            return out;
        }
    }

private:
    enum State { QUEUED = 0, INFLATED, STITCHED, PARSED, FAILED };
    struct Job {
        const uint8_t *base = nullptr; std::vector<std::pair<uint32_t, uint32_t>> blocks;   // (offset from base, total block size)
        std::vector<uint8_t> ubuf; std::vector<uint8_t> head_rec;                         // head_rec: a record completed from the previous job's tail
        std::vector<uint32_t> rec_off;                                                     // starts of whole records in ubuf (pointing at the length prefix)
        State state = QUEUED; std::string err;
        std::unique_ptr<Fragment> frag;
        // Speculative record chain, built by the inflating worker while the bytes are hot in its cache: the first offset
        // from which a few consecutive plausible record headers follow, and the chain walked from there.  The ordered
        // stitch step adopts it when the true first record start (known from the predecessor) is that offset — then the
        // serial work per job is O(1) — and otherwise walks the chain itself.
        size_t guess_start = (size_t) -1, guess_tail = 0; std::vector<uint32_t> guess_off;
        int32_t n_targets = 0;
        bool plausible(size_t o, int depth) const {
            const size_t U = ubuf.size(); const uint8_t *u = ubuf.data();
            for (int d = 0; d < depth; ++d) {
                if (o + 36 > U) return d > 0;                         // ran off the end after at least one full header
                const uint32_t bs = le32(u + o);
                if (bs < 32 || bs > (1u << 28)) return false;
                const int32_t tid = (int32_t) le32(u + o + 4), pos = (int32_t) le32(u + o + 8);
                const uint32_t lq = u[o + 12], ncig = le32(u + o + 16) & 0xffffu; const int32_t lseq = (int32_t) le32(u + o + 20);
                const int32_t mtid = (int32_t) le32(u + o + 24);
                if (tid < -1 || tid >= n_targets || mtid < -1 || mtid >= n_targets || pos < -1 || lq < 1 || lseq < 0) return false;
                const size_t need = 32 + (size_t) lq + 4 * (size_t) ncig + ((size_t) lseq + 1) / 2 + (size_t) lseq;
                if (need > bs) return false;
                if (o + 36 + lq <= U && u[o + 36 + lq - 1] != 0) return false;   // query name is NUL-terminated
                o += 4 + (size_t) bs;
                if (o == U) return true;
            }
            return true;
        }
        void speculate() {
            const size_t U = ubuf.size(); const uint8_t *u = ubuf.data();
            const size_t limit = std::min<size_t>(U, (size_t) 1 << 16);   // a first record start further in than this: leave it to the stitcher
            for (size_t o = 0; o + 36 <= limit; ++o) if (plausible(o, 4)) { guess_start = o; break; }
            if (guess_start == (size_t) -1) return;
            size_t off = guess_start;
            guess_off.reserve((U - off) / (4 + (size_t) le32(u + off)) + 16);
            while (off + 4 <= U) { const size_t len = 4 + (size_t) le32(u + off); if (off + len > U) break; guess_off.push_back((uint32_t) off); off += len; }
            guess_tail = off;
        }
        void inflate_all() {
            // sizes first (ISIZE trailer), then inflate block by block into place
            size_t tot = 0; std::vector<size_t> uoff(blocks.size());
            for (size_t b = 0; b < blocks.size(); ++b) { const uint8_t *p = base + blocks[b].first; uint32_t bs = blocks[b].second; uoff[b] = tot; tot += le32(p + bs - 4); }
            ubuf.resize(tot);
            z_stream zs; memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, -15) != Z_OK) throw std::runtime_error("inflateInit2");
            for (size_t b = 0; b < blocks.size(); ++b) {
                const uint8_t *p = base + blocks[b].first; uint32_t bs = blocks[b].second;
                int xlen = p[10] | (p[11] << 8);
                uint32_t isize = le32(p + bs - 4);
                if (isize == 0) continue;
                inflateReset(&zs);
                zs.next_in = (Bytef *)(p + 12 + xlen); zs.avail_in = bs - 12 - (uint32_t) xlen - 8;
                zs.next_out = ubuf.data() + uoff[b]; zs.avail_out = isize;
                if (inflate(&zs, Z_FINISH) != Z_STREAM_END || zs.total_out != isize) { inflateEnd(&zs); throw std::runtime_error("BGZF inflate failed"); }
            }
            inflateEnd(&zs);
            speculate();
        }
        void parse() {
            frag.reset(new Fragment());
            BamRec r;
            const size_t nrec = rec_off.size() + (head_rec.empty() ? 0 : 1);
            frag->tid.reserve(nrec);
            auto add = [&](const uint8_t *p) {
                uint32_t bs = le32(p);
                if (!parse_bam_record(p + 4, bs, r)) throw std::runtime_error("malformed BAM record");
                frag->soa.add(r); frag->tid.push_back(r.tid);
            };
            if (!head_rec.empty()) add(head_rec.data());
            for (uint32_t o : rec_off) add(ubuf.data() + o);
            std::vector<uint8_t>().swap(ubuf);
        }
    };

    // wait until no pool task of this reader is running or queued
    void quiesce() {
        std::unique_lock<std::mutex> l(m_);
        cancel_ = true;
        cv_.wait(l, [&] { return tasks_ == 0; });
    }
    // cut jobs of whole BGZF blocks from the mapping (header scan only) and queue their inflates; m_ held
    void top_up_locked() {
        const size_t want_jobs = (size_t) pool_.size() * 3 + 2;
        static const size_t target = [] { const char *e = getenv("MD_DECODE_JOB_BYTES"); size_t v = e ? (size_t) atol(e) : 0; return v ? v : (size_t) 1 << 20; }();   // ~1 MB of compressed blocks per job
        while (!eof_ && jobs_.size() < want_jobs) {
            std::shared_ptr<Job> j(new Job());
            j->base = map_ + file_off_; j->n_targets = (int32_t) hdr_.names.size();
            size_t used = 0;
            while (file_off_ + used + 18 <= size_) {
                const uint8_t *p = map_ + file_off_ + used;
                if (p[0] != 31 || p[1] != 139 || !(p[3] & 4)) throw std::runtime_error("not a BGZF block");
                int xlen = p[10] | (p[11] << 8), bsize = -1;
                if (file_off_ + used + 12 + (size_t) xlen > size_) throw std::runtime_error("truncated BGZF header");
                for (int off = 0; off + 4 <= xlen;) { int slen = p[12 + off + 2] | (p[12 + off + 3] << 8); if (p[12 + off] == 'B' && p[12 + off + 1] == 'C' && slen == 2) bsize = p[12 + off + 4] | (p[12 + off + 5] << 8); off += 4 + slen; }
                if (bsize < 0) throw std::runtime_error("BGZF block without BC field");
                const size_t bs = (size_t) bsize + 1;
                if (file_off_ + used + bs > size_) throw std::runtime_error("truncated BGZF block");
                j->blocks.emplace_back((uint32_t) used, (uint32_t) bs);
                used += bs;
                if (used >= target) break;
            }
            if (j->blocks.empty()) { eof_ = true; break; }
            file_off_ += used;
            jobs_.push_back(j);
            ++tasks_;
            pool_.post([this, j] { run_inflate(j); });
        }
    }
    void run_inflate(const std::shared_ptr<Job> &j) {
        bool ok = true; std::string err;
        if (!cancel_) { try { j->inflate_all(); } catch (std::exception &e) { ok = false; err = e.what(); } }
        std::lock_guard<std::mutex> g(m_);
        if (cancel_) { --tasks_; cv_.notify_all(); return; }
        if (!ok) { j->state = FAILED; j->err = err; } else j->state = INFLATED;
        // advance the stitch chain as far as the inflated prefix reaches
        while (stitch_next_ < jobs_.size()) {
            std::shared_ptr<Job> k = jobs_[stitch_next_];
            if (k->state == FAILED) break;
            if (k->state != INFLATED) break;
            try { stitch(*k); } catch (std::exception &e) { k->state = FAILED; k->err = e.what(); break; }
            k->state = STITCHED; ++stitch_next_;
            ++tasks_;
            pool_.post([this, k] { run_parse(k); });
        }
        --tasks_; cv_.notify_all();
    }
    void run_parse(const std::shared_ptr<Job> &j) {
        bool ok = true; std::string err;
        if (!cancel_) { try { j->parse(); } catch (std::exception &e) { ok = false; err = e.what(); } }
        std::lock_guard<std::mutex> g(m_);
        if (!cancel_) { if (ok) j->state = PARSED; else { j->state = FAILED; j->err = err; } }
        --tasks_; cv_.notify_all();
    }
    // record boundaries of one job; completes the previous job's straddling record; m_ held
    void stitch(Job &j) {
        size_t off = skip_; skip_ = 0;
        const size_t U = j.ubuf.size();
        if (off > U) throw std::runtime_error("BGZF seek past block end");
        if (!pending_.empty()) {
            // pending_ holds the first bytes of a record that began in an earlier job
            size_t have = pending_.size();
            if (have < 4) { size_t take = std::min(4 - have, U - off); pending_.insert(pending_.end(), j.ubuf.begin() + (ptrdiff_t) off, j.ubuf.begin() + (ptrdiff_t)(off + take)); off += take; have = pending_.size(); }
            if (have >= 4) {
                size_t need = 4 + (size_t) le32(pending_.data()) - have;
                size_t take = std::min(need, U - off);
                pending_.insert(pending_.end(), j.ubuf.begin() + (ptrdiff_t) off, j.ubuf.begin() + (ptrdiff_t)(off + take)); off += take;
                if (take == need) { j.head_rec.swap(pending_); pending_.clear(); }
            }
        }
        if (pending_.empty() && off == j.guess_start) {
            j.rec_off.swap(j.guess_off); off = j.guess_tail; ++n_adopted_;
            if (off < U) pending_.assign(j.ubuf.begin() + (ptrdiff_t) off, j.ubuf.end());
        } else if (pending_.empty()) {
            ++n_walked_;
            const uint8_t *u = j.ubuf.data();
            size_t nest = 0;
            if (U > off + 4) { size_t l0 = 4 + (size_t) le32(u + off); nest = l0 ? (U - off) / l0 + 16 : 0; }
            j.rec_off.reserve(std::min<size_t>(nest, (size_t) 1 << 22));
            while (off + 4 <= U) {
                size_t len = 4 + (size_t) le32(u + off);
                if (off + len > U) break;
                j.rec_off.push_back((uint32_t) off); off += len;
            }
            if (off < U) pending_.assign(j.ubuf.begin() + (ptrdiff_t) off, j.ubuf.end());
        }
    }

    ThreadPool pool_, aux_;
    BamHeader hdr_;
    int fd_ = -1; const uint8_t *map_ = nullptr; size_t size_ = 0;
    uint64_t start_voff_ = 0;
    std::mutex m_; std::condition_variable cv_;
    size_t file_off_ = 0, skip_ = 0; bool eof_ = false; std::atomic<bool> cancel_{false};
    size_t tasks_ = 0, stitch_next_ = 0, n_adopted_ = 0, n_walked_ = 0;
    std::vector<uint8_t> pending_;
    std::deque<std::shared_ptr<Job>> jobs_;
};

// Cuts the mapped BAM into segments of whole BGZF blocks for the device decoder (md_bam_push): only the 18-byte block
// headers are read on the host.
class BgzfSegmenter {
public:
    explicit BgzfSegmenter(const std::string &path) {
        fd_ = ::open(path.c_str(), O_RDONLY);
        if (fd_ < 0) throw std::runtime_error("Couldn't open " + path + " for reading!");
        struct stat st;
        if (fstat(fd_, &st) != 0 || st.st_size <= 0) { ::close(fd_); throw std::runtime_error("Couldn't stat " + path); }
        size_ = (size_t) st.st_size;
        void *m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (m == MAP_FAILED) { ::close(fd_); throw std::runtime_error("Couldn't map " + path); }
        map_ = (const uint8_t *) m;
        madvise((void *) map_, size_, MADV_SEQUENTIAL);
    }
    ~BgzfSegmenter() { if (map_) munmap((void *) map_, size_); if (fd_ >= 0) ::close(fd_); }
    void seek(uint64_t file_off) { off_ = (size_t) file_off; }
    bool eof() const { return off_ + 18 > size_; }
    size_t remaining() const { return off_ < size_ ? size_ - off_ : 0; }
    uint64_t tell() const { return off_; }
    // next segment of about `target` compressed bytes: pointer + length, block table with payload offsets relative to it
    bool next(size_t target, const uint8_t *&base, size_t &bytes, std::vector<md_bgzf_block> &blocks) {
        blocks.clear();
        base = map_ + off_; size_t used = 0;
        while (off_ + used + 18 <= size_) {
            const uint8_t *p = map_ + off_ + used;
            if (p[0] != 31 || p[1] != 139 || !(p[3] & 4)) throw std::runtime_error("not a BGZF block");
            int xlen = p[10] | (p[11] << 8), bsize = -1;
            if (off_ + used + 12 + (size_t) xlen > size_) throw std::runtime_error("truncated BGZF header");
            for (int o = 0; o + 4 <= xlen;) { int slen = p[12 + o + 2] | (p[12 + o + 3] << 8); if (p[12 + o] == 'B' && p[12 + o + 1] == 'C' && slen == 2) bsize = p[12 + o + 4] | (p[12 + o + 5] << 8); o += 4 + slen; }
            if (bsize < 0) throw std::runtime_error("BGZF block without BC field");
            const size_t bs = (size_t) bsize + 1;
            if (off_ + used + bs > size_ || bs < (size_t) 12 + xlen + 8) throw std::runtime_error("truncated BGZF block");
            md_bgzf_block b; b.comp_off = used + 12 + (size_t) xlen; b.comp_len = (uint32_t)(bs - 12 - (size_t) xlen - 8); b.isize = le32(p + bs - 4);
            blocks.push_back(b);
            used += bs;
            if (used >= target) break;
        }
        bytes = used; off_ += used;
        return !blocks.empty();
    }
private:
    int fd_ = -1; const uint8_t *map_ = nullptr; size_t size_ = 0, off_ = 0;
};

// Segments of the compressed file read into page-locked buffers ahead of the device: a copy from a file mapping is a
// pageable transfer (staged by the driver at ~11 GB/s and blocking the stream it is queued on), one from page-locked memory
// runs at PCIe rate and is asynchronous.  A background thread reads the next segments with pread — several slices in
// parallel; reading through the mapping instead costs a minor fault per 4 KB page and held the whole pipeline at ~2.7 GB/s —
// cuts them at BGZF block boundaries and hands them over in file order.  A segment's buffer is recycled `depth - 1` calls of
// next() later, which is what the overlapped device push needs (segment k counted, k+1 being decoded, k+2 staged).
class StagedSegments {
public:
    struct Seg { const uint8_t *base = nullptr; size_t bytes = 0; std::vector<md_bgzf_block> blocks; };
    StagedSegments(const std::string &path, uint64_t file_off, size_t target, void *(*alloc)(size_t), void (*release)(void *), int depth = 5, int io_threads = 6)
        : target_(target), alloc_(alloc), release_(release), io_threads_(io_threads), off_(file_off) {
        fd_ = ::open(path.c_str(), O_RDONLY);
        struct stat st;
        if (fd_ < 0 || fstat(fd_, &st) != 0) { if (fd_ >= 0) ::close(fd_); fd_ = -1; return; }
        size_ = (uint64_t) st.st_size;
        cap_ = target + (1u << 17);                       // a segment ends at the first block boundary at or beyond the target
        // the first buffer now (it also tells whether page-locked memory can be had at all); the others are allocated by the reader
        // when it first needs them, so the consumer gets segment 0 after one allocation instead of `depth`
        buf_.assign((size_t) std::max(depth, 4), nullptr);
        buf_[0] = alloc ? (uint8_t *) alloc(cap_) : nullptr;
        if (buf_[0]) th_ = std::thread([this] { run(); });
    }
    ~StagedSegments() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; } cv_.notify_all();
        if (th_.joinable()) th_.join();
        for (uint8_t *p : buf_) if (p && release_) release_(p);
        if (fd_ >= 0) ::close(fd_);
    }
    bool staged() const { return th_.joinable(); }
    // next segment in file order; false at the end of the file.  A result stays valid until the SECOND call after the one that
    // returned it (the driver holds two at a time: one being decoded, one being copied ahead — md_bam_prefetch).
    bool next(Seg &out) {
        std::unique_lock<std::mutex> l(m_);
        ++taken_; cv_.notify_all();
        cv_.wait(l, [&] { return !q_.empty() || done_; });
        if (!err_.empty()) throw std::runtime_error(err_);
        if (q_.empty()) return false;
        out = std::move(q_.front()); q_.pop_front();
        return true;
    }
private:
    void run() {
        try {
            for (size_t k = 0; off_ + 18 <= size_; ++k) {
                { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return stop_ || k + 2 < taken_ + buf_.size(); }); if (stop_) return; }   // segment k reuses the buffer of segment k - depth, which must have been returned at least three calls ago
                uint8_t *&slot = buf_[k % buf_.size()];
                if (!slot) { slot = (uint8_t *) alloc_(cap_); if (!slot) throw std::runtime_error("out of page-locked memory for the staging buffers"); }
                uint8_t *dst = slot;
                const size_t want = (size_t) std::min<uint64_t>(cap_, size_ - off_);
                const int nt = want > (8u << 20) ? io_threads_ : 1;
                std::atomic<bool> ok{true};
                auto slice = [&](int t) {
                    size_t a = want * (size_t) t / nt; const size_t b = want * (size_t)(t + 1) / nt;
                    while (a < b) { ssize_t r = pread(fd_, dst + a, b - a, (off_t)(off_ + a)); if (r <= 0) { ok = false; return; } a += (size_t) r; }
                };
                std::vector<std::thread> th;
                for (int t = 1; t < nt; ++t) th.emplace_back(slice, t);
                slice(0);
                for (auto &t : th) t.join();
                if (!ok) throw std::runtime_error("read error on the BAM file");
                // whole BGZF blocks up to the target
                Seg s; size_t used = 0;
                while (used + 18 <= want) {
                    const uint8_t *p = dst + used;
                    if (p[0] != 31 || p[1] != 139 || !(p[3] & 4)) throw std::runtime_error("not a BGZF block");
                    const int xlen = p[10] | (p[11] << 8); int bsize = -1;
                    if (used + 12 + (size_t) xlen > want) break;
                    for (int o = 0; o + 4 <= xlen;) { int slen = p[12 + o + 2] | (p[12 + o + 3] << 8); if (p[12 + o] == 'B' && p[12 + o + 1] == 'C' && slen == 2) bsize = p[12 + o + 4] | (p[12 + o + 5] << 8); o += 4 + slen; }
                    if (bsize < 0) throw std::runtime_error("BGZF block without BC field");
                    const size_t bs = (size_t) bsize + 1;
                    if (bs < (size_t) 12 + xlen + 8) throw std::runtime_error("truncated BGZF block");
                    if (used + bs > want) { if (off_ + used + bs > size_) throw std::runtime_error("truncated BGZF block"); break; }
                    md_bgzf_block b; b.comp_off = used + 12 + (size_t) xlen; b.comp_len = (uint32_t)(bs - 12 - (size_t) xlen - 8); b.isize = le32(p + bs - 4);
                    s.blocks.push_back(b);
                    used += bs;
                    if (used >= target_) break;
                }
                if (s.blocks.empty()) { if (off_ + 18 <= size_ && want == cap_) throw std::runtime_error("BGZF block larger than the staging buffer"); break; }
                s.base = dst; s.bytes = used; off_ += used;
                { std::lock_guard<std::mutex> g(m_); q_.push_back(std::move(s)); } cv_.notify_all();
            }
        } catch (std::exception &e) { std::lock_guard<std::mutex> g(m_); err_ = e.what(); }
        { std::lock_guard<std::mutex> g(m_); done_ = true; } cv_.notify_all();
    }
    size_t target_, cap_ = 0; void *(*alloc_)(size_t); void (*release_)(void *); int io_threads_; int fd_ = -1; uint64_t off_ = 0, size_ = 0;
    std::vector<uint8_t *> buf_; std::thread th_;
    std::mutex m_; std::condition_variable cv_; std::deque<Seg> q_; size_t taken_ = 0; bool stop_ = false, done_ = false; std::string err_;
};

// Same contract as Tiler (tiles.hpp), fed by fragments.  The caller's thread only scans positions to decide which runs of
// records belong to the tile; the column copies of those runs are done on the decode pool.
class FragTiler {
public:
    FragTiler(ParallelBam &pb, std::shared_ptr<Fragment> &cur, size_t &cur_i, int tid, uint32_t reg_beg, uint32_t reg_end, size_t target_reads)
        : pb_(pb), cur_(cur), i_(cur_i), tid_(tid), reg_beg_(reg_beg), reg_end_(reg_end), target_(target_reads), cur_beg_(reg_beg) {}
    // phred columns of the tiles are written as 2/4-bit codes when the tile's alphabet allows (md_reads_soa::qual_bits)
    void set_pack_quals(bool on) { pack_ = on; }
    bool next(SoaTile &t, SoaTile &carry) {
        if (done_) return false;
        t.clear(); t.tid = tid_; t.beg = cur_beg_;
        std::vector<size_t> keep;
        for (size_t i = 0; i < carry.n(); ++i) if (span_end(carry, i) > cur_beg_) keep.push_back(i);
        uint32_t cut = reg_end_;
        bool stream_end = false;
        struct Run { std::shared_ptr<Fragment> f; size_t a, b; SoaTile::Extent at; };
        std::vector<Run> plan; size_t have = keep.size();
        for (;;) {
            if (!cur_ || i_ >= cur_->soa.n()) { cur_ = pb_.next(); i_ = 0; if (!cur_) { stream_end = true; break; } if (cur_->soa.n() == 0) continue; }
            const SoaTile &f = cur_->soa;
            const int32_t rt = cur_->tid[i_];
            if (rt != tid_) { if (rt > tid_ || rt < 0) { stream_end = true; break; } ++i_; continue; }
            if ((uint32_t) f.pos[i_] >= reg_end_) { stream_end = true; break; }
            if (have >= target_ && (uint32_t) f.pos[i_] > cur_beg_ && f.pos[i_] > last_pos_) { cut = (uint32_t) f.pos[i_]; break; }
            // take a run of records of this contig that stays below the region end (and, once the tile is full, at one position)
            size_t j = i_;
            const size_t room = std::min<size_t>(have >= target_ ? 1 : target_ - have, 8192);
            while (j < f.n() && j - i_ < room && cur_->tid[j] == tid_ && (uint32_t) f.pos[j] < reg_end_) ++j;
            // records entirely left of the region start are dropped (index semantics: endpos > beg)
            size_t a = i_;
            while (a < j) {
                while (a < j && !((int64_t) std::max(f.rend[a], f.pos[a] + 1) > (int64_t) reg_beg_)) ++a;
                size_t b = a;
                while (b < j && (int64_t) std::max(f.rend[b], f.pos[b] + 1) > (int64_t) reg_beg_) ++b;
                if (b > a) { plan.push_back(Run{cur_, a, b, SoaTile::Extent()}); have += b - a; last_pos_ = f.pos[b - 1]; }
                a = b;
            }
            i_ = j;
        }
        // the tile's phred alphabet: what the contributing fragments saw while they were parsed, plus the carried reads
        if (pack_) {
            uint8_t seen[256] = {0};
            const Fragment *lastf = nullptr;
            for (const Run &r : plan) if (r.f.get() != lastf) { lastf = r.f.get(); for (int v = 0; v < 256; ++v) seen[v] |= r.f->soa.qseen[v]; }
            for (size_t i : keep) for (uint32_t j = 0, l = carry.l_qseq[i]; j < l; ++j) seen[carry.qual_value(i, j)] = 1;
            uint8_t lut[16]; int na = 0;
            for (int v = 0; v < 256; ++v) if (seen[v]) { if (na < 16) lut[na] = (uint8_t) v; ++na; }
            if (na <= 16) t.set_encoding(na <= 4 ? 2 : 4, lut, na);
        }
        for (size_t i : keep) t.add_from(carry, i);
        carry.clear();
        if (!plan.empty()) {
            SoaTile::Extent total;
            for (Run &r : plan) { r.at = total; const SoaTile::Extent e = t.extent_of(r.f->soa, r.a, r.b); total.n += e.n; total.c += e.c; total.s += e.s; total.q += e.q; }
            const SoaTile::Extent base = t.extend(total);
            pb_.parallel_for(plan.size(), [&](size_t k) {
                const Run &r = plan[k];
                SoaTile::Extent at; at.n = base.n + r.at.n; at.c = base.c + r.at.c; at.s = base.s + r.at.s; at.q = base.q + r.at.q;
                t.copy_range_at(r.f->soa, r.a, r.b, at);
            });
        }
        t.end = cut;
        // reads reaching beyond the cut (or beyond the region end) are handed on: to the next tile, or to the tiler of the adjacent region
        for (size_t i = 0; i < t.n(); ++i) if (span_end(t, i) > cut) carry.add_from(t, i);
        if (stream_end) done_ = true; else cur_beg_ = cut;
        return true;
    }
private:
    static uint32_t span_end(const SoaTile &t, size_t i) { return (uint32_t) std::max(t.rend[i], t.pos[i] + 1); }
    ParallelBam &pb_; std::shared_ptr<Fragment> &cur_; size_t &i_;
    int tid_; uint32_t reg_beg_, reg_end_; size_t target_; uint32_t cur_beg_; bool done_ = false; int32_t last_pos_ = -1; bool pack_ = false;
};

}  // namespace mdhost
