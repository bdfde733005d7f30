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

namespace mdhost {

class ThreadPool {
public:
    explicit ThreadPool(int n) {
        if (n < 1) n = 1;
        for (int i = 0; i < n; ++i) th_.emplace_back([this] { run(); });
    }
    ~ThreadPool() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    std::future<void> submit(std::function<void()> fn) {
        auto task = std::make_shared<std::packaged_task<void()>>(std::move(fn));
        std::future<void> f = task->get_future();
        { std::lock_guard<std::mutex> g(m_); q_.emplace_back([task] { (*task)(); }); }
        cv_.notify_one();
        return f;
    }
    int size() const { return (int) th_.size(); }
private:
    void run() {
        for (;;) {
            std::function<void()> fn;
            { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [this] { return stop_ || !q_.empty(); }); if (stop_ && q_.empty()) return; fn = std::move(q_.front()); q_.pop_front(); }
            fn();
        }
    }
    std::vector<std::thread> th_; std::deque<std::function<void()>> q_; std::mutex m_; std::condition_variable cv_; bool stop_ = false;
};

// A decoded run of consecutive records: SoA columns + the contig id of each record.
struct Fragment {
    SoaTile soa;
    std::vector<int32_t> tid;
};

class ParallelBam {
public:
    ParallelBam(const std::string &path, int nthreads) : pool_(nthreads) {
        // header through the sequential reader; remember where the records start
        BgzfReader rd(path);
        hdr_ = read_bam_header(rd);
        start_voff_ = rd.tell();
        fp_ = fopen(path.c_str(), "rb");
        if (!fp_) throw std::runtime_error("Couldn't open " + path + " for reading!");
        setvbuf(fp_, nullptr, _IOFBF, 8 << 20);
        seek(start_voff_);
    }
    ~ParallelBam() {
        for (auto &j : jobs_) { if (j->f_inflate.valid()) j->f_inflate.wait(); if (j->f_parse.valid()) j->f_parse.wait(); }
        if (fp_) fclose(fp_);
    }
    const BamHeader &header() const { return hdr_; }
    uint64_t first_record_voffset() const { return start_voff_; }
    // run fn(0..n-1) on the decode pool and wait (used for tile-level work such as phred packing)
    void parallel_for(size_t n, const std::function<void(size_t)> &fn) {
        std::vector<std::future<void>> fs;
        for (size_t k = 1; k < n; ++k) fs.push_back(pool_.submit([&fn, k] { fn(k); }));
        if (n) fn(0);
        for (auto &f : fs) f.get();
    }

    // Restart decoding at a BGZF virtual offset that points at a record start (from a BAI, or first_record_voffset()).
    void seek(uint64_t voff) {
        for (auto &j : jobs_) { if (j->f_inflate.valid()) j->f_inflate.wait(); if (j->f_parse.valid()) j->f_parse.wait(); }
        jobs_.clear(); pending_.clear();
        file_off_ = (int64_t)(voff >> 16); skip_ = (size_t)(voff & 0xffff); eof_ = false; file_eof_ = false; rolling_.clear(); rpos_ = 0; rolling_base_ = file_off_;
        if (fseeko(fp_, file_off_, SEEK_SET) != 0) throw std::runtime_error("seek failed");
    }

    // Next fragment in file order; nullptr at end of file.
    std::unique_ptr<Fragment> next() {
        for (;;) {
            top_up();
            if (jobs_.empty()) return nullptr;
            // stitch + launch the parse of every inflated job, in order
            for (auto &j : jobs_) {
                if (j->stitched) continue;
                if (j->f_inflate.wait_for(std::chrono::seconds(0)) != std::future_status::ready && &j != &jobs_.front()) break;
                j->f_inflate.get();
                stitch(*j);
                Job *jp = j.get();
                j->f_parse = pool_.submit([jp] { jp->parse(); });
            }
            Job &head = *jobs_.front();
            if (!head.stitched) continue;
            head.f_parse.get();
            std::unique_ptr<Fragment> out = std::move(head.frag);
            jobs_.pop_front();
            if (out->soa.n() == 0 && !(jobs_.empty() && eof_)) continue;   // a job that only completed a straddling record elsewhere
            return out;
        }
    }

private:
    struct Job {
        std::vector<uint8_t> comp; std::vector<std::pair<uint32_t, uint32_t>> blocks;   // (offset in comp, total block size)
        std::vector<uint8_t> ubuf; std::vector<uint8_t> head_rec;                         // head_rec: a record completed from the previous job's tail
        std::vector<uint32_t> rec_off;                                                     // starts of whole records in ubuf (pointing at the length prefix)
        std::future<void> f_inflate, f_parse; bool stitched = false;
        std::unique_ptr<Fragment> frag;
        void inflate_all() {
            // sizes first (ISIZE trailer), then inflate block by block into place
            size_t tot = 0; std::vector<size_t> uoff(blocks.size());
            for (size_t b = 0; b < blocks.size(); ++b) { const uint8_t *p = comp.data() + blocks[b].first; uint32_t bs = blocks[b].second; uoff[b] = tot; tot += le32(p + bs - 4); }
            ubuf.resize(tot);
            z_stream zs; memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, -15) != Z_OK) throw std::runtime_error("inflateInit2");
            for (size_t b = 0; b < blocks.size(); ++b) {
                const uint8_t *p = comp.data() + blocks[b].first; uint32_t bs = blocks[b].second;
                int xlen = p[10] | (p[11] << 8);
                uint32_t isize = le32(p + bs - 4);
                if (isize == 0) continue;
                inflateReset(&zs);
                zs.next_in = (Bytef *)(p + 12 + xlen); zs.avail_in = bs - 12 - (uint32_t) xlen - 8;
                zs.next_out = ubuf.data() + uoff[b]; zs.avail_out = isize;
                if (inflate(&zs, Z_FINISH) != Z_STREAM_END || zs.total_out != isize) { inflateEnd(&zs); throw std::runtime_error("BGZF inflate failed"); }
            }
            inflateEnd(&zs);
            std::vector<uint8_t>().swap(comp);
        }
        void parse() {
            frag.reset(new Fragment());
            BamRec r;
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

    // read compressed bytes and cut jobs at block boundaries
    void top_up() {
        const size_t want_jobs = (size_t) pool_.size() * 3 + 2;
        while (!eof_ && jobs_.size() < want_jobs) {
            std::unique_ptr<Job> j(new Job());
            static const size_t target = [] { const char *e = getenv("MD_DECODE_JOB_BYTES"); size_t v = e ? (size_t) atol(e) : 0; return v ? v : (size_t) 1 << 20; }();   // ~1 MB of compressed blocks per job
            size_t used = 0;
            for (;;) {
                // make sure a whole block header + block is in the rolling buffer
                if (rolling_.size() - rpos_ < 18) { if (!fill()) break; if (rolling_.size() - rpos_ < 18) break; }
                const uint8_t *p = rolling_.data() + rpos_;
                if (p[0] != 31 || p[1] != 139 || !(p[3] & 4)) throw std::runtime_error("not a BGZF block");
                int xlen = p[10] | (p[11] << 8), bsize = -1;
                if (rolling_.size() - rpos_ < (size_t) 12 + xlen) { if (!fill()) throw std::runtime_error("truncated BGZF header"); continue; }
                for (int off = 0; off + 4 <= xlen;) { int slen = p[12 + off + 2] | (p[12 + off + 3] << 8); if (p[12 + off] == 'B' && p[12 + off + 1] == 'C' && slen == 2) bsize = p[12 + off + 4] | (p[12 + off + 5] << 8); off += 4 + slen; }
                if (bsize < 0) throw std::runtime_error("BGZF block without BC field");
                const size_t bs = (size_t) bsize + 1;
                if (rolling_.size() - rpos_ < bs) { if (!fill()) throw std::runtime_error("truncated BGZF block"); continue; }
                j->blocks.emplace_back((uint32_t) j->comp.size(), (uint32_t) bs);
                j->comp.insert(j->comp.end(), rolling_.begin() + (ptrdiff_t) rpos_, rolling_.begin() + (ptrdiff_t)(rpos_ + bs));
                rpos_ += bs; used += bs;
                if (used >= target) break;
            }
            if (j->blocks.empty()) { eof_ = true; break; }
            Job *jp = j.get();
            j->f_inflate = pool_.submit([jp] { jp->inflate_all(); });
            jobs_.push_back(std::move(j));
        }
    }
    bool fill() {
        if (file_eof_) return false;
        if (rpos_ > 0) { rolling_.erase(rolling_.begin(), rolling_.begin() + (ptrdiff_t) rpos_); rpos_ = 0; }
        size_t old = rolling_.size();
        rolling_.resize(old + (4 << 20));
        size_t got = fread(rolling_.data() + old, 1, 4 << 20, fp_);
        rolling_.resize(old + got);
        if (got == 0) file_eof_ = true;
        return got > 0;
    }
    // record boundaries of one job; completes the previous job's straddling record
    void stitch(Job &j) {
        size_t off = skip_; skip_ = 0;
        const size_t U = j.ubuf.size();
        if (!pending_.empty()) {
            // pending_ holds the first bytes of a record that began in an earlier job
            size_t have = pending_.size();
            if (have < 4) { size_t take = std::min(4 - have, U - std::min(off, U)); pending_.insert(pending_.end(), j.ubuf.begin() + (ptrdiff_t) off, j.ubuf.begin() + (ptrdiff_t)(off + take)); off += take; have = pending_.size(); }
            if (have >= 4) {
                size_t need = 4 + (size_t) le32(pending_.data()) - have;
                size_t take = std::min(need, U - off);
                pending_.insert(pending_.end(), j.ubuf.begin() + (ptrdiff_t) off, j.ubuf.begin() + (ptrdiff_t)(off + take)); off += take;
                if (take == need) { j.head_rec.swap(pending_); pending_.clear(); }
            }
        }
        if (pending_.empty()) {
            while (off + 4 <= U) {
                size_t len = 4 + (size_t) le32(j.ubuf.data() + off);
                if (off + len > U) break;
                j.rec_off.push_back((uint32_t) off); off += len;
            }
            if (off < U) pending_.assign(j.ubuf.begin() + (ptrdiff_t) off, j.ubuf.end());
        }
        j.stitched = true;
    }

    ThreadPool pool_;
    BamHeader hdr_;
    FILE *fp_ = nullptr;
    uint64_t start_voff_ = 0;
    int64_t file_off_ = 0, rolling_base_ = 0; size_t skip_ = 0; bool eof_ = false, file_eof_ = false;
    std::vector<uint8_t> rolling_; size_t rpos_ = 0;
    std::vector<uint8_t> pending_;
    std::deque<std::unique_ptr<Job>> jobs_;
};

// Same contract as Tiler (tiles.hpp), fed by fragments.
class FragTiler {
public:
    FragTiler(ParallelBam &pb, std::unique_ptr<Fragment> &cur, size_t &cur_i, int tid, uint32_t reg_beg, uint32_t reg_end, size_t target_reads)
        : pb_(pb), cur_(cur), i_(cur_i), tid_(tid), reg_beg_(reg_beg), reg_end_(reg_end), target_(target_reads), cur_beg_(reg_beg) {}
    bool next(SoaTile &t, SoaTile &carry) {
        if (done_) return false;
        t.clear(); t.tid = tid_; t.beg = cur_beg_;
        for (size_t i = 0; i < carry.n(); ++i) if (span_end(carry, i) > cur_beg_) t.add_from(carry, i);
        carry.clear();
        uint32_t cut = reg_end_;
        bool stream_end = false;
        for (;;) {
            if (!cur_ || i_ >= cur_->soa.n()) { cur_ = pb_.next(); i_ = 0; if (!cur_) { stream_end = true; break; } if (cur_->soa.n() == 0) continue; }
            const SoaTile &f = cur_->soa;
            const int32_t rt = cur_->tid[i_];
            if (rt != tid_) { if (rt > tid_ || rt < 0) { stream_end = true; break; } ++i_; continue; }
            if ((uint32_t) f.pos[i_] >= reg_end_) { stream_end = true; break; }
            if (t.n() >= target_ && (uint32_t) f.pos[i_] > cur_beg_ && f.pos[i_] > last_pos_) { cut = (uint32_t) f.pos[i_]; break; }
            // take a run of records of this contig that stays below the region end (and, once the tile is full, at one position)
            size_t j = i_;
            const size_t room = t.n() >= target_ ? 1 : target_ - t.n();
            while (j < f.n() && j - i_ < room && cur_->tid[j] == tid_ && (uint32_t) f.pos[j] < reg_end_) ++j;
            // records entirely left of the region start are dropped (index semantics: endpos > beg)
            size_t a = i_;
            while (a < j) {
                while (a < j && !((int64_t) std::max(f.rend[a], f.pos[a] + 1) > (int64_t) reg_beg_)) ++a;
                size_t b = a;
                while (b < j && (int64_t) std::max(f.rend[b], f.pos[b] + 1) > (int64_t) reg_beg_) ++b;
                if (b > a) { t.add_range_from(f, a, b); last_pos_ = f.pos[b - 1]; }
                a = b;
            }
            i_ = j;
        }
        t.end = cut;
        // reads reaching beyond the cut (or beyond the region end) are handed on: to the next tile, or to the tiler of the adjacent region
        for (size_t i = 0; i < t.n(); ++i) if (span_end(t, i) > cut) carry.add_from(t, i);
        if (stream_end) done_ = true; else cur_beg_ = cut;
        return true;
    }
private:
    static uint32_t span_end(const SoaTile &t, size_t i) { return (uint32_t) std::max(t.rend[i], t.pos[i] + 1); }
    ParallelBam &pb_; std::unique_ptr<Fragment> &cur_; size_t &i_;
    int tid_; uint32_t reg_beg_, reg_end_; size_t target_; uint32_t cur_beg_; bool done_ = false; int32_t last_pos_ = -1;
};

}  // namespace mdhost
