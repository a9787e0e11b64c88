// Internal host-side declarations shared by pack.cpp, ingest.cpp, device_ctx.cu and report.cpp.
#pragma once
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/trew_b200.h"

namespace trew {

struct ReadRef { const char* ptr; uint32_t len; };

struct BatchView {
    uint32_t* bit_off; uint32_t* hi; uint32_t* lo; uint32_t* val;
    size_t plane_words; size_t bytes;
};

size_t batch_bytes(uint32_t n_reads, uint64_t total_bases);
void batch_layout(void* dst, uint32_t n_reads, uint64_t total_bases, BatchView* v);
void pack_prepare(const ReadRef* reads, uint32_t n, const uint32_t* range_starts, int n_ranges, const BatchView& v);
// buf_end: one past the last readable byte of the chunk the reads point into (nullptr: unknown)
void pack_range(const ReadRef* reads, uint32_t r0, uint32_t r1, const BatchView& v, const char* buf_end);

// Minimal fork-join pool: run(n, fn) calls fn(i) for i in [0, n) on the workers plus the caller.
class Pool {
public:
    explicit Pool(int n_threads) : stop_(false), gen_(0), next_(0), n_tasks_(0), pending_(0) {
        for (int i = 0; i + 1 < n_threads; i++) th_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> l(mu_); stop_ = true; gen_++; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    int size() const { return (int)th_.size() + 1; }
    void run(int n, const std::function<void(int)>& fn) {
        if (n <= 0) return;
        if (th_.empty() || n == 1) { for (int i = 0; i < n; i++) fn(i); return; }
        {
            std::lock_guard<std::mutex> l(mu_);
            fn_ = &fn; n_tasks_ = n; next_ = 0; pending_ = n; gen_++;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> l(mu_);
        done_.wait(l, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    void work() {
        for (;;) {
            int i;
            const std::function<void(int)>* f;
            {
                std::lock_guard<std::mutex> l(mu_);
                if (fn_ == nullptr || next_ >= n_tasks_) return;
                i = next_++; f = fn_;
            }
            (*f)(i);
            {
                std::lock_guard<std::mutex> l(mu_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            work();
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    bool stop_;
    uint64_t gen_;
    const std::function<void(int)>* fn_ = nullptr;
    int next_, n_tasks_, pending_;
};

// ingest.cpp: FASTQ / FASTQ.gz record reader with the reference's record semantics
struct IngestResult { int status; std::string message; };
typedef std::function<int(const char* buf1, const std::vector<int32_t>& locs1, const char* buf2,
                          const std::vector<int32_t>& locs2)> ChunkSink;
IngestResult ingest_file(int mode, int slice_length, const char* file1, bool gz1, const char* file2, bool gz2,
                         size_t chunk_bytes, const ChunkSink& sink);

}  // namespace trew
