// Internal host-side declarations shared by pack.cpp, ingest.cpp, device_ctx.cu and report.cpp.
#pragma once
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/trew_b200.h"

namespace trew {

// A raw chunk as the reference's reader threads produce it (QueueData / PairQueueData, src/kmer.h:93-103): text
// buffer(s) plus inclusive (st, nd) offsets.  unit == 2: read r is mate (r & 1) of pair (r >> 1).
struct ChunkView {
    const char* buf[2];
    const int32_t* locs[2];
    const char* end[2];   // one past the last byte the caller vouches for per buffer (nullptr: unknown)
    uint32_t unit;        // 1 single / long, 2 paired
    inline const char* start(uint32_t r) const {
        const int sd = unit == 2 ? (int)(r & 1u) : 0;
        return buf[sd] + locs[sd][2 * (size_t)(unit == 2 ? r >> 1 : r)];
    }
    inline void get(uint32_t r, const char*& p, uint32_t& len, size_t& slack) const {
        const int sd = unit == 2 ? (int)(r & 1u) : 0;
        const uint32_t i = unit == 2 ? r >> 1 : r;
        const int32_t st = locs[sd][2 * (size_t)i], nd = locs[sd][2 * (size_t)i + 1];
        len = nd >= st ? (uint32_t)(nd - st + 1) : 0u;
        p = buf[sd] + st;
        slack = end[sd] && end[sd] > p + len ? (size_t)(end[sd] - (p + len)) : 0;
    }
};

struct BatchView {
    uint32_t* bit_off; uint32_t* hi; uint32_t* lo; uint32_t* val;
    size_t plane_words; size_t bytes;
};

size_t batch_bytes(uint32_t n_reads, uint64_t total_bases);
void batch_layout(void* dst, uint32_t n_reads, uint64_t total_bases, BatchView* v);
// total bases and longest read of reads [r0, r1)
void chunk_stats(const ChunkView& cv, uint32_t r0, uint32_t r1, uint64_t* bases, uint32_t* max_len);
// Zero the 64-bit plane units nobody may store to (the unit holding each range's first bit) and the tail pad.
void pack_prepare(const uint64_t* range_bit0, int n_ranges, uint64_t total_bases, const BatchView& v);
// Pack reads [r0, r1) of the chunk as reads out0 .. of the batch, the first base at bit position bit0: writes
// bit_off[out0 ..] and the three planes.  Ranges may be packed concurrently: the bits that fall into the range's
// first 64-bit unit (possibly shared with the previous range) are returned in side[] instead of being stored.
// Sparse validity: instead of the val plane a packer can emit one 12-byte record (u32 bit position of a 64-base
// block, u64 mask of its invalid bases) per block that holds any base other than A/C/G/T.  Appending is branch-free
// -- the record is always written and the cursor advances only when the mask is non-zero -- so the cost does not
// depend on how the N's are spread; the buffer is sized for the worst case (every block) up front.
constexpr size_t kInvRecBytes = 12;
inline unsigned char* inv_record(unsigned char* p, uint32_t pos, uint64_t z) {
    memcpy(p, &pos, 4);
    memcpy(p + 4, &z, 8);
    return p + kInvRecBytes * (size_t)(z != 0);
}
struct InvList {
    std::unique_ptr<unsigned char[]> buf;
    size_t cap = 0;              // records
    unsigned char* p = nullptr;  // write cursor
    void start(size_t max_records) {
        if (max_records + 1 > cap) { cap = max_records + 1; buf.reset(new unsigned char[cap * kInvRecBytes]); }
        p = buf.get();
    }
    void put(uint32_t pos, uint64_t z) { p = inv_record(p, pos, z); }
    size_t size() const { return p ? (size_t)(p - buf.get()) / kInvRecBytes : 0; }
    size_t bytes() const { return size() * kInvRecBytes; }
    void get(size_t i, uint32_t* pos, uint64_t* z) const { memcpy(pos, buf.get() + i * kInvRecBytes, 4); memcpy(z, buf.get() + i * kInvRecBytes + 4, 8); }
};
// With inv != nullptr the range's blocks with invalid bases are recorded there (positions are plane bit
// coordinates); skip_val (needs inv) additionally allows the packer to leave the val plane unwritten where that
// saves work.  range_bases = the bases of reads [r0, r1) (sizes the record buffer).
void pack_chunk_range(const ChunkView& cv, uint32_t r0, uint32_t r1, uint32_t out0, uint64_t bit0, const BatchView& v,
                      uint64_t side[3], InvList* inv, bool skip_val, uint64_t range_bases);
// After all ranges are packed: OR every range's side bits into its first unit (single-threaded, n_ranges items).
void pack_fixup(const uint64_t* range_bit0, const uint64_t (*side)[3], int n_ranges, const BatchView& v);

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

// Growable byte buffer for file blocks: 2 MiB aligned and advised to use huge pages, never zero-filled, and kept
// across files by the owner (first-touch page faults on a fresh 256 MiB block cost more than reading it).
struct GrowBuf {
    char* data = nullptr;
    size_t cap = 0;
    void reserve(size_t n, size_t keep);   // at least n bytes; the first `keep` bytes survive a reallocation
    ~GrowBuf();
    GrowBuf() = default;
    GrowBuf(const GrowBuf&) = delete;
    GrowBuf& operator=(const GrowBuf&) = delete;
};
struct IngestScratch { GrowBuf a, b; };

// inflate.cpp: raw DEFLATE decoder, resumable between symbols.  run() consumes input and produces output until one
// of them runs out or the stream ends; unless in_final is set it wants at least 1 KiB of input in hand whenever a
// block header is due (it returns kNeedInput otherwise), so callers feed it from a buffer they top up.
class Inflater {
public:
    enum Status { kNeedInput, kOutputFull, kStreamEnd, kError };
    Inflater() { reset(); }
    void reset();   // start of a new DEFLATE stream
    Status run(const uint8_t* in, size_t in_len, bool in_final, size_t* in_used, uint8_t* out, size_t out_cap, size_t* out_used);
    // after kStreamEnd: the whole bytes that were pulled into the bit buffer but belong to whatever follows the stream
    size_t leftover(uint8_t out[8]);
    // continue a stream somebody else decoded so far: the next block header starts `bit` bits (0..7) into *first_byte, the
    // last hist_len (<= 32768) bytes produced were hist.  The caller feeds run() with the input behind first_byte.
    void resume(uint8_t first_byte, unsigned bit, const uint8_t* hist, size_t hist_len);

private:
    enum State { kHeader, kStored, kHuffman, kDone };
    static constexpr size_t kWindow = 32768, kLitCap = 4096, kDistCap = 1024;   // tables: worst cases are 2342 / 402 entries
    Status run_impl(const uint8_t* in, size_t in_len, bool in_final, size_t* in_used, uint8_t* out, size_t out_cap, size_t* out_used);
    Status run_generic(const uint8_t* in, size_t in_len, bool in_final, size_t* in_used, uint8_t* out, size_t out_cap, size_t* out_used);
    Status run_bmi2(const uint8_t* in, size_t in_len, bool in_final, size_t* in_used, uint8_t* out, size_t out_cap, size_t* out_used);
    bool read_header(const uint8_t*& ip, const uint8_t* in_end);
    void save_history(const uint8_t* out, size_t produced);
    uint64_t bitbuf_;
    int bitcnt_;
    State state_;
    bool final_;
    uint32_t stored_left_;
    size_t hist_len_;
    uint64_t lit_[kLitCap];
    uint32_t dist_[kDistCap];
    uint8_t hist_[kWindow];
};

// pinflate.cpp: one DEFLATE stream held in memory, decoded by all threads of a pool (speculative block starts, 16-bit
// symbols for what a segment copies from before its start; see the file's header)
struct RawBytes {   // growing byte array that is never zero-filled
    uint8_t* d = nullptr; size_t n = 0, cap = 0;
    RawBytes() {}
    RawBytes(const RawBytes&) = delete;
    RawBytes& operator=(const RawBytes&) = delete;
    ~RawBytes() { free(d); }
    bool reserve(size_t want) {
        if (want <= cap) return true;
        size_t c = cap ? cap : ((size_t)1 << 20);
        while (c < want) c += c / 2;
        uint8_t* p = (uint8_t*)realloc(d, c);
        if (!p) return false;
        d = p; cap = c;
        return true;
    }
};

class ParallelInflate {
public:
    ParallelInflate();
    ~ParallelInflate();
    // the stream starts at byte first_byte of in[0, n) (right behind a gzip header); no history
    void start(const uint8_t* in, size_t n, size_t first_byte);
    // Decodes the next part (about pool->size() segments of compressed input) and appends its bytes to `out`.
    // *member_end: the final block ended, *next_byte = offset of the byte behind the stream.  false: bad data, *err says what.
    bool next(Pool* pool, RawBytes& out, bool* member_end, size_t* next_byte, const char** err);
    // the same in two steps, for callers that place the bytes themselves: decode() keeps the part as symbols and says
    // how many bytes it is, emit() writes bytes [off, off + len) of it to dst (any ranges, any order, until the next decode())
    bool decode(Pool* pool, size_t* total, bool* member_end, size_t* next_byte, const char** err);
    void emit(Pool* pool, uint8_t* dst, size_t off, size_t len);
    size_t segment_bytes() const { return seg_bytes_; }
    // where the stream stands after the last decode(): next bit, and the last <= 32 KiB produced (to hand the rest of the
    // stream to the sequential decoder)
    uint64_t position() const { return pos_; }
    const std::vector<uint8_t>& window() const { return window_; }
    size_t last_chain() const { return chain_.size(); }
    // how the work went (tests, traces): calls of next(), segments with a block start found, segments on the chains
    uint64_t stat_calls = 0, stat_found = 0, stat_chained = 0;

private:
    struct Segment;
    std::vector<Segment*> seg_;        // kept between calls: their symbol arrays are large
    std::vector<int> chain_;           // the segments of the last decode() that follow each other
    std::vector<std::vector<uint8_t>> wins_;   // ... and each one's symbol -> byte table (built from the window before it)
    const uint8_t* in_ = nullptr;
    size_t n_ = 0;
    uint64_t pos_ = 0;                 // next bit of the stream
    std::vector<uint8_t> window_;      // the last <= 32 KiB produced
    size_t seg_bytes_;
};

// ingest.cpp: FASTQ / FASTQ.gz record reader with the reference's record semantics
struct IngestResult { int status; std::string message; };
typedef std::function<int(const char* buf1, const std::vector<int32_t>& locs1, const char* buf2,
                          const std::vector<int32_t>& locs2)> ChunkSink;
IngestResult ingest_file(int mode, int slice_length, const char* file1, bool gz1, const char* file2, bool gz2,
                         size_t chunk_bytes, const ChunkSink& sink, Pool* pool = nullptr, IngestScratch* scratch = nullptr);

}  // namespace trew
