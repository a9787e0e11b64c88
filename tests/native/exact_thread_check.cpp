// Host build of the thread-per-read exact routing (trew_b200/csrc/exact_thread.cuh is plain scalar code): packs reads
// with the same planar layout as the device batches and runs route_short_thread on each, so tests/test_exact_thread.py
// can diff the result against the CPU oracle without a GPU.  Test infrastructure.
#include <cstdint>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "../../trew_b200/csrc/exact_thread.cuh"

using namespace trew::et;

namespace {
int code_of(unsigned char ch) {   // codes[], src/kmer.cpp:14-31: T=0 G=1 C=2 A=3, anything else invalid
    switch (ch) {
        case 'T': case 't': return 0;
        case 'G': case 'g': return 1;
        case 'C': case 'c': return 2;
        case 'A': case 'a': return 3;
        default: return -1;
    }
}
typedef std::map<std::tuple<int, int, uint64_t, uint64_t>, uint64_t> Tables;   // (table, k, key hi, key lo) -> count
// bases [off, off + len) of mate `mate` become the current window (what the device's PlaneLoad does from the packed batch)
struct StringLoad {
    Mem m; const char* s[2];
    void operator()(int mate, int off, int len) const {
        for (int j = 0; j < kReadWords + 2; j++) { m[W_H + j] = 0; m[W_L + j] = 0; }
        for (int j = 0; j < kReadWords; j++) m[W_V + j] = 0;
        for (int i = 0; i < len; i++) {
            const int c = code_of((unsigned char)s[mate][off + i]);
            if (c >= 0) {
                m[W_V + (i >> 5)] |= 1u << (i & 31); m[W_H + (i >> 5)] |= (u32)(c >> 1) << (i & 31);
                m[W_L + (i >> 5)] |= (u32)(c & 1) << (i & 31);
            }
        }
    }
};
struct Collect {
    Tables* m;
    void operator()(int table, int k, u64 key, uint64_t count) { (*m)[std::make_tuple(table, k, (uint64_t)0, key)] += count; }
    void operator()(int table, int k, u128 key, uint64_t count) {
        (*m)[std::make_tuple(table, k, (uint64_t)(key >> 64), (uint64_t)key)] += count;
    }
};
long write_out(const Tables& tables, int32_t* out_table, int32_t* out_k, uint64_t* out_key, uint64_t* out_key_hi, uint64_t* out_count, long cap) {
    if ((long)tables.size() > cap) return -1;
    long i = 0;
    for (auto& kv : tables) {
        out_table[i] = std::get<0>(kv.first); out_k[i] = std::get<1>(kv.first); out_key[i] = std::get<3>(kv.first);
        if (out_key_hi) out_key_hi[i] = std::get<2>(kv.first);
        out_count[i] = kv.second;
        i++;
    }
    return i;
}

template <class K>
long scan_reads(const char* buf, const int32_t* locs, int n_reads, int min_mer, int max_mer, const unsigned short* thr_low,
                const unsigned short* thr_high, Tables& tables, int32_t* bailed_index) {
    Collect emit{&tables};
    long bailed = 0;
    for (int r = 0; r < n_reads; r++) {
        const int st = locs[2 * r], nd = locs[2 * r + 1];
        const int n = nd >= st ? nd - st + 1 : 0;
        u32 work[Lay<K>::WORDS];
        memset(work, 0xA5, sizeof(work));   // the workspace is not cleared between reads on the device either
        Mem m{work, 1};
        bool ok = true;
        if (n <= kMaxRead) {
            StringLoad load{m, {buf + st, nullptr}};
            ok = route_short_thread<K>(m, n, 7u, min_mer, max_mer, thr_low, thr_high, load, emit);
        } else {
            ok = n < 2 * min_mer;
        }
        if (!ok) { if (bailed_index) bailed_index[bailed] = r; bailed++; }
    }
    return bailed;
}
}  // namespace


namespace {
template <class K>
long scan_pairs(const char* buf1, const int32_t* locs1, const char* buf2, const int32_t* locs2, int n_pairs, int min_mer, int max_mer,
                const unsigned short* thr_low, const unsigned short* thr_high, Tables& tables, int32_t* bailed_index) {
    Collect emit{&tables};
    long bailed = 0;
    for (int r = 0; r < n_pairs; r++) {
        const int n1 = locs1[2 * r + 1] >= locs1[2 * r] ? locs1[2 * r + 1] - locs1[2 * r] + 1 : 0;
        const int n2 = locs2[2 * r + 1] >= locs2[2 * r] ? locs2[2 * r + 1] - locs2[2 * r] + 1 : 0;
        u32 work[Lay<K>::WORDS];
        memset(work, 0xA5, sizeof(work));
        Mem m{work, 1};
        StringLoad load{m, {buf1 + locs1[2 * r], buf2 + locs2[2 * r]}};
        const bool ok = route_pair_thread<K>(m, n1, n2, min_mer, max_mer, thr_low, thr_high, load, emit);
        if (!ok) { if (bailed_index) bailed_index[bailed] = r; bailed++; }
    }
    return bailed;
}
}  // namespace


extern "C" {

// thr_low / thr_high: 1025 entries each (min M with (double)M / T >= baseline), built by the caller with IEEE doubles.
// Returns the number of table entries written (up to cap), or -1 when cap is too small; *n_bailed = reads outside the
// thread path's limits (the device hands those to the warp kernel).
long etc_scan_reads(const char* buf, const int32_t* locs, int n_reads, int min_mer, int max_mer, const unsigned short* thr_low,
                    const unsigned short* thr_high, int32_t* out_table, int32_t* out_k, uint64_t* out_key, uint64_t* out_count, long cap,
                    long* n_bailed, int32_t* bailed_index) {
    Tables tables;
    const long bailed = scan_reads<u64>(buf, locs, n_reads, min_mer, max_mer, thr_low, thr_high, tables, bailed_index);
    if (n_bailed) *n_bailed = bailed;
    return write_out(tables, out_table, out_k, out_key, nullptr, out_count, cap);
}
// the 128-bit instantiation (MAX_MER <= 64); out_key_hi receives bits 64..127 of the keys
long etc_scan_reads_wide(const char* buf, const int32_t* locs, int n_reads, int min_mer, int max_mer, const unsigned short* thr_low,
                         const unsigned short* thr_high, int32_t* out_table, int32_t* out_k, uint64_t* out_key, uint64_t* out_key_hi,
                         uint64_t* out_count, long cap, long* n_bailed, int32_t* bailed_index) {
    Tables tables;
    const long bailed = scan_reads<u128>(buf, locs, n_reads, min_mer, max_mer, thr_low, thr_high, tables, bailed_index);
    if (n_bailed) *n_bailed = bailed;
    return write_out(tables, out_table, out_k, out_key, out_key_hi, out_count, cap);
}

// the same for pairs (buffer_task_pair): mate i of pair r is locs{1,2}[2r .. 2r+1] in buf{1,2}
long etc_scan_pairs(const char* buf1, const int32_t* locs1, const char* buf2, const int32_t* locs2, int n_pairs, int min_mer, int max_mer,
                    const unsigned short* thr_low, const unsigned short* thr_high, int32_t* out_table, int32_t* out_k, uint64_t* out_key,
                    uint64_t* out_count, long cap, long* n_bailed, int32_t* bailed_index) {
    Tables tables;
    const long bailed = scan_pairs<u64>(buf1, locs1, buf2, locs2, n_pairs, min_mer, max_mer, thr_low, thr_high, tables, bailed_index);
    if (n_bailed) *n_bailed = bailed;
    return write_out(tables, out_table, out_k, out_key, nullptr, out_count, cap);
}
long etc_scan_pairs_wide(const char* buf1, const int32_t* locs1, const char* buf2, const int32_t* locs2, int n_pairs, int min_mer, int max_mer,
                         const unsigned short* thr_low, const unsigned short* thr_high, int32_t* out_table, int32_t* out_k, uint64_t* out_key,
                         uint64_t* out_key_hi, uint64_t* out_count, long cap, long* n_bailed, int32_t* bailed_index) {
    Tables tables;
    const long bailed = scan_pairs<u128>(buf1, locs1, buf2, locs2, n_pairs, min_mer, max_mer, thr_low, thr_high, tables, bailed_index);
    if (n_bailed) *n_bailed = bailed;
    return write_out(tables, out_table, out_k, out_key, out_key_hi, out_count, cap);
}

// long reads (buffer_task_long) through the three steps of the thread path, run one after the other per read
long etc_scan_long(const char* buf, const int32_t* locs, int n_reads, int min_mer, int max_mer, int slice_len, const unsigned short* thr_low,
                   const unsigned short* thr_high, int32_t* out_table, int32_t* out_k, uint64_t* out_key, uint64_t* out_count, long cap,
                   long* n_bailed, int32_t* bailed_index) {
    Tables tables;
    Collect emit{&tables};
    long bailed = 0;
    for (int r = 0; r < n_reads; r++) {
        const int st = locs[2 * r], n = locs[2 * r + 1] >= st ? locs[2 * r + 1] - st + 1 : 0;
        if (n < slice_len) continue;   // the reader drops these (src/kmer.cpp:1184)
        u32 work[Lay<u64>::WORDS];
        memset(work, 0xA5, sizeof(work));
        Mem m{work, 1};
        StringLoad load{m, {buf + st, nullptr}};
        ClsSpill<u64> x;
        bool ok = max_mer <= 32 && slice_len <= kMaxRead;
        if (ok) {
            LongGeom g(n, slice_len);
            std::vector<u32> stats((size_t)g.snum + 1, 0u);
            for (int t = 1; t <= g.snum; t++)
                if (g.len(t) <= kMaxRead) stats[(size_t)t] = long_slice_stats(m, g, t, min_mer, max_mer, thr_low, thr_high, load, x);
            std::vector<LongTask> tasks;
            auto stat = [&](int t) { return stats[(size_t)t]; };
            auto task = [&](const LongTask& tk) { tasks.push_back(tk); };
            ok = long_walk(g, stat, task);
            if (ok) for (const LongTask& tk : tasks) long_emit(m, g, tk, load, x, emit);
        }
        if (!ok) { if (bailed_index) bailed_index[bailed] = r; bailed++; }
    }
    if (n_bailed) *n_bailed = bailed;
    return write_out(tables, out_table, out_k, out_key, nullptr, out_count, cap);
}

}  // extern "C"
