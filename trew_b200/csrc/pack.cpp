// Host side of the hot path's input: ASCII sequence lines -> planar 2-bit batches.
//
// Restates codes[] (src/kmer.cpp:14-31): T=0 G=1 C=2 A=3 for either case, every other byte invalid.
// The two code bits and the validity bit are written as three bit-planes (see include/trew_b200.h,
// trew_batch).  With AVX2 a 32-byte block becomes three 32-bit masks via movemask:
//     x1 = bit 2 of the byte, x0 = bit 1:   A -> 00, C -> 01, T -> 10, G -> 11
//     hi = ~x1, lo = ~(x1 ^ x0)             A -> 11, C -> 10, G -> 01, T -> 00   (the reference's codes)
#include "host_internal.h"

#include <atomic>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace trew {

namespace {

struct Lut {
    unsigned char v[256];  // bit0 = lo, bit1 = hi, bit2 = valid
    Lut() {
        memset(v, 0, sizeof(v));
        const char* s = "TGCA";
        for (int i = 0; i < 4; i++) {
            v[(unsigned char)s[i]] = (unsigned char)(4 | i);
            v[(unsigned char)(s[i] | 0x20)] = (unsigned char)(4 | i);
        }
    }
};
const Lut g_lut;

inline void masks_scalar(const unsigned char* p, int n, uint32_t& hi, uint32_t& lo, uint32_t& val) {
    uint32_t h = 0, l = 0, v = 0;
    for (int i = 0; i < n; i++) {
        unsigned c = g_lut.v[p[i]];
        l |= (uint32_t)(c & 1) << i;
        h |= (uint32_t)((c >> 1) & 1) << i;
        v |= (uint32_t)((c >> 2) & 1) << i;
    }
    hi = h; lo = l; val = v;
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) inline void masks_avx2(const unsigned char* p, uint32_t& hi, uint32_t& lo, uint32_t& val) {
    __m256i x = _mm256_loadu_si256((const __m256i*)p);
    uint32_t x1 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(x, 5));
    uint32_t x0 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(x, 6));
    __m256i lc = _mm256_or_si256(x, _mm256_set1_epi8(0x20));
    __m256i ok = _mm256_or_si256(
        _mm256_or_si256(_mm256_cmpeq_epi8(lc, _mm256_set1_epi8('a')), _mm256_cmpeq_epi8(lc, _mm256_set1_epi8('c'))),
        _mm256_or_si256(_mm256_cmpeq_epi8(lc, _mm256_set1_epi8('g')), _mm256_cmpeq_epi8(lc, _mm256_set1_epi8('t'))));
    uint32_t v = (uint32_t)_mm256_movemask_epi8(ok);
    hi = ~x1 & v; lo = ~(x1 ^ x0) & v; val = v;
}
#endif

bool have_avx2() {
#if defined(__x86_64__)
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
#else
    return false;
#endif
}

// Appends bits to the three planes starting at an arbitrary bit position.  The first and the last
// word it touches may be shared with a neighbouring range packed by another thread: those two are
// OR-ed atomically into pre-zeroed memory, interior words are plain stores.
struct BitWriter {
    uint32_t* hi; uint32_t* lo; uint32_t* val;
    uint64_t word;       // index of the next word to flush
    uint64_t ah, al, av; // accumulators
    int fill;            // valid bits in the accumulators
    bool first;

    BitWriter(uint32_t* h, uint32_t* l, uint32_t* v, uint64_t bitpos)
        : hi(h), lo(l), val(v), word(bitpos >> 5), ah(0), al(0), av(0), fill((int)(bitpos & 31)), first(true) {}

    inline void flush_word() {
        uint32_t h = (uint32_t)ah, l = (uint32_t)al, v = (uint32_t)av;
        if (first) {
            __atomic_fetch_or(&hi[word], h, __ATOMIC_RELAXED);
            __atomic_fetch_or(&lo[word], l, __ATOMIC_RELAXED);
            __atomic_fetch_or(&val[word], v, __ATOMIC_RELAXED);
            first = false;
        } else {
            hi[word] = h; lo[word] = l; val[word] = v;
        }
        word++; ah >>= 32; al >>= 32; av >>= 32; fill -= 32;
    }
    inline void put(uint32_t h, uint32_t l, uint32_t v, int nbits) {
        ah |= (uint64_t)h << fill; al |= (uint64_t)l << fill; av |= (uint64_t)v << fill;
        fill += nbits;
        if (fill >= 32) flush_word();
    }
    inline void finish() {
        if (fill > 0) {
            __atomic_fetch_or(&hi[word], (uint32_t)ah, __ATOMIC_RELAXED);
            __atomic_fetch_or(&lo[word], (uint32_t)al, __ATOMIC_RELAXED);
            __atomic_fetch_or(&val[word], (uint32_t)av, __ATOMIC_RELAXED);
        }
    }
};

// `slack` = readable bytes after the read's last byte (inside the caller's chunk): when at least 31 the tail
// block is loaded straight from the chunk and the excess lanes are masked off.
inline void pack_one(BitWriter& w, const unsigned char* s, int n, bool avx2, size_t slack) {
    int i = 0;
    uint32_t h, l, v;
#if defined(__x86_64__)
    if (avx2) {
        for (; i + 32 <= n; i += 32) { masks_avx2(s + i, h, l, v); w.put(h, l, v, 32); }
        if (i < n) {
            int m = n - i;
            uint32_t keep = (1u << m) - 1u;
            if (slack >= 31) {
                masks_avx2(s + i, h, l, v);
            } else {
                unsigned char tmp[32];
                memset(tmp, 0, sizeof(tmp));
                memcpy(tmp, s + i, (size_t)m);
                masks_avx2(tmp, h, l, v);
            }
            w.put(h & keep, l & keep, v & keep, m);
        }
        return;
    }
#endif
    for (; i < n; i += 32) {
        int m = n - i < 32 ? n - i : 32;
        masks_scalar(s + i, m, h, l, v);
        w.put(h, l, v, m);
    }
}

}  // namespace

size_t batch_bytes(uint32_t n_reads, uint64_t total_bases) {
    size_t off = ((size_t)(n_reads + 1) * 4 + 15) & ~(size_t)15;
    size_t words = (size_t)((total_bases + 31) / 32) + TREW_PLANE_PAD_WORDS;
    words = (words + 3) & ~(size_t)3;
    return off + 3 * words * 4;
}

void batch_layout(void* dst, uint32_t n_reads, uint64_t total_bases, BatchView* v) {
    size_t off = ((size_t)(n_reads + 1) * 4 + 15) & ~(size_t)15;
    size_t words = (size_t)((total_bases + 31) / 32) + TREW_PLANE_PAD_WORDS;
    words = (words + 3) & ~(size_t)3;
    v->bit_off = (uint32_t*)dst;
    v->hi = (uint32_t*)((char*)dst + off);
    v->lo = v->hi + words;
    v->val = v->lo + words;
    v->plane_words = words;
    v->bytes = off + 3 * words * 4;
}

// Pack reads[r0, r1) whose first base sits at bit position v.bit_off[r0].
void pack_range(const ReadRef* reads, uint32_t r0, uint32_t r1, const BatchView& v, const char* buf_end) {
    if (r0 >= r1) return;
    const bool avx2 = have_avx2();
    BitWriter w(v.hi, v.lo, v.val, v.bit_off[r0]);
    for (uint32_t r = r0; r < r1; r++) {
        const char* e = reads[r].ptr + reads[r].len;
        size_t slack = (buf_end && buf_end > e) ? (size_t)(buf_end - e) : 0;
        pack_one(w, (const unsigned char*)reads[r].ptr, (int)reads[r].len, avx2, slack);
    }
    w.finish();
}

// Fills bit_off and zeroes the words that pack_range() will OR into (range boundaries + tail pad).
void pack_prepare(const ReadRef* reads, uint32_t n, const uint32_t* range_starts, int n_ranges, const BatchView& v) {
    uint64_t pos = 0;
    for (uint32_t r = 0; r < n; r++) { v.bit_off[r] = (uint32_t)pos; pos += reads[r].len; }
    v.bit_off[n] = (uint32_t)pos;
    for (int i = 0; i < n_ranges; i++) {
        uint64_t wi = (uint64_t)v.bit_off[range_starts[i]] >> 5;
        v.hi[wi] = 0; v.lo[wi] = 0; v.val[wi] = 0;
    }
    for (size_t wi = (size_t)(pos >> 5); wi < v.plane_words; wi++) { v.hi[wi] = 0; v.lo[wi] = 0; v.val[wi] = 0; }
}

}  // namespace trew

extern "C" {

size_t trew_pack_bound(uint32_t n_reads, uint64_t total_bases) { return trew::batch_bytes(n_reads, total_bases); }

int trew_pack_reads(const char* buffer, const int32_t* locs, uint32_t n, void* dst, size_t dst_bytes, trew_batch* out) {
    if ((n && (!buffer || !locs)) || !dst || !out) return TREW_ERR_ARG;
    std::vector<trew::ReadRef> reads(n);
    uint64_t total = 0; uint32_t mx = 0;
    for (uint32_t i = 0; i < n; i++) {
        int32_t st = locs[2 * i], nd = locs[2 * i + 1];
        uint32_t len = nd >= st ? (uint32_t)(nd - st + 1) : 0u;
        reads[i] = trew::ReadRef{buffer + st, len};
        total += len; if (len > mx) mx = len;
    }
    if (total >= 0xffffffffULL) return TREW_ERR_ARG;
    if (trew::batch_bytes(n, total) > dst_bytes) return TREW_ERR_ARG;
    trew::BatchView v;
    trew::batch_layout(dst, n, total, &v);
    uint32_t zero = 0;
    trew::pack_prepare(reads.data(), n, &zero, n ? 1 : 0, v);
    trew::pack_range(reads.data(), 0, n, v, nullptr);
    out->n_reads = n; out->max_read_len = mx; out->bit_off = v.bit_off; out->hi = v.hi; out->lo = v.lo; out->val = v.val;
    return TREW_OK;
}

}  // extern "C"
