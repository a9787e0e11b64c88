// Host side of the hot path's input: ASCII sequence lines -> planar 2-bit batches.
//
// Restates codes[] (src/kmer.cpp:14-31): T=0 G=1 C=2 A=3 for either case, every other byte invalid.
// The two code bits and the validity bit are written as three bit-planes (see include/trew_b200.h,
// trew_batch).  A 32-byte (AVX2) or 64-byte (AVX-512BW) block becomes three bit masks:
//     x1 = bit 2 of the byte, x0 = bit 1:   A -> 00, C -> 01, T -> 10, G -> 11
//     hi = ~x1, lo = ~(x1 ^ x0)             A -> 11, C -> 10, G -> 01, T -> 00   (the reference's codes)
// The instruction set is picked at run time (scalar fallback for other CPUs; the DEVICE path has no fallback).
#include "host_internal.h"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifndef TREW_PACK_PREFETCH
// How far ahead the SIMD packers prefetch into L1 (0 = off): TREW_PACK_PREFETCH_READS reads ahead for reads of up to
// 512 bases, TREW_PACK_PREFETCH bytes ahead within a longer read.  The packers stream ~150 B per
// read in and 56 B out on every core; alone, the hardware prefetchers keep a core at ~5 GB/s of input on the
// B200 hosts measured, and an explicit prefetch 2-8 KB ahead shortens a 1 M-read batch from 1.9 to 1.3 ms on 16
// cores (host packer).
#define TREW_PACK_PREFETCH 3072
#endif
#ifndef TREW_PACK_PREFETCH_READS
#define TREW_PACK_PREFETCH_READS 16
#endif
#ifndef TREW_PACK_PREFETCH_HINT
#define TREW_PACK_PREFETCH_HINT _MM_HINT_T0
#endif

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace trew {

namespace {

constexpr uint32_t kPrefetchReads = TREW_PACK_PREFETCH_READS;

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

inline void masks_scalar(const unsigned char* p, int n, uint64_t& hi, uint64_t& lo, uint64_t& val) {
    uint64_t h = 0, l = 0, v = 0;
    for (int i = 0; i < n; i++) {
        uint64_t c = g_lut.v[p[i]];
        l |= (c & 1) << i;
        h |= ((c >> 1) & 1) << i;
        v |= ((c >> 2) & 1) << i;
    }
    hi = h; lo = l; val = v;
}

// Appends bits to the three planes starting at an arbitrary bit position, 64 bits at a time (the planes are
// 8-byte aligned; a little-endian u64 store is two consecutive u32 plane words).  Branch-free in the steady
// state: every put stores the current (possibly still partial) unit and a later put overwrites it with the
// complete value.  Ranges are packed concurrently by different threads, and a range's FIRST unit may also hold
// the last bits of the previous range: those bits are never stored here but kept in `side` and OR-ed in by
// the caller once all ranges are done (pack_fixup); every other unit has exactly one writer.
struct BitWriter {
    uint64_t* hi; uint64_t* lo; uint64_t* val;
    uint64_t unit;        // index of the 64-bit unit being filled
    uint64_t first_unit;
    uint64_t ah, al, av;  // accumulators
    int fill;             // valid bits in the accumulators
    uint64_t side[3];     // hi / lo / val bits of the first unit

    BitWriter(uint32_t* h, uint32_t* l, uint32_t* v, uint64_t bitpos)
        : hi((uint64_t*)h), lo((uint64_t*)l), val((uint64_t*)v), unit(bitpos >> 6), first_unit(bitpos >> 6), ah(0), al(0),
          av(0), fill((int)(bitpos & 63)) { side[0] = side[1] = side[2] = 0; }

    // FULL: nbits == 64.  PEEL: the first unit may still be the current one.  Bits above nbits must be zero.
    template <bool FULL, bool PEEL>
    __attribute__((always_inline)) inline void put(uint64_t h, uint64_t l, uint64_t v, int nbits) {
        const uint64_t th = ah | (h << fill), tl = al | (l << fill), tv = av | (v << fill);
        if (PEEL && unit == first_unit) { side[0] = th; side[1] = tl; side[2] = tv; }
        else { hi[unit] = th; lo[unit] = tl; val[unit] = tv; }
        const int back = 63 - fill;  // x >> (64 - fill) without the undefined shift by 64
        const uint64_t ch = (h >> 1) >> back, cl = (l >> 1) >> back, cvv = (v >> 1) >> back;
        if (FULL) {
            ah = ch; al = cl; av = cvv; unit++;
        } else {
            const int nf = fill + nbits;
            const bool adv = nf >= 64;
            ah = adv ? ch : th; al = adv ? cl : tl; av = adv ? cvv : tv;
            unit += adv ? 1 : 0;
            fill = nf & 63;
        }
    }
    bool peeling() const { return unit == first_unit; }
    // the bits carried past the last completed unit
    void finish() {
        if (fill != 0) {
            if (unit == first_unit) { side[0] = ah; side[1] = al; side[2] = av; }
            else { hi[unit] = ah; lo[unit] = al; val[unit] = av; }
        }
    }
};

template <bool PEEL>
inline void pack_read_scalar(BitWriter& w, const unsigned char* s, uint32_t len, uint32_t rpos, InvList* inv) {
    uint64_t h, l, v;
    uint32_t i = 0;
    for (; i + 64 <= len; i += 64) {
        masks_scalar(s + i, 64, h, l, v);
        if (inv) inv->put(rpos + i, ~v);
        w.put<true, PEEL>(h, l, v, 64);
    }
    if (i < len) {
        masks_scalar(s + i, (int)(len - i), h, l, v);
        const uint64_t keep = (1ULL << (len - i)) - 1ULL;
        if (inv) inv->put(rpos + i, ~v & keep);
        w.put<false, PEEL>(h, l, v, (int)(len - i));
    }
}

void pack_reads_scalar(const ChunkView& cv, uint32_t r0, uint32_t r1, uint32_t* off, uint64_t pos, BitWriter& w_out,
                       InvList* inv) {
    BitWriter w = w_out;  // local copy: its address never escapes, so the state stays in registers across the plane stores
    const char* p; uint32_t len; size_t slack;
    uint32_t r = r0;
    for (; r < r1 && w.peeling(); r++) {
        cv.get(r, p, len, slack);
        *off++ = (uint32_t)pos; pos += len;
        pack_read_scalar<true>(w, (const unsigned char*)p, len, (uint32_t)pos - len, inv);
    }
    for (; r < r1; r++) {
        cv.get(r, p, len, slack);
        *off++ = (uint32_t)pos; pos += len;
        pack_read_scalar<false>(w, (const unsigned char*)p, len, (uint32_t)pos - len, inv);
    }
    w_out = w;
}

#if defined(__x86_64__)
#define TREW_AVX2 __attribute__((target("avx2")))
#define TREW_AVX512 __attribute__((target("avx512f,avx512bw")))

// AVX2: a 32-byte block becomes three 32-bit masks.  x1 = bit 2 of the byte, x0 = bit 1: A -> 00, C -> 01,
// T -> 10, G -> 11; hi = ~x1, lo = ~(x1 ^ x0) give the reference's codes.  Validity: a 16-entry shuffle table
// keyed by the low nibble returns the one lower-case base letter with that nibble (or 0); the byte is a base
// iff it equals that letter once its case bit is set.
TREW_AVX2 __attribute__((always_inline)) inline void masks_avx2(const unsigned char* q, uint64_t keep, __m256i lut, __m256i case_bit,
                                                                uint64_t& h, uint64_t& l, uint64_t& v) {
    __m256i x = _mm256_loadu_si256((const __m256i*)q);
    uint32_t x1 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(x, 5));
    uint32_t x0 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(x, 6));
    uint32_t vv = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, x), _mm256_or_si256(x, case_bit))) & (uint32_t)keep;
    h = ~x1 & vv; l = ~(x1 ^ x0) & vv; v = vv;
}

template <bool PEEL>
TREW_AVX2 __attribute__((always_inline)) inline void pack_read_avx2(BitWriter& w, const unsigned char* s, uint32_t len, size_t slack,
                                                                    __m256i lut, __m256i case_bit, uint32_t rpos,
                                                                    InvList* inv) {
    uint64_t h, l, v, h2, l2, v2;
    uint32_t i = 0;
    for (; i + 64 <= len; i += 64) {
        masks_avx2(s + i, 0xffffffffu, lut, case_bit, h, l, v);
        masks_avx2(s + i + 32, 0xffffffffu, lut, case_bit, h2, l2, v2);
        v |= v2 << 32;
        if (inv) inv->put(rpos + i, ~v);
        w.put<true, PEEL>(h | (h2 << 32), l | (l2 << 32), v, 64);
    }
    if (i < len) {
        const int m = (int)(len - i);
        const uint64_t keep = m >= 64 ? ~0ULL : ((1ULL << m) - 1ULL);
        unsigned char tmp[64];
        const unsigned char* q = s + i;
        if (slack + (size_t)m < 64) {  // the 64-byte window would leave the caller's chunk: copy the tail
            memset(tmp, 0, sizeof(tmp));
            memcpy(tmp, q, (size_t)m);
            q = tmp;
        }
        masks_avx2(q, keep, lut, case_bit, h, l, v);
        h2 = l2 = v2 = 0;
        if (m > 32) masks_avx2(q + 32, keep >> 32, lut, case_bit, h2, l2, v2);
        v |= v2 << 32;
        if (inv) inv->put(rpos + i, ~v & keep);
        w.put<false, PEEL>(h | (h2 << 32), l | (l2 << 32), v, m);
    }
}

TREW_AVX2 void pack_reads_avx2(const ChunkView& cv, uint32_t r0, uint32_t r1, uint32_t* off, uint64_t pos, BitWriter& w_out,
                               InvList* inv) {
    BitWriter w = w_out;  // local copy: its address never escapes, so the state stays in registers across the plane stores
    const __m256i lut = _mm256_setr_epi8(0, 'a', 0, 'c', 't', 0, 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0,
                                         0, 'a', 0, 'c', 't', 0, 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i case_bit = _mm256_set1_epi8(0x20);
    const char* p; uint32_t len; size_t slack;
    uint32_t r = r0;
    for (; r < r1 && w.peeling(); r++) {
        cv.get(r, p, len, slack);
        *off++ = (uint32_t)pos; pos += len;
        pack_read_avx2<true>(w, (const unsigned char*)p, len, slack, lut, case_bit, (uint32_t)pos - len, inv);
    }
    for (; r < r1; r++) {
        cv.get(r, p, len, slack);
        *off++ = (uint32_t)pos; pos += len;
#if TREW_PACK_PREFETCH
        {
            const uint32_t ra = r + kPrefetchReads * cv.unit;
            const char* pf = len <= 512u ? cv.start(ra < r1 ? ra : r) : p + TREW_PACK_PREFETCH;
            for (uint32_t q = 0; q < len; q += 64) _mm_prefetch(pf + q, TREW_PACK_PREFETCH_HINT);
        }
#endif
        pack_read_avx2<false>(w, (const unsigned char*)p, len, slack, lut, case_bit, (uint32_t)pos - len, inv);
    }
    w_out = w;
}

// AVX-512BW: 64-byte blocks, mask registers instead of movemask, masked loads for read tails.
template <bool PEEL>
TREW_AVX512 __attribute__((always_inline)) inline void pack_read_avx512(BitWriter& w, const unsigned char* s, uint32_t len, __m512i lut,
                                                                        __m512i case_bit, __m512i b2, __m512i b1, uint32_t rpos,
                                                                        InvList* inv) {
    uint32_t i = 0;
    for (; i + 64 <= len; i += 64) {
        __m512i x = _mm512_loadu_si512((const void*)(s + i));
        uint64_t x1 = _mm512_test_epi8_mask(x, b2), x0 = _mm512_test_epi8_mask(x, b1);
        uint64_t v = _mm512_cmpeq_epi8_mask(_mm512_shuffle_epi8(lut, x), _mm512_or_si512(x, case_bit));
        if (inv) inv->put(rpos + i, ~v);
        w.put<true, PEEL>(~x1 & v, ~(x1 ^ x0) & v, v, 64);
    }
    if (i < len) {
        const int m = (int)(len - i);
        const uint64_t keep = (1ULL << m) - 1ULL;
        __m512i x = _mm512_maskz_loadu_epi8((__mmask64)keep, (const void*)(s + i));
        uint64_t x1 = _mm512_test_epi8_mask(x, b2), x0 = _mm512_test_epi8_mask(x, b1);
        uint64_t v = _mm512_cmpeq_epi8_mask(_mm512_shuffle_epi8(lut, x), _mm512_or_si512(x, case_bit)) & keep;
        if (inv) inv->put(rpos + i, ~v & keep);
        w.put<false, PEEL>(~x1 & v, ~(x1 ^ x0) & v, v, m);
    }
}

// Steady state of the AVX-512 packer for single-buffer chunks, written against register pressure (x86-64 has 16
// general registers and the generic writer keeps a dozen values alive): one plane pointer that advances plus two
// constant strides, no first-unit test.  Precondition: the writer has left its first unit.
// LIST: record the blocks with invalid bases in *inv (branch-free: see inv_record).  VAL == false (needs LIST): the val plane is not written at all -- a third
// less store traffic; the list is then the only record of validity.
// PAIR: read r is mate (r & 1) of pair (r >> 1), the mates coming from the chunk's two buffers.
template <bool LIST, bool VAL, bool PAIR>
TREW_AVX512 void pack_lean_avx512(const ChunkView& cv, uint32_t r0, uint32_t r1, uint32_t* off, BitWriter& w_out, InvList* inv) {
    const char* const buf0 = cv.buf[0];
    const char* const buf1 = PAIR ? cv.buf[1] : buf0;
    const int32_t* const locs0 = cv.locs[0];
    const int32_t* const locs1 = PAIR ? cv.locs[1] : locs0;
    constexpr uint32_t kAhead = PAIR ? 2u * kPrefetchReads : kPrefetchReads;   // the same mate, kPrefetchReads pairs on
    const __m512i lut = _mm512_broadcast_i32x4(_mm_setr_epi8(0, 'a', 0, 'c', 't', 0, 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0));
    const __m512i case_bit = _mm512_set1_epi8(0x20), b2 = _mm512_set1_epi8(4), b1 = _mm512_set1_epi8(2);
    long long* ph = (long long*)(w_out.hi + w_out.unit);
    const ptrdiff_t s1 = w_out.lo - w_out.hi, s2 = w_out.val - w_out.hi;
    uint64_t ah = w_out.ah, al = w_out.al, av = w_out.av;
    unsigned fill = (unsigned)w_out.fill;
    uint32_t pos = (uint32_t)(w_out.unit * 64 + fill);
    unsigned char* ip = LIST ? inv->p : nullptr;
    for (uint32_t r = r0; r < r1; r++) {
        const char* const buf = PAIR && (r & 1u) ? buf1 : buf0;
        const int32_t* const locs = (PAIR && (r & 1u) ? locs1 : locs0) + 2 * (size_t)(PAIR ? r >> 1 : r);
        const int32_t st = locs[0], nd = locs[1];
        const uint32_t len = nd >= st ? (uint32_t)(nd - st + 1) : 0u;
        const unsigned char* s = (const unsigned char*)buf + st;
        *off++ = pos;
#if TREW_PACK_PREFETCH
        {   // short reads: the read kPrefetchReads ahead (the bytes between reads -- FASTQ headers, qualities -- are
            // not wanted); long reads: further along the same read
            const unsigned char* pf = len <= 512u ? (const unsigned char*)buf + locs[r + kAhead < r1 ? 2 * (size_t)kPrefetchReads : 0]
                                                  : s + TREW_PACK_PREFETCH;
            for (uint32_t q = 0; q < len; q += 64) _mm_prefetch((const char*)pf + q, TREW_PACK_PREFETCH_HINT);
        }
#endif
        uint32_t i = 0;
        for (; i + 64 <= len; i += 64) {
            const __m512i x = _mm512_loadu_si512((const void*)(s + i));
            const uint64_t x1 = _mm512_test_epi8_mask(x, b2), x0 = _mm512_test_epi8_mask(x, b1);
            const uint64_t v = _mm512_cmpeq_epi8_mask(_mm512_shuffle_epi8(lut, x), _mm512_or_si512(x, case_bit));
            const uint64_t h = ~x1 & v, l = ~(x1 ^ x0) & v;
            if (LIST) ip = inv_record(ip, pos + i, ~v);
            ph[0] = (long long)(ah | (h << fill));
            ph[s1] = (long long)(al | (l << fill));
            if (VAL) ph[s2] = (long long)(av | (v << fill));
            const unsigned back = 63u - fill;
            ah = (h >> 1) >> back; al = (l >> 1) >> back;
            if (VAL) av = (v >> 1) >> back;
            ph++;
        }
        if (i < len) {
            const unsigned m = len - i;
            const uint64_t keep = (1ULL << m) - 1ULL;
            const __m512i x = _mm512_maskz_loadu_epi8((__mmask64)keep, (const void*)(s + i));
            const uint64_t x1 = _mm512_test_epi8_mask(x, b2), x0 = _mm512_test_epi8_mask(x, b1);
            const uint64_t v = _mm512_cmpeq_epi8_mask(_mm512_shuffle_epi8(lut, x), _mm512_or_si512(x, case_bit)) & keep;
            const uint64_t h = ~x1 & v, l = ~(x1 ^ x0) & v;
            if (LIST) ip = inv_record(ip, pos + i, ~v & keep);
            const uint64_t th = ah | (h << fill), tl = al | (l << fill), tv = VAL ? av | (v << fill) : 0;
            ph[0] = (long long)th;
            ph[s1] = (long long)tl;
            if (VAL) ph[s2] = (long long)tv;
            const unsigned back = 63u - fill, nf = fill + m;
            const bool adv = nf >= 64;
            ah = adv ? (h >> 1) >> back : th; al = adv ? (l >> 1) >> back : tl;
            if (VAL) av = adv ? (v >> 1) >> back : tv;
            ph += adv ? 1 : 0;
            fill = nf & 63u;
        }
        pos += len;
    }
    w_out.unit = (uint64_t)((uint64_t*)ph - w_out.hi);
    w_out.ah = ah; w_out.al = al; w_out.av = VAL ? av : 0; w_out.fill = (int)fill;
    if (LIST) inv->p = ip;
}

TREW_AVX512 void pack_reads_avx512(const ChunkView& cv, uint32_t r0, uint32_t r1, uint32_t* off, uint64_t pos, BitWriter& w_out,
                                   InvList* inv, bool skip_val) {
    BitWriter w = w_out;  // local copy: its address never escapes, so the state stays in registers across the plane stores
    const __m512i lut = _mm512_broadcast_i32x4(_mm_setr_epi8(0, 'a', 0, 'c', 't', 0, 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0));
    const __m512i case_bit = _mm512_set1_epi8(0x20), b2 = _mm512_set1_epi8(4), b1 = _mm512_set1_epi8(2);
    const char* p; uint32_t len; size_t slack;
    uint32_t r = r0;
    for (; r < r1 && w.peeling(); r++) {
        cv.get(r, p, len, slack);
        *off++ = (uint32_t)pos; pos += len;
        pack_read_avx512<true>(w, (const unsigned char*)p, len, lut, case_bit, b2, b1, (uint32_t)pos - len, inv);
    }
    if (r < r1) {   // the lean steady-state loop
        if (cv.unit == 1) {
            if (!inv) pack_lean_avx512<false, true, false>(cv, r, r1, off, w, nullptr);
            else if (skip_val) pack_lean_avx512<true, false, false>(cv, r, r1, off, w, inv);
            else pack_lean_avx512<true, true, false>(cv, r, r1, off, w, inv);
        } else {
            if (!inv) pack_lean_avx512<false, true, true>(cv, r, r1, off, w, nullptr);
            else if (skip_val) pack_lean_avx512<true, false, true>(cv, r, r1, off, w, inv);
            else pack_lean_avx512<true, true, true>(cv, r, r1, off, w, inv);
        }
    }
    w_out = w;
}
#endif

int simd_level() {  // 0 scalar, 1 AVX2, 2 AVX-512BW; TREW_PACK_SIMD=0/1/2 caps it (tests)
#if defined(__x86_64__)
    static const int v = [] {
        int lvl = 0;
        if (__builtin_cpu_supports("avx2")) lvl = 1;
        if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw")) lvl = 2;
        if (const char* e = getenv("TREW_PACK_SIMD")) { int cap = atoi(e); if (cap < lvl) lvl = cap < 0 ? 0 : cap; }
        return lvl;
    }();
    return v;
#else
    return 0;
#endif
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

void chunk_stats(const ChunkView& cv, uint32_t r0, uint32_t r1, uint64_t* bases, uint32_t* max_len) {
    uint64_t t = 0; uint32_t mx = 0;
    const char* p; uint32_t len; size_t slack;
    for (uint32_t r = r0; r < r1; r++) { cv.get(r, p, len, slack); t += len; if (len > mx) mx = len; }
    *bases = t; *max_len = mx;
}

void pack_chunk_range(const ChunkView& cv, uint32_t r0, uint32_t r1, uint32_t out0, uint64_t bit0, const BatchView& v,
                      uint64_t side[3], InvList* inv, bool skip_val, uint64_t range_bases) {
    side[0] = side[1] = side[2] = 0;
    if (inv) inv->start((size_t)(range_bases >> 6) + (size_t)(r1 > r0 ? r1 - r0 : 0u) + 2);   // >= the range's 64-base blocks
    if (r0 >= r1) return;
    BitWriter w(v.hi, v.lo, v.val, bit0);
    const int lvl = simd_level();
#if defined(__x86_64__)
    if (lvl == 2) pack_reads_avx512(cv, r0, r1, v.bit_off + out0, bit0, w, inv, inv && skip_val);
    else if (lvl == 1) pack_reads_avx2(cv, r0, r1, v.bit_off + out0, bit0, w, inv);
    else
#endif
        pack_reads_scalar(cv, r0, r1, v.bit_off + out0, bit0, w, inv);
    (void)lvl; (void)skip_val;
    w.finish();
    side[0] = w.side[0]; side[1] = w.side[1]; side[2] = w.side[2];
}

void pack_fixup(const uint64_t* range_bit0, const uint64_t (*side)[3], int n_ranges, const BatchView& v) {
    uint64_t *hi = (uint64_t*)v.hi, *lo = (uint64_t*)v.lo, *val = (uint64_t*)v.val;
    for (int i = 0; i < n_ranges; i++) {
        uint64_t u = range_bit0[i] >> 6;
        hi[u] |= side[i][0]; lo[u] |= side[i][1]; val[u] |= side[i][2];
    }
}

void pack_prepare(const uint64_t* range_bit0, int n_ranges, uint64_t total_bases, const BatchView& v) {
    // 64-bit units: the one holding each range's first bit, and everything from the unit of the last bit on
    for (int i = 0; i < n_ranges; i++) {
        uint64_t wi = (range_bit0[i] >> 6) * 2;
        v.hi[wi] = 0; v.lo[wi] = 0; v.val[wi] = 0; v.hi[wi + 1] = 0; v.lo[wi + 1] = 0; v.val[wi + 1] = 0;
    }
    for (size_t wi = (size_t)(total_bases >> 6) * 2; wi < v.plane_words; wi++) { v.hi[wi] = 0; v.lo[wi] = 0; v.val[wi] = 0; }
}

}  // namespace trew

extern "C" {

size_t trew_pack_bound(uint32_t n_reads, uint64_t total_bases) { return trew::batch_bytes(n_reads, total_bases); }

int trew_pack_reads(const char* buffer, const int32_t* locs, uint32_t n, void* dst, size_t dst_bytes, trew_batch* out) {
    if ((n && (!buffer || !locs)) || !dst || !out || ((uintptr_t)dst & 7) != 0) return TREW_ERR_ARG;
    trew::ChunkView cv{{buffer, nullptr}, {locs, nullptr}, {nullptr, nullptr}, 1u};
    uint64_t total = 0; uint32_t mx = 0;
    trew::chunk_stats(cv, 0, n, &total, &mx);
    if (total >= 0xffffffffULL) return TREW_ERR_ARG;
    if (trew::batch_bytes(n, total) > dst_bytes) return TREW_ERR_ARG;
    trew::BatchView v;
    trew::batch_layout(dst, n, total, &v);
    uint64_t zero = 0;
    trew::pack_prepare(&zero, 1, total, v);
    uint64_t side[1][3];
    trew::pack_chunk_range(cv, 0, n, 0, 0, v, side[0], nullptr, false, total);
    trew::pack_fixup(&zero, side, 1, v);
    v.bit_off[n] = (uint32_t)total;
    out->n_reads = n; out->max_read_len = mx; out->bit_off = v.bit_off; out->hi = v.hi; out->lo = v.lo; out->val = v.val;
    return TREW_OK;
}

int trew_pack_reads_ranges(const char* buffer, const int32_t* locs, const char* buffer2, const int32_t* locs2, uint32_t n,
                           uint32_t n_ranges, uint32_t n_threads, uint32_t flags, void* dst, size_t dst_bytes, trew_batch* out,
                           uint32_t* inv, size_t inv_cap, size_t* n_inv) {
    if ((n && (!buffer || !locs)) || !dst || !out || ((uintptr_t)dst & 7) != 0 || (inv && !n_inv)) return TREW_ERR_ARG;
    if ((flags & TREW_PACK_NO_VAL) && !inv) return TREW_ERR_ARG;
    if ((buffer2 != nullptr) != (locs2 != nullptr)) return TREW_ERR_ARG;
    if (n_ranges == 0) n_ranges = 1;
    const uint32_t unit = buffer2 ? 2u : 1u;   // pairs: the batch holds 2 n reads, mate 1 and mate 2 alternating
    if (unit == 2 && n > 0x7fffffffu) return TREW_ERR_ARG;
    trew::ChunkView cv{{buffer, buffer2}, {locs, locs2}, {nullptr, nullptr}, unit};
    const uint32_t n_units = n;
    n *= unit;
    std::vector<uint32_t> r0((size_t)n_ranges + 1);
    std::vector<uint64_t> bases(n_ranges), bit0(n_ranges);
    for (uint32_t i = 0; i <= n_ranges; i++) r0[i] = (uint32_t)((uint64_t)n_units * i / n_ranges) * unit;
    trew::Pool pool((int)(n_threads ? n_threads : 1u));
    std::vector<uint32_t> mxs(n_ranges);
    pool.run((int)n_ranges, [&](int i) { trew::chunk_stats(cv, r0[(size_t)i], r0[(size_t)i + 1], &bases[(size_t)i], &mxs[(size_t)i]); });
    uint64_t total = 0; uint32_t mx = 0;
    for (uint32_t i = 0; i < n_ranges; i++) { bit0[i] = total; total += bases[i]; if (mxs[i] > mx) mx = mxs[i]; }
    if (total >= 0xffffffffULL) return TREW_ERR_ARG;
    if (trew::batch_bytes(n, total) > dst_bytes) return TREW_ERR_ARG;
    trew::BatchView v;
    trew::batch_layout(dst, n, total, &v);
    trew::pack_prepare(bit0.data(), (int)n_ranges, total, v);
    v.bit_off[n] = (uint32_t)total;
    std::vector<uint64_t> side_flat((size_t)3 * n_ranges);
    uint64_t (*side)[3] = (uint64_t (*)[3])side_flat.data();
    std::vector<trew::InvList> lists(inv ? n_ranges : 0u);
    pool.run((int)n_ranges, [&](int i) {
        trew::pack_chunk_range(cv, r0[(size_t)i], r0[(size_t)i + 1], r0[(size_t)i], bit0[(size_t)i], v, side[i],
                               inv ? &lists[(size_t)i] : nullptr, (flags & TREW_PACK_NO_VAL) != 0, bases[(size_t)i]);
    });
    trew::pack_fixup(bit0.data(), side, (int)n_ranges, v);
    if (inv) {
        size_t k = 0;
        for (const auto& l : lists)
            for (size_t j = 0; j < l.size(); j++) {
                uint32_t base; uint64_t z;
                l.get(j, &base, &z);
                for (; z; z &= z - 1) { if (k < inv_cap) inv[k] = base + (uint32_t)__builtin_ctzll(z); k++; }
            }
        *n_inv = k;
    }
    out->n_reads = n; out->max_read_len = mx; out->bit_off = v.bit_off; out->hi = v.hi; out->lo = v.lo; out->val = v.val;
    return TREW_OK;
}

}  // extern "C"
