// Parallel decode of ONE DEFLATE stream (the body of a plain gzip member) -- the stream the reference reads through a
// single gzread (FileReader, src/kmer.h:157-204) and that pins a .gz input to one core.
//
// DEFLATE has no index, but it is a sequence of blocks, and a dynamic-Huffman block header is so constrained (code
// length sets that must be complete prefix codes) that a block START can be recognised at an arbitrary bit position.
// So the compressed bytes ahead are cut into segments; the workers
//   1. search each segment for the first position that parses as a non-final dynamic block followed by another valid
//      block header (`find_block`),
//   2. decode each segment from its position until they land EXACTLY on a later segment's position at a block
//      boundary.  What a segment copies from the 32 KiB before its start is not known yet, so it decodes into 16-bit
//      symbols: 0..255 a byte, 0x8000 | i "byte i of the window before my start" (copies of such symbols stay symbols),
//   3. the chain of segments that really follow each other is walked from the first one (whose position and window are
//      known): each segment's last 32 KiB are resolved against its predecessor's -- serial, 32 K look-ups per segment --
//   4. and all segments are translated to bytes in parallel.
// A position the search got wrong is harmless: nobody lands on it, its segment is dropped, and the predecessor simply
// decodes on to the next position it does hit.  Everything is checked the way the sequential path checks it (the gzip
// layer in ingest.cpp compares CRC-32 and ISIZE of every member), and the result is the same bytes.  The idea is
// the one of pugz (Kerbiriou & Chikhi 2019); this is an independent implementation on this library's tables.
#include "host_internal.h"
#include "deflate_tables.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace trew {

namespace {

constexpr size_t kWin = 32768;
constexpr uint16_t kMark = 0x8000u;
constexpr size_t kLitCap32 = 4096, kDistCap32 = 1024;
constexpr size_t kMaxSegmentSymbols = (size_t)32 << 20;   // a segment that has produced this much stops at its next block boundary

inline uint64_t load_le64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

// bit reader over an in-memory stream; reads beyond the end deliver zeros (and are noticed: `over`)
struct Bits {
    const uint8_t* in; size_t n;       // the stream and its length in bytes
    uint64_t pos;                      // next bit
    bool over = false;
    Bits(const uint8_t* in_, size_t n_, uint64_t pos_) : in(in_), n(n_), pos(pos_) {}
    inline uint64_t peek() {           // >= 56 valid bits from pos (zeros beyond the end)
        const size_t b = (size_t)(pos >> 3);
        uint64_t v;
        if (b + 8 <= n) v = load_le64(in + b);
        else { v = 0; for (size_t i = b; i < n; i++) v |= (uint64_t)in[i] << (8 * (i - b)); }
        return v >> (pos & 7);
    }
    inline void skip(unsigned k) { pos += k; }
    inline uint32_t take(unsigned k) { const uint32_t v = (uint32_t)(peek() & (((uint64_t)1 << k) - 1)); pos += k; return v; }
    inline bool past_end() const { return pos > (uint64_t)n * 8; }
};

struct Tables {
    uint32_t lit[kLitCap32];
    uint32_t dist[kDistCap32];
};

// strict: what the block search demands beyond what a decoder must accept (complete code sets), to keep chance hits rare
bool complete(const uint8_t* lens, int n, bool allow_single) {
    int count[16] = {0};
    for (int i = 0; i < n; i++) count[lens[i]]++;
    int used = 0;
    for (int l = 1; l <= 15; l++) used += count[l];
    if (used == 0) return allow_single;
    if (used == 1) return allow_single && count[1] == 1;
    long left = 1;
    for (int l = 1; l <= 15; l++) { left = (left << 1) - count[l]; if (left < 0) return false; }
    return left == 0;
}

// reads a block header at br.pos.  type: 0 stored (br is left at the first data byte, *stored_len set), 1 / 2 Huffman
// (tables built).  false: not a valid header.
bool read_block_header(Bits& br, bool strict, bool* final_block, int* type, uint32_t* stored_len, Tables& t) {
    const uint64_t w = br.peek();
    *final_block = (w & 1u) != 0;
    *type = (int)((w >> 1) & 3u);
    br.skip(3);
    if (*type == 3) return false;
    if (*type == 0) {
        br.skip((unsigned)((8 - (br.pos & 7)) & 7));
        const uint32_t len = br.take(16), nlen = br.take(16);
        if ((len ^ nlen) != 0xFFFFu || br.past_end()) return false;
        *stored_len = len;
        return true;
    }
    uint8_t lens[320 + 140];
    int hlit, hdist;
    if (*type == 1) {
        int i = 0;
        for (; i < 144; i++) lens[i] = 8;
        for (; i < 256; i++) lens[i] = 9;
        for (; i < 280; i++) lens[i] = 7;
        for (; i < 288; i++) lens[i] = 8;
        for (i = 0; i < 32; i++) lens[288 + i] = 5;
        hlit = 288; hdist = 32;
    } else {
        hlit = (int)br.take(5) + 257; hdist = (int)br.take(5) + 1;
        const int hclen = (int)br.take(4) + 4;
        if (hlit > 286 || hdist > 30) return false;
        uint8_t pre_lens[19] = {0};
        for (int i = 0; i < hclen; i++) pre_lens[kPreOrder[i]] = (uint8_t)br.take(3);
        if (strict && !complete(pre_lens, 19, false)) return false;
        uint32_t pre[1 << kPreBits];
        if (!build_table(pre_lens, 19, kPre, kPreBits, pre, (size_t)1 << kPreBits)) return false;
        int i = 0;
        while (i < hlit + hdist) {
            const uint32_t e = pre[br.peek() & ((1u << kPreBits) - 1)];
            if (e & kExceptional) return false;
            br.skip(e & 0xFFu);
            const int sym = (int)(e >> 16);
            if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
            int rep;
            uint8_t val = 0;
            if (sym == 16) { if (i == 0) return false; val = lens[i - 1]; rep = 3 + (int)br.take(2); }
            else if (sym == 17) rep = 3 + (int)br.take(3);
            else rep = 11 + (int)br.take(7);
            if (i + rep > hlit + hdist) return false;
            memset(lens + i, val, (size_t)rep);
            i += rep;
        }
        if (br.past_end() || lens[256] == 0) return false;
        if (strict && (!complete(lens, hlit, false) || !complete(lens + hlit, hdist, true))) return false;
    }
    if (!build_table(lens, hlit, kLitLen, kLitBits, t.lit, kLitCap32)) return false;
    if (!build_table(lens + hlit, hdist, kDist, kDistBits, t.dist, kDistCap32)) return false;
    return true;
}

// growing array of 16-bit symbols
struct Sym {
    uint16_t* d = nullptr; size_t n = 0, cap = 0;
    ~Sym() { free(d); }
    bool reserve(size_t want) {
        if (want <= cap) return true;
        size_t c = cap ? cap : ((size_t)1 << 20);
        while (c < want) c += c / 2;
        uint16_t* p = (uint16_t*)realloc(d, c * sizeof(uint16_t));
        if (!p) return false;
        d = p; cap = c;
        return true;
    }
};

// One Huffman block from br into out (nullptr: only walk it, the block search).  limit: give up after this many symbols of
// output (the search).  Returns 1 end of block, 0 bad data, -1 limit / out of memory.
// Fast path (while 16 bytes of input remain): one unaligned 8-byte load gives >= 57 bits, enough for several literals in a
// row or for one length + distance pair (15 + 5 + 15 + 13 bits), so the loop reloads once per match or per few literals.
int huffman_block(Bits& br, const Tables& t, Sym* out, size_t limit) {
    const uint32_t lit_mask = (1u << kLitBits) - 1, dist_mask = (1u << kDistBits) - 1;
    const uint8_t* const in = br.in;
    const uint64_t fast_end = br.n >= 16 ? (uint64_t)(br.n - 16) * 8 : 0;
    uint64_t pos = br.pos;
    size_t produced = 0;
    int rc = 2;
    while (rc == 2) {
        if (out && out->cap - out->n < 600 && !out->reserve(out->n + 4096)) { rc = -1; break; }
        if (pos >= fast_end) {   // near the end of the input: one symbol at a time through the checked reader
            br.pos = pos;
            if (br.past_end()) { rc = 0; break; }
            uint64_t w = br.peek();
            uint32_t e = t.lit[w & lit_mask];
            unsigned used = 0;
            if (e & kExceptional) {
                if (e & kSubtable) {
                    used = kLitBits;
                    e = t.lit[(e >> 16) + ((w >> kLitBits) & ((1u << ((e >> 8) & 0x1Fu)) - 1))];
                }
                if (e & kExceptional) {
                    if (e & kEndOfBlock) { pos += used + (e & 0xFFu); rc = 1; break; }
                    rc = 0; break;
                }
            }
            if (e & kLiteral) {
                pos += used + (e & 0xFFu);
                if (out) out->d[out->n++] = (uint16_t)((e >> 16) & 0xFFu);
                if (++produced > limit) rc = -1;
                continue;
            }
            const unsigned lbits = e & 0xFFu, lx = (e >> 8) & 0x1Fu;
            const uint32_t len = (e >> 16) + (uint32_t)(((w >> used) >> (lbits - lx)) & (((uint64_t)1 << lx) - 1));
            pos += used + lbits;
            br.pos = pos;
            w = br.peek();
            uint32_t d = t.dist[w & dist_mask];
            used = 0;
            if (d & kExceptional) {
                if (!(d & kSubtable)) { rc = 0; break; }
                used = kDistBits;
                d = t.dist[(d >> 16) + ((w >> kDistBits) & ((1u << ((d >> 8) & 0x1Fu)) - 1))];
                if (d & kExceptional) { rc = 0; break; }
            }
            const unsigned dbits = d & 0xFFu, dx = (d >> 8) & 0x1Fu;
            const size_t dist = (d >> 16) + (size_t)(((w >> used) >> (dbits - dx)) & (((uint64_t)1 << dx) - 1));
            pos += used + dbits;
            if (out) {
                uint16_t* o = out->d + out->n;
                const size_t have = out->n;
                for (uint32_t i = 0; i < len; i++) {
                    const size_t at = have + i;
                    o[i] = at >= dist ? out->d[at - dist] : (uint16_t)(kMark | (uint16_t)(kWin - (dist - at)));
                }
                out->n += len;
            }
            produced += len;
            if (produced > limit) rc = -1;
            continue;
        }
        // ---- fast path
        uint64_t w = load_le64(in + (pos >> 3)) >> (pos & 7);   // >= 57 valid bits
        unsigned used = 0;
        uint32_t e = t.lit[w & lit_mask];
        uint16_t* o = out ? out->d + out->n : nullptr;
        unsigned nl = 0;
        while ((e & kLiteral) && used <= 42) {   // 57 - 15: the next code still fits
            if (out) o[nl] = (uint16_t)((e >> 16) & 0xFFu);
            nl++;
            used += e & 0xFFu;
            w >>= e & 0xFFu;
            e = t.lit[w & lit_mask];
        }
        if (nl) {
            if (out) out->n += nl;
            produced += nl;
            pos += used;
            if (produced > limit) rc = -1;
            continue;   // reload: whatever follows gets a full buffer
        }
        if (e & kExceptional) {
            if (e & kSubtable) {
                used = kLitBits;
                e = t.lit[(e >> 16) + ((w >> kLitBits) & ((1u << ((e >> 8) & 0x1Fu)) - 1))];
            }
            if (e & kExceptional) {
                if (e & kEndOfBlock) { pos += used + (e & 0xFFu); rc = 1; break; }
                rc = 0; break;
            }
            if (e & kLiteral) {
                pos += used + (e & 0xFFu);
                if (out) out->d[out->n++] = (uint16_t)((e >> 16) & 0xFFu);
                if (++produced > limit) rc = -1;
                continue;
            }
        }
        // length + distance from the same buffer
        const unsigned lbits = e & 0xFFu, lx = (e >> 8) & 0x1Fu;
        w >>= used;
        const uint32_t len = (e >> 16) + (uint32_t)((w >> (lbits - lx)) & (((uint64_t)1 << lx) - 1));
        w >>= lbits;
        used += lbits;
        uint32_t d = t.dist[w & dist_mask];
        if (d & kExceptional) {
            if (!(d & kSubtable)) { rc = 0; break; }
            w >>= kDistBits; used += kDistBits;
            d = t.dist[(d >> 16) + (w & ((1u << ((d >> 8) & 0x1Fu)) - 1))];
            if (d & kExceptional) { rc = 0; break; }
        }
        const unsigned dbits = d & 0xFFu, dx = (d >> 8) & 0x1Fu;
        const size_t dist = (d >> 16) + (size_t)((w >> (dbits - dx)) & (((uint64_t)1 << dx) - 1));
        pos += used + dbits;
        if (out) {
            const size_t have = out->n;
            if (dist <= have) {
                const uint16_t* s = o - dist;
                if (dist >= 4) {   // four symbols per copy; may write up to three symbols beyond len (room is kept)
                    for (uint32_t i = 0; i < len; i += 4) memcpy(o + i, s + i, 8);
                } else {
                    for (uint32_t i = 0; i < len; i++) o[i] = s[i];   // forward: a symbol may be read right after it was written
                }
            } else {
                for (uint32_t i = 0; i < len; i++) {
                    const size_t at = have + i;                        // position of the symbol; its source is at - dist
                    o[i] = at >= dist ? out->d[at - dist] : (uint16_t)(kMark | (uint16_t)(kWin - (dist - at)));
                }
            }
            out->n += len;
        }
        produced += len;
        if (produced > limit) rc = -1;
    }
    br.pos = pos;
    return rc;
}

// Cheap necessary conditions for "a non-final dynamic block starts at bit pos", straight from the bits: BFINAL = 0,
// BTYPE = 10, HLIT <= 29, HDIST <= 29 and a code-length code that is a complete prefix code (Kraft sum exactly 1 over
// the HCLEN + 4 three-bit lengths).  One position in a few hundred passes; only those are parsed for real.
inline bool maybe_block_start(const uint8_t* in, size_t n, uint64_t pos) {
    if ((pos >> 3) + 24 > n) return true;   // near the end of the input: let the checked parser decide
    const uint64_t w = load_le64(in + (pos >> 3)) >> (pos & 7);
    if ((w & 7u) != 4u) return false;
    if (((w >> 3) & 31u) > 29u || ((w >> 8) & 31u) > 29u) return false;
    const int hclen = (int)((w >> 13) & 15u) + 4;
    const uint64_t p2 = pos + 17;
    uint64_t v = load_le64(in + (p2 >> 3)) >> (p2 & 7);   // 57 bits = 19 fields
    int kraft = 0;
    for (int i = 0; i < hclen; i++) { const unsigned l = (unsigned)(v & 7u); v >>= 3; kraft += l ? 128 >> l : 0; }
    return kraft == 128;
}

// Does a non-final dynamic block start at bit `pos`?  It must parse strictly, decode to its end, and be followed by
// another valid block header.
bool block_starts_at(const uint8_t* in, size_t n, uint64_t pos, Tables& t) {
    Bits br(in, n, pos);
    const uint64_t w = br.peek();
    if ((w & 7u) != 4u) return false;                       // BFINAL = 0, BTYPE = 10
    if (((w >> 3) & 31u) > 29u || ((w >> 8) & 31u) > 29u) return false;   // HLIT, HDIST
    bool fin; int type; uint32_t sl;
    if (!read_block_header(br, true, &fin, &type, &sl, t)) return false;
    if (huffman_block(br, t, nullptr, (size_t)1 << 22) != 1) return false;
    if (br.past_end()) return false;
    Tables* t2 = new Tables();
    const bool ok = read_block_header(br, false, &fin, &type, &sl, *t2) && !br.past_end();
    delete t2;
    return ok;
}

}  // namespace

struct ParallelInflate::Segment {
    uint64_t start = 0, end = 0;     // bit positions
    bool found = false;              // a block start was found (segment 0: given)
    int stop = 0;                    // 1 landed on a later segment's start / the end of the span, 2 the final block ended,
                                     // 3 stopped early at a block boundary (output bound), 0 error
    bool eof = false;                // ... because the input ended
    Sym out;
    size_t out_off = 0;              // where its bytes go
};

ParallelInflate::ParallelInflate() {
    // 512 KiB of compressed input per segment: the symbols of one round (16 segments x ~2 M symbols x 2 bytes) stay in the
    // last-level cache between decode() and emit(), and the search for a block start is a small part of a segment's time.
    // Measured on the 16-core hosts, 2 M reads, gzip -6 / -1 (before the table translation and the search prefilter):
    // 256 KiB 1.73 / 1.21 Gbases/s, 512 KiB 2.24 / 1.28, 1 MiB 1.86 / 1.16, 2 MiB 1.61 / 1.04; now 2.92 / 2.31 at 512 KiB
    // (one stream through the sequential decoder: 0.70 / 0.58)
    seg_bytes_ = (size_t)512 << 10;
    if (const char* e = getenv("TREW_PGZ_SEGMENT")) { const long v = atol(e); if (v >= 1024) seg_bytes_ = (size_t)v; }
}
ParallelInflate::~ParallelInflate() { for (Segment* s : seg_) delete s; }

void ParallelInflate::start(const uint8_t* in, size_t n, size_t first_byte) {
    in_ = in; n_ = n; pos_ = (uint64_t)first_byte * 8; window_.clear();
}

bool ParallelInflate::next(Pool* pool, RawBytes& out, bool* member_end, size_t* next_byte, const char** err) {
    size_t total = 0;
    if (!decode(pool, &total, member_end, next_byte, err)) return false;
    const size_t base = out.n;
    if (!out.reserve(base + total)) { *err = "out of memory"; return false; }
    out.n = base + total;
    emit(pool, out.d + base, 0, total);
    return true;
}

bool ParallelInflate::decode(Pool* pool, size_t* total_out, bool* member_end, size_t* next_byte, const char** err) {
    *member_end = false;
    *total_out = 0;
    static const bool trace = getenv("TREW_PGZ_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = trace ? now() : 0;
    const int P = std::max(1, pool ? pool->size() : 1);
    const uint64_t span_end = std::min<uint64_t>((uint64_t)n_ * 8, (pos_ & ~(uint64_t)7) + (uint64_t)P * seg_bytes_ * 8);
    while (seg_.size() < (size_t)P) seg_.push_back(new Segment());
    struct SegView { std::vector<Segment*>& v; Segment& operator[](size_t i) { return *v[i]; } } seg{seg_};
    for (int i = 0; i < P; i++) { Segment& s = seg[(size_t)i]; s.start = s.end = 0; s.found = false; s.stop = 0; s.eof = false; s.out.n = 0; s.out_off = 0; }
    seg[0].start = pos_; seg[0].found = true;
    // 1. block starts
    auto search = [&](int i) {
        if (i == 0) return;
        const uint64_t a = (pos_ & ~(uint64_t)7) + (uint64_t)i * seg_bytes_ * 8, b = std::min(span_end, a + (uint64_t)seg_bytes_ * 8);
        Tables* t = new Tables();
        for (uint64_t p = a; p < b; p++) {
            if (maybe_block_start(in_, n_, p) && block_starts_at(in_, n_, p, *t)) { seg[(size_t)i].start = p; seg[(size_t)i].found = true; break; }
        }
        delete t;
    };
    if (pool && P > 1) pool->run(P, search);
    const double t1 = trace ? now() : 0;
    // 2. decode every segment until it lands on a later one (or reaches the end of the span / of the stream)
    std::vector<uint64_t> starts;
    for (int i = 1; i < P; i++) if (seg[(size_t)i].found) starts.push_back(seg[(size_t)i].start);
    auto decode = [&](int i) {
        Segment& s = seg[(size_t)i];
        if (!s.found) return;
        Tables* t = new Tables();
        Bits br(in_, n_, s.start);
        s.out.reserve((size_t)seg_bytes_ * 4);
        s.stop = 0;
        for (bool first = true;; first = false) {
            if (!first) {
                if (br.pos >= span_end || std::binary_search(starts.begin(), starts.end(), br.pos)) { s.stop = 1; break; }
                // very compressible data (runs of one base, of N): the round ends here and the next one goes on from this
                // boundary, so that the symbols of a round stay bounded (16 segments x 64 MB) whatever the ratio
                if (s.out.n > kMaxSegmentSymbols) { s.stop = 3; break; }
            }
            bool fin; int type; uint32_t sl = 0;
            if (!read_block_header(br, false, &fin, &type, &sl, *t)) break;
            if (type == 0) {
                const size_t b = (size_t)(br.pos >> 3);
                if (b + sl > n_ || !s.out.reserve(s.out.n + sl)) break;
                for (uint32_t k = 0; k < sl; k++) s.out.d[s.out.n + k] = in_[b + k];
                s.out.n += sl;
                br.skip(8 * sl);
            } else if (huffman_block(br, *t, &s.out, ~(size_t)0) != 1) break;
            if (fin) { s.stop = 2; break; }
        }
        s.end = br.pos;
        s.eof = s.stop == 0 && br.pos >= (uint64_t)n_ * 8;
        delete t;
    };
    if (pool && P > 1) pool->run(P, decode); else decode(0);
    const double t2 = trace ? now() : 0;
    // 3. the chain, and every link's last 32 KiB as bytes
    std::vector<int>& chain = chain_;
    std::vector<std::vector<uint8_t>>& wins = wins_;   // wins[c]: the translation table of chain[c] (from the window BEFORE it)
    chain.clear(); wins.clear();
    {
        int cur = 0;
        std::vector<uint8_t> w = window_;
        size_t total = 0;
        for (;;) {
            Segment& s = seg[(size_t)cur];
            if (s.stop == 0) { *err = s.eof ? "unexpected end of file" : "invalid compressed data"; return false; }
            // references into the window must not reach before the start of the member (a full window: every index is fine)
            if (w.size() < kWin) {
                for (size_t k = 0; k < s.out.n; k++) {
                    const uint16_t v = s.out.d[k];
                    if (v >= kMark && (size_t)(v & 0x7FFFu) < kWin - w.size()) { *err = "invalid compressed data"; return false; }
                }
            }
            chain.push_back(cur);
            // the link's translation table: symbol -> byte (0..255 themselves, 0x8000 | i byte i of the window before it).
            // In FASTQ most symbols ARE marks (28 % at gzip -6, 95 % at -1 on synthetic reads: headers and bases keep
            // being copied from copies), so emit() is one table load per symbol, not a test
            wins.emplace_back((size_t)65536);
            {
                std::vector<uint8_t>& lut = wins.back();
                for (int v = 0; v < 256; v++) lut[(size_t)v] = (uint8_t)v;
                if (!w.empty()) memcpy(lut.data() + 0x8000 + (kWin - w.size()), w.data(), w.size());
            }
            s.out_off = total;
            total += s.out.n;
            // the window after this segment
            std::vector<uint8_t> nw;
            const size_t take = std::min(kWin, s.out.n), keep = std::min(w.size(), kWin - take);
            nw.reserve(keep + take);
            nw.insert(nw.end(), w.end() - (ptrdiff_t)keep, w.end());
            for (size_t k = s.out.n - take; k < s.out.n; k++) {
                const uint16_t v = s.out.d[k];
                nw.push_back(v < kMark ? (uint8_t)v : w[w.size() - (kWin - (size_t)(v & 0x7FFFu))]);
            }
            w.swap(nw);
            if (s.stop == 2) { *member_end = true; break; }
            if (s.end >= span_end || s.stop == 3) break;
            int nxt = -1;
            for (int j = cur + 1; j < P; j++) if (seg[(size_t)j].found && seg[(size_t)j].start == s.end) { nxt = j; break; }
            if (nxt < 0) { *err = "invalid compressed data"; return false; }   // cannot happen: it stopped there because of j
            cur = nxt;
        }
        window_.swap(w);
        stat_calls++; stat_chained += chain.size();
        for (int i = 0; i < P; i++) stat_found += seg[(size_t)i].found ? 1 : 0;
        const Segment& last = seg[(size_t)chain.back()];
        pos_ = last.end;
        if (*member_end) *next_byte = (size_t)((pos_ + 7) >> 3);
        *total_out = total;
        if (trace) fprintf(stderr, "[pgz] %zu bytes out, %zu of %d segments: search %.1f ms, decode %.1f, chain %.1f\n", total, chain.size(), P,
                           t1 - t0, t2 - t1, now() - t2);
    }
    return true;
}

// bytes [off, off + len) of what the last decode() produced: symbols -> bytes.  The range is cut into equal parts, one per
// thread, whatever links of the chain they fall into (a caller's block often ends in the middle of a round: by link, the
// few links of such a tail would be all the parallelism).
void ParallelInflate::emit(Pool* pool, uint8_t* dst, size_t off, size_t len) {
    auto narrow = [&](size_t lo, size_t hi) {   // bytes [lo, hi) of the round
        for (size_t c = 0; c < chain_.size() && lo < hi; c++) {
            const Segment& s = *seg_[(size_t)chain_[c]];
            const size_t a = std::max(lo, s.out_off), b = std::min(hi, s.out_off + s.out.n);
            if (a >= b) continue;
            const uint8_t* lut = wins_[c].data();
            uint8_t* o = dst + (a - off);
            const uint16_t* d = s.out.d + (a - s.out_off);
            const size_t n = b - a;
            for (size_t k = 0; k < n; k++) o[k] = lut[d[k]];
        }
    };
    const int P = pool && len >= ((size_t)1 << 20) ? pool->size() : 1;
    if (P > 1) pool->run(P, [&](int i) { narrow(off + len * (size_t)i / (size_t)P, off + len * (size_t)(i + 1) / (size_t)P); });
    else narrow(off, off + len);
}

}  // namespace trew
