// Raw DEFLATE (RFC 1951) decoder for the FASTQ.gz ingest path.
//
// The reference reads .gz input through zlib's gzread (FileReader, src/kmer.h:157-204); with the scan on the GPU a
// gzip file is bound by single-stream inflate, so this path gets a decoder built for speed on a 64-bit machine:
// a 64-bit bit buffer refilled with one unaligned load, an 11-bit first-level table for literals/lengths (8-bit
// for distances) with second-level tables for longer codes, table entries that carry the base value and the
// extra-bit count so a symbol costs one lookup, first-level literal entries that decode up to four literals at once
// (sequence lines are ~2-bit codes: the lookup -> shift -> lookup latency chain is the limit, so each link should
// carry as many symbols as fit in 11 bits), up to three such lookups per refill, and word-wise match copies.
// It is resumable between symbols (input and output arrive in chunks; the last 32 KiB are kept as history), which
// the gzip stream layer in ingest.cpp relies on.  zlib stays in use for crc32 and as the fallback for inputs that
// are not gzip at all (gzread's transparent mode).
#include "host_internal.h"
#include "deflate_tables.h"

#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace trew {

namespace {

// Literal/length table as the decoder uses it: 64-bit entries.  Bits 0-15 as in the 32-bit leaf (bit count, extra
// bits, flags), 16-47 up to four literals (or the length base / subtable index in 16-31), 48-51 the code length of
// the first literal alone (the careful loop takes one symbol at a time), 56-58 the number of literals.  A first-
// level entry packs every further literal whose code still fits into the 11 index bits.
void finalize_litlen(const uint32_t* t32, size_t used, uint64_t* t64) {
    const uint32_t mask = (1u << kLitBits) - 1;
    for (uint32_t idx = 0; idx <= mask; idx++) {
        const uint32_t e = t32[idx];
        if (!(e & kLiteral)) { t64[idx] = e; continue; }
        const uint32_t l1 = e & 0xFFu;
        uint32_t bits = l1, cnt = 1;
        uint64_t lits = (e >> 16) & 0xFFu;
        while (cnt < 4 && bits < (uint32_t)kLitBits) {
            const uint32_t e2 = t32[(idx >> bits) & mask];
            if (!(e2 & kLiteral) || (e2 & 0xFFu) > (uint32_t)kLitBits - bits) break;
            lits |= (uint64_t)((e2 >> 16) & 0xFFu) << (8 * cnt);
            bits += e2 & 0xFFu;
            cnt++;
        }
        t64[idx] = (uint64_t)bits | kLiteral | (lits << 16) | ((uint64_t)l1 << 48) | ((uint64_t)cnt << 56);
    }
    for (size_t i = (size_t)mask + 1; i < used; i++) {
        const uint32_t e = t32[i];
        t64[i] = (e & kLiteral) ? ((uint64_t)e | ((uint64_t)(e & 0xFFu) << 48) | ((uint64_t)1 << 56)) : (uint64_t)e;
    }
}

inline uint64_t load_u64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

}  // namespace

void Inflater::reset() {
    bitbuf_ = 0; bitcnt_ = 0; state_ = kHeader; final_ = false; stored_left_ = 0; hist_len_ = 0;
}

void Inflater::resume(uint8_t first_byte, unsigned bit, const uint8_t* hist, size_t hist_len) {
    reset();
    bitbuf_ = (uint64_t)first_byte >> bit;
    bitcnt_ = 8 - (int)bit;
    if (hist_len > kWindow) { hist += hist_len - kWindow; hist_len = kWindow; }
    if (hist_len) memcpy(hist_, hist, hist_len);
    hist_len_ = hist_len;
}

size_t Inflater::leftover(uint8_t out[8]) {
    size_t n = 0;
    while (bitcnt_ >= 8) { out[n++] = (uint8_t)bitbuf_; bitbuf_ >>= 8; bitcnt_ -= 8; }
    bitbuf_ = 0; bitcnt_ = 0;
    return n;
}

// the last 32 KiB of everything produced so far (matches of the next call may reach back into it)
void Inflater::save_history(const uint8_t* out, size_t produced) {
    if (produced >= kWindow) {
        memcpy(hist_, out + produced - kWindow, kWindow);
        hist_len_ = kWindow;
    } else if (produced) {
        const size_t keep = hist_len_ + produced > kWindow ? kWindow - produced : hist_len_;
        memmove(hist_, hist_ + hist_len_ - keep, keep);
        memcpy(hist_ + keep, out, produced);
        hist_len_ = keep + produced;
    }
}

bool Inflater::read_header(const uint8_t*& ip, const uint8_t* in_end) {
    uint64_t bb = bitbuf_;
    int bc = bitcnt_;
    auto need = [&](int nbits) {
        while (bc < nbits) { if (ip == in_end) return false; bb |= (uint64_t)*ip++ << bc; bc += 8; }
        return true;
    };
    auto take = [&](int nbits) { const uint32_t v = (uint32_t)(bb & (((uint64_t)1 << nbits) - 1)); bb >>= nbits; bc -= nbits; return v; };
    bool ok = false;
    do {
        if (!need(3)) break;
        final_ = take(1) != 0;
        const uint32_t type = take(2);
        if (type == 0) {
            take(bc & 7);
            if (!need(32)) break;
            const uint32_t len = take(16), nlen = take(16);
            if ((len ^ nlen) != 0xFFFFu) break;
            stored_left_ = len;
            state_ = kStored;
            ok = true;
        } else if (type == 1) {
            uint8_t lens[320];
            int i = 0;
            for (; i < 144; i++) lens[i] = 8;
            for (; i < 256; i++) lens[i] = 9;
            for (; i < 280; i++) lens[i] = 7;
            for (; i < 288; i++) lens[i] = 8;
            uint32_t t32[kLitCap];
            size_t used = 0;
            if (!build_table(lens, 288, kLitLen, kLitBits, t32, kLitCap, &used)) break;
            finalize_litlen(t32, used, lit_);
            for (i = 0; i < 32; i++) lens[i] = 5;
            if (!build_table(lens, 32, kDist, kDistBits, dist_, kDistCap)) break;
            state_ = kHuffman;
            ok = true;
        } else if (type == 2) {
            if (!need(14)) break;
            const int hlit = (int)take(5) + 257, hdist = (int)take(5) + 1, hclen = (int)take(4) + 4;
            if (hlit > 286 || hdist > 30) break;
            uint8_t pre_lens[19] = {0};
            bool fine = true;
            for (int i = 0; i < hclen; i++) { if (!need(3)) { fine = false; break; } pre_lens[kPreOrder[i]] = (uint8_t)take(3); }
            if (!fine) break;
            uint32_t pre[1 << kPreBits];
            if (!build_table(pre_lens, 19, kPre, kPreBits, pre, (size_t)1 << kPreBits)) break;
            uint8_t lens[320 + 140];
            int i = 0;
            while (i < hlit + hdist) {
                if (!need(7 + 7)) {   // the very end of the input: fewer bits are fine as long as the code fits
                    if (bc <= 0) { fine = false; break; }
                }
                const uint32_t e = pre[bb & ((1u << kPreBits) - 1)];
                const int bits = (int)(e & 0xFFu);
                if ((e & kExceptional) || bits > bc) { fine = false; break; }
                take(bits);
                const int sym = (int)(e >> 16);
                if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
                int rep, extra;
                uint8_t val = 0;
                if (sym == 16) { if (i == 0) { fine = false; break; } val = lens[i - 1]; extra = 2; rep = 3; }
                else if (sym == 17) { extra = 3; rep = 3; }
                else { extra = 7; rep = 11; }
                if (bc < extra) { fine = false; break; }
                rep += (int)take(extra);
                if (i + rep > hlit + hdist) { fine = false; break; }
                memset(lens + i, val, (size_t)rep);
                i += rep;
            }
            if (!fine) break;
            if (lens[256] == 0) break;   // no end-of-block code
            uint32_t t32[kLitCap];
            size_t used = 0;
            if (!build_table(lens, hlit, kLitLen, kLitBits, t32, kLitCap, &used)) break;
            finalize_litlen(t32, used, lit_);
            if (!build_table(lens + hlit, hdist, kDist, kDistBits, dist_, kDistCap)) break;
            state_ = kHuffman;
            ok = true;
        }
    } while (false);
    bitbuf_ = bb; bitcnt_ = bc;
    return ok;
}

// One body, compiled twice: with BMI2 the variable shifts of the bit buffer are single three-operand instructions
// (about 10 % on the whole decode); run() picks at run time.
Inflater::Status Inflater::run(const uint8_t* in, size_t in_len, bool in_final, size_t* in_used, uint8_t* out, size_t out_cap,
                               size_t* out_used) {
#if defined(__x86_64__)
    static const bool bmi2 = __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("bmi") && !getenv("TREW_NO_BMI2");   // (tests run both builds)
    if (bmi2) return run_bmi2(in, in_len, in_final, in_used, out, out_cap, out_used);
#endif
    return run_generic(in, in_len, in_final, in_used, out, out_cap, out_used);
}

#if defined(__x86_64__)
__attribute__((target("bmi,bmi2"))) Inflater::Status Inflater::run_bmi2(const uint8_t* in, size_t in_len, bool in_final, size_t* in_used,
                                                                        uint8_t* out, size_t out_cap, size_t* out_used) {
    return run_impl(in, in_len, in_final, in_used, out, out_cap, out_used);
}
#endif

Inflater::Status Inflater::run_generic(const uint8_t* in, size_t in_len, bool in_final, size_t* in_used, uint8_t* out, size_t out_cap,
                                       size_t* out_used) {
    return run_impl(in, in_len, in_final, in_used, out, out_cap, out_used);
}

__attribute__((always_inline)) inline Inflater::Status Inflater::run_impl(const uint8_t* in, size_t in_len, bool in_final, size_t* in_used,
                                                                          uint8_t* out, size_t out_cap, size_t* out_used) {
    const uint8_t* ip = in;
    const uint8_t* const in_end = in + in_len;
    uint8_t* op = out;
    uint8_t* const out_end = out + out_cap;
    Status st = kError;
    for (;;) {
        if (state_ == kDone) { st = kStreamEnd; break; }
        if (state_ == kHeader) {
            // a dynamic block header is at most 562 bytes: with that much input in hand it can be parsed in one go
            if (!in_final && (size_t)(in_end - ip) < 1024) { st = kNeedInput; break; }
            if (!read_header(ip, in_end)) { st = kError; break; }
            continue;
        }
        if (state_ == kStored) {
            while (stored_left_ && bitcnt_ >= 8 && op < out_end) { *op++ = (uint8_t)bitbuf_; bitbuf_ >>= 8; bitcnt_ -= 8; stored_left_--; }
            if (stored_left_ && bitcnt_ < 8) {
                const size_t m = std::min<size_t>(std::min<size_t>(stored_left_, (size_t)(in_end - ip)), (size_t)(out_end - op));
                memcpy(op, ip, m);
                op += m; ip += m; stored_left_ -= (uint32_t)m;
            }
            if (stored_left_ == 0) { state_ = final_ ? kDone : kHeader; continue; }
            if (op == out_end) { st = kOutputFull; break; }
            st = in_final ? kError : kNeedInput;
            break;
        }
        // ---- Huffman block: fast loop while both buffers have slack
        uint64_t bb = bitbuf_;
        int bc = bitcnt_;
        bool block_done = false, bad = false;
        const uint32_t lit_mask = (1u << kLitBits) - 1, dist_mask = (1u << kDistBits) - 1;
        if (in_end - ip >= 48 && out_end - op >= 400) {
            // Per turn: up to three literal entries, or up to two and one match.  At most three refills (each loads 8
            // bytes and advances < 8) and 12 + 258 + 31 output bytes, hence the margins.  The entry for the next turn
            // is looked up before the match bytes are copied, so the copy overlaps the next decode.
            const uint8_t* const in_fast = in_end - 40;
            uint8_t* const out_fast = out_end - 340;
#define TREW_REFILL() do { bb |= load_u64(ip) << bc; ip += (63 - bc) >> 3; bc |= 56; } while (0)
#define TREW_LITS(e) do { const uint32_t lits_ = (uint32_t)((e) >> 16); memcpy(op, &lits_, 4); op += (e) >> 56; \
                          bb >>= ((e) & 0xFFu); bc -= (int)((e) & 0xFFu); } while (0)
            TREW_REFILL();
            uint64_t e = lit_[bb & lit_mask];
            for (;;) {
                if (e & kLiteral) {
                    TREW_LITS(e);
                    e = lit_[bb & lit_mask];
                    if (e & kLiteral) {
                        TREW_LITS(e);
                        e = lit_[bb & lit_mask];
                        if (e & kLiteral) {
                            TREW_LITS(e);
                            TREW_REFILL();
                            e = lit_[bb & lit_mask];
                            if (ip > in_fast || op > out_fast) break;
                            continue;
                        }
                    }
                    TREW_REFILL();
                }
                if (e & kExceptional) {
                    if (e & kSubtable) {
                        bb >>= kLitBits; bc -= kLitBits;
                        e = lit_[(uint32_t)(e >> 16) + (bb & ((1u << ((e >> 8) & 0x1Fu)) - 1))];
                        if (e & kLiteral) {
                            bb >>= (e & 0xFFu); bc -= (int)(e & 0xFFu); *op++ = (uint8_t)(e >> 16);
                            TREW_REFILL();
                            e = lit_[bb & lit_mask];
                            if (ip > in_fast || op > out_fast) break;
                            continue;
                        }
                    }
                    if (e & kExceptional) {
                        if (e & kEndOfBlock) { bb >>= (e & 0xFFu); bc -= (int)(e & 0xFFu); block_done = true; break; }
                        bad = true; break;
                    }
                }
                uint64_t sv = bb;   // the extra bits are cut out of the saved buffer, off the critical path
                bb >>= (e & 0xFFu); bc -= (int)(e & 0xFFu);
                const uint32_t lx = (uint32_t)((e >> 8) & 0x1Fu);
                const uint32_t len = (uint32_t)(e >> 16) + (uint32_t)((sv >> ((uint32_t)(e & 0xFFu) - lx)) & (((uint64_t)1 << lx) - 1));
                uint32_t d = dist_[bb & dist_mask];
                if (d & kExceptional) {
                    if (!(d & kSubtable)) { bad = true; break; }
                    bb >>= kDistBits; bc -= kDistBits;
                    d = dist_[(d >> 16) + (bb & ((1u << ((d >> 8) & 0x1Fu)) - 1))];
                    if (d & kExceptional) { bad = true; break; }
                }
                sv = bb;
                bb >>= (d & 0xFFu); bc -= (int)(d & 0xFFu);
                const uint32_t dx = (d >> 8) & 0x1Fu;
                const size_t dist = (d >> 16) + (size_t)((sv >> ((d & 0xFFu) - dx)) & (((uint64_t)1 << dx) - 1));
                TREW_REFILL();
                e = lit_[bb & lit_mask];
                const size_t produced = (size_t)(op - out);
                if (__builtin_expect(dist > produced, 0)) {
                    if (dist - produced > hist_len_) { bad = true; break; }
                    for (uint32_t i = 0; i < len; i++) {
                        const ptrdiff_t pos = (ptrdiff_t)produced + (ptrdiff_t)i - (ptrdiff_t)dist;
                        op[i] = pos < 0 ? hist_[(ptrdiff_t)hist_len_ + pos] : out[pos];
                    }
                    op += len;
                } else {
                    const uint8_t* src = op - dist;
                    uint8_t* const stop = op + len;
                    if (dist >= 8) {   // word copies in order: a word may read what the previous one wrote
                        memcpy(op, src, 8);
                        memcpy(op + 8, src + 8, 8);
                        memcpy(op + 16, src + 16, 8);
                        memcpy(op + 24, src + 24, 8);
                        if (len > 32) {
                            op += 32; src += 32;
                            do { memcpy(op, src, 8); op += 8; src += 8; } while (op < stop);
                        }
                    } else if (dist == 1) {
                        const uint64_t v = 0x0101010101010101ULL * (uint64_t)*src;
                        memcpy(op, &v, 8);
                        memcpy(op + 8, &v, 8);
                        memcpy(op + 16, &v, 8);
                        memcpy(op + 24, &v, 8);
                        if (len > 32) {
                            op += 32;
                            do { memcpy(op, &v, 8); op += 8; } while (op < stop);
                        }
                    } else {
                        do { *op++ = *src++; } while (op < stop);
                    }
                    op = stop;
                }
                if (ip > in_fast || op > out_fast) break;
            }
#undef TREW_REFILL
#undef TREW_LITS
            bb &= bc >= 64 ? ~(uint64_t)0 : (((uint64_t)1 << bc) - 1);   // drop the look-ahead bits above bitcnt
        }
        // ---- careful loop near the end of either buffer: one symbol at a time, committed only if it fits
        while (!block_done && !bad) {
            while (bc < 56 && ip < in_end) { bb |= (uint64_t)*ip++ << bc; bc += 8; }
            if (bc < 48 && !in_final) { st = kNeedInput; goto suspend; }
            if (ip + 48 <= in_end && op + 400 <= out_end) break;   // (after a refill by the caller) back to the fast loop
            uint64_t b2 = bb;
            int c2 = bc;
            uint64_t e = lit_[b2 & lit_mask];
            if ((e & kExceptional) && (e & kSubtable)) {
                if (c2 < kLitBits) { bad = true; break; }
                b2 >>= kLitBits; c2 -= kLitBits;
                e = lit_[(uint32_t)(e >> 16) + (b2 & ((1u << ((e >> 8) & 0x1Fu)) - 1))];
            }
            if ((e & kExceptional) && !(e & kEndOfBlock)) { bad = true; break; }
            if (e & kLiteral) {   // the first literal of the entry only
                const int l1 = (int)((e >> 48) & 0xFu);
                if (l1 > c2) { bad = true; break; }
                if (op == out_end) { st = kOutputFull; goto suspend; }
                *op++ = (uint8_t)(e >> 16);
                bb = b2 >> l1; bc = c2 - l1;
                continue;
            }
            if ((int)(e & 0xFFu) > c2) { bad = true; break; }   // input ends inside a symbol
            uint64_t sv2 = b2;
            b2 >>= (e & 0xFFu); c2 -= (int)(e & 0xFFu);
            if (e & kEndOfBlock) { bb = b2; bc = c2; block_done = true; break; }
            const uint32_t lx = (uint32_t)((e >> 8) & 0x1Fu);   // (the entry's bit count, consumed above, includes them)
            const uint32_t len = (uint32_t)(e >> 16) + (uint32_t)((sv2 >> ((uint32_t)(e & 0xFFu) - lx)) & (((uint64_t)1 << lx) - 1));
            uint32_t d = dist_[b2 & dist_mask];
            if ((d & kExceptional) && (d & kSubtable)) {
                if (c2 < kDistBits) { bad = true; break; }
                b2 >>= kDistBits; c2 -= kDistBits;
                d = dist_[(d >> 16) + (b2 & ((1u << ((d >> 8) & 0x1Fu)) - 1))];
            }
            if (d & kExceptional) { bad = true; break; }
            if ((int)(d & 0xFFu) > c2) { bad = true; break; }
            sv2 = b2;
            b2 >>= (d & 0xFFu); c2 -= (int)(d & 0xFFu);
            const uint32_t dx = (d >> 8) & 0x1Fu;
            const size_t dist = (d >> 16) + (size_t)((sv2 >> ((d & 0xFFu) - dx)) & (((uint64_t)1 << dx) - 1));
            if ((size_t)(out_end - op) < len) { st = kOutputFull; goto suspend; }
            const size_t produced = (size_t)(op - out);
            if (dist > produced && dist - produced > hist_len_) { bad = true; break; }
            for (uint32_t i = 0; i < len; i++) {
                const ptrdiff_t pos = (ptrdiff_t)produced + (ptrdiff_t)i - (ptrdiff_t)dist;
                op[i] = pos < 0 ? hist_[(ptrdiff_t)hist_len_ + pos] : out[pos];
            }
            op += len;
            bb = b2; bc = c2;
        }
        bitbuf_ = bb; bitcnt_ = bc;
        if (bad) { st = kError; break; }
        if (block_done) state_ = final_ ? kDone : kHeader;
        continue;
    suspend:
        bitbuf_ = bb; bitcnt_ = bc;
        break;
    }
    if (st == kStreamEnd) { bitbuf_ >>= (bitcnt_ & 7); bitcnt_ -= (bitcnt_ & 7); }
    const size_t produced = (size_t)(op - out);
    save_history(out, produced);
    *in_used = (size_t)(ip - in);
    *out_used = produced;
    return st;
}

}  // namespace trew
