// Huffman decoding tables of the DEFLATE decoders (inflate.cpp: the sequential byte decoder; pinflate.cpp: the
// speculative 16-bit decoder of the parallel gzip path).  Internal header.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>

namespace trew {
namespace {

constexpr int kLitBits = 11, kDistBits = 8, kPreBits = 7;
constexpr uint32_t kLiteral = 0x8000u, kExceptional = 0x4000u, kSubtable = 0x2000u, kEndOfBlock = 0x1000u;
constexpr uint32_t kInvalid = kExceptional;

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kPreOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

inline uint32_t reverse_bits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; i++) { r = (r << 1) | (code & 1u); code >>= 1; }
    return r;
}

enum Kind { kLitLen, kDist, kPre };

// Table entry (32 bits): bits 0-7 the input bits the symbol uses in this table -- code bits plus, for lengths and
// distances, the extra bits, so that one shift consumes the symbol --, 8-12 the number of extra bits (or the index
// width of a second-level table), 12-15 flags, 16-31 literal / base value / second-level table start.

inline uint32_t leaf(Kind kind, int sym, int bits) {
    if (kind == kLitLen) {
        if (sym < 256) return kLiteral | ((uint32_t)sym << 16) | (uint32_t)bits;
        if (sym == 256) return kExceptional | kEndOfBlock | (uint32_t)bits;
        if (sym < 286) return ((uint32_t)kLenBase[sym - 257] << 16) | ((uint32_t)kLenExtra[sym - 257] << 8) | (uint32_t)(bits + kLenExtra[sym - 257]);
        return kInvalid;
    }
    if (kind == kDist) {
        if (sym < 30) return ((uint32_t)kDistBase[sym] << 16) | ((uint32_t)kDistExtra[sym] << 8) | (uint32_t)(bits + kDistExtra[sym]);
        return kInvalid;
    }
    return ((uint32_t)sym << 16) | (uint32_t)bits;
}

// Canonical Huffman code -> lookup table indexed by the next table_bits input bits (LSB first).  Codes longer than
// table_bits go through a second-level table whose size fits the longest code with that prefix.  Over-subscribed
// length sets are rejected; unused code space decodes to kInvalid.
bool build_table(const uint8_t* lens, int n, Kind kind, int table_bits, uint32_t* table, size_t cap, size_t* used_out = nullptr) {
    int count[16] = {0};
    for (int i = 0; i < n; i++) count[lens[i]]++;
    count[0] = 0;
    int left = 1;
    for (int l = 1; l <= 15; l++) { left = (left << 1) - count[l]; if (left < 0) return false; }
    const size_t main_size = (size_t)1 << table_bits;
    for (size_t i = 0; i < main_size; i++) table[i] = kInvalid;
    uint32_t next[16];
    uint32_t code = 0;
    for (int l = 1; l <= 15; l++) { code = (code + (uint32_t)count[l - 1]) << 1; next[l] = code; }
    // pass 1 (long codes only): the longest code behind every first-level prefix
    uint8_t longest[1 << kLitBits];
    bool any_long = false;
    for (int l = table_bits + 1; l <= 15; l++) any_long |= count[l] != 0;
    uint32_t codes[320];
    if (any_long) memset(longest, 0, main_size);
    for (int s = 0; s < n; s++) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t r = reverse_bits(next[l]++, l);
        codes[s] = r;
        if (l <= table_bits) {
            const uint32_t e = leaf(kind, s, l);
            for (size_t i = r; i < main_size; i += (size_t)1 << l) table[i] = e;
        } else {
            uint8_t& m = longest[r & (main_size - 1)];
            if (l > m) m = (uint8_t)l;
        }
    }
    if (used_out) *used_out = main_size;
    if (!any_long) return true;
    // pass 2: allocate the second-level tables and fill them
    size_t used = main_size;
    for (int s = 0; s < n; s++) {
        const int l = lens[s];
        if (l <= table_bits) continue;
        const uint32_t r = codes[s];
        const size_t prefix = r & (main_size - 1);
        uint32_t e = table[prefix];
        if (!(e & kSubtable)) {
            const int sub_bits = longest[prefix] - table_bits;
            if (used + ((size_t)1 << sub_bits) > cap) return false;
            e = kExceptional | kSubtable | ((uint32_t)used << 16) | ((uint32_t)sub_bits << 8) | (uint32_t)table_bits;
            table[prefix] = e;
            for (size_t i = 0; i < ((size_t)1 << sub_bits); i++) table[used + i] = kInvalid;
            used += (size_t)1 << sub_bits;
        }
        const size_t start = e >> 16;
        const int sub_bits = (int)((e >> 8) & 0x1Fu);
        const uint32_t le = leaf(kind, s, l - table_bits);
        for (size_t i = r >> table_bits; i < ((size_t)1 << sub_bits); i += (size_t)1 << (l - table_bits)) table[start + i] = le;
    }
    if (used_out) *used_out = used;
    return true;
}


}  // namespace
}  // namespace trew
