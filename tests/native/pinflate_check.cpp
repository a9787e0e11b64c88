// Test driver for trew_b200/csrc/pinflate.cpp (parallel decode of one DEFLATE stream): FASTQ-like and other data
// compressed by zlib at several levels / strategies / with flush points must come back byte for byte, whatever the
// segment size and thread count; corrupted and truncated streams must fail (or decode) without touching memory outside
// the buffers (build with -fsanitize=address,undefined).  zlib is the oracle.
#include "../../trew_b200/csrc/host_internal.h"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

typedef std::vector<uint8_t> Bytes;

static Bytes deflate_raw(const Bytes& src, int level, int strategy, std::mt19937_64& rng, bool flushes) {
    z_stream zs{};
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, strategy) != Z_OK) abort();
    Bytes out(deflateBound(&zs, src.size()) + 64 + (flushes ? src.size() / 50 * 16 + 1024 : 0));
    zs.next_out = out.data(); zs.avail_out = (uInt)out.size();
    size_t pos = 0;
    while (pos < src.size() && flushes) {
        size_t n = std::min<size_t>(src.size() - pos, 1 + rng() % 50000);
        zs.next_in = const_cast<uint8_t*>(src.data()) + pos; zs.avail_in = (uInt)n;
        const int fl[3] = {Z_SYNC_FLUSH, Z_FULL_FLUSH, Z_NO_FLUSH};
        if (deflate(&zs, fl[rng() % 3]) != Z_OK) abort();
        pos += n;
    }
    zs.next_in = const_cast<uint8_t*>(src.data()) + pos; zs.avail_in = (uInt)(src.size() - pos);
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) abort();
    out.resize(zs.total_out);
    deflateEnd(&zs);
    return out;
}

static Bytes fastq_like(std::mt19937_64& rng, size_t n_reads) {
    Bytes b;
    char hdr[96];
    for (size_t r = 0; r < n_reads; r++) {
        int hl = snprintf(hdr, sizeof(hdr), "@SIM:1:FCX:%d:%d:%d:%d 1:N:0:ATCACG\n", (int)(r % 8) + 1, (int)(1100 + r / 5000), (int)(rng() % 20000), (int)(rng() % 20000));
        b.insert(b.end(), hdr, hdr + hl);
        const size_t len = 100 + rng() % 60;
        const bool tel = rng() % 50 == 0;
        for (size_t i = 0; i < len; i++) b.push_back(tel ? "TTAGGG"[i % 6] : (rng() % 1000 == 0 ? 'N' : "ACGT"[rng() & 3]));
        b.push_back('\n'); b.push_back('+'); b.push_back('\n');
        for (size_t i = 0; i < len; i++) b.push_back((uint8_t)("FFFFF:FF,F#"[rng() % 11]));
        b.push_back('\n');
    }
    return b;
}

// decode comp (raw DEFLATE followed by `trailing` junk bytes) with the parallel decoder; true iff it reports the end of the
// stream at the right byte and the bytes equal `want`
static int decode(const Bytes& comp, size_t trailing, const Bytes& want, trew::Pool* pool, bool expect_ok) {
    Bytes in = comp;
    in.resize(comp.size() + trailing, 0x5A);
    trew::ParallelInflate pz;
    pz.start(in.data(), in.size(), 0);
    trew::RawBytes out;
    bool end = false;
    size_t next = 0;
    const char* err = nullptr;
    int rounds = 0;
    while (!end) {
        if (!pz.next(pool, out, &end, &next, &err)) return expect_ok ? 1 : 0;
        if (++rounds > 1000000) return 2;
    }
    if (!expect_ok) return 0;   // corrupted input that still decodes to something: fine, the CRC catches it
    if (next != comp.size()) return 3;
    if (out.n != want.size() || (out.n && memcmp(out.d, want.data(), out.n) != 0)) return 4;
    return 0;
}

int main(int argc, char** argv) {
    const uint64_t seed = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
    const int threads = argc > 2 ? atoi(argv[2]) : 4;
    std::mt19937_64 rng(seed);
    trew::Pool pool(threads);
    int cases = 0;
    for (int round = 0; round < 6; round++) {
        Bytes src;
        switch (round) {
            case 0: src = fastq_like(rng, 6000); break;
            case 1: src = fastq_like(rng, 1500); break;
            case 2: src.resize(150000); for (auto& c : src) c = (uint8_t)rng(); break;                      // incompressible: stored blocks
            case 3: src.assign(200000, 'A'); break;                                                          // one long run
            case 4: src = fastq_like(rng, 50); break;                                                        // smaller than one segment
            default: src.clear(); break;                                                                     // empty stream
        }
        const int levels[4] = {1, 6, 9, 4};
        const int strategies[3] = {Z_DEFAULT_STRATEGY, Z_FIXED, Z_HUFFMAN_ONLY};
        for (int li = 0; li < 4; li++) {
            for (int si = 0; si < 3; si++) {
                if (si && li > 1) continue;
                for (int fl = 0; fl < 2; fl++) {
                    const Bytes comp = deflate_raw(src, levels[li], strategies[si], rng, fl != 0);
                    int rc = decode(comp, 8 + rng() % 64, src, &pool, true);
                    if (rc) { printf("FAIL round %d level %d strategy %d flush %d: rc %d\n", round, levels[li], si, fl, rc); return 1; }
                    rc = decode(comp, 0, src, nullptr, true);   // no pool: one segment at a time
                    if (rc) { printf("FAIL (serial) round %d level %d strategy %d flush %d: rc %d\n", round, levels[li], si, fl, rc); return 1; }
                    cases += 2;
                    // corruption: flipped bits / truncation must not crash
                    for (int c = 0; c < 3 && comp.size() > 16; c++) {
                        Bytes bad = comp;
                        if (c < 2) bad[rng() % bad.size()] ^= (uint8_t)(1u << (rng() % 8));
                        else bad.resize(rng() % bad.size());
                        if (decode(bad, c == 2 ? 0 : 16, src, &pool, false) != 0) { printf("FAIL corrupt\n"); return 1; }
                        cases++;
                    }
                }
            }
        }
    }
    {   // very compressible: a segment reaches its output bound and the round ends at a block boundary
        Bytes src((size_t)80 << 20, 'N');
        for (size_t i = 0; i < src.size(); i += 4096) src[i] = (uint8_t)"ACGT"[(i >> 12) & 3];
        const Bytes comp = deflate_raw(src, 6, Z_DEFAULT_STRATEGY, rng, false);
        const int rc = decode(comp, 8, src, &pool, true);
        if (rc) { printf("FAIL compressible: rc %d\n", rc); return 1; }
        cases++;
    }
    printf("ok %d cases\n", cases);
    return 0;
}
