// Test driver for trew_b200/csrc/inflate.cpp: every stream zlib's deflate can produce here (levels, strategies,
// flush points, stored / fixed / dynamic blocks) must decode to the original bytes -- in one piece, and with input
// and output arriving in arbitrary chunks -- and corrupted streams must fail or finish without touching memory
// outside the buffers (build with -fsanitize=address,undefined).  zlib is the oracle.
#include "../../trew_b200/csrc/host_internal.h"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

using trew::Inflater;
typedef std::vector<uint8_t> Bytes;

static Bytes deflate_raw(const Bytes& src, int level, int strategy, std::mt19937_64& rng, bool flushes) {
    z_stream zs{};
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, strategy) != Z_OK) abort();
    Bytes out(deflateBound(&zs, src.size()) + 64 + (flushes ? src.size() / 50 * 16 + 1024 : 0));
    zs.next_out = out.data(); zs.avail_out = (uInt)out.size();
    size_t pos = 0;
    while (pos < src.size() && flushes) {
        size_t n = std::min<size_t>(src.size() - pos, 1 + rng() % 5000);
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

// whole buffers: the BGZF use (input complete, output exactly as large as the content)
static bool decode_whole(const Bytes& comp, const Bytes& want, size_t trailing) {
    Inflater inf;
    Bytes in = comp;
    in.resize(comp.size() + trailing, 0xA5);
    Bytes out(want.size());
    size_t iu = 0, ou = 0;
    Inflater::Status st = inf.run(in.data(), in.size(), true, &iu, out.data(), out.size(), &ou);
    if (st != Inflater::kStreamEnd || ou != want.size() || out != want) return false;
    uint8_t lo[8];
    const size_t nl = inf.leftover(lo);
    // consumed bytes minus the ones handed back = the stream's length
    return iu - nl == comp.size();
}

// chunks: the gzip stream use -- a sliding input buffer that is topped up, output buffers of arbitrary sizes
static bool decode_chunked(const Bytes& comp, const Bytes& want, std::mt19937_64& rng) {
    Inflater inf;
    Bytes got;
    Bytes win;   // input window
    size_t fed = 0;
    bool final_in = false;
    const size_t top = 1024 + rng() % 3000;
    for (int guard = 0; guard < 10000000; guard++) {
        if (!final_in && win.size() < top) {
            size_t n = std::min<size_t>(comp.size() - fed, top + rng() % 4096);
            win.insert(win.end(), comp.begin() + (ptrdiff_t)fed, comp.begin() + (ptrdiff_t)(fed + n));
            fed += n;
            final_in = fed == comp.size();
        }
        Bytes out(1 + rng() % (rng() % 4 ? 70000 : 300));
        size_t iu = 0, ou = 0;
        Inflater::Status st = inf.run(win.data(), win.size(), final_in, &iu, out.data(), out.size(), &ou);
        got.insert(got.end(), out.begin(), out.begin() + (ptrdiff_t)ou);
        win.erase(win.begin(), win.begin() + (ptrdiff_t)iu);
        if (st == Inflater::kError) return false;
        if (st == Inflater::kStreamEnd) return got == want;
        if (st == Inflater::kNeedInput && final_in) return false;
        if (got.size() > want.size()) return false;
    }
    return false;
}

static Bytes make_data(int kind, size_t n, std::mt19937_64& rng) {
    Bytes d(n);
    switch (kind) {
    case 0: for (auto& b : d) b = (uint8_t)rng(); break;                                  // incompressible
    case 1: for (auto& b : d) b = "ACGT"[rng() & 3]; break;                                // 2 bits of entropy
    case 2: {                                                                              // FASTQ-like
        std::string s;
        size_t r = 0;
        while (s.size() < n) {
            s += "@read" + std::to_string(r++) + " 1:N:0\n";
            const int L = 100 + (int)(rng() % 60);
            for (int i = 0; i < L; i++) s += "ACGTN"[rng() % 100 == 0 ? 4 : rng() & 3];
            s += "\n+\n";
            for (int i = 0; i < L; i++) s += (char)('#' + (rng() % 8 ? 37 : (int)(rng() % 40)));
            s += "\n";
        }
        if (n) memcpy(d.data(), s.data(), n);
        break;
    }
    case 3: {                                                                              // long runs and short periods
        size_t i = 0;
        while (i < n) {
            const size_t run = 1 + rng() % 2000, period = 1 + rng() % 9;
            uint8_t pat[9];
            for (auto& p : pat) p = (uint8_t)rng();
            for (size_t j = 0; j < run && i < n; j++) d[i++] = pat[j % period];
        }
        break;
    }
    default: {                                                                             // skewed alphabet: long Huffman codes
        for (auto& b : d) { int k = 0; while (k < 250 && (rng() & 1)) k++; b = (uint8_t)k; }
        if (n > 300) for (int i = 0; i < 256; i++) d[rng() % n] = (uint8_t)i;
    }
    }
    return d;
}

int main(int argc, char** argv) {
    const uint64_t seed = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
    const int rounds = argc > 2 ? atoi(argv[2]) : 40;
    std::mt19937_64 rng(seed);
    size_t n_streams = 0, n_fuzz = 0;
    const size_t sizes[] = {0, 1, 2, 7, 64, 300, 5000, 70000, 400000};
    for (int round = 0; round < rounds; round++) {
        for (int kind = 0; kind < 5; kind++) {
            const size_t n = round < (int)(sizeof(sizes) / sizeof(sizes[0])) ? sizes[round] : rng() % 300000;
            const Bytes src = make_data(kind, n, rng);
            const int levels[] = {0, 1, 4, 6, 9};
            const int strategies[] = {Z_DEFAULT_STRATEGY, Z_FIXED, Z_HUFFMAN_ONLY, Z_RLE, Z_FILTERED};
            const int level = levels[rng() % 5], strategy = strategies[rng() % 5];
            const Bytes comp = deflate_raw(src, level, strategy, rng, rng() % 3 == 0);
            n_streams++;
            if (!decode_whole(comp, src, rng() % 2 ? 0 : 1 + rng() % 40)) { printf("FAIL whole kind %d n %zu level %d strategy %d\n", kind, n, level, strategy); return 1; }
            for (int rep = 0; rep < 2; rep++)
                if (!decode_chunked(comp, src, rng)) { printf("FAIL chunked kind %d n %zu level %d strategy %d\n", kind, n, level, strategy); return 1; }
            // corrupted copies: any outcome but a crash / out-of-bounds access
            for (int f = 0; f < 6 && !comp.empty(); f++) {
                Bytes bad = comp;
                const int hits = 1 + (int)(rng() % 3);
                for (int h = 0; h < hits; h++) bad[rng() % bad.size()] ^= (uint8_t)(1u << (rng() % 8));
                if (rng() % 4 == 0) bad.resize(rng() % bad.size());
                Inflater inf;
                Bytes out(src.size() + rng() % 1000);
                size_t iu = 0, ou = 0;
                inf.run(bad.data(), bad.size(), true, &iu, out.data(), out.size(), &ou);
                if (iu > bad.size() || ou > out.size()) { printf("FAIL fuzz bounds\n"); return 1; }
                decode_chunked(bad, src, rng);
                n_fuzz++;
            }
        }
    }
    printf("ok %zu streams %zu corrupted\n", n_streams, n_fuzz);
    return 0;
}
