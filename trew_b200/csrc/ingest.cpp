// FASTQ / FASTQ.gz ingest with the reference's record semantics (read_fastq_thread,
// read_pair_fastq_thread, read_fastq_long_thread: src/kmer.cpp:987-1213; FileReader: src/kmer.h:157-204).
//
//   * lines are delimited by '\n' only; the 2nd line of every group of four is a sequence line; its
//     length counts every byte up to the '\n' (a trailing '\r' is an invalid base, not stripped);
//   * a last line without '\n' is never seen (the reference only acts on newlines);
//   * short mode: a sequence line longer than MAX_SEQ = 1000 aborts with the reference's message
//     (src/kmer.cpp:1006-1008);  long mode: lines shorter than SLICE_LENGTH are dropped (:1184);
//   * paired mode: records are paired index-wise; different line totals at EOF are an error (:1111-1115).
// The chunk size is an I/O detail with no observable effect (the reference uses 4 MiB - 1).
#include "host_internal.h"

#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace trew {

void GrowBuf::reserve(size_t n, size_t keep) {
    if (n <= cap) return;
    size_t want = (n + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    char* p = (char*)aligned_alloc((size_t)2 << 20, want);
    if (!p) throw std::bad_alloc();
#ifdef MADV_HUGEPAGE
    madvise(p, want, MADV_HUGEPAGE);
#endif
    if (data && keep) memcpy(p, data, keep);
    free(data);
    data = p; cap = want;
}

GrowBuf::~GrowBuf() { free(data); }

namespace {

// uninitialised, grow-only array of uint32 (newline positions); untouched pages of a generous capacity cost nothing
struct RawU32 {
    uint32_t* d = nullptr;
    size_t cap = 0, n = 0;
    void ensure(size_t want) {
        if (want <= cap) return;
        uint32_t* q = (uint32_t*)malloc(want * sizeof(uint32_t));
        if (!q) throw std::bad_alloc();
        if (n) memcpy(q, d, n * sizeof(uint32_t));
        free(d);
        d = q; cap = want;
    }
    RawU32() = default;
    RawU32(const RawU32&) = delete;
    RawU32& operator=(const RawU32&) = delete;
    RawU32(RawU32&& o) noexcept : d(o.d), cap(o.cap), n(o.n) { o.d = nullptr; o.cap = o.n = 0; }
    ~RawU32() { free(d); }
};

// Offsets (relative to p) of the '\n' bytes in p[a, b), appended at o; returns the new end.  o must have room for
// b - a + 3 entries.  FASTQ has a newline every ~80 bytes, so a memchr call per line is mostly call overhead; here a
// 64-byte block becomes a bit mask and its first three set bits are stored without a branch (the cursor advances
// only past real ones).
inline uint32_t* nl_from_mask(uint32_t* o, uint32_t base, uint64_t m) {
    const uint64_t top = 1ULL << 63;
    *o = base + (uint32_t)__builtin_ctzll(m | top); o += (m != 0); m &= m - 1;
    *o = base + (uint32_t)__builtin_ctzll(m | top); o += (m != 0); m &= m - 1;
    *o = base + (uint32_t)__builtin_ctzll(m | top); o += (m != 0); m &= m - 1;
    while (__builtin_expect(m != 0, 0)) { *o++ = base + (uint32_t)__builtin_ctzll(m); m &= m - 1; }
    return o;
}

uint32_t* find_newlines_scalar(const char* p, size_t a, size_t b, uint32_t* o) {
    while (a < b) {
        const char* q = (const char*)memchr(p + a, '\n', b - a);
        if (!q) break;
        *o++ = (uint32_t)(q - p);
        a = (size_t)(q - p) + 1;
    }
    return o;
}

#if defined(__x86_64__)
__attribute__((target("avx512f,avx512bw"))) uint32_t* find_newlines_avx512(const char* p, size_t a, size_t b, uint32_t* o) {
    const __m512i nl = _mm512_set1_epi8('\n');
    size_t i = a;
    for (; i + 64 <= b; i += 64)
        o = nl_from_mask(o, (uint32_t)i, _mm512_cmpeq_epi8_mask(_mm512_loadu_si512((const void*)(p + i)), nl));
    if (i < b) {
        const __mmask64 keep = (__mmask64)((1ULL << (b - i)) - 1ULL);
        o = nl_from_mask(o, (uint32_t)i, _mm512_mask_cmpeq_epi8_mask(keep, _mm512_maskz_loadu_epi8(keep, (const void*)(p + i)), nl));
    }
    return o;
}

__attribute__((target("avx2"))) uint32_t* find_newlines_avx2(const char* p, size_t a, size_t b, uint32_t* o) {
    const __m256i nl = _mm256_set1_epi8('\n');
    size_t i = a;
    for (; i + 64 <= b; i += 64) {
        const uint32_t m0 = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i*)(p + i)), nl));
        const uint32_t m1 = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i*)(p + i + 32)), nl));
        o = nl_from_mask(o, (uint32_t)i, (uint64_t)m0 | ((uint64_t)m1 << 32));
    }
    return find_newlines_scalar(p, i, b, o);
}
#endif

uint32_t* find_newlines(const char* p, size_t a, size_t b, uint32_t* o) {
#if defined(__x86_64__)
    static const int lvl = [] {
        int l = 0;
        if (__builtin_cpu_supports("avx2")) l = 1;
        if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw")) l = 2;
        if (const char* e = getenv("TREW_PACK_SIMD")) { int cap = atoi(e); if (cap < l) l = cap < 0 ? 0 : cap; }
        return l;
    }();
    if (lvl == 2) return find_newlines_avx512(p, a, b, o);
    if (lvl == 1) return find_newlines_avx2(p, a, b, o);
#endif
    return find_newlines_scalar(p, a, b, o);
}

struct Reader {
    bool gz = false;
    int fd = -1;
    gzFile gfp = nullptr;
    uint64_t offset = 0;    // plain files: next byte to read
    int64_t size = -1;      // plain regular files: total size, else -1
    // Plain regular files are not read at all but mapped: the newline index and the packer work on the page cache
    // directly (3x the throughput of pread + scan on the hosts measured -- the copy was the cost).  As with any
    // mapping, a file that is truncated while it is being processed raises SIGBUS; TREW_NO_MMAP=1 reads instead.
    const char* map = nullptr;
    // BGZF (bgzip) files are a series of independent gzip members of at most 64 KiB, each announcing its compressed
    // size in a 'BC' extra field: the members of a batch are inflated in parallel.  Plain gzip stays on zlib's gzread.
    bool bgzf = false;
    const unsigned char* cmap = nullptr;   // the compressed file, mapped (the workers fault it in while inflating)
    std::vector<char> pending;         // inflated bytes not yet handed out (a block that did not fit the caller's buffer)
    size_t pending_pos = 0;
    // Ordinary gzip files (regular file, gzip magic) go through the decoder in inflate.cpp instead of zlib's gzread:
    // same stream semantics -- concatenated members, trailing garbage after a member ignored -- checked against
    // each member's CRC-32 and ISIZE.  (Stricter than gzread in one point: a file that ends inside a member is an
    // I/O error here; gzread hands out what it could decode and then reports end of file.)
    bool ownz = false;
    std::vector<unsigned char> zbuf;   // window of the compressed file
    size_t z_pos = 0, z_have = 0;
    bool z_eof = false, z_any_member = false;
    enum ZPhase { kGzHeader, kGzBody, kGzTrailer, kGzEnd } zphase = kGzHeader;
    std::unique_ptr<Inflater> inf;
    uint64_t z_len = 0;                // bytes of the current member so far
    const char* zerr = nullptr;
    // CRC-32 work is deferred to finish_crc (it runs on the pool): output pieces in order, a member's last piece
    // carrying the CRC its trailer announced
    struct CrcPiece { const char* p; size_t len; bool member_end; uint32_t want; };
    std::vector<CrcPiece> crc_todo;
    uint32_t z_crc = 0;
    // With a pool of at least four threads the members' DEFLATE streams are decoded by all of them (pinflate.cpp:
    // speculative block starts in the mapped compressed file); same stream semantics and checks as read_gz.
    bool pgz = false, pgz_tried = false;
    ParallelInflate pz;
    size_t pz_pos = 0;            // next member header
    bool pz_in_member = false, pz_trailer_due = false, pz_done = false;
    size_t pz_total = 0, pz_emitted = 0, pz_next_byte = 0;   // the part decode() holds, what was handed out of it, the byte behind the stream
    // A stream whose blocks are not found (stored / fixed blocks only, blocks larger than a segment) leaves one thread
    // decoding 16-bit symbols: slower than the byte decoder.  Two such rounds in a row hand the rest of the member to it.
    int pz_poor_rounds = 0;
    bool pz_seq = false;          // the rest of this member goes through `inf`
    size_t pz_seq_pos = 0;        // its next input byte
    // Files made of many small members (a member smaller than a segment is one round for one thread, and the others search
    // in vain): after two such members in a row the members are decoded sequentially from their start, until one turns out
    // to be large again.
    int pz_small_members = 0;
    bool pz_members_seq = false;
    size_t pz_member_start = 0;

    static bool bgzf_header(const unsigned char* p, size_t avail, uint32_t* bsize, uint32_t* hdr_len) {
        if (avail < 18 || p[0] != 31 || p[1] != 139 || p[2] != 8 || !(p[3] & 4)) return false;
        const uint32_t xlen = p[10] | ((uint32_t)p[11] << 8);
        if (avail < 12 + xlen) return false;
        for (uint32_t q = 12; q + 4 <= 12 + xlen;) {
            const uint32_t slen = p[q + 2] | ((uint32_t)p[q + 3] << 8);
            if (p[q] == 'B' && p[q + 1] == 'C' && slen == 2 && q + 6 <= 12 + xlen) {
                *bsize = (p[q + 4] | ((uint32_t)p[q + 5] << 8)) + 1u;
                *hdr_len = 12 + xlen;
                return *bsize >= *hdr_len + 8;
            }
            q += 4 + slen;
        }
        return false;
    }

    bool open(const char* name, bool is_gz) {
        gz = is_gz;
        if (gz) {
            int probe = ::open(name, O_RDONLY);
            if (probe >= 0) {
                unsigned char h[64];
                ssize_t got = pread(probe, h, sizeof(h), 0);
                struct stat st;
                uint32_t bs, hl;
                if (got >= 18 && bgzf_header(h, (size_t)got, &bs, &hl) && fstat(probe, &st) == 0 && S_ISREG(st.st_mode) &&
                    !getenv("TREW_NO_BGZF")) {
                    void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_SHARED, probe, 0);
                    if (m != MAP_FAILED) {
                        bgzf = true; fd = probe; size = (int64_t)st.st_size; cmap = (const unsigned char*)m;
                        return true;
                    }   // cannot map: the file is still a valid multi-member gzip for the sequential reader below
                }
                if (got >= 2 && h[0] == 31 && h[1] == 139 && fstat(probe, &st) == 0 && S_ISREG(st.st_mode) && !getenv("TREW_ZLIB_GZ")) {
                    ownz = true; fd = probe; size = (int64_t)st.st_size;
                    zbuf.resize(((size_t)8 << 20) + 64);
                    inf.reset(new Inflater());
                    return true;
                }
                ::close(probe);
            }
            gfp = gzopen(name, "r");
            if (gfp) gzbuffer(gfp, 1 << 20);
            return gfp != nullptr;
        }
        fd = ::open(name, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode)) size = (int64_t)st.st_size;
        if (size > 0 && !getenv("TREW_NO_MMAP")) {
            void* m = mmap(nullptr, (size_t)size, PROT_READ, MAP_SHARED, fd, 0);
            if (m != MAP_FAILED) map = (const char*)m;
        }
        return true;
    }

    struct BgzfBlock { const unsigned char* cdata; uint32_t clen, isize; size_t out_off; };

    // every member's CRC-32 (the trailer word right behind its DEFLATE data) is checked like gzread does: a flipped
    // payload bit is "incorrect data check", not silently different records
    static bool inflate_blocks(const BgzfBlock* blk, size_t n, char* out) {
        std::unique_ptr<Inflater> inf(new Inflater());
        for (size_t i = 0; i < n; i++) {
            const unsigned char* t = blk[i].cdata + blk[i].clen;
            const uint32_t want = t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
            if (blk[i].isize == 0) { if (want != 0) return false; continue; }
            inf->reset();
            size_t iu = 0, ou = 0;
            const Inflater::Status st = inf->run(blk[i].cdata, blk[i].clen, true, &iu, (uint8_t*)out + blk[i].out_off, blk[i].isize, &ou);
            if (st != Inflater::kStreamEnd || ou != blk[i].isize) return false;
            if ((uint32_t)crc32_z(0L, (const Bytef*)out + blk[i].out_off, blk[i].isize) != want) return false;
        }
        return true;
    }

    long read_bgzf(char* buf, size_t n, Pool* pool) {
        size_t done = 0;
        if (pending_pos < pending.size()) {
            size_t m = std::min(n, pending.size() - pending_pos);
            memcpy(buf, pending.data() + pending_pos, m);
            pending_pos += m; done = m;
            if (pending_pos == pending.size()) { pending.clear(); pending_pos = 0; }
        }
        if (done == n || offset >= (uint64_t)size) return (long)done;
        // plan the blocks that fit the caller's buffer
        std::vector<BgzfBlock> plan;
        uint64_t off = offset;
        size_t out = done;
        bool spill = false;
        while (off < (uint64_t)size) {
            const unsigned char* p = cmap + off;
            const size_t avail = (size_t)((uint64_t)size - off);
            uint32_t bs = 0, hl = 0;
            if (!bgzf_header(p, avail, &bs, &hl) || bs > avail) return -1;   // malformed or truncated
            const uint32_t isize = p[bs - 4] | ((uint32_t)p[bs - 3] << 8) | ((uint32_t)p[bs - 2] << 16) | ((uint32_t)p[bs - 1] << 24);
            if (isize > 65536) return -1;
            if (out + isize > n) { spill = plan.empty(); break; }
            plan.push_back(BgzfBlock{p + hl, bs - hl - 8, isize, out});
            out += isize; off += bs;
        }
        if (!plan.empty()) {
            const int P = pool ? std::min<int>(pool->size() * 4, (int)(plan.size() / 8) + 1) : 1;
            std::vector<char> okv((size_t)P, 1);
            auto work = [&](int i) {
                size_t a = plan.size() * (size_t)i / (size_t)P, b = plan.size() * (size_t)(i + 1) / (size_t)P;
                okv[(size_t)i] = inflate_blocks(plan.data() + a, b - a, buf) ? 1 : 0;
            };
            if (P > 1) pool->run(P, work); else work(0);
            for (char o : okv) if (!o) return -1;
            done = out; offset = off;
        } else if (spill) {
            // the next block alone is larger than what is left of the caller's buffer: inflate it aside
            const unsigned char* p = cmap + offset;
            uint32_t bs = 0, hl = 0;
            if (!bgzf_header(p, (size_t)((uint64_t)size - offset), &bs, &hl)) return -1;
            const uint32_t isize = p[bs - 4] | ((uint32_t)p[bs - 3] << 8) | ((uint32_t)p[bs - 2] << 16) | ((uint32_t)p[bs - 1] << 24);
            pending.resize(isize); pending_pos = 0;
            BgzfBlock one{p + hl, bs - hl - 8, isize, 0};
            if (!inflate_blocks(&one, 1, pending.data())) return -1;
            offset += bs;
            size_t m = std::min(n - done, pending.size());
            memcpy(buf + done, pending.data(), m);
            pending_pos = m; done += m;
            if (pending_pos == pending.size()) { pending.clear(); pending_pos = 0; }
        }
        return (long)done;
    }

    // refill the compressed window; keeps the 8 bytes before z_pos (the decoder may hand bytes back at a member's end)
    bool z_topup() {
        if (z_eof) return true;
        const size_t keep_from = z_pos >= 8 ? z_pos - 8 : 0;
        if (keep_from) { memmove(zbuf.data(), zbuf.data() + keep_from, z_have - keep_from); z_have -= keep_from; z_pos -= keep_from; }
        while (z_have < zbuf.size()) {
            ssize_t r = pread(fd, zbuf.data() + z_have, zbuf.size() - z_have, (off_t)offset);
            if (r < 0) { if (errno == EINTR) continue; zerr = strerror(errno); return false; }
            if (r == 0) { z_eof = true; break; }
            z_have += (size_t)r; offset += (uint64_t)r;
        }
        return true;
    }

    // gzip member header (RFC 1952) at p: its length, 0 if incomplete within avail, -1 if malformed
    static long gz_header_len(const unsigned char* p, size_t avail) {
        if (avail < 10) return 0;
        if (p[0] != 31 || p[1] != 139 || p[2] != 8 || (p[3] & 0xE0)) return -1;
        const int flg = p[3];
        size_t q = 10;
        if (flg & 4) { if (avail < q + 2) return 0; q += 2 + (size_t)(p[q] | (p[q + 1] << 8)); if (avail < q) return 0; }
        for (int bit : {8, 16})
            if (flg & bit) { while (q < avail && p[q]) q++; if (q >= avail) return 0; q++; }
        if (flg & 2) q += 2;
        return avail < q ? 0 : (long)q;
    }

    long read_gz(char* buf, size_t n) {
        size_t done = 0;
        while (done < n) {
            if (pending_pos < pending.size()) {   // bytes decoded aside (below) go out first, in order
                const size_t m = std::min(n - done, pending.size() - pending_pos);
                memcpy(buf + done, pending.data() + pending_pos, m);
                crc_todo.push_back(CrcPiece{buf + done, m, false, 0});
                pending_pos += m; done += m;
                if (pending_pos == pending.size()) { pending.clear(); pending_pos = 0; }
                continue;
            }
            if (zphase == kGzEnd) break;
            if (!z_eof && z_have - z_pos < ((size_t)64 << 10) && !z_topup()) return -1;
            const unsigned char* p = zbuf.data() + z_pos;
            const size_t avail = z_have - z_pos;
            if (zphase == kGzHeader) {
                if (avail < 2 || p[0] != 31 || p[1] != 139) {
                    // end of file, or bytes that are not another member: ignored once a member was read (as gzread does)
                    if (avail == 0 || z_any_member) { zphase = kGzEnd; break; }
                    zerr = "not in gzip format"; return -1;
                }
                const long hl = gz_header_len(p, avail);
                if (hl < 0) { zerr = "unknown compression method or header flags"; return -1; }
                if (hl == 0) { zerr = z_eof ? "unexpected end of file" : "gzip header too large"; return -1; }
                z_pos += (size_t)hl;
                inf->reset();
                z_len = 0; z_any_member = true;
                zphase = kGzBody;
            } else if (zphase == kGzBody) {
                size_t iu = 0, ou = 0;
                Inflater::Status st;
                if (n - done < 1024) {
                    // too little room left for the decoder to be sure of progress (a match is up to 258 bytes): decode aside
                    pending.resize((size_t)64 << 10); pending_pos = 0;
                    st = inf->run(p, avail, z_eof, &iu, (uint8_t*)pending.data(), pending.size(), &ou);
                    pending.resize(ou);
                    z_len += ou;
                } else {
                    st = inf->run(p, avail, z_eof, &iu, (uint8_t*)buf + done, n - done, &ou);
                    if (ou) crc_todo.push_back(CrcPiece{buf + done, ou, false, 0});
                    done += ou; z_len += ou;
                }
                z_pos += iu;
                if (st == Inflater::kError) { zerr = z_eof && z_pos >= z_have ? "unexpected end of file" : "invalid compressed data"; return -1; }
                if (st == Inflater::kStreamEnd) {
                    uint8_t back[8];
                    z_pos -= inf->leftover(back);
                    zphase = kGzTrailer;
                } else if (st == Inflater::kOutputFull) {
                    if (pending.empty()) break;   // the caller's buffer is as full as it gets
                } else if (z_eof) {   // kNeedInput although the decoder was told the input is complete
                    zerr = "unexpected end of file"; return -1;
                }
            } else {   // kGzTrailer: CRC-32 and ISIZE, little endian
                if (avail < 8) {
                    if (z_eof) { zerr = "unexpected end of file"; return -1; }
                    if (!z_topup()) return -1;
                    continue;
                }
                const uint32_t crc = p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
                const uint32_t isize = p[4] | ((uint32_t)p[5] << 8) | ((uint32_t)p[6] << 16) | ((uint32_t)p[7] << 24);
                if (isize != (uint32_t)z_len) { zerr = "incorrect length check"; return -1; }
                crc_todo.push_back(CrcPiece{buf + done, 0, true, crc});
                z_pos += 8;
                zphase = kGzHeader;
            }
        }
        return (long)done;
    }

    void try_parallel_gz(Pool* pool) {
        if (!pool) return;   // (asked again when a pool comes along)
        pgz_tried = true;
        size_t min_bytes = (size_t)1 << 20;
        if (const char* e = getenv("TREW_PGZ_MIN_BYTES")) min_bytes = (size_t)std::max(0L, atol(e));
        if (!pool || pool->size() < 4 || size < (int64_t)min_bytes || offset != 0 || z_have != 0 || getenv("TREW_NO_PARALLEL_GZ")) return;
        void* m = mmap(nullptr, (size_t)size, PROT_READ, MAP_SHARED, fd, 0);
        if (m == MAP_FAILED) return;
        cmap = (const unsigned char*)m;
        pgz = true;
    }

    static uint32_t le32(const unsigned char* p) { return p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

    long read_pgz(char* buf, size_t n, Pool* pool) {
        size_t done = 0;
        while (done < n) {
            if (pending_pos < pending.size()) {   // bytes the sequential decoder produced aside (below) go out first
                const size_t m = std::min(n - done, pending.size() - pending_pos);
                memcpy(buf + done, pending.data() + pending_pos, m);
                crc_todo.push_back(CrcPiece{buf + done, m, false, 0});
                pending_pos += m; done += m;
                if (pending_pos == pending.size()) { pending.clear(); pending_pos = 0; }
                continue;
            }
            if (pz_emitted < pz_total) {   // straight into the caller's block; what does not fit waits in the decoder as symbols
                const size_t m = std::min(n - done, pz_total - pz_emitted);
                pz.emit(pool, (uint8_t*)buf + done, pz_emitted, m);
                crc_todo.push_back(CrcPiece{buf + done, m, false, 0});
                pz_emitted += m; done += m; z_len += m;
                continue;
            }
            if (pz_trailer_due) {   // CRC-32 and ISIZE, little endian
                if (pz_next_byte + 8 > (size_t)size) { zerr = "unexpected end of file"; return -1; }
                if (le32(cmap + pz_next_byte + 4) != (uint32_t)z_len) { zerr = "incorrect length check"; return -1; }
                crc_todo.push_back(CrcPiece{buf + done, 0, true, le32(cmap + pz_next_byte)});
                {
                    const size_t member_bytes = pz_next_byte - pz_member_start, seg = pz.segment_bytes();
                    if (member_bytes < seg) pz_small_members++; else pz_small_members = 0;
                    if (pz_small_members >= 2) pz_members_seq = true;
                    else if (member_bytes >= 16 * seg) pz_members_seq = false;
                }
                pz_pos = pz_next_byte + 8;
                pz_in_member = false; pz_trailer_due = false;
                continue;
            }
            if (pz_done) break;
            if (pz_seq) {
                size_t iu = 0, ou = 0;
                Inflater::Status st;
                if (n - done < 1024) {
                    // too little room left for the decoder to be sure of progress (a match is up to 258 bytes): decode aside
                    pending.resize((size_t)64 << 10); pending_pos = 0;
                    st = inf->run(cmap + pz_seq_pos, (size_t)size - pz_seq_pos, true, &iu, (uint8_t*)pending.data(), pending.size(), &ou);
                    pending.resize(ou);
                } else {
                    st = inf->run(cmap + pz_seq_pos, (size_t)size - pz_seq_pos, true, &iu, (uint8_t*)buf + done, n - done, &ou);
                    if (ou) crc_todo.push_back(CrcPiece{buf + done, ou, false, 0});
                    done += ou;
                }
                pz_seq_pos += iu;
                z_len += ou;
                if (st == Inflater::kError || st == Inflater::kNeedInput) {
                    zerr = pz_seq_pos >= (size_t)size ? "unexpected end of file" : "invalid compressed data";
                    return -1;
                }
                if (st == Inflater::kStreamEnd) {
                    uint8_t back[8];
                    pz_next_byte = pz_seq_pos - inf->leftover(back);
                    pz_trailer_due = true; pz_seq = false;
                    continue;
                }
                if (pending.empty()) break;   // kOutputFull: the caller's block is as full as it gets
                continue;
            }
            if (!pz_in_member) {
                const unsigned char* p = cmap + pz_pos;
                const size_t avail = (size_t)size - pz_pos;
                if (avail < 2 || p[0] != 31 || p[1] != 139) {
                    // end of file, or bytes that are not another member: ignored once a member was read (as gzread does)
                    if (avail == 0 || z_any_member) { pz_done = true; break; }
                    zerr = "not in gzip format"; return -1;
                }
                const long hl = gz_header_len(p, avail);
                if (hl < 0) { zerr = "unknown compression method or header flags"; return -1; }
                if (hl == 0) { zerr = "unexpected end of file"; return -1; }
                pz_member_start = pz_pos;
                pz_in_member = true; z_any_member = true; z_len = 0; pz_poor_rounds = 0;
                if (pz_members_seq) {
                    if (!inf) inf.reset(new Inflater());
                    inf->reset();
                    pz_seq = true; pz_seq_pos = pz_pos + (size_t)hl;
                    continue;
                }
                pz.start(cmap, (size_t)size, pz_pos + (size_t)hl);
            }
            const char* e = nullptr;
            bool mend = false;
            size_t nb = 0, total = 0;
            if (!pz.decode(pool, &total, &mend, &nb, &e)) { zerr = e ? e : "invalid compressed data"; return -1; }
            pz_total = total; pz_emitted = 0; pz_trailer_due = mend; pz_next_byte = nb;
            if (!mend && (!pool || pz.last_chain() == 1 || pz.last_chain() * 8 <= (size_t)pool->size())) {
                if (++pz_poor_rounds >= 2) {
                    const uint64_t bit = pz.position();
                    if ((bit >> 3) >= (uint64_t)size) { zerr = "unexpected end of file"; return -1; }
                    const std::vector<uint8_t>& w = pz.window();
                    if (!inf) inf.reset(new Inflater());
                    inf->resume(cmap[bit >> 3], (unsigned)(bit & 7), w.data(), w.size());
                    pz_seq = true; pz_seq_pos = (size_t)(bit >> 3) + 1;
                    if (getenv("TREW_PGZ_TRACE")) fprintf(stderr, "[pgz] block starts are not found: the rest of the member goes to the sequential decoder (byte %zu)\n", pz_seq_pos);
                }
            } else pz_poor_rounds = 0;
        }
        return (long)done;
    }

    // CRC-32 of what read_gz handed out since the last call (parallel slices combined with crc32_combine)
    bool finish_crc(Pool* pool) {
        bool ok = true;
        for (const CrcPiece& c : crc_todo) {
            if (c.len) {
                const int P = pool && c.len >= ((size_t)1 << 20) ? pool->size() : 1;
                std::vector<uint32_t> part((size_t)P);
                auto work = [&](int i) {
                    const size_t a = c.len * (size_t)i / (size_t)P, b = c.len * (size_t)(i + 1) / (size_t)P;
                    part[(size_t)i] = (uint32_t)crc32_z(0L, (const Bytef*)c.p + a, b - a);
                };
                if (P > 1) pool->run(P, work); else work(0);
                for (int i = 0; i < P; i++) {
                    const size_t a = c.len * (size_t)i / (size_t)P, b = c.len * (size_t)(i + 1) / (size_t)P;
                    z_crc = (uint32_t)crc32_combine(z_crc, part[(size_t)i], (z_off_t)(b - a));
                }
            }
            if (c.member_end) {
                if (z_crc != c.want) { zerr = "incorrect data check"; ok = false; }
                z_crc = 0;
            }
        }
        crc_todo.clear();
        return ok;
    }

    // returns bytes read (0 at EOF), -1 on error.  Plain regular files are read with pread in parallel slices when a
    // pool is given and the request is large (the copy out of the page cache is the cost, and it scales with cores).
    // `hook` (parallel plain-file reads only): begin(slice, bytes) once per slice, then data(slice, a, b) after every
    // ~1 MiB that landed in buf[a, b) -- the caller can look at the bytes while they are still in that core's cache;
    // *hook_slices = the number of slices (0: the hook was not used).
    struct Hook {
        std::function<void(int, size_t)> begin;
        std::function<void(int, size_t, size_t)> data;
    };
    long read(char* buf, size_t n, Pool* pool, size_t par_min, const Hook* hook = nullptr, int* hook_slices = nullptr) {
        if (hook_slices) *hook_slices = 0;
        if (bgzf) return read_bgzf(buf, std::min<size_t>(n, (size_t)1 << 30), pool);
        if (ownz && !pgz_tried) try_parallel_gz(pool);
        if (pgz) return read_pgz(buf, std::min<size_t>(n, (size_t)1 << 30), pool);
        if (ownz) return read_gz(buf, std::min<size_t>(n, (size_t)1 << 30));
        if (gz) {
            int r = gzread(gfp, buf, (unsigned)std::min<size_t>(n, 1u << 30));
            return r < 0 ? -1 : r;
        }
        if (size >= 0) {
            size_t want = (size_t)std::min<uint64_t>(n, (uint64_t)size > offset ? (uint64_t)size - offset : 0);
            if (want == 0) return 0;
            const int P = pool && want >= par_min ? std::min(pool->size(), (int)(want / (par_min / 4 + 1)) + 1) : 1;
            std::vector<long> got((size_t)P, 0);
            const bool hooked = hook && P > 1;
            auto slice = [&](int i) {
                size_t a = want * (size_t)i / (size_t)P, b = want * (size_t)(i + 1) / (size_t)P;
                if (hooked) hook->begin(i, b - a);
                while (a < b) {
                    const size_t step = hooked ? std::min(b - a, (size_t)1 << 20) : b - a;
                    ssize_t r = pread(fd, buf + a, step, (off_t)(offset + a));
                    if (r < 0) { if (errno == EINTR) continue; got[(size_t)i] = -1; return; }
                    if (r == 0) break;
                    if (hooked) hook->data(i, a, a + (size_t)r);
                    a += (size_t)r; got[(size_t)i] += (long)r;
                }
            };
            if (P > 1) pool->run(P, slice); else slice(0);
            if (hooked && hook_slices) *hook_slices = P;
            long total = 0;
            for (long g : got) { if (g < 0) return -1; total += g; }
            offset += (uint64_t)total;
            return total;
        }
        for (;;) {
            ssize_t r = ::read(fd, buf, n);
            if (r < 0 && errno == EINTR) continue;
            return r < 0 ? -1 : (long)r;
        }
    }
    std::string error() {
        if (bgzf) return "malformed BGZF block or incorrect data check";
        if (ownz) return zerr ? zerr : "gzip stream error";
        if (gz) { int e; return gzerror(gfp, &e); }
        return strerror(errno);
    }
    void close() {
        if (map) munmap(const_cast<char*>(map), (size_t)size);
        map = nullptr;
        if (cmap) munmap(const_cast<unsigned char*>(cmap), (size_t)size);
        cmap = nullptr;
        if (gfp) gzclose(gfp);
        gfp = nullptr;
        if (fd >= 0) ::close(fd);
        fd = -1;
    }
    ~Reader() { close(); }
};

// One side of the ingest: a buffer holding [carried bytes | fresh bytes] and the line phase.
struct Side {
    Reader rd;
    GrowBuf own;
    GrowBuf* buf = &own;   // the caller's scratch buffer when it provides one
    size_t have = 0;        // bytes in buf
    size_t scanned = 0;     // bytes already examined for newlines
    size_t line_start = 0;  // start of the line being assembled
    uint64_t num = 0;       // newlines seen so far (the reference's `num`)
    uint64_t total_lines = 0;
    bool eof = false;
    std::vector<int32_t> locs;  // sequence lines found in buf, inclusive (st, nd)

    // read more bytes; returns false on I/O error
    // With a pool, plain files are read in parallel slices and each slice's newlines are indexed right behind the
    // read, 1 MiB at a time, instead of in a second pass over memory (scan() then only assigns the line roles).
    size_t win = 0;   // mapped files: file offset of the window's first byte (data() == map + win)
    const char* data() const { return rd.map ? rd.map + win : buf->data; }
    bool fill(size_t chunk, Pool* pool, size_t par_min, bool defer_crc = false) {
        nl_slices = 0;
        if (chunk == 0) return true;   // paired mode: enough carried records to pair with, nothing to read this round
        if (rd.map) {   // widen the window over the mapping; nothing is copied
            const size_t left = (size_t)rd.size - win - have;
            const size_t add = std::min(chunk, left);
            if (add == 0) eof = true;
            have += add;
            if (win + have == (size_t)rd.size) eof = true;
            else posix_fadvise(rd.fd, (off_t)(win + have), (off_t)chunk, POSIX_FADV_WILLNEED);   // cold files: start reading the next window
            return true;
        }
        buf->reserve(have + chunk, have);
        Reader::Hook hook;
        const bool fuse = pool && pool->size() > 1 && scanned == have;
        if (fuse) {
            if (nl.size() < (size_t)pool->size()) nl.resize((size_t)pool->size());
            const char* base = buf->data;
            const size_t have0 = have;
            hook.begin = [this](int i, size_t bytes) { nl[(size_t)i].n = 0; nl[(size_t)i].ensure(bytes + 3); };
            hook.data = [this, base, have0](int i, size_t a, size_t b) {
                RawU32& v = nl[(size_t)i];
                v.n = (size_t)(find_newlines(base, have0 + a, have0 + b, v.d + v.n) - v.d);
            };
        }
        long r = rd.read(buf->data + have, chunk, pool, par_min, fuse ? &hook : nullptr, &nl_slices);
        if (r < 0) return false;
        if (!defer_crc && !rd.finish_crc(pool)) return false;
        if (r == 0) eof = true;
        have += (size_t)r;
        return true;
    }
    inline void line_done(size_t st, size_t nl, int mode, int slice, bool* too_long, std::vector<int32_t>& out) {
        size_t len = nl - st;
        if (mode == TREW_MODE_SHORT && len > 1000) *too_long = true;
        if (!(mode == TREW_MODE_LONG && len < (size_t)slice)) {
            out.push_back((int32_t)st);
            out.push_back((int32_t)nl - 1);
        }
    }
    std::vector<RawU32> nl;   // per slice: newline offsets found by the parallel search
    int nl_slices = 0;        // > 0: fill() already indexed the fresh bytes into nl[0 .. nl_slices)
    // Turn the newline offsets of slices [0, P) into sequence-line locations.  A line's role depends only on the
    // ordinal of the newline that ends it, so a slice needs nothing from the others but their newline counts; except
    // in long mode (which drops short lines) the number of sequence lines per slice follows from those counts too and
    // every slice writes straight into its place in `locs`.
    void finish_scan(int P, int mode, int slice, bool* too_long, Pool* pool) {
        std::vector<uint64_t> base((size_t)P + 1), first((size_t)P + 1);
        std::vector<int64_t> prev((size_t)P);
        base[0] = num;
        int64_t last = (int64_t)line_start - 1;
        const size_t locs0 = locs.size();
        first[0] = locs0;
        for (int i = 0; i < P; i++) {
            const uint64_t g0 = base[(size_t)i], g1 = g0 + nl[(size_t)i].n;
            base[(size_t)i + 1] = g1;
            first[(size_t)i + 1] = first[(size_t)i] + 2 * (((g1 + 2) >> 2) - ((g0 + 2) >> 2));   // ordinals g in (g0, g1] with g % 4 == 2
            prev[(size_t)i] = last;
            if (nl[(size_t)i].n) last = (int64_t)nl[(size_t)i].d[nl[(size_t)i].n - 1];
        }
        std::vector<char> tl((size_t)P, 0);
        if (mode != TREW_MODE_LONG) {
            locs.resize((size_t)first[(size_t)P]);
            int32_t* out = locs.data();
            auto work = [&](int i) {
                const RawU32& v = nl[(size_t)i];
                int32_t* o = out + first[(size_t)i];
                const uint64_t g0 = base[(size_t)i];
                bool t = false;
                // the first newline of the slice with ordinal % 4 == 2, then every fourth
                for (size_t j = (size_t)((2 - (g0 + 1)) & 3); j < v.n; j += 4) {
                    const int64_t st = (j ? (int64_t)v.d[j - 1] : prev[(size_t)i]) + 1;
                    const int64_t e = (int64_t)v.d[j];
                    t |= mode == TREW_MODE_SHORT && e - st > 1000;
                    *o++ = (int32_t)st; *o++ = (int32_t)e - 1;
                }
                tl[(size_t)i] = t ? 1 : 0;
            };
            if (pool && P > 1) pool->run(P, work); else for (int i = 0; i < P; i++) work(i);
        } else {
            std::vector<std::vector<int32_t>> out((size_t)P);
            auto work = [&](int i) {
                const RawU32& v = nl[(size_t)i];
                int64_t pv = prev[(size_t)i];
                uint64_t g = base[(size_t)i];
                bool t = false;
                auto& o = out[(size_t)i];
                o.reserve(v.n / 2 + 8);
                for (size_t j = 0; j < v.n; j++) {
                    g++;
                    if ((g & 3) == 2) line_done((size_t)(pv + 1), v.d[j], mode, slice, &t, o);
                    pv = (int64_t)v.d[j];
                }
                tl[(size_t)i] = t ? 1 : 0;
            };
            if (pool && P > 1) pool->run(P, work); else for (int i = 0; i < P; i++) work(i);
            size_t add = 0;
            for (auto& o : out) add += o.size();
            locs.reserve(locs.size() + add);
            for (int i = 0; i < P; i++) locs.insert(locs.end(), out[(size_t)i].begin(), out[(size_t)i].end());
        }
        for (int i = 0; i < P; i++) if (tl[(size_t)i]) *too_long = true;
        total_lines += base[(size_t)P] - num;
        num = base[(size_t)P];
        line_start = (size_t)(last + 1);
    }
    // examine fresh bytes; mode 0: too_long set when a short read exceeds 1000; mode 2: drop < slice.
    // With a pool and enough fresh bytes the newline search runs in parallel slices: a line's role depends only on the
    // ordinal of the newline that ends it, so the slices need nothing from each other but their newline counts.
    void scan(int mode, int slice, bool* too_long, Pool* pool, size_t par_min) {
        const char* p = data();
        if (nl_slices > 0) {
            finish_scan(nl_slices, mode, slice, too_long, pool);
            nl_slices = 0;
            scanned = have;
            return;
        }
        if (pool && pool->size() > 1 && have - scanned >= par_min) {
            const size_t begin = scanned, end = have;
            const int P = std::min(pool->size() * 2, (int)((end - begin) / (par_min / 8 + 1)) + 1);
            if (nl.size() < (size_t)P) nl.resize((size_t)P);
            pool->run(P, [&](int i) {
                const size_t a = begin + (end - begin) * (size_t)i / (size_t)P, b = begin + (end - begin) * (size_t)(i + 1) / (size_t)P;
                RawU32& v = nl[(size_t)i];
                v.n = 0;
                v.ensure(b - a + 3);
                v.n = (size_t)(find_newlines(p, a, b, v.d) - v.d);
            });
            finish_scan(P, mode, slice, too_long, pool);
            scanned = have;
            return;
        }
        while (scanned < have) {
            const char* nl = (const char*)memchr(p + scanned, '\n', have - scanned);
            if (!nl) { scanned = have; break; }
            size_t i = (size_t)(nl - p);
            num++; total_lines++;
            if ((num & 3) == 2) line_done(line_start, i, mode, slice, too_long, locs);
            line_start = i + 1;
            scanned = i + 1;
        }
    }
    // drop everything before `from` (a line start); the line phase is kept
    void compact(size_t from) {
        if (rd.map) win += from;   // slide the window
        else memmove(buf->data, buf->data + from, have - from);
        have -= from; scanned -= from; line_start -= from;
        locs.clear();
    }
};

}  // namespace

IngestResult ingest_file(int mode, int slice_length, const char* file1, bool gz1, const char* file2, bool gz2,
                         size_t chunk_bytes, const ChunkSink& sink, Pool* pool, IngestScratch* scratch) {
    IngestResult res{TREW_OK, ""};
    static const std::vector<int32_t> kEmpty;
    const bool auto_chunk = chunk_bytes == 0;
    if (chunk_bytes > ((size_t)1 << 30)) chunk_bytes = (size_t)1 << 30;  // offsets are int32 like the reference's
    // fresh bytes from which reading / newline indexing go parallel (TREW_INGEST_PAR_MIN: tests force the path)
    size_t par_min = (size_t)4 << 20;
    if (const char* e = getenv("TREW_INGEST_PAR_MIN")) par_min = (size_t)std::max(1L, atol(e));
    Side a, b;
    if (scratch) { a.buf = &scratch->a; b.buf = &scratch->b; }
    if (!a.rd.open(file1, gz1)) return IngestResult{TREW_ERR_IO, "File open failed"};
    const bool pair = mode == TREW_MODE_PAIR;
    if (pair && !b.rd.open(file2, gz2)) return IngestResult{TREW_ERR_IO, "File open failed"};
    if (auto_chunk) {
        // blocks small enough that the packer finds part of what the index just read in the last-level cache, large enough
        // that the fork-join hand-offs do not show
        // measured on the 16-core GPU hosts (4 M x 150 bp): plain 22.8 Gbases/s at 64 MiB against 20.4 at 256 MiB and
        // 18.0 at 32 MiB; BGZF 5.5 at 64 MiB against 5.0-5.3 at 32 and 256 MiB
        chunk_bytes = (size_t)64 << 20;
        if (const char* e = getenv("TREW_CHUNK_MB")) { long v = atol(e); if (v > 0 && v <= 1024) chunk_bytes = (size_t)v << 20; }   // experiments
    }
    bool too_long = false;
    const bool trace = getenv("TREW_INGEST_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    if (!pair) {
        for (;;) {
            double t0 = trace ? now() : 0;
            if (!a.fill(chunk_bytes, pool, par_min)) return IngestResult{TREW_ERR_IO, "File-IO Error: " + a.rd.error() + "."};
            double t1 = trace ? now() : 0;
            a.scan(mode, slice_length, &too_long, pool, par_min);
            double t2 = trace ? now() : 0;
            if (too_long) return IngestResult{TREW_ERR_TOO_LONG, trew_status_string(TREW_ERR_TOO_LONG)};
            if (!a.locs.empty()) {
                int rc = sink(a.data(), a.locs, nullptr, kEmpty);
                if (rc) return IngestResult{rc, ""};
            }
            if (trace) fprintf(stderr, "[ingest] block %zu bytes: read %.1f ms, index %.1f ms, sink %.1f ms (%zu reads)\n", a.have, t1 - t0,
                               t2 - t1, now() - t2, a.locs.size() / 2);
            if (a.eof) break;
            a.compact(a.line_start);
        }
        return res;
    }
    // Records are paired index-wise, so when the two files' records differ in size (trimmed mates, 28 + 90 bp
    // libraries) the side with the smaller records finds more of them per block and carries the surplus over.  The
    // carry is bounded: a side that still holds complete unpaired records only tops its buffer up to one block (the
    // reference bounds it the same way by reading LENGTH - 1 - shift, src/kmer.cpp:1062-1066), so `have` stays below
    // two blocks and the int32 offsets cannot wrap.
    size_t carried[2] = {0, 0};   // complete records carried over from the last round, per side
    for (;;) {
        Side* sides[2] = {&a, &b};
        size_t want[2];
        for (int i = 0; i < 2; i++) {
            const size_t have = sides[i]->have;
            want[i] = sides[i]->eof ? 0 : have < chunk_bytes ? chunk_bytes - have : (carried[i] ? 0 : chunk_bytes);
            if (have + want[i] >= (size_t)0x7fffffff)
                return IngestResult{TREW_ERR_PAIRING, "Error: a paired-end record does not fit a 2 GiB block."};
        }
        for (int i = 0; i < 2; i++) if (sides[i]->rd.ownz && !sides[i]->rd.pgz_tried) sides[i]->rd.try_parallel_gz(pool);
        const bool one_stream_per_thread = (gz1 && !a.rd.pgz && !a.rd.bgzf) || (gz2 && !b.rd.pgz && !b.rd.bgzf);
        if (pool && pool->size() > 1 && (gz1 || gz2) && one_stream_per_thread) {
            // two inflate streams are independent: read both mates' blocks at the same time (streams that all threads
            // decode together -- parallel gzip, BGZF -- take the other branch, one file after the other)
            bool ok[2] = {true, true};
            pool->run(2, [&](int i) { if (!sides[i]->eof) ok[i] = sides[i]->fill(want[i], nullptr, par_min, true); });
            for (int i = 0; i < 2; i++) ok[i] = ok[i] && sides[i]->rd.finish_crc(pool);   // the members' CRC-32, on all threads
            if (!ok[0]) return IngestResult{TREW_ERR_IO, "File 1 IO Error: " + a.rd.error() + "."};
            if (!ok[1]) return IngestResult{TREW_ERR_IO, "File 2 IO Error: " + b.rd.error() + "."};
        } else {
            if (!a.eof && !a.fill(want[0], pool, par_min)) return IngestResult{TREW_ERR_IO, "File 1 IO Error: " + a.rd.error() + "."};
            if (!b.eof && !b.fill(want[1], pool, par_min)) return IngestResult{TREW_ERR_IO, "File 2 IO Error: " + b.rd.error() + "."};
        }
        a.scan(mode, slice_length, &too_long, pool, par_min);
        b.scan(mode, slice_length, &too_long, pool, par_min);
        const bool done = a.eof && b.eof;
        if (done && a.total_lines != b.total_lines) {
            char msg[160];
            snprintf(msg, sizeof(msg), "Error: Mismatched record counts between files (num1: %llu, num2: %llu).",
                     (unsigned long long)a.total_lines, (unsigned long long)b.total_lines);
            return IngestResult{TREW_ERR_PAIRING, msg};
        }
        size_t n = std::min(a.locs.size(), b.locs.size()) / 2;
        if (n) {
            int rc;
            if (a.locs.size() == 2 * n && b.locs.size() == 2 * n) {   // the usual case: the blocks hold the same records
                rc = sink(a.data(), a.locs, b.data(), b.locs);
            } else {
                std::vector<int32_t> la(a.locs.begin(), a.locs.begin() + 2 * n), lb(b.locs.begin(), b.locs.begin() + 2 * n);
                rc = sink(a.data(), la, b.data(), lb);
            }
            if (rc) return IngestResult{rc, ""};
        }
        if (done) break;
        // keep unpaired records: restart at the sequence line of the first unpaired record, phase = "header seen"
        for (int i = 0; i < 2; i++) {
            Side* s = sides[i];
            Side* o = sides[1 - i];
            carried[i] = s->locs.size() / 2 - n;
            // the other file is used up: the rest of this one cannot be paired any more and is only read on for the
            // line totals of the mismatch message, so nothing of it is kept
            const bool other_spent = o->eof && o->locs.size() / 2 <= n;
            if (other_spent) carried[i] = 0;
            if (carried[i]) {
                size_t from = (size_t)s->locs[2 * n];
                uint64_t dropped = 0;  // newlines between `from` and the scan position are re-counted
                for (size_t i = from; i < s->scanned; i++) dropped += s->data()[i] == '\n';
                s->num -= dropped; s->total_lines -= dropped;
                s->scanned = from; s->line_start = from;
                s->compact(from);
            } else {
                s->compact(s->line_start);
            }
        }
    }
    return res;
}

}  // namespace trew

extern "C" int trew_ingest_file(int mode, int slice_length, const char* file1, int is_gz1, const char* file2, int is_gz2,
                                uint64_t chunk_bytes, trew_chunk_sink sink, void* user, char* message, size_t message_cap) {
    if (!file1 || !sink || mode < 0 || mode > 2 || (mode == TREW_MODE_PAIR) != (file2 != nullptr)) return TREW_ERR_ARG;
    trew::Pool pool(4);
    trew::IngestResult r = trew::ingest_file(
        mode, slice_length, file1, is_gz1 != 0, file2, is_gz2 != 0, (size_t)chunk_bytes,
        [&](const char* b1, const std::vector<int32_t>& l1, const char* b2, const std::vector<int32_t>& l2) {
            return sink(user, b1, l1.data(), (uint32_t)(l1.size() / 2), b2, b2 ? l2.data() : nullptr, b2 ? (uint32_t)(l2.size() / 2) : 0u);
        },
        &pool);
    if (message && message_cap) { snprintf(message, message_cap, "%s", r.message.c_str()); }
    return r.status;
}
