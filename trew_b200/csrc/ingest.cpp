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
#include <cstdio>
#include <cstring>

#include <zlib.h>

namespace trew {

namespace {

struct Reader {
    bool gz = false;
    FILE* fp = nullptr;
    gzFile gfp = nullptr;
    bool open(const char* name, bool is_gz) {
        gz = is_gz;
        if (gz) { gfp = gzopen(name, "r"); if (gfp) gzbuffer(gfp, 1 << 20); return gfp != nullptr; }
        fp = fopen(name, "r");
        return fp != nullptr;
    }
    // returns bytes read (0 at EOF), -1 on error
    long read(char* buf, size_t n) {
        if (gz) {
            int r = gzread(gfp, buf, (unsigned)std::min<size_t>(n, 1u << 30));
            return r < 0 ? -1 : r;
        }
        size_t r = fread(buf, 1, n, fp);
        if (r == 0 && ferror(fp)) return -1;
        return (long)r;
    }
    std::string error() {
        if (gz) { int e; return gzerror(gfp, &e); }
        return strerror(errno);
    }
    void close() { if (gz) { if (gfp) gzclose(gfp); gfp = nullptr; } else { if (fp) fclose(fp); fp = nullptr; } }
    ~Reader() { close(); }
};

// One side of the ingest: a buffer holding [carried bytes | fresh bytes] and the line phase.
struct Side {
    Reader rd;
    std::vector<char> buf;
    size_t have = 0;        // bytes in buf
    size_t scanned = 0;     // bytes already examined for newlines
    size_t line_start = 0;  // start of the line being assembled
    uint64_t num = 0;       // newlines seen so far (the reference's `num`)
    uint64_t total_lines = 0;
    bool eof = false;
    std::vector<int32_t> locs;  // sequence lines found in buf, inclusive (st, nd)

    // read more bytes; returns false on I/O error
    bool fill(size_t chunk) {
        if (buf.size() < have + chunk) buf.resize(have + chunk);
        long r = rd.read(buf.data() + have, chunk);
        if (r < 0) return false;
        if (r == 0) eof = true;
        have += (size_t)r;
        return true;
    }
    // examine fresh bytes; mode 0: too_long set when a short read exceeds 1000; mode 2: drop < slice
    void scan(int mode, int slice, bool* too_long) {
        const char* p = buf.data();
        while (scanned < have) {
            const char* nl = (const char*)memchr(p + scanned, '\n', have - scanned);
            if (!nl) { scanned = have; break; }
            size_t i = (size_t)(nl - p);
            num++; total_lines++;
            if ((num & 3) == 2) {
                size_t len = i - line_start;
                if (mode == TREW_MODE_SHORT && len > 1000) *too_long = true;
                if (!(mode == TREW_MODE_LONG && len < (size_t)slice)) {
                    locs.push_back((int32_t)line_start);
                    locs.push_back((int32_t)i - 1);
                }
            }
            line_start = i + 1;
            scanned = i + 1;
        }
    }
    // drop everything before `from` (a line start); the line phase is kept
    void compact(size_t from) {
        memmove(buf.data(), buf.data() + from, have - from);
        have -= from; scanned -= from; line_start -= from;
        locs.clear();
    }
};

}  // namespace

IngestResult ingest_file(int mode, int slice_length, const char* file1, bool gz1, const char* file2, bool gz2,
                         size_t chunk_bytes, const ChunkSink& sink) {
    IngestResult res{TREW_OK, ""};
    static const std::vector<int32_t> kEmpty;
    if (chunk_bytes > ((size_t)1 << 30)) chunk_bytes = (size_t)1 << 30;  // offsets are int32 like the reference's
    Side a, b;
    if (!a.rd.open(file1, gz1)) return IngestResult{TREW_ERR_IO, "File open failed"};
    const bool pair = mode == TREW_MODE_PAIR;
    if (pair && !b.rd.open(file2, gz2)) return IngestResult{TREW_ERR_IO, "File open failed"};
    bool too_long = false;
    if (!pair) {
        for (;;) {
            if (!a.fill(chunk_bytes)) return IngestResult{TREW_ERR_IO, "File-IO Error: " + a.rd.error() + "."};
            a.scan(mode, slice_length, &too_long);
            if (too_long) return IngestResult{TREW_ERR_TOO_LONG, trew_status_string(TREW_ERR_TOO_LONG)};
            if (!a.locs.empty()) {
                int rc = sink(a.buf.data(), a.locs, nullptr, kEmpty);
                if (rc) return IngestResult{rc, ""};
            }
            if (a.eof) break;
            a.compact(a.line_start);
        }
        return res;
    }
    for (;;) {
        if (!a.eof && !a.fill(chunk_bytes)) return IngestResult{TREW_ERR_IO, "File 1 IO Error: " + a.rd.error() + "."};
        if (!b.eof && !b.fill(chunk_bytes)) return IngestResult{TREW_ERR_IO, "File 2 IO Error: " + b.rd.error() + "."};
        a.scan(mode, slice_length, &too_long);
        b.scan(mode, slice_length, &too_long);
        const bool done = a.eof && b.eof;
        if (done && a.total_lines != b.total_lines) {
            char msg[160];
            snprintf(msg, sizeof(msg), "Error: Mismatched record counts between files (num1: %llu, num2: %llu).",
                     (unsigned long long)a.total_lines, (unsigned long long)b.total_lines);
            return IngestResult{TREW_ERR_PAIRING, msg};
        }
        size_t n = std::min(a.locs.size(), b.locs.size()) / 2;
        if (n) {
            std::vector<int32_t> la(a.locs.begin(), a.locs.begin() + 2 * n), lb(b.locs.begin(), b.locs.begin() + 2 * n);
            int rc = sink(a.buf.data(), la, b.buf.data(), lb);
            if (rc) return IngestResult{rc, ""};
        }
        if (done) break;
        // keep unpaired records: restart at the sequence line of the first unpaired record, phase = "header seen"
        for (Side* s : {&a, &b}) {
            if (s->locs.size() / 2 > n) {
                size_t from = (size_t)s->locs[2 * n];
                uint64_t dropped = 0;  // newlines between `from` and the scan position are re-counted
                for (size_t i = from; i < s->scanned; i++) dropped += s->buf[i] == '\n';
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
    trew::IngestResult r = trew::ingest_file(
        mode, slice_length, file1, is_gz1 != 0, file2, is_gz2 != 0, chunk_bytes ? (size_t)chunk_bytes : ((size_t)32 << 20),
        [&](const char* b1, const std::vector<int32_t>& l1, const char* b2, const std::vector<int32_t>& l2) {
            return sink(user, b1, l1.data(), (uint32_t)(l1.size() / 2), b2, b2 ? l2.data() : nullptr, b2 ? (uint32_t)(l2.size() / 2) : 0u);
        });
    if (message && message_cap) { snprintf(message, message_cap, "%s", r.message.c_str()); }
    return r.status;
}
