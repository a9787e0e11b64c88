// Host restatement of the reference's per-file report and cross-file scoring:
//   process_output        src/kmer.cpp:1478-1634   (RC fold, FinalData build, check_ans_seq, sort, >H: / >L:)
//   check_ans_seq         src/kmer.cpp:2549-2569
//   main's accumulation   src/trew.cpp:454-467     (add_data per (k, seq))
//   final_process_output  src/kmer.cpp:2571-2691   (>Putative_TRM)
//   get_score_map         src/kmer.cpp:2693-2761
// Table sizes are O(distinct repeat units), so this stays on the host.
//
// Determinism: the reference sorts with std::sort on keys that tie (and iterates salted hash maps), so the
// order of tied rows -- and, through the top-4 cuts of get_score_map, even some Putative_TRM scores -- is
// not a function of its input.  Here every sort is made total by appending (k, seq)
// ascending, so the output is deterministic; rows and counts are identical to the reference's, row
// order can differ only inside groups the reference leaves unordered.
#include "host_internal.h"

#include <algorithm>
#include <cinttypes>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

namespace {

typedef unsigned __int128 u128;

struct Key {
    int k; u128 seq;
    bool operator<(const Key& o) const { return k != o.k ? k < o.k : seq < o.seq; }
    bool operator==(const Key& o) const { return k == o.k && seq == o.seq; }
};
struct Fin { int64_t forward = 0, backward = 0, both = 0; };
typedef std::map<Key, Fin> FinMap;
typedef std::vector<std::pair<Key, Fin>> FinVec;

u128 canon(u128 w, int k) {  // get_rot_seq_128, src/kmer.cpp:1825-1833
    u128 best = w, cur = w;
    for (int r = 1; r < k; r++) {
        cur = ((cur & 3) << (2 * (k - 1))) | (cur >> 2);
        if (cur < best) best = cur;
    }
    return best;
}

u128 crc(u128 w, int k) {  // rot_reverse_complement, src/kmer.cpp:72-74
    u128 r = 0;
    for (int i = 0; i < k; i++) { r = (r << 2) | (3 - (w & 3)); w >>= 2; }
    return canon(r, k);
}

int dna_count(u128 w, int k) {  // get_dna_count, src/kmer.cpp:1869-1884
    int seen[4] = {0, 0, 0, 0};
    for (int i = 0; i < k; i++) { seen[(int)(w & 3)] = 1; w >>= 2; }
    return seen[0] + seen[1] + seen[2] + seen[3];
}

void int_to_four(char* buf, u128 seq, int n) {  // src/kmer.cpp:1886-1892
    static const char t[4] = {'T', 'G', 'C', 'A'};
    for (int i = 0; i < n; i++) { buf[n - 1 - i] = t[(int)(seq & 3)]; seq >>= 2; }
    buf[n] = 0;
}

// check_ans_seq: false when the unit is itself k'-periodic for some 3 <= k' < MIN_MER
bool check_ans_seq(const Key& key, int min_mer) {
    for (int k = 3; k < min_mer; k++) {
        u128 mask = (((u128)1) << (2 * k)) - 1;
        u128 num = key.seq, bef = 0;
        int i;
        for (i = 0; i < key.k - k + 1; i++) {
            u128 cur = canon(num & mask, k);
            if (i > 0 && cur != bef) break;
            bef = cur;
            num >>= 2;
        }
        if (i == key.k - k + 1) return false;
    }
    return true;
}

void sort_report(FinVec& v) {
    std::sort(v.begin(), v.end(), [](const std::pair<Key, Fin>& a, const std::pair<Key, Fin>& b) {
        if (a.second.forward != b.second.forward) return a.second.forward > b.second.forward;
        if (a.second.both != b.second.both) return a.second.both > b.second.both;
        return a.first < b.first;
    });
}

void append_rows(std::string& out, const FinVec& v) {
    char seq[65], line[256];
    for (auto& kv : v) {
        const Fin& f = kv.second;
        if (f.forward + f.backward + f.both >= 10) {  // ABS_MIN_PRINT_COUNT
            int_to_four(seq, kv.first.seq, kv.first.k);
            snprintf(line, sizeof(line), "%d,%s,%" PRId64 ",%" PRId64 ",%" PRId64 ",%c\n", kv.first.k, seq,
                     std::max(f.forward, f.backward), std::min(f.forward, f.backward), f.both,
                     f.forward > f.backward ? '+' : (f.forward < f.backward ? '-' : '?'));
            out += line;
        }
    }
}

// one class (high or low) of process_output: src/kmer.cpp:1518-1613
FinVec build_class(std::map<Key, uint64_t>& F, const std::map<Key, uint64_t>& B, const std::map<Key, uint64_t>& O, int min_mer) {
    for (auto& kv : B) F[Key{kv.first.k, crc(kv.first.seq, kv.first.k)}] += kv.second;
    FinMap fin;
    for (auto& kv : F) {
        u128 t = crc(kv.first.seq, kv.first.k);
        u128 kseq = std::min(t, kv.first.seq);
        Key key{kv.first.k, kseq};
        auto it = fin.find(key);
        if (it == fin.end()) { Fin f; f.backward = (t == kv.first.seq) ? -1 : 0; it = fin.emplace(key, f).first; }
        if (kseq == kv.first.seq) it->second.forward = (int64_t)kv.second;
        else it->second.backward = (int64_t)kv.second;
    }
    for (auto& kv : O) {  // looked up by the RAW key, assignment not += (src/kmer.cpp:1541-1549)
        auto it = fin.find(kv.first);
        if (it != fin.end()) it->second.both = (int64_t)kv.second;
        else { Fin f; f.backward = (crc(kv.first.seq, kv.first.k) == kv.first.seq) ? -1 : 0; f.both = (int64_t)kv.second; fin.emplace(kv.first, f); }
    }
    FinVec v;
    for (auto& kv : fin) if (check_ans_seq(kv.first, min_mer)) v.emplace_back(kv.first, kv.second);
    sort_report(v);
    return v;
}

// get_score_map, src/kmer.cpp:2693-2761
std::map<Key, uint32_t> score_map(const FinMap& total) {
    FinVec vec;
    for (auto& kv : total) {
        const Fin& v = kv.second;
        if (v.forward + v.backward + v.both >= 10) {
            if (v.backward > v.forward) { Fin s; s.forward = v.backward; s.backward = v.forward; s.both = v.both; vec.emplace_back(kv.first, s); }
            else vec.emplace_back(kv.first, v);
        }
    }
    FinMap ratio;
    std::map<Key, uint32_t> score;
    std::sort(vec.begin(), vec.end(), [](auto& a, auto& b) {
        if (a.second.forward != b.second.forward) return a.second.forward > b.second.forward;
        return a.first < b.first;
    });
    int cnt = 0;
    for (auto& kv : vec) {
        if (kv.second.forward == 0 || cnt >= 20) break;  // NUM_RAT_CAND
        if (kv.second.backward >= 0) { cnt++; ratio[kv.first] = kv.second; }
    }
    for (size_t i = 0; i < std::min<size_t>(4, vec.size()); i++) {  // NUM_FOR_MAX_COUNT
        if (vec[i].second.forward == 0) break;
        score[vec[i].first] += 1;
    }
    std::sort(vec.begin(), vec.end(), [](auto& a, auto& b) {
        int64_t ta = a.second.forward + a.second.backward + a.second.both, tb = b.second.forward + b.second.backward + b.second.both;
        if (ta != tb) return ta > tb;
        return a.first < b.first;
    });
    cnt = 0;
    for (auto& kv : vec) {
        if (cnt >= 20) break;
        if (kv.second.forward > 0 && kv.second.backward >= 0) { cnt++; ratio[kv.first] = kv.second; }
    }
    for (size_t i = 0; i < std::min<size_t>(4, vec.size()); i++) score[vec[i].first] += 1;  // NUM_TOT_MAX_COUNT
    FinVec rv(ratio.begin(), ratio.end());
    std::sort(rv.begin(), rv.end(), [](auto& a, auto& b) {
        double ra = (double)a.second.backward / a.second.forward, rb = (double)b.second.backward / b.second.forward;
        if (ra != rb) return ra < rb;
        return a.first < b.first;
    });
    for (size_t i = 0; i < std::min<size_t>(4, rv.size()); i++) score[rv[i].first] += 1;  // NUM_RAT_MAX_COUNT
    return score;
}

}  // namespace

struct trew_report {
    int min_mer = 5;
    FinMap total_high, total_low;
    std::string text;
    bool finished = false;
};

extern "C" {

int trew_report_create(int min_mer, trew_report** out) {
    if (!out || min_mer < 3) return TREW_ERR_ARG;
    *out = new trew_report();
    (*out)->min_mer = min_mer;
    return TREW_OK;
}

void trew_report_destroy(trew_report* r) { delete r; }

int trew_report_add_file(trew_report* r, const char* file_name, const trew_entry* entries, uint64_t n) {
    if (!r || !file_name || (n && !entries)) return TREW_ERR_ARG;
    std::map<Key, uint64_t> maps[6];
    for (uint64_t i = 0; i < n; i++) {
        const trew_entry& e = entries[i];
        if (e.table < 0 || e.table > 5) return TREW_ERR_ARG;
        maps[e.table][Key{e.k, ((u128)e.seq_hi << 64) | e.seq_lo}] += e.count;
    }
    FinVec low = build_class(maps[TREW_TABLE_FORWARD_LOW], maps[TREW_TABLE_BACKWARD_LOW], maps[TREW_TABLE_BOTH_LOW], r->min_mer);
    FinVec high = build_class(maps[TREW_TABLE_FORWARD_HIGH], maps[TREW_TABLE_BACKWARD_HIGH], maps[TREW_TABLE_BOTH_HIGH], r->min_mer);
    r->text += ">H:"; r->text += file_name; r->text += "\n";
    append_rows(r->text, high);
    r->text += ">L:"; r->text += file_name; r->text += "\n";
    append_rows(r->text, low);
    // src/trew.cpp:454-467: ALL kept entries (not only printed ones) accumulate across files
    for (auto& kv : high) { Fin& t = r->total_high[kv.first]; t.forward += kv.second.forward; t.backward += kv.second.backward; t.both += kv.second.both; }
    for (auto& kv : low) { Fin& t = r->total_low[kv.first]; t.forward += kv.second.forward; t.backward += kv.second.backward; t.both += kv.second.both; }
    return TREW_OK;
}

int trew_report_text(trew_report* r, const char** text, size_t* len) {
    if (!r) return TREW_ERR_ARG;
    if (text) *text = r->text.c_str();
    if (len) *len = r->text.size();
    return TREW_OK;
}

int trew_report_finish(trew_report* r, const char** text, size_t* len) {
    if (!r) return TREW_ERR_ARG;
    if (!r->finished) {
        r->finished = true;
        bool any = false;
        for (auto& kv : r->total_high) if (kv.second.forward + kv.second.backward + kv.second.both >= 20) { any = true; break; }  // ABS_MIN_ANS_COUNT
        for (auto& kv : r->total_low) if (kv.second.forward + kv.second.backward + kv.second.both >= 20) { any = true; break; }
        r->text += ">Putative_TRM\n";
        if (any) {
            std::map<Key, uint32_t> score = score_map(r->total_low);
            for (auto& kv : score_map(r->total_high)) score[kv.first] += kv.second;
            struct Row { Key key; uint32_t score; int dna; int dir; };
            std::vector<Row> rows;
            for (auto& kv : score) {
                Fin low = r->total_low[kv.first], high = r->total_high[kv.first];  // operator[] default-inserts like the reference
                int bonus = 0;
                int high_dir = high.forward > high.backward ? 1 : (high.forward < high.backward ? -1 : 0);
                int low_dir = low.forward > low.backward ? 1 : (low.forward < low.backward ? -1 : 0);
                int final_dir;
                if (low_dir != 0 && low_dir == high_dir) { bonus += 1; final_dir = low_dir; }
                else if (low_dir == 0 && high_dir != 0) final_dir = high_dir;
                else if (low_dir != 0 && high_dir == 0) final_dir = low_dir;
                else if (low_dir != high_dir && (low.forward > 0 || low.backward > 0 || high.forward > 0 || high.backward > 0)) {
                    if (low.forward < low.backward) std::swap(low.forward, low.backward);
                    if (high.forward < high.backward) std::swap(high.forward, high.backward);
                    if (low.backward * high.forward == high.backward * low.forward)
                        final_dir = (low.forward + low.backward > high.forward + high.backward) ? low_dir : high_dir;
                    else if (low.backward * high.forward < high.backward * low.forward) final_dir = low_dir;
                    else final_dir = high_dir;
                } else final_dir = 0;
                int dc = dna_count(kv.first.seq, kv.first.k);
                if (dc > 2) bonus += 1;
                rows.push_back(Row{kv.first, kv.second + (uint32_t)bonus, dc, final_dir});
            }
            std::sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) {
                if (a.score != b.score) return a.score > b.score;
                if (a.dna != b.dna) return a.dna > b.dna;
                return a.key < b.key;
            });
            char seq[65], line[160];
            for (size_t i = 0; i < std::min<size_t>(10, rows.size()); i++) {  // ABS_MAX_ANS_NUM
                int_to_four(seq, rows[i].key.seq, rows[i].key.k);
                snprintf(line, sizeof(line), "%d,%s,%" PRIu32 ",%c\n", rows[i].key.k, seq, rows[i].score,
                         rows[i].dir == 1 ? '+' : (rows[i].dir == -1 ? '-' : '?'));
                r->text += line;
            }
        } else {
            r->text += "NO_PUTATIVE_TRM,-1\n";
        }
    }
    if (text) *text = r->text.c_str();
    if (len) *len = r->text.size();
    return TREW_OK;
}

}  // extern "C"
