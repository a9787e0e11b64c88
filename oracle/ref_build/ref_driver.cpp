// TEST INFRASTRUCTURE ONLY (oracle/_ref): a driver around the UNMODIFIED reference sources
// /root/reference/src/kmer.cpp + kmer.h, compiled where they lie (see oracle/ref_build/Makefile).
//
// The reference's own main (src/trew.cpp) needs p-ranav/argparse, which is not installed, so this
// file replaces it:
//   * it defines the nine globals of src/trew.cpp:10-20,
//   * `trew_ref` (built with -DREF_MAIN) mirrors src/trew.cpp:382-477: table setup -> per-file
//     process_kmer* -> add_data accumulation -> final_process_output, with a hand-rolled parser
//     for the same flags, plus `--dump-tables FILE` which re-creates the worker orchestration of
//     src/kmer.cpp:1266-1476 around the reference's buffer_task* so the six raw count maps can be
//     written out before process_output folds them,
//   * `libtrew_ref.so` exports a small C ABI (ref_*) used by tests/ and tools/ through ctypes to
//     drive the reference's primitives and buffer_task* on in-memory chunks.
// Nothing here is part of the product; nothing here is copied from the reference.
#include "kmer.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <map>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

int MAX_MER;
int MIN_MER;
int TABLE_MAX_MER;
int NUM_THREAD;
int SLICE_LENGTH;
int QUEUE_SIZE;
double LOW_BASELINE;
double HIGH_BASELINE;
bool INDEX = true;

namespace {

struct Tables {
    uint8_t** repeat_check_table = nullptr;
    uint32_t** rot_table = nullptr;
    uint64_t* extract_k_mer = nullptr;
    uint128_t* extract_k_mer_128 = nullptr;
    uint128_t* extract_k_mer_ans = nullptr;
    int built_table_top = -1;  // MIN(MAX_MER, TABLE_MAX_MER) the rot tables were built for
};
Tables g_tab;

// mirrors src/trew.cpp:382-401
void setup_tables() {
    int top = MIN(MAX_MER, TABLE_MAX_MER);
    if (MIN_MER <= TABLE_MAX_MER) {
        if (g_tab.built_table_top != top) {
            if (g_tab.rot_table != nullptr) {
                for (int i = 0; i <= g_tab.built_table_top - ABS_MIN_MER; i++) {
                    free(g_tab.rot_table[i]);
                    free(g_tab.repeat_check_table[i]);
                }
                free(g_tab.rot_table);
                free(g_tab.repeat_check_table);
            }
            g_tab.repeat_check_table = set_repeat_check_table();
            g_tab.rot_table = set_rotation_table(g_tab.repeat_check_table);
            g_tab.built_table_top = top;
        }
    }
    free(g_tab.extract_k_mer); g_tab.extract_k_mer = nullptr;
    free(g_tab.extract_k_mer_128); g_tab.extract_k_mer_128 = nullptr;
    free(g_tab.extract_k_mer_ans); g_tab.extract_k_mer_ans = nullptr;
    if (MAX_MER <= ABS_UINT64_MAX_MER) g_tab.extract_k_mer = set_extract_k_mer();
    else g_tab.extract_k_mer_128 = set_extract_k_mer_128();
    if (MIN_MER > ABS_MIN_MER) g_tab.extract_k_mer_ans = set_extract_k_mer_ans();
}

struct Entry { int table; int k; uint64_t lo, hi; uint64_t count; };

// table ids: 0 forward.high 1 forward.low 2 backward.high 3 backward.low 4 both.high 5 both.low
void collect(const ResultMapData& r, std::map<std::tuple<int, int, uint64_t, uint64_t>, uint64_t>& acc) {
    ResultMap* maps[6] = {r.forward.first, r.forward.second, r.backward.first, r.backward.second,
                          r.both.first, r.both.second};
    for (int t = 0; t < 6; t++) {
        for (auto& [key, v] : *maps[t]) {
            acc[{t, key.first, (uint64_t)(key.second >> 64), (uint64_t)key.second}] += v;
        }
    }
}

void free_result(ResultMapData& r) {
    delete r.forward.first; delete r.forward.second;
    delete r.backward.first; delete r.backward.second;
    delete r.both.first; delete r.both.second;
}

std::vector<Entry> g_result;

void publish(const std::map<std::tuple<int, int, uint64_t, uint64_t>, uint64_t>& acc) {
    g_result.clear();
    for (auto& [key, v] : acc) {
        g_result.push_back(Entry{std::get<0>(key), std::get<1>(key), std::get<3>(key), std::get<2>(key), v});
    }
}

// One ThreadData per distinct (MIN, MAX, TABLE_MAX) config; the reference never frees its scratch.
std::map<std::tuple<int, int, int>, std::vector<ThreadData*>> g_thread_data;
ThreadData* thread_data_for(int idx) {
    auto& v = g_thread_data[{MIN_MER, MAX_MER, TABLE_MAX_MER}];
    while ((int)v.size() <= idx) v.push_back(new ThreadData());
    return v[idx];
}

}  // namespace

extern "C" {

int ref_init(int min_mer, int max_mer, int table_max_mer, double low, double high, int slice_length) {
    MIN_MER = min_mer; MAX_MER = max_mer; TABLE_MAX_MER = table_max_mer;
    LOW_BASELINE = low; HIGH_BASELINE = high; SLICE_LENGTH = slice_length;
    NUM_THREAD = 1; QUEUE_SIZE = -1;
    setup_tables();
    return 0;
}

// mode: 0 short single (buffer_task), 1 paired (buffer_task_pair), 2 long (buffer_task_long).
// locs are inclusive (st, nd) pairs, flattened; the chunk is copied because the consumer frees it.
// Returns the number of (table, k, seq) entries; fetch them with ref_result_copy.
int ref_scan(int mode, const char* buf1, long len1, const int* locs1, int n1,
             const char* buf2, long len2, const int* locs2, int n2) {
    std::map<std::tuple<int, int, uint64_t, uint64_t>, uint64_t> acc;
    ResultMapData r;
    if (mode == 1) {
        TBBPairQueue q;
        char* b1 = (char*)malloc(len1 + 1); memcpy(b1, buf1, len1); b1[len1] = 0;
        char* b2 = (char*)malloc(len2 + 1); memcpy(b2, buf2, len2); b2[len2] = 0;
        auto* l1 = new LocationVector(); auto* l2 = new LocationVector();
        for (int i = 0; i < n1; i++) l1->emplace_back(locs1[2 * i], locs1[2 * i + 1]);
        for (int i = 0; i < n2; i++) l2->emplace_back(locs2[2 * i], locs2[2 * i + 1]);
        q.push(PairQueueData{b1, b2, l1, l2});
        q.push(PairQueueData{nullptr, nullptr, nullptr, nullptr});
        r = buffer_task_pair(&q, thread_data_for(0), g_tab.rot_table, g_tab.extract_k_mer, g_tab.extract_k_mer_128, g_tab.repeat_check_table);
    } else {
        TBBQueue q;
        char* b1 = (char*)malloc(len1 + 1); memcpy(b1, buf1, len1); b1[len1] = 0;
        auto* l1 = new LocationVector();
        for (int i = 0; i < n1; i++) l1->emplace_back(locs1[2 * i], locs1[2 * i + 1]);
        q.push(QueueData{b1, l1});
        q.push(QueueData{nullptr, nullptr});
        if (mode == 0) r = buffer_task(&q, thread_data_for(0), g_tab.rot_table, g_tab.extract_k_mer, g_tab.extract_k_mer_128, g_tab.repeat_check_table);
        else r = buffer_task_long(&q, thread_data_for(0), g_tab.rot_table, g_tab.extract_k_mer, g_tab.extract_k_mer_128, g_tab.repeat_check_table);
    }
    collect(r, acc);
    free_result(r);
    publish(acc);
    return (int)g_result.size();
}

int ref_result_size() { return (int)g_result.size(); }

void ref_result_copy(int* table, int* k, uint64_t* lo, uint64_t* hi, uint64_t* count) {
    for (size_t i = 0; i < g_result.size(); i++) {
        table[i] = g_result[i].table; k[i] = g_result[i].k;
        lo[i] = g_result[i].lo; hi[i] = g_result[i].hi; count[i] = g_result[i].count;
    }
}

// The primitive: k_mer_check / k_mer_check_128 on seq[st..nd] with k in [min_mer, max_mer].
// out[0]=target_k_high out[1]=target_k_low; seqs[0..1]=S_high (lo,hi) seqs[2..3]=S_low (lo,hi).
// Emissions (table 0 = high map, 1 = low map) are left in the result buffer.
int ref_k_mer_check(const char* seq, int st, int nd, int min_mer, int max_mer, int* out, uint64_t* seqs) {
    ThreadData* td = thread_data_for(0);
    auto [k_mer_counter, k_mer_data, k_mer_data_128, k_mer_counter_list] = td->init_check();
    std::vector<int16_t> total(MAX_MER - MIN_MER + 2);
    ResultMapPair rp = {new ResultMap{}, new ResultMap{}};
    KmerData kd;
    if (MAX_MER <= ABS_UINT64_MAX_MER) {
        CounterMap* cm = TABLE_MAX_MER < MAX_MER ? new CounterMap[MAX_MER - TABLE_MAX_MER] : nullptr;
        std::pair<uint64_t, uint64_t> rs{0, 0};
        kd = k_mer_check(seq, st, nd, g_tab.rot_table, g_tab.extract_k_mer, k_mer_counter, cm, k_mer_data,
                         k_mer_counter_list, g_tab.repeat_check_table, rp, total.data(), min_mer, max_mer, &rs);
        seqs[0] = rs.first; seqs[1] = 0; seqs[2] = rs.second; seqs[3] = 0;
        delete[] cm;
    } else {
        CounterMap_128* cm = new CounterMap_128[MAX_MER - TABLE_MAX_MER];
        std::pair<uint128_t, uint128_t> rs{0, 0};
        kd = k_mer_check_128(seq, st, nd, g_tab.rot_table, g_tab.extract_k_mer_128, k_mer_counter, cm, k_mer_data_128,
                             k_mer_counter_list, g_tab.repeat_check_table, rp, total.data(), min_mer, max_mer, &rs);
        seqs[0] = (uint64_t)rs.first; seqs[1] = (uint64_t)(rs.first >> 64);
        seqs[2] = (uint64_t)rs.second; seqs[3] = (uint64_t)(rs.second >> 64);
        delete[] cm;
    }
    out[0] = kd.first; out[1] = kd.second;
    std::map<std::tuple<int, int, uint64_t, uint64_t>, uint64_t> acc;
    for (auto& [key, v] : *rp.first) acc[{0, key.first, (uint64_t)(key.second >> 64), (uint64_t)key.second}] += v;
    for (auto& [key, v] : *rp.second) acc[{1, key.first, (uint64_t)(key.second >> 64), (uint64_t)key.second}] += v;
    delete rp.first; delete rp.second;
    publish(acc);
    return (int)g_result.size();
}

void ref_get_rot_seq(uint64_t lo, uint64_t hi, int k, uint64_t* out) {
    uint128_t v = get_rot_seq_128(absl::MakeUint128(hi, lo), k);
    out[0] = (uint64_t)v; out[1] = (uint64_t)(v >> 64);
}

void ref_rot_reverse_complement(uint64_t lo, uint64_t hi, int k, uint64_t* out) {
    KmerSeq r = rot_reverse_complement(KmerSeq{k, absl::MakeUint128(hi, lo)});
    out[0] = (uint64_t)r.second; out[1] = (uint64_t)(r.second >> 64);
}

int ref_get_repeat_check(uint64_t lo, uint64_t hi, int k) {
    return get_repeat_check(absl::MakeUint128(hi, lo), k);
}

void ref_int_to_four(uint64_t lo, uint64_t hi, int k, char* buffer) {
    int_to_four(buffer, absl::MakeUint128(hi, lo), k);
}

}  // extern "C"

#ifdef REF_MAIN
namespace {

// The worker orchestration of src/kmer.cpp:1266-1476, re-created so the raw maps can be dumped.
template <class Queue, class Reader, class Worker>
void run_workers(Queue& q, Reader reader, Worker worker, ResultMapData* result_list) {
    if (QUEUE_SIZE >= 4) q.set_capacity(QUEUE_SIZE / 4);
    std::vector<std::thread> th;
    for (int i = 0; i < NUM_THREAD; i++) th.emplace_back([&, i] { result_list[i] = worker(i); });
    reader();
    for (int i = 0; i < NUM_THREAD; i++) q.push({});
    for (auto& t : th) t.join();
}

FileReader open_reader(const char* name, bool is_gz) {
    if (is_gz) {
        gzFile fp = gzopen(name, "r");
        if (fp == nullptr) { fprintf(stderr, "File open failed\n"); exit(EXIT_FAILURE); }
        return FileReader(fp);
    }
    FILE* fp = fopen(name, "r");
    if (fp == nullptr) { fprintf(stderr, "File open failed\n"); exit(EXIT_FAILURE); }
    return FileReader(fp);
}

FILE* g_dump = nullptr;

void dump_tables(const char* file_name, ResultMapData* result_list) {
    std::map<std::tuple<int, int, uint64_t, uint64_t>, uint64_t> acc;
    for (int i = 0; i < NUM_THREAD; i++) collect(result_list[i], acc);
    static const char* names[6] = {"F_h", "F_l", "B_h", "B_l", "O_h", "O_l"};
    char buffer[ABS_MAX_MER + 1];
    fprintf(g_dump, "#file %s\n", file_name);
    for (auto& [key, v] : acc) {
        int_to_four(buffer, absl::MakeUint128(std::get<2>(key), std::get<3>(key)), std::get<1>(key));
        fprintf(g_dump, "%s %d %s %llu\n", names[std::get<0>(key)], std::get<1>(key), buffer, (unsigned long long)v);
    }
}

FinalFastqOutput capture_file(int mode, const char* f1, const char* f2, bool gz1, bool gz2) {
    ResultMapData* result_list = (ResultMapData*)malloc(sizeof(ResultMapData) * NUM_THREAD);
    if (mode == 1) {
        TBBPairQueue q;
        FileReader r1 = open_reader(f1, gz1), r2 = open_reader(f2, gz2);
        run_workers(q, [&] { read_pair_fastq_thread(r1, r2, &q); r1.close(); r2.close(); },
                    [&](int i) { return buffer_task_pair(&q, thread_data_for(i), g_tab.rot_table, g_tab.extract_k_mer, g_tab.extract_k_mer_128, g_tab.repeat_check_table); },
                    result_list);
    } else {
        TBBQueue q;
        FileReader r = open_reader(f1, gz1);
        if (mode == 0) {
            run_workers(q, [&] { read_fastq_thread(r, &q); r.close(); },
                        [&](int i) { return buffer_task(&q, thread_data_for(i), g_tab.rot_table, g_tab.extract_k_mer, g_tab.extract_k_mer_128, g_tab.repeat_check_table); },
                        result_list);
        } else {
            run_workers(q, [&] { read_fastq_long_thread(r, &q); r.close(); },
                        [&](int i) { return buffer_task_long(&q, thread_data_for(i), g_tab.rot_table, g_tab.extract_k_mer, g_tab.extract_k_mer_128, g_tab.repeat_check_table); },
                        result_list);
        }
    }
    dump_tables(f1, result_list);
    return process_output(f1, result_list, g_tab.rot_table, g_tab.extract_k_mer_ans);
}

bool has_gz_ext(const std::filesystem::path& p) {
    std::string e = p.extension().string();
    return e == ".gz" || e == ".bgz";
}

[[noreturn]] void usage() {
    fprintf(stderr, "usage: trew_ref short|long MIN_MER MAX_MER [FASTQ...] [-t N] [-m M] [-L x] [-H x] [-s S] [-q Q]\n"
                    "                [--paired_end --fq1 A... --fq2 B...] [--dump-tables FILE]\n");
    exit(1);
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 4) usage();
    std::string cmd = argv[1];
    if (cmd != "short" && cmd != "long") usage();
    bool is_short = cmd == "short";
    MIN_MER = atoi(argv[2]); MAX_MER = atoi(argv[3]);
    NUM_THREAD = 2; TABLE_MAX_MER = 12; LOW_BASELINE = 0.5; HIGH_BASELINE = 0.8; SLICE_LENGTH = 150; QUEUE_SIZE = -1;
    bool paired = false;
    std::vector<std::string> files, fq1, fq2;
    std::string dump_path;
    std::vector<std::string>* sink = &files;
    for (int i = 4; i < argc; i++) {
        std::string a = argv[i];
        auto need = [&]() -> const char* { if (i + 1 >= argc) usage(); return argv[++i]; };
        if (a == "-t" || a == "--thread") { NUM_THREAD = atoi(need()); sink = &files; }
        else if (a == "-m" || a == "--table_max_mer") { TABLE_MAX_MER = atoi(need()); sink = &files; }
        else if (a == "-L" || a == "--low_baseline") { LOW_BASELINE = atof(need()); sink = &files; }
        else if (a == "-H" || a == "--high_baseline") { HIGH_BASELINE = atof(need()); sink = &files; }
        else if (a == "-s" || a == "--slice_length") { SLICE_LENGTH = atoi(need()); sink = &files; }
        else if (a == "-q" || a == "--queue_size") { QUEUE_SIZE = atoi(need()); sink = &files; }
        else if (a == "--dump-tables") { dump_path = need(); sink = &files; }
        else if (a == "--paired_end") { paired = true; sink = &files; }
        else if (a == "--fq1") sink = &fq1;
        else if (a == "--fq2") sink = &fq2;
        else sink->push_back(a);
    }
    if (MIN_MER > MAX_MER || MIN_MER < ABS_MIN_MER || MAX_MER > ABS_MAX_MER || TABLE_MAX_MER > ABS_TABLE_MAX_MER ||
        TABLE_MAX_MER <= 0 || NUM_THREAD < 1 || !(0 < LOW_BASELINE && LOW_BASELINE <= 1) ||
        !(0 < HIGH_BASELINE && HIGH_BASELINE <= 1) || LOW_BASELINE > HIGH_BASELINE ||
        (!is_short && SLICE_LENGTH < 2 * MAX_MER)) usage();
    bool is_pair = is_short && paired;
    std::vector<std::filesystem::path> fastq_path_list;
    if (is_pair) {
        if (fq1.size() != fq2.size() || fq1.empty() || !files.empty()) usage();
        for (size_t i = 0; i < fq1.size(); i++) { fastq_path_list.emplace_back(fq1[i]); fastq_path_list.emplace_back(fq2[i]); }
    } else {
        if (files.empty()) usage();
        for (auto& f : files) fastq_path_list.emplace_back(f);
    }
    for (auto& p : fastq_path_list) {
        if (!std::filesystem::is_regular_file(p)) { fprintf(stderr, "%s : file not found\n", p.c_str()); return 1; }
    }
    if (!dump_path.empty()) {
        g_dump = fopen(dump_path.c_str(), "w");
        if (g_dump == nullptr) { fprintf(stderr, "cannot open %s\n", dump_path.c_str()); return 1; }
    }

    setup_tables();  // src/trew.cpp:382-401

    // src/trew.cpp:403-476
    FinalFastqData* total_result_low = new FinalFastqData{};
    FinalFastqData* total_result_high = new FinalFastqData{};
    ThreadData* thread_data_list = new ThreadData[NUM_THREAD];
    for (size_t i = 0; i < fastq_path_list.size() / (is_pair ? 2 : 1); ++i) {
        FinalFastqOutput out;
        if (is_pair) {
            std::string a = std::filesystem::canonical(fastq_path_list[2 * i]).string();
            std::string b = std::filesystem::canonical(fastq_path_list[2 * i + 1]).string();
            bool g1 = has_gz_ext(fastq_path_list[2 * i]), g2 = has_gz_ext(fastq_path_list[2 * i + 1]);
            out = g_dump ? capture_file(1, a.c_str(), b.c_str(), g1, g2)
                         : process_kmer_pair(a.c_str(), b.c_str(), g_tab.repeat_check_table, g_tab.rot_table, g_tab.extract_k_mer,
                                             g_tab.extract_k_mer_128, g_tab.extract_k_mer_ans, thread_data_list, g1, g2);
        } else {
            std::string a = std::filesystem::canonical(fastq_path_list[i]).string();
            bool g1 = has_gz_ext(fastq_path_list[i]);
            if (is_short) {
                out = g_dump ? capture_file(0, a.c_str(), nullptr, g1, false)
                             : process_kmer(a.c_str(), g_tab.repeat_check_table, g_tab.rot_table, g_tab.extract_k_mer,
                                            g_tab.extract_k_mer_128, g_tab.extract_k_mer_ans, thread_data_list, g1);
            } else {
                out = g_dump ? capture_file(2, a.c_str(), nullptr, g1, false)
                             : process_kmer_long(a.c_str(), g_tab.repeat_check_table, g_tab.rot_table, g_tab.extract_k_mer,
                                                 g_tab.extract_k_mer_128, g_tab.extract_k_mer_ans, thread_data_list, g1);
            }
        }
        for (auto& [k, v] : *out.high) {
            if (total_result_high->contains(k)) (*total_result_high)[k] = add_data((*total_result_high)[k], v);
            else (*total_result_high)[k] = v;
        }
        for (auto& [k, v] : *out.low) {
            if (total_result_low->contains(k)) (*total_result_low)[k] = add_data((*total_result_low)[k], v);
            else (*total_result_low)[k] = v;
        }
        delete out.high;
        delete out.low;
    }
    final_process_output(total_result_high, total_result_low);
    if (g_dump) fclose(g_dump);
    return 0;
}
#endif
