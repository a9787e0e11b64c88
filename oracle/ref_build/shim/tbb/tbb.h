// Shim: tbb::task_group::run/wait on std::thread.
#pragma once
#include <thread>
#include <vector>
#include "concurrent_queue.h"
namespace tbb {
class task_group {
    std::vector<std::thread> th_;
public:
    template <class F> void run(F f) { th_.emplace_back(std::move(f)); }
    void wait() { for (auto& t : th_) t.join(); th_.clear(); }
    ~task_group() { wait(); }
};
}  // namespace tbb
