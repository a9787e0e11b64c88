// Shim: tbb::concurrent_bounded_queue with the four members the reference calls.
#pragma once
#include <condition_variable>
#include <cstddef>
#include <deque>
#include <mutex>
namespace tbb {
template <class T>
class concurrent_bounded_queue {
    std::mutex mu_;
    std::condition_variable not_empty_, not_full_;
    std::deque<T> q_;
    std::ptrdiff_t cap_ = -1;
public:
    void set_capacity(std::ptrdiff_t c) { std::lock_guard<std::mutex> l(mu_); cap_ = c; }
    void push(const T& v) {
        std::unique_lock<std::mutex> l(mu_);
        not_full_.wait(l, [&] { return cap_ < 0 || (std::ptrdiff_t)q_.size() < cap_; });
        q_.push_back(v);
        not_empty_.notify_one();
    }
    void pop(T& out) {
        std::unique_lock<std::mutex> l(mu_);
        not_empty_.wait(l, [&] { return !q_.empty(); });
        out = q_.front();
        q_.pop_front();
        not_full_.notify_one();
    }
    bool empty() { std::lock_guard<std::mutex> l(mu_); return q_.empty(); }
};
}  // namespace tbb
