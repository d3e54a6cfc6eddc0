// stand-in for moodycamel::ConcurrentQueue (external, not vendored): mutex + std::queue
#pragma once
#include <queue>
#include <mutex>
#include <cstddef>
namespace moodycamel {
template <typename T>
class ConcurrentQueue {
  std::queue<T> q_; std::mutex m_;
public:
  ConcurrentQueue(size_t = 0) {}
  bool enqueue(const T& v) { std::lock_guard<std::mutex> g(m_); q_.push(v); return true; }
  bool try_enqueue(const T& v) { return enqueue(v); }
  bool try_dequeue(T& v) {
    std::lock_guard<std::mutex> g(m_); if (q_.empty()) return false;
    v = std::move(q_.front()); q_.pop(); return true;
  }
  size_t size_approx() { std::lock_guard<std::mutex> g(m_); return q_.size(); }
};
}
