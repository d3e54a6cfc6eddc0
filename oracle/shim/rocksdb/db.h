// Test-infrastructure stand-in for the external RocksDB dependency (not vendored in the
// reference tree): an in-memory, mutex-guarded string map with the handful of calls the
// reference's kvdb wrapper makes.  Used only to build oracle/_ref/sortmerna_ref.
#pragma once
#include <string>
#include <unordered_map>
#include <mutex>
#include <cassert>
namespace rocksdb {
enum CompressionType { kNoCompression, kZlibCompression, kXpressCompression };
struct Options {
  bool create_if_missing = false;
  CompressionType compression = kNoCompression;
  void IncreaseParallelism() {}
};
struct WriteOptions {};
struct ReadOptions {};
struct Status { bool ok() const { return true; } std::string ToString() const { return "OK"; } };
class DB {
  std::unordered_map<std::string, std::string> kv_;
  std::mutex mx_;
public:
  static Status Open(const Options&, const std::string&, DB** out) { *out = new DB(); return Status(); }
  Status Put(const WriteOptions&, const std::string& k, const std::string& v) {
    std::lock_guard<std::mutex> g(mx_); kv_[k] = v; return Status();
  }
  Status Get(const ReadOptions&, const std::string& k, std::string* v) {
    std::lock_guard<std::mutex> g(mx_);
    auto it = kv_.find(k); if (it != kv_.end()) *v = it->second; return Status();
  }
};
}
