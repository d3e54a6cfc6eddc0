// stand-in for rapidgzip::ParallelGzipReader (external, not vendored) on top of zlib gz* calls
#pragma once
#include <zlib.h>
#include <memory>
#include <cstddef>
#include <filereader/Standard.hpp>
namespace rapidgzip {
template <typename T = void>
class ParallelGzipReader {
  gzFile f_ = nullptr;
public:
  ParallelGzipReader(std::unique_ptr<StandardFileReader> fr, std::size_t) { f_ = gzopen(fr->path.c_str(), "rb"); }
  ~ParallelGzipReader() { if (f_) gzclose(f_); }
  long long read(char* buf, std::size_t n) { return gzread(f_, buf, (unsigned)n); }
  long long seek(long long off) { return gzseek(f_, off, SEEK_SET); }
};
}
