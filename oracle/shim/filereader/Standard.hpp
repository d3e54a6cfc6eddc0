// stand-in for rapidgzip's StandardFileReader (external, not vendored): carries the path only
#pragma once
#include <string>
namespace rapidgzip {
struct StandardFileReader {
  std::string path;
  explicit StandardFileReader(const std::string& p) : path(p) {}
};
}
