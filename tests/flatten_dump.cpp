// Dumps what flatten_index (sortmerna_b200/csrc/smr_index.cpp) makes of an on-disk index part, for tests/test_index_order_model.py:
//   flatten_dump <prefix> <part> <lnwin> <out_dir>   -> flookup.u32, flist.u32 (text,id pairs), pos_off.u32, pos.u32 (pos,seq pairs)
#include <cstdio>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>
#include "../sortmerna_b200/csrc/smr_index.h"

static std::vector<char> slurp(const std::string& p) { std::ifstream f(p, std::ios::binary); return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>()); }
static void dump(const std::string& p, const void* d, size_t n) { FILE* f = fopen(p.c_str(), "wb"); fwrite(d, 1, n, f); fclose(f); }

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  const std::string pre = argv[1], part = argv[2], out = argv[4];
  auto k = slurp(pre + ".kmer_" + part + ".dat"), t = slurp(pre + ".bursttrie_" + part + ".dat"), p = slurp(pre + ".pos_" + part + ".dat");
  smr::FlatIndex fx;
  const std::string e = smr::flatten_index(k.data(), k.size(), t.data(), t.size(), p.data(), p.size(), (uint32_t)atoi(argv[3]), fx);
  if (!e.empty()) { fprintf(stderr, "%s\n", e.c_str()); return 1; }
  dump(out + "/flookup.u32", fx.flookup.data(), fx.flookup.size() * 4);
  dump(out + "/flist.u32", fx.flist.data(), fx.flist.size() * 8);
  dump(out + "/pos_off.u32", fx.pos_off.data(), fx.pos_off.size() * 4);
  dump(out + "/pos.u32", fx.pos.data(), fx.pos.size() * 8);
  return 0;
}
