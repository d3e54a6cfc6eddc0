// Host-side run of sortmerna_b200/csrc/smr_inflate.h: the same FIND / COUNT / chain / WRITE / WINDOW / RESOLVE steps the CUDA
// kernels perform (smr_inflate.cuh), serially on the CPU.  tests/test_inflate_host.py compares the output with zlib's.
//   inflate_check in.gz out.bin chunk_bytes   -> prints "ok bytes N spans S candidates C" or "error <status>"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../sortmerna_b200/csrc/smr_inflate.h"
using namespace smr;

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  std::vector<uint8_t> raw;
  { uint8_t buf[65536]; size_t k; while ((k = fread(buf, 1, sizeof buf, f)) > 0) raw.insert(raw.end(), buf, buf + k); fclose(f); }
  const uint64_t nbytes = raw.size(), CH = strtoull(argv[3], nullptr, 10);
  std::vector<uint32_t> w((nbytes + 64 + 3) / 4 + 1, 0);
  memcpy(w.data(), raw.data(), nbytes);
  const uint64_t nbits = nbytes * 8;
  // FIND
  std::vector<uint64_t> cand;
  for (uint64_t c = CH; c < nbytes; c += CH) {
    const uint64_t end = std::min(nbits, (c + CH) * 8);
    for (uint64_t p = c * 8; p < end; ++p) if (inf_probe_block(w.data(), nbits, p)) { cand.push_back(p); break; }
  }
  // COUNT
  const uint32_t ns = (uint32_t)cand.size() + 1;
  std::vector<SpanResult> res(ns);
  HuffTabs T;
  for (uint32_t i = 0; i < ns; ++i)
    inflate_span<false>(w.data(), nbytes, i ? cand[i - 1] : 0, i == 0, cand.data(), (uint32_t)cand.size(), i, T, nullptr, 0, nullptr, res[i]);
  std::vector<uint32_t> real(ns); std::vector<uint64_t> off(ns);
  uint32_t nreal = 0, why = 0;
  const uint64_t total = inf_chain(cand.data(), (uint32_t)cand.size(), res.data(), real.data(), off.data(), nreal, &why);
  if (total == kInfNone) { printf("error %u\n", why); return 1; }
  // WRITE
  std::vector<uint16_t> sym(total + 1);
  std::vector<MemberEnd> ends;
  for (uint32_t k = 0; k < nreal; ++k) {
    const uint32_t i = real[k];
    SpanResult r;
    std::vector<MemberEnd> mine(res[i].members + 1);
    inflate_span<true>(w.data(), nbytes, i ? cand[i - 1] : 0, i == 0, cand.data(), (uint32_t)cand.size(), i, T, sym.data() + off[k], res[i].out_n, mine.data(), r);
    if (r.status != res[i].status || r.out_n != res[i].out_n || r.end_bit != res[i].end_bit || r.members != res[i].members) { printf("error write pass differs\n"); return 1; }
    for (uint32_t m = 0; m < r.members; ++m) { mine[m].out_end += off[k]; ends.push_back(mine[m]); }
  }
  // WINDOW (front to back) + RESOLVE
  std::vector<uint8_t> win((size_t)(nreal + 1) * kInfWindow, 0), out(total);
  for (uint32_t k = 0; k < nreal; ++k) {
    const uint8_t* prev = win.data() + (size_t)k * kInfWindow;
    uint8_t* cur = win.data() + (size_t)(k + 1) * kInfWindow;
    const uint64_t n = res[real[k]].out_n;
    for (uint32_t j = 0; j < kInfWindow; ++j) cur[j] = inf_window_byte(sym.data() + off[k], n, prev, j);
  }
  for (uint32_t k = 0; k < nreal; ++k) {
    const uint8_t* prev = win.data() + (size_t)k * kInfWindow;
    const uint64_t n = res[real[k]].out_n;
    for (uint64_t j = 0; j < n; ++j) out[off[k] + j] = inf_resolve(sym[off[k] + j], prev);
  }
  // CRC-32 and ISIZE of every member, in pieces as on the device
  {
    std::vector<uint64_t> poff; std::vector<uint32_t> plen, first, crcs, tab(256);
    for (uint32_t i = 0; i < 256; ++i) tab[i] = crc_table_entry(i);
    inf_crc_plan(ends, 32768, poff, plen, first);
    crcs.resize(poff.size());
    for (size_t k = 0; k < poff.size(); ++k) crcs[k] = crc_piece(out.data() + poff[k], plen[k], tab.data());
    const uint32_t bad = inf_crc_verify(ends, plen, first, crcs.data());
    if (bad) { printf("error %u\n", bad); return 1; }
  }
  f = fopen(argv[2], "wb");
  fwrite(out.data(), 1, out.size(), f); fclose(f);
  printf("ok bytes %llu spans %u candidates %zu\n", (unsigned long long)total, nreal, cand.size());
  return 0;
}
