// Host-side proof obligations for sortmerna_b200/csrc/smr_levbits.h (run by tests/test_lev_equivalence.py):
//   1. the reference's table-driven LEV(1) automaton (re-encoded in oracle/smr_oracle.cpp) accepts / is
//      alive / reaches state 9 exactly according to edit distance <= 1 -- so replacing it is sound;
//   2. classify_bits / viable_bits compute exactly those edit-distance predicates.
// Prints "cases N table_vs_ed M1 bits_vs_ed M2 viable_table M3 viable_bits M4"; all M must be 0.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include "../sortmerna_b200/csrc/smr_levbits.h"

static const char* const kLev0[16] = {"3eeeeeeeeeeeee", "3eeeeeeeeeeeee", "7eee4444eeeeee", "7eee4444eeeeee", "0e22ee22eeeeee", "0e22ee22eeeeee",
                                      "0e224466eeeeee", "0e224466eeeeee", "31e1e1e1eeeeee", "31e1e1e1eeeeee", "71e14545eeeeee", "71e14545eeeeee",
                                      "0123e123eeeeee", "0123e123eeeeee", "01234567eeeeee", "01234567eeeeee"};
static const char* const kLev1[8] = {"3eeeeeeeeeeeee", "deeeaaaaeeeeee", "8e22ee22eeeeee", "8e22aacceeeeee", "31e1e1e1eeeeee", "d1e1ababeeeeee",
                                     "8123e123eeeeee", "8123abcdeeeeee"};
static const char* const kLev2[4] = {"ceeeeeeeceeeee", "9eaaeeaa9eeeaa", "c1e1e1e1cee1e1", "91ace1ac9ee1ac"};
static const char* const kLev3[2] = {"aeeeeeeeeaeeee", "aaeaeaeaeaeeae"};
static uint32_t hexv(char c) { return c <= '9' ? c - '0' : c - 'a' + 10; }
static uint32_t lev_step(uint32_t t, uint32_t bv, uint32_t s) {
  switch (t) { case 0: return hexv(kLev0[bv & 15][s]); case 1: return bv < 8 ? hexv(kLev1[bv][s]) : 0; case 2: return bv < 4 ? hexv(kLev2[bv][s]) : 0;
               default: return bv < 2 ? hexv(kLev3[bv][s]) : 0; }
}
static const int pw = 9;
static uint32_t bvrow(const uint8_t* p, int d, int c) {  // bitvector.cpp:56-132
  uint32_t v = 0;
  for (int b = 0; b < 4; b++) { int k = d + 2 - b; if (d == 0 && b == 3) continue; if (k >= 0 && k < pw && p[k] == c) v |= 1u << b; }
  return v;
}
static uint32_t lev_next(const uint8_t* p, int depth, int c, uint32_t lev) {  // traverse_bursttrie.cpp:131-139
  if (depth < pw - 2) return lev_step(0, bvrow(p, depth, c), lev);
  return lev_step(3 - pw + depth, bvrow(p, pw - 3, c) & ((2u << (pw - depth)) - 1), lev);
}
static int ed(const uint8_t* a, int n, const uint8_t* b, int m) {
  int D[16][16];
  for (int i = 0; i <= n; i++) D[i][0] = i;
  for (int j = 0; j <= m; j++) D[0][j] = j;
  for (int i = 1; i <= n; i++) for (int j = 1; j <= m; j++) D[i][j] = std::min({D[i - 1][j] + 1, D[i][j - 1] + 1, D[i - 1][j - 1] + (a[i - 1] != b[j - 1])});
  return D[n][m];
}
int main(int argc, char** argv) {
  const long N = argc > 1 ? atol(argv[1]) : 1000000;
  std::mt19937_64 rng(20260924);
  long m1 = 0, m2 = 0, m3 = 0, m4 = 0;
  for (long it = 0; it < N; it++) {
    uint8_t p[9], T[12], tmp[16];
    for (int i = 0; i < 9; i++) p[i] = rng() & 3;
    int ne = rng() % 4, len = 9;
    memcpy(tmp, p, 9);
    for (int e = 0; e < ne && ne < 3; e++) {
      int typ = rng() % 3, pos = rng() % len;
      if (typ == 0) tmp[pos] = rng() & 3;
      else if (typ == 1 && len > 5) { memmove(tmp + pos, tmp + pos + 1, len - pos - 1); len--; }
      else if (len < 12) { memmove(tmp + pos + 1, tmp + pos, len - pos); tmp[pos] = rng() & 3; len++; }
    }
    if (ne == 3) { for (int i = 0; i < 10; i++) tmp[i] = rng() & 3; len = 10; }
    for (int i = 0; i < 10; i++) T[i] = i < len ? tmp[i] : (rng() & 3);
    uint32_t Pb = 0, Tb = 0;
    for (int i = 0; i < 9; i++) Pb |= (uint32_t)p[i] << (2 * i);
    for (int i = 0; i < 10; i++) Tb |= (uint32_t)T[i] << (2 * i);
    // table automaton over the 10 text characters
    uint32_t s = 0; int d1 = 0, z = 0; int alive[10];
    for (int d = 0; d < 10; d++) alive[d] = 0;
    for (int d = 0; d < 10; d++) {
      s = lev_next(p, d, T[d], s);
      if (s == 14) break;
      alive[d] = 1;
      if (d >= pw - 2) { if (!d1 && s >= 8) d1 = d - (pw - 3); if (d == pw - 1 && s == 9) z = 4; }
    }
    const int e7 = ed(T, 8, p, 9) <= 1, e8 = ed(T, 9, p, 9) <= 1, e9 = ed(T, 10, p, 9) <= 1;
    const uint32_t want = (uint32_t)(e7 ? 1 : (e8 ? 2 : (e9 ? 3 : 0))) | (memcmp(T, p, 9) == 0 ? 4u : 0u);
    if ((uint32_t)(d1 | z) != want) m1++;
    if (smr::classify_bits(Pb, Tb, pw) != want) m2++;
    if (smr::within_one_edit(Pb, Tb, smr::lev_masks(pw)) != ((want & 3u) != 0)) m2++;   // the seed kernel's streaming test
    for (int k = 1; k <= 8; k++) {
      int v = 0;
      for (int j = 0; j <= 9; j++) if (ed(T, k, p, j) <= 1) v = 1;
      if (alive[k - 1] != v) { m3++; break; }
      if (smr::viable_bits(Pb, Tb, k) != (bool)v || smr::viable_bits(Pb, Tb & ((1u << (2 * k)) - 1), k) != (bool)v) { m4++; break; }
    }
  }
  printf("cases %ld table_vs_ed %ld bits_vs_ed %ld viable_table %ld viable_bits %ld\n", N, m1, m2, m3, m4);
  return (m1 | m2 | m3 | m4) ? 1 : 0;
}
