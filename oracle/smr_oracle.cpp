/*
 * oracle/smr_oracle.cpp -- CPU restatement of SortMeRNA v5.0.0's per-read alignment hot path
 * (align -> align2 -> traverse -> traversetrie_align -> compute_lis_alignment -> ssw_align).
 *
 * TEST INFRASTRUCTURE ONLY.  The product (sortmerna_b200/, libsmr_b200.so) never includes, links
 * or calls this file; it is used by tests/, __graft_entry__.smoke() and the CPU legs of bench.py
 * as the checker.  Written from the behaviour described in SURVEY.md sections 3/8/Appendix A;
 * every function cites the reference file:line (relative to /root/reference) it restates.
 *
 * Parity of this restatement is PINNED (tests/test_oracle_pin.py, tests/golden/):
 *   - ora_ssw_align against the reference's own ssw.c compiled as oracle/_ref/libssw_ref.so,
 *   - ora_align end to end against SAM/BLAST/log output of oracle/_ref/sortmerna_ref (the
 *     unmodified reference built by oracle/Makefile.ref) on the bundled data sets.
 */
#include "smr_oracle.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace {

/* ------------------------------------------------------------------------------------------
 * Index (index.cpp:143-357; include/indexdb.hpp:67-104)
 * ------------------------------------------------------------------------------------------ */
struct Elem {          /* one of the 4 elements (A,C,G,T) of a trie node; NodeElement, indexdb.hpp:67-84 */
  uint8_t flag;        /* 0 empty, 1 child trie node, 2 bucket */
  uint32_t child;      /* flag 1: index of the child node in Index::nodes */
  uint32_t boff, bsize;/* flag 2: byte offset/size of the bucket in Index::buckets */
};
struct Node { Elem e[4]; };
struct SeqPos { uint32_t pos, seq; }; /* indexdb.hpp:87-91 */

} // namespace

struct ora_index {
  uint32_t lnwin = 0, partialwin = 0;
  std::vector<uint32_t> count;          /* kmer::count per 9-mer */
  std::vector<int64_t> rootF, rootR;    /* node index of the mini-trie root, -1 = NULL */
  std::vector<Node> nodes;
  std::vector<uint8_t> buckets;         /* 8-byte entries {u32 tail, u32 id} */
  std::vector<uint64_t> pos_off;        /* id -> [pos_off[id], pos_off[id+1]) */
  std::vector<SeqPos> pos;
  uint64_t n_buckets = 0, max_bucket = 0, max_pos = 0;
};

namespace {

struct Cursor {
  const uint8_t* p; size_t n, o = 0; bool bad = false;
  uint32_t u32() { if (o + 4 > n) { bad = true; return 0; } uint32_t v; memcpy(&v, p + o, 4); o += 4; return v; }
  uint8_t u8() { if (o + 1 > n) { bad = true; return 0; } return p[o++]; }
  const uint8_t* bytes(size_t k) { if (o + k > n) { bad = true; return nullptr; } const uint8_t* r = p + o; o += k; return r; }
};

bool read_file(const std::string& path, std::vector<uint8_t>& out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
  out.resize((size_t)sz);
  size_t got = sz ? fread(out.data(), 1, (size_t)sz, f) : 0;
  fclose(f);
  return got == (size_t)sz;
}

/* BFS stream of one mini burst trie (index.cpp:193-300): 4 flag bytes of the root, then for every
 * node popped in FIFO order and each of its 4 elements: flag 1 -> the child's 4 flag bytes follow,
 * flag 2 -> u32 bucket size + bucket bytes follow. */
int64_t parse_trie(Cursor& c, ora_index& ix) {
  struct Pending { uint32_t node; uint8_t flags[4]; };
  std::deque<Pending> q;
  Pending root; root.node = (uint32_t)ix.nodes.size();
  ix.nodes.push_back(Node{});
  for (int k = 0; k < 4; ++k) root.flags[k] = c.u8();
  q.push_back(root);
  while (!q.empty() && !c.bad) {
    Pending cur = q.front(); q.pop_front();
    for (int k = 0; k < 4; ++k) {
      Elem el{}; el.flag = cur.flags[k];
      if (el.flag == 1) {
        Pending ch; ch.node = (uint32_t)ix.nodes.size();
        ix.nodes.push_back(Node{});
        for (int t = 0; t < 4; ++t) ch.flags[t] = c.u8();
        el.child = ch.node;
        q.push_back(ch);
      } else if (el.flag == 2) {
        uint32_t sz = c.u32();
        const uint8_t* b = c.bytes(sz);
        if (!b) return -2;
        el.boff = (uint32_t)ix.buckets.size(); el.bsize = sz;
        ix.buckets.insert(ix.buckets.end(), b, b + sz);
        ix.n_buckets++; ix.max_bucket = std::max<uint64_t>(ix.max_bucket, sz);
      } else if (el.flag != 0) {
        return -2; /* index.cpp:282-286: unknown flag is fatal */
      }
      ix.nodes[cur.node].e[k] = el;
    }
  }
  return c.bad ? -2 : (int64_t)root.node;
}

/* The universal Levenshtein automaton for d=1 used by the reference (15 states, 14 = dead),
 * traverse_bursttrie.cpp:68-98, re-encoded one hex digit per state: kLev[t][bitvector][state]. */
const char* const kLev0[16] = {"3eeeeeeeeeeeee", "3eeeeeeeeeeeee", "7eee4444eeeeee", "7eee4444eeeeee",
                               "0e22ee22eeeeee", "0e22ee22eeeeee", "0e224466eeeeee", "0e224466eeeeee",
                               "31e1e1e1eeeeee", "31e1e1e1eeeeee", "71e14545eeeeee", "71e14545eeeeee",
                               "0123e123eeeeee", "0123e123eeeeee", "01234567eeeeee", "01234567eeeeee"};
const char* const kLev1[8] = {"3eeeeeeeeeeeee", "deeeaaaaeeeeee", "8e22ee22eeeeee", "8e22aacceeeeee",
                              "31e1e1e1eeeeee", "d1e1ababeeeeee", "8123e123eeeeee", "8123abcdeeeeee"};
const char* const kLev2[4] = {"ceeeeeeeceeeee", "9eaaeeaa9eeeaa", "c1e1e1e1cee1e1", "91ace1ac9ee1ac"};
const char* const kLev3[2] = {"aeeeeeeeeaeeee", "aaeaeaeaeaeeae"};

inline uint32_t hexv(char ch) { return ch <= '9' ? (uint32_t)(ch - '0') : (uint32_t)(ch - 'a' + 10); }
inline uint32_t lev_step(uint32_t tbl, uint32_t bv, uint32_t state) {
  /* the reference's table[1..3] rows beyond 8/4/2 are zero-initialised; masked indices never reach them */
  switch (tbl) {
    case 0: return hexv(kLev0[bv & 15][state]);
    case 1: return bv < 8 ? hexv(kLev1[bv][state]) : 0;
    case 2: return bv < 4 ? hexv(kLev2[bv][state]) : 0;
    default: return bv < 2 ? hexv(kLev3[bv][state]) : 0;
  }
}

/* bitvector.cpp:56-132.  bv[d*4+c], d=0..partialwin-3, c in ACGT.  Row 0 has 3 bits (chars 0,1,2 of
 * the half window -> bits 2,1,0); row d>=1 = (row d-1 << 1) & 15 with bit 0 set for letter half[d+2].
 * dir=+1: half window read ascending from p (init_win_f); dir=-1: descending (init_win_r). */
void build_bitvectors(const uint8_t* p, int dir, uint32_t partialwin, uint8_t* bv) {
  uint32_t rows = partialwin - 2; /* bitvec_size = (partialwin-2)<<2, paralleltraversal.cpp:107 */
  memset(bv, 0, rows * 4);
  const uint8_t* q = p;
  for (int bit = 2; bit >= 0; --bit) { bv[*q] |= (uint8_t)(1u << bit); q += dir; }
  for (uint32_t d = 1; d < rows; ++d) {
    for (int c = 0; c < 4; ++c) bv[d * 4 + c] = (uint8_t)((bv[(d - 1) * 4 + c] << 1) & 15);
    bv[d * 4 + *q] |= 1; q += dir;
  }
}

struct Instr { uint64_t windows = 0, nodes = 0, entries = 0, buckets = 0, pos_entries = 0, sw_calls = 0, sw_cells = 0; };

struct IdWin { uint32_t id, win; }; /* traverse_bursttrie.hpp:57-90 */

struct SeedCtx {
  const ora_index* ix;
  const uint8_t* bv;      /* win_k1_ptr */
  const uint8_t* bv_last; /* win_k1_full = bv + ((partialwin-3)<<2), paralleltraversal.cpp:110 */
  bool full_search;
  uint32_t win;
  std::vector<IdWin>* hits;
  bool accept_zero;
  Instr* ins;
};

inline uint32_t lev_next(const SeedCtx& c, uint32_t depth, uint32_t letter, uint32_t lev) {
  uint32_t pw = c.ix->partialwin;
  if (depth < pw - 2) return lev_step(0, c.bv[(depth << 2) + letter], lev);   /* traverse_bursttrie.cpp:131-135 */
  return lev_step(3 - pw + depth, c.bv_last[letter] & ((2u << (pw - depth)) - 1), lev); /* :136-139 */
}

/* traverse_bursttrie.cpp:100-298: DFS of one mini-trie in lock step with the automaton. */
void walk_trie(SeedCtx& c, uint32_t node_idx, uint32_t lev_in, uint32_t depth) {
  const ora_index& ix = *c.ix;
  const uint32_t pw = ix.partialwin;
  if (c.ins) c.ins->nodes++;
  for (uint32_t letter = 0; letter < 4; ++letter) {
    const Elem& el = ix.nodes[node_idx].e[letter];
    if (el.flag == 0) continue;
    uint32_t lev = lev_next(c, depth, letter, lev_in);
    if (lev == 14) continue;
    if (el.flag == 1) {
      walk_trie(c, el.child, lev, depth + 1);
      if (c.accept_zero) return;               /* :167 */
      continue;
    }
    /* bucket (:176-292) */
    if (c.ins) c.ins->buckets++;
    const uint32_t nchars = pw - depth;        /* :184 */
    const uint8_t* b = ix.buckets.data() + el.boff;
    const uint8_t* bend = b + el.bsize;
    for (; b != bend; b += 8) {
      if (c.ins) c.ins->entries++;
      uint32_t tail, id; memcpy(&tail, b, 4); memcpy(&id, b + 4, 4);
      uint32_t depth_b = depth, st = lev;
      bool local_accept = false;
      for (uint32_t j = 0; j < nchars; ++j) {
        uint32_t nt = tail & 3;
        ++depth_b;
        st = lev_next(c, depth_b, nt, st);
        if (st == 14) break;
        if (depth_b >= pw - 2) {               /* :229 */
          if (st >= 8) local_accept = true;    /* 1-error match */
          if (depth_b == pw - 1 && st == 9) {  /* 0-error match, :237-246 */
            c.accept_zero = true;
            if (c.full_search) c.accept_zero = false;
          }
        }
        if (local_accept) {
          if (c.accept_zero) {                 /* :256-262 */
            c.hits->clear();
            c.hits->push_back(IdWin{id, c.win});
            return;
          }
          bool dup = false;                    /* :265-277 */
          for (const IdWin& h : *c.hits) if (h.id == id) { dup = true; break; }
          if (dup) break;
          c.hits->push_back(IdWin{id, c.win});
        }
        tail >>= 2;
      }
    }
  }
}

inline uint32_t hash_kmer(const uint8_t* s, uint32_t len) { /* read.cpp:601-611 */
  uint32_t h = 0;
  for (uint32_t i = 0; i < len; ++i) h = (h << 2) | s[i];
  return h;
}

/* paralleltraversal.cpp:129-249: both sub-searches of one window; returns accept_zero_kmer */
bool seed_window(const ora_index& ix, const uint8_t* seq03, uint32_t win_pos, bool full_search, int minoccur,
                 std::vector<IdWin>& id_hits, Instr* ins) {
  uint8_t bv[64];
  const uint32_t pw = ix.partialwin;
  const uint32_t last = (pw - 3) << 2;
  SeedCtx c{&ix, bv, bv + last, full_search, win_pos, &id_hits, false, ins};
  if (ins) ins->windows++;
  build_bitvectors(seq03 + win_pos + pw, +1, pw, bv);           /* :141-142 */
  uint32_t keyf = hash_kmer(seq03 + win_pos, pw);               /* :145 */
  if ((int64_t)ix.count[keyf] > (int64_t)minoccur && ix.rootF[keyf] >= 0) /* :161 */
    walk_trie(c, (uint32_t)ix.rootF[keyf], 0, 0);
  if (!c.accept_zero) {                                         /* :188 */
    build_bitvectors(seq03 + win_pos + pw - 1, -1, pw, bv);     /* :194-195 */
    uint32_t keyr = hash_kmer(seq03 + win_pos + pw, pw);        /* :198 */
    if ((int64_t)ix.count[keyr] > (int64_t)minoccur && ix.rootR[keyr] >= 0) /* :215 */
      walk_trie(c, (uint32_t)ix.rootR[keyr], 0, 0);
  }
  return c.accept_zero;
}

/* ------------------------------------------------------------------------------------------
 * Smith-Waterman (ssw.c).  Scores are those of an affine-gap local alignment (Gotoh); the SSE2
 * byte kernel is only a saturating prefilter for the word kernel (ssw.c:862-869), so outputs are
 * defined by true scores + the tie-breaks below.
 * ------------------------------------------------------------------------------------------ */
struct SwEnd { int32_t score, ref, read; };

/* sw_sse2_byte / sw_sse2_word restated column by column (ssw.c:150-373, 399-575).
 * ref_dir 0: columns 0..refLen-1; 1: refLen-1..0.  terminate < 0: never terminates early.
 * Column loop order, "first column whose max strictly exceeds the running max" (ssw.c:310-318,
 * 516-526), "smallest read index holding that max in the saved column" (:328-336, :537-545) and the
 * early exit when a column max equals `terminate` (:324, :528) are the observable semantics. */
SwEnd sw_scan(const int8_t* ref, int ref_dir, int32_t refLen, const int8_t* read, int32_t readLen,
              const int8_t* mat, int32_t go, int32_t ge, int32_t terminate) {
  std::vector<int32_t> H(readLen + 1, 0), E(readLen + 1, 0), Hbest(readLen, 0);
  int32_t best = 0, end_ref = 0, end_read = readLen - 1;
  bool have_best = false;
  int32_t begin = 0, end = refLen, step = 1;
  if (ref_dir == 1) { begin = refLen - 1; end = -1; step = -1; }
  for (int32_t j = begin; j != end; j += step) {
    const int8_t* mrow = mat + 5 * ref[j];
    int32_t diag = 0, F = 0, colmax = 0; /* H(-1, .) = 0 */
    for (int32_t i = 0; i < readLen; ++i) {
      /* H[i] currently holds H(i, prev column); E[i] holds E(i, this column) */
      int32_t h = diag + mrow[read[i]];
      if (h < E[i]) h = E[i];
      if (h < F) h = F;
      if (h < 0) h = 0;
      diag = H[i];
      H[i] = h;
      if (h > colmax) colmax = h;
      int32_t open = h - go;
      int32_t e = E[i] - ge; E[i] = e > open ? e : open; if (E[i] < 0) E[i] = 0; /* saturating, ssw.c:257-260 */
      F = F - ge; if (F < open) F = open; if (F < 0) F = 0;
    }
    if (colmax > best) {
      best = colmax; end_ref = j; have_best = true;
      for (int32_t i = 0; i < readLen; ++i) Hbest[i] = H[i];
    }
    if (terminate >= 0 && colmax == terminate) break;
  }
  if (have_best) {
    for (int32_t i = 0; i < readLen; ++i) if (Hbest[i] == best) { if (i < end_read) end_read = i; break; }
  } else {
    end_read = 0; /* all-zero saved column: index 0 matches max==0 (ssw.c:331-336) */
    if (readLen - 1 < end_read) end_read = readLen - 1;
  }
  return SwEnd{best, end_ref, end_read};
}

/* banded_sw (ssw.c:577-773): same arrays, same band coordinates, same direction codes. */
inline int32_t band_u(int32_t w, int32_t i, int32_t j) { int32_t x = i - w; if (x < 0) x = 0; return j - x + 1; }   /* set_u, ssw.c:70 */
inline int32_t band_d(int32_t w, int32_t i, int32_t j, int32_t p) { int32_t x = i - w; if (x < 0) x = 0; return (j - x) * 3 + p; } /* set_d, :73 */

bool banded_traceback(const int8_t* ref, const int8_t* read, int32_t refLen, int32_t readLen, int32_t score,
                      int32_t go, int32_t ge, int32_t band_width, const int8_t* mat, std::vector<uint32_t>& cig) {
  std::vector<int32_t> h_b, e_b, h_c;
  std::vector<int8_t> dir;
  int32_t maxv = 0, width = 0, width_d = 0;
  do {
    width = band_width * 2 + 3; width_d = band_width * 2 + 1;
    if ((int64_t)width_d * readLen * 3 > (int64_t)1 << 31) return false; /* ssw.c:608-612 */
    /* realloc keeps old contents in the reference (:601-606,:626): grow without clearing */
    if ((int32_t)h_b.size() < width + 1) { h_b.resize(width + 1, 0); e_b.resize(width + 1, 0); h_c.resize(width + 1, 0); }
    if (dir.size() < (size_t)width_d * readLen * 3 + 8) dir.resize((size_t)width_d * readLen * 3 + 8, 0);
    for (int32_t j = 1; j < width - 1; ++j) h_b[j] = 0;
    for (int32_t i = 0; i < readLen; ++i) {
      int32_t beg = std::max(0, i - band_width), end = std::min(refLen - 1, i + band_width);
      int32_t edge = end + 1 < width - 1 ? end + 1 : width - 1;
      int32_t f = 0, u = 0;
      h_b[0] = e_b[0] = h_b[edge] = e_b[edge] = h_c[0] = 0;
      int8_t* dl = dir.data() + (size_t)width_d * i * 3;
      for (int32_t j = beg; j <= end; ++j) {
        u = band_u(band_width, i, j);
        int32_t up = band_u(band_width, i - 1, j), lf = band_u(band_width, i, j - 1), dg = band_u(band_width, i - 1, j - 1);
        int32_t de = band_d(band_width, i, j, 0), df = de + 1, dh = de + 2;
        int32_t t1 = i == 0 ? -go : h_b[up] - go;
        int32_t t2 = i == 0 ? -ge : e_b[up] - ge;
        e_b[u] = t1 > t2 ? t1 : t2;
        dl[de] = t1 > t2 ? 3 : 2;
        t1 = h_c[lf] - go; t2 = f - ge;
        f = t1 > t2 ? t1 : t2;
        dl[df] = t1 > t2 ? 5 : 4;
        int32_t e1 = e_b[u] > 0 ? e_b[u] : 0, f1 = f > 0 ? f : 0;
        t1 = e1 > f1 ? e1 : f1;
        t2 = h_b[dg] + mat[ref[j] * 5 + read[i]];
        h_c[u] = t1 > t2 ? t1 : t2;
        if (h_c[u] > maxv) maxv = h_c[u];
        if (t1 <= t2) dl[dh] = 1; else dl[dh] = e1 > f1 ? dl[de] : dl[df];
      }
      for (int32_t j = 1; j <= u; ++j) h_b[j] = h_c[j];
    }
    band_width *= 2;
  } while (maxv < score);
  band_width /= 2;

  /* trace back (ssw.c:674-747) */
  int32_t i = readLen - 1, j = refLen - 1, run = 0, cur_op = 0, op = 0, which = 2;
  const int8_t* dl = dir.data() + (size_t)width_d * (readLen - 1) * 3;
  std::vector<uint32_t> c;
  while (i > 0) {
    int32_t t = band_d(band_width, i, j, which);
    switch (dl[t]) {
      case 1: --i; --j; which = 2; dl -= width_d * 3; op = 0; break;
      case 2: --i; which = 0; dl -= width_d * 3; op = 1; break;
      case 3: --i; which = 2; dl -= width_d * 3; op = 1; break;
      case 4: --j; which = 1; op = 2; break;
      case 5: --j; which = 2; op = 2; break;
      default: return false; /* "Trace back error" -> exit(1) in the reference */
    }
    if (op == cur_op) ++run;
    else { c.push_back((uint32_t)run << 4 | (uint32_t)cur_op); cur_op = op; run = 1; }
  }
  if (op == 0) c.push_back((uint32_t)(run + 1) << 4);
  else { c.push_back((uint32_t)run << 4 | (uint32_t)op); c.push_back(16); }
  cig.assign(c.rbegin(), c.rend());
  return true;
}

struct SwResult { int32_t score1 = 0, ref_begin1 = -1, ref_end1 = 0, read_begin1 = -1, read_end1 = 0; std::vector<uint32_t> cigar; bool ok = true; };

/* ssw_align with flag=2, filterd=0, maskLen=0 (ssw.c:834-941, call site alignment.cpp:371-381) */
SwResult ssw_align_restated(const int8_t* read, int32_t readLen, const int8_t* ref, int32_t refLen,
                            const int8_t* mat, int32_t go, int32_t ge, uint16_t filters) {
  SwResult r;
  SwEnd fwd = sw_scan(ref, 0, refLen, read, readLen, mat, go, ge, -1);
  r.score1 = fwd.score; r.ref_end1 = fwd.ref; r.read_end1 = fwd.read;
  if (fwd.score == 0) r.ref_end1 = -1; /* the byte kernel (used when there is no overflow) starts end_ref at -1, ssw.c:179 */
  if ((uint16_t)r.score1 < filters) return r; /* ssw.c:897 */
  std::vector<int8_t> rev(read, read + r.read_end1 + 1);
  std::reverse(rev.begin(), rev.end());      /* seq_reverse, ssw.c:775-786 */
  SwEnd bwd = sw_scan(ref, 1, r.ref_end1 + 1, rev.data(), r.read_end1 + 1, mat, go, ge, r.score1);
  r.ref_begin1 = bwd.ref; r.read_begin1 = r.read_end1 - bwd.read; /* :914-915 */
  int32_t rl = r.ref_end1 - r.ref_begin1 + 1, ql = r.read_end1 - r.read_begin1 + 1;
  int32_t band = std::abs(rl - ql) + 1;      /* :924 */
  r.ok = banded_traceback(ref + r.ref_begin1, read + r.read_begin1, rl, ql, r.score1, go, ge, band, mat, r.cigar);
  return r;
}

/* ------------------------------------------------------------------------------------------
 * Per-read state (include/read.hpp; read.cpp)
 * ------------------------------------------------------------------------------------------ */
struct Aln { /* s_align2, ssw.hpp:44-56 */
  std::vector<uint32_t> cigar; uint32_t ref_num = 0; int32_t ref_begin1 = 0, ref_end1 = 0, read_begin1 = 0, read_end1 = 0;
  uint32_t readlen = 0; uint16_t score1 = 0, part = 0, index_num = 0; bool strand = false;
};
struct Persist { /* what Read::toBinString stores (read.cpp:429-462) and load_db restores (:467-539) */
  bool present = false;
  uint32_t lastIndex = 0, lastPart = 0, hit_seeds = 0, min_index = 0, max_index = 0;
  uint16_t max_SW_count = 0; bool is_done = false, is_hit = false;
  std::vector<Aln> alignv;
};
struct ReadSt {
  uint32_t len = 0;
  std::vector<uint8_t> iseq;          /* isequence */
  std::vector<uint32_t> ambiguous;    /* ambiguous_nt */
  bool is03 = true, is04 = false, reversed = false;
  uint32_t lastIndex = 0, lastPart = 0, hit_seeds = 0, min_index = 0, max_index = 0;
  int32_t best = 0;
  uint16_t max_SW_count = 0;
  bool is_done = false, is_hit = false, is_new_hit = false;
  std::vector<Aln> alignv;
  std::vector<IdWin> id_win_hits;
  int8_t mat[25];
};

void flip34(ReadSt& r) { /* read.cpp:379-401 */
  if (r.ambiguous.empty()) return;
  uint8_t val = r.is03 ? 4 : 0;
  for (uint32_t p : r.ambiguous) r.iseq[r.reversed ? r.len - p - 1 : p] = val;
  r.is03 = !r.is03; r.is04 = !r.is04;
}
void rev_int_str(ReadSt& r) { /* read.cpp:350-357 */
  static const uint8_t comp[5] = {3, 2, 1, 0, 4};
  std::reverse(r.iseq.begin(), r.iseq.end());
  for (auto& c : r.iseq) c = comp[c];
  r.reversed = !r.reversed;
}

void find_lis(const std::deque<std::pair<uint32_t, uint32_t>>& a, std::vector<uint32_t>& b) { /* alignment.cpp:58-98 */
  std::vector<uint32_t> p(a.size());
  if (a.empty()) return;
  b.push_back(0);
  for (size_t i = 1; i < a.size(); ++i) {
    if (a[b.back()].second < a[i].second) { p[i] = b.back(); b.push_back((uint32_t)i); continue; }
    size_t u = 0, v = b.size() - 1;
    while (u < v) { size_t c = (u + v) / 2; if (a[b[c]].second < a[i].second) u = c + 1; else v = c; }
    if (a[i].second < a[b[u]].second) { if (u > 0) p[i] = b[u - 1]; b[u] = (uint32_t)i; }
  }
  size_t u = b.size(); uint32_t v = b.back();
  while (u--) { b[u] = v; v = p[v]; }
}

uint32_t find_min_index(const std::vector<Aln>& v) { /* alignment.cpp:533-546 */
  uint32_t ms = v[0].score1, mi = 0;
  for (uint32_t i = 0; i < v.size(); ++i) if (v[i].score1 < ms) { ms = v[i].score1; mi = i; }
  return mi;
}
uint32_t find_max_index(const std::vector<Aln>& v) { /* alignment.cpp:548-561 */
  uint32_t ms = v[0].score1, mi = 0;
  for (uint32_t i = 0; i < v.size(); ++i) if (v[i].score1 > ms) { ms = v[i].score1; mi = i; }
  return mi;
}

struct PassCtx {
  const ora_index* ix; uint16_t index_num, part;
  const uint8_t* refseq; const uint64_t* refoff; uint32_t nref;
  uint32_t minimal_score; const uint32_t* skip; bool is_last_idx;
  const ora_params* prm;
  std::atomic<uint64_t>* num_aligned; std::atomic<uint64_t>* matched_per_db;
};

/* alignment.cpp:100-509 */
void compute_lis_alignment(ReadSt& read, const PassCtx& pc, bool& search, uint32_t max_SW_score, Instr& ins) {
  const ora_params& o = *pc.prm;
  const ora_index& ix = *pc.ix;
  bool is_aligned = false;
  std::map<uint32_t, uint32_t> cnt;
  for (const IdWin& h : read.id_win_hits) {          /* :118-130 */
    for (uint64_t k = ix.pos_off[h.id]; k < ix.pos_off[h.id + 1]; ++k) { cnt[ix.pos[k].seq]++; ins.pos_entries++; }
  }
  std::vector<std::pair<uint32_t, uint32_t>> cand;   /* (ref, count) */
  for (auto& kv : cnt) if (kv.second >= (uint32_t)o.num_seeds) cand.push_back(kv);
  std::sort(cand.begin(), cand.end(), [](const std::pair<uint32_t, uint32_t>& a, const std::pair<uint32_t, uint32_t>& b) {
    if (a.second == b.second) return a.first < b.first;
    return a.second > b.second; });                  /* :143-148 */

  bool is_search_candidates = true;
  for (uint32_t k = 0; k < cand.size() && is_search_candidates; ++k) {
    uint32_t max_ref = cand[k].first, max_occur = cand[k].second;
    if (max_occur < (uint32_t)o.num_seeds) break;    /* :158 */
    if (is_aligned && o.min_lis > 0 && k > 0 && max_occur < cand[k - 1].second) { /* :165-169 */
      --read.best;
      if (read.best < 1) break;
    }
    std::vector<std::pair<uint32_t, uint32_t>> hits_on_ref; /* (ref pos, read pos) :181-194 */
    for (const IdWin& h : read.id_win_hits)
      for (uint64_t q = ix.pos_off[h.id]; q < ix.pos_off[h.id + 1]; ++q)
        if (ix.pos[q].seq == max_ref) hits_on_ref.emplace_back(ix.pos[q].pos, h.win);
    std::sort(hits_on_ref.begin(), hits_on_ref.end());  /* (first asc, second asc) :197-201 */

    size_t it = 0; const size_t nh = hits_on_ref.size();
    std::deque<std::pair<uint32_t, uint32_t>> match_set;
    uint32_t begin_ref = hits_on_ref[0].first, begin_read = hits_on_ref[0].second;
    const uint64_t rlen = read.len;
    const uint64_t lnwin = ix.lnwin;

    while (it != nh && is_search_candidates) {        /* :217 */
      uint64_t end_ref_max = (uint64_t)begin_ref + rlen - begin_read - lnwin + 1; /* :231 (size_t arithmetic) */
      bool push = false;
      while (it != nh && (uint64_t)hits_on_ref[it].first <= end_ref_max) { match_set.push_back(hits_on_ref[it]); push = true; ++it; }
      bool skip_to_pop = false;
      if (!push && is_aligned) skip_to_pop = true;    /* heuristic 1, :244 */
      else is_aligned = false;
      if (!skip_to_pop && match_set.size() >= (size_t)o.num_seeds) {
        std::vector<uint32_t> lis;
        find_lis(match_set, lis);
        if (lis.size() >= (size_t)o.min_lis) {        /* :261 */
          uint32_t lcs_ref_start = match_set[lis[0]].first, lcs_que_start = match_set[lis[0]].second;
          uint64_t head = 0, tail = 0, ars = 0, aqs = 0, alen = 0;
          uint64_t reflen = pc.refoff[max_ref + 1] - pc.refoff[max_ref];
          uint32_t edges = o.edges_is_percent ? (uint32_t)((o.edges / 100.0) * rlen) : (uint32_t)o.edges; /* :278-282 */
          if (lcs_ref_start < lcs_que_start) {        /* :288-330 */
            ars = 0; aqs = lcs_que_start - lcs_ref_start; head = 0;
            if (reflen < rlen) {
              tail = 0;
              if (aqs > (rlen - reflen)) alen = reflen - (aqs - (rlen - reflen));
              else alen = reflen;
            } else {
              tail = reflen - ars - rlen;
              if (tail > (uint64_t)(uint32_t)(edges - 1)) tail = edges;
              alen = rlen + head + tail - aqs;
            }
          } else {                                    /* :331-357 */
            ars = lcs_ref_start - lcs_que_start; aqs = 0;
            if (ars > (uint64_t)(uint32_t)(edges - 1)) head = edges;
            if (ars + rlen > reflen) { tail = 0; alen = reflen - ars - head; }
            else {
              tail = reflen - ars - rlen;
              if (tail > (uint64_t)(uint32_t)(edges - 1)) tail = edges;
              alen = rlen + head + tail;
            }
          }
          if (read.is03) flip34(read);                /* :360-361 */
          const int8_t* q = (const int8_t*)read.iseq.data() + aqs;
          int32_t qlen = (int32_t)(alen - head - tail);
          const int8_t* t = (const int8_t*)pc.refseq + pc.refoff[max_ref] + ars - head;
          ins.sw_calls++; ins.sw_cells += (uint64_t)alen * (uint64_t)qlen;
          SwResult res = ssw_align_restated(q, qlen, t, (int32_t)alen, read.mat, o.gap_open, o.gap_ext, (uint16_t)pc.minimal_score);
          is_aligned = res.ok && (uint32_t)res.score1 > pc.minimal_score; /* :388 */
          if (is_aligned) {
            if ((uint32_t)res.score1 == max_SW_score) ++read.max_SW_count; /* :391 */
            Aln a;
            a.cigar = res.cigar;
            a.ref_begin1 = res.ref_begin1 + (int32_t)(ars - head); a.ref_end1 = res.ref_end1 + (int32_t)(ars - head);
            a.read_begin1 = res.read_begin1 + (int32_t)aqs; a.read_end1 = res.read_end1 + (int32_t)aqs;
            a.readlen = (uint32_t)rlen; a.ref_num = max_ref; a.index_num = pc.index_num; a.part = pc.part;
            a.strand = !read.reversed; a.score1 = (uint16_t)res.score1;
            if (!read.is_hit) {                       /* :411-416 */
              read.is_hit = true;
              pc.num_aligned->fetch_add(1);
              pc.matched_per_db[pc.index_num].fetch_add(1);
            }
            const uint32_t N = (uint32_t)o.num_alignments;
            if (N == 0 || !o.is_best || (o.is_best && read.alignv.size() < N)) { /* :420-424 */
              read.alignv.push_back(a); read.is_new_hit = true;
            } else if (o.is_best && read.alignv.size() == N && read.alignv[read.min_index].score1 < a.score1) { /* :425-459 */
              if (N > 1 && read.max_index == 0 && read.min_index == 0) {
                read.min_index = find_min_index(read.alignv); read.max_index = find_max_index(read.alignv);
              }
              uint32_t mn = read.min_index, mx = read.max_index;
              read.alignv[mn] = a; read.is_new_hit = true;
              if (a.score1 > read.alignv[mx].score1 && read.alignv.size() > 1) {
                read.max_index = mn; read.min_index = find_min_index(read.alignv);
              }
              /* :454-457: -- then ++ of reads_matched_per_db on the SAME (already overwritten) index_num: net no-op */
            }
            if (N > 0) {                              /* :462-469 */
              if (o.is_best) { if (N == read.max_SW_count) is_search_candidates = false; }
              else if (N == read.alignv.size()) is_search_candidates = false;
            }
            search = false;                           /* :472 */
          }
        }
      }
      /* pop: (:486-506) */
      if (!match_set.empty()) match_set.pop_front();
      if (match_set.empty()) {
        if (it != nh) { begin_ref = hits_on_ref[it].first; begin_read = hits_on_ref[it].second; }
        else break;
      } else { begin_ref = match_set.front().first; begin_read = match_set.front().second; }
    }
  }
}

/* paralleltraversal.cpp:81-297 */
void traverse(ReadSt& read, const PassCtx& pc, bool is_last_strand, Instr& ins) {
  const ora_params& o = *pc.prm;
  const ora_index& ix = *pc.ix;
  read.lastIndex = pc.index_num; read.lastPart = pc.part;
  uint32_t win_shift = pc.skip[0];
  std::vector<bool> searched(read.len, false);
  size_t pass_n = 0;
  uint32_t max_SW_score = read.len * (uint32_t)o.match;
  for (bool search = true; search;) {
    uint32_t numwin = (read.len - ix.lnwin + win_shift) / win_shift;
    uint32_t win_pos = 0;
    for (uint32_t w = 0; w < numwin; ++w) {
      if (read.is04) flip34(read);                    /* :126 */
      if (!searched[win_pos]) {
        searched[win_pos] = true;
        std::vector<IdWin> id_hits;
        seed_window(ix, read.iseq.data(), win_pos, o.is_full_search != 0, o.minoccur, id_hits, &ins);
        if (!id_hits.empty()) { for (auto& h : id_hits) read.id_win_hits.push_back(h); ++read.hit_seeds; } /* :242-249 */
      }
      if (w == numwin - 1) {                          /* :253-279 */
        if (read.hit_seeds >= (uint32_t)o.num_seeds) compute_lis_alignment(read, pc, search, max_SW_score, ins);
        if (search) {
          if (pass_n == 2) search = false;
          else {
            while (pass_n < 2 && pc.skip[pass_n] == pc.skip[pass_n + 1]) ++pass_n; /* :269-272 (bounded to the 3 entries) */
            if (++pass_n > 2) search = false; else win_shift = pc.skip[pass_n];
          }
        }
        break;
      }
      win_pos += win_shift;
    }
  }
  const uint32_t N = (uint32_t)o.num_alignments;      /* :286-297 */
  if (N > 0) {
    if ((o.is_best && N == read.max_SW_count) || (!o.is_best && read.alignv.size() == N)) read.is_done = true;
  } else {
    if (pc.is_last_idx && is_last_strand && !read.alignv.empty()) read.is_done = true;
  }
}

/* the per-read body of align2 (processor.cpp:104-162) */
void align_one(const uint8_t* seq04, uint32_t len, Persist& db, const PassCtx& pc, std::atomic<uint64_t>* num_short, Instr& ins) {
  const ora_params& o = *pc.prm;
  ReadSt read; read.len = len;
  /* Read::init (read.cpp:264-271) */
  read.best = o.min_lis > 0 ? o.min_lis : 0;
  read.iseq.resize(len);
  for (uint32_t i = 0; i < len; ++i) {              /* seqToIntStr, read.cpp:334-347 */
    uint8_t c = seq04[i];
    if (c >= 4) { read.ambiguous.push_back(i); c = 0; }
    read.iseq[i] = c;
  }
  for (int l = 0, k = 0; l < 4; ++l) {              /* initScoringMatrix, read.cpp:274-288 */
    for (int m = 0; m < 4; ++m) read.mat[k++] = (int8_t)(l == m ? o.match : o.mismatch);
    read.mat[k++] = (int8_t)o.score_N;
  }
  for (int m = 0; m < 5; ++m) read.mat[20 + m] = (int8_t)o.score_N;
  if (len < pc.ix->lnwin) { num_short->fetch_add(1); return; }   /* processor.cpp:109-114 */
  if (len == 0) return;
  if (db.present) {                                 /* load_db, read.cpp:467-539 */
    read.lastIndex = db.lastIndex; read.lastPart = db.lastPart; read.is_done = db.is_done; read.is_hit = db.is_hit;
    read.max_SW_count = db.max_SW_count; read.hit_seeds = db.hit_seeds;
    read.min_index = db.min_index; read.max_index = db.max_index; read.alignv = db.alignv;
  }
  if (read.is_done) return;                         /* processor.cpp:120-126 */
  bool single = (o.is_forward != 0) ^ (o.is_reverse != 0);
  int num_strands = single ? 1 : 2;
  for (int count = 0; count < num_strands && !read.is_done; ++count) {
    if ((single && o.is_reverse) || count == 1) { if (!read.reversed) rev_int_str(read); }
    traverse(read, pc, single || count == 1, ins);
    read.id_win_hits.clear();
  }
  if (read.is_new_hit && !read.alignv.empty()) {    /* kvdb.put(read.id, toBinString()), processor.cpp:150-155 */
    db.present = true; db.lastIndex = read.lastIndex; db.lastPart = read.lastPart; db.is_done = read.is_done;
    db.is_hit = read.is_hit; db.max_SW_count = read.max_SW_count; db.hit_seeds = read.hit_seeds;
    db.min_index = read.min_index; db.max_index = read.max_index; db.alignv = read.alignv;
  }
}

} // namespace

static thread_local uint32_t g_all_slots = 16, g_need_slots = 0;   /* per calling thread (the stand-in of the C ABI runs one context per thread) */

extern "C" {

void ora_set_aln_slots(uint32_t slots) { g_all_slots = slots ? slots : 1; g_need_slots = 0; }
uint32_t ora_aln_slots_needed(void) { return g_need_slots; }

ora_index* ora_index_load(const char* prefix, uint32_t part, uint32_t lnwin, char* err, size_t errlen) {
  auto fail = [&](const std::string& m) -> ora_index* { if (err && errlen) snprintf(err, errlen, "%s", m.c_str()); return nullptr; };
  std::string pre(prefix), sfx = "_" + std::to_string(part) + ".dat";
  std::vector<uint8_t> kmer, trie, posf;
  if (!read_file(pre + ".kmer" + sfx, kmer)) return fail("cannot read " + pre + ".kmer" + sfx);
  if (!read_file(pre + ".bursttrie" + sfx, trie)) return fail("cannot read " + pre + ".bursttrie" + sfx);
  if (!read_file(pre + ".pos" + sfx, posf)) return fail("cannot read " + pre + ".pos" + sfx);
  ora_index* ix = new ora_index();
  ix->lnwin = lnwin; ix->partialwin = lnwin / 2;
  const uint32_t limit = 1u << lnwin;               /* index.cpp:155 */
  if (kmer.size() < (size_t)limit * 4) { delete ix; return fail("kmer file too short"); }
  ix->count.resize(limit);
  memcpy(ix->count.data(), kmer.data(), (size_t)limit * 4);
  ix->rootF.assign(limit, -1); ix->rootR.assign(limit, -1);
  Cursor c{trie.data(), trie.size()};
  for (uint32_t i = 0; i < limit; ++i) {
    uint32_t sz[2] = {c.u32(), c.u32()};
    if (c.bad) { delete ix; return fail("bursttrie file truncated"); }
    if (ix->count[i] == 0) continue;                /* index.cpp:187: tries are only present when count != 0 */
    for (int j = 0; j < 2; ++j) {
      if (sz[j] == 0) continue;
      int64_t root = parse_trie(c, *ix);
      if (root < 0) { delete ix; return fail("bursttrie stream corrupt at 9-mer " + std::to_string(i)); }
      (j == 0 ? ix->rootF : ix->rootR)[i] = root;
    }
  }
  if (c.o != trie.size()) { delete ix; return fail("bursttrie file has trailing bytes"); }
  Cursor p{posf.data(), posf.size()};
  uint32_t n = p.u32();
  ix->pos_off.assign((size_t)n + 1, 0);
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t sz = p.u32();
    const uint8_t* b = p.bytes((size_t)sz * 8);
    if (p.bad) { delete ix; return fail("pos file truncated"); }
    size_t o = ix->pos.size();
    ix->pos.resize(o + sz);
    if (sz) memcpy(ix->pos.data() + o, b, (size_t)sz * 8);
    ix->pos_off[i + 1] = ix->pos.size();
    ix->max_pos = std::max<uint64_t>(ix->max_pos, sz);
  }
  if (p.o != posf.size()) { delete ix; return fail("pos file has trailing bytes"); }
  return ix;
}

void ora_index_free(ora_index* ix) { delete ix; }
uint32_t ora_index_num_ids(const ora_index* ix) { return (uint32_t)(ix->pos_off.size() - 1); }

void ora_index_stats(const ora_index* ix, uint64_t out[8]) {
  uint64_t ne = 0; for (uint32_t v : ix->count) ne += v != 0;
  out[0] = ne; out[1] = ix->nodes.size(); out[2] = ix->n_buckets; out[3] = ix->buckets.size() / 8;
  out[4] = ix->pos_off.size() - 1; out[5] = ix->pos.size(); out[6] = ix->max_bucket; out[7] = ix->max_pos;
}

int ora_seed_window(const ora_index* ix, const uint8_t* seq03, uint32_t win_pos, int is_full_search, int minoccur,
                    uint32_t* ids, int max_ids, int* accept_zero) {
  std::vector<IdWin> hits;
  bool az = seed_window(*ix, seq03, win_pos, is_full_search != 0, minoccur, hits, nullptr);
  if (accept_zero) *accept_zero = az;
  int n = 0;
  for (auto& h : hits) { if (n < max_ids) ids[n] = h.id; ++n; }
  return n;
}

int ora_ssw_align(const int8_t* read, int32_t readLen, const int8_t* ref, int32_t refLen, const int8_t* mat,
                  int32_t gap_open, int32_t gap_ext, uint16_t filters, int32_t out[6], uint32_t* cigar, int32_t max_cigar) {
  SwResult r = ssw_align_restated(read, readLen, ref, refLen, mat, gap_open, gap_ext, filters);
  out[0] = r.score1; out[1] = r.ref_begin1; out[2] = r.ref_end1; out[3] = r.read_begin1; out[4] = r.read_end1;
  out[5] = (int32_t)r.cigar.size();
  for (int32_t i = 0; i < (int32_t)r.cigar.size() && i < max_cigar; ++i) cigar[i] = r.cigar[i];
  return r.ok ? 0 : 1;
}

int ora_align(const ora_index* const* idx, const uint16_t* index_num, const uint16_t* part, uint32_t nidx,
              uint32_t n_index_files,
              const uint8_t* const* refseq, const uint64_t* const* refoff, const uint32_t* nref,
              const uint32_t* minimal_score, const uint32_t* skiplengths, const ora_params* prm,
              const uint8_t* reads, const uint64_t* readoff, uint32_t nreads,
              ora_read_result* res, ora_aln* alns, uint32_t* cigar_pool, uint64_t cigar_cap, uint64_t* cigar_used,
              uint64_t* reads_matched_per_db, ora_counters* counters, int nthreads) {
  std::vector<Persist> db(nreads);
  std::atomic<uint64_t> num_aligned{0}, num_short{0};
  std::vector<std::atomic<uint64_t>> matched(n_index_files);
  for (auto& m : matched) m = 0;
  if (nthreads < 1) nthreads = 1;
  std::vector<Instr> ins(nthreads);
  for (uint32_t k = 0; k < nidx; ++k) {
    num_short = 0;                                   /* processor.cpp:228 */
    PassCtx pc{idx[k], index_num[k], part[k], refseq[k], refoff[k], nref[k], minimal_score[k], skiplengths + 3 * k,
               k == nidx - 1, prm, &num_aligned, matched.data()};
    std::atomic<uint32_t> next{0};
    auto worker = [&](int tid) {
      for (;;) {
        uint32_t b = next.fetch_add(64);
        if (b >= nreads) break;
        uint32_t e = std::min(nreads, b + 64);
        for (uint32_t r = b; r < e; ++r)
          align_one(reads + readoff[r], (uint32_t)(readoff[r + 1] - readoff[r]), db[r], pc, &num_short, ins[tid]);
      }
    };
    if (nthreads == 1) worker(0);
    else {
      std::vector<std::thread> th;
      for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
      for (auto& t : th) t.join();
    }
  }
  const uint32_t slots = prm->num_alignments > 0 ? (uint32_t)prm->num_alignments : g_all_slots;   /* 0 = all alignments: caller-set stride */
  uint64_t used = 0; int rc = 0;
  for (uint32_t r = 0; r < nreads; ++r) {
    const Persist& p = db[r];
    ora_read_result& o = res[r];
    memset(&o, 0, sizeof(o));
    if (!p.present) continue;
    o.lastIndex = p.lastIndex; o.lastPart = p.lastPart; o.hit_seeds = p.hit_seeds; o.min_index = p.min_index; o.max_index = p.max_index;
    o.max_SW_count = p.max_SW_count; o.is_done = p.is_done; o.is_hit = p.is_hit;
    o.n_align = (uint32_t)p.alignv.size();
    if (o.n_align > slots) { rc = 2; if (o.n_align > g_need_slots) g_need_slots = o.n_align; o.n_align = slots; } /* all-alignments mode: stride too small, ora_aln_slots_needed() */
    for (uint32_t a = 0; a < o.n_align; ++a) {
      const Aln& s = p.alignv[a]; ora_aln& d = alns[(uint64_t)r * slots + a];
      memset(&d, 0, sizeof(d));
      d.cigar_off = (uint32_t)used; d.cigar_len = (uint32_t)s.cigar.size();
      if (used + s.cigar.size() > cigar_cap) { rc = 3; d.cigar_len = 0; }
      else { memcpy(cigar_pool + used, s.cigar.data(), s.cigar.size() * 4); used += s.cigar.size(); }
      d.ref_num = s.ref_num; d.ref_begin1 = s.ref_begin1; d.ref_end1 = s.ref_end1; d.read_begin1 = s.read_begin1; d.read_end1 = s.read_end1;
      d.readlen = s.readlen; d.score1 = s.score1; d.part = s.part; d.index_num = s.index_num; d.strand = s.strand;
    }
  }
  if (cigar_used) *cigar_used = used;
  if (reads_matched_per_db) for (uint32_t i = 0; i < n_index_files; ++i) reads_matched_per_db[i] = matched[i];
  if (counters) {
    memset(counters, 0, sizeof(*counters));
    counters->num_aligned = num_aligned; counters->num_short_last = num_short;
    for (auto& i : ins) {
      counters->sw_calls += i.sw_calls; counters->sw_cells += i.sw_cells; counters->windows += i.windows; counters->trie_nodes += i.nodes;
      counters->bucket_entries += i.entries; counters->buckets += i.buckets; counters->pos_entries += i.pos_entries;
    }
  }
  return rc;
}

} // extern "C"
