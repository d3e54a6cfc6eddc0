// Flattener of the indexdb on-disk format (see smr_index.h).
// Format facts (src/sortmerna/index.cpp:143-357, writer src/sortmerna/indexdb.cpp:714-870):
//   .kmer_P.dat      : 2^lnwin x u32 count                                  (index.cpp:155-161)
//   .bursttrie_P.dat : for each 9-mer: u32 bytes_F, u32 bytes_R; if count != 0, for each non-zero
//                      size a BFS stream: 4 root flag bytes, then for every node in FIFO order and
//                      each of its 4 elements: flag 1 -> the child's 4 flag bytes follow, flag 2 ->
//                      u32 bucket bytes + the bucket (8-byte entries)          (index.cpp:176-315)
//   .pos_P.dat       : u32 N, then N x { u32 size, size x {u32 pos, u32 seq} } (index.cpp:328-352)
#include "smr_index.h"

#include <algorithm>
#include <cstring>
#include <deque>

namespace smr {
namespace {

struct Reader {
  const uint8_t* p; size_t n; size_t o = 0; bool bad = false;
  uint32_t u32() { if (o + 4 > n) { bad = true; return 0; } uint32_t v; memcpy(&v, p + o, 4); o += 4; return v; }
  uint8_t u8() { if (o >= n) { bad = true; return 0; } return p[o++]; }
  const uint8_t* take(size_t k) { if (o + k > n) { bad = true; return nullptr; } const uint8_t* r = p + o; o += k; return r; }
};

// Parses one mini burst trie; returns root node index or kNone on error.
uint32_t parse_mini_trie(Reader& rd, FlatIndex& fx) {
  struct Open { uint32_t node; uint8_t flag[4]; };
  std::deque<Open> fifo;
  Open root; root.node = (uint32_t)fx.nodes.size();
  fx.nodes.push_back(FlatNode{});
  for (auto& f : root.flag) f = rd.u8();
  fifo.push_back(root);
  while (!fifo.empty()) {
    if (rd.bad) return kNone;
    Open cur = fifo.front(); fifo.pop_front();
    for (int k = 0; k < 4; ++k) {
      uint32_t w0 = 0, w1 = 0;
      switch (cur.flag[k]) {
        case 0: break;
        case 1: {
          Open ch; ch.node = (uint32_t)fx.nodes.size();
          fx.nodes.push_back(FlatNode{});
          for (auto& f : ch.flag) f = rd.u8();
          fifo.push_back(ch);
          w0 = 1; w1 = ch.node;
          break;
        }
        case 2: {
          uint32_t bytes = rd.u32();
          const uint8_t* b = rd.take(bytes);
          if (!b || (bytes & 7)) return kNone;
          uint32_t cnt = bytes / 8;
          w0 = 2u | (cnt << 2); w1 = (uint32_t)fx.entries.size();
          size_t o = fx.entries.size();
          fx.entries.resize(o + cnt);
          if (cnt) memcpy(&fx.entries[o], b, bytes);
          fx.n_buckets++;
          fx.max_bucket_entries = std::max(fx.max_bucket_entries, cnt);
          break;
        }
        default: return kNone;  // index.cpp:282-286: fatal in the reference
      }
      fx.nodes[cur.node].w[2 * k] = w0;
      fx.nodes[cur.node].w[2 * k + 1] = w1;
    }
  }
  return rd.bad ? kNone : root.node;
}

// appends the entries of the mini trie rooted at `node` to fx.flist in the reference's DFS order
bool dfs_flatten(const FlatIndex& fx, std::vector<Entry>& out, uint32_t node, uint32_t depth, uint32_t path) {
  if (depth > fx.partialwin) return false;   // a mini trie is at most partialwin + 1 characters deep (malformed file otherwise)
  for (uint32_t c = 0; c < 4; ++c) {
    const uint32_t w0 = fx.nodes[node].w[2 * c], w1 = fx.nodes[node].w[2 * c + 1];
    const uint32_t flag = w0 & 3u, tp = path | (c << (2 * depth));
    if (flag == 1) { if (!dfs_flatten(fx, out, w1, depth + 1, tp)) return false; }
    else if (flag == 2) {
      const uint32_t cnt = w0 >> 2;
      for (uint32_t k = 0; k < cnt; ++k) out.push_back(Entry{tp | (fx.entries[w1 + k].tail << (2 * (depth + 1))), fx.entries[w1 + k].id});
    }
  }
  return true;
}

}  // namespace

std::string flatten_index(const void* kmer_file, size_t kmer_bytes, const void* trie_file, size_t trie_bytes,
                          const void* pos_file, size_t pos_bytes, uint32_t lnwin, FlatIndex& fx) {
  if (lnwin < 8 || lnwin > 26 || (lnwin & 1)) return "unsupported lnwin";
  fx = FlatIndex{};
  fx.lnwin = lnwin; fx.partialwin = lnwin / 2;
  const uint32_t limit = 1u << lnwin;  // index.cpp:155 (= 4^partialwin)
  if (kmer_bytes < (size_t)limit * 4) return "kmer file too short";
  fx.kmer_count.resize(limit);
  memcpy(fx.kmer_count.data(), kmer_file, (size_t)limit * 4);
  fx.lookup.assign((size_t)limit * 2, kNone);
  Reader rd{(const uint8_t*)trie_file, trie_bytes};
  for (uint32_t i = 0; i < limit; ++i) {
    uint32_t sz[2] = {rd.u32(), rd.u32()};
    if (rd.bad) return "bursttrie file truncated";
    if (fx.kmer_count[i] == 0) continue;  // index.cpp:187
    for (int j = 0; j < 2; ++j) {
      if (sz[j] == 0) continue;
      uint32_t root = parse_mini_trie(rd, fx);
      if (root == kNone) return "bursttrie stream corrupt at 9-mer " + std::to_string(i);
      fx.lookup[(size_t)i * 2 + j] = root;
    }
  }
  if (rd.o != trie_bytes) return "bursttrie file has trailing bytes";
  // DFS-ordered flat lists (text = path + tail needs (partialwin+1)*2 <= 32 bits)
  if (fx.partialwin + 1 > 16) return "lnwin too large for 32-bit packed texts";
  fx.flookup.assign((size_t)limit * 4, 0);
  fx.flist.reserve(fx.entries.size());
  for (uint32_t i = 0; i < limit; ++i) {
    for (int j = 0; j < 2; ++j) {
      const uint32_t root = fx.lookup[(size_t)i * 2 + j];
      const size_t o = fx.flist.size();
      if (root != kNone && !dfs_flatten(fx, fx.flist, root, 0, 0)) return "bursttrie stream corrupt: trie deeper than the seed half at 9-mer " + std::to_string(i);
      fx.flookup[(size_t)i * 4 + 2 * j] = (uint32_t)o;
      fx.flookup[(size_t)i * 4 + 2 * j + 1] = (uint32_t)(fx.flist.size() - o);
      fx.max_list = std::max(fx.max_list, (uint32_t)(fx.flist.size() - o));
    }
  }

  Reader pr{(const uint8_t*)pos_file, pos_bytes};
  uint32_t n = pr.u32();
  if (pr.bad) return "pos file truncated";
  if (4 + (uint64_t)n * 4 > pos_bytes) return "pos file: id count larger than the file";
  fx.pos_off.assign((size_t)n + 1, 0);
  fx.pos.reserve((pos_bytes - 4 - (size_t)n * 4) / 8);
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t sz = pr.u32();
    const uint8_t* b = pr.take((size_t)sz * 8);
    if (pr.bad) return "pos file truncated";
    size_t o = fx.pos.size();
    if (o + sz > 0xFFFFFFFFull) return "more than 2^32 positions in one index part";
    fx.pos.resize(o + sz);
    if (sz) memcpy(&fx.pos[o], b, (size_t)sz * 8);
    // candidate gathering binary-searches each list by seq: lists are written in sequence-scan
    // order (indexdb.cpp:1723, add_kmer_to_table :318-348), i.e. already sorted; re-sort defensively
    // (order inside a list never matters to the reference: alignment.cpp:118-130,181-201 count / re-sort).
    auto lt = [](const SeqPos& a, const SeqPos& c) { return a.seq != c.seq ? a.seq < c.seq : a.pos < c.pos; };
    if (!std::is_sorted(fx.pos.begin() + o, fx.pos.end(), lt)) std::sort(fx.pos.begin() + o, fx.pos.end(), lt);
    fx.pos_off[i + 1] = (uint32_t)fx.pos.size();
    fx.max_positions = std::max(fx.max_positions, sz);
  }
  if (pr.o != pos_bytes) return "pos file has trailing bytes";
  // every id stored in a bucket must address the positions table
  for (const Entry& e : fx.entries) if (e.id >= n) return "bucket entry id out of range";
  return std::string();
}

}  // namespace smr
