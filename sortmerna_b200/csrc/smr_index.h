// Host-side flattener of the on-disk indexdb format into contiguous arrays for HBM.
// Stands in for Index::load (src/sortmerna/index.cpp:143-357), which builds pointer-linked
// NodeElement tries (include/indexdb.hpp:67-84); here every mini burst trie becomes 32-byte nodes
// plus 8-byte bucket entries in two flat arrays, addressed by 32-bit indices.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace smr {

constexpr uint32_t kNone = 0xFFFFFFFFu;

// One trie node = 4 elements (A,C,G,T) x {w0, w1} = 32 bytes = one DRAM sector.
//   w0 bits[1:0] flag (0 empty, 1 child node, 2 bucket); w0 >> 2 = number of bucket entries (flag 2)
//   w1 = child node index (flag 1) or first entry index in `entries` (flag 2)
struct FlatNode { uint32_t w[8]; };
struct Entry { uint32_t tail, id; };  // bucket entry: 2-bit packed tail (LSB first) + k-mer id (indexdb.hpp:57)
struct SeqPos { uint32_t pos, seq; }; // indexdb.hpp:87-91

struct FlatIndex {
  uint32_t lnwin = 0, partialwin = 0;
  std::vector<uint32_t> lookup;   // 2 words per 9-mer: root node index of trie_F, trie_R (kNone = absent or count <= minoccur(0))
  std::vector<uint32_t> kmer_count; // kmer::count per 9-mer (kept for the minoccur test, paralleltraversal.cpp:161)
  std::vector<FlatNode> nodes;
  std::vector<Entry> entries;
  // DFS-ordered flat form of every mini burst trie (what the kernels read): for 9-mer k, direction d
  // (0 = trie_F, 1 = trie_R) the list flist[flookup[4k+2d] .. +flookup[4k+2d+1]) holds one item per bucket
  // entry in the order the reference's DFS visits them (elements A,C,G,T; child before next sibling; bucket
  // order; traverse_bursttrie.cpp:117-295): {text, id} with text = trie path letters + bucket tail =
  // partialwin+1 characters, 2 bits each, first character in the lowest bits.
  std::vector<uint32_t> flookup;  // 4 words per 9-mer: offF, cntF, offR, cntR
  std::vector<Entry> flist;       // {text, id}
  uint32_t max_list = 0;
  std::vector<uint32_t> pos_off;  // id -> [pos_off[id], pos_off[id+1])
  std::vector<SeqPos> pos;        // every list sorted by (seq, pos)
  uint64_t n_buckets = 0;
  uint32_t max_bucket_entries = 0, max_positions = 0;
};

// Returns empty string on success, else the error text.
std::string flatten_index(const void* kmer_file, size_t kmer_bytes, const void* trie_file, size_t trie_bytes,
                          const void* pos_file, size_t pos_bytes, uint32_t lnwin, FlatIndex& out);

}  // namespace smr
